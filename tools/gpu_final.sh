O=gpurun_out; mkdir -p $O
timeout 900 python bench.py > $O/bench_final2.json 2> $O/bench_final2.err; echo "bench rc=$?"; cut -c1-250 $O/bench_final2.json
VLN_PDL=0 timeout 300 python tools/trace_step.py $O/trace_final2.md > /dev/null 2>&1; head -14 $O/trace_final2.md
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/launches.csv python tools/profile_step.py > $O/profile_step.log 2>&1; echo "launchlist rc=$?"
python tools/summarize_launches.py $O/launches.csv $O/launches_final2.md > /dev/null 2>&1
timeout 300 python tools/lstm_stamps.py > $O/lstm_stamps2.log 2>&1; tail -12 $O/lstm_stamps2.log
