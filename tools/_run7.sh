O=gpurun_out; mkdir -p $O
timeout 900 python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$?"; cut -c1-300 $O/bench_final.json
VLN_PDL=0 timeout 300 python tools/trace_step.py $O/trace_final.md > $O/trace_final.log 2>&1; head -16 $O/trace_final.md
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/launches.csv python tools/profile_step.py > $O/profile_step.log 2>&1; echo "launchlist rc=$?"
python tools/summarize_launches.py $O/launches.csv $O/launches_final.md > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_bwd -c 1 -f -o $O/ncu_lstm_tc_bwd \
  python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "test_lstm_layer_matches_oracle and tcgen05 and 256-256-64-80" > $O/ncu_lstm_bwd.log 2>&1; echo "ncu lstm bwd rc=$?"
ncu -i $O/ncu_lstm_tc_bwd.ncu-rep --page raw --csv > $O/ncu_lstm_tc_bwd.raw.csv 2>/dev/null; wc -c $O/ncu_lstm_tc_bwd.raw.csv
