O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest.log
timeout 200 python tools/lstm_stamps.py > $O/lstm_stamps.log 2>&1; cat $O/lstm_stamps.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench2.json 2> $O/bench2.err; echo "bench rc=$?"; cut -c1-330 $O/bench2.json
VLN_PDL=0 timeout 300 python tools/trace_step.py $O/trace2.md > $O/trace2.log 2>&1; head -30 $O/trace2.md
