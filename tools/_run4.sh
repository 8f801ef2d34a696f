O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench3.json 2> $O/bench3.err; echo "bench rc=$?"; cut -c1-300 $O/bench3.json
VLN_PDL=0 timeout 300 python tools/trace_step.py $O/trace3.md > $O/trace3.log 2>&1; head -12 $O/trace3.md
