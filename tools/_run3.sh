O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "lstm_layer" > $O/pytest_lstm.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_lstm.log
VLN_PDL=0 timeout 300 python tools/trace_step.py $O/trace3.md > $O/trace3.log 2>&1; grep "lstm_tc\|iterations" $O/trace3.md
