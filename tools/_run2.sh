O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "lstm_layer" > $O/pytest_lstm.log 2>&1; echo "pytest rc=$?"; tail -25 $O/pytest_lstm.log
timeout 200 python tools/lstm_stamps.py 2>&1 | grep -v "^mma\|^kernel:" | tee $O/lstm_tc_times.log
