O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_agents_gpu.py -m gpu -x -q -k "trainer_end_to_end" > $O/pytest_tr.log 2>&1; echo "pytest rc=$?"; tail -30 $O/pytest_tr.log | cut -c1-300
