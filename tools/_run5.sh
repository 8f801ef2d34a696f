O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_agents_gpu.py -m gpu -x -q > $O/pytest_agents.log 2>&1; echo "pytest rc=$?"; tail -30 $O/pytest_agents.log | cut -c1-400
