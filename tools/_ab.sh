timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "fused_step_tail or policy_env_act" 2>&1 | grep -v Warn | tail -8 | cut -c1-300
timeout 600 python -m pytest tests/test_agents_gpu.py -m gpu -x -q 2>&1 | tail -2
for v in 1 0; do VLN_FUSE_TAIL=$v python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('fuse_tail=$v', d['value'], d['ms_per_step'], d['e2e']['value'])"; done
