#!/bin/bash
# The BASELINE.json parity configurations at full size on one GPU (outputs gpurun_out/configs.jsonl)
O=gpurun_out; mkdir -p $O; : > $O/configs.jsonl
python tools/config_bench.py --agent FOLLOWER --batch 16 2>$O/cfg1.err | grep '^{' | tee -a $O/configs.jsonl
python tools/config_bench.py --agent ENVDROP --batch 64 2>$O/cfg2.err | grep '^{' | tee -a $O/configs.jsonl
python tools/config_bench.py --agent SELF-MONITOR --clmode NAIVE --batch 64 2>$O/cfg3.err | grep '^{' | tee -a $O/configs.jsonl
python tools/config_bench.py --agent ENVDROP --clmode SELF-PACE --batch 128 2>$O/cfg4.err | grep '^{' | tee -a $O/configs.jsonl
tail -3 $O/cfg1.err $O/cfg3.err $O/cfg4.err | cut -c1-300
