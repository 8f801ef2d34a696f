#!/bin/bash
# Round 2: 1 -> 8 GPU weak-scaling lines of bench.py on one box (no CPU legs) + configs 3 / 4 on 8 GPUs
O=gpurun_out; mkdir -p $O
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2> $O/r2_scale_1.err | grep '^{' > $O/r2_scale_1.json; echo "N=1 rc=$?"
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2> $O/r2_scale_$N.err | grep '^{' > $O/r2_scale_$N.json; echo "N=$N rc=$?"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 \
  tools/config_bench.py --agent ENVDROP --clmode SELF-PACE --batch 128 2> $O/r2_cfg4_dp8.err | grep '^{' > $O/r2_cfg4_dp8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701 \
  tools/config_bench.py --agent SELF-MONITOR --clmode NAIVE --batch 64 2> $O/r2_cfg3_dp8.err | grep '^{' > $O/r2_cfg3_dp8.json
for f in r2_scale_1 r2_scale_2 r2_scale_4 r2_scale_8 r2_cfg4_dp8 r2_cfg3_dp8; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$f.json"))
    print("$f", d.get("value", d.get("episodes_per_s")), d.get("ms_per_step", d.get("ms_per_iteration")), (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("$f", "FAILED", e)
PY
done
