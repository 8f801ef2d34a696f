#!/bin/bash
# 1 -> 8 GPU weak-scaling lines of bench.py on one box + the reference arm with per-op host times
O=gpurun_out; mkdir -p $O
python bench.py --impl reference --steps 3 --warmup 1 --per-op 2> $O/ref.err | grep '^{' > $O/scale_ref.json; echo "reference rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 2> $O/scale_1.err | grep '^{' > $O/scale_1.json; echo "N=1 rc=$?"
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
    bench.py --gpus $N --steps 20 --warmup 5 2> $O/scale_$N.err | grep '^{' > $O/scale_$N.json; echo "N=$N rc=$?"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 \
  tools/config_bench.py --agent ENVDROP --clmode SELF-PACE --batch 128 2> $O/cfg4_dp8.err | grep '^{' > $O/cfg4_dp8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701 \
  tools/config_bench.py --agent SELF-MONITOR --clmode NAIVE --batch 64 2> $O/cfg3_dp8.err | grep '^{' > $O/cfg3_dp8.json
for f in scale_ref scale_1 scale_2 scale_4 scale_8 cfg4_dp8 cfg3_dp8; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$f.json"))
    print("$f", d.get("value", d.get("episodes_per_s")), d.get("ms_per_step", d.get("ms_per_iteration")), (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("$f", "FAILED", e)
PY
done
