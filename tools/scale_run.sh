#!/bin/bash
# 1 -> 8 GPU weak-scaling lines of bench.py on one box (outputs gpurun_out/scale_N.json)
O=gpurun_out; mkdir -p $O
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/scale_1.json 2> $O/scale_1.err; echo "N=1 rc=$?"
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
    bench.py --gpus $N --steps 20 --warmup 5 > $O/scale_$N.json 2> $O/scale_$N.err; echo "N=$N rc=$?"
done
for N in 1 2 4 8; do python - <<PY
import json
d=json.load(open("gpurun_out/scale_$N.json"))
print($N, d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
done
