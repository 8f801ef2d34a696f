#!/bin/bash
# One GPU-box pass: parity tests, bench line, config-5 microbench, warm trace, ncu launch list and
# ncu --set full captures of the two pano-attention variants.  Outputs under gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
cat $O/bench.json | cut -c1-400
timeout 400 python tools/microbench.py > $O/microbench.jsonl 2> $O/microbench.err; echo "microbench rc=$?"
VLN_PDL=0 timeout 300 python tools/trace_step.py $O/trace.md > $O/trace.log 2>&1; echo "trace rc=$?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/launches.csv python tools/profile_step.py > $O/profile_step.log 2>&1; echo "launchlist rc=$?"
python tools/summarize_launches.py $O/launches.csv $O/launches.md > /dev/null 2>&1
for B in 64 2048; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pano -c 2 -f -o $O/ncu_pano_fwd_B$B \
    python tools/microbench.py --batches $B --splits 1 --drops 0.3 --modes 0 --no-cand > $O/ncu_pano_B$B.log 2>&1; echo "ncu B=$B rc=$?"
  ncu -i $O/ncu_pano_fwd_B$B.ncu-rep --page raw --csv > $O/ncu_pano_fwd_B$B.raw.csv 2>/dev/null
done
ls -la $O
