#!/bin/bash
# Round-2 evidence pass on one GPU box (outputs under gpurun_out/, copied to profiles/r02_* afterwards):
# bench line, warm per-kernel trace, ncu launch list of one iteration, ncu --set full of the weight-gradient kernel,
# compute-sanitizer racecheck + memcheck of the cluster / tcgen05 kernels added or touched this round.
O=gpurun_out; mkdir -p $O
timeout 400 python bench.py > $O/r2_bench_final.json 2> $O/r2_bench_final.err; echo "bench rc=$?"; cut -c1-300 $O/r2_bench_final.json
VLN_PDL=0 timeout 300 python tools/trace_step.py $O/r2_trace_final.md > $O/r2_trace_final.log 2>&1; echo "trace rc=$?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/r2_launches.csv python tools/profile_step.py > $O/r2_profile_step.log 2>&1; echo "launchlist rc=$?"
python tools/summarize_launches.py $O/r2_launches.csv $O/r2_launches.md > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tf32 -s 2 -c 1 -f -o $O/r2_ncu_wgrad \
  python tools/wgrad_bench.py > $O/r2_ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
ncu -i $O/r2_ncu_wgrad.ncu-rep --page raw --csv > $O/r2_ncu_wgrad.raw.csv 2>/dev/null
python tools/ncu_trim.py $O/r2_ncu_wgrad.raw.csv $O/r2_ncu_wgrad.csv
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_kernels_gpu.py -q -m gpu -x \
  -k "(wgrad_tcgen05 and (300 or 77)) or (dgrad and 129) or (seq_outer and 17) or pano_attn_mask_bits or ctx_step_equals" > $O/r2_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -5 $O/r2_racecheck.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_kernels_gpu.py tests/test_speaker_gpu.py -q -m gpu -x \
  -k "(wgrad_tcgen05 and (300 or 77 or 2688-512)) or dgrad or seq_outer or (speaker_matches_oracle and eval)" > $O/r2_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 $O/r2_memcheck.log
ls -la $O | grep r2_ | tail -20
