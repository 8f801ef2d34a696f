#!/bin/bash
# ncu captures for profiles/: launch list of one iteration + --set full of the tensor-core kernels.
O=gpurun_out; mkdir -p $O
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/launches.csv python tools/profile_step.py > $O/profile_step.log 2>&1; echo "launchlist rc=$?"
python tools/summarize_launches.py $O/launches.csv $O/launches.md > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_fwd -s 3 -c 1 -f -o $O/ncu_lstm_tc_fwd \
  python tools/lstm_stamps.py > $O/ncu_lstm_fwd.log 2>&1; echo "ncu lstm fwd rc=$?"
ncu -i $O/ncu_lstm_tc_fwd.ncu-rep --page raw --csv > $O/ncu_lstm_tc_fwd.raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_bf16x3 -s 3 -c 1 -f -o $O/ncu_linear_gates \
  python tools/gemm_bench.py --only gates > $O/ncu_linear.log 2>&1; echo "ncu linear rc=$?"
ncu -i $O/ncu_linear_gates.ncu-rep --page raw --csv > $O/ncu_linear_gates.raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_bwd -s 1 -c 1 -f -o $O/ncu_lstm_tc_bwd \
  python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "lstm_layer and tcgen05 and 256-256-64" > $O/ncu_lstm_bwd.log 2>&1; echo "ncu lstm bwd rc=$?"
ncu -i $O/ncu_lstm_tc_bwd.ncu-rep --page raw --csv > $O/ncu_lstm_tc_bwd.raw.csv 2>/dev/null
ls -la $O | grep ncu_
