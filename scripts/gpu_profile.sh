#!/bin/bash
# ncu evidence of one round (run under gpurun, ONE GPU): plain run first, then the launch list, then `--set full` of the hot kernels.
set -x
R=${1:-r02}
python scripts/profile_step.py 3 > gpurun_out/profile_plain.log 2>&1 || exit 1
python scripts/profile_step.py 2 512 >> gpurun_out/profile_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$R.csv python scripts/profile_step.py 3 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lstm_fwdx|lstm_bwd|ctc_kernel|gemm_tn_pair|gemm_atb_pair|ctc_greedy|gemm_tn_kernel" -s 30 -c 30 -o gpurun_out/ncu_full_$R -f python scripts/profile_step.py 2 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lstm_bwd" -s 2 -c 2 -o gpurun_out/ncu_full_${R}_b512 -f python scripts/profile_step.py 2 512 > gpurun_out/ncu_full_b512.log 2>&1
ls -la gpurun_out/*.ncu-rep
