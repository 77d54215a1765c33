#!/bin/bash
# One `ncu --set full` capture of the hot kernels of the second (warm) train step -> gpurun_out/full.ncu-rep + CSV summary.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
K='regex:conv_zs_kernel|conv_tc_kernel|wgrad_tc_kernel|norm_bwd_apply|norm_bwd_reduce_kernel|norm_act_kernel'
# matching launches per step: conv_zs 9 + conv_tc 14 + wgrad 12 + apply 12 + reduce 12 + norm_act 10 = 69
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 69 -c 69 -f -o gpurun_out/full \
  python bench.py --profile-steps 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/full.ncu-rep > gpurun_out/ncu_full_summary.csv 2> gpurun_out/ncu_summary.err
head -80 gpurun_out/ncu_full_summary.csv
ls -la gpurun_out/full.ncu-rep
