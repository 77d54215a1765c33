#!/bin/bash
# One `ncu --set full` capture of the hot kernels of a warm train step -> gpurun_out/ncu_full_summary.csv.
# The .ncu-rep of ~70 launches is > 100 MB and gpurun_out/ is capped at 64 MiB: the report stays in /tmp on the box,
# only the CSV summary (scripts/ncu_summary.py) and the raw page of the dominant kernel travel back.
# Costs ~6 GPU-minutes (ncu replays every kernel ~40 times).   usage: scripts/gpu_ncu.sh [kernel regex] [count]
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
K=${1:-'regex:conv_zs_kernel|conv_tc_kernel|wgrad_tc_kernel|norm_bwd_apply|norm_bwd_reduce_kernel|norm_act_kernel'}
# matching launches per eager step: conv_zs 9 + conv_tc 14 + wgrad 12 + apply 12 + reduce 12 + norm_act 10 = 69
CNT=${2:-69}
E3B_BENCH_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip "$CNT" -c "$CNT" -f \
  -o /tmp/full python bench.py --profile-steps 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
python scripts/ncu_summary.py /tmp/full.ncu-rep > gpurun_out/ncu_full_summary.csv 2> gpurun_out/ncu_summary.err
ncu -i /tmp/full.ncu-rep --page raw --csv -k regex:conv_zs_kernel 2>/dev/null | head -8 > gpurun_out/ncu_conv_zs_raw.csv
head -80 gpurun_out/ncu_full_summary.csv
