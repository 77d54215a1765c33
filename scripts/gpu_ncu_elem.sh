#!/bin/bash
# ncu --set full (with source) of the HBM-bound norm kernels of one eager train step: the full-resolution launches
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
E3B_BENCH_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"${KERN:-norm_bwd_fused_kernel|norm_act_kernel}" --launch-skip ${SKIP:-42} --launch-count ${CNT:-8} -f -o gpurun_out/${OUT:-r02_norm} \
  python bench.py --profile-steps 1 --no-predictor > gpurun_out/ncu_norm.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_norm.log; ls -la gpurun_out/${OUT:-r02_norm}.ncu-rep
python scripts/ncu_summary.py gpurun_out/${OUT:-r02_norm}.ncu-rep | cut -c1-260
