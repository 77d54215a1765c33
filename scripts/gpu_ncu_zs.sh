#!/bin/bash
# ncu --set full (with source) of the dominant conv_zs launch (up_convs.1.conv1 forward: 32+32 -> 32 @ 4x64^3) via scripts/zs_bench.py
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
ZS_SHAPES=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_zs_kernel -s 8 -c 2 -f -o gpurun_out/r02_zs_concat \
  python scripts/zs_bench.py 0 > gpurun_out/ncu_zs.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_zs.log; ls -la gpurun_out/r02_zs_concat.ncu-rep
