#!/bin/bash
# Round-2 probe visit: micro-tests that decide kernel design questions + per-role counters of the conv kernels.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
{
for sw in 0 1; do for x0 in 0 8 1 7 -1 -3 199; do timeout 30 scripts/tma_shift_test $x0 $sw; done; done
} > gpurun_out/tma_shift.txt 2>&1
cat gpurun_out/tma_shift.txt
E3B_ZS_PROF=1 ZS_SHAPES=1,2 timeout 300 python scripts/zs_bench.py 0 1 2 8 11 > gpurun_out/zs_bench.txt 2>&1
cat gpurun_out/zs_bench.txt
timeout 300 python scripts/conv_pipeline_debug.py > gpurun_out/conv_tc_roles.txt 2>&1
cat gpurun_out/conv_tc_roles.txt
timeout 300 python scripts/layer_bench.py > gpurun_out/layer_bench.txt 2>&1
cat gpurun_out/layer_bench.txt
