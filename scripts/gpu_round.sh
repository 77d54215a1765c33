#!/bin/bash
# One GPU-box visit: GPU test-suite, bench line (both arms).  Logs -> gpurun_out/.   usage: scripts/gpu_round.sh [pytest -k expr]
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
rm -f gpurun_out/r02_parity_fullsize.json
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 300 ${1:+-k "$1"} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/round_summary.txt
tail -40 gpurun_out/pytest_gpu.log
if [ -z "$SKIP_BENCH" ]; then
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" | tee -a gpurun_out/round_summary.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
