#!/bin/bash
# Run the GPU test-suite one test function per process (a trapped kernel poisons its CUDA context),
# each under its own timeout; logs -> gpurun_out/.   usage: scripts/gpu_check.sh [pytest -k expression]
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
sel="$1"
summary=gpurun_out/check_summary.txt
: > $summary
for t in $(python -m pytest tests -m gpu --collect-only -q ${sel:+-k "$sel"} 2>/dev/null | grep '::' | sed 's/\[.*//' | sort -u); do
  name=$(echo "$t" | tr '/:' '__')
  timeout 600 python -m pytest "$t" -m gpu -q --tb=short -x > "gpurun_out/$name.log" 2>&1
  rc=$?
  echo "$rc $t $(tail -1 gpurun_out/$name.log)" | tee -a $summary
done
