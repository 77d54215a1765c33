#!/bin/bash
# z-stacked conv bring-up: its tests first (own process, bounded), then the full round.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -x -k zstacked --timeout 120 > gpurun_out/pytest_zs.log 2>&1
rc=$?
echo "zs pytest rc=$rc"; tail -30 gpurun_out/pytest_zs.log
if [ $rc -eq 0 ]; then bash scripts/gpu_round.sh; fi
