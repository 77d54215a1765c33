#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -x -k zstacked --timeout 120 > gpurun_out/pytest_zs.log 2>&1
rc=$?
echo "zs pytest rc=$rc"; tail -5 gpurun_out/pytest_zs.log
if [ $rc -eq 0 ]; then E3B_ZS_PROF=1 timeout 300 python scripts/zs_bench.py 0 2>&1 | tee gpurun_out/zs_bench.log; fi
