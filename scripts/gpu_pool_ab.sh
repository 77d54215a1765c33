#!/bin/bash
# Same-box A/B of the inference pooling kernel (16-byte units, default) against the generic one (E3B_POOL_GENERIC=1): tests,
# then the Predictor bench twice each.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -m gpu -q --tb=short --timeout 120 -k "pooling_of or golden or predictor" 2>&1 | tail -4
for rep in 1 2; do for g in 0 1; do
  if [ $g = 1 ]; then export E3B_POOL_GENERIC=1; else unset E3B_POOL_GENERIC; fi
  echo "generic=$g"; timeout 200 python scripts/pred_bench.py 2>/dev/null | head -1 | cut -c1-200
done; done
