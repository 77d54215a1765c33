#!/bin/bash
# Same-box A/B of the fused norm backward with the un-pooling inputs staged with the item (FusedVar<5>, default) against the
# per-voxel L2 gathers (E3B_FUSED_NO_STAGED_POOL=1): kernel tests, then the train-step bench line twice each.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short --timeout 120 -k "direct_plus_skip or norm_act_pool" 2>&1 | tail -6
for v in 1 0; do
  if [ $v = 0 ]; then export E3B_FUSED_NO_STAGED_POOL=1; else unset E3B_FUSED_NO_STAGED_POOL; fi
  echo "staged=$v"; timeout 300 python bench.py --no-cpu-baseline --no-ref-gpu --no-predictor --steps 40 --warmup 10 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('train ms/step %.4f e2e %.4f' % (b['ms_per_step'], b['e2e']['ms_per_step']))"
done
for v in 1 0; do
  if [ $v = 0 ]; then export E3B_FUSED_NO_STAGED_POOL=1; else unset E3B_FUSED_NO_STAGED_POOL; fi
  echo "staged=$v"; timeout 300 python bench.py --no-cpu-baseline --no-ref-gpu --no-predictor --steps 40 --warmup 10 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('train ms/step %.4f e2e %.4f' % (b['ms_per_step'], b['e2e']['ms_per_step']))"
done
