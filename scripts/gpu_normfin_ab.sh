#!/bin/bash
# GPU test-suite, then the same-box A/B of the norm finalisation inside the activation kernel (default) against the separate
# launch (E3B_NORM_FIN=split): train-step bench line twice each.
cd "$(dirname "$0")/.."
SKIP_BENCH=1 bash scripts/gpu_round.sh
for rep in 1 2; do for m in fused split; do
  export E3B_NORM_FIN=$m
  echo "norm_fin=$m"; timeout 300 python bench.py --no-cpu-baseline --no-ref-gpu --no-predictor --steps 40 --warmup 10 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('train ms/step %.4f e2e %.4f launches/step %d' % (b['ms_per_step'], b['e2e']['ms_per_step'], b['gpu_launches'] // b['steps']))"
done; done
