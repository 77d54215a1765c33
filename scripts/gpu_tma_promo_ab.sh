#!/bin/bash
# Same-box A/B of the TMA L2 promotion size of the conv / wgrad tensor maps (E3B_TMA_PROMO: 0 none, 1 64 B, 2 128 B = default,
# 3 256 B): the dominant conv launch alone, then the train-step bench line.  -> gpurun_out/tma_promo_ab.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/tma_promo_ab.txt
for rep in 1 2; do
for promo in 2 0 1; do
  export E3B_TMA_PROMO=$promo
  (echo "promo=$promo"; timeout 200 python scripts/zs_size_sweep.py 2>/dev/null | grep -E "^N= 4|^N= 8"
   timeout 300 python bench.py --no-cpu-baseline --no-ref-gpu --no-predictor --steps 40 --warmup 10 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('train ms/step %.4f e2e %.4f' % (b['ms_per_step'], b['e2e']['ms_per_step']))") | tee -a gpurun_out/tma_promo_ab.txt
done
done
