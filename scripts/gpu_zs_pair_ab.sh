#!/bin/bash
# Same-box A/B of the z-stacked conv with the two chains of a CTA on x-adjacent tile columns (default) against the two halves
# of one contiguous run (E3B_ZS_NO_PAIR=1): conv parity tests, the dominant launch, the bench line.  -> gpurun_out/zs_pair_ab.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short --timeout 120 -k "conv" 2>&1 | tail -4
: > gpurun_out/zs_pair_ab.txt
for rep in 1 2; do
for nopair in 0 1; do
  if [ $nopair = 1 ]; then export E3B_ZS_NO_PAIR=1; else unset E3B_ZS_NO_PAIR; fi
  (echo "no_pair=$nopair"; timeout 200 python scripts/zs_size_sweep.py 2>/dev/null | grep -E "^N= 4|^N= 8"
   timeout 300 python bench.py --no-cpu-baseline --no-ref-gpu --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read()); p=b['predictor']; print('train %.4f pred %.4f e2e %.4f domfrac %.3f preddom %.3f' % (b['ms_per_step'], p['seconds_per_volume'], p['e2e']['seconds_per_volume'], b['roofline']['frac'], p['roofline']['frac']))") | tee -a gpurun_out/zs_pair_ab.txt
done
done
