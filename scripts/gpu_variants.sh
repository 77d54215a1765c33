#!/bin/bash
# GPU test-suite, then the bench line without the CPU arm / Predictor: headline step + the variants list
cd "$(dirname "$0")/.."
SKIP_BENCH=1 bash scripts/gpu_round.sh
timeout 300 python bench.py --no-cpu-baseline --no-predictor 2>/dev/null > gpurun_out/bench_variants.json
python - <<'PY'
import json
b = json.load(open('gpurun_out/bench_variants.json'))
print('train', b['ms_per_step'], 'e2e', b['e2e']['ms_per_step'])
for v in b['variants']:
    print(v['workload'][:70], round(v['ms_per_step'], 3), round(v['ref_gpu_ms_per_step'], 2), round(v['speedup_device'], 2))
PY
