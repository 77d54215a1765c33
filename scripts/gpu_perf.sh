#!/bin/bash
# Kernel-iteration visit: conv / network parity tests, then the per-layer benches and a bench line (no CPU / cuDNN arms).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_protocol_gpu.py -m gpu -q --tb=short --timeout 300 -x ${1:+-k "$1"} > gpurun_out/pytest_perf.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_perf.log
E3B_ZS_PROF=1 ZS_SHAPES=1,2 timeout 300 python scripts/zs_bench.py 0 > gpurun_out/zs_bench.txt 2>&1; cat gpurun_out/zs_bench.txt
timeout 300 python scripts/layer_bench.py > gpurun_out/layer_bench.txt 2>&1; cat gpurun_out/layer_bench.txt
timeout 600 python bench.py --no-cpu-baseline --no-ref-gpu > gpurun_out/bench_perf.json 2> gpurun_out/bench_perf.err
python - <<'PY'
import json
try:
    b = json.load(open('gpurun_out/bench_perf.json'))
    print('train ms/step', b['ms_per_step'], 'e2e', b['e2e']['ms_per_step'], 'roofline', b['roofline']['frac'], b['roofline']['ms_per_launch'])
    p = b.get('predictor') or {}
    print('predictor s/vol', p.get('seconds_per_volume'), 'e2e', (p.get('e2e') or {}).get('seconds_per_volume'), 'roofline', (p.get('roofline') or {}).get('frac'))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_perf.err').read()[-2000:])
PY
