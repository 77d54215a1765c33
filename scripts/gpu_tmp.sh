mkdir -p gpurun_out
timeout 300 python scripts/debug_case.py 2>&1 | tail -30
timeout 600 python -m pytest tests/test_protocol_gpu.py -m gpu -q --tb=short --timeout 300 -k "dice" 2>&1 | tail -12
timeout 600 python bench.py --no-cpu-baseline --no-ref-gpu --no-predictor 2>/dev/null | python -c "import json,sys; b=json.load(sys.stdin); print('train ms', b['ms_per_step'], 'e2e', b['e2e']['ms_per_step'], 'launches', b['gpu_launches'])"
E3B_BENCH_TORCH_LOSS=1 timeout 600 python bench.py --no-cpu-baseline --no-ref-gpu --no-predictor 2>/dev/null | python -c "import json,sys; b=json.load(sys.stdin); print('torch-loss train ms', b['ms_per_step'], 'e2e', b['e2e']['ms_per_step'])"
