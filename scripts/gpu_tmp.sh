mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short --timeout 120 -x -k "norm" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_protocol_gpu.py tests/test_parity_fullsize_gpu.py -m gpu -q --tb=line --timeout 300 2>&1 | tail -6
E3B_FUSED_PROF=1 timeout 300 python scripts/normbwd_bench.py 2>&1 | grep -v "round [123]\|cta last" | tail -20
timeout 600 python bench.py --no-cpu-baseline --no-ref-gpu --no-predictor 2>gpurun_out/b.err | python -c "import json,sys; b=json.load(sys.stdin); print('fused train ms', b['ms_per_step'], 'e2e', b['e2e']['ms_per_step'], 'launches', b['gpu_launches'])" || tail -20 gpurun_out/b.err
E3B_NORM_BWD=split timeout 600 python bench.py --no-cpu-baseline --no-ref-gpu --no-predictor 2>/dev/null | python -c "import json,sys; b=json.load(sys.stdin); print('split train ms', b['ms_per_step'], 'e2e', b['e2e']['ms_per_step'], 'launches', b['gpu_launches'])"
