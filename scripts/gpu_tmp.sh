timeout 900 python -m pytest tests/test_protocol_gpu.py tests/test_unet_gpu.py -m gpu -q --tb=short --timeout 300 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline --no-ref-gpu --no-predictor 2>gpurun_out/b.err | python -c "import json,sys; b=json.load(sys.stdin); print('train ms', b['ms_per_step'], 'e2e', b['e2e']['ms_per_step'], b['e2e']['mode'])" || tail -20 gpurun_out/b.err
