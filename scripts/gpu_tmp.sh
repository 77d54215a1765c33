mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short --timeout 120 -x -k "wgrad or transposed" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_protocol_gpu.py -m gpu -q --tb=line --timeout 300 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --no-ref-gpu --no-predictor 2>gpurun_out/b.err | python -c "import json,sys; b=json.load(sys.stdin); print('train ms', b['ms_per_step'], 'e2e', b['e2e']['ms_per_step'])" || tail -20 gpurun_out/b.err
bash scripts/gpu_launches.sh 2>&1 | grep "wgrad\|total"
