#!/bin/bash
# Two-GPU visit: the tests that need a second device, then bench.py under torchrun (train: weak scaling, Predictor: sharded).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
NG=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n$NG.txt 2>&1
timeout 900 python -m pytest tests/test_protocol_gpu.py -m gpu -q --tb=short --timeout 600 -k "sharded or second_device or data_parallel" > gpurun_out/pytest_n$NG.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_n$NG.log
E3B_BENCH_TIMEOUT=600 timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29500 \
  bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err
echo "bench rc=$?"; cat gpurun_out/bench_n$NG.json | cut -c1-3000; tail -5 gpurun_out/bench_n$NG.err | cut -c1-300
