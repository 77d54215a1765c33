#!/bin/bash
# 2-GPU check of bench.py (graph for forward+backward, eager NCCL all-reduce + optimizer), tightly bounded
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 100 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "graphed" --timeout 80 2>&1 | tail -3
E3B_BENCH_TIMEOUT=70 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2g.log 2>&1
echo rc=$?; grep -a "^{" gpurun_out/bench_n2g.log | cut -c1-700; grep -a -i "error\|Traceback" gpurun_out/bench_n2g.log | head -5
