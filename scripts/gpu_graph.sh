#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -q --tb=short -x -k "graphed or dice" --timeout 200 2>&1 | tail -15
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "bench rc=$?"; cat gpurun_out/bench_graph.json; tail -3 gpurun_out/bench_graph.err
E3B_BENCH_GRAPH=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_eager.json 2>/dev/null; cat gpurun_out/bench_eager.json | cut -c1-400
