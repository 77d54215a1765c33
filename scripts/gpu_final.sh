#!/bin/bash
# Round-end check: GPU test-suite, smoke(), bench line, Predictor (cfg 4) timing.  Logs -> gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests -m gpu -q --tb=short --timeout 120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench.json
timeout 150 python scripts/pred_bench.py > gpurun_out/pred_bench.jsonl 2> gpurun_out/pred_bench.err; echo "pred rc=$?"; cat gpurun_out/pred_bench.jsonl; tail -2 gpurun_out/pred_bench.err
