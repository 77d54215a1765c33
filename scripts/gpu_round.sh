#!/bin/bash
# One GPU-box visit: GPU test-suite, bench line, ncu launch list of one train step.  Logs -> gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1000 python -m pytest tests -m gpu -q --tb=short --timeout 180 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/round_summary.txt
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" | tee -a gpurun_out/round_summary.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
# (eager launches for the launch list: one warm-up step + one listed step)
E3B_BENCH_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --profile-steps 1 > gpurun_out/ncu_launch.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/round_summary.txt
python scripts/launch_summary.py gpurun_out/launches.csv 2 2>&1 | tail -40
