#!/bin/bash
# ncu launch list (device time per launch, cold-cache and serialised: compare SHARES) of one eager train step -> gpurun_out/launches.csv
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
E3B_BENCH_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --profile-steps 1 --no-predictor > gpurun_out/ncu_launch.log 2>&1
echo "ncu rc=$?"
python scripts/launch_summary.py gpurun_out/launches.csv 2 2>&1 | tee gpurun_out/launch_summary.txt | head -45
