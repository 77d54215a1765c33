#!/bin/bash
# ncu launch list of one warm Predictor pass (cfg-4 model, 128^3 volume = 8 tiles of 80^3 in one batch) -> gpurun_out/launches_pred.csv
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_pred.csv \
  python bench.py --profile-predictor > gpurun_out/ncu_launch_pred.log 2>&1
echo "ncu rc=$?"
python scripts/launch_summary.py gpurun_out/launches_pred.csv 1 2>&1 | tee gpurun_out/launch_summary_pred.txt | head -30
