mkdir -p gpurun_out
for ty in 8 4 2; do echo "TY=$ty"; E3B_WGRAD_TY=$ty timeout 200 python scripts/layer_bench.py 2>&1 | awk -F'|' '{print $1 "|" $4}'; done
