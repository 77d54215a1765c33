bash scripts/gpu_round.sh 2>&1 | tail -25
bash scripts/gpu_ncu_zs.sh 2>&1 | tail -3
python scripts/ncu_summary.py gpurun_out/r02_zs_concat.ncu-rep > gpurun_out/r02_ncu_zs_concat.csv; cut -c1-260 gpurun_out/r02_ncu_zs_concat.csv
KERN=norm_bwd_fused_kernel OUT=r02_norm_fused CNT=14 SKIP=0 timeout 600 bash scripts/gpu_ncu_elem.sh > gpurun_out/r02_ncu_norm_fused.csv 2>&1; tail -15 gpurun_out/r02_ncu_norm_fused.csv | cut -c1-200
bash scripts/gpu_launches.sh > /dev/null 2>&1; head -30 gpurun_out/launch_summary.txt
