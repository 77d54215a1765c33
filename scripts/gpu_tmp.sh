bash scripts/gpu_round.sh 2>&1 | tail -12
KERN=norm_bwd_fused_kernel OUT=r02_norm_fused CNT=14 SKIP=0 timeout 600 bash scripts/gpu_ncu_elem.sh > gpurun_out/r02_ncu_norm_fused.csv 2>&1; tail -14 gpurun_out/r02_ncu_norm_fused.csv | cut -c1-120
bash scripts/gpu_launches.sh > /dev/null 2>&1; head -24 gpurun_out/launch_summary.txt
E3B_FUSED_PROF=1 timeout 300 python scripts/normbwd_bench.py > gpurun_out/r02_norm_bwd_fused_timeline.txt 2>&1
timeout 300 python scripts/layer_bench.py > gpurun_out/r02_layer_bench_final.txt 2>&1
E3B_CONV_DEBUG=1 timeout 300 python scripts/upconv_bench.py > gpurun_out/r02_upconv_bench.txt 2>&1
