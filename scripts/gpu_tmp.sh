mkdir -p gpurun_out
NG=8
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n$NG.txt 2>&1
E3B_BENCH_TIMEOUT=500 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29500 \
  bench.py --gpus $NG --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err
echo "bench rc=$?"; grep '^{' gpurun_out/bench_n$NG.json | tail -1 | python -c "
import json,sys
b=json.loads(sys.stdin.read()); p=b.get('predictor',{})
print('N',b['n_gpus'],'train ms',b['ms_per_step'],'Mvox/s',b['value']/1e6,'e2e ms',b['e2e']['ms_per_step'])
print('pred s',p.get('seconds_per_volume'),'e2e',p.get('e2e',{}).get('seconds_per_volume'),'Mvox/s',p.get('value',0)/1e6)
"; tail -3 gpurun_out/bench_n$NG.err | cut -c1-300
NG=4
E3B_BENCH_TIMEOUT=500 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29501 \
  bench.py --gpus $NG --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err
echo "bench rc=$?"; grep '^{' gpurun_out/bench_n$NG.json | tail -1 | python -c "
import json,sys
b=json.loads(sys.stdin.read()); p=b.get('predictor',{})
print('N',b['n_gpus'],'train ms',b['ms_per_step'],'Mvox/s',b['value']/1e6,'e2e ms',b['e2e']['ms_per_step'])
print('pred s',p.get('seconds_per_volume'),'e2e',p.get('e2e',{}).get('seconds_per_volume'),'Mvox/s',p.get('value',0)/1e6)
"; tail -3 gpurun_out/bench_n$NG.err | cut -c1-300
