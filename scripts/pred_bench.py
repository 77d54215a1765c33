"""BASELINE cfg 4: Predictor tiled inference of UNet(n_blocks=4) over a 512x512x256 synthetic volume,
tile_shape=(64,64,64), overlap=(8,8,8); host volume in, host result out (end to end), plus the
device-resident part alone.   python scripts/pred_bench.py [tile_batch | 0 = default] [D H W]   -> one JSON line
Run under torchrun for the sharded variant (one process per GPU, NCCL all_gather of the slabs)."""
import json
import os
import sys
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
import torch.distributed as dist

import elektronn3_b200 as e3

rank = int(os.environ.get('RANK', 0))
world = int(os.environ.get('WORLD_SIZE', 1))
lr = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
tb = int(sys.argv[1]) if len(sys.argv) > 1 and int(sys.argv[1]) > 0 else None      # None: the Predictor's default (a row of tiles)
vol = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (512, 512, 256)
torch.manual_seed(0)
m = e3.UNet(n_blocks=4, start_filts=32).to(dev).eval()
x = torch.randn((1, 1) + vol).pin_memory()
for argmax in (False, True):
    p = e3.Predictor(m, device=dev, tile_shape=(64, 64, 64), overlap_shape=(8, 8, 8), offset=(0, 0, 0),
                     out_shape=((1 if argmax else 2),) + vol, apply_softmax=True, apply_argmax=argmax, tile_batch=tb)
    p.predict(x)       # warm-up (weights packed, kernels loaded)
    ts = []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        out = p.predict(x)
        torch.cuda.synchronize()
        ts.append(time.time() - t0)
    if rank == 0:
        nvox = vol[0] * vol[1] * vol[2]
        print(json.dumps(dict(what=f'Predictor cfg4 volume {vol} tile 64^3 overlap 8, n_gpus={world}, tile_batch={tb}, '
                                   f'{"argmax uint8" if argmax else "softmax fp32"} output, host->host',
                              seconds=min(ts), all=ts, out_voxels_per_s=nvox / min(ts), tiles=p.last_stats['tiles'])))
if world > 1:
    dist.destroy_process_group()
