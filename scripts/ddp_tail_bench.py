"""What follows the graph replay of a data-parallel train step (bench.py, N > 1), timed alone: flatten -> all-reduce ->
divide -> un-flatten -> fused SGD over the cfg-2 parameter set.  torchrun --nproc-per-node N scripts/ddp_tail_bench.py"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
import torch.distributed as dist

import elektronn3_b200 as e3

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
m = e3.UNet(n_blocks=3, start_filts=32, normalization='group').to(dev)
params = list(m.parameters())
for p in params:
    p.grad = torch.randn_like(p)
opt = torch.optim.SGD(params, lr=1e-3, momentum=0.9, fused=True)
grads = [p.grad for p in params]
flat0 = torch._utils._flatten_dense_tensors(grads)


def timed(fn, k=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / k * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def current():
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat)
    flat.div_(world)
    torch._foreach_copy_(grads, torch._utils._unflatten_dense_tensors(flat, grads))
    opt.step()


def avg():
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    torch._foreach_copy_(grads, torch._utils._unflatten_dense_tensors(flat, grads))
    opt.step()


res = dict(current=timed(current), avg=timed(avg), allreduce_only=timed(lambda: dist.all_reduce(flat0)),
           flatten_unflatten=timed(lambda: torch._foreach_copy_(grads, torch._utils._unflatten_dense_tensors(torch._utils._flatten_dense_tensors(grads), grads))),
           sgd=timed(opt.step))
if rank == 0:
    print(f'N={world}: ' + '  '.join(f'{k} {v:.1f} us' for k, v in res.items()), flush=True)
dist.destroy_process_group()
