"""Kernel time of one forward+backward of the bench model (cfg 2) by kernel, from the torch profiler (no ncu needed,
warm caches, no serialisation): python scripts/step_profile.py [top_n]"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
from torch.profiler import ProfilerActivity, profile

import elektronn3_b200 as e3

top = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
m = e3.UNet(n_blocks=3, start_filts=32, normalization='group').cuda().train()
x = torch.randn(4, 1, 64, 64, 64, device='cuda')
for _ in range(3):
    m.zero_grad()
    m(x).sum().backward()
torch.cuda.synchronize()
reps = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(reps):
        m.zero_grad()
        m(x).sum().backward()
    torch.cuda.synchronize()
rows = sorted(((e.key, e.device_time_total / reps, e.count / reps) for e in prof.key_averages()), key=lambda r: -r[1])
print(f'total kernel time {sum(r[1] for r in rows):.0f} us per forward+backward')
for k, us, c in rows[:top]:
    print(f'  {k[:70]:70s} {us:8.1f} us  x{c:.0f}')
