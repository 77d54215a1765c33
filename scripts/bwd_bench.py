"""Timings of the norm backward kernels on a full-resolution cfg-2 layer (4 x 32ch x 64^3) under the tuning switches
E3B_RED_VPT / E3B_X4_ITER (read once per process): python scripts/bwd_bench.py"""
import os
import subprocess
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)

if len(sys.argv) > 1:
    import torch
    import elektronn3_b200 as e3
    from elektronn3_b200 import engine as eng
    torch.manual_seed(0)
    m = e3.UNet(n_blocks=3, start_filts=32, normalization='group').cuda().train()
    x = torch.randn(4, 1, 64, 64, 64, device='cuda')
    t = torch.randint(0, 2, (4, 64, 64, 64), device='cuda')
    for i in range(3):
        m.zero_grad(); m(x).sum().backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(5):
            m.zero_grad(); m(x).sum().backward()
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total / 5, e.count / 5) for e in prof.key_averages()]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    print(f'total kernel time {tot:.0f} us/step')
    for k, us, c in rows[:14]:
        print(f'  {k[:60]:60s} {us:8.1f} us  x{c:.0f}')
    sys.exit(0)

for env in [{}, {'E3B_RED_VPT': '2'}, {'E3B_RED_VPT': '4'}, {'E3B_X4_ITER': '4'}]:
    r = subprocess.run([sys.executable, __file__, 'child'], env=dict(os.environ, **env), capture_output=True, text=True)
    print('==', env)
    print(r.stdout[-1600:] if r.returncode == 0 else r.stderr[-800:], flush=True)
