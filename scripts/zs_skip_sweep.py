import os, sys
sys.path.insert(0, '/root/repo')
import torch
from elektronn3_b200 import engine as eng
dev = torch.device('cuda')
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for (C0, C1) in ((32, 32), (32, 0)):
    w = torch.randn(32, C0 + C1, 3, 3, 3, device=dev) * 0.05
    var = eng.conv_variant(C0, C1, 32, (3, 3, 3))
    wpk = eng.pack_weights(4, w, None, C0, C1, 32, (3, 3, 3))
    q0 = eng.QP.empty_half(4, 32, 64, 64, 64, dev); q0.t.normal_()
    q1 = None
    if C1:
        q1 = eng.QP.empty_half(4, 32, 64, 64, 64, dev); q1.t.normal_()
    for rep in range(2):
        for skip in (0, 8, 1, 2, 3, 4):
            os.environ['E3B_ZS_SKIP'] = str(skip)
            t = timed(lambda: eng.conv_forward(q0, wpk, 32, 32, (3, 3, 3), (1, 1, 1), src1=q1, stats_channels=32, variant=var))
            print(f'{C0}+{C1}->32 skip={skip}: {t:7.1f} us', flush=True)
