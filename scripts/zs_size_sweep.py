"""Fixed cost vs rate of the dominant conv launch (conv_zs, 32+32 -> 32 @ N x 64^3): time against the batch size N, in the
training form (fp32 output + statistics) and the inference form (bias + ReLU, fp16 output).  python scripts/zs_size_sweep.py"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch

from elektronn3_b200 import engine as eng

dev = torch.device('cuda')
PEAK = 1686.5


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


w = torch.randn(32, 64, 3, 3, 3, device=dev) * 0.05
bias = torch.zeros(32, device=dev)
var = eng.conv_variant(32, 32, 32, (3, 3, 3))
wpk = eng.pack_weights(4 if var else 0, w, None, 32, 32, 32, (3, 3, 3))
rows = []
for N in (1, 2, 4, 8, 16, 32):
    q0 = eng.QP.empty_half(N, 32, 64, 64, 64, dev)
    q1 = eng.QP.empty_half(N, 32, 64, 64, 64, dev)
    q0.t.normal_(), q1.t.normal_()
    gf = 2 * N * 64 ** 3 * 32 * 64 * 27 / 1e9
    t_train = timed(lambda: eng.conv_forward(q0, wpk, 32, 32, (3, 3, 3), (1, 1, 1), src1=q1, stats_channels=32, variant=var))
    t_eval = timed(lambda: eng.conv_forward(q0, wpk, 32, 32, (3, 3, 3), (1, 1, 1), src1=q1, bias=bias, relu=True, half_out=True, variant=var))
    rows.append((N, gf, t_train, t_eval))
    print(f'N={N:2d} {gf:7.1f} GF | train form {t_train * 1e3:7.1f} us {gf / t_train / PEAK:.3f} of peak | inference form {t_eval * 1e3:7.1f} us {gf / t_eval / PEAK:.3f}', flush=True)
    del q0, q1
# least-squares line t = a + b * GF over N >= 4
import numpy as np
for name, col in (('train form', 2), ('inference form', 3)):
    x = np.array([r[1] for r in rows if r[0] >= 4]); y = np.array([r[col] * 1e3 for r in rows if r[0] >= 4])
    b, a = np.polyfit(x, y, 1)
    print(f'{name}: t = {a:.1f} us + {b:.4f} us/GF  ->  asymptote {1e3 / b / PEAK:.3f} of peak')
