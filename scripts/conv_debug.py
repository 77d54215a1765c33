"""Diagnostics for the tcgen05 conv kernel (run on the GPU box): where do one-hot inputs land?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from elektronn3_b200 import engine as eng
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_ops_gpu import to_qp_ref, from_qp_ref, dyadic

def qp(x):
    N, C, D, H, W = x.shape
    return eng.QP(to_qp_ref(x), N, C, D, H, W)

def run(x, w, k, pad, **kw):
    Co, C0 = w.shape[:2]
    wpk = eng.pack_weights(0, w, None, C0, 0, Co, k)
    y, _, _ = eng.conv_forward(qp(x), wpk, eng.cpad16(Co), Co, k, pad, **kw)
    torch.cuda.synchronize()
    return from_qp_ref(y.t, Co)

def report(tag, got, ref):
    err = (got.double() - ref.double()).abs()
    print(f'{tag}: max err {err.max().item():.4g} ; n_bad {(err > 1e-5).sum().item()} / {err.numel()}', flush=True)
    return err.max().item() < 1e-5

try:
    # A: 1 tap identity
    C, sp = 8, (2, 16, 8)
    x = dyadic((1, C) + sp, 1)
    w = torch.zeros((16, C, 1, 1, 1), device='cuda')
    for i in range(C): w[i, i] = 1
    got = run(x, w, (1, 1, 1), (0, 0, 0))
    ok = report('A identity 1-tap', got[:, :C], x)
    if not ok:
        for (c, z, y_, x_) in [(0, 0, 0, 0), (1, 0, 0, 0), (4, 0, 0, 0), (0, 0, 0, 1), (0, 0, 1, 0), (0, 1, 0, 0), (5, 1, 3, 2)]:
            xo = torch.zeros((1, C) + sp, device='cuda'); xo[0, c, z, y_, x_] = 1
            g = run(xo, w, (1, 1, 1), (0, 0, 0))
            nz = g.nonzero().tolist()
            print('  one-hot in (c,z,y,x)=', (c, z, y_, x_), '-> nonzero out (n,c,z,y,x):', nz[:8], [g[tuple(i)].item() for i in nz[:8]])
    # B: 1 tap, random weights
    w = dyadic((16, C, 1, 1, 1), 2)
    report('B random 1-tap', run(x, w, (1, 1, 1), (0, 0, 0)), F.conv3d(x, w))
    # C: 3x3x3 single-tap weights: only tap t nonzero (identity), should be a shifted copy
    for t in [13, 14, 12, 16, 10, 22, 4, 0, 26]:
        w = torch.zeros((16, C, 27), device='cuda')
        for i in range(C): w[i, i, t] = 1
        w = w.view(16, C, 3, 3, 3)
        report(f'C shift tap {t}', run(x, w, (3, 3, 3), (1, 1, 1)), F.conv3d(x, w, padding=1))
    # D: full 3x3x3
    w = dyadic((16, C, 3, 3, 3), 3, scale=4, lo=-2, hi=3)
    report('D 3x3x3', run(x, w, (3, 3, 3), (1, 1, 1)), F.conv3d(x, w, padding=1))
    # E: more channels / multiple chunks
    x = dyadic((1, 32, 4, 16, 16), 4, scale=4, lo=-4, hi=5)
    w = dyadic((32, 32, 3, 3, 3), 5, scale=4, lo=-2, hi=3)
    report('E 32->32', run(x, w, (3, 3, 3), (1, 1, 1)), F.conv3d(x, w, padding=1))
    report('E 32->32 tz1', run(x, w, (3, 3, 3), (1, 1, 1), force_tz=1), F.conv3d(x, w, padding=1))
except Exception as e:
    print('EXC', repr(e), flush=True)
