"""Diagnostics for the tcgen05 wgrad kernel (run on the GPU box)."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
import torch.nn.functional as F
from elektronn3_b200 import engine as eng
from test_ops_gpu import to_qp_ref, from_qp_ref, dyadic

def qp(x):
    N, C, D, H, W = x.shape
    return eng.QP(to_qp_ref(x), N, C, D, H, W, pl=eng.planar_from_ncdhw(x))

def ref_wgrad(x, dy, k, pad, Co):
    w = torch.zeros((Co, x.shape[1]) + k, dtype=torch.float64, device='cuda', requires_grad=True)
    F.conv3d(x.double(), w, None, padding=pad).backward(dy.double())
    return w.grad

def go(tag, x, dy, k, pad):
    Co = dy.shape[1]
    ref = ref_wgrad(x, dy, k, pad, Co)
    got = eng.wgrad(qp(x), qp(dy), Co, k, pad, tuple(ref.shape))
    torch.cuda.synchronize()
    err = (got.double() - ref).abs()
    print(f'{tag}: max err {err.max().item():.4g} n_bad {(err > 1e-5).sum().item()}/{err.numel()}  |got|max {got.abs().max().item():.4g} |ref|max {ref.abs().max().item():.4g}', flush=True)
    return got, ref

try:
    sp = (2, 8, 8)
    # 1-tap GEMM, one-hot
    for (c, v, co, v2) in [((0), (0, 0, 0), 0, (0, 0, 0)), (1, (0, 0, 0), 0, (0, 0, 0)), (0, (0, 0, 0), 1, (0, 0, 0)),
                           (5, (0, 0, 3), 9, (0, 0, 3)), (0, (1, 2, 3), 4, (1, 2, 3)), (2, (0, 0, 1), 3, (0, 0, 0))]:
        x = torch.zeros((1, 8) + sp, device='cuda'); x[(0, c) + v] = 1
        dy = torch.zeros((1, 16) + sp, device='cuda'); dy[(0, co) + v2] = 1
        got, ref = go(f'1tap onehot x[c={c},v={v}] dy[co={co},v={v2}]', x, dy, (1, 1, 1), (0, 0, 0))
        print('   got nz (co,c,..):', got.nonzero().tolist()[:6], ' ref nz:', ref.nonzero().tolist()[:6])
    x = dyadic((1, 8) + sp, 1, scale=2, lo=-2, hi=3); dy = dyadic((1, 16) + sp, 2, scale=2, lo=-2, hi=3)
    go('1tap random', x, dy, (1, 1, 1), (0, 0, 0))
    go('3x3x3 random', x, dy, (3, 3, 3), (1, 1, 1))
    for (v, v2) in [((0, 3, 3), (0, 3, 3)), ((0, 3, 4), (0, 3, 3)), ((0, 4, 3), (0, 3, 3)), ((1, 3, 3), (0, 3, 3))]:
        x = torch.zeros((1, 8) + sp, device='cuda'); x[(0, 2) + v] = 1
        dy = torch.zeros((1, 16) + sp, device='cuda'); dy[(0, 5) + v2] = 1
        got, ref = go(f'3x3x3 onehot x v={v} dy v={v2}', x, dy, (3, 3, 3), (1, 1, 1))
        print('   got nz:', got.nonzero().tolist()[:6], ' ref nz:', ref.nonzero().tolist()[:6])
    x = dyadic((2, 32, 6, 20, 18), 3, scale=2, lo=-2, hi=3); dy = dyadic((2, 32, 6, 20, 18), 4, scale=2, lo=-2, hi=3)
    go('32x32 3x3x3', x, dy, (3, 3, 3), (1, 1, 1))
except Exception as e:
    import traceback; traceback.print_exc()
