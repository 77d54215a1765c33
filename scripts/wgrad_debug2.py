"""wgrad diagnostics, one configuration per process: python scripts/wgrad_debug2.py kd kh kw C0 Co D H W"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import torch
import torch.nn.functional as F
from elektronn3_b200 import engine as eng
from test_ops_gpu import to_qp_ref, dyadic, shifted_planar
kd, kh, kw, C0, Co, D, H, W = [int(v) for v in sys.argv[1:9]]
k = (kd, kh, kw); pad = (kd // 2, kh // 2, kw // 2)
x = dyadic((1, C0, D, H, W), 1, scale=2, lo=-2, hi=3); dy = dyadic((1, Co, D, H, W), 2, scale=2, lo=-2, hi=3)
w = torch.zeros((Co, C0) + k, dtype=torch.float64, device='cuda', requires_grad=True)
F.conv3d(x.double(), w, None, padding=pad).backward(dy.double())
def qp(t):
    N, C, d, h, ww = t.shape
    return eng.QP(to_qp_ref(t), N, C, d, h, ww, pl=eng.planar_from_ncdhw(t))
try:
    qd = qp(dy); qd.pl = shifted_planar(eng, dy, kw, pad[2], W)
    got = eng.wgrad(qp(x), qd, Co, k, pad, tuple(w.shape))
    torch.cuda.synchronize()
    err = (got.double() - w.grad).abs()
    print(sys.argv[1:], 'max err', err.max().item(), 'n_bad', (err > 1e-5).sum().item(), '/', err.numel(), flush=True)
    if err.max().item() > 1e-5:
        bad = (err > 1e-5).nonzero()
        print('  first bad idx (co,ci,kd,kh,kw):', bad[:10].tolist())
        print('  taps with errors:', sorted(set((b[2], b[3], b[4]) for b in bad.tolist())))
except Exception as e:
    print(sys.argv[1:], 'EXC', str(e).split('\n')[0], flush=True)
