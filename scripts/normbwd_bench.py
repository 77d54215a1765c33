"""Norm-backward alone at the shapes of BASELINE cfg 2 (GroupNorm, batch 4): fused (one persistent kernel) vs split (reduce /
finalize / apply), CUDA-event timing with an L2 flush between iterations, and -- with E3B_FUSED_PROF=1 -- the fused kernel's
phase timeline (first and last CTA)."""
import ctypes
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elektronn3_b200 import engine as eng, _lib as L


def qp32(x):
    """NCDHW float32 -> QP (N, C/4, D, H, W, 4)"""
    N, C, D, H, W = x.shape
    Cp = (C + 7) & ~7
    xp = torch.zeros((N, Cp, D, H, W), dtype=x.dtype, device=x.device)
    xp[:, :C] = x
    return eng.QP(xp.view(N, Cp // 4, 4, D, H, W).permute(0, 1, 3, 4, 5, 2).contiguous(), N, C, D, H, W)


def setup(N, C, sp, G=8, g1=False):
    torch.manual_seed(0)
    y = torch.randn((N, C) + sp, device='cuda')
    yq = qp32(y)
    stats = torch.stack((y.double().sum(dim=(2, 3, 4)), (y.double() ** 2).sum(dim=(2, 3, 4))), dim=-1).contiguous()
    gamma, beta = torch.ones(C, device='cuda'), torch.zeros(C, device='cuda')
    S = sp[0] * sp[1] * sp[2]
    nstate = eng.norm_finalize(stats, 1, G, N, C, S, gamma, beta, 1e-5, None, None, 0.1, y.device)
    a, _ = eng.norm_act(yq, nstate.scale, nstate.shift, save=True)
    u = eng.Unit()

    class Spec:
        pass
    u.spec = Spec()
    u.spec.norm = torch.nn.GroupNorm(G, C).cuda()
    u.a, u.y, u.pool, u.mode, u.G, u.nstate, u.stats, u.pooled = a, yq, None, 1, G, nstate, stats, None
    g0 = qp32(torch.randn((N, C) + sp, device='cuda'))
    g1q = qp32(torch.randn((N, C) + sp, device='cuda')) if g1 else None
    return u, C, g0, g1q


def timeit(fn, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for name, N, C, sp, g1 in [('full-res 32ch', 4, 32, (64, 64, 64), False), ('full-res 32ch + skip grad', 4, 32, (64, 64, 64), True),
                           ('half-res 64ch', 4, 64, (32, 32, 32), False), ('quarter-res 128ch', 4, 128, (16, 16, 16), False)]:
    u, C, g0, g1q = setup(N, C, sp, g1=g1)
    out = {}
    for path in ('fused', 'split'):
        os.environ['E3B_NORM_BWD'] = path
        out[path] = timeit(lambda: eng._norm_bwd(u, C, g0, g1=g1q))
    print('%-28s fused %7.1f us   split %7.1f us' % (name, out['fused'], out['split']))
    if os.environ.get('E3B_FUSED_PROF'):
        os.environ['E3B_NORM_BWD'] = 'fused'
        eng._norm_bwd(u, C, g0, g1=g1q); torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 64)()
        L.check(L.lib().e3b_debug_fused_prof(buf), 'prof')
        for cta in range(2):
            t00 = buf[cta * 32]
            for r in range(4):
                st = [buf[(cta * 4 + r) * 8 + k] for k in range(8)]
                if st[0] == 0:
                    continue
                print('   cta %s round %d: start %6.1f | A %5.1f (wait %4.1f) | barrier %5.1f | B %4.1f | C %5.1f (wait %4.1f) us' % (
                    'first' if cta == 0 else 'last ', r, (st[0] - t00) / 1e3, (st[1] - st[0]) / 1e3, st[5] / 1e3, (st[2] - st[1]) / 1e3,
                    (st[3] - st[2]) / 1e3, (st[4] - st[3]) / 1e3, st[6] / 1e3))
