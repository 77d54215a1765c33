"""Timings of the z-stacked conv kernel on the cfg-2 full-resolution shapes (CUDA events), with the tuning
switches of E3B_ZS_SKIP (re-read per launch).   python scripts/zs_bench.py [skip masks ...]"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch

from elektronn3_b200 import engine as eng

N = 4
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


def qh(C, S):
    q = eng.QP.empty_half(N, C, S, S, S, dev)
    q.t.normal_()
    return q


masks = [int(m) for m in sys.argv[1:]] or [0]
SHAPES = [(1, 0, 32, 64), (32, 0, 32, 64), (32, 32, 32, 64), (32, 0, 64, 32)]
if os.environ.get('ZS_SHAPES'):
    SHAPES = [SHAPES[int(i)] for i in os.environ['ZS_SHAPES'].split(',')]
for (C0, C1, Co, S) in SHAPES:
    gf = 2 * N * S ** 3 * Co * (C0 + C1) * 27 / 1e9
    x0 = qh(C0, S)
    x1 = qh(C1, S) if C1 else None
    w = torch.randn(Co, C0 + C1, 3, 3, 3, device=dev) * 0.05
    dy = qh(Co, S)
    nd = eng.cpad16(eng.cpad8(C0) + (eng.cpad8(C1) if C1 else 0))
    for var in (0, 1):
        wpk = eng.pack_weights(4 if var else 0, w, None, C0, C1, Co, (3, 3, 3))
        wpd = eng.pack_weights(5 if var else 1, w, None, C0, C1, Co, (3, 3, 3))
        for m in (masks if var else [0]):
            os.environ['E3B_ZS_SKIP'] = str(m)
            ms = timed(lambda: eng.conv_forward(x0, wpk, eng.cpad16(Co), Co, (3, 3, 3), (1, 1, 1), src1=x1, stats_channels=Co, variant=var))
            ms_d = timed(lambda: eng.conv_forward(dy, wpd, nd, C0, (3, 3, 3), (1, 1, 1), dst1_C=C1, variant=var))
            if var and os.environ.get('E3B_ZS_PROF'):
                import ctypes
                from elektronn3_b200 import _lib
                buf = (ctypes.c_ulonglong * 16)()
                _lib.lib().e3b_debug_zs_prof(buf, 1)
                eng.conv_forward(x0, wpk, eng.cpad16(Co), Co, (3, 3, 3), (1, 1, 1), src1=x1, stats_channels=Co, variant=var)
                torch.cuda.synchronize()
                _lib.lib().e3b_debug_zs_prof(buf, 1)
                names = ['prod_total', 'prod_wait_empty', 'iss_total', 'iss_wait_free', 'iss_wait_full', 'iss_issue', 'epi_total',
                         'epi_wait_full', 'epi_ld', 'epi_store', 'epi_stats', 'epi_release', 'iss_setup', 'iss_mma', 'iss_commit', 'iss_other']
                print('   fwd kcycles/CTA:', ' '.join(f'{n}={buf[i] / 148e3:.1f}' for i, n in enumerate(names) if n))
            print(f'{C0}+{C1}->{Co} @{S}^3 {gf:6.1f} GF variant {var} skip {m:2d} | fwd {ms * 1e3:7.1f} us {gf / ms:6.1f} TF/s ({gf / ms / PEAK:.3f}) | '
                  f'dgrad {ms_d * 1e3:7.1f} us {gf / ms_d:6.1f} TF/s ({gf / ms_d / PEAK:.3f})', flush=True)
