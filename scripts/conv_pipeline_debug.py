"""Cycle accounting of conv_tc_kernel's pipeline roles on representative layer shapes (GPU box).
E3B_CONV_DEBUG=1 python scripts/conv_pipeline_debug.py"""
import ctypes, os, sys
os.environ['E3B_CONV_DEBUG'] = '1'
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from elektronn3_b200 import engine as eng, _lib as L

def counters(reset=True):
    buf = (ctypes.c_ulonglong * 16)()
    torch.cuda.synchronize()
    L.check(L.lib().e3b_debug_conv_counters(ctypes.cast(buf, ctypes.c_void_p), 1 if reset else 0))
    return list(buf)

names = ['prod:a_empty', 'prod:b_empty', 'prod:total', 'mma:acc_empty', 'mma:a_full', 'mma:b_full', 'mma:issue', 'mma:total',
         'epi:acc_full', 'epi:total']
cases = [(4, 32, 0, 32, 64), (4, 32, 32, 32, 64), (4, 64, 0, 64, 32), (4, 64, 64, 64, 32), (4, 128, 0, 128, 16), (4, 1, 0, 32, 64)]
for (N, C0, C1, Co, S) in cases:
    dev = 'cuda'
    q0 = eng.QP.empty_half(N, C0, S, S, S, dev); q0.t.normal_()
    q1 = None
    if C1:
        q1 = eng.QP.empty_half(N, C1, S, S, S, dev); q1.t.normal_()
    w = torch.randn(Co, C0 + C1, 3, 3, 3, device=dev) * 0.05
    wpk = eng.pack_weights(0, w, None, C0, C1, Co, (3, 3, 3))
    for stats in (0, Co):
        for _ in range(2):
            eng.conv_forward(q0, wpk, eng.cpad16(Co), Co, (3, 3, 3), (1, 1, 1), src1=q1, stats_channels=stats)
        counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.conv_forward(q0, wpk, eng.cpad16(Co), Co, (3, 3, 3), (1, 1, 1), src1=q1, stats_channels=stats)
        e1.record()
        c = counters()
        ms = e0.elapsed_time(e1)
        gf = 2 * N * S ** 3 * Co * (C0 + C1) * 27 / 1e9
        print(f'case N={N} C={C0}+{C1}->{Co} S={S} stats={stats}: {ms*1e3:.0f} us, {gf/ms:.0f} TF/s')
        tot = c[7] or 1
        print('   ' + '  '.join(f'{n}={100*v/tot:.0f}%' for n, v in zip(names, c)))
