"""Transposed conv (k = s = 2, scatter epilogue of conv_tc) at the cfg-2 / cfg-4 shapes: CUDA-event timings (L2 flushed),
output bandwidth, and the per-role cycle shares (needs E3B_CONV_DEBUG=1)."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from elektronn3_b200 import engine as eng, _lib as L

names = ['prod:a_empty', 'prod:b_empty', 'prod:total', 'mma:acc_empty', 'mma:a_full', 'mma:b_full', 'mma:issue', 'mma:total',
         'epi:acc_full', 'epi:total']


def counters():
    buf = (ctypes.c_ulonglong * 16)()
    torch.cuda.synchronize()
    L.check(L.lib().e3b_debug_conv_counters(ctypes.cast(buf, ctypes.c_void_p), 1))
    return list(buf)


def timed(fn, iters=10):
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


for (N, Ci, Co, S) in [(4, 64, 32, 32), (4, 128, 64, 16), (8, 64, 32, 40)]:
    dev = 'cuda'
    x = eng.QP.empty_half(N, Ci, S, S, S, dev); x.t.normal_()
    up = torch.nn.ConvTranspose3d(Ci, Co, 2, 2).cuda()
    s = (2, 2, 2)
    wpk = eng.pack_weights(2, up.weight.detach(), None, Ci, 0, Co, s)
    n_total = 8 * eng.cpad16(Co)
    out_sp = (2 * S, 2 * S, 2 * S)
    vox = N * (2 * S) ** 3
    for what, kw, obytes in (('train (fp32 y + stats)', dict(stats_channels=Co), 4), ('eval (relu, fp16 out)', dict(relu=True, half_out=True), 2)):
        f = lambda: eng.conv_forward(x, wpk, n_total, Co, (1, 1, 1), (0, 0, 0), bias=up.bias.detach(), scatter=s, out_spatial=out_sp, **kw)
        us = timed(f)
        mb = vox * Co * obytes / 1e6 + N * S ** 3 * Ci * 2 / 1e6
        line = f'upconv {Ci}->{Co} {N}x{S}^3 -> {2 * S}^3 {what}: {us:7.1f} us  {mb:6.1f} MB  {mb / us / 1e3:5.2f} TB/s'
        if os.environ.get('E3B_CONV_DEBUG'):
            counters(); f(); c = counters()
            tot = c[7] or 1
            line += '   ' + ' '.join(f'{n}={100 * v / tot:.0f}%' for n, v in zip(names, c))
        print(line, flush=True)
