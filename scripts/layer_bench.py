"""Per-layer kernel timings (CUDA events) of the cfg-2 conv shapes: forward / dgrad (the kernel variant the network uses) and wgrad.
  python scripts/layer_bench.py [N]        N = batch (default 4)"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch

from elektronn3_b200 import engine as eng

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device('cuda')
PEAK = 1686.5     # measured bf16 / f16 burst peak (MEASURED_PEAKS.json)


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


def qp_rand(C, S):
    q = eng.QP.empty_half(N, C, S, S, S, dev)
    q.t.normal_()
    return q


for (C0, C1, Co, S) in [(32, 0, 32, 64), (32, 32, 32, 64), (32, 0, 64, 32), (64, 0, 64, 32), (64, 64, 64, 32),
                        (64, 0, 128, 16), (128, 0, 128, 16)]:
    gf = 2 * N * S ** 3 * Co * (C0 + C1) * 27 / 1e9
    x0 = qp_rand(C0, S)
    x1 = qp_rand(C1, S) if C1 else None
    w = torch.randn(Co, C0 + C1, 3, 3, 3, device=dev) * 0.05
    var = eng.conv_variant(C0, C1, eng.cpad16(Co), (3, 3, 3))
    wpk = eng.pack_weights(4 if var else 0, w, None, C0, C1, Co, (3, 3, 3))
    ms = timed(lambda: eng.conv_forward(x0, wpk, eng.cpad16(Co), Co, (3, 3, 3), (1, 1, 1), src1=x1, stats_channels=Co, variant=var))
    dy = qp_rand(Co, S)
    ms_w = timed(lambda: eng.wgrad(x0, dy, Co, (3, 3, 3), (1, 1, 1), (Co, C0 + C1, 3, 3, 3), src1=x1))
    nd = eng.cpad16(eng.cpad8(C0) + (eng.cpad8(C1) if C1 else 0))
    dvar = eng.conv_variant(Co, 0, nd, (3, 3, 3))
    wpd = eng.pack_weights(5 if dvar else 1, w, None, C0, C1, Co, (3, 3, 3))
    ms_d = timed(lambda: eng.conv_forward(dy, wpd, nd, C0, (3, 3, 3), (1, 1, 1), dst1_C=C1, variant=dvar))
    print(f'N={N} {C0}+{C1}->{Co} @{S}^3  {gf:7.2f} GF | fwd {ms * 1e3:7.1f} us {gf / ms:6.1f} TF/s ({gf / ms / PEAK:.3f}) | '
          f'dgrad {ms_d * 1e3:7.1f} us {gf / ms_d:6.1f} TF/s ({gf / ms_d / PEAK:.3f}) | '
          f'wgrad {ms_w * 1e3:7.1f} us {gf / ms_w:6.1f} TF/s ({gf / ms_w / PEAK:.3f})', flush=True)
