"""Bring-up aid for the z-stacked conv: runs one configuration per subprocess (a trapped kernel poisons its
context) and reports pass / fail / seconds.   python scripts/zs_debug.py"""
import os
import subprocess
import sys
import time

CASES = [  # grid cap, N, C0, Co, D, H, W, SA
    (1, 1, 16, 16, 6, 8, 8, 3), (1, 1, 16, 16, 4, 8, 8, 3), (1, 1, 16, 16, 9, 8, 8, 8),
]

if len(sys.argv) > 1:
    import torch
    import torch.nn.functional as F
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
    from elektronn3_b200 import engine as eng
    import test_ops_gpu as T
    g, N, C0, Co, D, H, W, SA = [int(v) for v in sys.argv[1:]]
    x = T.dyadic((N, C0, D, H, W), 1, scale=4, lo=-4, hi=5)
    w = T.dyadic((Co, C0, 3, 3, 3), 2, scale=4, lo=-2, hi=3)
    ref = F.conv3d(x.double(), w.double(), None, padding=1)
    wpk = eng.pack_weights(4, w, None, C0, 0, Co, (3, 3, 3))
    y, _, _ = eng.conv_forward(T.qp(eng, x), wpk, eng.cpad16(Co), Co, (3, 3, 3), (1, 1, 1), variant=1)
    try:
        torch.cuda.synchronize()
    except Exception as e:
        import ctypes
        from elektronn3_b200 import _lib
        buf = (ctypes.c_uint32 * 16)()
        _lib.lib().e3b_debug_zs_read(buf, 16)
        print('DBG', ' '.join(f'{v:x}' for v in buf))
        sys.exit(4)
    got = T.from_qp_ref(y.t, Co)
    err = (got.double() - ref).abs().amax(dim=(0, 1, 3, 4))
    print('max err per z plane:', [float(f'{e:.3g}') for e in err.tolist()])
    sys.exit(0 if float(err.max()) < 1e-5 else 3)

for c in CASES:
    env = dict(os.environ, E3B_ZS_GRID=str(c[0]), E3B_ZS_SA=str(c[7]), E3B_ZS_DEBUG='1')
    t0 = time.time()
    r = subprocess.run([sys.executable, __file__] + [str(v) for v in c], env=env, capture_output=True, text=True, timeout=120)
    tail = (r.stdout.strip().splitlines() or [''])[-1] if r.returncode in (0, 3, 4) else ' | '.join(l for l in r.stderr.splitlines() if 'rror' in l)
    print(c, 'rc', r.returncode, f'{time.time() - t0:.1f}s', tail[:300], flush=True)
