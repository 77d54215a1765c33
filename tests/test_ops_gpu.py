"""Per-kernel parity tests of libe3b.so through the C ABI (ctypes) on a B200.

Inputs are small dyadic rationals (k/8), exactly representable in fp16 / TF32, so that tensor-core products
are exact and the comparison against a float64 torch evaluation of the same operator is tight: any
indexing / layout / pipeline bug shows up as an O(1) error instead of hiding under TF32 noise.
Real-valued inputs (TF32 tolerance) are covered by the golden-vector tests in test_unet_gpu.py.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from elektronn3_b200 import engine
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return engine


def dyadic(shape, seed, scale=8, lo=-8, hi=9):
    rs = np.random.RandomState(seed)
    return torch.from_numpy(rs.randint(lo, hi, size=shape).astype(np.float32) / scale).cuda()


def to_qp_ref(x):
    """pure-torch NCDHW -> QP (test-side restatement of the layout in csrc/common.cuh)"""
    N, C, D, H, W = x.shape
    Cp = (C + 7) & ~7
    xp = torch.zeros((N, Cp, D, H, W), dtype=x.dtype, device=x.device)
    xp[:, :C] = x
    return xp.view(N, Cp // 4, 4, D, H, W).permute(0, 1, 3, 4, 5, 2).contiguous()


def from_qp_ref(t, C):
    N, Cq, D, H, W, _ = t.shape
    return t.permute(0, 1, 5, 2, 3, 4).reshape(N, Cq * 4, D, H, W)[:, :C].contiguous()


def to_qh_ref(x):
    """pure-torch NCDHW -> QH, the fp16 operand layout (N, ceil16(C)/8, D, H, W, 8) of csrc/common.cuh"""
    N, C, D, H, W = x.shape
    Cp = (C + 15) & ~15
    xp = torch.zeros((N, Cp, D, H, W), dtype=torch.float16, device=x.device)
    xp[:, :C] = x.to(torch.float16)
    return xp.view(N, Cp // 8, 8, D, H, W).permute(0, 1, 3, 4, 5, 2).contiguous()


def from_qh_ref(q, C):
    """QH operand tensor -> float32 NCDHW; a scaled gradient (q.scale) is un-scaled"""
    t = q.t
    N, Ch, D, H, W, _ = t.shape
    out = t.permute(0, 1, 5, 2, 3, 4).reshape(N, Ch * 8, D, H, W)[:, :C].float().contiguous()
    if q.scale is not None:
        out = out * q.scale[2]
    return out


def qp(eng, x):
    """MMA operand tensor (QH)"""
    N, C, D, H, W = x.shape
    return eng.QP(to_qh_ref(x), N, C, D, H, W)


def qp32(eng, x):
    """float32 QP tensor (conv outputs, gradients w.r.t. activations)"""
    N, C, D, H, W = x.shape
    return eng.QP(to_qp_ref(x), N, C, D, H, W)


def assert_close(got, ref, tol, what=''):
    ref = ref.to(torch.float64)
    err = (got.to(torch.float64) - ref).abs().max().item()
    scale = max(ref.abs().max().item(), 1e-6)
    assert err / scale < tol, f'{what}: max err {err:.3e} (scale {scale:.3e})'


def test_pack_unpack_layout(eng):
    x = dyadic((2, 5, 3, 6, 7), 0)
    q = eng.pack_input(x)
    assert q.half and torch.equal(q.t, to_qh_ref(x))
    assert torch.equal(eng.unpack(qp32(eng, x)), x)
    x1 = dyadic((1, 1, 4, 4, 9), 1)
    assert torch.equal(eng.pack_input(x1).t, to_qh_ref(x1))
    x2 = dyadic((2, 19, 3, 4, 5), 2)                      # two 16-channel chunks, the second one ragged
    qq = eng.pack_input(x2)
    assert torch.equal(qq.t, to_qh_ref(x2))


CONV_CASES = [
    # N, C0, Co, (D,H,W), k, pad
    (1, 8, 16, (4, 16, 8), (3, 3, 3), (1, 1, 1)),
    (1, 8, 16, (4, 16, 8), (1, 1, 1), (0, 0, 0)),
    (2, 1, 32, (16, 16, 16), (3, 3, 3), (1, 1, 1)),
    (1, 32, 32, (16, 18, 20), (3, 3, 3), (1, 1, 1)),
    (1, 64, 64, (8, 16, 16), (3, 3, 3), (1, 1, 1)),
    (1, 128, 128, (6, 8, 8), (3, 3, 3), (1, 1, 1)),
    (1, 256, 256, (4, 8, 8), (3, 3, 3), (1, 1, 1)),
    (1, 16, 24, (5, 32, 32), (1, 3, 3), (0, 1, 1)),       # planar
    (2, 8, 8, (1, 40, 24), (1, 3, 3), (0, 1, 1)),         # 2D
    (1, 3, 8, (11, 13, 18), (3, 3, 3), (1, 1, 1)),        # odd extents, ragged channels
    (1, 8, 8, (12, 20, 20), (3, 3, 3), (0, 0, 0)),        # VALID
    (1, 40, 48, (6, 10, 10), (3, 3, 3), (1, 1, 1)),
]


@pytest.mark.parametrize('case', CONV_CASES, ids=[str(c) for c in CONV_CASES])
def test_conv_forward_bias_relu_stats(eng, case):
    N, C0, Co, sp, k, pad = case
    x = dyadic((N, C0) + sp, 1, scale=4, lo=-4, hi=5)
    w = dyadic((Co, C0) + k, 2, scale=4, lo=-2, hi=3)
    b = dyadic((Co,), 3)
    wpk = eng.pack_weights(0, w, None, C0, 0, Co, k)
    ref = F.conv3d(x.double(), w.double(), b.double(), padding=pad)
    y, _, stats = eng.conv_forward(qp(eng, x), wpk, eng.cpad16(Co), Co, k, pad, bias=b, stats_channels=Co)
    torch.cuda.synchronize()
    got = from_qp_ref(y.t, Co)
    assert got.shape == ref.shape
    assert_close(got, ref, 1e-6, 'conv')
    # padding channels must be exactly zero
    Cp = (Co + 7) & ~7
    if Cp != Co:
        full = y.t.permute(0, 1, 5, 2, 3, 4).reshape(N, Cp, *ref.shape[2:])
        assert full[:, Co:].abs().max().item() == 0.0
    assert_close(stats[:, :, 0], ref.sum(dim=(2, 3, 4)), 1e-6, 'sum')
    assert_close(stats[:, :, 1], (ref * ref).sum(dim=(2, 3, 4)), 1e-6, 'sumsq')
    yr, _, _ = eng.conv_forward(qp(eng, x), wpk, eng.cpad16(Co), Co, k, pad, bias=b, relu=True)
    assert_close(from_qp_ref(yr.t, Co), ref.clamp_min(0), 1e-6, 'conv+relu')
    # the same, written directly as the next layer's fp16 operand (eval path)
    yh, _, _ = eng.conv_forward(qp(eng, x), wpk, eng.cpad16(Co), Co, k, pad, bias=b, relu=True, half_out=True)
    assert yh.half and yh.t.shape[1] == eng.cpad16(Co) // 8
    assert_close(from_qh_ref(yh, Co), ref.clamp_min(0), 1e-3, 'conv+relu -> QH')
    if eng.cpad16(Co) != Co:
        assert from_qh_ref(yh, eng.cpad16(Co))[:, Co:].abs().max().item() == 0.0


ZS_CASES = [
    # N, C0, C1, Co, (D,H,W), pad       -- the z-stacked kernel (csrc/conv_zs.cu, e3b_conv_args.variant = 1)
    (1, 8, 0, 16, (4, 16, 8), (1, 1, 1)),
    (2, 1, 0, 32, (16, 16, 16), (1, 1, 1)),
    (1, 32, 0, 32, (16, 18, 20), (1, 1, 1)),
    (1, 32, 32, 32, (9, 16, 16), (1, 1, 1)),              # virtual concat
    (1, 16, 24, 16, (5, 16, 8), (1, 1, 1)),               # virtual concat, ragged second source
    (1, 3, 0, 8, (11, 13, 18), (1, 1, 1)),                # odd extents, ragged channels
    (1, 8, 0, 8, (12, 20, 20), (0, 0, 0)),                # VALID
    (1, 8, 0, 8, (6, 12, 12), (2, 2, 2)),                 # full correlation (the dgrad geometry of VALID)
    (1, 40, 0, 48, (6, 10, 10), (1, 1, 1)),
    (1, 32, 0, 64, (8, 16, 16), (1, 1, 1)),
    (2, 16, 0, 80, (3, 8, 8), (1, 1, 1)),
    (1, 16, 0, 16, (70, 8, 8), (1, 1, 1)),                # one chain longer than the 32-block TMEM ring
    (1, 16, 0, 16, (1, 8, 8), (1, 1, 1)),                 # a single plane
    (2, 32, 0, 32, (40, 48, 40), (1, 1, 1)),              # more work than SMs: runs cut mid-chain, ring wrap-around
    (1, 64, 0, 64, (10, 20, 12), (1, 1, 1)),              # wide output: two N tiles of 32 columns (one CTA group each)
    (2, 64, 0, 128, (6, 16, 16), (1, 1, 1)),              # four N tiles
    (1, 24, 0, 96, (5, 9, 11), (1, 1, 1)),                # three N tiles, ragged extents
    (1, 32, 32, 64, (4, 16, 8), (1, 1, 1)),               # virtual concat into two N tiles
]


@pytest.mark.parametrize('case', ZS_CASES, ids=[str(c) for c in ZS_CASES])
def test_conv_zstacked_forward_bias_relu_stats(eng, case):
    N, C0, C1, Co, sp, pad = case
    k = (3, 3, 3)
    assert eng.conv_variant(C0, C1, eng.cpad16(Co), k) == 1
    x0 = dyadic((N, C0) + sp, 31, scale=4, lo=-4, hi=5)
    x1 = dyadic((N, C1) + sp, 32, scale=4, lo=-4, hi=5) if C1 else None
    w = dyadic((Co, C0 + C1) + k, 33, scale=4, lo=-2, hi=3)
    b = dyadic((Co,), 34)
    xin = x0 if x1 is None else torch.cat((x0, x1), 1)
    ref = F.conv3d(xin.double(), w.double(), b.double(), padding=pad)
    wpk = eng.pack_weights(4, w, None, C0, C1, Co, k)
    src1 = qp(eng, x1) if C1 else None
    y, _, stats = eng.conv_forward(qp(eng, x0), wpk, eng.cpad16(Co), Co, k, pad, src1=src1, bias=b, stats_channels=Co, variant=1)
    torch.cuda.synchronize()
    got = from_qp_ref(y.t, Co)
    assert got.shape == ref.shape
    assert_close(got, ref, 1e-6, 'conv')
    Cp = (Co + 7) & ~7
    if Cp != Co:
        full = y.t.permute(0, 1, 5, 2, 3, 4).reshape(N, Cp, *ref.shape[2:])
        assert full[:, Co:].abs().max().item() == 0.0
    assert_close(stats[:, :, 0], ref.sum(dim=(2, 3, 4)), 1e-6, 'sum')
    assert_close(stats[:, :, 1], (ref * ref).sum(dim=(2, 3, 4)), 1e-6, 'sumsq')
    # the halo-tile kernel computes the same function
    y0, _, _ = eng.conv_forward(qp(eng, x0), eng.pack_weights(0, w, None, C0, C1, Co, k), eng.cpad16(Co), Co, k, pad, src1=src1,
                                bias=b)
    assert torch.equal(y0.t, y.t)
    yh, _, _ = eng.conv_forward(qp(eng, x0), wpk, eng.cpad16(Co), Co, k, pad, src1=src1, bias=b, relu=True, half_out=True,
                                variant=1)
    assert yh.half and yh.t.shape[1] == eng.cpad16(Co) // 8
    assert_close(from_qh_ref(yh, Co), ref.clamp_min(0), 1e-3, 'conv+relu -> QH')
    if eng.cpad16(Co) != Co:
        assert from_qh_ref(yh, eng.cpad16(Co))[:, Co:].abs().max().item() == 0.0


@pytest.mark.parametrize('case', [(1, 16, 0, 16, (4, 16, 8), (1, 1, 1)), (2, 32, 0, 64, (5, 10, 12), (1, 1, 1)),
                                  (1, 8, 0, 8, (8, 12, 12), (0, 0, 0)), (1, 16, 24, 32, (4, 16, 8), (1, 1, 1)),
                                  (1, 32, 32, 32, (20, 24, 16), (1, 1, 1)),
                                  (1, 64, 0, 64, (6, 12, 16), (1, 1, 1)),          # N tiles: two of 32 columns
                                  (1, 64, 64, 64, (5, 16, 8), (1, 1, 1)),          # concat dgrad: four N tiles, two per destination
                                  (1, 24, 40, 32, (4, 8, 8), (1, 1, 1))])          # destinations that end inside an N tile
def test_conv_zstacked_dgrad(eng, case):
    N, C0, C1, Co, sp, pad = case
    k = (3, 3, 3)
    nt = eng.cpad16(eng.cpad8(C0) + (eng.cpad8(C1) if C1 else 0))
    assert eng.conv_variant(Co, 0, nt, k) == 1
    x = dyadic((N, C0 + C1) + sp, 35, scale=4, lo=-4, hi=5).double().requires_grad_(True)
    w = dyadic((Co, C0 + C1) + k, 36, scale=4, lo=-2, hi=3)
    y = F.conv3d(x, w.double(), None, padding=pad)
    dy = dyadic(tuple(y.shape), 37, scale=4, lo=-4, hi=5)
    y.backward(dy.double())
    wpk = eng.pack_weights(5, w, None, C0, C1, Co, k)
    dpad = tuple(2 - pp for pp in pad)
    d0, d1, _ = eng.conv_forward(qp(eng, dy), wpk, nt, C0, k, dpad, dst1_C=C1, variant=1)
    assert_close(from_qp_ref(d0.t, C0), x.grad[:, :C0], 1e-6, 'dgrad 0')
    if C1:
        assert_close(from_qp_ref(d1.t, C1), x.grad[:, C0:], 1e-6, 'dgrad 1')


@pytest.mark.parametrize('tz', [1, 2, 3, 4])
def test_conv_forward_tile_depths(eng, tz):
    N, C0, Co, sp, k, pad = 1, 16, 32, (7, 20, 12), (3, 3, 3), (1, 1, 1)
    x = dyadic((N, C0) + sp, 5, scale=4, lo=-4, hi=5)
    w = dyadic((Co, C0) + k, 6, scale=4, lo=-2, hi=3)
    wpk = eng.pack_weights(0, w, None, C0, 0, Co, k)
    ref = F.conv3d(x.double(), w.double(), None, padding=pad)
    y, _, _ = eng.conv_forward(qp(eng, x), wpk, eng.cpad16(Co), Co, k, pad, force_tz=tz)
    assert_close(from_qp_ref(y.t, Co), ref, 1e-6, f'conv tz={tz}')


def test_conv_many_tiles_persistent(eng):
    """more tiles than SMs: exercises the persistent loop, smem ring wrap-around and TMEM ping-pong"""
    N, C0, Co, sp = 2, 32, 32, (40, 48, 40)
    x = dyadic((N, C0) + sp, 7, scale=4, lo=-4, hi=5)
    w = dyadic((Co, C0, 3, 3, 3), 8, scale=4, lo=-2, hi=3)
    wpk = eng.pack_weights(0, w, None, C0, 0, Co, (3, 3, 3))
    ref = F.conv3d(x.double(), w.double(), None, padding=1)
    y, _, _ = eng.conv_forward(qp(eng, x), wpk, eng.cpad16(Co), Co, (3, 3, 3), (1, 1, 1))
    assert_close(from_qp_ref(y.t, Co), ref, 1e-6, 'conv persistent')


@pytest.mark.parametrize('crop', [(0, 0, 0), (1, 2, 3)])
def test_conv_virtual_concat(eng, crop):
    """torch.cat((updec, enc), 1) -> conv (unet.py:399-402) with the skip tensor centre-cropped (autocrop)"""
    N, C0, C1, Co, sp = 1, 16, 16, 16, (6, 16, 16)
    x0 = dyadic((N, C0) + sp, 9, scale=4, lo=-4, hi=5)
    big = tuple(s + 2 * c for s, c in zip(sp, crop))
    x1 = dyadic((N, C1) + big, 10, scale=4, lo=-4, hi=5)
    w = dyadic((Co, C0 + C1, 3, 3, 3), 11, scale=4, lo=-2, hi=3)
    x1c = x1[:, :, crop[0]:crop[0] + sp[0], crop[1]:crop[1] + sp[1], crop[2]:crop[2] + sp[2]]
    ref = F.conv3d(torch.cat((x0, x1c), 1).double(), w.double(), None, padding=1)
    wpk = eng.pack_weights(0, w, None, C0, C1, Co, (3, 3, 3))
    y, _, _ = eng.conv_forward(qp(eng, x0), wpk, eng.cpad16(Co), Co, (3, 3, 3), (1, 1, 1), src1=qp(eng, x1), off1=crop)
    assert_close(from_qp_ref(y.t, Co), ref, 1e-6, 'concat conv')


@pytest.mark.parametrize('case', [(1, 16, 8, (4, 8, 8), (2, 2, 2), None), (2, 32, 16, (3, 8, 16), (1, 2, 2), None),
                                  (1, 64, 32, (3, 5, 6), (2, 2, 2), (5, 9, 12)), (1, 128, 64, (4, 8, 8), (2, 2, 2), None),
                                  (1, 8, 3, (2, 8, 8), (2, 2, 2), None)])
def test_transposed_conv_scatter(eng, case):
    """ConvTranspose k=s (unet.py:152-165) incl. the autocrop of from_up (unet.py:294-301)"""
    N, Ci, Co, sp, s, crop_to = case
    x = dyadic((N, Ci) + sp, 12, scale=4, lo=-4, hi=5)
    w = dyadic((Ci, Co) + s, 13, scale=4, lo=-2, hi=3)
    b = dyadic((Co,), 14)
    ref = F.conv_transpose3d(x.double(), w.double(), b.double(), stride=s)
    out_sp = tuple(ref.shape[2:]) if crop_to is None else crop_to
    ref = ref[:, :, :out_sp[0], :out_sp[1], :out_sp[2]]
    wpk = eng.pack_weights(2, w, None, Ci, 0, Co, s)
    taps = s[0] * s[1] * s[2]
    y, _, stats = eng.conv_forward(qp(eng, x), wpk, taps * eng.cpad16(Co), Co, (1, 1, 1), (0, 0, 0), bias=b,
                                   stats_channels=Co, scatter=s, out_spatial=out_sp)
    assert_close(from_qp_ref(y.t, Co), ref, 1e-6, 'convT')
    assert_close(stats[:, :, 0], ref.sum(dim=(2, 3, 4)), 1e-6, 'convT sum')
    assert_close(stats[:, :, 1], (ref * ref).sum(dim=(2, 3, 4)), 1e-6, 'convT sumsq')


@pytest.mark.parametrize('case', [(1, 16, 16, (4, 16, 8), (3, 3, 3), (1, 1, 1)), (2, 32, 64, (5, 10, 12), (3, 3, 3), (1, 1, 1)),
                                  (1, 8, 24, (3, 16, 16), (1, 3, 3), (0, 1, 1)), (1, 8, 8, (8, 12, 12), (3, 3, 3), (0, 0, 0)),
                                  (1, 128, 128, (4, 8, 8), (3, 3, 3), (1, 1, 1))])
def test_conv_dgrad(eng, case):
    N, C0, Co, sp, k, pad = case
    x = dyadic((N, C0) + sp, 15, scale=4, lo=-4, hi=5).double().requires_grad_(True)
    w = dyadic((Co, C0) + k, 16, scale=4, lo=-2, hi=3)
    y = F.conv3d(x, w.double(), None, padding=pad)
    dy = dyadic(tuple(y.shape), 17, scale=4, lo=-4, hi=5)
    y.backward(dy.double())
    wpk = eng.pack_weights(1, w, None, C0, 0, Co, k)
    dpad = tuple(kk - 1 - pp for kk, pp in zip(k, pad))
    dx, _, _ = eng.conv_forward(qp(eng, dy), wpk, eng.cpad16(eng.cpad8(C0)), C0, k, dpad)
    assert_close(from_qp_ref(dx.t, C0), x.grad, 1e-6, 'dgrad')


def test_conv_dgrad_concat_split(eng):
    N, C0, C1, Co, sp = 1, 16, 24, 32, (4, 16, 8)
    w = dyadic((Co, C0 + C1, 3, 3, 3), 18, scale=4, lo=-2, hi=3)
    x = dyadic((N, C0 + C1) + sp, 19, scale=4, lo=-4, hi=5).double().requires_grad_(True)
    y = F.conv3d(x, w.double(), None, padding=1)
    dy = dyadic(tuple(y.shape), 20, scale=4, lo=-4, hi=5)
    y.backward(dy.double())
    wpk = eng.pack_weights(1, w, None, C0, C1, Co, (3, 3, 3))
    d0, d1, _ = eng.conv_forward(qp(eng, dy), wpk, eng.cpad16(eng.cpad8(C0) + eng.cpad8(C1)), C0, (3, 3, 3), (1, 1, 1),
                                 dst1_C=C1)
    assert_close(from_qp_ref(d0.t, C0), x.grad[:, :C0], 1e-6, 'dgrad split 0')
    assert_close(from_qp_ref(d1.t, C1), x.grad[:, C0:], 1e-6, 'dgrad split 1')


WGRAD_CASES = [
    (1, 8, 0, 16, (4, 16, 8), (3, 3, 3), (1, 1, 1)),
    (2, 32, 0, 32, (6, 20, 18), (3, 3, 3), (1, 1, 1)),
    (1, 1, 0, 32, (8, 16, 16), (3, 3, 3), (1, 1, 1)),
    (1, 64, 0, 64, (4, 8, 8), (3, 3, 3), (1, 1, 1)),
    (1, 16, 16, 16, (4, 16, 8), (3, 3, 3), (1, 1, 1)),     # virtual concat
    (1, 8, 0, 24, (3, 16, 16), (1, 3, 3), (0, 1, 1)),      # planar
    (1, 8, 0, 8, (8, 12, 12), (3, 3, 3), (0, 0, 0)),       # VALID
    (1, 160, 0, 16, (3, 8, 8), (3, 3, 3), (1, 1, 1)),      # > 128 input channels: two M chunks
    (1, 3, 0, 5, (5, 7, 9), (3, 3, 3), (1, 1, 1)),
    (2, 32, 32, 32, (5, 40, 70), (3, 3, 3), (1, 1, 1)),    # virtual concat at the width of the BASELINE layers, ragged tiles
    (1, 16, 0, 64, (4, 12, 36), (3, 3, 3), (1, 1, 1)),     # two N chunks
    (1, 64, 0, 128, (2, 20, 40), (1, 3, 3), (0, 1, 1)),    # planar, 128 output channels
    (1, 24, 0, 40, (6, 9, 11), (3, 3, 3), (0, 0, 0)),      # VALID, ragged channel counts
    (1, 32, 0, 16, (1, 33, 65), (1, 3, 3), (0, 1, 1)),     # D = 1 (the 2D path)
    (1, 256, 0, 32, (2, 6, 6), (3, 3, 3), (1, 1, 1)),      # 8 M chunks, tiny extent
    (1, 8, 0, 8, (8, 8, 8), (3, 3, 3), (0, 0, 0)),         # VALID at the bottom of a small net (merge_mode='add' golden)
    (1, 8, 0, 16, (8, 8, 8), (3, 3, 3), (0, 0, 0)),
    (1, 8, 0, 8, (6, 6, 6), (3, 3, 3), (0, 0, 0)),
]


@pytest.mark.parametrize('case', WGRAD_CASES, ids=[str(c) for c in WGRAD_CASES])
def test_conv_wgrad(eng, case):
    N, C0, C1, Co, sp, k, pad = case
    x = dyadic((N, C0 + C1) + sp, 21, scale=2, lo=-2, hi=3)
    w = torch.zeros((Co, C0 + C1) + k, dtype=torch.float64, device='cuda', requires_grad=True)
    y = F.conv3d(x.double(), w, None, padding=pad)
    dy = dyadic(tuple(y.shape), 22, scale=2, lo=-2, hi=3)
    y.backward(dy.double())
    src0 = qp(eng, x[:, :C0].contiguous())
    src1 = qp(eng, x[:, C0:].contiguous()) if C1 else None
    dw = eng.wgrad(src0, qp(eng, dy), Co, k, pad, tuple(w.shape), src1=src1)
    assert_close(dw, w.grad, 1e-6, 'wgrad')


def test_conv_wgrad_cropped_second_source(eng):
    """the skip tensor of a VALID network enters the conv as a centre-cropped view (autocrop, unet.py:303-324); any voxel
    offset is a legal start (16-byte units), including x offsets that are not multiples of 8"""
    N, C0, C1, Co, sp, off = 1, 16, 16, 32, (4, 10, 12), (2, 3, 5)
    big = tuple(s + 2 * o for s, o in zip(sp, off))
    x0 = dyadic((N, C0) + sp, 41, scale=2, lo=-2, hi=3)
    x1 = dyadic((N, C1) + big, 42, scale=2, lo=-2, hi=3)
    crop = x1[:, :, off[0]:off[0] + sp[0], off[1]:off[1] + sp[1], off[2]:off[2] + sp[2]]
    w = torch.zeros((Co, C0 + C1, 3, 3, 3), dtype=torch.float64, device='cuda', requires_grad=True)
    y = F.conv3d(torch.cat((x0, crop), 1).double(), w, None)
    dy = dyadic(tuple(y.shape), 43, scale=2, lo=-2, hi=3)
    y.backward(dy.double())
    dw = eng.wgrad(qp(eng, x0), qp(eng, dy), Co, (3, 3, 3), (0, 0, 0), tuple(w.shape), src1=qp(eng, x1), off1=off)
    assert_close(dw, w.grad, 1e-6, 'wgrad cropped concat')


@pytest.mark.parametrize('case', [(1, 16, 8, (4, 8, 8), (2, 2, 2)), (2, 32, 16, (3, 8, 16), (1, 2, 2)),
                                  (1, 64, 32, (2, 8, 8), (2, 2, 2))])
def test_transposed_conv_backward(eng, case):
    """dx and dW of ConvTranspose k=s via the space-to-depth gradient"""
    N, Ci, Co, sp, s = case
    x = dyadic((N, Ci) + sp, 23, scale=2, lo=-2, hi=3).double().requires_grad_(True)
    w = dyadic((Ci, Co) + s, 24, scale=2, lo=-2, hi=3).double().requires_grad_(True)
    y = F.conv_transpose3d(x, w, None, stride=s)
    dy = dyadic(tuple(y.shape), 25, scale=2, lo=-2, hi=3)
    y.backward(dy.double())
    # space-to-depth of dy: channel = tap * pad8(Co) + co on the coarse grid
    taps = s[0] * s[1] * s[2]
    Cp = eng.cpad8(Co)
    D, H, W = sp
    d = dy.view(N, Co, D, s[0], H, s[1], W, s[2]).permute(0, 3, 5, 7, 1, 2, 4, 6).reshape(N, taps, Co, D, H, W)
    dpad = torch.zeros((N, taps, Cp, D, H, W), device='cuda')
    dpad[:, :, :Co] = d
    dyq = qp(eng, dpad.view(N, taps * Cp, D, H, W))
    wpk = eng.pack_weights(3, w.detach().float(), None, Ci, 0, Co, s)
    dx, _, _ = eng.conv_forward(dyq, wpk, eng.cpad16(Ci), Ci, (1, 1, 1), (0, 0, 0))
    assert_close(from_qp_ref(dx.t, Ci), x.grad, 1e-6, 'convT dgrad')
    dw = eng.wgrad(qp(eng, x.detach().float()), dyq, taps * Cp, (1, 1, 1), (0, 0, 0), tuple(w.shape), layout=1,
                   up_taps=taps, up_co=Co)
    assert_close(dw, w.grad, 1e-6, 'convT wgrad')


# ---------------------------------------------------------------------------------------- norm / act / pool
def _norm_ref(y, mode, G, gamma, beta, rm, rv, eps=1e-5):
    if mode == 1:
        return F.group_norm(y, G, gamma, beta, eps)
    if mode == 2:
        return F.batch_norm(y, rm, rv, gamma, beta, True, 0.1, eps)
    return y


@pytest.mark.parametrize('mode,G,C,sp,pool', [(1, 8, 32, (6, 8, 10), None), (1, 8, 16, (5, 7, 9), (2, 2, 2)),
                                              (2, 1, 8, (4, 6, 8), (2, 2, 2)), (2, 1, 24, (3, 9, 8), (1, 2, 2)),
                                              (0, 1, 8, (4, 6, 8), (2, 2, 2)), (1, 3, 3, (4, 4, 4), None),
                                              (1, 4, 8, (2, 3, 136), None), (2, 1, 8, (2, 4, 132), (1, 2, 2)),
                                              (1, 2, 12, (4, 5, 6), None), (1, 4, 64, (4, 8, 8), None)])
@pytest.mark.parametrize('path', ['fused', 'split', 'fused-rescale'])
def test_norm_act_pool_forward_backward(eng, monkeypatch, mode, G, C, sp, pool, path):
    """`path`: the one-kernel backward (reduce -> grid barrier -> apply per sample) or the three-kernel one; 'fused-rescale'
    makes the second sample's gradient 2^12 times larger, so the per-tensor fp16 scale chosen after sample 0 has to be
    lowered and sample 0 re-scaled in place."""
    monkeypatch.setenv('E3B_NORM_BWD', 'split' if path == 'split' else 'fused')
    N = 3 if path == 'fused-rescale' else 2
    rs = np.random.RandomState(31)
    y = torch.from_numpy(rs.standard_normal((N, C) + sp).astype(np.float32)).cuda()
    gamma = torch.from_numpy((1 + 0.2 * rs.standard_normal(C)).astype(np.float32)).cuda()
    beta = torch.from_numpy((0.1 * rs.standard_normal(C)).astype(np.float32)).cuda()
    rm = torch.zeros(C, device='cuda')
    rv = torch.ones(C, device='cuda')
    S = sp[0] * sp[1] * sp[2]
    yd = y.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rmd, rvd = rm.double().clone(), rv.double().clone()
    a_ref = F.relu(_norm_ref(yd, mode, G, gd if mode else None, bd if mode else None, rmd, rvd))
    outs = [a_ref]
    if pool is not None:
        outs.append(F.max_pool3d(a_ref, pool, pool, ceil_mode=True))
    # forward through the kernels
    yq = qp32(eng, y)
    stats = torch.stack((y.double().sum(dim=(2, 3, 4)), (y.double() ** 2).sum(dim=(2, 3, 4))), dim=-1).contiguous()
    if mode == 0:
        a, pooled = eng.norm_act(yq, None, None, pool=pool, save=True)
        nstate = None
    else:
        nstate = eng.norm_finalize(stats, mode, G, N, C, S, gamma, beta, 1e-5, rm if mode == 2 else None,
                                   rv if mode == 2 else None, 0.1, y.device)
        a, pooled = eng.norm_act(yq, nstate.scale, nstate.shift, pool=pool, save=True)
    # activations / gradients that feed an MMA are stored rounded to TF32 (2^-11 relative)
    assert a.half and a.t.shape[1] == eng.cpad16(C) // 8
    assert_close(from_qh_ref(a, C), a_ref, 6e-4, 'norm+relu')
    if eng.cpad16(C) != C:
        assert from_qh_ref(a, eng.cpad16(C))[:, C:].abs().max().item() == 0.0
    if pool is not None:
        assert_close(from_qh_ref(pooled, C), outs[1], 6e-4, 'pool')
    if mode == 2:
        assert_close(rm, rmd, 1e-5, 'running_mean')
        assert_close(rv, rvd, 1e-5, 'running_var')
    # backward: g0 on a, gp on pooled
    g0 = torch.from_numpy(rs.standard_normal(tuple(a_ref.shape)).astype(np.float32)).cuda()
    if path == 'fused-rescale':
        g0[1] *= 4096.0
    loss = (a_ref * g0.double()).sum()
    gpq = None
    if pool is not None:
        gp = torch.from_numpy(rs.standard_normal(tuple(outs[1].shape)).astype(np.float32)).cuda()
        if path == 'fused-rescale':
            gp[1] *= 4096.0
        loss = loss + (outs[1] * gp.double()).sum()
        gpq = qp32(eng, gp)
    loss.backward()

    class Spec:
        pass
    u = eng.Unit()
    u.spec = Spec()
    u.spec.norm = torch.nn.GroupNorm(1, 1) if mode else None
    if mode:
        u.spec.norm.weight = torch.nn.Parameter(gamma)
        u.spec.norm.eps = 1e-5
    u.a, u.y, u.pool, u.mode, u.G, u.nstate, u.stats = a, yq, pool, mode, G, nstate, (stats if mode else None)
    u.pooled = pooled
    dy, dgamma, dbeta, dbias = eng._norm_bwd(u, C, qp32(eng, g0), gp=gpq)
    assert dy.half and dy.scale is not None
    if path == 'fused-rescale' and mode != 2:
        # per sample: the small samples sit 12 bits below the large one in the shared fp16 scale (fp16 subnormal steps)
        for n in range(N):
            assert_close(from_qh_ref(dy, C)[n], yd.grad[n], 1e-3 if n == 1 else 2e-2, 'norm bwd dy sample %d' % n)
    else:
        assert_close(from_qh_ref(dy, C), yd.grad, 1e-3, 'norm bwd dy')
    # the fp16 scale is a power of two that brings the largest |dy| into (2^10, 2^14]
    k = torch.log2(dy.scale[1]).item()
    assert k == round(k) and abs(dy.scale[1].item() * dy.scale[2].item() - 1.0) < 1e-6
    assert 2.0 ** 10 < dy.t.float().abs().max().item() <= 2.0 ** 14
    if mode:
        assert_close(dgamma, gd.grad, 2e-4, 'dgamma')
        assert_close(dbeta, bd.grad, 2e-4, 'dbeta')
    ref_dbias = yd.grad.sum(dim=(0, 2, 3, 4))
    scale = max(yd.grad.abs().sum(dim=(0, 2, 3, 4)).max().item(), 1e-6)
    assert (dbias.double() - ref_dbias).abs().max().item() / scale < 1e-4


@pytest.mark.parametrize('pool,direct,sp', [(None, True, (4, 6, 8)), ((2, 2, 2), True, (4, 6, 8)), ((2, 2, 2), False, (4, 6, 8)),
                                            # skip + pooled gradient on items made of whole pooling windows (variant 5: the
                                            # pooled gradient and the slots are staged with the item): rows of one plane
                                            # (H W = 512 and 1024), whole planes (H W = 128), planar pooling
                                            ((2, 2, 2), False, (4, 16, 32)), ((2, 2, 2), False, (2, 32, 32)),
                                            ((2, 2, 2), False, (8, 8, 16)), ((1, 2, 2), False, (3, 16, 32)),
                                            ((2, 2, 2), False, (4, 64, 64))])
def test_norm_backward_direct_plus_skip_gradient(eng, pool, direct, sp):
    """two un-cropped gradients on the same activation (g0 + g1: the fused kernel's variant 1; with a pooled gradient on
    top: the run-time variant; skip + pooled gradient without a direct one: variant 3 / 5, the encoder's last conv)"""
    N, C, G = 2, 16, 4
    rs = np.random.RandomState(77)
    y = torch.from_numpy(rs.standard_normal((N, C) + sp).astype(np.float32)).cuda()
    gamma = torch.from_numpy((1 + 0.2 * rs.standard_normal(C)).astype(np.float32)).cuda()
    beta = torch.from_numpy((0.1 * rs.standard_normal(C)).astype(np.float32)).cuda()
    yd = y.double().requires_grad_(True)
    a_ref = F.relu(F.group_norm(yd, G, gamma.double(), beta.double(), 1e-5))
    g0 = torch.from_numpy(rs.standard_normal((N, C) + sp).astype(np.float32)).cuda()
    g1 = torch.from_numpy(rs.standard_normal((N, C) + sp).astype(np.float32)).cuda()
    if not direct:
        g0.zero_()
    loss = (a_ref * (g0 + g1).double()).sum()
    yq = qp32(eng, y)
    S = sp[0] * sp[1] * sp[2]
    stats = torch.stack((y.double().sum(dim=(2, 3, 4)), (y.double() ** 2).sum(dim=(2, 3, 4))), dim=-1).contiguous()
    nstate = eng.norm_finalize(stats, 1, G, N, C, S, gamma, beta, 1e-5, None, None, 0.1, y.device)
    a, pooled = eng.norm_act(yq, nstate.scale, nstate.shift, pool=pool, save=True)
    gpq = None
    if pool is not None:
        # un-pool along the arg-max of the activations the kernels stored (rounded to 10 mantissa bits: a near-tie of the
        # fp64 reference may resolve differently, which is not what this test is about)
        a_k = from_qh_ref(a, C).double()
        p_k, idx = F.max_pool3d(a_k, pool, pool, ceil_mode=True, return_indices=True)
        gp = torch.from_numpy(rs.standard_normal(tuple(p_k.shape)).astype(np.float32)).cuda()
        loss = loss + (a_ref * F.max_unpool3d(gp.double(), idx, pool, pool, output_size=sp)).sum()
        gpq = qp32(eng, gp)
    loss.backward()

    class Spec:
        pass
    u = eng.Unit()
    u.spec = Spec()
    u.spec.norm = torch.nn.GroupNorm(G, C)
    u.spec.norm.weight = torch.nn.Parameter(gamma)
    u.a, u.y, u.pool, u.mode, u.G, u.nstate, u.stats, u.pooled = a, yq, pool, 1, G, nstate, stats, pooled
    dy, dgamma, dbeta, dbias = eng._norm_bwd(u, C, qp32(eng, g0) if direct else None, g1=qp32(eng, g1), gp=gpq)
    assert_close(from_qh_ref(dy, C), yd.grad, 1e-3, 'dy')


@pytest.mark.parametrize('fine', [(5, 7, 8), (4, 6, 8)])
def test_norm_backward_space_to_depth(eng, fine):
    """norm0/act0 backward of UpConv written tap-major for the transposed conv's dgrad GEMM; (5, 7, 8): cropped fine
    grid (three-kernel path), (4, 6, 8): the one-kernel path"""
    N, C, s = 2, 8, (2, 2, 2)
    rs = np.random.RandomState(5)
    y = torch.from_numpy(rs.standard_normal((N, C) + fine).astype(np.float32)).cuda()
    g0 = torch.from_numpy(rs.standard_normal((N, C) + fine).astype(np.float32)).cuda()
    yd = y.double().requires_grad_(True)
    a_ref = F.relu(yd)
    (a_ref * g0.double()).sum().backward()
    yq = qp32(eng, y)
    a, _ = eng.norm_act(yq, None, None)
    u = eng.Unit()

    class Spec:
        norm = None
    u.spec = Spec()
    u.a, u.y, u.pool, u.mode, u.G, u.nstate, u.stats = a, yq, None, 0, 1, None, None
    dy, _, _, dbias = eng._norm_bwd(u, C, qp32(eng, g0), s2d=s)
    coarse = tuple(-(-f // k) for f, k in zip(fine, s))
    full = torch.zeros((N, C) + tuple(c * k for c, k in zip(coarse, s)), dtype=torch.float64, device='cuda')
    full[:, :, :fine[0], :fine[1], :fine[2]] = yd.grad
    D, H, W = coarse
    ref = full.view(N, C, D, 2, H, 2, W, 2).permute(0, 3, 5, 7, 1, 2, 4, 6).reshape(N, 8 * C, D, H, W)
    assert_close(from_qh_ref(dy, 8 * C), ref, 6e-4, 's2d')


# ---------------------------------------------------------------------------------------- head
@pytest.mark.parametrize('C,Co', [(32, 2), (8, 4), (64, 3)])
def test_head_modes_and_backward(eng, C, Co):
    N, sp = 2, (5, 6, 7)
    rs = np.random.RandomState(41)
    a = torch.from_numpy(rs.standard_normal((N, C) + sp).astype(np.float32)).cuda().half().float()   # features are fp16
    conv = torch.nn.Conv3d(C, Co, 1).cuda()
    ad = a.double().requires_grad_(True)
    wd, bd = conv.weight.detach().double().requires_grad_(True), conv.bias.detach().double().requires_grad_(True)
    ref = F.conv3d(ad, wd, bd)
    aq = qp(eng, a)
    assert_close(eng.head(aq, conv, 0), ref, 1e-5, 'logits')
    assert_close(eng.head(aq, conv, 1), ref.softmax(1), 1e-5, 'softmax')
    am = eng.head(aq, conv, 2)
    assert am.dtype == torch.uint8 and am.shape == (N, 1) + sp
    top2 = ref.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 1e-4
    assert torch.equal(am[:, 0][decided].long(), ref.argmax(1)[decided])
    # crop-and-place (inference.py:147-151,188-197)
    dst = torch.full((1, Co, 9, 9, 9), -1.0, device='cuda')
    org = torch.tensor([[0, 0, 0], [3, 4, 1]], dtype=torch.int32, device='cuda')
    eng.head(aq, conv, 0, dst=dst, crop=((1, 1, 2), (3, 4, 4)), dst_origin=org, dst_single=True)
    assert_close(dst[0, :, 3:6, 4:8, 1:5], ref[1, :, 1:4, 1:5, 2:6], 1e-5, 'placed tile 1')
    assert_close(dst[0, :, 0:3, 0:4, 0:4], ref[0, :, 1:4, 1:5, 2:6], 1e-5, 'placed tile 0')
    assert dst[0, :, 6:, :, :].eq(-1).all()
    # backward
    dl = torch.from_numpy(rs.standard_normal(tuple(ref.shape)).astype(np.float32)).cuda()
    ref.backward(dl.double())
    da = eng.QP.empty(N, C, *sp, a.device)
    dw = torch.empty((Co, C), device='cuda')
    db = torch.empty((Co,), device='cuda')
    ws = torch.empty((Co * eng.cpad8(C) + Co,), dtype=torch.float64, device='cuda')
    from elektronn3_b200 import _lib as L
    L.check(L.lib().e3b_head_bwd(dl.data_ptr(), aq.ptr, conv.weight.detach().data_ptr(), da.ptr, dw.data_ptr(),
                                 db.data_ptr(), ws.data_ptr(), N, C, Co, *sp, torch.cuda.current_stream().cuda_stream))
    assert_close(from_qp_ref(da.t, C), ad.grad, 1e-5, 'head da')
    assert_close(dw, wd.grad.view(Co, C), 1e-5, 'head dw')
    assert_close(db, bd.grad, 1e-5, 'head db')


def test_gather_tiles_zero_padding(eng):
    vol = dyadic((2, 6, 7, 8), 51)
    org = torch.tensor([[-2, -1, -3], [3, 4, 5], [0, 0, 0]], dtype=torch.int32, device='cuda')
    q = eng.gather_tiles(vol, org, 3, 2, (5, 6, 7))
    pad = F.pad(vol, (8, 8, 8, 8, 8, 8))
    for b, (z, y, x) in enumerate(org.tolist()):
        ref = pad[:, z + 8:z + 13, y + 8:y + 14, x + 8:x + 15]
        assert torch.equal(from_qh_ref(eng.QP(q.t[b:b + 1], 1, 2, 5, 6, 7), 2)[0], ref)


def test_errors_surface_as_exceptions(eng):
    x = dyadic((1, 8, 4, 8, 8), 61)
    w = dyadic((8, 8, 5, 5, 5), 62)
    with pytest.raises(RuntimeError):
        eng.conv_forward(qp(eng, x), w.flatten(), 16, 8, (5, 5, 5), (2, 2, 2))


# ---------------------------------------------------------------------------------------- options
_ACTS = {'leaky': ((1, 0.1), lambda t: F.leaky_relu(t, 0.1)), 'lin': ((0, 0.0), lambda t: t), 'prelu': (None, None),
         'rrelu_eval': ((1, (1 / 8 + 1 / 3) / 2), lambda t: F.rrelu(t, training=False)), 'silu': ((2, 0.0), F.silu)}


@pytest.mark.parametrize('act', list(_ACTS))
@pytest.mark.parametrize('mode,pool,path', [(1, None, 'fused'), (1, (2, 2, 2), 'fused'), (1, (2, 2, 2), 'split'),
                                            (2, None, 'fused'), (0, (1, 2, 2), 'split')])
def test_activations_forward_backward(eng, monkeypatch, act, mode, pool, path):
    """get_activation's other choices (models/unet.py:183-199) through norm_act and both norm-backward paths (SiLU always
    takes the three-kernel one) against float64 autograd"""
    monkeypatch.setenv('E3B_NORM_BWD', 'split' if path == 'split' else 'fused')
    code, fn = _ACTS[act]
    slope = sd = None
    if act == 'prelu':                    # the learned slope of nn.PReLU lives in device memory; its gradient is a third sum
        slope = torch.nn.Parameter(torch.tensor([0.3], device='cuda'))
        sd = slope.detach().double().requires_grad_(True)
        code, fn = (1, 0.0, slope), (lambda t: F.prelu(t, sd))
    N, C, G, sp = 2, 16, 4, (4, 6, 8)
    rs = np.random.RandomState(5)
    y = torch.from_numpy(rs.standard_normal((N, C) + sp).astype(np.float32)).cuda()
    gamma = torch.from_numpy((1 + 0.2 * rs.standard_normal(C)).astype(np.float32)).cuda()
    beta = torch.from_numpy((0.1 * rs.standard_normal(C)).astype(np.float32)).cuda()
    rm, rv = torch.zeros(C, device='cuda'), torch.ones(C, device='cuda')
    yd = y.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    a_ref = fn(_norm_ref(yd, mode, G, gd if mode else None, bd if mode else None, rm.double(), rv.double()))
    p_ref = F.max_pool3d(a_ref, pool, pool, ceil_mode=True) if pool is not None else None
    yq = qp32(eng, y)
    S = sp[0] * sp[1] * sp[2]
    stats = torch.stack((y.double().sum(dim=(2, 3, 4)), (y.double() ** 2).sum(dim=(2, 3, 4))), dim=-1).contiguous()
    nstate = None
    if mode:
        nstate = eng.norm_finalize(stats, mode, G, N, C, S, gamma, beta, 1e-5, rm if mode == 2 else None,
                                   rv if mode == 2 else None, 0.1, y.device)
    a, pooled = eng.norm_act(yq, nstate.scale if mode else None, nstate.shift if mode else None, pool=pool, save=True, act=code)
    assert_close(from_qh_ref(a, C), a_ref, 6e-4, 'norm+act')
    g0 = torch.from_numpy(rs.standard_normal((N, C) + sp).astype(np.float32)).cuda()
    loss = (a_ref * g0.double()).sum()
    gpq = None
    if pool is not None:
        assert_close(from_qh_ref(pooled, C), p_ref, 6e-4, 'pool')
        gp = torch.from_numpy(rs.standard_normal(tuple(p_ref.shape)).astype(np.float32)).cuda()
        loss = loss + (p_ref * gp.double()).sum()
        gpq = qp32(eng, gp)
    loss.backward()

    class Spec:
        pass
    u = eng.Unit()
    u.spec = Spec()
    u.spec.norm = torch.nn.GroupNorm(1, 1) if mode else None
    if mode:
        u.spec.norm.weight = torch.nn.Parameter(gamma)
        u.spec.norm.eps = 1e-5
    u.a, u.y, u.pool, u.mode, u.G, u.nstate, u.stats, u.pooled = a, yq, pool, mode, G, nstate, (stats if mode else None), pooled
    u.act = code
    dy, dgamma, dbeta, dbias = eng._norm_bwd(u, C, qp32(eng, g0), gp=gpq)
    assert_close(from_qh_ref(dy, C), yd.grad, 1e-3, 'norm bwd dy (%s)' % act)
    if mode:
        assert_close(dgamma, gd.grad, 3e-4, 'dgamma')
        assert_close(dbeta, bd.grad, 3e-4, 'dbeta')
    if act == 'prelu':
        assert_close(u.dslope, sd.grad, 3e-4, 'dslope')
    else:
        assert u.dslope is None


@pytest.mark.parametrize('C,sp,sp1,off', [(16, (4, 6, 8), (4, 6, 8), (0, 0, 0)), (8, (3, 5, 7), (7, 9, 9), (2, 2, 1)),
                                          (24, (1, 6, 10), (1, 8, 12), (0, 1, 1))])
def test_add_qh_with_centre_crop(eng, C, sp, sp1, off):
    """merge_mode='add' (models/unet.py:399-401) after autocrop's centre crop of the skip tensor (:303-324)"""
    a = dyadic((2, C) + sp, 1)
    b = dyadic((2, C) + sp1, 2)
    out = eng.add_qh(qp(eng, a), qp(eng, b), off)
    ref = a + b[:, :, off[0]:off[0] + sp[0], off[1]:off[1] + sp[1], off[2]:off[2] + sp[2]]
    assert_close(from_qh_ref(out, C), ref, 1e-6, 'add')
    if eng.cpad16(C) != C:
        assert from_qh_ref(out, eng.cpad16(C))[:, C:].abs().max().item() == 0.0


@pytest.mark.parametrize('mode', ['nearest', 'trilinear'])
@pytest.mark.parametrize('C,sp,s,cp,crop', [(16, (3, 4, 5), (2, 2, 2), (1, 1, 1), (0, 0, 0)),
                                            (8, (3, 4, 5), (2, 2, 2), (1, 1, 1), (1, 0, 1)),
                                            (24, (4, 3, 6), (1, 2, 2), (0, 1, 1), (0, 1, 0)),
                                            (8, (2, 5, 4), (2, 2, 2), (0, 0, 0), (1, 1, 0))])
def test_upsample_into_padded_tensor_and_its_transpose(eng, mode, C, sp, s, cp, crop):
    """nn.Upsample of ResizeConv (models/unet.py:411-449) written with the conv's zero padding explicit and autocrop's
    trailing crop folded in; the backward kernel is checked against autograd of F.interpolate + F.pad."""
    import ctypes  # noqa: F401
    from elektronn3_b200 import _lib as L
    N = 2
    x = dyadic((N, C) + sp, 3)
    full = tuple(a * b for a, b in zip(sp, s))
    out_sp = tuple(f - c for f, c in zip(full, crop))                  # extents of the (auto)cropped conv output
    k = tuple(2 * p + 1 for p in cp)
    ext = tuple(o + kk - 1 for o, kk in zip(out_sp, k))
    R = tuple(min(f, e - p) for f, e, p in zip(full, ext, cp))
    xd = x.double().requires_grad_(True)
    up = F.interpolate(xd, scale_factor=tuple(float(v) for v in s), mode=mode) if mode == 'nearest' else \
        F.interpolate(xd, scale_factor=tuple(float(v) for v in s), mode=mode, align_corners=False)
    up = up[:, :, :R[0], :R[1], :R[2]]
    ref = F.pad(up, (cp[2], ext[2] - cp[2] - R[2], cp[1], ext[1] - cp[1] - R[1], cp[0], ext[0] - cp[0] - R[0]))
    src = qp(eng, x)
    dst = eng.QP.empty_half(N, C, ext[0], ext[1], ext[2], x.device)
    geom = tuple(sp) + ext + tuple(s) + tuple(cp) + R + (0 if mode == 'nearest' else 1,)
    L.check(L.lib().e3b_upsample_qh(src.ptr, dst.ptr, N, C, *geom, eng._stream()), 'upsample_qh')
    assert_close(from_qh_ref(dst, C), ref, 1e-3 if mode != 'nearest' else 1e-6, 'upsample')
    g = dyadic((N, C) + ext, 4)
    (ref * g.double()).sum().backward()
    u = eng.Unit()
    u.resize = geom
    got = eng._resize_bwd(u, qp32(eng, g))
    assert_close(from_qp_ref(got.t, C), xd.grad, 1e-6, 'upsample backward')


# ---------------------------------------------------------------------------------------- resunet shortcut
@pytest.mark.parametrize('C,sp,half', [(16, (4, 6, 8), True), (8, (3, 5, 7), False), (20, (2, 9, 11), True), (32, (8, 16, 16), False)])
def test_residual_add_and_statistics(eng, C, sp, half):
    """y += proj(inp) of ConvBlock (models/resunet.py:257-258) with the statistics of the sum; then the gradient accumulation"""
    from elektronn3_b200 import _lib as L
    N = 2
    y = dyadic((N, C) + sp, 11)
    r = dyadic((N, C) + sp, 12)
    yq = qp32(eng, y)
    rq = qp(eng, r) if half else qp32(eng, r)
    stats = torch.empty((N, C, 2), dtype=torch.float64, device='cuda')
    S = sp[0] * sp[1] * sp[2]
    L.check(L.lib().e3b_residual_add(yq.ptr, rq.ptr, 1 if half else 0, stats.data_ptr(), N, C, S, eng._stream()), 'residual_add')
    ref = (y + r).double()
    assert_close(from_qp_ref(yq.t, C), ref, 1e-7, 'residual sum')
    assert_close(stats[..., 0], ref.sum(dim=(2, 3, 4)), 1e-6, 'sum')
    assert_close(stats[..., 1], (ref ** 2).sum(dim=(2, 3, 4)), 1e-6, 'sum of squares')
    if eng.cpad8(C) != C:
        assert from_qp_ref(yq.t, eng.cpad8(C))[:, C:].abs().max().item() == 0.0
    # gradient accumulation: dst += alpha * src
    g = dyadic((N, C) + sp, 13)
    gq = qp32(eng, g)
    alpha = torch.tensor([0.25], device='cuda')
    eng._axpy(gq, rq, alpha.data_ptr())
    assert_close(from_qp_ref(gq.t, C), g + 0.25 * r, 1e-7, 'axpy')
    eng._axpy(gq, rq)
    assert_close(from_qp_ref(gq.t, C), g + 1.25 * r, 1e-7, 'axpy (alpha = 1)')


@pytest.mark.parametrize('C0,C1,Co,sp', [(8, 0, 16, (4, 8, 8)), (16, 16, 16, (4, 8, 16)), (32, 32, 32, (3, 10, 18)), (8, 8, 8, (5, 7, 9))])
def test_projection_shortcut_conv_forward_backward(eng, C0, C1, Co, sp):
    """the 1x1x1 projection of a residual ConvBlock (models/resunet.py:246-250) over the (virtual) concat: forward, dgrad
    into both sources and the weight gradient in torch layout"""
    N = 2
    x = dyadic((N, C0 + C1) + sp, 31, scale=2, lo=-2, hi=3).double().requires_grad_(True)
    w = dyadic((Co, C0 + C1, 1, 1, 1), 32, scale=2, lo=-2, hi=3).double().requires_grad_(True)
    b = dyadic((Co,), 33)
    y = F.conv3d(x, w, b.double())
    dy = dyadic(tuple(y.shape), 34, scale=2, lo=-2, hi=3)
    y.backward(dy.double())
    k, pad = (1, 1, 1), (0, 0, 0)
    src0 = qp(eng, x.detach()[:, :C0].float().contiguous())
    src1 = qp(eng, x.detach()[:, C0:].float().contiguous()) if C1 else None
    wpk = eng.pack_weights(0, w.detach().float(), None, C0, C1, Co, k)
    out, _, _ = eng.conv_forward(src0, wpk, eng.cpad16(Co), Co, k, pad, src1=src1, bias=b)
    assert_close(from_qp_ref(out.t, Co), y, 1e-6, 'projection forward')
    dyq = qp(eng, dy)
    dw = eng.wgrad(src0, dyq, Co, k, pad, tuple(w.shape), src1=src1)
    assert_close(dw, w.grad, 1e-6, 'projection wgrad')
    nd = eng.cpad16(eng.cpad8(C0) + (eng.cpad8(C1) if C1 else 0))
    wpd = eng.pack_weights(1, w.detach().float(), None, C0, C1, Co, k)
    d0, d1, _ = eng.conv_forward(dyq, wpd, nd, C0, k, pad, dst1_C=C1)
    assert_close(from_qp_ref(d0.t, C0), x.grad[:, :C0], 1e-6, 'projection dgrad 0')
    if C1:
        assert_close(from_qp_ref(d1.t, C1), x.grad[:, C0:], 1e-6, 'projection dgrad 1')


def test_weight_scale_table_kernel_matches_the_torch_formula(eng):
    """e3b_weight_scales (one launch for all weight tensors of a training step) == engine.weight_scale_table (torch ops)"""
    from elektronn3_b200 import _lib as L
    torch.manual_seed(3)
    ws = [torch.randn(32, 32, 3, 3, 3, device='cuda') * 0.03, torch.randn(7, 5, 1, 3, 3, device='cuda') * 40.0,
          torch.zeros(4, 4, 1, 1, 1, device='cuda'), torch.full((3, 3, 3), 2.0, device='cuda'),
          torch.randn(128, 128, 3, 3, 3, device='cuda') * 1e-6, torch.tensor([0.99999994], device='cuda')]
    jobs = (L.WsJob * len(ws))(*[L.WsJob(w.data_ptr(), w.numel()) for w in ws])
    dev_jobs = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).cuda()
    table = torch.empty((len(ws), 2), device='cuda')
    scratch = torch.zeros((2 * len(ws),), dtype=torch.int32, device='cuda')
    ref = eng.weight_scale_table(ws)
    for _ in range(2):                           # (the second call finds the scratch words zeroed by the first)
        table.fill_(-1.0)
        L.check(L.lib().e3b_weight_scales(dev_jobs.data_ptr(), len(ws), table.data_ptr(), scratch.data_ptr(), eng._stream()),
                'weight_scales')
        assert torch.equal(table, ref), (table, ref)
    assert int(scratch.abs().sum()) == 0
    assert table[2].tolist() == [1.0, 1.0]
    for w, (up, down) in zip(ws, table.tolist()):
        if float(w.abs().max()) > 0:
            assert 1.0 <= float(w.abs().max()) * up < 2.0 + 1e-6 and up * down == 1.0


@pytest.mark.parametrize('C,sp,pool', [(16, (4, 6, 8), (2, 2, 2)), (8, (5, 7, 9), (2, 2, 2)), (24, (3, 9, 10), (1, 2, 2)), (40, (1, 11, 13), (1, 2, 2))])
@pytest.mark.parametrize('generic', [False, True])
def test_pooling_of_an_fp16_activation(eng, monkeypatch, C, sp, pool, generic):
    """inference path: the conv epilogue wrote the activated QH tensor, norm_act only pools it (ceil mode, MaxPool of
    models/unet.py:225-229); the 16-byte-unit kernel and the generic one"""
    if generic:
        monkeypatch.setenv('E3B_POOL_GENERIC', '1')
    x = dyadic((2, C) + sp, 17)
    _, pooled = eng.norm_act(qp(eng, x), None, None, write_a=False, pool=pool)
    ref = F.max_pool3d(x, pool, pool, ceil_mode=True)
    assert (pooled.D, pooled.H, pooled.W) == tuple(ref.shape[2:])
    assert_close(from_qh_ref(pooled, C), ref, 1e-7, 'pool')
    if eng.cpad16(C) != C:
        assert from_qh_ref(pooled, eng.cpad16(C))[:, C:].abs().max().item() == 0.0
