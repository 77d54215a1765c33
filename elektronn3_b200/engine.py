"""Host-side orchestration of the sm_100a kernels (libe3b.so) for the UNet forward / backward.

This is the glue between the reference-shaped ``torch.nn.Module`` (elektronn3_b200/unet.py) and the
C ABI (include/e3b.h): it owns no arithmetic.  PyTorch is used for device memory (caching allocator),
the current stream and the parameter tensors only.

fp32 tensors (conv outputs, activation gradients) live in the QP layout (see csrc/common.cuh): float32
``(N, ceil8(C)/4, D, H, W, 4)``; everything an MMA reads (input, activations, conv-output gradients) lives in
the QH operand layout: float16 ``(N, ceil16(C)/8, D, H, W, 8)``.
The sequence of operations follows the reference ``UNet.forward`` (models/unet.py:894-916),
``DownConv.forward`` (:244-253) and ``UpConv.forward`` (:384-408); the backward is what autograd
derives from them.
"""
import ctypes
import os
import threading

import torch

from . import _lib as L


def cpad8(c):
    return (c + 7) & ~7


def cpad16(c):
    return (c + 15) & ~15


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class QP:
    """A device tensor in quad-planar layout: float32 QP, or -- when ``half`` -- the float16 QH operand layout."""
    __slots__ = ('t', 'N', 'C', 'D', 'H', 'W', 'idx', 'half', 'scale', 'crop_off')

    def __init__(self, t, N, C, D, H, W):
        self.t, self.N, self.C, self.D, self.H, self.W = t, N, C, D, H, W
        self.idx = None            # pooled tensors: arg-max slots of the pooling windows (uint8)
        self.half = t.dtype == torch.float16
        self.scale = None          # gradients: device float[4] (bound bits, 2^k, 2^-k, -) of the fp16 scale
        self.crop_off = None       # gradient of a centre-cropped view: offset of the view inside the full tensor

    @staticmethod
    def empty(N, C, D, H, W, device):
        return QP(torch.empty((N, cpad8(C) // 4, D, H, W, 4), dtype=torch.float32, device=device), N, C, D, H, W)

    @staticmethod
    def empty_half(N, C, D, H, W, device, all_planes_written=False):
        """QH operand tensor; padding planes that no kernel writes (C % 16 in 1..8) must read as zero.
        all_planes_written: the producing kernel writes every 16-byte plane itself (pack / gather / add / upsample)."""
        alloc = torch.zeros if (cpad16(C) != cpad8(C) and not all_planes_written) else torch.empty
        return QP(alloc((N, cpad16(C) // 8, D, H, W, 8), dtype=torch.float16, device=device), N, C, D, H, W)

    @property
    def ptr(self):
        return self.t.data_ptr()

    @property
    def spatial(self):
        return (self.D, self.H, self.W)


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f'elektronn3_b200: {what} must be a CUDA tensor (this path has no CPU implementation)')
    if t.dtype != torch.float32:
        raise RuntimeError(f'elektronn3_b200: {what} must be float32, got {t.dtype}')


# ------------------------------------------------------------------------------------------ layout
def pack_input(x5):
    """NCDHW float32 -> QH (reference: the tensor `Trainer._train_step` moves to the device, trainer.py:515)"""
    _require_cuda(x5, 'input')
    x5 = x5.contiguous()
    N, C, D, H, W = x5.shape
    q = QP.empty_half(N, C, D, H, W, x5.device, all_planes_written=True)
    L.check(L.lib().e3b_pack_ncdhw(x5.data_ptr(), q.ptr, N, C, D, H, W, D, H, W, 0, 0, 0, _stream()), 'pack_ncdhw')
    return q


def unpack(q):
    if q.half:
        raise RuntimeError('unpack: expects a float32 QP tensor')
    out = torch.empty((q.N, q.C, q.D, q.H, q.W), dtype=torch.float32, device=q.t.device)
    L.check(L.lib().e3b_unpack_qp(q.ptr, out.data_ptr(), q.N, q.C, q.D, q.H, q.W, _stream()), 'unpack_qp')
    return out


def gather_tiles(vol, origins, B, C, tile, flip=0):
    """Predictor tile gather (inference.py:179-189): vol (C, Dv, Hv, Wv) device tensor, origins int32 (B,3).
    flip: bit mask (D, H, W) of the test-time-augmentation mirror applied while gathering (inference.py:215-223)."""
    D, H, W = tile
    q = QP.empty_half(B, C, D, H, W, vol.device, all_planes_written=True)
    L.check(L.lib().e3b_gather_tiles(vol.data_ptr(), origins.data_ptr(), q.ptr, B, C, D, H, W,
                                     vol.shape[-3], vol.shape[-2], vol.shape[-1], flip, _stream()), 'gather_tiles')
    return q


# ------------------------------------------------------------------------------------------ weights
def pack_weights(mode, w, scale, C0, C1, Co, k, wscale=None):
    """wscale: optional WeightScale -- the image is packed times its power of two 2^k (undone by the conv epilogue
    through conv_forward(w_unscale=...)), so that tiny / huge weights keep their 10 mantissa bits in fp16."""
    n = L.lib().e3b_packed_weight_floats(mode, C0, C1, Co, *k)
    if n <= 0:
        raise RuntimeError(f'elektronn3_b200: unsupported channel configuration C0={C0} C1={C1} Co={Co}')
    dst = torch.empty((n,), dtype=torch.float32, device=w.device)
    wc = w.detach()
    if not wc.is_contiguous():
        wc = wc.contiguous()
    L.check(L.lib().e3b_pack_weights(mode, wc.data_ptr(), _p(scale), wscale.up_ptr if wscale is not None else None,
                                     dst.data_ptr(), C0, C1, Co, *k, _stream()), 'pack_weights')
    return dst


class WeightScale:
    """One row (2^k, 2^-k) of a device table of per-tensor power-of-two weight scales."""
    __slots__ = ('table', 'row')

    def __init__(self, table, row):
        self.table, self.row = table, row

    @property
    def up_ptr(self):
        return self.table.data_ptr() + 8 * self.row

    @property
    def down_ptr(self):
        return self.table.data_ptr() + 8 * self.row + 4


def weight_scale_table(weights, channel_scales=None):
    """(P, 2) float32 device table: for every weight tensor the power of two 2^k that brings max|w| into [1, 2) and its
    inverse.  A handful of multi-tensor torch kernels for ALL tensors of the network (no host synchronisation).
    channel_scales: optional per-tensor (factors [Co], output-channel axis) folded into the weights at pack time
    (eval-mode BatchNorm); those tensors are measured after the fold."""
    ws = [w.detach() for w in weights]
    if channel_scales is not None and any(c is not None for c in channel_scales):
        amaxes = []
        for w, c in zip(ws, channel_scales):
            if c is None:
                amaxes.append(w.abs().max())
            else:
                f, axis = c
                per_channel = w.abs().amax([d for d in range(w.dim()) if d != axis])
                amaxes.append((per_channel * f.abs()).max())
        amax = torch.stack(amaxes)
    else:
        amax = torch.stack(torch._foreach_norm(ws, float('inf')))
    k = -torch.floor(torch.log2(amax.clamp_min(1e-37)))
    k = torch.where((amax > 0) & torch.isfinite(amax), k, torch.zeros_like(k)).clamp_(-100.0, 100.0)
    return torch.stack((torch.exp2(k), torch.exp2(-k)), 1).contiguous()


def conv_variant(C0, C1, n_total, k):
    """1 if the z-stacked kernel serves this convolution (weights then packed with mode 4/5), else 0"""
    return int(L.lib().e3b_conv_variant(C0, C1 or 0, n_total, k[0], k[1], k[2], 0))


class WeightCache:
    """Packed-weight images of one UNet instance.

    Training mode: NEVER served from the cache.  Parameters change every step, and the reference's own code
    writes them in ways no version counter sees (``p.data.copy_`` in training/swa.py:201 ``swap_swa_sgd``,
    ``p.data.addcdiv_`` in training/padam.py:94), so every training forward / backward re-packs (23 small launches
    for cfg 2 -- what an optimizer step cost before, too).
    Eval mode: an image is reused while (storage, version) of every tensor that entered it AND the cache epoch are
    unchanged.  The epoch moves on every training-mode forward (BatchNorm running statistics are written through raw
    pointers by e3b_norm_finalize), on ``train()`` / ``eval()``, ``load_state_dict`` and ``_apply`` of the module,
    and once per ``Predictor.predict`` call."""

    def __init__(self):
        self.d = {}
        self.epoch = 0

    def invalidate(self):
        self.epoch += 1
        self.d.clear()

    def get(self, key, params, make, training=False):
        if training:
            return make()
        sig = (self.epoch,) + tuple((p.data_ptr(), p._version) for p in params if p is not None)
        hit = self.d.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        val = make()
        self.d[key] = (sig, val)
        return val


# ------------------------------------------------------------------------------------------ conv
def conv_forward(src0, wpk, n_total, Co, k, pad, *, src1=None, off1=(0, 0, 0), bias=None, relu=False,
                 stats_channels=0, dst1_C=0, scatter=None, out_spatial=None, force_tz=0, half_out=False, variant=0,
                 w_unscale=None):
    """One launch of the implicit-GEMM kernel.  Sources are QH operand tensors; the output is float32 QP, or
    QH again with half_out (it is the next layer's operand).  A scaled gradient source (src0.scale) is
    un-scaled in the epilogue.  Returns (dst0, dst1, stats)."""
    a = L.ConvArgs()
    dev = src0.t.device
    if not src0.half or (src1 is not None and not src1.half):
        raise RuntimeError('conv: sources must be float16 QH operand tensors')
    if src0.scale is not None:
        a.out_scale = src0.scale.data_ptr() + 8
    if w_unscale is not None:
        a.w_unscale = w_unscale.down_ptr           # the image in `wpk` was packed times 2^k (WeightScale)
    a.src0, a.C0 = src0.ptr, src0.C
    a.N, a.D, a.H, a.W = src0.N, src0.D, src0.H, src0.W
    if src1 is not None:
        a.src1, a.C1 = src1.ptr, src1.C
        a.D1, a.H1, a.W1 = src1.D, src1.H, src1.W
        a.off1_d, a.off1_h, a.off1_w = off1
    a.kd, a.kh, a.kw = k
    a.pd, a.ph, a.pw = pad
    a.wpk = wpk.data_ptr()
    if bias is not None:
        a.bias, a.n_bias = bias.data_ptr(), bias.numel()
    a.n_total = n_total
    if scatter is None:
        Do, Ho, Wo = (src0.D + 2 * pad[0] - k[0] + 1, src0.H + 2 * pad[1] - k[1] + 1, src0.W + 2 * pad[2] - k[2] + 1)
    else:
        Do, Ho, Wo = out_spatial
        a.scatter = 1
        a.sd, a.sh, a.sw = scatter
        a.Ds, a.Hs, a.Ws = out_spatial
    dst0 = (QP.empty_half if half_out else QP.empty)(src0.N, Co, Do, Ho, Wo, dev)
    a.dst0, a.Cd0 = dst0.ptr, Co
    dst1 = None
    if dst1_C:
        dst1 = QP.empty(src0.N, dst1_C, Do, Ho, Wo, dev)
        a.dst1, a.Cd1 = dst1.ptr, dst1_C
    a.relu = 1 if relu else 0
    a.half_out = 1 if half_out else 0
    stats = None
    if stats_channels:
        stats = torch.empty((src0.N, stats_channels, 2), dtype=torch.float64, device=dev)
        a.stats, a.stats_channels = stats.data_ptr(), stats_channels
    a.force_tz = force_tz
    a.variant = variant            # must match the mode `wpk` was packed with (conv_variant)
    L.check(L.lib().e3b_conv(ctypes.byref(a), _stream()), 'conv')
    return dst0, dst1, stats


class _WgradDefer(threading.local):
    """per thread (nn.DataParallel replicas run in threads): the deferred split-K reductions of the backward pass in flight"""
    jobs = None


_wg_defer = _WgradDefer()


def flush_wgrad():
    """reduce the split-K partials of every wgrad() call since begin_wgrad_defer() in one launch (per 16 layers)"""
    jobs, _wg_defer.jobs = _wg_defer.jobs, None
    if jobs:
        arr = (L.WgradArgs * len(jobs))(*[j[0] for j in jobs])
        L.check(L.lib().e3b_wgrad_reduce_batched(arr, len(jobs), _stream()), 'wgrad_reduce_batched')


def wgrad(src0, dy, Co, k, pad, dw_shape, *, src1=None, off1=(0, 0, 0), layout=0, up_taps=0, up_co=0):
    """dW of a convolution: src0 / src1 the QH activations it read, dy the (scaled) QH gradient of its output.  The kernel
    reads the operand tensors of the forward / dgrad kernels directly (MN-major MMA operands): no extra copies.
    Inside a network backward pass (engine._backward) the split-K reduction is deferred: the returned tensor is written by
    flush_wgrad() at the end of the pass."""
    a = L.WgradArgs()
    dev = src0.t.device
    if not (src0.half and dy.half and (src1 is None or src1.half)):
        raise RuntimeError('wgrad: operands must be float16 QH tensors')
    a.src0, a.C0 = src0.ptr, src0.C
    a.N, a.D, a.H, a.W = src0.N, src0.D, src0.H, src0.W
    if src1 is not None:
        a.src1, a.C1 = src1.ptr, src1.C
        a.D1, a.H1, a.W1 = src1.D, src1.H, src1.W
        a.off1_d, a.off1_h, a.off1_w = off1
    a.dy, a.Co = dy.ptr, Co
    a.kd, a.kh, a.kw = k
    a.pd, a.ph, a.pw = pad
    dw = torch.empty(dw_shape, dtype=torch.float32, device=dev)
    a.dw, a.layout, a.up_taps, a.up_co = dw.data_ptr(), layout, up_taps, up_co
    n = L.lib().e3b_wgrad_workspace_floats(ctypes.byref(a))
    if n <= 0:
        L.check(1, 'wgrad_workspace_floats')
    ws = torch.empty((n,), dtype=torch.float32, device=dev)
    a.workspace = ws.data_ptr()
    if dy.scale is not None:
        a.dy_unscale = dy.scale.data_ptr() + 8          # the gradient tensor holds 2^k * dy
    if _wg_defer.jobs is not None:
        a.defer_reduce = 1
        _wg_defer.jobs.append((a, ws, dw, dy.scale))    # (keeps the workspace and the scale alive until the flush)
    L.check(L.lib().e3b_wgrad(ctypes.byref(a), _stream()), 'wgrad')
    return dw


# ------------------------------------------------------------------------------------------ norm
MODE_NONE, MODE_GROUP, MODE_BATCH, MODE_BATCH_EVAL = 0, 1, 2, 3


class NormState:
    __slots__ = ('scale', 'shift', 'mean', 'rstd')


def norm_finalize(stats, mode, G, N, C, S, gamma, beta, eps, rm, rv, momentum, device):
    Cp = cpad8(C)
    buf = torch.empty((4, N, Cp), dtype=torch.float32, device=device)
    st = NormState()
    st.scale, st.shift, st.mean, st.rstd = buf[0], buf[1], buf[2], buf[3]
    L.check(L.lib().e3b_norm_finalize(_p(stats), mode, G, N, C, S, _p(gamma), _p(beta), eps, _p(rm), _p(rv), momentum,
                                      st.scale.data_ptr(), st.shift.data_ptr(), st.mean.data_ptr(),
                                      st.rstd.data_ptr(), _stream()), 'norm_finalize')
    return st


ACT_RELU = (1, 0.0, None)


def act_code(act, training):
    """-> (code, negative slope, slope parameter or None) as the kernels take them (include/e3b.h) for an activation module
    produced by get_activation (models/unet.py:183-199): 'relu', 'leaky', 'prelu', 'rrelu' (eval mode), 'silu', 'lin' or a
    module of those types.  nn.PReLU's learned slope stays in device memory (third entry: the parameter)."""
    import torch.nn as nn
    if act is None or isinstance(act, nn.ReLU):
        return ACT_RELU
    if isinstance(act, nn.Identity):
        return (0, 0.0, None)
    if isinstance(act, nn.LeakyReLU):
        return (1, float(act.negative_slope), None)
    if isinstance(act, nn.PReLU):
        if act.weight.numel() != 1:
            raise NotImplementedError('nn.PReLU with one slope per channel is not on the B200 path (num_parameters=1 is)')
        return (1, 0.0, act.weight)
    if isinstance(act, nn.RReLU):
        if training:
            raise NotImplementedError('nn.RReLU draws a random slope per element in training mode: not on the B200 path '
                                      '(eval mode, with the mean slope, is)')
        return (1, (float(act.lower) + float(act.upper)) / 2, None)
    if isinstance(act, nn.SiLU):
        return (2, 0.0, None)
    raise NotImplementedError(f'activation module {type(act).__name__} is not on the B200 path')


def norm_act(y, scale, shift, *, write_a=True, pool=None, relu=True, save=False, act=None):
    """a = act(y*scale+shift) (QH) and optionally the ceil-mode max-pooled tensor (QH).  save=True (a backward pass will
    follow) also records the arg-max slots of the pooling windows.  With write_a=False `y` is already a QH activation
    (eval path) and is only pooled.  act: (code, slope) of act_code; default ReLU (relu=False: identity)."""
    if act is None:
        act = ACT_RELU if relu else (0, 0.0, None)
    slope_dev = act[2].detach() if len(act) > 2 and act[2] is not None else None
    dev = y.t.device
    if write_a == y.half:
        raise RuntimeError('norm_act: y must be float32 QP when a is written, a QH activation otherwise')
    a = QP.empty_half(y.N, y.C, y.D, y.H, y.W, dev) if write_a else None
    pooled = None
    pk = (1, 1, 1)
    if pool is not None:
        pk = pool
        pooled = QP.empty_half(y.N, y.C, -(-y.D // pk[0]), -(-y.H // pk[1]), -(-y.W // pk[2]), dev)
    pidx = None
    if save and pooled is not None:
        # arg-max slots for the backward pass (what MaxPool(return_indices=True) keeps in the reference)
        pidx = pooled.idx = torch.empty(pooled.t.shape[:1] + (cpad8(y.C) // 4,) + pooled.t.shape[2:5] + (4,), dtype=torch.uint8,
                                        device=dev)
    L.check(L.lib().e3b_norm_act(y.ptr, _p(scale), _p(shift), a.ptr if a else None, pooled.ptr if pooled else None,
                                 _p(pidx), y.N, y.C, y.D, y.H, y.W, pk[0], pk[1], pk[2],
                                 act[0], act[1], _p(slope_dev), 1 if y.half else 0, _stream()), 'norm_act')
    return a, pooled


# ------------------------------------------------------------------------------------------ network description
class ConvSpec:
    """A conv3 layer (models/unet.py:131-149) + the norm/activation that follows it."""

    def __init__(self, name, conv, norm, C0, C1, act=None, explicit_pad=False):
        self.name, self.conv, self.norm, self.act = name, conv, norm, act
        self.residual = False      # a shortcut is added between this conv and its norm (resunet ConvBlock.conv2)
        w = conv.weight
        self.Co = w.shape[0]
        self.C0, self.C1 = C0, C1
        ks = tuple(w.shape[2:])
        pads = tuple(conv.padding)
        if len(ks) == 2:
            ks, pads = (1,) + ks, (0,) + pads
        if explicit_pad:           # the source tensor already carries the zero padding (ResizeSpec)
            pads = (0, 0, 0)
        self.k, self.pad = ks, pads
        self.n_total = cpad16(self.Co)
        self.n_total_dgrad = cpad16(cpad8(C0) + (cpad8(C1) if C1 else 0))
        self._variants = None

    @property
    def variants(self):
        """(forward, dgrad) kernel variant: asked once from the library (it knows its shared-memory plans)"""
        if self._variants is None:
            self._variants = (conv_variant(self.C0, self.C1, self.n_total, self.k),
                              conv_variant(self.Co, 0, self.n_total_dgrad, self.k))
        return self._variants

    def w5(self):
        w = self.conv.weight
        return w


class UpSpec:
    """upconv2 'transpose' (models/unet.py:152-165) + norm0/act0 of UpConv"""

    def __init__(self, name, up, norm, act=None):
        self.name, self.up, self.norm, self.act = name, up, norm, act
        w = up.weight
        self.Ci, self.Co = w.shape[0], w.shape[1]
        s = tuple(w.shape[2:])
        if len(s) == 2:
            s = (1,) + s
        self.s = s
        self.taps = s[0] * s[1] * s[2]
        self.n_total = self.taps * cpad16(self.Co)


class ResizeSpec(ConvSpec):
    """upconv2 'resizeconv_*' (ResizeConv, models/unet.py:411-449): nn.Upsample + conv3 / conv1, + norm0/act0 of UpConv.
    The up-sampled tensor is produced with the conv's zero padding made explicit, so the conv is a ConvSpec with pad 0."""

    def __init__(self, name, rc, norm, Ci, act=None):
        super().__init__(name + '.conv', rc.conv, norm, Ci, 0, act=act, explicit_pad=True)
        sf = rc.scale_factor
        sf = (sf,) * 3 if isinstance(sf, int) else tuple(int(v) for v in sf)
        if len(sf) == 2:
            sf = (1,) + sf
        if rc.dim == 2:
            sf = (1, sf[1], sf[2])
        self.s = sf
        self.linear = 0 if rc.upsampling_mode == 'nearest' else 1
        cpads = tuple(rc.conv.padding)
        if len(cpads) == 2:
            cpads = (0,) + cpads
        self.conv_pad = cpads          # the padding of rc.conv (1 on 3-tap axes, 0 for conv1): where data starts in the padded tensor


def norm_mode(norm, training):
    """-> (mode, G) for a module produced by get_normalization (models/unet.py:77-111)"""
    import torch.nn as nn
    if norm is None or isinstance(norm, nn.Identity):
        return MODE_NONE, 1
    if isinstance(norm, nn.GroupNorm):
        return MODE_GROUP, norm.num_groups
    if isinstance(norm, nn.modules.instancenorm._InstanceNorm):
        if norm.track_running_stats:
            raise NotImplementedError('InstanceNorm with running stats is not supported')
        return MODE_GROUP, norm.num_features
    if isinstance(norm, nn.modules.batchnorm._BatchNorm):
        use_batch = training or (norm.running_mean is None)
        return (MODE_BATCH if use_batch else MODE_BATCH_EVAL), 1
    raise NotImplementedError(f'normalization module {type(norm).__name__} is not supported')


class Unit:
    """Everything one conv -> norm -> relu [-> pool] stage leaves behind for the backward pass."""
    __slots__ = ('spec', 'src0', 'src1', 'off1', 'y', 'a', 'pooled', 'pool', 'mode', 'G', 'nstate', 'stats', 'dec', 'act',
                 'resize', 'res', 'res_grads', 'dslope')


class WeightSet:
    """What one forward (and the backward that follows it) needs besides the raw parameters: the per-tensor power-of-two
    weight scales, in eval mode the folded BatchNorm factors, in training mode the freshly packed weight images."""
    __slots__ = ('scales', 'folds', 'images')


class TrainImages:
    """Training mode re-packs every weight image every step (WeightCache): persistent destination buffers, one persistent
    scale table and ONE launch for all images (forward and dgrad layouts) through the library's job table."""

    def __init__(self, net, specs):
        lib = L.lib()
        mods = [sp.conv if isinstance(sp, ConvSpec) else sp.up for sp in specs]
        dev = mods[0].weight.device
        self.sig = tuple(m.weight.data_ptr() for m in mods)
        self.table = torch.empty((len(specs), 2), dtype=torch.float32, device=dev)
        self.scales = {sp.name: WeightScale(self.table, i) for i, sp in enumerate(specs)}
        self.img = {}
        jobs = []
        for i, (sp, mod) in enumerate(zip(specs, mods)):
            w = mod.weight
            if not w.is_contiguous():
                raise RuntimeError(f'elektronn3_b200: parameter {sp.name}.weight must be contiguous')
            if isinstance(sp, ConvSpec):
                variants = ((4 if sp.variants[0] else 0, 'fwd'), (5 if sp.variants[1] else 1, 'bwd'))
                dims = (sp.C0, sp.C1, sp.Co) + tuple(sp.k)
            else:
                variants = ((2, 'fwd'), (3, 'bwd'))
                dims = (sp.Ci, 0, sp.Co) + tuple(sp.s)
            for mode, key in variants:
                n = lib.e3b_packed_weight_floats(mode, *dims)
                if n <= 0:
                    raise RuntimeError(f'elektronn3_b200: unsupported channel configuration for {sp.name}')
                dst = torch.empty((n,), dtype=torch.float32, device=dev)
                self.img[(sp.name, key)] = dst
                j = L.PackJob()
                j.w, j.scale, j.wscale, j.dst = w.data_ptr(), None, self.scales[sp.name].up_ptr, dst.data_ptr()
                j.mode, j.C0, j.C1, j.Co, j.kd, j.kh, j.kw = (mode,) + dims
                jobs.append(j)
        self.njobs = len(jobs)
        arr = (L.PackJob * self.njobs)(*jobs)
        host = torch.empty((int(lib.e3b_pack_job_table_bytes(self.njobs)),), dtype=torch.uint8)
        blocks = ctypes.c_int64(0)
        L.check(lib.e3b_pack_jobs_fill(arr, self.njobs, host.data_ptr(), ctypes.byref(blocks)), 'pack_jobs_fill')
        self.blocks = int(blocks.value)
        self.dev_table = host.to(dev)
        self.weights = [m.weight for m in mods]
        # descriptors of the weight tensors for the one-launch scale table (e3b_weight_scales)
        ws = (L.WsJob * len(mods))(*[L.WsJob(m.weight.data_ptr(), m.weight.numel()) for m in mods])
        self.ws_jobs = torch.frombuffer(bytearray(bytes(ws)), dtype=torch.uint8).to(dev)
        self.ws_scratch = torch.zeros((2 * len(mods),), dtype=torch.int32, device=dev)

    def pack(self):
        L.check(L.lib().e3b_weight_scales(self.ws_jobs.data_ptr(), len(self.weights), self.table.data_ptr(),
                                          self.ws_scratch.data_ptr(), _stream()), 'weight_scales')
        L.check(L.lib().e3b_pack_weights_batched(self.dev_table.data_ptr(), self.njobs, self.blocks, _stream()),
                'pack_weights_batched')


class Block:
    """conv1 -> norm -> act -> conv2 [+ shortcut(block input)] -> norm -> act: DownConv / UpConv of models/unet.py (no shortcut)
    and ConvBlock of models/resunet.py:212-261.  res: None, 'identity' or the ConvSpec of the 1x1x1 projection."""

    def __init__(self, c1, c2, res=None):
        self.c1, self.c2, self.res = c1, c2, res
        c2.residual = res is not None


class Net:
    """Flat description of a UNet instance (built by elektronn3_b200.unet.UNet / resunet.UNet): down = [(blocks, pool kernel)],
    up = [(UpSpec | ResizeSpec, blocks)]."""

    def __init__(self, down, up, final_conv, dim, cache, merge_add=False):
        self.down, self.up, self.final, self.dim, self.cache = down, up, final_conv, dim, cache
        self.merge_add = merge_add       # merge_mode='add' (models/unet.py:399-401) instead of the channel concat
        self.wset = None
        self.train_images = None

    def specs(self):
        def of_blocks(blocks):
            for b in blocks:
                yield b.c1
                yield b.c2
                if isinstance(b.res, ConvSpec):
                    yield b.res
        for blocks, _ in self.down:
            yield from of_blocks(blocks)
        for ups, blocks in self.up:
            yield ups
            yield from of_blocks(blocks)


def prepare_weights(net, training):
    """-> WeightSet; cached like the packed images (never in training mode)."""
    specs = list(net.specs())
    if training and not any(norm_mode(sp.norm, True)[0] == MODE_BATCH_EVAL for sp in specs):
        ti = net.train_images
        sig = tuple((sp.conv if isinstance(sp, ConvSpec) else sp.up).weight.data_ptr() for sp in specs)
        if ti is None or ti.sig != sig:
            ti = net.train_images = TrainImages(net, specs)
        ti.pack()
        ws = WeightSet()
        ws.scales, ws.folds, ws.images = ti.scales, {}, ti.img
        return ws
    params = []
    for sp in specs:
        mod = sp.conv if isinstance(sp, ConvSpec) else sp.up
        n = sp.norm
        params += [mod.weight, mod.bias, getattr(n, 'weight', None), getattr(n, 'bias', None),
                   getattr(n, 'running_mean', None), getattr(n, 'running_var', None)]

    def make():
        ws = WeightSet()
        ws.folds = {}
        weights, cscales = [], []
        for sp in specs:
            is_conv = isinstance(sp, ConvSpec)
            mod = sp.conv if is_conv else sp.up
            weights.append(mod.weight)
            if norm_mode(sp.norm, training)[0] == MODE_BATCH_EVAL and not getattr(sp, 'residual', False):
                f, b = _bn_fold(mod, sp.norm)
                ws.folds[sp.name] = (f, b)
                cscales.append((f, 0 if is_conv else 1))
            else:
                cscales.append(None)
        table = weight_scale_table(weights, cscales)
        ws.scales = {sp.name: WeightScale(table, i) for i, sp in enumerate(specs)}
        ws.images = None
        return ws
    return net.cache.get(('weightset',), params, make, training)


def _bn_fold(conv, norm):
    """eval-mode BatchNorm folded into the conv that precedes it: per-channel
    scale for the weights and the new bias.  O(C) torch ops, cached with the packed weights."""
    s = torch.rsqrt(norm.running_var.detach() + norm.eps)
    if norm.weight is not None:
        s = s * norm.weight.detach()
    b = conv.bias.detach() if conv.bias is not None else torch.zeros_like(s)
    b = (b - norm.running_mean.detach()) * s
    if norm.bias is not None:
        b = b + norm.bias.detach()
    return s.contiguous(), b.contiguous()


def _conv_weights(net, spec, mode, training):
    """-> (wpk, bias, WeightScale) for forward; BN-eval folding applied when the following norm allows it"""
    nm, _ = norm_mode(spec.norm, training)
    conv = spec.conv
    pmode = 4 if spec.variants[0] else 0
    wsc = net.wset.scales[spec.name]
    if net.wset.images is not None:
        return net.wset.images[(spec.name, 'fwd')], (conv.bias.detach() if conv.bias is not None else None), wsc
    if nm == MODE_BATCH_EVAL and not spec.residual:
        n = spec.norm
        params = (conv.weight, conv.bias, n.weight, n.bias, n.running_mean, n.running_var)
        s, b = net.wset.folds[spec.name]
        wpk = net.cache.get((spec.name, 'fwd_fold'), params,
                            lambda: pack_weights(pmode, conv.weight, s, spec.C0, spec.C1, spec.Co, spec.k, wscale=wsc), training)
        return wpk, b, wsc
    wpk = net.cache.get((spec.name, 'fwd'), (conv.weight,),
                        lambda: pack_weights(pmode, conv.weight, None, spec.C0, spec.C1, spec.Co, spec.k, wscale=wsc), training)
    return wpk, (conv.bias.detach() if conv.bias is not None else None), wsc


def _shortcut(net, res, training):
    """the shortcut branch of a resunet ConvBlock (models/resunet.py:246-250,257-258): -> (tensor, is_half).  Identity: the
    block input itself (a QH activation); otherwise the 1x1x1 projection of the (virtually concatenated) block input, fp32 QP."""
    kind, r0, r1, roff = res
    if not isinstance(kind, ConvSpec):
        return r0, 1
    wp, bp, wscp = _conv_weights(net, kind, 0, training)
    r, _, _ = conv_forward(r0, wp, kind.n_total, kind.Co, kind.k, kind.pad, src1=r1, off1=roff, bias=bp,
                           variant=kind.variants[0], w_unscale=wscp)
    return r, 0


def _run_unit(net, spec, src0, src1, off1, pool, training, save, res=None):
    """conv -> [+ shortcut] -> norm -> act [-> pool]  (DownConv.forward models/unet.py:244-253, UpConv :402-407; with
    res = (kind, block input ...): ConvBlock.forward of models/resunet.py:252-261)"""
    mode, G = norm_mode(spec.norm, training)
    act = act_code(spec.act, training)
    wpk, bias, wsc = _conv_weights(net, spec, 0, training)
    u = Unit()
    u.spec, u.src0, u.src1, u.off1, u.pool, u.mode, u.G, u.act = spec, src0, src1, off1, pool, mode, G, act
    u.pooled = u.nstate = u.stats = u.dec = u.resize = u.res = u.res_grads = None
    var = spec.variants[0]
    if res is not None:
        # y = conv2(..) + b; y += shortcut; the statistics of the SUM feed the norm (never folded into the weights)
        y, _, _ = conv_forward(src0, wpk, spec.n_total, spec.Co, spec.k, spec.pad, src1=src1, off1=off1, bias=bias,
                               variant=var, w_unscale=wsc)
        r, r_half = _shortcut(net, res, training)
        if (r.N, r.C) + r.spatial != (y.N, y.C) + y.spatial:
            raise RuntimeError(f'residual shortcut of shape {(r.N, r.C) + r.spatial} cannot be added to the conv output '
                               f'{(y.N, y.C) + y.spatial} (models/resunet.py:257-258; VALID convolutions shrink it)')
        stats = None
        if mode in (MODE_GROUP, MODE_BATCH):
            stats = torch.empty((y.N, spec.Co, 2), dtype=torch.float64, device=y.t.device)
        L.check(L.lib().e3b_residual_add(y.ptr, r.ptr, r_half, _p(stats), y.N, spec.Co, y.D * y.H * y.W, _stream()),
                'residual_add')
        u.y, u.stats, u.res = y, stats, res
        if mode == MODE_NONE:
            u.a, u.pooled = norm_act(y, None, None, pool=pool, save=save, act=act)
        else:
            n = spec.norm
            rm = rv = None
            mom = 0.0
            if mode == MODE_BATCH:
                rm, rv, mom = _bn_running(n, training)
            elif mode == MODE_BATCH_EVAL:
                rm, rv = n.running_mean, n.running_var
            u.nstate = norm_finalize(stats, mode, G, y.N, spec.Co, y.D * y.H * y.W, _affine(n, 'weight'), _affine(n, 'bias'),
                                     n.eps, rm, rv, mom, y.t.device)
            u.a, u.pooled = norm_act(y, u.nstate.scale, u.nstate.shift, pool=pool, save=save, act=act)
    elif (mode == MODE_BATCH_EVAL or (mode == MODE_NONE and not save)) and act == ACT_RELU:
        # inference: the conv epilogue (folded BN, bias, ReLU) writes the next layer's operand directly
        a, _, _ = conv_forward(src0, wpk, spec.n_total, spec.Co, spec.k, spec.pad, src1=src1, off1=off1, bias=bias,
                               relu=True, half_out=True, variant=var, w_unscale=wsc)
        u.y = u.a = a
        if pool is not None:
            _, u.pooled = norm_act(a, None, None, write_a=False, pool=pool)
    elif mode == MODE_NONE or mode == MODE_BATCH_EVAL:
        # no normalisation (or eval-mode BatchNorm folded into the weights) with a backward pass to follow / an activation
        # the conv epilogue does not fuse: y is kept in float32 (identity affine), the activation kernel follows
        y, _, stats = conv_forward(src0, wpk, spec.n_total, spec.Co, spec.k, spec.pad, src1=src1, off1=off1, bias=bias,
                                   variant=var, w_unscale=wsc)
        u.y = y
        u.a, u.pooled = norm_act(y, None, None, pool=pool, save=save, act=act)
    else:
        n = spec.norm
        y, _, stats = conv_forward(src0, wpk, spec.n_total, spec.Co, spec.k, spec.pad, src1=src1, off1=off1,
                                   bias=bias, stats_channels=spec.Co, variant=var, w_unscale=wsc)
        S = y.D * y.H * y.W
        rm = rv = None
        mom = 0.0
        if mode == MODE_BATCH:
            rm, rv, mom = _bn_running(n, training)
        u.nstate = norm_finalize(stats, mode, G, y.N, spec.Co, S, _affine(n, 'weight'), _affine(n, 'bias'), n.eps,
                                 rm, rv, mom, y.t.device)
        u.y, u.stats = y, stats
        u.a, u.pooled = norm_act(y, u.nstate.scale, u.nstate.shift, pool=pool, save=save, act=act)
    if not save:
        u.y = u.src0 = u.src1 = u.res = None
    return u


def _run_blocks(net, blocks, src0, src1, off1, pool, training, save):
    """a stack of Blocks; the last one is followed by the pooling.  -> [(unit of conv1, unit of conv2)]"""
    units = []
    for bi, blk in enumerate(blocks):
        u1 = _run_unit(net, blk.c1, src0, src1, off1, None, training, save)
        res = (blk.res, src0, src1, off1) if blk.res is not None else None
        u2 = _run_unit(net, blk.c2, u1.a, None, (0, 0, 0), pool if bi == len(blocks) - 1 else None, training, save, res=res)
        units.append((u1, u2))
        src0, src1, off1 = u2.a, None, (0, 0, 0)
    return units


def _affine(n, name):
    t = getattr(n, name, None)
    return None if t is None else t.detach()


def _bn_running(n, training):
    """BatchNorm train-mode side effects (running stats, num_batches_tracked): torch semantics"""
    if not training or n.running_mean is None or not n.track_running_stats:
        return None, None, 0.0
    if n.momentum is None:
        raise NotImplementedError('BatchNorm(momentum=None) (cumulative average) is not supported')
    if n.num_batches_tracked is not None:
        n.num_batches_tracked.add_(1)
    return n.running_mean, n.running_var, float(n.momentum)


def _autocrop(full, enc_spatial):
    """autocrop (models/unet.py:256-325) -> (extents the up-sampled tensor is cropped to, offset of the centre crop of the
    skip tensor): from_up loses one trailing voxel where (u - d) is odd (:294-301), from_down is centre-cropped (:303-324)"""
    out_sp = tuple(u_ - ((u_ - d_) % 2) for u_, d_ in zip(full, enc_spatial))
    for u_, d_ in zip(out_sp, enc_spatial):
        if u_ > d_:
            raise RuntimeError(f'autocrop: upsampled extent {out_sp} exceeds the skip tensor {enc_spatial} '
                               '(models/unet.py:303-324 cannot crop from_down to a larger shape)')
    return out_sp, tuple((d_ - u_) // 2 for u_, d_ in zip(out_sp, enc_spatial))


def _run_resize(net, spec, dec, enc, training, save):
    """ResizeConv -> autocrop -> norm0 -> act0  (UpConv.forward models/unet.py:385-398 with up_mode='resizeconv_*',
    ResizeConv :411-449).  The up-sampled tensor carries the conv's zero padding explicitly: fine voxel f sits at
    f + conv_pad, and only the fine voxels the (auto)cropped conv output reads are stored -- the conv is then VALID."""
    s, k, cp = spec.s, spec.k, spec.conv_pad
    full = (dec.D * s[0], dec.H * s[1], dec.W * s[2])
    out_sp, off1 = _autocrop(full, enc.spatial)
    ext = tuple(o + kk - 1 for o, kk in zip(out_sp, k))
    R = tuple(min(f, e - p) for f, e, p in zip(full, ext, cp))
    up = QP.empty_half(dec.N, dec.C, ext[0], ext[1], ext[2], dec.t.device, all_planes_written=True)
    geom = (dec.D, dec.H, dec.W) + ext + tuple(s) + tuple(cp) + R + (spec.linear,)
    L.check(L.lib().e3b_upsample_qh(dec.ptr, up.ptr, dec.N, dec.C, *geom, _stream()), 'upsample_qh')
    u = _run_unit(net, spec, up, None, (0, 0, 0), None, training, save)
    u.resize = geom
    return u, off1


def _resize_bwd(u, dup):
    """gradient of the padded up-sampled tensor (dgrad output, QP) -> gradient of the coarse decoder tensor (QP)"""
    geom = u.resize
    g = QP.empty(dup.N, dup.C, geom[0], geom[1], geom[2], dup.t.device)
    L.check(L.lib().e3b_upsample_bwd_qp(dup.ptr, g.ptr, dup.N, dup.C, *geom, _stream()), 'upsample_bwd_qp')
    return g


def add_qh(a, b, off):
    """merge_mode='add' (models/unet.py:399-401): a + b[centre crop at off], QH operand tensors"""
    out = QP.empty_half(a.N, a.C, a.D, a.H, a.W, a.t.device, all_planes_written=True)
    L.check(L.lib().e3b_add_qh(a.ptr, b.ptr, out.ptr, a.N, a.C, a.D, a.H, a.W, b.D, b.H, b.W, off[0], off[1], off[2],
                               _stream()), 'add_qh')
    return out


def _run_up(net, spec, dec, enc, training, save):
    """upconv -> autocrop -> norm0 -> act0  (UpConv.forward models/unet.py:385-398)"""
    mode, G = norm_mode(spec.norm, training)
    act = act_code(spec.act, training)
    up = spec.up
    full = (dec.D * spec.s[0], dec.H * spec.s[1], dec.W * spec.s[2])
    out_sp, off1 = _autocrop(full, enc.spatial)
    u = Unit()
    u.spec, u.src0, u.src1, u.off1, u.pool, u.mode, u.G, u.act = spec, dec, None, (0, 0, 0), None, mode, G, act
    u.pooled = u.nstate = u.stats = u.resize = u.res = u.res_grads = None
    u.dec = dec
    bias = up.bias.detach() if up.bias is not None else None
    wsc = net.wset.scales[spec.name]
    if net.wset.images is not None:
        wpk = net.wset.images[(spec.name, 'fwd')]
    elif mode == MODE_BATCH_EVAL:
        n = spec.norm
        params = (up.weight, up.bias, n.weight, n.bias, n.running_mean, n.running_var)
        fs, bias = net.wset.folds[spec.name]

        def make():
            # the transposed-conv weight is (Ci, Co, ...): scale its output channel axis on the host
            shape = (1, -1) + (1,) * (up.weight.dim() - 2)
            return pack_weights(2, (up.weight.detach() * fs.view(shape)).contiguous(), None, spec.Ci, 0, spec.Co,
                                spec.s, wscale=wsc)
        wpk = net.cache.get((spec.name, 'up_fold'), params, make, training)
    else:
        wpk = net.cache.get((spec.name, 'up'), (up.weight,),
                            lambda: pack_weights(2, up.weight, None, spec.Ci, 0, spec.Co, spec.s, wscale=wsc), training)
    if (mode == MODE_BATCH_EVAL or (mode == MODE_NONE and not save)) and act == ACT_RELU:
        a, _, _ = conv_forward(dec, wpk, spec.n_total, spec.Co, (1, 1, 1), (0, 0, 0), bias=bias, relu=True,
                               scatter=spec.s, out_spatial=out_sp, half_out=True, w_unscale=wsc)
        u.y = u.a = a
    elif mode == MODE_NONE or mode == MODE_BATCH_EVAL:
        y, _, _ = conv_forward(dec, wpk, spec.n_total, spec.Co, (1, 1, 1), (0, 0, 0), bias=bias,
                               scatter=spec.s, out_spatial=out_sp, w_unscale=wsc)
        u.y = y
        u.a, _ = norm_act(y, None, None, act=act)
    else:
        n = spec.norm
        y, _, stats = conv_forward(dec, wpk, spec.n_total, spec.Co, (1, 1, 1), (0, 0, 0), bias=bias,
                                   stats_channels=spec.Co, scatter=spec.s, out_spatial=out_sp, w_unscale=wsc)
        rm = rv = None
        mom = 0.0
        if mode == MODE_BATCH:
            rm, rv, mom = _bn_running(n, training)
        u.nstate = norm_finalize(stats, mode, G, y.N, spec.Co, y.D * y.H * y.W, _affine(n, 'weight'),
                                 _affine(n, 'bias'), n.eps, rm, rv, mom, y.t.device)
        u.y, u.stats = y, stats
        u.a, _ = norm_act(y, u.nstate.scale, u.nstate.shift, act=act)
    if not save:
        u.y = u.dec = u.src0 = None
    return u, off1


class Tape:
    __slots__ = ('down', 'up', 'final_in', 'in_shape', 'squeeze', 'wset')


def forward_features(net, x, training, save):
    """Everything up to (not including) conv_final.  x: (N,C,D,H,W) or (N,C,H,W) float32 CUDA.
    Returns (QP features, Tape)."""
    squeeze = x.dim() == 4
    x5 = x.unsqueeze(2) if squeeze else x
    if x5.dim() != 5:
        raise RuntimeError(f'expected a {net.dim + 2}-dimensional input, got shape {tuple(x.shape)}')
    cin = net.down[0][0][0].c1.C0
    if x5.shape[1] != cin:
        raise RuntimeError(f'expected {cin} input channels, got {x5.shape[1]}')
    cur = pack_input(x5)
    return forward_features_qp(net, cur, training, save, squeeze, tuple(x.shape))


def forward_features_qp(net, cur, training, save, squeeze=False, in_shape=None):
    tape = Tape()
    tape.down, tape.up, tape.squeeze, tape.in_shape = [], [], squeeze, in_shape
    net.wset = tape.wset = prepare_weights(net, training)
    enc = []
    for blocks, pool in net.down:
        units = _run_blocks(net, blocks, cur, None, (0, 0, 0), pool, training, save)
        last = units[-1][1]
        enc.append(last.a)
        cur = last.pooled if pool is not None else last.a
        tape.down.append(units)
    for i, (ups, blocks) in enumerate(net.up):
        e = enc[-(i + 2)]
        u0, off1 = (_run_resize if isinstance(ups, ResizeSpec) else _run_up)(net, ups, cur, e, training, save)
        add_info = None
        if net.merge_add:
            units = _run_blocks(net, blocks, add_qh(u0.a, e, off1), None, (0, 0, 0), None, training, save)
            add_info = (off1, e.spatial)
        else:
            units = _run_blocks(net, blocks, u0.a, e, off1, None, training, save)
        cur = units[-1][1].a
        tape.up.append((u0, units, len(net.down) - 2 - i, add_info))
    tape.final_in = cur
    return cur, tape


def head(feat, conv_final, out_mode=0, dst=None, crop=None, dst_origin=None, dst_single=False, flip=0, accumulate=False,
         acc_scale=1.0, threshold=None, round_half=False):
    """conv_final (models/unet.py:881,912) [+ Softmax(1) / Argmax of Predictor, inference.py:443-456] and
    the Predictor's crop-and-place.  out_mode 0 logits, 1 softmax, 2 argmax(uint8).  flip / accumulate / acc_scale:
    test-time augmentation (un-mirror, mean); threshold: nn.Threshold before Argmax; round_half: float16=True."""
    w, b = conv_final.weight.detach(), conv_final.bias
    Co = w.shape[0]
    a = L.HeadArgs()
    a.a, a.N, a.C, a.D, a.H, a.W = feat.ptr, feat.N, feat.C, feat.D, feat.H, feat.W
    a.w, a.b, a.Co = w.data_ptr(), _p(b.detach() if b is not None else None), Co
    a.out_mode = out_mode
    if crop is None:
        crop = ((0, 0, 0), feat.spatial)
    (a.c0_d, a.c0_h, a.c0_w), (a.cn_d, a.cn_h, a.cn_w) = crop
    if dst is None:
        oc = 1 if out_mode == 2 else Co
        dst = torch.empty((feat.N, oc) + tuple(crop[1]), dtype=torch.uint8 if out_mode == 2 else torch.float32,
                          device=feat.t.device)
    a.dst = dst.data_ptr()
    a.Dd, a.Hd, a.Wd = dst.shape[-3], dst.shape[-2], dst.shape[-1]
    a.dst_origin = _p(dst_origin)
    a.dst_single = 1 if dst_single else 0
    a.flip, a.accumulate, a.acc_scale = flip, 1 if accumulate else 0, float(acc_scale)
    if threshold is not None:
        a.use_threshold, a.threshold = 1, float(threshold)
    a.round_half = 1 if round_half else 0
    L.check(L.lib().e3b_head(ctypes.byref(a), _stream()), 'head')
    return dst


def forward(net, x, training, save):
    _require_cuda(x, 'input')
    # the kernels are launched on the CURRENT device's current stream: make that the tensor's device
    # (model.to('cuda:1') without torch.cuda.set_device(1), nn.DataParallel replica threads)
    with torch.cuda.device(x.device):
        if training:
            net.cache.invalidate()       # eval-mode images (folded BatchNorm statistics) are stale after this pass
        feat, tape = forward_features(net, x, training, save)
        logits = head(feat, net.final)
    if tape.squeeze:
        logits = logits.squeeze(2)
    return logits, tape


# ------------------------------------------------------------------------------------------ backward
def _norm_bwd(u, C, g0, g1=None, gp=None, s2d=None, want_bias=True):
    """Backward of norm -> relu [-> pool] for unit `u`; returns (dy QP or s2d QP, dgamma, dbeta, dbias).
    g1 may be the gradient of a centre-cropped view of this unit's activation (``g1.crop_off`` set by
    _conv_unit_bwd): it is then added inside that box only (the backward of autocrop's slice is a zero pad)."""
    if u.mode == MODE_BATCH_EVAL:
        raise NotImplementedError('backward through eval-mode BatchNorm is not on the accelerated path '
                                  '(call model.train() or wrap evaluation in torch.no_grad())')
    a = u.a
    dev = a.t.device
    N, Cp = a.N, cpad8(C)
    args = L.NormBwdArgs()
    args.y = u.y.ptr
    if u.nstate is not None:
        args.scale, args.shift = u.nstate.scale.data_ptr(), u.nstate.shift.data_ptr()
    args.g0, args.g1, args.gp = (g0.ptr if g0 is not None else None, g1.ptr if g1 is not None else None,
                                 gp.ptr if gp is not None else None)
    args.N, args.C, args.D, args.H, args.W = N, C, a.D, a.H, a.W
    if g1 is not None and (g1.crop_off is not None or g1.spatial != a.spatial):
        args.g1_crop = 1
        args.g1_od, args.g1_oh, args.g1_ow = g1.crop_off or (0, 0, 0)
        args.g1_D, args.g1_H, args.g1_W = g1.spatial
    if gp is not None:
        args.pk_d, args.pk_h, args.pk_w = u.pool
        if u.pooled is None or u.pooled.idx is None:
            raise RuntimeError('norm backward: the forward pass kept no pooling indices')
        args.pool_idx = u.pooled.idx.data_ptr()
    args.mode, args.G = u.mode, u.G
    n = u.spec.norm
    args.eps = float(getattr(n, 'eps', 0.0) or 0.0)
    gamma = _affine(n, 'weight') if u.mode != MODE_NONE else None
    args.gamma = _p(gamma)
    if u.nstate is None:       # mode none: identity statistics
        ident = torch.zeros((2, N, Cp), dtype=torch.float32, device=dev)
        ident[1].fill_(1.0)
        mean, rstd = ident[0], ident[1]
    else:
        mean, rstd = u.nstate.mean, u.nstate.rstd
    args.mean, args.rstd = mean.data_ptr(), rstd.data_ptr()
    args.fwd_stats = _p(u.stats)
    # sums (fp64), amax (u32), dy_scale (4 floats) back to back: e3b_norm_bwd_reduce clears them with one memset
    acc = torch.empty((N * Cp * 2 * 8 + N * Cp * 2 * 4 + 16,), dtype=torch.uint8, device=dev)
    sums = acc[:N * Cp * 16].view(torch.float64)
    amax = acc[N * Cp * 16:N * Cp * 24].view(torch.int32)
    dy_scale = acc[N * Cp * 24:].view(torch.float32)
    m = torch.empty((2, N, Cp), dtype=torch.float32, device=dev)
    args.sums, args.m1, args.m2 = sums.data_ptr(), m[0].data_ptr(), m[1].data_ptr()
    args.amax, args.dy_scale = amax.data_ptr(), dy_scale.data_ptr()
    pg = torch.empty((3, C), dtype=torch.float32, device=dev)
    has_affine = gamma is not None
    args.dgamma = pg[0].data_ptr() if has_affine else None
    args.dbeta = pg[1].data_ptr() if has_affine else None
    args.dbias = pg[2].data_ptr() if want_bias else None
    if s2d is not None:
        args.s2d = 1
        args.sd, args.sh, args.sw = s2d
        nsl = s2d[0] * s2d[1] * s2d[2]
        Dw, Hw, Ww = -(-a.D // s2d[0]), -(-a.H // s2d[1]), -(-a.W // s2d[2])
        dy = QP.empty_half(N, nsl * Cp, Dw, Hw, Ww, dev)
    else:
        dy = QP.empty_half(N, C, a.D, a.H, a.W, dev)
    dy.scale = dy_scale            # the tensor holds 2^k * dy; conv_forward undoes it through dy_scale[2]
    args.dy = dy.ptr
    act = tuple(getattr(u, 'act', ACT_RELU)) + (None,)
    args.relu, args.act_slope = act[0], act[1]
    dslope = None
    if act[2] is not None:                     # nn.PReLU: the slope is read from device memory; its gradient is one more sum
        slope_ws = torch.empty((N * Cp,), dtype=torch.float64, device=dev)
        dslope = torch.empty((1,), dtype=torch.float32, device=dev)
        args.act_slope_dev, args.slope_sums, args.dslope = act[2].detach().data_ptr(), slope_ws.data_ptr(), dslope.data_ptr()
    u.dslope = dslope
    lib = L.lib()
    st = _stream()
    # (PReLU's slope gradient lives in the three-kernel path only; SiLU is a compile-time variant of the persistent kernel)
    fused = Cp <= 512 and dslope is None and os.environ.get('E3B_NORM_BWD', 'fused') != 'split'
    if s2d is not None and (a.D % s2d[0] or a.H % s2d[1] or a.W % s2d[2]):
        fused = False              # autocrop dropped fine voxels: the three-kernel path zero-fills them
    if fused:
        # one persistent kernel: per sample reduce -> grid barrier -> apply, the second read served by L2
        L.check(lib.e3b_norm_bwd_fused(ctypes.byref(args), st), 'norm_bwd_fused')
    else:
        L.check(lib.e3b_norm_bwd_reduce(ctypes.byref(args), st), 'norm_bwd_reduce')
        L.check(lib.e3b_norm_bwd_finalize(ctypes.byref(args), st), 'norm_bwd_finalize')
        L.check(lib.e3b_norm_bwd_apply(ctypes.byref(args), st), 'norm_bwd_apply')
    return dy, (pg[0] if has_affine else None), (pg[1] if has_affine else None), (pg[2] if want_bias else None)


def _put(grads, param, val):
    if param is not None and param.requires_grad:
        grads[id(param)] = val.view(param.shape) if val.shape != param.shape else val


def _conv_unit_bwd(net, u, g0, g1, gp, grads, need_dx):
    """Backward of one conv -> norm -> relu [-> pool] unit: returns (dsrc0, dsrc1)."""
    spec = u.spec
    conv, n = spec.conv, spec.norm
    dy, dgamma, dbeta, dbias = _norm_bwd(u, spec.Co, g0, g1, gp, want_bias=conv.bias is not None)
    if dgamma is not None:
        _put(grads, n.weight, dgamma)
        _put(grads, n.bias, dbeta)
    if conv.bias is not None:
        _put(grads, conv.bias, dbias)
    if u.dslope is not None:
        _put(grads, spec.act.weight, u.dslope)
    if getattr(u, 'res', None) is not None:
        _shortcut_bwd(net, u, dy, dbias, grads)
    cropped = u.src1 is not None and (tuple(u.off1) != (0, 0, 0) or u.src1.spatial != u.src0.spatial)
    if conv.weight.requires_grad:
        dw = wgrad(u.src0, dy, spec.Co, spec.k, spec.pad, tuple(conv.weight.shape), src1=u.src1, off1=u.off1)
        _put(grads, conv.weight, dw)
    if not need_dx:
        return None, None
    dvar = spec.variants[1]
    wsc = net.wset.scales[spec.name]
    if net.wset.images is not None:
        wpk = net.wset.images[(spec.name, 'bwd')]
    else:
        wpk = pack_weights(5 if dvar else 1, conv.weight, None, spec.C0, spec.C1, spec.Co, spec.k, wscale=wsc)
    dpad = tuple(kk - 1 - pp for kk, pp in zip(spec.k, spec.pad))
    d0, d1, _ = conv_forward(dy, wpk, spec.n_total_dgrad, spec.C0, spec.k, dpad,
                             dst1_C=spec.C1 if u.src1 is not None else 0, variant=dvar, w_unscale=wsc)
    if cropped:
        d1.crop_off = tuple(u.off1)       # gradient of autocrop's slice of the skip tensor (models/unet.py:303-324)
    return d0, d1


def _shortcut_bwd(net, u, dy, dbias, grads):
    """gradient of the shortcut branch of a resunet ConvBlock: parameters of the projection, and u.res_grads = what has to
    be added to the gradient of the block input (the scaled fp16 dy itself for the identity shortcut)."""
    kind, r0, r1, roff = u.res
    if not isinstance(kind, ConvSpec):
        u.res_grads = ('dy', dy)
        return
    pc = kind.conv
    if pc.bias is not None and dbias is not None:
        _put(grads, pc.bias, dbias.clone())        # both biases are added to the same y: the same gradient
    if pc.weight.requires_grad:
        _put(grads, pc.weight, wgrad(r0, dy, kind.Co, kind.k, kind.pad, tuple(pc.weight.shape), src1=r1, off1=roff))
    wsc = net.wset.scales[kind.name]
    if net.wset.images is not None:
        wpk = net.wset.images[(kind.name, 'bwd')]
    else:
        wpk = pack_weights(5 if kind.variants[1] else 1, pc.weight, None, kind.C0, kind.C1, kind.Co, kind.k, wscale=wsc)
    p0, p1, _ = conv_forward(dy, wpk, kind.n_total_dgrad, kind.C0, kind.k, (0, 0, 0),
                             dst1_C=kind.C1 if r1 is not None else 0, variant=kind.variants[1], w_unscale=wsc)
    u.res_grads = ('qp', p0, p1)


def _axpy(dst, src, alpha=None):
    L.check(L.lib().e3b_qp_axpy(dst.ptr, src.ptr, 1 if src.half else 0, alpha, dst.N, dst.C, dst.D * dst.H * dst.W, _stream()),
            'qp_axpy')


def _blocks_bwd(net, units, g0, g1, gp, grads, need_dx):
    """backward of a stack of Blocks; (g0, g1, gp) arrive at the last block's output.  -> gradients w.r.t. the two sources
    of the first block"""
    d0 = d1 = None
    for bi in range(len(units) - 1, -1, -1):
        u1, u2 = units[bi]
        ga, _ = _conv_unit_bwd(net, u2, g0, g1, gp, grads, True)
        need = bi > 0 or need_dx
        d0, d1 = _conv_unit_bwd(net, u1, ga, None, None, grads, need)
        rg = u2.res_grads
        if rg is not None and need:
            if rg[0] == 'dy':
                _axpy(d0, rg[1], rg[1].scale.data_ptr() + 8)       # the tensor holds 2^k * dy: times dy_scale[2] = 2^-k
            else:
                _axpy(d0, rg[1])
                if d1 is not None:
                    _axpy(d1, rg[2])
        u2.res_grads = None
        g0, g1, gp = d0, None, None
    return d0, d1


def backward(net, tape, dlogits, need_dx):
    """-> (dict id(param) -> grad tensor, dx or None)"""
    _require_cuda(dlogits, 'grad_output')
    with torch.cuda.device(dlogits.device):
        defer = os.environ.get('E3B_WGRAD_DEFER', '1') != '0'
        _wg_defer.jobs = [] if defer else None
        try:
            out = _backward(net, tape, dlogits, need_dx)
            flush_wgrad()
            return out
        finally:
            _wg_defer.jobs = None


def _backward(net, tape, dlogits, need_dx):
    grads = {}
    net.wset = tape.wset                # the weight scales of the forward this tape belongs to
    dl = dlogits.unsqueeze(2) if tape.squeeze else dlogits
    dl = dl.contiguous()
    feat = tape.final_in
    fc = net.final
    Co = fc.weight.shape[0]
    dev = dl.device
    da = QP.empty(feat.N, feat.C, feat.D, feat.H, feat.W, dev)
    dw = torch.empty((Co, feat.C), dtype=torch.float32, device=dev)
    db = torch.empty((Co,), dtype=torch.float32, device=dev)
    ws = torch.empty((Co * cpad8(feat.C) + Co,), dtype=torch.float64, device=dev)
    L.check(L.lib().e3b_head_bwd(dl.data_ptr(), feat.ptr, fc.weight.detach().data_ptr(), da.ptr, dw.data_ptr(),
                                 db.data_ptr(), ws.data_ptr(), feat.N, feat.C, Co, feat.D, feat.H, feat.W, _stream()),
            'head_bwd')
    _put(grads, fc.weight, dw)
    _put(grads, fc.bias, db)

    g = da
    skip = {}
    for (u0, units, enc_index, add_info), (ups, blocks) in zip(reversed(tape.up), reversed(net.up)):
        du, denc = _blocks_bwd(net, units, g, None, None, grads, True)
        if add_info is not None:
            # merge_mode='add': the gradient of the sum goes to both summands; the skip tensor sees it through autocrop's slice
            denc = QP(du.t, du.N, du.C, du.D, du.H, du.W)
            if tuple(add_info[0]) != (0, 0, 0) or add_info[1] != du.spatial:
                denc.crop_off = tuple(add_info[0])
        skip[enc_index] = denc
        if isinstance(ups, ResizeSpec):
            dup, _ = _conv_unit_bwd(net, u0, du, None, None, grads, True)
            g = _resize_bwd(u0, dup)
            continue
        # norm0/act0 backward, written space-to-depth for the transposed conv's GEMMs
        up = ups.up
        dy, dgamma, dbeta, dbias = _norm_bwd(u0, ups.Co, du, s2d=ups.s, want_bias=up.bias is not None)
        if dgamma is not None:
            _put(grads, ups.norm.weight, dgamma)
            _put(grads, ups.norm.bias, dbeta)
        if up.bias is not None:
            _put(grads, up.bias, dbias)
        if u0.dslope is not None:
            _put(grads, ups.act.weight, u0.dslope)
        dec = u0.dec
        if up.weight.requires_grad:
            dwu = wgrad(dec, dy, dy.C, (1, 1, 1), (0, 0, 0), tuple(up.weight.shape), layout=1, up_taps=ups.taps,
                        up_co=ups.Co)
            _put(grads, up.weight, dwu)
        wsc = net.wset.scales[ups.name]
        if net.wset.images is not None:
            wpk = net.wset.images[(ups.name, 'bwd')]
        else:
            wpk = pack_weights(3, up.weight, None, ups.Ci, 0, ups.Co, ups.s, wscale=wsc)
        g, _, _ = conv_forward(dy, wpk, cpad16(ups.Ci), ups.Ci, (1, 1, 1), (0, 0, 0), w_unscale=wsc)
    nd = len(net.down)
    dx = None
    for i in range(nd - 1, -1, -1):
        units = tape.down[i]
        # (the skip gradient always travels as g1: it may be the gradient of a centre-cropped view)
        if units[-1][1].pool is not None:
            g, _ = _blocks_bwd(net, units, None, skip.get(i), g, grads, i > 0 or need_dx)
        else:
            g, _ = _blocks_bwd(net, units, g, skip.get(i), None, grads, i > 0 or need_dx)
    if need_dx:
        dx = unpack(g)
        if tape.squeeze:
            dx = dx.squeeze(2)
    return grads, dx
