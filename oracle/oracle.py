"""CPU oracle for the elektronn3 UNet / Predictor hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``elektronn3_b200``) never does.

It composes the plain-C operator restatements of ``oracle/e3_oracle.c`` (numpy +
ctypes, no torch) into the reference's network structure:

* ``UNetOracle``      restates ``elektronn3/models/unet.py``: channel plan
  (:840-881), ``DownConv.forward`` (:244-253), ``autocrop`` (:256-325),
  ``UpConv.forward`` (:384-408), ``UNet.forward`` (:894-916) and the backward
  that torch autograd derives from them.
* ``tiled_apply``     restates ``elektronn3/inference/inference.py:45-199``.
* ``predictor_apply`` restates the model wrapping of ``Predictor``
  (``inference.py:443-458,496-525``): softmax(1) -> optional argmax -> crop.

Parity pin: the reference has no golden vectors for this path, so the oracle is
pinned against outputs of the reference itself (``oracle/gen_golden.py`` ->
``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).
"""
import ctypes
import itertools
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libe3oracle.so')
_SRC = os.path.join(_HERE, 'e3_oracle.c')

_f = ctypes.POINTER(ctypes.c_float)
_i64 = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile oracle/e3_oracle.c -> oracle/libe3oracle.so (gcc, OpenMP)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(['gcc', '-O2', '-fopenmp', '-shared', '-fPIC', '-o', _SO, _SRC, '-lm'])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(_f)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _as5(x):
    """(N,C,H,W) -> (N,C,1,H,W) view so that 2D nets run through the 3D operators."""
    return x[:, :, None] if x.ndim == 4 else x


# ----------------------------------------------------------------------------- operators
def conv_fwd(x, w, b, pad):
    x, w = _c(x), _c(w)
    N, Ci, D, H, W = x.shape
    Co, _, kd, kh, kw = w.shape
    y = np.empty((N, Co, D + 2 * pad[0] - kd + 1, H + 2 * pad[1] - kh + 1, W + 2 * pad[2] - kw + 1), np.float32)
    lib().e3o_conv3d_fwd(_p(x), _p(w), _p(_c(b)) if b is not None else None, _p(y),
                         N, Ci, D, H, W, Co, kd, kh, kw, *pad)
    return y


def conv_bwd(x, w, dy, pad, need_dx=True):
    x, w, dy = _c(x), _c(w), _c(dy)
    N, Ci, D, H, W = x.shape
    Co, _, kd, kh, kw = w.shape
    dx = np.empty_like(x) if need_dx else None
    dw = np.empty_like(w)
    db = np.empty((Co,), np.float32)
    lib().e3o_conv3d_bwd(_p(x), _p(w), _p(dy), _p(dx), _p(dw), _p(db), N, Ci, D, H, W, Co, kd, kh, kw, *pad)
    return dx, dw, db


def convT_fwd(x, w, b):
    x, w = _c(x), _c(w)
    N, Ci, D, H, W = x.shape
    _, Co, sd, sh, sw = w.shape
    y = np.empty((N, Co, D * sd, H * sh, W * sw), np.float32)
    lib().e3o_convT_fwd(_p(x), _p(w), _p(_c(b)), _p(y), N, Ci, D, H, W, Co, sd, sh, sw)
    return y


def convT_bwd(x, w, dy):
    x, w, dy = _c(x), _c(w), _c(dy)
    N, Ci, D, H, W = x.shape
    _, Co, sd, sh, sw = w.shape
    dx, dw, db = np.empty_like(x), np.empty_like(w), np.empty((Co,), np.float32)
    lib().e3o_convT_bwd(_p(x), _p(w), _p(dy), _p(dx), _p(dw), _p(db), N, Ci, D, H, W, Co, sd, sh, sw)
    return dx, dw, db


def groupnorm_fwd(x, gamma, beta, G, eps=1e-5):
    x = _c(x)
    N, C = x.shape[:2]
    S = int(np.prod(x.shape[2:]))
    y = np.empty_like(x)
    lib().e3o_groupnorm_fwd(_p(x), _p(gamma), _p(beta), _p(y), None, None, N, C, ctypes.c_int64(S), G,
                            ctypes.c_float(eps))
    return y


def groupnorm_bwd(x, gamma, dy, G, eps=1e-5):
    x, dy = _c(x), _c(dy)
    N, C = x.shape[:2]
    S = int(np.prod(x.shape[2:]))
    dx, dg, dbt = np.empty_like(x), np.empty((C,), np.float32), np.empty((C,), np.float32)
    lib().e3o_groupnorm_bwd(_p(x), _p(gamma), _p(dy), _p(dx), _p(dg), _p(dbt), N, C, ctypes.c_int64(S), G,
                            ctypes.c_float(eps))
    return dx, dg, dbt


def batchnorm_fwd(x, gamma, beta, rm, rv, training, eps=1e-5, momentum=0.1):
    """rm / rv are updated IN PLACE when training (like torch)."""
    x = _c(x)
    N, C = x.shape[:2]
    S = int(np.prod(x.shape[2:]))
    y = np.empty_like(x)
    lib().e3o_batchnorm_fwd(_p(x), _p(gamma), _p(beta), _p(rm), _p(rv), _p(y), None, None, N, C,
                            ctypes.c_int64(S), ctypes.c_float(eps), ctypes.c_float(momentum), int(training))
    return y


def batchnorm_bwd(x, gamma, dy, eps=1e-5):
    x, dy = _c(x), _c(dy)
    N, C = x.shape[:2]
    S = int(np.prod(x.shape[2:]))
    dx, dg, dbt = np.empty_like(x), np.empty((C,), np.float32), np.empty((C,), np.float32)
    lib().e3o_batchnorm_bwd(_p(x), _p(gamma), _p(dy), _p(dx), _p(dg), _p(dbt), N, C, ctypes.c_int64(S),
                            ctypes.c_float(eps))
    return dx, dg, dbt


def maxpool_fwd(x, k):
    x = _c(x)
    N, C, D, H, W = x.shape
    Do, Ho, Wo = -(-D // k[0]), -(-H // k[1]), -(-W // k[2])
    y = np.empty((N, C, Do, Ho, Wo), np.float32)
    idx = np.empty((N, C, Do, Ho, Wo), np.int64)
    lib().e3o_maxpool_fwd(_p(x), _p(y), idx.ctypes.data_as(_i64), N, C, D, H, W, *k)
    return y, idx


def maxpool_bwd(dy, idx, xshape, k):
    dy = _c(dy)
    N, C, D, H, W = xshape
    dx = np.empty(xshape, np.float32)
    lib().e3o_maxpool_bwd(_p(dy), idx.ctypes.data_as(_i64), _p(dx), N, C, D, H, W, *k)
    return dx


def softmax1(x):
    m = x.max(axis=1, keepdims=True)
    e = np.exp((x - m).astype(np.float64))
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


# ----------------------------------------------------------------------------- UNet
def parse_norm(normalization, C):
    """get_normalization models/unet.py:77-111 -> ('none'|'group'|'batch', groups)."""
    if normalization is None or normalization == 'none':
        return 'none', 0
    if normalization.startswith('group'):
        return 'group', (8 if normalization == 'group' else int(normalization[len('group'):]))
    if normalization == 'instance':
        return 'instance', C
    if normalization == 'batch':
        return 'batch', 0
    raise ValueError(normalization)


class _Norm:
    """norm -> ReLU unit (unet.py:246-247 etc.) with cached tensors for backward."""

    def __init__(self, sd, prefix, normalization, C, training):
        self.kind, self.G = parse_norm(normalization, C)
        self.sd, self.prefix, self.training = sd, prefix, training
        self.g = sd.get(prefix + '.weight')
        self.b = sd.get(prefix + '.bias')

    def fwd(self, x):
        self.x = x
        if self.kind == 'none':
            y = x
        elif self.kind in ('group', 'instance'):
            y = groupnorm_fwd(x, self.g, self.b, self.G)
        else:
            rm, rv = self.sd[self.prefix + '.running_mean'], self.sd[self.prefix + '.running_var']
            y = batchnorm_fwd(x, self.g, self.b, rm, rv, self.training)
            if self.training and (self.prefix + '.num_batches_tracked') in self.sd:
                self.sd[self.prefix + '.num_batches_tracked'] += 1
        self.a = np.maximum(y, 0)
        return self.a

    def bwd(self, da, grads):
        dy = np.where(self.a > 0, da, 0).astype(np.float32)
        if self.kind == 'none':
            return dy
        if self.kind in ('group', 'instance'):
            dx, dg, db = groupnorm_bwd(self.x, self.g, dy, self.G)
        else:
            assert self.training, 'eval-mode BN backward is not on the reference hot path'
            dx, dg, db = batchnorm_bwd(self.x, self.g, dy)
        if self.g is not None:
            grads[self.prefix + '.weight'] = dg
            grads[self.prefix + '.bias'] = db
        return dx


class UNetOracle:
    """Restatement of elektronn3.models.unet.UNet (transpose up-mode, concat merge,
    ReLU, SAME or VALID convs, group/batch/instance/none normalisation, planar
    blocks, dim 2 or 3) over numpy state_dict arrays (reference key names)."""

    def __init__(self, state_dict, in_channels=1, out_channels=2, n_blocks=3, start_filts=32,
                 planar_blocks=(), normalization='batch', full_norm=True, dim=3, conv_mode='same',
                 training=False):
        self.sd = {k: (np.array(v, copy=True) if np.ndim(v) == 0 or v.dtype != np.float32
                       else np.ascontiguousarray(v, dtype=np.float32).copy()) for k, v in state_dict.items()}
        self.n_blocks, self.planar_blocks, self.dim = n_blocks, tuple(planar_blocks), dim
        self.normalization, self.full_norm, self.training = normalization, full_norm, training
        self.pad1 = 1 if 'same' in conv_mode else 0
        self.chans = [start_filts * 2 ** i for i in range(n_blocks)]

    # -- helpers
    def _w(self, key):
        w = self.sd[key]
        return w[:, :, None] if w.ndim == 4 else w  # 2D kernels -> depth-1 3D kernels

    def _pad(self, planar):
        p = self.pad1
        return (0, p, p) if (planar or self.dim == 2) else (p, p, p)

    def _k2(self, planar):
        return (1, 2, 2) if (planar or self.dim == 2) else (2, 2, 2)

    def _norm(self, prefix, C, enabled=True):
        return _Norm(self.sd, prefix, self.normalization if enabled else 'none', C, self.training)

    @staticmethod
    def _autocrop(enc, up):
        """unet.py:256-325.  Returns cropped views and the slices used."""
        ds, us = enc.shape[2:], up.shape[2:]
        if ds == us:
            return enc, up, None, None
        upcrop = [u - ((u - d) % 2) for d, u in zip(ds, us)]
        us_sl = (slice(None), slice(None)) + tuple(slice(0, c) for c in upcrop)
        up = up[us_sl]
        us = up.shape[2:]
        en_sl = (slice(None), slice(None)) + tuple(slice((d - u) // 2, (d + u) // 2) for d, u in zip(ds, us))
        return enc[en_sl], up, en_sl, us_sl

    # -- forward (unet.py:894-916)
    def forward(self, x):
        squeeze = x.ndim == 4
        x = _c(_as5(x))
        self.tape = []
        enc = []
        for i in range(self.n_blocks):
            planar = i in self.planar_blocks
            C = self.chans[i]
            p = f'down_convs.{i}'
            n0, n1 = self._norm(p + '.norm0', C, self.full_norm), self._norm(p + '.norm1', C)
            x_in = x
            y = n0.fwd(conv_fwd(x_in, self._w(p + '.conv1.weight'), self.sd[p + '.conv1.bias'], self._pad(planar)))
            y2 = n1.fwd(conv_fwd(y, self._w(p + '.conv2.weight'), self.sd[p + '.conv2.bias'], self._pad(planar)))
            enc.append(y2)
            rec = dict(kind='down', p=p, planar=planar, x_in=x_in, a1=y, n0=n0, n1=n1, pool=None)
            if i < self.n_blocks - 1:
                x, idx = maxpool_fwd(y2, self._k2(planar))
                rec['pool'] = (idx, y2.shape)
            else:
                x = y2
            self.tape.append(rec)
        for i in range(self.n_blocks - 1):
            planar = (self.n_blocks - 2 - i) in self.planar_blocks
            C = self.chans[self.n_blocks - 2 - i]
            p = f'up_convs.{i}'
            before_pool = enc[-(i + 2)]
            dec = x
            up_full = convT_fwd(dec, self._w(p + '.upconv.weight'), self.sd[p + '.upconv.bias'])
            e, up, en_sl, us_sl = self._autocrop(before_pool, up_full)
            n0 = self._norm(p + '.norm0', C, self.full_norm)
            n1 = self._norm(p + '.norm1', C, self.full_norm)
            n2 = self._norm(p + '.norm2', C)
            u = n0.fwd(_c(up))
            mrg = np.concatenate((u, e), axis=1)           # unet.py:399: (updec, enc)
            y1 = n1.fwd(conv_fwd(mrg, self._w(p + '.conv1.weight'), self.sd[p + '.conv1.bias'], self._pad(planar)))
            y2 = n2.fwd(conv_fwd(y1, self._w(p + '.conv2.weight'), self.sd[p + '.conv2.bias'], self._pad(planar)))
            self.tape.append(dict(kind='up', p=p, planar=planar, dec=dec, up_shape=up_full.shape, en_sl=en_sl,
                                  us_sl=us_sl, enc_index=self.n_blocks - 2 - i, enc_shape=before_pool.shape,
                                  mrg=mrg, a1=y1, n0=n0, n1=n1, n2=n2, C=C))
            x = y2
        self.final_in = x
        out = conv_fwd(x, self._w('conv_final.weight'), self.sd['conv_final.bias'], (0, 0, 0))
        return out[:, :, 0] if squeeze else out

    # -- backward (autograd of the above)
    def backward(self, dout):
        dout = _c(_as5(dout))
        g = {}

        def put(key, val):
            ref = self.sd[key]
            g[key] = val.reshape(ref.shape)

        dx, dw, db = conv_bwd(self.final_in, self._w('conv_final.weight'), dout, (0, 0, 0))
        put('conv_final.weight', dw), put('conv_final.bias', db)
        d_enc = [None] * self.n_blocks     # gradient flowing into each before_pool via the skip
        for rec in reversed(self.tape):
            p = rec['p']
            if rec['kind'] == 'up':
                pad = self._pad(rec['planar'])
                dy2 = rec['n2'].bwd(dx, g)
                da1, dw, db = conv_bwd(rec['a1'], self._w(p + '.conv2.weight'), dy2, pad)
                put(p + '.conv2.weight', dw), put(p + '.conv2.bias', db)
                dy1 = rec['n1'].bwd(da1, g)
                dmrg, dw, db = conv_bwd(rec['mrg'], self._w(p + '.conv1.weight'), dy1, pad)
                put(p + '.conv1.weight', dw), put(p + '.conv1.bias', db)
                C = rec['C']
                du, de = dmrg[:, :C], dmrg[:, C:]
                full = np.zeros(rec['enc_shape'], np.float32)
                if rec['en_sl'] is None:
                    full += de
                else:
                    full[rec['en_sl']] += de
                d_enc[rec['enc_index']] = full
                dup = rec['n0'].bwd(_c(du), g)
                dup_full = np.zeros(rec['up_shape'], np.float32)
                if rec['us_sl'] is None:
                    dup_full += dup
                else:
                    dup_full[rec['us_sl']] = dup
                dx, dw, db = convT_bwd(rec['dec'], self._w(p + '.upconv.weight'), dup_full)
                put(p + '.upconv.weight', dw), put(p + '.upconv.bias', db)
            else:
                i = int(p.split('.')[1])
                pad = self._pad(rec['planar'])
                if rec['pool'] is not None:
                    idx, shp = rec['pool']
                    dx = maxpool_bwd(dx, idx, shp, self._k2(rec['planar']))
                if d_enc[i] is not None:
                    dx = dx + d_enc[i]
                dy2 = rec['n1'].bwd(dx, g)
                da1, dw, db = conv_bwd(rec['a1'], self._w(p + '.conv2.weight'), dy2, pad)
                put(p + '.conv2.weight', dw), put(p + '.conv2.bias', db)
                dy1 = rec['n0'].bwd(da1, g)
                dx, dw, db = conv_bwd(rec['x_in'], self._w(p + '.conv1.weight'), dy1, pad, need_dx=(i > 0))
                put(p + '.conv1.weight', dw), put(p + '.conv1.bias', db)
        return g


# ----------------------------------------------------------------------------- tiled inference
def tiled_apply(func, inp, tile_shape, overlap_shape, offset, out_shape):
    """inference.py:45-199 (SAME nets: pad by overlap, crop the centre; VALID nets
    (offset given): input is pre-padded, no crop).  ``func(tile, crop_slices)``."""
    tile_shape, overlap_shape = np.array(tile_shape), np.array(overlap_shape)
    out_shape = np.array(out_shape)
    if not np.all(np.mod(out_shape[2:], tile_shape) == 0):
        raise ValueError('out_shape not divisible by tile_shape')
    crop = None
    if np.array_equal(out_shape[2:], np.array(inp.shape[2:])):
        padded = np.zeros(tuple(np.array(inp.shape) + np.array((0, 0, *overlap_shape * 2))), inp.dtype)
        padded[(slice(None), slice(None)) + tuple(slice(l, h) for l, h in
                                                   zip(overlap_shape, np.array(padded.shape[2:]) - overlap_shape))] = inp
        crop = (slice(None), slice(None)) + tuple(slice(l, h) for l, h in
                                                   zip(overlap_shape, tile_shape + overlap_shape))
    else:
        padded = inp
    if offset is not None:
        crop = None
    out = None
    tiles = np.ceil(out_shape[2:] / tile_shape).astype(int)
    for pos in itertools.product(*[range(t) for t in tiles]):
        pos = np.array(pos)
        lo, hi = tile_shape * pos, tile_shape * (pos + 1)
        isl = (slice(None), slice(None)) + tuple(slice(l, h) for l, h in zip(lo, hi + 2 * overlap_shape))
        osl = (slice(None), slice(None)) + tuple(slice(l, h) for l, h in zip(lo, hi))
        t = func(np.ascontiguousarray(padded[isl]), crop)
        if out is None:
            out = np.empty(tuple(out_shape), t.dtype)
        out[osl] = t
    return out


def predictor_apply(net, tile, crop, apply_softmax=True, apply_argmax=False, out_dtype=None):
    """Predictor model wrapping + _predict (inference.py:443-458, 496-525), no TTA."""
    o = net.forward(tile)
    if apply_softmax:
        o = softmax1(o)
    if apply_argmax:
        o = np.argmax(o, axis=1)[:, None]
        out_dtype = out_dtype or np.uint8
    if crop is not None:
        o = o[crop]
    return o.astype(out_dtype) if out_dtype is not None else o
