"""Plain-torch functional restatement of the reference UNet.forward (models/unet.py:244-253, 384-408,
894-916) over the parameters of an elektronn3_b200.UNet -- TEST INFRASTRUCTURE.

Used on the GPU box (where /root/reference does not exist) to measure the TF32 noise floor: the same
network evaluated by torch/cuDNN in fp32 and in TF32 (the reference's own default GPU arithmetic).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _norm(n, t):
    return t if isinstance(n, nn.Identity) else n(t)


def tf32_round(t):
    """round-to-nearest (ties away, like cvt.rna) fp32 -> tf32, as a float32 tensor"""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class _RoundSTE(torch.autograd.Function):
    """forward: round to TF32; backward: identity (the kernels do not differentiate the rounding)"""

    @staticmethod
    def forward(ctx, t):
        return tf32_round(t)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):
    """forward: identity; backward: round the gradient to TF32 (dy is stored rounded: it is an MMA operand)"""

    @staticmethod
    def forward(ctx, t):
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return tf32_round(g)


def unet_forward(m, x, emulate=False):
    """emulate=True reproduces the ARITHMETIC of the sm_100a kernels in fp32 torch ops: every MMA operand
    (network input, activations, weights, conv-output gradients) is rounded to TF32 where the kernels store
    it rounded, products are then exact and accumulation is fp32 -- so the result should agree with
    libe3b to ~1e-5 (logits) and isolates implementation bugs from TF32 noise."""
    conv_ = F.conv3d if m.dim == 3 else F.conv2d
    convT_ = F.conv_transpose3d if m.dim == 3 else F.conv_transpose2d
    pool = F.max_pool3d if m.dim == 3 else F.max_pool2d
    rs = _RoundSTE.apply if emulate else (lambda t: t)
    rg = _RoundGrad.apply if emulate else (lambda t: t)

    def conv(t, w, b, **kw):
        return rg(conv_(t, rs(w), b, **kw))

    def convT(t, w, b, **kw):
        return rg(convT_(t, rs(w), b, **kw))

    def act(t):
        return rs(F.relu(t))
    x = rs(x)
    enc = []
    for b in m.down_convs:
        y = act(_norm(b.norm0, conv(x, b.conv1.weight, b.conv1.bias, padding=b.conv1.padding)))
        y = act(_norm(b.norm1, conv(y, b.conv2.weight, b.conv2.bias, padding=b.conv2.padding)))
        enc.append(y)
        x = pool(y, b.pool.kernel_size, ceil_mode=True) if b.pooling else y
    for i, b in enumerate(m.up_convs):
        e = enc[-(i + 2)]
        u = convT(x, b.upconv.weight, b.upconv.bias, stride=b.upconv.stride)
        # autocrop (models/unet.py:256-325)
        ds, us = e.shape[2:], u.shape[2:]
        if ds != us:
            u = u[(slice(None), slice(None)) + tuple(slice(0, a - ((a - d) % 2)) for a, d in zip(us, ds))]
            us = u.shape[2:]
            e = e[(slice(None), slice(None)) + tuple(slice((d - a) // 2, (d + a) // 2) for a, d in zip(us, ds))]
        u = act(_norm(b.norm0, u))
        y = act(_norm(b.norm1, conv(torch.cat((u, e), 1), b.conv1.weight, b.conv1.bias, padding=b.conv1.padding)))
        x = act(_norm(b.norm2, conv(y, b.conv2.weight, b.conv2.bias, padding=b.conv2.padding)))
    return conv_(x, m.conv_final.weight, m.conv_final.bias)      # the 1x1x1 head runs in fp32 on CUDA cores


def grads_with(m, x, dlogits, mode):
    """parameter gradients of the functional forward; mode 'fp32' | 'tf32' (cuDNN conv math) |
    'emulate' (fp32 math on TF32-rounded operands: the kernels' arithmetic)"""
    import copy
    m = copy.deepcopy(m)          # keeps BatchNorm running statistics of the caller untouched
    m.zero_grad()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = mode == 'tf32'
    try:
        out = unet_forward(m, x, emulate=(mode == 'emulate'))
        out.backward(dlogits)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    return out.detach(), {k: p.grad.detach() for k, p in m.named_parameters()}
