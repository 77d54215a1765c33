"""Drop-in for ``elektronn3.modules.loss.DiceLoss`` (modules/loss.py:165-233) whose forward and backward are two fused
CUDA passes over the logits (libe3b.so ``e3b_dice_fwd`` / ``e3b_dice_bwd``) instead of ~25 torch launches over
(N, C, spatial) tensors.  Same constructor, same buffer (``weight``), same value: the generalised Dice loss

    mean_c  w_c * (1 - (2 * sum(p_c * t_c) + smooth) / (sum(p_c) + sum(t_c) + smooth + 1e-4))

with p = softmax(output) (``apply_softmax=True``) and t the one-hot target (dense class-index targets are converted on
the fly, one-hot targets of the output's shape are accepted as they are).  Like the rest of the package it has no CPU
path: the output must be a float32 CUDA tensor.
"""
from typing import Optional

import torch

from . import _lib as L


class _DiceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, weight, apply_softmax, smooth):
        out = output.detach().contiguous()
        N, C = int(out.shape[0]), int(out.shape[1])
        S = out[0, 0].numel()
        dense = tuple(target.shape) != tuple(out.shape)
        if dense:
            if not (target.shape[0] == out.shape[0] and tuple(target.shape[1:]) == tuple(out.shape[2:])):
                raise ValueError(f'Target shape {target.shape} is not compatible with output shape {output.shape}.')
            tgt = target.contiguous() if target.dtype == torch.int64 else target.long().contiguous()
        else:
            tgt = target.to(torch.float32).contiguous()
        dev = out.device
        w = weight.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        if w.numel() not in (1, C):
            raise ValueError(f'weight has to have {C} elements (one per class) or one, got {w.numel()}')
        sums = torch.empty((3 * C,), dtype=torch.float64, device=dev)
        res = torch.empty((1 + 2 * C,), dtype=torch.float32, device=dev)       # loss, then the backward coefficients
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            L.check(L.lib().e3b_dice_fwd(out.data_ptr(), tgt.data_ptr() if dense else None, None if dense else tgt.data_ptr(),
                                         w.data_ptr(), w.numel(), N, C, S, 1 if apply_softmax else 0, float(smooth), 1e-4,
                                         sums.data_ptr(), res.data_ptr(), res.data_ptr() + 4, st), 'dice_fwd')
        ctx.save_for_backward(out, tgt, res)
        ctx.dense, ctx.apply_softmax = dense, apply_softmax
        return res[0].clone()

    @staticmethod
    def backward(ctx, gout):
        out, tgt, res = ctx.saved_tensors
        N, C = int(out.shape[0]), int(out.shape[1])
        S = out[0, 0].numel()
        dx = torch.empty_like(out)
        g = gout.detach().to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(out.device):
            st = torch.cuda.current_stream(out.device).cuda_stream
            L.check(L.lib().e3b_dice_bwd(out.data_ptr(), tgt.data_ptr() if ctx.dense else None,
                                         None if ctx.dense else tgt.data_ptr(), res.data_ptr() + 4, g.data_ptr(), dx.data_ptr(),
                                         N, C, S, 1 if ctx.apply_softmax else 0, st), 'dice_bwd')
        return dx, None, None, None, None


class DiceLoss(torch.nn.Module):
    """Generalized Dice Loss with the constructor of the reference class (modules/loss.py:192-233)."""

    def __init__(self, apply_softmax: bool = True, weight: Optional[torch.Tensor] = None, smooth: float = 0.):
        super().__init__()
        self.apply_softmax = bool(apply_softmax)
        if weight is None:
            weight = torch.tensor(1.)
        self.register_buffer('weight', weight)
        self.smooth = smooth

    def forward(self, output, target):
        if not output.is_cuda or output.dtype != torch.float32:
            raise RuntimeError('elektronn3_b200.DiceLoss expects the float32 CUDA output of the network (no CPU path)')
        if output.shape[1] > 16:
            raise NotImplementedError('elektronn3_b200.DiceLoss supports up to 16 classes')
        return _DiceFunction.apply(output, target, self.weight, self.apply_softmax, self.smooth)
