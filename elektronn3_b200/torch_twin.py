"""Plain-torch twin of ``elektronn3_b200.UNet`` for EXPORT (TorchScript / ONNX / running a checkpoint where libe3b.so does
not exist).

``Trainer._save_model`` scripts or traces the model when ``save_jit`` is set (training/trainer.py:876-887; the reference's
example scripts default to ``--jit onsave``), and ``Predictor`` loads such ``.pts`` archives (inference.py:418-421).  The
sm_100a kernels cannot live in a TorchScript archive, so for export -- and only for export -- the module hands out this
twin: the same parameter / buffer objects under the same ``state_dict`` keys, and a ``forward`` that issues the ATen call
sequence of the reference network (models/unet.py:244-253 DownConv, :384-408 UpConv, :256-325 autocrop, :894-916 UNet).
``torch.jit.script(model)`` finds it through ``UNet.__prepare_scriptable__``; ``torch.jit.trace`` through the tracing
branch of ``UNet.forward``.  It is never used to compute on the product path.
"""
from typing import List, Tuple

import torch
import torch.nn as nn


def _crop_pair(enc: torch.Tensor, up: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """autocrop: drop one trailing voxel of the up-sampled tensor where the extents differ by an odd number, then cut the
    skip tensor's centre to its shape."""
    nd = up.dim() - 2
    if enc.shape[2:] == up.shape[2:]:
        return enc, up
    us: List[int] = []
    for i in range(nd):
        u, d = up.shape[2 + i], enc.shape[2 + i]
        us.append(u - ((u - d) % 2))
    if nd == 3:
        up = up[:, :, :us[0], :us[1], :us[2]]
    else:
        up = up[:, :, :us[0], :us[1]]
    lo: List[int] = []
    for i in range(nd):
        lo.append((enc.shape[2 + i] - us[i]) // 2)
    if nd == 3:
        enc = enc[:, :, lo[0]:lo[0] + us[0], lo[1]:lo[1] + us[1], lo[2]:lo[2] + us[2]]
    else:
        enc = enc[:, :, lo[0]:lo[0] + us[0], lo[1]:lo[1] + us[1]]
    return enc, up


class TwinDown(nn.Module):
    def __init__(self, b):
        super().__init__()
        self.conv1, self.conv2, self.norm0, self.norm1, self.pool = b.conv1, b.conv2, b.norm0, b.norm1, b.pool
        self.act1, self.act2 = b.act1, b.act2

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        y = self.act1(self.norm0(self.conv1(x)))
        y = self.act2(self.norm1(self.conv2(y)))
        return self.pool(y), y


class TwinUp(nn.Module):
    def __init__(self, b):
        super().__init__()
        self.upconv, self.conv1, self.conv2 = b.upconv, b.conv1, b.conv2
        self.norm0, self.norm1, self.norm2 = b.norm0, b.norm1, b.norm2
        self.act0, self.act1, self.act2 = b.act0, b.act1, b.act2
        self.add = b.merge_mode == 'add'

    def forward(self, enc: torch.Tensor, dec: torch.Tensor) -> torch.Tensor:
        up = self.upconv(dec)
        enc, up = _crop_pair(enc, up)
        up = self.act0(self.norm0(up))
        if self.add:
            mrg = up + enc
        else:
            mrg = torch.cat((up, enc), 1)
        y = self.act1(self.norm1(self.conv1(mrg)))
        return self.act2(self.norm2(self.conv2(y)))


class TwinUNet(nn.Module):
    """Shares every layer object with the ``elektronn3_b200.UNet`` it was built from: same ``state_dict`` keys and
    storage; scriptable and traceable."""

    def __init__(self, unet):
        super().__init__()
        self.down_convs = nn.ModuleList([TwinDown(b) for b in unet.down_convs])
        self.up_convs = nn.ModuleList([TwinUp(b) for b in unet.up_convs])
        self.conv_final = unet.conv_final

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        skips: List[torch.Tensor] = []
        for d in self.down_convs:
            x, before_pool = d(x)
            skips.append(before_pool)
        i = 0
        for u in self.up_convs:
            x = u(skips[-(i + 2)], x)
            i += 1
        return self.conv_final(x)
