"""Drop-in replacement for ``elektronn3.models.resunet.UNet`` (models/resunet.py:598-990) on the sm_100a kernels of libe3b.so.

The residual U-Net is the plain one (``elektronn3_b200.unet.UNet``) with stacks of ``ConvBlock``s per level and a shortcut
around each block when ``enc_res_blocks`` / ``dec_res_blocks`` >= 1 (models/resunet.py:212-261):

    y = conv2(act1(norm1(conv1(inp))));  y += proj(inp);  y = act2(norm2(y))

with ``proj`` a 1x1x1 convolution where the channel counts differ and the identity elsewhere.  The convolutions run on the
same tensor-core kernels; the shortcut is ``e3b_residual_add`` (the sum and the statistics of the sum for ``norm2`` in one
pass; the projection is a 1-tap launch of the conv kernel, over the virtual concat in the decoder), its backward
``e3b_qp_axpy``.  Parameters live in ordinary torch layers under the reference's names
(``down_convs.{i}.convs.{j}.conv1|norm1|conv2|norm2|proj``, ``up_convs.{i}.upconv|norm0|convs.{j}...``), so checkpoints
interchange.  The block classes below carry plain-torch ``forward``s: they serve the export twin (TorchScript,
``UNet.torch_twin``) only -- ``UNet.forward`` never calls them.
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import engine
from . import unet as _u
from .torch_twin import _crop_pair


class ConvBlock(nn.Module):
    """models/resunet.py:212-261"""

    def __init__(self, in_channels, out_channels, planar=False, activation='relu', normalization=None, conv_mode='same',
                 residual=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.normalization, self.conv_mode, self.residual, self.dim = normalization, conv_mode, residual, 3
        pad = 1 if conv_mode == 'same' else 0
        k3, p3 = (3, pad) if not planar else ((1, 3, 3), (0, pad, pad))
        self.conv1 = nn.Conv3d(in_channels, out_channels, kernel_size=k3, padding=p3)
        self.norm1 = _u._make_norm(normalization, out_channels, 3)
        self.act1 = _u._make_activation(activation)
        self.conv2 = nn.Conv3d(out_channels, out_channels, kernel_size=k3, padding=p3)
        self.norm2 = _u._make_norm(normalization, out_channels, 3)
        self.act2 = _u._make_activation(activation)
        if residual and in_channels != out_channels:
            self.proj = nn.Conv3d(in_channels, out_channels, kernel_size=1)      # "projection" to match the channel counts
        else:
            self.proj = nn.Identity()

    def forward(self, inp: torch.Tensor) -> torch.Tensor:
        y = self.act1(self.norm1(self.conv1(inp)))
        y = self.conv2(y)
        if self.residual:
            y = y + self.proj(inp)
        return self.act2(self.norm2(y))


class DownBlock(nn.Module):
    """models/resunet.py:264-311"""

    def __init__(self, in_channels, out_channels, pooling=True, planar=False, activation='relu', normalization=None,
                 conv_mode='same', res_blocks=0, skip_first_residual=False):
        super().__init__()
        self.in_channels, self.out_channels, self.pooling = in_channels, out_channels, pooling
        self.normalization, self.res_blocks, self.dim = normalization, res_blocks, 3
        enable_residual = res_blocks >= 1
        convs = [ConvBlock(in_channels, out_channels, planar=planar, activation=activation, normalization=normalization,
                           conv_mode=conv_mode, residual=(enable_residual and not skip_first_residual))]
        for _ in range(res_blocks - 1):
            convs.append(ConvBlock(out_channels, out_channels, planar=planar, activation=activation,
                                   normalization=normalization, conv_mode=conv_mode, residual=enable_residual))
        self.convs = nn.Sequential(*convs)
        if pooling:
            self.pool_ks = (1, 2, 2) if planar else 2
            self.pool = nn.MaxPool3d(kernel_size=self.pool_ks, ceil_mode=True)
        else:
            self.pool_ks = -123
            self.pool = nn.Identity()

    def pool_kernel(self):
        if not self.pooling:
            return None
        ks = self.pool_ks
        return (ks, ks, ks) if isinstance(ks, int) else tuple(ks)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        y = self.convs(x)
        return self.pool(y), y


class DummyAttention(nn.Module):
    """models/resunet.py:592-595 (attention=False): the skip tensor passes unchanged"""

    def forward(self, x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
        return x


class UpBlock(nn.Module):
    """models/resunet.py:386-456"""

    att: Optional[torch.Tensor]

    def __init__(self, in_channels, out_channels, merge_mode='concat', up_mode='transpose', planar=False, activation='relu',
                 normalization=None, conv_mode='same', res_blocks=0):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.merge_mode, self.up_mode, self.normalization, self.res_blocks, self.dim = merge_mode, up_mode, normalization, res_blocks, 3
        enable_residual = res_blocks >= 1
        if up_mode == 'transpose':
            k2 = (1, 2, 2) if planar else 2
            self.upconv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size=k2, stride=k2)
        else:
            self.upconv = _u.ResizeConv(in_channels, out_channels, kernel_size=1 if up_mode.endswith('1') else 3, planar=planar,
                                        dim=3, upsampling_mode='trilinear' if 'linear' in up_mode else 'nearest')
        self.act0 = _u._make_activation(activation)
        self.norm0 = _u._make_norm(normalization, out_channels, 3)
        self.attention = DummyAttention()
        self.att = None
        convs = [ConvBlock(2 * out_channels if merge_mode == 'concat' else out_channels, out_channels, planar=planar,
                           activation=activation, normalization=normalization, conv_mode=conv_mode, residual=enable_residual)]
        for _ in range(res_blocks - 1):
            convs.append(ConvBlock(out_channels, out_channels, planar=planar, activation=activation, normalization=normalization,
                                   conv_mode=conv_mode, residual=enable_residual))
        self.convs = nn.Sequential(*convs)

    def forward(self, enc: torch.Tensor, dec: torch.Tensor) -> torch.Tensor:
        up = self.upconv(dec)
        enc, up = _crop_pair(enc, up)
        up = self.act0(self.norm0(up))
        if self.merge_mode == 'concat':
            mrg = torch.cat((up, enc), 1)
        else:
            mrg = up + enc
        return self.convs(mrg)


class TwinResUNet(nn.Module):
    """Plain-torch twin for export (see torch_twin.py): shares every layer object with the ``UNet`` it was built from."""

    def __init__(self, net):
        super().__init__()
        self.down_convs, self.up_convs, self.conv_final = net.down_convs, net.up_convs, net.conv_final

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


class UNet(_u.UNet):
    """B200-native residual U-Net with the constructor of ``elektronn3.models.resunet.UNet`` (models/resunet.py:598-780)."""

    def __init__(
            self,
            in_channels: int = 1,
            out_channels: int = 2,
            n_blocks: int = 3,
            start_filts: int = 32,
            up_mode: str = 'transpose',
            merge_mode: str = 'concat',
            enc_res_blocks: int = 0,
            dec_res_blocks: int = 0,
            planar_blocks: Sequence = (),
            batch_norm: str = 'unset',
            attention: bool = False,
            activation='relu',
            normalization: str = 'batch',
            full_norm: bool = True,
            dim: int = 3,
            conv_mode: str = 'same',
    ):
        nn.Module.__init__(self)
        _u.validate_unet_args(n_blocks, dim, planar_blocks, up_mode, merge_mode, batch_norm, attention)
        if dim != 3:
            # (the reference's DownBlock / UpBlock do not hand `dim` to their ConvBlocks, which therefore always hold
            # nn.Conv3d layers, models/resunet.py:292-301: its 2D mode does not run)
            raise NotImplementedError('resunet.UNet(dim=2): the reference builds 3D ConvBlocks whatever `dim` says')
        if (enc_res_blocks or dec_res_blocks) and conv_mode != 'same':
            raise NotImplementedError('residual blocks need conv_mode="same": a VALID conv output cannot be added to the block '
                                      'input (models/resunet.py:257-258 fails with a shape mismatch)')
        self.up_mode, self.merge_mode = up_mode, merge_mode
        self.out_channels, self.in_channels = out_channels, in_channels
        self.start_filts, self.n_blocks = start_filts, n_blocks
        self.normalization, self.attention = normalization, attention
        self.conv_mode, self.activation, self.dim = conv_mode, activation, dim
        self.enc_res_blocks, self.dec_res_blocks = enc_res_blocks, dec_res_blocks
        self.planar_blocks = planar_blocks

        self.down_convs = nn.ModuleList()
        self.up_convs = nn.ModuleList()
        outs = in_channels
        for i in range(n_blocks):                       # models/resunet.py:722-741
            ins = in_channels if i == 0 else outs
            outs = start_filts * (2 ** i)
            self.down_convs.append(DownBlock(ins, outs, pooling=i < n_blocks - 1, planar=i in planar_blocks,
                                             activation=activation, normalization=normalization, conv_mode=conv_mode,
                                             res_blocks=enc_res_blocks, skip_first_residual=(i == 0)))
        for i in range(n_blocks - 1):                   # models/resunet.py:745-764
            ins = outs
            outs = ins // 2
            self.up_convs.append(UpBlock(ins, outs, up_mode=up_mode, merge_mode=merge_mode,
                                         planar=(n_blocks - 2 - i) in planar_blocks, activation=activation,
                                         normalization=normalization, conv_mode=conv_mode, res_blocks=dec_res_blocks))
        self.conv_final = nn.Conv3d(outs, out_channels, kernel_size=1)
        self.apply(self.weight_init)

    def _net(self):
        net = self.__dict__.get('_e3b_net')
        if net is None:
            def blocks_of(prefix, convs):
                out = []
                for j, cb in enumerate(convs):
                    p = f'{prefix}.convs.{j}'
                    first_two = prefix.startswith('up_convs') and j == 0 and self.merge_mode == 'concat'
                    c0, c1 = (cb.out_channels, cb.out_channels) if first_two else (cb.in_channels, 0)
                    res = None
                    if cb.residual:
                        res = 'identity' if isinstance(cb.proj, nn.Identity) else engine.ConvSpec(p + '.proj', cb.proj, None, c0, c1)
                    out.append(engine.Block(engine.ConvSpec(p + '.conv1', cb.conv1, cb.norm1, c0, c1, act=cb.act1),
                                            engine.ConvSpec(p + '.conv2', cb.conv2, cb.norm2, cb.out_channels, 0, act=cb.act2),
                                            res))
                return out
            down = [(blocks_of(f'down_convs.{i}', b.convs), b.pool_kernel()) for i, b in enumerate(self.down_convs)]
            up = []
            for i, b in enumerate(self.up_convs):
                p = f'up_convs.{i}'
                if isinstance(b.upconv, _u.ResizeConv):
                    ups = engine.ResizeSpec(p + '.upconv', b.upconv, b.norm0, b.in_channels, act=b.act0)
                else:
                    ups = engine.UpSpec(p + '.upconv', b.upconv, b.norm0, act=b.act0)
                up.append((ups, blocks_of(p, b.convs)))
            net = engine.Net(down, up, self.conv_final, self.dim, engine.WeightCache(), merge_add=self.merge_mode == 'add')
            self.__dict__['_e3b_net'] = net
        return net

    def _convs_per_block(self, i, down):
        return max(1, self.enc_res_blocks if down else self.dec_res_blocks)

    def torch_twin(self):
        return TwinResUNet(self).train(self.training)
