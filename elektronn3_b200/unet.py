"""Drop-in replacement for ``elektronn3.models.unet.UNet`` whose forward and backward run on the
hand-written sm_100a kernels of libe3b.so.

Boundary: the reference has no plugin API; the seam is the ``torch.nn.Module``
protocol as used by ``Trainer`` (training/trainer.py:309,519-520,863-874) and ``Predictor``
(inference/inference.py:402-458).  This class therefore

* takes the reference constructor arguments (models/unet.py:755-771) with the same validation errors,
* holds its parameters in ordinary torch layers (``nn.Conv3d`` / ``nn.ConvTranspose3d`` / ``nn.GroupNorm``
  / ``nn.BatchNorm3d`` ...) under the reference's attribute names, so ``state_dict()`` keys and shapes,
  ``isinstance(m, _BatchNorm)`` scans (training/swa.py:317-331), ``torch.save(model)`` and
  ``copy.deepcopy`` behave as before and checkpoints interchange with the reference,
* but never calls those layers: ``forward`` hands the input to ``engine.forward`` (one
  ``torch.autograd.Function`` for the whole network) which launches the CUDA kernels through the C ABI.

There is no CPU / cuDNN fallback: a non-CUDA input or a missing ``libe3b.so`` raises.
Options: ``up_mode`` 'transpose' and the four 'resizeconv_*' modes, ``merge_mode`` 'concat' and 'add', activations
'relu' / 'leaky' / 'prelu' / 'silu' / 'lin' / 'rrelu' (eval mode) or modules of those types.  What is still outside the
accelerated path raises ``NotImplementedError`` at construction (``attention=True``, activation modules of other types).
"""
import copy
from typing import Sequence

import torch
import torch.nn as nn

from . import engine


def _conv_cls(dim):
    return nn.Conv3d if dim == 3 else nn.Conv2d


def _make_norm(normtype, C, dim):
    """same mapping as get_normalization (models/unet.py:77-111)"""
    if normtype is None or normtype == 'none':
        return nn.Identity()
    if normtype.startswith('group'):
        if normtype == 'group':
            groups = 8
        elif len(normtype) > len('group') and normtype[len('group'):].isdigit():
            groups = int(normtype[len('group'):])
        else:
            raise ValueError(f'normtype "{normtype}" not understood. It should be "group<G>", where <G> is the '
                             'number of groups.')
        return nn.GroupNorm(num_groups=groups, num_channels=C)
    if normtype == 'instance':
        return nn.InstanceNorm3d(C) if dim == 3 else nn.InstanceNorm2d(C)
    if normtype == 'batch':
        return nn.BatchNorm3d(C) if dim == 3 else nn.BatchNorm2d(C)
    raise ValueError(f'Unknown normalization type "{normtype}".\nValid choices are "batch", "instance", '
                     '"group" or "group<G>", where <G> is the number of groups.')


def _make_activation(activation):
    """get_activation (models/unet.py:183-199): one module per call site; modules are deep-copied"""
    if isinstance(activation, str):
        table = {'relu': nn.ReLU, 'leaky': lambda: nn.LeakyReLU(negative_slope=0.1), 'prelu': lambda: nn.PReLU(num_parameters=1),
                 'rrelu': nn.RReLU, 'silu': nn.SiLU, 'lin': nn.Identity}
        if activation not in table:
            # (the reference returns None here and fails at the first forward)
            raise ValueError(f'Unknown activation "{activation}"')
        return table[activation]()
    if not isinstance(activation, nn.Module):
        raise ValueError(f'activation must be a string or a torch module, got {activation!r}')
    engine.act_code(activation, False)          # NotImplementedError for module types without a kernel
    return copy.deepcopy(activation)


class _Container(nn.Module):
    """Parameter container mirroring a reference sub-block; the kernels are driven by UNet.forward."""

    def forward(self, *args, **kwargs):
        raise RuntimeError(f'{type(self).__name__} is a parameter container of elektronn3_b200.UNet; '
                           'call the UNet itself')


class DownConv(_Container):
    """Parameters of DownConv (models/unet.py:202-253): conv1, conv2, norm0, norm1, pool."""

    def __init__(self, cin, cout, pooling, planar, normalization, full_norm, dim, conv_mode, activation='relu'):
        super().__init__()
        self.in_channels, self.out_channels, self.pooling, self.dim = cin, cout, pooling, dim
        self.normalization = normalization
        pad = 1 if 'same' in conv_mode else 0
        k3, p3 = 3, pad
        if planar and dim == 3:
            k3, p3 = (1, 3, 3), (0, pad, pad)
        C = _conv_cls(dim)
        self.conv1 = C(cin, cout, kernel_size=k3, padding=p3)
        self.conv2 = C(cout, cout, kernel_size=k3, padding=p3)
        if pooling:
            ks = (1, 2, 2) if (planar and dim == 3) else 2
            self.pool = (nn.MaxPool3d if dim == 3 else nn.MaxPool2d)(kernel_size=ks, ceil_mode=True)
            self.pool_ks = ks
        else:
            self.pool = nn.Identity()
            self.pool_ks = -123
        self.act1, self.act2 = _make_activation(activation), _make_activation(activation)
        self.norm0 = _make_norm(normalization, cout, dim) if full_norm else nn.Identity()
        self.norm1 = _make_norm(normalization, cout, dim)

    def pool_kernel(self):
        if not self.pooling:
            return None
        ks = self.pool_ks
        if isinstance(ks, int):
            return (1, ks, ks) if self.dim == 2 else (ks, ks, ks)
        return tuple(ks)


class DummyAttention(_Container):
    pass


class ResizeConv(nn.Module):
    """Parameters of ResizeConv (models/unet.py:411-449): 2x ``nn.Upsample`` + conv3 (padding 1) / conv1.  ``forward`` is
    plain torch and serves the export twin only (torch_twin.py); the network runs e3b_upsample_qh + the conv kernels."""

    def __init__(self, in_channels, out_channels, kernel_size=3, planar=False, dim=3, upsampling_mode='nearest'):
        super().__init__()
        self.upsampling_mode = upsampling_mode
        self.scale_factor = 2
        if dim == 3 and planar:
            self.scale_factor = (1, 2, 2)
        self.dim = dim
        self.upsample = nn.Upsample(scale_factor=self.scale_factor, mode=self.upsampling_mode)
        C = _conv_cls(dim)
        if kernel_size == 3:
            k3, p3 = (3, 1) if not (planar and dim == 3) else ((1, 3, 3), (0, 1, 1))
            self.conv = C(in_channels, out_channels, kernel_size=k3, padding=p3)
        elif kernel_size == 1:
            self.conv = C(in_channels, out_channels, kernel_size=1)
        else:
            raise ValueError(f'kernel_size={kernel_size} is not supported. Choose 1 or 3.')

    def forward(self, x):
        return self.conv(self.upsample(x))


class UpConv(_Container):
    """Parameters of UpConv (models/unet.py:328-408): upconv, conv1, conv2, norm0..2."""

    def __init__(self, cin, cout, planar, normalization, full_norm, dim, conv_mode, activation='relu', merge_mode='concat',
                 up_mode='transpose'):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.merge_mode, self.up_mode, self.normalization = merge_mode, up_mode, normalization
        pad = 1 if 'same' in conv_mode else 0
        k3, p3, k2 = 3, pad, 2
        if planar and dim == 3:
            k3, p3, k2 = (1, 3, 3), (0, pad, pad), (1, 2, 2)
        C = _conv_cls(dim)
        CT = nn.ConvTranspose3d if dim == 3 else nn.ConvTranspose2d
        if up_mode == 'transpose':               # upconv2, models/unet.py:152-175
            self.upconv = CT(cin, cout, kernel_size=k2, stride=k2)
        else:
            mode = ('trilinear' if dim == 3 else 'bilinear') if 'linear' in up_mode else 'nearest'
            self.upconv = ResizeConv(cin, cout, kernel_size=1 if up_mode.endswith('1') else 3, planar=planar, dim=dim,
                                     upsampling_mode=mode)
        self.conv1 = C(2 * cout if merge_mode == 'concat' else cout, cout, kernel_size=k3, padding=p3)
        self.conv2 = C(cout, cout, kernel_size=k3, padding=p3)
        self.act0, self.act1, self.act2 = (_make_activation(activation), _make_activation(activation),
                                           _make_activation(activation))
        if full_norm:
            self.norm0 = _make_norm(normalization, cout, dim)
            self.norm1 = _make_norm(normalization, cout, dim)
        else:
            self.norm0, self.norm1 = nn.Identity(), nn.Identity()
        self.norm2 = _make_norm(normalization, cout, dim)
        self.attention = DummyAttention()
        self.att = None


def _called_from_trace_check():
    import sys
    f = sys._getframe(1)
    for _ in range(24):
        if f is None:
            return False
        if f.f_code.co_name == '_check_trace' and 'jit' in f.f_code.co_filename:
            return True
        f = f.f_back
    return False


def validate_unet_args(n_blocks, dim, planar_blocks, up_mode, merge_mode, batch_norm, attention):
    """the reference's argument validation (models/unet.py:773-833, models/resunet.py:640-700), same exception types"""
    if n_blocks < 1:
        raise ValueError('n_blocks must be > 1.')
    if dim not in {2, 3}:
        raise ValueError('dim has to be 2 or 3')
    if dim == 2 and tuple(planar_blocks) != ():
        raise ValueError('If dim=2, you can\'t use planar_blocks since everything will be planar '
                         '(2-dimensional) anyways.\nEither set dim=3 or set planar_blocks=().')
    valid_up = ('transpose', 'upsample', 'resizeconv_nearest', 'resizeconv_linear', 'resizeconv_nearest1',
                'resizeconv_linear1')
    if up_mode not in valid_up:
        raise ValueError(f'"{up_mode}" is not a valid mode for upsampling')
    if merge_mode not in ('concat', 'add'):
        raise ValueError(f'"{merge_mode}" is not a valid mode for merging up and down paths. '
                         'Only "concat" and "add" are allowed.')
    if 'resizeconv' in up_mode and merge_mode == 'add':
        raise ValueError('up_mode "resizeconv" is incompatible with merge_mode "add"')
    if len(planar_blocks) > n_blocks:
        raise ValueError('planar_blocks can\'t be longer than n_blocks.')
    if planar_blocks and (max(planar_blocks) >= n_blocks or min(planar_blocks) < 0):
        raise ValueError('planar_blocks has invalid value range. All values have to be block indices, '
                         'meaning integers between 0 and (n_blocks - 1).')
    if batch_norm != 'unset':
        raise RuntimeError('The `batch_norm` option has been replaced with the more general `normalization` '
                           'option.\nIf you still want to use batch normalization, set `normalization=batch` '
                           'instead.')
    # --- options outside the accelerated hot path
    if up_mode == 'upsample':
        # (valid by the reference's check, but its upconv2 has no branch for it and returns None: unet.py:152-175)
        raise NotImplementedError('up_mode="upsample" is not on the B200 path')
    if attention:
        raise NotImplementedError('attention=True (GridAttention) is not on the B200 path')


class _UNetFunction(torch.autograd.Function):
    """The whole encoder/decoder as ONE autograd node: forward saves the QP activations on the ctx,
    backward launches dgrad / wgrad / norm-backward kernels and returns per-parameter gradients."""

    @staticmethod
    def forward(ctx, model, save, x, *params):
        # `save` is decided by the caller: grad mode is always off in here, and ctx.needs_input_grad reflects
        # requires_grad of the inputs whatever the grad mode (so it is True under torch.no_grad(), too)
        net = model._net()
        logits, tape = engine.forward(net, x.detach(), model.training, save=save)
        ctx.model, ctx.tape, ctx.params = model, (tape if save else None), params
        ctx.need_dx = x.requires_grad
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        if ctx.tape is None:
            raise RuntimeError('backward called on a forward that saved no activations')
        grads, dx = engine.backward(ctx.model._net(), ctx.tape, dlogits, ctx.need_dx)
        ctx.tape = None
        return (None, None, dx) + tuple(grads.get(id(p)) if p.requires_grad else None for p in ctx.params)


class UNet(nn.Module):
    """B200-native U-Net with the constructor of ``elektronn3.models.unet.UNet`` (models/unet.py:550-771)."""

    def __init__(
            self,
            in_channels: int = 1,
            out_channels: int = 2,
            n_blocks: int = 3,
            start_filts: int = 32,
            up_mode: str = 'transpose',
            merge_mode: str = 'concat',
            planar_blocks: Sequence = (),
            batch_norm: str = 'unset',
            attention: bool = False,
            activation='relu',
            normalization: str = 'batch',
            full_norm: bool = True,
            dim: int = 3,
            conv_mode: str = 'same',
    ):
        super().__init__()
        validate_unet_args(n_blocks, dim, planar_blocks, up_mode, merge_mode, batch_norm, attention)

        self.up_mode, self.merge_mode = up_mode, merge_mode
        self.out_channels, self.in_channels = out_channels, in_channels
        self.start_filts, self.n_blocks = start_filts, n_blocks
        self.normalization, self.attention = normalization, attention
        self.conv_mode, self.activation, self.dim = conv_mode, activation, dim
        self.planar_blocks = planar_blocks

        self.down_convs = nn.ModuleList()
        self.up_convs = nn.ModuleList()
        outs = in_channels
        for i in range(n_blocks):                       # channel plan: models/unet.py:840-857
            ins = in_channels if i == 0 else outs
            outs = start_filts * (2 ** i)
            self.down_convs.append(DownConv(ins, outs, pooling=i < n_blocks - 1, planar=i in planar_blocks,
                                            normalization=normalization, full_norm=full_norm, dim=dim,
                                            conv_mode=conv_mode, activation=activation))
        for i in range(n_blocks - 1):                   # models/unet.py:861-879
            ins = outs
            outs = ins // 2
            self.up_convs.append(UpConv(ins, outs, planar=(n_blocks - 2 - i) in planar_blocks,
                                        normalization=normalization, full_norm=full_norm, dim=dim,
                                        conv_mode=conv_mode, activation=activation, merge_mode=merge_mode,
                                        up_mode=up_mode))
        self.conv_final = _conv_cls(dim)(outs, out_channels, kernel_size=1)
        self.apply(self.weight_init)

    @staticmethod
    def weight_init(m):
        """Xavier-normal weights, zero biases (models/unet.py:885-892)"""
        if isinstance(m, (nn.Conv3d, nn.Conv2d, nn.ConvTranspose3d, nn.ConvTranspose2d)):
            nn.init.xavier_normal_(m.weight)
            if getattr(m, 'bias') is not None:
                nn.init.constant_(m.bias, 0)

    # ---- engine description; never pickled, rebuilt lazily (torch.save(model), deepcopy, DataParallel replicas)
    def _net(self):
        net = self.__dict__.get('_e3b_net')
        if net is None:
            cache = engine.WeightCache()
            down, up = [], []
            for i, b in enumerate(self.down_convs):
                p = f'down_convs.{i}'
                down.append(([engine.Block(engine.ConvSpec(p + '.conv1', b.conv1, b.norm0, b.in_channels, 0, act=b.act1),
                                           engine.ConvSpec(p + '.conv2', b.conv2, b.norm1, b.out_channels, 0, act=b.act2))],
                             b.pool_kernel()))
            for i, b in enumerate(self.up_convs):
                p = f'up_convs.{i}'
                if isinstance(b.upconv, ResizeConv):
                    ups = engine.ResizeSpec(p + '.upconv', b.upconv, b.norm0, b.in_channels, act=b.act0)
                else:
                    ups = engine.UpSpec(p + '.upconv', b.upconv, b.norm0, act=b.act0)
                add = b.merge_mode == 'add'
                up.append((ups, [engine.Block(
                    engine.ConvSpec(p + '.conv1', b.conv1, b.norm1, b.out_channels, 0 if add else b.out_channels, act=b.act1),
                    engine.ConvSpec(p + '.conv2', b.conv2, b.norm2, b.out_channels, 0, act=b.act2))]))
            net = engine.Net(down, up, self.conv_final, self.dim, cache, merge_add=self.merge_mode == 'add')
            self.__dict__['_e3b_net'] = net
        return net

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop('_e3b_net', None)
        return state

    def _replicate_for_data_parallel(self):
        replica = super()._replicate_for_data_parallel()
        replica.__dict__.pop('_e3b_net', None)
        return replica

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop('_e3b_net', None)      # parameters may be re-created (.to(), .half(), ...)
        return super()._apply(fn, *args, **kwargs)

    def output_spatial(self, in_spatial):
        """Spatial extents of ``forward``'s result for an input of spatial extents ``in_spatial``: the layer arithmetic
        of models/unet.py (conv3 :131-149, ceil-mode pooling :225-229, transposed conv :160-165, autocrop :256-325)
        without running anything.  SAME nets return ``in_spatial``; VALID nets shrink (unet.py:714-753)."""
        dim = self.dim
        cur = [1] * (3 - dim) + [int(v) for v in in_spatial]
        if len(cur) != 3:
            raise ValueError(f'expected {dim} spatial extents, got {tuple(in_spatial)}')
        pad = 1 if 'same' in self.conv_mode else 0

        def conv2x(sp, planar, nblocks=1):
            lose = 2 * 2 * (1 - pad) * nblocks             # two 3-tap convolutions per block
            out = [sp[0] - (0 if (planar or dim == 2) else lose), sp[1] - lose, sp[2] - lose]
            if min(out) < 1:
                raise RuntimeError(f'input extents {tuple(in_spatial)} are too small for this network')
            return out
        enc = []
        for i, b in enumerate(self.down_convs):
            planar = i in self.planar_blocks
            cur = conv2x(cur, planar, self._convs_per_block(i, True))
            enc.append(list(cur))
            if b.pooling:
                k = b.pool_kernel()
                cur = [-(-c // kk) for c, kk in zip(cur, k)]
        for i, b in enumerate(self.up_convs):
            planar = (self.n_blocks - 2 - i) in self.planar_blocks
            s = (1, 2, 2) if (planar or dim == 2) else (2, 2, 2)
            e = enc[-(i + 2)]
            up = [c * ss for c, ss in zip(cur, s)]
            up = [u - ((u - d) % 2) for u, d in zip(up, e)]
            if any(u > d for u, d in zip(up, e)):
                raise RuntimeError('autocrop: the upsampled tensor exceeds the skip tensor')
            cur = conv2x(up, planar, self._convs_per_block(i, False))
        return tuple(cur[3 - dim:])

    def _convs_per_block(self, i, down):
        """two-conv blocks per level (resunet.UNet stacks several)"""
        return 1

    def invalidate_weight_cache(self):
        """Drop the packed weight images (eval mode keeps them between calls).  Needed only after writing parameters
        or BatchNorm statistics through ``.data`` / raw pointers while staying in eval mode; ``train()``, ``eval()``,
        ``load_state_dict`` and every training-mode forward do it themselves."""
        net = self.__dict__.get('_e3b_net')
        if net is not None:
            net.cache.invalidate()

    def train(self, mode: bool = True):
        self.invalidate_weight_cache()
        return super().train(mode)

    def load_state_dict(self, *args, **kwargs):
        self.invalidate_weight_cache()
        return super().load_state_dict(*args, **kwargs)

    def _param_tensors(self):
        """The tensors autograd must see as inputs of the network node, in a fixed order.  On an ``nn.DataParallel``
        replica ``parameters()`` is empty (``Module._replicate_for_data_parallel`` clears ``_parameters``): the broadcast
        copies hang on the submodules as plain attributes / ``_former_parameters`` (torch/nn/parallel/replicate.py), and
        gradients flow through them back to the real parameters."""
        if not getattr(self, '_is_replica', False):
            return list(self.parameters())
        out = []
        for mod in self.modules():
            former = getattr(mod, '_former_parameters', None)
            if former:
                out.extend(t for t in former.values() if t is not None)
        return out

    # ---- export (TorchScript): a plain-torch twin that shares the parameters, see torch_twin.py
    def torch_twin(self):
        """Plain-torch module with the same parameters / buffers / ``state_dict`` keys that issues the reference's ATen
        call sequence: for ``torch.jit.script`` / ``trace`` / ONNX export and for running a checkpoint without libe3b.so."""
        from .torch_twin import TwinUNet
        return TwinUNet(self).train(self.training)

    def __prepare_scriptable__(self):
        # torch.jit.script(model) -- what Trainer._save_model does with save_jit='script' (training/trainer.py:876-881)
        return self.torch_twin()

    def forward(self, x):
        if torch.jit.is_tracing():
            # torch.jit.trace(model, example) (trainer.py:882-887): the archive must hold ATen ops, not ctypes calls
            self.__dict__['_e3b_traced'] = True
            return self.torch_twin()(x)
        if self.__dict__.get('_e3b_traced'):
            # torch.jit.trace then re-runs the Python module and demands agreement with the trace to 1e-5: that one call
            # (recognised by its caller, torch/jit/_trace.py::_check_trace) must see the arithmetic that was traced
            if _called_from_trace_check():
                return self.torch_twin()(x)
            self.__dict__['_e3b_traced'] = False
        if x.dim() != self.dim + 2:
            raise RuntimeError(f'Expected {self.dim + 2}D input (N, C{", D" if self.dim == 3 else ""}, H, W), '
                               f'got shape {tuple(x.shape)}')
        if torch.is_autocast_enabled():
            x = x.float()                        # fp16 operands / fp32 accumulate regardless of autocast
        params = self._param_tensors()
        # activations (fp32 conv outputs, planar copies, pooling indices) are kept only when a backward can follow
        save = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        with torch.autocast(device_type='cuda', enabled=False):
            return _UNetFunction.apply(self, save, x, *params)

    @torch.jit.unused
    def forward_gradcp(self, x):
        """The reference trades compute for memory with per-block checkpointing (models/unet.py:918-935);
        here activations are kept (180 GB HBM), the result is identical."""
        return self.forward(x)
