"""Tiled sliding-window inference on B200: drop-in for ``elektronn3.inference.Predictor``
(inference/inference.py:246-691) when the model is an ``elektronn3_b200.UNet``.

What the reference does per tile in a serial Python loop (``tiled_apply`` inference.py:134-197: host
slice -> pageable H2D -> forward at batch 1 -> centre crop -> implicit-sync D2H) is restructured:

* the (virtually zero-padded) volume is moved to HBM once;
* tiles are gathered ON DEVICE in batches straight into the kernels' QP layout (``e3b_gather_tiles``;
  out-of-volume voxels read as 0 == the zero padding of inference.py:137-145 and :645-687);
* the network runs on the whole tile batch;
* the head kernel fuses conv_final + Softmax(1) [+ Argmax] (inference.py:443-456) with the centre crop
  and writes each tile's result at its place in the device output volume (inference.py:188-197);
* one D2H copy (pinned) returns the CPU tensor the reference would have returned.

Multi-GPU (one process per GPU, ``torch.distributed`` initialised): tiles are independent, so the tile
grid is split into contiguous slabs along the first tiled axis; every rank fills its slab and a single
``all_gather`` assembles the volume.  No collective runs during compute.
"""
import itertools
import os
import time
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import engine
from .unet import UNet


# --------------------------------------------------------------------------------------- host logic
def plan_tiles(spatial, tile, world_size=1, rank=0):
    """Tile positions (row-major, last axis fastest, like itertools.product in inference.py:159-165)
    of the tile grid covering ``spatial`` and the slab of it owned by ``rank``.

    The grid is split along the first axis with more than one tile row... specifically axis 0: rank r
    owns tile rows [r0, r1).  Returns (positions owned by rank as an (n,3) int array, (r0, r1), rows).
    """
    spatial, tile = np.asarray(spatial), np.asarray(tile)
    grid = -(-spatial // tile)
    rows = int(grid[0])
    per = -(-rows // world_size)
    r0, r1 = min(rank * per, rows), min((rank + 1) * per, rows)
    pos = [p for p in itertools.product(range(r0, r1), range(int(grid[1])), range(int(grid[2])))]
    return np.asarray(pos, dtype=np.int64).reshape(-1, 3), (r0, r1), rows


def slab_rows_per_rank(rows, world_size):
    return -(-rows // world_size)


def assemble_slabs(gathered, rows, per, tile0, extent0):
    """Concatenate equally-sized per-rank slabs (axis = first spatial axis, index -3) and cut to extent0."""
    full = torch.cat(list(gathered), dim=-3)
    return full[..., :extent0, :, :]


def set_state_dict(model, state_dict):
    """inference.py:698-710: also accepts state dicts saved from nn.DataParallel wrappers."""
    try:
        model.load_state_dict(state_dict)
    except RuntimeError:
        model.load_state_dict(OrderedDict((k.replace('module.', ''), v) for k, v in state_dict.items()))


def _is_set(a):
    return a is not None and np.any(a)


class Predictor:
    """Same constructor and ``predict`` contract as the reference ``Predictor`` (inference.py:368-388,569).

    Extra keyword arguments (not in the reference): ``tile_batch`` = tiles per forward pass;
    ``distributed`` = shard tiles over the ranks of the default process group (default: on if
    ``torch.distributed`` is initialised with more than one rank).
    """

    def __init__(self, model, state_dict_src=None, device=None, batch_size=None, tile_shape=None,
                 overlap_shape=None, offset=None, out_shape=None, out_dtype=None, float16=False,
                 apply_softmax=True, transform=None, augmentations=None, strict_shapes=False,
                 apply_argmax=False, argmax_with_threshold=None, verbose=False, report_inp_stats=False,
                 tile_batch=8, distributed=None):
        if device is None:
            device = torch.device('cuda')
        elif isinstance(device, str):
            device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('elektronn3_b200.Predictor runs on CUDA devices only (no CPU path)')
        self.device = device
        self.batch_size = batch_size
        self.out_dtype = out_dtype
        if float16:
            raise NotImplementedError('float16=True is not on the B200 path (the kernels already multiply fp16 operands with fp32 '
                                      'accumulation; a .half() module is not needed)')
        if augmentations is not None:
            raise NotImplementedError('test-time augmentations are not on the B200 path yet')
        if argmax_with_threshold is not None:
            raise NotImplementedError('argmax_with_threshold is not on the B200 path yet')
        self.float16, self.dtype = False, torch.float32
        self.transform = transform
        self.augmentations = None
        self.strict_shapes = strict_shapes
        self.apply_softmax = apply_softmax
        self.apply_argmax = apply_argmax
        self.argmax_with_threshold = None
        self.verbose = verbose
        self.report_inp_stats = report_inp_stats
        if isinstance(model, os.PathLike):
            model = str(model)
        if isinstance(model, str):
            if not os.path.isfile(model):
                raise ValueError(f'Model path {model} not found.')
            if model.endswith('.pt'):
                model = torch.load(model, map_location=device, weights_only=False)
            elif model.endswith('.pts'):
                raise NotImplementedError('TorchScript archives (.pts) cannot carry the B200 kernels; load the .pt '
                                          'or a state_dict into elektronn3_b200.UNet')
            else:
                raise ValueError(f'{model} has an unkown file extension. Supported are .pt and .pts')
        if isinstance(model, (nn.DataParallel, nn.parallel.DistributedDataParallel)):
            model = model.module
        if not isinstance(model, UNet):
            raise NotImplementedError(f'elektronn3_b200.Predictor accelerates elektronn3_b200.UNet models; got '
                                      f'{type(model).__name__} (use the reference Predictor for other models)')
        if isinstance(state_dict_src, str):
            state_dict = torch.load(state_dict_src, map_location=device)
            if 'model_state_dict' in state_dict:
                state_dict = state_dict['model_state_dict']
        elif isinstance(state_dict_src, dict) or state_dict_src is None:
            state_dict = state_dict_src
        else:
            raise ValueError('"state_dict_src" has to be either a path to a .pth file (str), a state_dict object '
                             '(dict) or None.')
        if state_dict is not None:
            set_state_dict(model, state_dict)
        self.model = model
        if apply_argmax and self.out_dtype is None:
            self.out_dtype = torch.uint8
        self.model.eval()                        # inference.py:458

        if _is_set(overlap_shape) and _is_set(offset):
            raise ValueError(f'overlap_shape={overlap_shape} and offet={offset} are both specified, but this is not '
                             'supported.\nEither specify overlap_shape (if the spatial shape of inputs and outputs '
                             'are the same)\nor offset (if the output is smaller).')
        if not _is_set(tile_shape):
            assert not (_is_set(out_shape) or _is_set(overlap_shape) or _is_set(offset)), \
                'If tile_shape is not set, out_shape, overlap_shape and offset should not be set either.'
            self.enable_tiling = False
        else:
            assert _is_set(out_shape), 'If tile_shape is set, out_shape is required to be set, too.'
            self.enable_tiling = True
            if offset is None:
                # the reference probes with a forward pass (data/utils.py:63-78); for this model class the
                # answer is known from conv_mode
                if 'same' not in model.conv_mode:
                    raise NotImplementedError('tiled inference with conv_mode="valid" is not on the B200 path yet')
                offset = (0,) * model.dim
            if np.count_nonzero(offset) != 0:
                raise NotImplementedError('tiled inference with a non-zero offset (VALID convolutions) is not on '
                                          'the B200 path yet')
        self.offset = None
        self.overlap_shape = np.array(overlap_shape) if overlap_shape is not None else None
        self.tile_shape = np.array(tile_shape) if tile_shape is not None else None
        self.out_shape = np.array(out_shape) if out_shape is not None else None
        self.tile_batch = int(tile_batch)
        self.distributed = distributed
        self.last_stats = {}

    # ------------------------------------------------------------------------------------------------
    def _dist(self):
        import torch.distributed as dist
        use = self.distributed
        if use is None:
            use = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if use:
            return dist.get_world_size(), dist.get_rank()
        return 1, 0

    @torch.no_grad()
    def _predict_volume(self, dvol, n_index, spatial3, tile3, ovl3, out_mode, dout):
        """All tiles of this rank for one sample.  dvol (C, D, H, W) device; dout (Co|1, Dslab, H, W)."""
        world, rank = self._dist()
        pos, (r0, r1), rows = plan_tiles(spatial3, tile3, world, rank)
        if len(pos) == 0:
            return 0
        net = self.model._net()
        C = dvol.shape[0]
        src_org = pos * tile3 - ovl3                          # may be negative: zero padding
        dst_org = pos * tile3
        dst_org[:, 0] -= r0 * tile3[0]                        # slab-local
        org = torch.as_tensor(np.concatenate([src_org, dst_org], axis=1).astype(np.int32)).to(self.device,
                                                                                               non_blocking=True)
        in_tile = tuple(int(t + 2 * o) for t, o in zip(tile3, ovl3))
        for b0 in range(0, len(pos), self.tile_batch):
            o = org[b0:b0 + self.tile_batch]
            B = o.shape[0]
            so = o[:, :3].contiguous()
            do = o[:, 3:].contiguous()
            q = engine.gather_tiles(dvol, so, B, C, in_tile)
            feat, _ = engine.forward_features_qp(net, q, training=False, save=False)
            engine.head(feat, net.final, out_mode=out_mode, dst=dout, crop=(tuple(int(v) for v in ovl3),
                                                                            tuple(int(v) for v in tile3)),
                        dst_origin=do, dst_single=True)
        return len(pos)

    def predict(self, inp):
        """inference.py:569-643.  ``inp``: np.ndarray or torch.Tensor (N, C, [D,] H, W); returns a CPU tensor."""
        if self.transform is not None:
            if isinstance(inp, torch.Tensor):
                inp = inp.numpy()
            transformed = np.empty_like(inp)
            for i in range(inp.shape[0]):
                transformed[i], _ = self.transform(inp[i], None)
            inp = transformed
        start = time.time()
        model = self.model
        dim = model.dim
        inp = torch.as_tensor(inp)
        if inp.dim() != dim + 2:
            raise ValueError(f'expected input of shape (N, C, {"D, " if dim == 3 else ""}H, W), got {tuple(inp.shape)}')
        N, C = int(inp.shape[0]), int(inp.shape[1])
        spatial = np.array(inp.shape[2:])
        if self.enable_tiling:
            out_shape = np.array(self.out_shape)
            if np.any(out_shape[1:] % self.tile_shape) and self.strict_shapes:
                raise ValueError('Make sure that out_shape is divisible by tile_shape or relax this constraint by '
                                 'setting strict_shapes=False.')
            if not np.array_equal(out_shape[1:], spatial):
                raise ValueError(f'out_shape {tuple(out_shape)} does not match the input extents {tuple(spatial)}')
            tile = np.array(self.tile_shape)
            ovl = np.array(self.overlap_shape) if self.overlap_shape is not None else np.zeros_like(tile)
        else:
            out_shape = None
            tile, ovl = spatial.copy(), np.zeros_like(spatial)
        if self.out_dtype is None:
            self.out_dtype = torch.float32 if inp.dtype not in (torch.float16, torch.float64) else inp.dtype
        n_out = model.out_channels
        if out_shape is not None and out_shape[0] > 255 and self.out_dtype == torch.uint8:
            raise ValueError(f'C = out_shape[0] = {out_shape[0]}, but out_dtype torch.uint8 can only hold values up '
                             'to 255.')
        if self.tile_shape is None:
            self.tile_shape = spatial
        if self.overlap_shape is None:
            self.overlap_shape = np.zeros_like(spatial)
        if self.batch_size is None:
            self.batch_size = N
        num_batches = int(np.ceil(N / self.batch_size))

        # unify 2D / 3D: internal spatial rank is 3
        if dim == 2:
            spatial3, tile3, ovl3 = np.array([1, *spatial]), np.array([1, *tile]), np.array([0, *ovl])
        else:
            spatial3, tile3, ovl3 = spatial, tile, ovl
        out_mode = 2 if self.apply_argmax else (1 if self.apply_softmax else 0)
        oc = 1 if out_mode == 2 else n_out
        odt = torch.uint8 if out_mode == 2 else torch.float32

        world, rank = self._dist()
        _, (r0, r1), rows = plan_tiles(spatial3, tile3, world, rank)
        per = slab_rows_per_rank(rows, world)
        slab0 = int(per * tile3[0]) if world > 1 else int(spatial3[0])

        t_h2d = time.time()
        dinp = inp.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
        dinp5 = dinp.view(N, C, *[int(s) for s in spatial3])
        dout = torch.empty((N, oc, slab0, int(spatial3[1]), int(spatial3[2])), dtype=odt, device=self.device)
        ntiles = 0
        for n in range(N):
            ntiles += self._predict_volume(dinp5[n], n, spatial3, tile3, ovl3, out_mode, dout[n])
        if world > 1:
            import torch.distributed as dist
            parts = [torch.empty_like(dout) for _ in range(world)]
            dist.all_gather(parts, dout)
            dout = assemble_slabs(parts, rows, per, int(tile3[0]), int(spatial3[0]))
        if dout.dtype != self.out_dtype:
            dout = dout.to(self.out_dtype)
        host = torch.empty(dout.shape, dtype=dout.dtype, pin_memory=True)
        host.copy_(dout, non_blocking=True)
        torch.cuda.synchronize(self.device)
        out = host
        if dim == 2:
            out = out.squeeze(2)
        if out_mode == 2 and out_shape is not None and int(out_shape[0]) != 1:
            # reference quirk (inference.py:195-197): the (N,1,...) argmax tile is broadcast into all
            # out_shape[0] channels of the preallocated output
            out = out.expand(-1, int(out_shape[0]), *([-1] * dim)).contiguous()
        if num_batches > 1:
            out = out.to(self.dtype)             # _splitbatch_predict buffers in self.dtype (inference.py:561)
        self.last_stats = dict(tiles=ntiles, seconds=time.time() - start, h2d_bytes=inp.numel() * 4,
                               d2h_bytes=host.numel() * host.element_size())
        if self.verbose:
            dt = time.time() - start
            print(f'Inference speed: {out.numel() / dt / 1e6:.2f} MVox/s, time: {dt:.2f}.')
        return out

    def predict_proba(self, inp):
        return self.predict(inp)
