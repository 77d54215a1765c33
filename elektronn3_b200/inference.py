"""Tiled sliding-window inference on B200: drop-in for ``elektronn3.inference.Predictor``
(inference/inference.py:246-691) when the model is an ``elektronn3_b200.UNet``.

What the reference does per tile in a serial Python loop (``tiled_apply`` inference.py:134-197: host
slice -> pageable H2D -> forward at batch 1 -> centre crop -> implicit-sync D2H) is restructured:

* the (virtually zero-padded) volume is moved to HBM in z-chunks on a copy stream while the first tile rows
  already compute;
* tiles are gathered ON DEVICE in batches straight into the kernels' QH layout (``e3b_gather_tiles``;
  out-of-volume voxels read as 0 == the zero padding of inference.py:137-145 and :645-687; the test-time
  augmentation flips of inference.py:215-243 are index transforms of this gather, not copies);
* the network runs on the whole tile batch;
* the head kernel fuses conv_final + Softmax(1) [+ Threshold + Argmax] (inference.py:443-456) with the centre
  crop, the un-mirroring and mean of the test-time augmentations (inference.py:507-517) and writes each tile's
  result at its place in the device output volume (inference.py:188-197);
* finished tile rows go back to pinned host memory on the copy stream while later rows compute.

Multi-GPU (one process per GPU, ``torch.distributed`` initialised): tiles are independent, so the tile
grid is split into contiguous slabs along the first tiled axis; every rank uploads only its slab plus halo,
fills its output slab, and ONE gather of the slabs (uint8 for label maps) assembles the volume on rank 0
(SURVEY.md section 8e).  No collective runs during compute.
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

    The grid is split along axis 0: rank r owns tile rows [r0, r1).
    Returns (positions owned by rank as an (n,3) int array, (r0, r1), rows).
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


def input_slab(r0, r1, tile0, halo0, src_shift0, extent0):
    """Input planes [lo, hi) along axis 0 that the tile rows [r0, r1) read: their output range widened by the halo
    (SAME nets: the overlap, shifted by -overlap; VALID nets: the offset on the high side only), clipped to the volume."""
    if r1 <= r0:
        return 0, 0
    lo = max(r0 * tile0 + src_shift0, 0)
    hi = min(r1 * tile0 + src_shift0 + 2 * halo0, extent0)
    return int(lo), int(max(hi, lo))


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


class FlipAugment:
    """inference.py:215-226 (same constructor / methods; the accelerated Predictor only reads ``spatial_dims``)."""

    def __init__(self, dims):
        self.spatial_dims = tuple(int(d) for d in dims)
        self.dims = tuple(np.array(dims) + 2)        # dim offset to skip (N, C)

    def forward(self, inp):
        return torch.flip(inp, dims=self.dims)

    def backward(self, inp):
        return self.forward(inp)


DEFAULT_AUGMENTATIONS_3D = [FlipAugment(dims) for dims in [(0,), (1,), (0, 1), (2,), (0, 2), (1, 2), (0, 1, 2)]]
DEFAULT_AUGMENTATIONS_2D = DEFAULT_AUGMENTATIONS_3D[:3]


def _is_set(a):
    return a is not None and np.any(a)


class Predictor:
    """Same constructor and ``predict`` contract as the reference ``Predictor`` (inference.py:368-388,569).

    Extra keyword arguments (not in the reference): ``tile_batch`` = tiles per forward pass (default ``None``: a whole row
    of tiles when the free device memory allows -- the deep, small levels of the network only fill the GPU with many tiles
    per launch: 0.097 -> 0.086 s per cfg-4 volume from 8 to 32 tiles);
    ``distributed`` = shard tiles over the ranks of the default process group (default: on if
    ``torch.distributed`` is initialised with more than one rank); ``result_on`` = ``'rank0'`` (the assembled
    volume is returned by rank 0, the other ranks return ``None``) or ``'all'``; ``return_device`` = leave the result
    in HBM (a CUDA tensor) instead of copying it to the host like the reference does; ``predict`` also accepts a CUDA
    tensor as input.
    """

    def __init__(self, model, state_dict_src=None, device=None, batch_size=None, tile_shape=None,
                 overlap_shape=None, offset=None, out_shape=None, out_dtype=None, float16=False,
                 apply_softmax=True, transform=None, augmentations=None, strict_shapes=False,
                 apply_argmax=False, argmax_with_threshold=None, verbose=False, report_inp_stats=False,
                 tile_batch=None, distributed=None, result_on='rank0', return_device=False):
        if device is None:
            device = torch.device('cuda')
        elif isinstance(device, str):
            device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('elektronn3_b200.Predictor runs on CUDA devices only (no CPU path)')
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = device
        self.batch_size = batch_size
        self.out_dtype = out_dtype
        # float16=True (inference.py:402-408,445-446): the reference runs a .half() copy of the model on fp16 inputs.
        # The kernels here already multiply fp16 operands (with fp32 accumulation), so the same launches serve both;
        # the flag only rounds the returned values through fp16 like the .half() model's outputs are.
        self.float16 = bool(float16)
        self.dtype = torch.float16 if float16 else torch.float32
        self.transform = transform
        if isinstance(augmentations, int):
            augmentations = DEFAULT_AUGMENTATIONS_3D[:augmentations]
        self.augmentations = augmentations
        self.strict_shapes = strict_shapes
        self.apply_softmax = apply_softmax
        self.apply_argmax = apply_argmax
        self.argmax_with_threshold = argmax_with_threshold
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
        if not apply_softmax and augmentations is not None:
            raise ValueError('When augmentations are enabled, apply_softmax cannot be False.')
        self.model = model
        if (apply_argmax or argmax_with_threshold is not None) and self.out_dtype is None:
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
            offset = None
        else:
            assert _is_set(out_shape), 'If tile_shape is set, out_shape is required to be set, too.'
            self.enable_tiling = True
            if offset is None:
                # the reference probes with a forward pass on a 90^3 / 186^2 input (data/utils.py:63-78); for this
                # model class the same numbers follow from the layer arithmetic alone
                probe = (90,) * 3 if model.dim == 3 else (186,) * 2
                offset = tuple((i - o) // 2 for i, o in zip(probe, model.output_spatial(probe)))
            if np.count_nonzero(offset) == 0:    # no valid conv -> disable offset (inference.py:482-483)
                offset = None
            else:
                offset = np.array(offset)
                overlap_shape = offset           # inference.py:486-488
                out_shape = np.array([*out_shape[:-len(offset)], *(np.array(out_shape[-len(offset):]) - 2 * offset)])
        self.offset = offset
        self.overlap_shape = np.array(overlap_shape) if overlap_shape is not None else None
        self.tile_shape = np.array(tile_shape) if tile_shape is not None else None
        self.out_shape = np.array(out_shape) if out_shape is not None else None
        self.tile_batch = None if tile_batch is None else max(1, int(tile_batch))
        self.distributed = distributed
        if result_on not in ('rank0', 'all'):
            raise ValueError("result_on must be 'rank0' or 'all'")
        self.result_on = result_on
        self.return_device = bool(return_device)
        self.last_stats = {}
        self._copy_stream = None

    # ------------------------------------------------------------------------------------------------
    def _dist(self):
        import torch.distributed as dist
        use = self.distributed
        if use is None:
            use = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if use:
            return dist.get_world_size(), dist.get_rank()
        return 1, 0

    def _flip_masks(self, dim):
        """TTA passes as (D, H, W) bit masks; pass 0 is the un-augmented one (inference.py:507-517)."""
        masks = [0]
        for aug in self.augmentations or []:
            dims = getattr(aug, 'spatial_dims', None)
            if dims is None:
                dims = tuple(int(d) - 2 for d in aug.dims)
            m = 0
            for d in dims:
                m |= 1 << (d + (3 - dim))       # 2D data lives in (1, H, W): spatial dim 0 is H
            masks.append(m)
        return masks

    @torch.no_grad()
    def _predict_rows(self, dvol, z_lo, rows, grid12, tile3, ovl3, src_shift, crop0, out_mode, dout, r0, h2d_events,
                      chunk0, on_row_done):
        """Tile rows ``rows`` of this rank for one sample.  dvol (C, Dslab, H, W) device slab starting at input plane
        z_lo; dout (Co|1, Dslab_out, H, W).  h2d_events[i] fires when input planes up to chunk i are resident."""
        net = self.model._net()
        C = dvol.shape[0]
        in_tile = tuple(int(t + 2 * o) for t, o in zip(tile3, ovl3))
        masks = self._flip_masks(self.model.dim)
        tta = len(masks) > 1
        thr = self.argmax_with_threshold
        head_mode = 1 if tta else out_mode
        ntiles = 0
        cur = torch.cuda.current_stream(self.device)
        for r in rows:
            pos = np.asarray([p for p in itertools.product([r], range(grid12[0]), range(grid12[1]))], dtype=np.int64)
            src_org = pos * tile3 + src_shift
            src_org[:, 0] -= z_lo                              # slab-local input coordinates
            dst_org = pos * tile3
            dst_org[:, 0] -= r0 * tile3[0]                     # slab-local output coordinates
            org = torch.as_tensor(np.concatenate([src_org, dst_org], axis=1).astype(np.int32)).to(self.device, non_blocking=True)
            if h2d_events:
                need = min(((r + 1) * int(tile3[0]) + int(src_shift[0]) + 2 * int(ovl3[0]) - z_lo + chunk0 - 1) // chunk0,
                           len(h2d_events)) - 1
                if need >= 0:
                    cur.wait_event(h2d_events[need])
            tb = self.tile_batch or self._auto_tile_batch(len(pos), C, in_tile)
            for b0 in range(0, len(pos), tb):
                o = org[b0:b0 + tb]
                B = o.shape[0]
                so, do = o[:, :3].contiguous(), o[:, 3:].contiguous()
                for i, m in enumerate(masks):
                    q = engine.gather_tiles(dvol, so, B, C, in_tile, flip=m)
                    feat, _ = engine.forward_features_qp(net, q, training=False, save=False)
                    if crop0 is None:                          # VALID net: its output must BE the tile (inference.py:152-153)
                        if tuple(feat.spatial) != tuple(int(t) for t in tile3):
                            raise ValueError(f'the model maps input tiles {in_tile} to {tuple(feat.spatial)}, expected the '
                                             f'tile shape {tuple(int(t) for t in tile3)}: offset does not match the network')
                        crop = ((0, 0, 0), tuple(int(v) for v in tile3))
                    else:
                        crop = (crop0, tuple(int(v) for v in tile3))
                    engine.head(feat, net.final, out_mode=head_mode, dst=dout, crop=crop, dst_origin=do, dst_single=True,
                                flip=m, accumulate=tta and i > 0, acc_scale=(1.0 / len(masks)) if tta else 1.0,
                                threshold=thr if (not tta and out_mode == 2) else None,
                                round_half=self.float16 and head_mode != 2)
            ntiles += len(pos)
            if on_row_done is not None:
                on_row_done(r)
        return ntiles

    def _auto_tile_batch(self, row_tiles, C, in_tile):
        """tiles per forward pass when the caller did not choose: the whole tile row, capped by a quarter of the free device
        memory at ~12 fp16 activation tensors of the widest full-resolution layer per tile (a conservative bound on what
        a no-grad forward keeps alive)"""
        tb = self.__dict__.get('_tb_auto')
        if tb is None:
            vox = int(in_tile[0]) * int(in_tile[1]) * int(in_tile[2])
            width = max(int(getattr(self.model, 'start_filts', 32)), int(C))
            per_tile = vox * width * 2 * 12
            free, _ = torch.cuda.mem_get_info(self.device)
            tb = self.__dict__['_tb_auto'] = int(max(1, min(64, (free // 4) // max(per_tile, 1))))
        return max(1, min(tb, row_tiles))

    def predict(self, inp):
        """inference.py:569-643.  ``inp``: np.ndarray or torch.Tensor (N, C, [D,] H, W); returns a CPU tensor
        (on rank 0; ``None`` on the other ranks of a sharded run unless ``result_on='all'``)."""
        if self.transform is not None:
            if isinstance(inp, torch.Tensor):
                inp = inp.cpu().numpy()
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
        if inp.dtype not in (torch.float32,):
            inp = inp.to(torch.float32)
        inp = inp.contiguous()
        N, C = int(inp.shape[0]), int(inp.shape[1])
        spatial = np.array(inp.shape[2:])
        valid = self.offset is not None
        if self.enable_tiling:
            out_shape = np.array(self.out_shape)
            if np.any(out_shape[1:] % self.tile_shape) and self.strict_shapes:
                raise ValueError('Make sure that out_shape is divisible by tile_shape or relax this constraint by '
                                 'setting strict_shapes=False.')
            halo = np.array(self.offset) if valid else np.zeros_like(spatial)
            if not np.array_equal(out_shape[1:] + 2 * halo, spatial):
                raise ValueError(f'out_shape {tuple(out_shape)} does not match the input extents {tuple(spatial)}'
                                 + (f' minus 2 * offset {tuple(halo)}' if valid else ''))
            tile = np.array(self.tile_shape)
            ovl = np.array(self.overlap_shape) if self.overlap_shape is not None else np.zeros_like(tile)
            out_spatial = out_shape[1:]
        else:
            out_shape = None
            out_spatial = np.array(model.output_spatial(tuple(int(s) for s in spatial)))
            tile, ovl = out_spatial.copy(), (spatial - out_spatial) // 2
            valid = bool(np.any(ovl))
        if self.out_dtype is None:                # inference.py:613-614 (`inp` there already carries self.dtype)
            self.out_dtype = torch.uint8 if self.argmax_with_threshold is not None else self.dtype
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
        def three(a, fill):
            return np.array([fill, *a]) if dim == 2 else np.array(a)
        in3, out3, tile3, ovl3 = three(spatial, 1), three(out_spatial, 1), three(tile, 1), three(ovl, 0)
        # SAME nets: tiles read [pos*tile - overlap, ...) of the un-padded volume (virtual zero padding) and the centre is
        # cropped; VALID nets: the input already carries the halo, tiles read [pos*tile, pos*tile + tile + 2*offset)
        src_shift = np.zeros(3, dtype=np.int64) if valid else -ovl3
        crop0 = None if valid else tuple(int(v) for v in ovl3)
        want_argmax = self.apply_argmax or self.argmax_with_threshold is not None
        tta = self.augmentations is not None and len(self.augmentations) > 0
        out_mode = 2 if want_argmax else (1 if self.apply_softmax else 0)
        work_mode = 1 if tta else out_mode                    # TTA accumulates softmax maps, argmax is deferred
        oc = 1 if work_mode == 2 else n_out
        odt = torch.uint8 if work_mode == 2 else torch.float32

        world, rank = self._dist()
        _, (r0, r1), rows = plan_tiles(out3, tile3, world, rank)
        per = slab_rows_per_rank(rows, world)
        slab0 = int(per * tile3[0]) if world > 1 else int(out3[0])
        grid12 = (int(-(-out3[1] // tile3[1])), int(-(-out3[2] // tile3[2])))
        z_lo, z_hi = input_slab(r0, r1, int(tile3[0]), int(ovl3[0]), int(src_shift[0]), int(in3[0]))

        dev = self.device
        model.invalidate_weight_cache()                         # weights cannot change during one predict call
        with torch.cuda.device(dev):
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
            cs, cur = self._copy_stream, torch.cuda.current_stream(dev)
            final_on_host = world == 1 and not tta and not self.return_device     # rows can stream back as they finish
            dout = torch.empty((N, oc, slab0, int(out3[1]), int(out3[2])), dtype=odt, device=dev)
            host = None
            if final_on_host:
                host = torch.empty(dout.shape, dtype=odt, pin_memory=True)
            inp5 = inp.view(N, C, *[int(s) for s in in3])
            dslab = torch.empty((N, C, max(z_hi - z_lo, 1), int(in3[1]), int(in3[2])), dtype=torch.float32, device=dev)
            chunk0 = max(int(tile3[0]), 1)
            ntiles = 0
            h2d_bytes = 0
            cs.wait_stream(cur)
            for n in range(N):
                # H2D of this rank's slab (+ halo) in z-chunks on the copy stream; tile row r waits for its chunks only
                events = []
                with torch.cuda.stream(cs):
                    for c0 in range(0, z_hi - z_lo, chunk0):
                        c1 = min(c0 + chunk0, z_hi - z_lo)
                        for ch in range(C):        # (one contiguous block per channel: a strided host copy would be staged)
                            dslab[n, ch, c0:c1].copy_(inp5[n, ch, z_lo + c0:z_lo + c1], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(cs)
                        events.append(ev)
                        h2d_bytes += (c1 - c0) * C * int(in3[1]) * int(in3[2]) * 4
                on_row = None
                if final_on_host:
                    def on_row(r, n=n):
                        a, b = r * int(tile3[0]), min((r + 1) * int(tile3[0]), slab0)
                        ev = torch.cuda.Event()
                        ev.record(cur)
                        cs.wait_event(ev)
                        with torch.cuda.stream(cs):
                            for ch in range(oc):    # contiguous blocks: plain asynchronous DMA into the pinned result
                                host[n, ch, a:b].copy_(dout[n, ch, a:b], non_blocking=True)
                ntiles += self._predict_rows(dslab[n], z_lo, range(r0, r1), grid12, tile3, ovl3, src_shift, crop0, work_mode,
                                             dout[n], r0, events, chunk0, on_row)
            cur.wait_stream(cs)
            if tta and want_argmax:
                lab = torch.empty((N, 1) + tuple(dout.shape[2:]), dtype=torch.uint8, device=dev)
                thr = self.argmax_with_threshold
                engine.L.check(engine.L.lib().e3b_prob_argmax(dout.data_ptr(), lab.data_ptr(), N, n_out, int(np.prod(dout.shape[2:])),
                                                             1 if thr is not None else 0, float(thr or 0.0), engine._stream()),
                               'prob_argmax')
                dout = lab
            if world > 1:
                import torch.distributed as dist
                if self.result_on == 'all':
                    parts = [torch.empty_like(dout) for _ in range(world)]
                    dist.all_gather(parts, dout)
                else:
                    parts = [torch.empty_like(dout) for _ in range(world)] if rank == 0 else None
                    dist.gather(dout, parts, dst=0)
                if parts is None:
                    torch.cuda.synchronize(dev)
                    self.last_stats = dict(tiles=ntiles, seconds=time.time() - start, h2d_bytes=h2d_bytes, d2h_bytes=0)
                    return None
                dout = assemble_slabs(parts, rows, per, int(tile3[0]), int(out3[0]))
            if self.return_device:
                host = dout
            elif host is None:
                host = torch.empty(dout.shape, dtype=dout.dtype, pin_memory=True)
                host.copy_(dout, non_blocking=True)
            torch.cuda.synchronize(dev)
        out = host
        if out.dtype != self.out_dtype:
            out = out.to(self.out_dtype)
        if dim == 2:
            out = out.squeeze(2)
        if want_argmax and out_shape is not None and int(out_shape[0]) != 1:
            # reference quirk (inference.py:195-197): the (N,1,...) argmax tile is broadcast into all
            # out_shape[0] channels of the preallocated output
            out = out.expand(-1, int(out_shape[0]), *([-1] * dim)).contiguous()
        if num_batches > 1:
            out = out.to(self.dtype)             # _splitbatch_predict buffers in self.dtype (inference.py:561)
        self.last_stats = dict(tiles=ntiles, seconds=time.time() - start, h2d_bytes=0 if inp.is_cuda else h2d_bytes,
                               d2h_bytes=0 if self.return_device else host.numel() * host.element_size())
        if self.verbose:
            dt = time.time() - start
            print(f'Inference speed: {out.numel() / dt / 1e6:.2f} MVox/s, time: {dt:.2f}.')
        return out

    def predict_proba(self, inp):
        return self.predict(inp)
