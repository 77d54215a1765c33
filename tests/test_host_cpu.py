"""CPU-only tests (no GPU, no compute calls into libe3b): the C-ABI library loads and exports every
symbol include/e3b.h declares, the ctypes structs match the C structs, the host-side tiling / sharding
logic of the Predictor, and the N>1 paths over `gloo` with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'e3b.h')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(e3b_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ('e3b_version', 'e3b_last_error', 'e3b_conv', 'e3b_wgrad', 'e3b_norm_act', 'e3b_norm_bwd_apply',
                 'e3b_head', 'e3b_head_bwd', 'e3b_gather_tiles', 'e3b_pack_weights'):
        assert must in names


def test_library_loads_and_exports_every_declared_symbol():
    from elektronn3_b200 import _lib
    assert os.path.isfile(_lib.LIB_PATH), 'libe3b.so not built: run __graft_entry__.build()'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f'{name} declared in include/e3b.h but not exported by libe3b.so'
    lib.e3b_version.restype = ctypes.c_int
    header_version = int(re.search(r'#define\s+E3B_VERSION\s+(\d+)', open(HEADER).read()).group(1))
    assert lib.e3b_version() == header_version


def test_ctypes_structs_match_the_c_structs(tmp_path):
    """sizeof / offsetof of every argument struct as gcc sees include/e3b.h == the ctypes mirror"""
    from elektronn3_b200 import _lib
    structs = {'e3b_conv_args': _lib.ConvArgs, 'e3b_wgrad_args': _lib.WgradArgs,
               'e3b_norm_bwd_args': _lib.NormBwdArgs, 'e3b_head_args': _lib.HeadArgs, 'e3b_pack_job': _lib.PackJob,
               'e3b_ws_job': _lib.WsJob}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){']
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['return 0;}']
    src = tmp_path / 'abi.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'abi'
    subprocess.check_call(['gcc', '-o', str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).split('\n')
    for line in out:
        if not line:
            continue
        cname, field, val = line.split()
        cls = structs[cname]
        if field == 'size':
            assert ctypes.sizeof(cls) == int(val), cname
        else:
            assert getattr(cls, field).offset == int(val), (cname, field)


def test_missing_library_fails_loudly(monkeypatch):
    from elektronn3_b200 import _lib
    monkeypatch.setattr(_lib, 'LIB_PATH', os.path.join(ROOT, 'does_not_exist.so'))
    monkeypatch.setattr(_lib, '_lib', None)
    with pytest.raises((RuntimeError, OSError)):
        _lib.lib()


def test_cpu_input_is_rejected_not_silently_computed():
    import elektronn3_b200 as e3
    m = e3.UNet(n_blocks=2, start_filts=8)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 8, 8, 8))


def test_constructor_validation_matches_reference_messages():
    import elektronn3_b200 as e3
    with pytest.raises(ValueError):
        e3.UNet(dim=4)
    with pytest.raises(ValueError):
        e3.UNet(dim=2, planar_blocks=(0,))
    with pytest.raises(ValueError):
        e3.UNet(up_mode='bogus')
    with pytest.raises(ValueError):
        e3.UNet(n_blocks=2, planar_blocks=(0, 1, 2))
    with pytest.raises(RuntimeError):
        e3.UNet(batch_norm=True)
    with pytest.raises(NotImplementedError):
        e3.UNet(attention=True)
    with pytest.raises(ValueError):
        e3.UNet(up_mode='resizeconv_nearest', merge_mode='add')        # models/unet.py:791-800
    with pytest.raises(NotImplementedError):
        e3.UNet(activation=torch.nn.Tanh())
    with pytest.raises(NotImplementedError):
        e3.UNet(activation=torch.nn.PReLU(num_parameters=8))           # one slope per channel


def test_option_modules_follow_the_reference_layout():
    """up_mode / merge_mode / activation variants: state_dict keys, shapes and activation modules as
    upconv2 / ResizeConv / get_activation build them (models/unet.py:152-199,411-449)"""
    import elektronn3_b200 as e3
    m = e3.UNet(n_blocks=2, start_filts=8, up_mode='resizeconv_linear', activation='leaky')
    sd = m.state_dict()
    assert tuple(sd['up_convs.0.upconv.conv.weight'].shape) == (8, 16, 3, 3, 3)
    assert 'up_convs.0.upconv.weight' not in sd
    assert isinstance(m.up_convs[0].act0, torch.nn.LeakyReLU) and m.up_convs[0].act0.negative_slope == 0.1
    assert m.up_convs[0].act0 is not m.up_convs[0].act1
    m1 = e3.UNet(n_blocks=2, start_filts=8, up_mode='resizeconv_nearest1', planar_blocks=(0,))
    assert tuple(m1.state_dict()['up_convs.0.upconv.conv.weight'].shape) == (8, 16, 1, 1, 1)
    assert m1.up_convs[0].upconv.scale_factor == (1, 2, 2)
    ma = e3.UNet(n_blocks=2, start_filts=8, merge_mode='add', activation=torch.nn.SiLU())
    assert tuple(ma.state_dict()['up_convs.0.conv1.weight'].shape) == (8, 8, 3, 3, 3)
    assert isinstance(ma.down_convs[0].act1, torch.nn.SiLU)
    mp = e3.UNet(n_blocks=2, start_filts=8, activation='prelu')
    assert tuple(mp.state_dict()['down_convs.0.act1.weight'].shape) == (1,) and 'up_convs.0.act0.weight' in mp.state_dict()
    from elektronn3_b200 import engine
    assert engine.act_code(torch.nn.RReLU(), False) == (1, (1 / 8 + 1 / 3) / 2, None)
    assert engine.act_code(mp.down_convs[0].act1, True)[2] is mp.down_convs[0].act1.weight
    with pytest.raises(NotImplementedError):
        engine.act_code(torch.nn.RReLU(), True)


def test_state_dict_keys_follow_the_reference_naming():
    import elektronn3_b200 as e3
    m = e3.UNet(n_blocks=3, start_filts=32, normalization='group')
    sd = m.state_dict()
    assert sum(v.numel() for v in m.parameters()) == 1356866          # SURVEY 8(a) a10, cfg 2
    assert tuple(sd['up_convs.1.upconv.weight'].shape) == (64, 32, 2, 2, 2)
    assert tuple(sd['up_convs.1.conv1.weight'].shape) == (32, 64, 3, 3, 3)
    assert tuple(sd['conv_final.weight'].shape) == (2, 32, 1, 1, 1)
    m4 = e3.UNet(n_blocks=4, start_filts=32, planar_blocks=(0, 1))
    assert sum(v.numel() for v in m4.parameters()) == 5155970          # cfg 3
    assert tuple(m4.state_dict()['down_convs.0.conv1.weight'].shape) == (32, 1, 1, 3, 3)
    assert 'down_convs.0.norm0.running_mean' in m4.state_dict()


# ------------------------------------------------------------------------------------------ tiling
@pytest.mark.parametrize('spatial,tile,world', [((512, 512, 256), (64, 64, 64), 1), ((512, 512, 256), (64, 64, 64), 8),
                                                ((16, 24, 16), (8, 8, 8), 2), ((13, 21, 29), (8, 8, 16), 3),
                                                ((1, 64, 64), (1, 16, 16), 4)])
def test_plan_tiles_partitions_the_grid_in_reference_order(spatial, tile, world):
    import itertools
    from elektronn3_b200.inference import plan_tiles
    grid = [-(-s // t) for s, t in zip(spatial, tile)]
    want = list(itertools.product(*[range(g) for g in grid]))        # inference.py:159-165 order
    got = []
    for r in range(world):
        pos, (r0, r1), rows = plan_tiles(spatial, tile, world, r)
        assert rows == grid[0] and 0 <= r0 <= r1 <= rows
        assert all(r0 <= p[0] < r1 for p in pos)
        got += [tuple(int(v) for v in p) for p in pos]
    assert got == want


def _identity_tiled(vol, tile, ovl, world, rank):
    """what Predictor.predict does per rank, with the network replaced by the identity: gather each
    (tile + 2*overlap) box with zero padding, keep its centre, place it in the rank's slab."""
    from elektronn3_b200.inference import plan_tiles, slab_rows_per_rank
    spatial = np.array(vol.shape)
    tile, ovl = np.array(tile), np.array(ovl)
    pos, (r0, r1), rows = plan_tiles(spatial, tile, world, rank)
    per = slab_rows_per_rank(rows, world)
    slab = np.zeros((per * tile[0] if world > 1 else spatial[0], spatial[1], spatial[2]), np.float32)
    padded = np.pad(vol, [(o, o + t) for o, t in zip(ovl, tile)])
    for p in pos:
        src = p * tile                                                 # origin in padded coordinates (= -ovl + ovl)
        box = padded[src[0]:src[0] + tile[0] + 2 * ovl[0], src[1]:src[1] + tile[1] + 2 * ovl[1],
                     src[2]:src[2] + tile[2] + 2 * ovl[2]]
        centre = box[ovl[0]:ovl[0] + tile[0], ovl[1]:ovl[1] + tile[1], ovl[2]:ovl[2] + tile[2]]
        d = p * tile
        d[0] -= r0 * tile[0]
        e = np.minimum(d + tile, slab.shape)
        slab[d[0]:e[0], d[1]:e[1], d[2]:e[2]] = centre[:e[0] - d[0], :e[1] - d[1], :e[2] - d[2]]
    return slab, rows, per


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from elektronn3_b200.inference import assemble_slabs
        rs = np.random.RandomState(0)
        vol = rs.standard_normal((24, 20, 19)).astype(np.float32)
        tile, ovl = (8, 8, 8), (4, 4, 4)
        slab, rows, per = _identity_tiled(vol, tile, ovl, world, rank)
        t = torch.from_numpy(slab)[None, None]
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        full = assemble_slabs(parts, rows, per, tile[0], vol.shape[0])[0, 0].numpy()
        ok_pred = bool(np.array_equal(full, vol))

        # DDP wraps the module (parameter broadcast from rank 0, bucketed all-reduce hooks registered)
        import elektronn3_b200 as e3
        torch.manual_seed(100 + rank)
        m = e3.UNet(n_blocks=2, start_filts=8)
        ddp = torch.nn.parallel.DistributedDataParallel(m)
        flat = torch.cat([p.detach().flatten() for p in ddp.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, 0)
        ok_ddp = bool(torch.equal(flat, ref))
        # gradient averaging over ranks as DDP does it for the single autograd node's outputs
        g = torch.full((4,), float(rank + 1))
        dist.all_reduce(g)
        ok_ar = bool(torch.equal(g / world, torch.full((4,), (world + 1) / 2)))
        q.put((rank, ok_pred, ok_ddp, ok_ar))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_predictor_slabs_and_ddp():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok_pred, ok_ddp, ok_ar in res:
        assert ok_pred, f'rank {rank}: gathered slabs differ from the volume'
        assert ok_ddp, f'rank {rank}: DDP did not synchronise the parameters'
        assert ok_ar


def test_single_rank_identity_tiling_equals_reference_tiled_apply():
    """the same identity run against the oracle restatement of tiled_apply (inference.py:45-199)"""
    from oracle import oracle as orc
    rs = np.random.RandomState(1)
    vol = rs.standard_normal((16, 24, 16)).astype(np.float32)
    slab, _, _ = _identity_tiled(vol, (8, 8, 8), (4, 4, 4), 1, 0)
    out = orc.tiled_apply(lambda t, c: t[c] if c is not None else t, vol[None, None],
                          (8, 8, 8), (4, 4, 4), None, (1, 1, 16, 24, 16))
    assert np.array_equal(out[0, 0], slab)


def test_conv_kernel_selection_and_weight_image_sizes():
    """e3b_conv_variant is pure host logic (shared-memory plan of the z-stacked kernel): 3x3x3 taps, narrow
    outputs, weight image resident.  The packed-weight sizes of modes 4/5 equal those of modes 0/1."""
    from elektronn3_b200 import _lib
    lib = _lib.lib()
    v = lambda C0, C1, nt, k=(3, 3, 3), sc=0: lib.e3b_conv_variant(C0, C1, nt, k[0], k[1], k[2], sc)
    assert v(32, 0, 32) == 1 and v(32, 32, 32) == 1 and v(1, 0, 32) == 1      # cfg-2 full-resolution layers
    assert v(32, 0, 64) == 1 and v(64, 0, 32) == 1                            # 32 -> 64 and its dgrad
    assert v(64, 0, 64) == 1 and v(64, 0, 128) == 1 and v(32, 0, 96) == 1    # wide outputs: N tiles of 32 columns
    assert v(128, 0, 64) == 0 and v(128, 0, 128) == 0 and v(64, 64, 64) == 0  # the image of a 32-column tile does not fit
    assert v(64, 0, 48) == 0                                                  # neither whole (image too big) nor a multiple of 32
    assert v(32, 0, 32, (1, 3, 3)) == 0 and v(32, 0, 32, (1, 1, 1)) == 0      # planar / 1x1x1: halo-tile kernel
    assert v(32, 0, 32, (3, 3, 3), 1) == 0                                    # transposed conv (scatter)
    f = lib.e3b_packed_weight_floats
    for (C0, C1, Co) in [(32, 0, 32), (32, 32, 32), (3, 0, 8), (40, 0, 48)]:
        assert f(4, C0, C1, Co, 3, 3, 3) == f(0, C0, C1, Co, 3, 3, 3) > 0
        assert f(5, C0, C1, Co, 3, 3, 3) == f(1, C0, C1, Co, 3, 3, 3) > 0
    assert f(4, 64, 0, 64, 3, 3, 3) == f(0, 64, 0, 64, 3, 3, 3) > 0          # two N tiles, same number of elements
    assert f(4, 32, 0, 32, 1, 3, 3) == -1 and f(4, 128, 0, 128, 3, 3, 3) == -1


def test_graphed_train_step_needs_cuda():
    import elektronn3_b200 as e3
    m = e3.UNet(n_blocks=2, start_filts=8)
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    with pytest.raises(RuntimeError):
        e3.GraphedTrainStep(m, torch.nn.functional.cross_entropy, opt, (1, 1, 8, 8, 8), (1, 8, 8, 8))


# ------------------------------------------------------------------------------------------ TorchScript export
@pytest.mark.parametrize('kw,shape', [(dict(n_blocks=2, start_filts=8), (1, 1, 16, 16, 16)),
                                      (dict(n_blocks=3, start_filts=4, normalization='group4', planar_blocks=(0,)), (2, 1, 5, 13, 18)),
                                      (dict(dim=2, n_blocks=3, start_filts=4), (1, 1, 24, 20)),
                                      (dict(n_blocks=2, start_filts=4, conv_mode='valid', normalization='none'), (1, 1, 20, 20, 20)),
                                      (dict(n_blocks=3, start_filts=4, normalization='group4', up_mode='resizeconv_linear',
                                            activation='leaky'), (1, 1, 9, 12, 14)),
                                      (dict(n_blocks=2, start_filts=4, merge_mode='add', activation='silu'), (1, 1, 8, 16, 16))])
def test_torch_twin_and_torchscript_export(kw, shape, tmp_path):
    """Trainer._save_model scripts / traces the model when save_jit is set (training/trainer.py:876-887): the module hands
    out a plain-torch twin with the SAME parameters and state_dict keys that computes the reference forward."""
    import elektronn3_b200 as e3
    from oracle import torch_ref
    torch.manual_seed(0)
    m = e3.UNet(**kw).eval()
    x = torch.randn(shape)
    with torch.no_grad():
        want = torch_ref.unet_forward(m, x)
        twin = m.torch_twin()
        assert list(twin.state_dict().keys()) == list(m.state_dict().keys())
        assert all(a.data_ptr() == b.data_ptr() for a, b in zip(twin.parameters(), m.parameters()))
        assert torch.allclose(twin(x), want, atol=1e-6)
        scripted = torch.jit.script(m)                       # goes through UNet.__prepare_scriptable__
        assert torch.allclose(scripted(x), want, atol=1e-6)
        path = str(tmp_path / 'model.pts')
        scripted.save(path)
        assert torch.allclose(torch.jit.load(path)(x), want, atol=1e-6)
        traced = torch.jit.trace(m, x)                       # the tracing branch of UNet.forward
        assert torch.allclose(traced(x), want, atol=1e-6)
    # the twin trains, too (gradients land in the shared parameters)
    m.train()
    m.torch_twin()(x).square().mean().backward()
    assert all(p.grad is not None for p in m.parameters())


# ------------------------------------------------------------------------------------------ resunet
def test_resunet_module_follows_the_reference_layout():
    """state_dict keys / order / shapes of elektronn3.models.resunet.UNet (the golden files carry the key list of the real
    reference module), constructor errors, shortcut kinds"""
    import json
    import elektronn3_b200 as e3
    from conftest import load_golden
    from oracle import fixtures as fx
    for name, case in fx.RESUNET_CASES.items():
        g, sd = load_golden(name)
        m = e3.resunet.UNet(**case['model'])
        assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s, _ in json.loads(str(g['keys']))], name
    m = e3.resunet.UNet(n_blocks=2, start_filts=8, enc_res_blocks=2, dec_res_blocks=1)
    assert isinstance(m, e3.UNet)                                      # Predictor / GraphedTrainStep accept it
    cb = m.down_convs[0].convs
    assert not cb[0].residual and cb[1].residual and isinstance(cb[1].proj, torch.nn.Identity)   # no shortcut from the image
    assert isinstance(m.down_convs[1].convs[0].proj, torch.nn.Conv3d)                            # 8 -> 16: projection
    assert tuple(m.up_convs[0].convs[0].proj.weight.shape) == (8, 16, 1, 1, 1)                   # over the concat
    net = m._net()
    assert [type(b.res).__name__ for b in net.down[0][0]] == ['NoneType', 'str']
    assert net.down[1][0][0].res.name == 'down_convs.1.convs.0.proj' and net.down[1][0][0].c2.residual
    assert m.output_spatial((16, 16, 16)) == (16, 16, 16)
    with pytest.raises(NotImplementedError):
        e3.resunet.UNet(dim=2)
    with pytest.raises(NotImplementedError):
        e3.resunet.UNet(enc_res_blocks=1, conv_mode='valid')
    with pytest.raises(ValueError):
        e3.resunet.UNet(up_mode='bogus')


@pytest.mark.parametrize('kw,shape', [(dict(n_blocks=2, start_filts=4), (1, 1, 8, 16, 16)),
                                      (dict(n_blocks=3, start_filts=4, normalization='group4', enc_res_blocks=2, dec_res_blocks=1,
                                            planar_blocks=(0,)), (2, 1, 5, 13, 18)),
                                      (dict(n_blocks=2, start_filts=4, enc_res_blocks=1, dec_res_blocks=1, merge_mode='add',
                                            activation='silu'), (1, 1, 8, 8, 16))])
def test_resunet_twin_and_torchscript_export(kw, shape, tmp_path):
    import elektronn3_b200 as e3
    from oracle import torch_ref
    torch.manual_seed(0)
    m = e3.resunet.UNet(**kw).eval()
    x = torch.randn(shape)
    with torch.no_grad():
        want = torch_ref.unet_forward(m, x)
        twin = m.torch_twin()
        assert list(twin.state_dict().keys()) == list(m.state_dict().keys())
        assert torch.allclose(twin(x), want, atol=1e-6)
        scripted = torch.jit.script(m)
        assert torch.allclose(scripted(x), want, atol=1e-6)
        path = str(tmp_path / 'res.pts')
        scripted.save(path)
        assert torch.allclose(torch.jit.load(path)(x), want, atol=1e-6)
        traced = torch.jit.trace(m, x)
        assert torch.allclose(traced(x), want, atol=1e-6)
