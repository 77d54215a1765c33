"""Parity at the REAL sizes of BASELINE.json's configs, on the GPU box.

Comparators, both evaluated on the same B200 with the plain-torch restatement of the reference's ATen call sequence
(oracle/torch_ref.py, pinned against reference-generated goldens by tests/test_oracle_golden.py):
  * fp32   : torch with TF32 disabled -- the arithmetic the reference's CPU path (and its golden vectors) has;
  * cuDNN-TF32 : torch's default GPU arithmetic on this build, i.e. what the reference itself computes on a GPU.
The contract (BASELINE.json north_star) is 1e-3 relative against the reference's PyTorch/cuDNN path.  TF32-class operands
(10 mantissa bits; ours are stored as fp16 = the same bits) cost ~1-2e-3 of the output scale on these deep random nets for
cuDNN itself, so the assertion is "no worse than the reference's own GPU arithmetic":

    rms_err(ours, fp32) <= max(1e-3, 1.25 * rms_err(cuDNN-TF32, fp32))
    max_err(ours, fp32) <= max(1e-3, 1.5  * max_err(cuDNN-TF32, fp32))       (a max over ~10^7 values fluctuates more)

with err = |a - b| / max|b| (max) and ||a - b|| / ||b|| (rms).  Argmax maps: the number of voxels whose label differs from
the fp32 result is printed next to cuDNN-TF32's own count, must not exceed 1.25x that count (+ 16), and every differing
voxel must be a numerical tie (fp32 logit margin below twice the measured logit error).  Gradients: per parameter,
err(ours, fp32) <= 3 * max(err(cuDNN-TF32, fp32), err(TF32-operand emulation, fp32)) + 5e-3 of the gradient scale.
Every number is written to gpurun_out/r02_parity_fullsize.json (committed under profiles/ by the round script).
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, 'gpurun_out', 'r02_parity_fullsize.json')


@pytest.fixture(scope='module')
def e3():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    import elektronn3_b200
    return elektronn3_b200


def errs(a, b):
    a, b = a.double(), b.double()
    d = (a - b)
    return float(d.abs().max() / b.abs().max()), float(d.norm() / b.norm())


def report(name, rec):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    data = {}
    if os.path.exists(REPORT):
        try:
            data = json.load(open(REPORT))
        except ValueError:
            data = {}
    data[name] = rec
    json.dump(data, open(REPORT, 'w'), indent=1, sort_keys=True)
    print(f'\n[parity] {name}: ' + json.dumps({k: v for k, v in rec.items() if k != 'grad_noise'}))


def ref_forward(m, x, tf32):
    from oracle import torch_ref
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = tf32
    try:
        with torch.no_grad():
            return torch_ref.unet_forward(m, x)
    finally:
        torch.backends.cudnn.allow_tf32 = old


def check_logits(name, ours, l32, ltf, rec):
    mo, ro = errs(ours, l32)
    mt, rt = errs(ltf, l32)
    rec.update(logits_max_err_ours=mo, logits_rms_err_ours=ro, logits_max_err_cudnn_tf32=mt, logits_rms_err_cudnn_tf32=rt)
    # label maps against the fp32 result
    a32, ao, at = l32.argmax(1), ours.argmax(1), ltf.argmax(1)
    mis_o, mis_t = (ao != a32), (at != a32)
    top2 = l32.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])
    abs_err = float((ours - l32).abs().max())
    rec.update(argmax_voxels=int(a32.numel()), argmax_mismatch_ours=int(mis_o.sum()), argmax_mismatch_cudnn_tf32=int(mis_t.sum()),
               argmax_mismatch_max_fp32_margin=float(margin[mis_o].max()) if mis_o.any() else 0.0, logit_abs_err_ours=abs_err)
    report(name, rec)
    assert ro <= max(1e-3, 1.25 * rt), (name, 'rms', ro, rt)
    assert mo <= max(1e-3, 1.5 * mt), (name, 'max', mo, mt)
    assert rec['argmax_mismatch_ours'] <= 1.25 * rec['argmax_mismatch_cudnn_tf32'] + 16, name
    assert rec['argmax_mismatch_max_fp32_margin'] <= 2 * abs_err + 1e-12, name


def check_grads(name, m, x, dlogits, ours, rec):
    from oracle import torch_ref
    _, g32 = torch_ref.grads_with(m, x, dlogits, 'fp32')
    _, gtf = torch_ref.grads_with(m, x, dlogits, 'tf32')
    _, gem = torch_ref.grads_with(m, x, dlogits, 'emulate')
    gmax = max(v.abs().max().item() for v in g32.values())
    table, worst = {}, 0.0
    for k, ref in g32.items():
        sc = max(ref.abs().max().item(), 1e-2 * gmax)
        e_t = ((gtf[k] - ref).abs().max() / sc).item()
        e_e = ((gem[k] - ref).abs().max() / sc).item()
        e_o = ((ours[k] - ref).abs().max() / sc).item()
        table[k] = dict(ours=e_o, cudnn_tf32=e_t, tf32_operand_emulation=e_e)
        worst = max(worst, e_o / (3 * max(e_t, e_e) + 5e-3))
    rec['grad_noise'] = table
    rec['grad_worst_ratio_to_bound'] = worst
    rec['grad_max_err_ours'] = max(v['ours'] for v in table.values())
    rec['grad_max_err_cudnn_tf32'] = max(v['cudnn_tf32'] for v in table.values())
    report(name, rec)
    for k, v in table.items():
        assert v['ours'] <= 3 * max(v['cudnn_tf32'], v['tf32_operand_emulation']) + 5e-3, (name, k, v)


def make_model(e3, seed, train, **kw):
    torch.manual_seed(seed)
    m = e3.UNet(**kw).cuda()
    # non-trivial biases / affine parameters / running statistics (the reference inits them to 0 / 1 / (0, 1))
    with torch.no_grad():
        for k, p in m.named_parameters():
            if p.dim() == 1:
                p.copy_((1.0 if k.endswith('weight') else 0.0) + 0.1 * torch.randn_like(p))
        for k, b in m.named_buffers():
            if k.endswith('running_mean'):
                b.copy_(0.1 * torch.randn_like(b))
            elif k.endswith('running_var'):
                b.copy_(0.5 + torch.rand_like(b))
    return m.train(train)


def dice(logits, target):
    from oracle import torch_ref
    return torch_ref.dice_loss(logits, target)


def run_train_case(e3, name, kw, shape, seed):
    m = make_model(e3, seed, True, **kw)
    bufs0 = {k: v.clone() for k, v in m.named_buffers()}
    torch.manual_seed(seed + 1)
    x = torch.randn(shape, device='cuda')
    t = torch.randint(0, 2, (shape[0],) + tuple(shape[2:]), device='cuda')
    logits = m(x)
    loss = dice(logits, t)
    loss.backward()
    ours_g = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    with torch.no_grad():                      # the comparators start from the same BatchNorm statistics
        for k, v in m.named_buffers():
            v.copy_(bufs0[k])
    import copy
    mref = copy.deepcopy(m)
    l32 = ref_forward(copy.deepcopy(mref), x, False)
    ltf = ref_forward(copy.deepcopy(mref), x, True)
    rec = dict(model=kw, input=list(shape), loss_ours=float(loss), loss_fp32=float(dice(l32, t)))
    check_logits(name, logits.detach(), l32, ltf, rec)
    # gradients of the same upstream gradient (dlogits of the fp32 logits under the reference DiceLoss formula)
    lt = l32.clone().requires_grad_(True)
    dice(lt, t).backward()
    m.zero_grad()
    logits2 = m(x)
    logits2.backward(lt.grad)
    ours_g = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    with torch.no_grad():
        for k, v in m.named_buffers():
            v.copy_(bufs0[k])
    check_grads(name, mref, x, lt.grad, ours_g, rec)
    assert abs(rec['loss_ours'] - rec['loss_fp32']) <= 2e-3 * abs(rec['loss_fp32']) + 1e-6


def test_cfg2_train_step_full_size(e3):
    """BASELINE configs[1]: UNet(n_blocks=3, start_filts=32, GroupNorm) on (4,1,64,64,64) with the Dice loss"""
    run_train_case(e3, 'cfg2_train_4x64^3', dict(n_blocks=3, start_filts=32, normalization='group'), (4, 1, 64, 64, 64), 11)


def test_cfg3_planar_train_step_full_size(e3):
    """BASELINE configs[2]: UNet(n_blocks=4, start_filts=32, planar_blocks=(0,1)) (BatchNorm, train) on (2,1,32,128,128)"""
    run_train_case(e3, 'cfg3_planar_train_2x32x128x128', dict(n_blocks=4, start_filts=32, planar_blocks=(0, 1)), (2, 1, 32, 128, 128), 12)


def test_cfg5_2d_train_step(e3):
    """BASELINE configs[4]: UNet(dim=2, n_blocks=4, start_filts=64) (up to 512 channels), a (4,1,512,512) slice of the batch"""
    run_train_case(e3, 'cfg5_2d_train_4x512^2', dict(dim=2, n_blocks=4, start_filts=64), (4, 1, 512, 512), 13)


def test_cfg4_tiles_and_predictor_full_model(e3):
    """BASELINE configs[3]: UNet(n_blocks=4) in eval mode (BatchNorm folded) on a batch of 80^3 tiles, and the Predictor
    (tile 64^3, overlap 8) over a 192x128x128 volume against the reference's tiled loop (tile for tile)."""
    from oracle import torch_ref
    m = make_model(e3, 14, False, n_blocks=4, start_filts=32)
    torch.manual_seed(15)
    x = torch.randn(4, 1, 80, 80, 80, device='cuda')
    with torch.no_grad():
        ours = m(x)
    l32, ltf = ref_forward(m, x, False), ref_forward(m, x, True)
    check_logits('cfg4_tiles_4x80^3', ours, l32, ltf, dict(model='UNet(n_blocks=4,start_filts=32) eval', input=[4, 1, 80, 80, 80]))

    vol = torch.randn(1, 1, 192, 128, 128)
    kw = dict(device='cuda', tile_shape=(64, 64, 64), overlap_shape=(8, 8, 8), offset=(0, 0, 0), apply_softmax=True)
    out = e3.Predictor(m, out_shape=(2, 192, 128, 128), **kw).predict(vol)
    lab = e3.Predictor(m, out_shape=(1, 192, 128, 128), apply_argmax=True, **kw).predict(vol)
    assert not out.is_cuda and out.dtype == torch.float32 and lab.dtype == torch.uint8

    def ref_tiled(tf32):
        old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = tf32
        try:
            with torch.no_grad():
                return torch_ref.tiled_apply(lambda tl: torch_ref.unet_forward(m, tl.cuda()).softmax(1), vol, (64, 64, 64), (8, 8, 8),
                                             (1, 2, 192, 128, 128))
        finally:
            torch.backends.cudnn.allow_tf32 = old
    p32, ptf = ref_tiled(False), ref_tiled(True)
    mo, ro = errs(out, p32)
    mt, rt = errs(ptf, p32)
    a32 = p32.argmax(1)
    mis_o, mis_t = (lab[:, 0].long() != a32), (ptf.argmax(1) != a32)
    margin = (p32[:, 0] - p32[:, 1]).abs()
    rec = dict(volume=[192, 128, 128], tiles=12, softmax_max_err_ours=mo, softmax_rms_err_ours=ro, softmax_max_err_cudnn_tf32=mt,
               softmax_rms_err_cudnn_tf32=rt, argmax_voxels=int(a32.numel()), argmax_mismatch_ours=int(mis_o.sum()),
               argmax_mismatch_cudnn_tf32=int(mis_t.sum()),
               argmax_mismatch_max_fp32_margin=float(margin[mis_o].max()) if mis_o.any() else 0.0,
               softmax_abs_err_ours=float((out - p32).abs().max()))
    report('cfg4_predictor_192x128x128', rec)
    assert ro <= max(1e-3, 1.25 * rt) and mo <= max(1e-3, 1.5 * mt), rec
    assert rec['argmax_mismatch_ours'] <= 1.25 * rec['argmax_mismatch_cudnn_tf32'] + 16, rec
    assert rec['argmax_mismatch_max_fp32_margin'] <= 4 * rec['softmax_abs_err_ours'] + 1e-12, rec


# ---------------------------------------------------------------------------------------------------- operand range
def test_operand_range_raw_uint8_input_and_tiny_weights(e3):
    """fp16 operand storage has 5 exponent bits where the reference's TF32 has 8: inputs on the raw uint8 scale
    (legal for Predictor) and parameters 1e-4 times the usual size must not cost accuracy (per-tensor power-of-two
    weight scales; activations are re-normalised by GroupNorm)."""
    import copy
    m = make_model(e3, 21, True, n_blocks=3, start_filts=16, normalization='group')
    torch.manual_seed(22)
    x = torch.randint(0, 256, (2, 1, 32, 32, 32), device='cuda').float()
    rec = {}
    with torch.no_grad():
        ours = m(x)
        l32, ltf = ref_forward(copy.deepcopy(m), x, False), ref_forward(copy.deepcopy(m), x, True)
        rec['raw_uint8_input'] = dict(zip(('max_ours', 'rms_ours', 'max_tf32', 'rms_tf32'), errs(ours, l32) + errs(ltf, l32)))
        m2 = copy.deepcopy(m)
        for k, p in m2.named_parameters():
            if p.dim() > 1 and 'conv_final' not in k:
                p.mul_(1e-4)
        xs = torch.randn(2, 1, 32, 32, 32, device='cuda')
        ours2 = m2(xs)
        l32b, ltfb = ref_forward(copy.deepcopy(m2), xs, False), ref_forward(copy.deepcopy(m2), xs, True)
        rec['weights_x1e-4'] = dict(zip(('max_ours', 'rms_ours', 'max_tf32', 'rms_tf32'), errs(ours2, l32b) + errs(ltfb, l32b)))
    report('operand_range', rec)
    for k, v in rec.items():
        # (cuDNN runs these narrow layers in true fp32, hence the absolute floor of a TF32-class result)
        assert v['rms_ours'] <= max(2e-3, 1.25 * v['rms_tf32']), (k, v)
        assert v['max_ours'] <= max(4e-3, 1.5 * v['max_tf32']), (k, v)


def test_operand_range_bn_eval_small_running_var(e3):
    """eval path with a folded BatchNorm whose running_var is 1e-6 (fold factor ~300) in one layer: the folded weights are
    re-scaled per tensor, the 300x larger activations stay far inside the fp16 range"""
    import copy
    m = make_model(e3, 23, False, n_blocks=2, start_filts=16)
    with torch.no_grad():
        m.down_convs[0].norm1.running_var.fill_(1e-6)
        m.invalidate_weight_cache()
        x = torch.randn(1, 1, 32, 32, 32, device='cuda')
        ours = m(x)
        l32, ltf = ref_forward(copy.deepcopy(m), x, False), ref_forward(copy.deepcopy(m), x, True)
    v = dict(zip(('max_ours', 'rms_ours', 'max_tf32', 'rms_tf32'), errs(ours, l32) + errs(ltf, l32)))
    report('operand_range_bn_var_1e-6', v)
    assert torch.isfinite(ours).all()
    assert v['rms_ours'] <= max(2e-3, 1.25 * v['rms_tf32']) and v['max_ours'] <= max(4e-3, 1.5 * v['max_tf32']), v
