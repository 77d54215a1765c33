"""Pin the CPU oracle (oracle/) against golden vectors produced by the real reference
(oracle/gen_golden.py).  CPU only; this is the gate that lets the GPU parity tests trust
the oracle at sizes / shapes the goldens do not cover."""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import fixtures as fx
from oracle import oracle as orc

TOL = 2e-5          # fp32 reference (oneDNN) vs double-accumulating oracle


@pytest.mark.parametrize('name', list(fx.CASES))
def test_unet_oracle_matches_reference(name):
    case = fx.CASES[name]
    g, sd = load_golden(name)
    net = orc.UNetOracle(sd, training=case['train'], **case['model'])
    x = fx.make_input(case['x'])
    logits = net.forward(x)
    assert logits.shape == g['logits'].shape
    assert rel_err(logits, g['logits']) < TOL
    if not case['train']:
        return
    grads = net.backward(g['dlogits'])
    # conv biases in front of a train-mode BN have an analytically ZERO gradient (pure
    # rounding noise in the reference), hence the absolute floor tied to the global scale
    gmax = max(np.abs(v[:-3]).max() for k, v in g.items() if k.startswith('grad_digest/'))
    for k in sd:
        if ('grad_digest/' + k) in g:
            ref = g['grad_digest/' + k]
            got = fx.digest(grads[k])
            scale = max(np.abs(ref[:-3]).max(), 1e-2 * gmax)
            assert np.abs(got[:-3] - ref[:-3]).max() / scale < 2e-4, k
            assert abs(got[-1] - ref[-1]) / max(ref[-1], 1e-2 * gmax) < 2e-4, k
        if ('grad/' + k) in g:
            ref = g['grad/' + k]
            assert np.abs(grads[k] - ref).max() / max(np.abs(ref).max(), 1e-2 * gmax) < 2e-4, k
    for k in sd:                      # BN running statistics after one training forward
        if ('buf/' + k) in g and not k.endswith('num_batches_tracked'):
            assert rel_err(net.sd[k], g['buf/' + k]) < 1e-5, k
        if ('buf/' + k) in g and k.endswith('num_batches_tracked'):
            assert int(net.sd[k]) == int(g['buf/' + k])


@pytest.mark.parametrize('name', list(fx.PRED_CASES))
def test_tiled_predictor_oracle_matches_reference(name):
    case = fx.PRED_CASES[name]
    g, sd = load_golden(name)
    net = orc.UNetOracle(sd, training=False, **case['model'])
    vol = fx.make_input(case['vol'], kind='neuro')
    oc = case['out_channels']
    out = orc.tiled_apply(lambda t, c: orc.predictor_apply(net, t, c), vol, case['tile'], case['overlap'],
                          None, (vol.shape[0], oc, *vol.shape[2:]))
    assert rel_err(out, g['softmax']) < TOL
    am = orc.tiled_apply(lambda t, c: orc.predictor_apply(net, t, c, apply_argmax=True), vol, case['tile'],
                         case['overlap'], None, (vol.shape[0], 1, *vol.shape[2:]))
    assert am.dtype == np.uint8
    # argmax must agree wherever the reference's own softmax margin is not a numerical tie
    p = g['softmax']
    margin = np.abs(p[:, 0] - p[:, 1])
    decided = margin > 1e-4
    assert np.array_equal(am[:, 0][decided], g['argmax'][:, 0][decided])
    assert decided.mean() > 0.999


def test_maxpool_ceil_and_argmax_semantics():
    x = np.arange(2 * 5 * 5 * 5, dtype=np.float32).reshape(1, 2, 5, 5, 5)
    y, idx = orc.maxpool_fwd(x, (2, 2, 2))
    assert y.shape == (1, 2, 3, 3, 3)
    assert y[0, 0, 2, 2, 2] == x[0, 0, 4, 4, 4]
    dx = orc.maxpool_bwd(np.ones_like(y), idx, x.shape, (2, 2, 2))
    assert dx.sum() == y.size and dx[0, 0, 1, 1, 1] == 1 and dx[0, 0, 0, 0, 0] == 0
    # ties: first maximum in scan order wins
    z = np.zeros((1, 1, 2, 2, 2), np.float32)
    _, idx = orc.maxpool_fwd(z, (2, 2, 2))
    assert idx.ravel()[0] == 0


def test_convT_is_gemm_plus_pixel_shuffle():
    rs = np.random.RandomState(0)
    x = rs.standard_normal((1, 3, 2, 3, 4)).astype(np.float32)
    w = rs.standard_normal((3, 5, 2, 2, 2)).astype(np.float32)
    b = rs.standard_normal((5,)).astype(np.float32)
    y = orc.convT_fwd(x, w, b)
    ref = np.einsum('ncdhw,coijk->nodihjwk', x, w).reshape(1, 5, 4, 6, 8) + b[None, :, None, None, None]
    assert rel_err(y, ref) < 1e-6


# ---------------------------------------------------------------------------------------------------------------
# The plain-torch restatement (oracle/torch_ref.py: the comparator of the full-size GPU parity tests and the CPU /
# cuDNN arms of bench.py) pinned against the same reference-generated goldens.
def _torch_model(case, sd, train):
    import torch
    import elektronn3_b200 as e3
    m = (e3.resunet.UNet if case.get('arch') == 'resunet' else e3.UNet)(**case['model'])
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    return m.train(train)


@pytest.mark.parametrize('name', list(fx.CASES) + list(fx.OPTION_CASES) + list(fx.RESUNET_CASES))
def test_torch_restatement_matches_reference(name):
    import torch
    from oracle import torch_ref
    case = {**fx.CASES, **fx.OPTION_CASES, **fx.RESUNET_CASES}[name]
    g, sd = load_golden(name)
    m = _torch_model(case, sd, case['train'])
    x = torch.from_numpy(fx.make_input(case['x']))
    if not case['train']:
        with torch.no_grad():
            assert rel_err(torch_ref.unet_forward(m, x).numpy(), g['logits']) < TOL
        return
    out = torch_ref.unet_forward(m, x)
    assert rel_err(out.detach().numpy(), g['logits']) < TOL
    out.backward(torch.from_numpy(g['dlogits']))
    gmax = max(np.abs(v[:-3]).max() for k, v in g.items() if k.startswith('grad_digest/'))
    for k, p in m.named_parameters():
        if ('grad_digest/' + k) in g:
            ref = g['grad_digest/' + k]
            got = fx.digest(p.grad.numpy())
            assert np.abs(got[:-3] - ref[:-3]).max() / max(np.abs(ref[:-3]).max(), 1e-2 * gmax) < 2e-4, k


@pytest.mark.parametrize('name', list(fx.PRED_CASES))
def test_torch_tiled_restatement_matches_reference(name):
    import torch
    from oracle import torch_ref
    case = fx.PRED_CASES[name]
    g, sd = load_golden(name)
    m = _torch_model(case, sd, False)
    vol = torch.from_numpy(fx.make_input(case['vol'], kind='neuro'))
    with torch.no_grad():
        out = torch_ref.tiled_apply(lambda t: torch_ref.unet_forward(m, t).softmax(1), vol, case['tile'], case['overlap'],
                                    (vol.shape[0], case['out_channels'], *vol.shape[2:]))
    assert rel_err(out.numpy(), g['softmax']) < TOL


def test_torch_dice_restatement_matches_reference_formula():
    """dice_loss of oracle/torch_ref.py against the closed form of modules/loss.py:165-233 on a case small enough
    to evaluate by hand: perfect prediction -> loss ~ 0, uniform prediction with 2 balanced classes -> 0.5."""
    import torch
    from oracle import torch_ref
    t = torch.tensor([[[0, 1], [1, 0]]])                        # (N=1, 2, 2)
    big = torch.zeros(1, 2, 2, 2)
    big.scatter_(1, t.unsqueeze(1), 50.0)
    assert float(torch_ref.dice_loss(big, t)) < 1e-4
    assert abs(float(torch_ref.dice_loss(torch.zeros(1, 2, 2, 2), t)) - 0.5) < 1e-4
