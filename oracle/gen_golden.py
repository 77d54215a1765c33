"""Generate tests/golden/*.npz by running the REAL reference (CPU, fp32) in the build
container -- TEST INFRASTRUCTURE ONLY.  Usage: ``python oracle/gen_golden.py``.

The reference is imported from /root/reference by file path (``import elektronn3``
itself fails here: colorlog / h5py / generated _version.py are absent, SURVEY.md
section 8c), with stub parent packages for ``elektronn3.data.utils`` and
``elektronn3.modules``.  Nothing under /root/reference is copied; only inputs-derived
outputs are stored.  The GPU box has no /root/reference: tests read the .npz files.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import fixtures as fx  # noqa: E402

REF = os.environ.get('E3_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    unet = _load('e3ref_unet', f'{REF}/elektronn3/models/unet.py')
    for pkg in ('elektronn3', 'elektronn3.data', 'elektronn3.modules'):
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m
    utils = types.ModuleType('elektronn3.data.utils')
    utils.calculate_offset = lambda model, tile_shape=None: (_ for _ in ()).throw(
        RuntimeError('pass offset explicitly'))
    sys.modules['elektronn3.data.utils'] = utils
    sys.modules['elektronn3.data'].utils = utils
    _load('elektronn3.modules.lovasz_losses', f'{REF}/elektronn3/modules/lovasz_losses.py')
    loss = _load('elektronn3.modules.loss', f'{REF}/elektronn3/modules/loss.py')
    inference = _load('e3ref_inference', f'{REF}/elektronn3/inference/inference.py')
    global RESUNET
    RESUNET = _load('e3ref_resunet', f'{REF}/elektronn3/models/resunet.py')
    return unet, loss, inference


RESUNET = None


def build_model(unet, kwargs, arch='unet'):
    torch.manual_seed(0)
    model = (RESUNET if arch == 'resunet' else unet).UNet(**kwargs)
    shapes = fx.state_shapes_from_torch(model)
    sd = fx.make_state(shapes)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    return model, shapes, sd


def main():
    torch.set_num_threads(8)
    torch.backends.mkldnn.enabled = True
    unet, loss_mod, inference = load_reference()
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])           # optional: names of the cases to (re)generate
    for name, case in list(fx.CASES.items()) + list(fx.OPTION_CASES.items()) + list(fx.RESUNET_CASES.items()):
        if only and name not in only:
            continue
        model, shapes, sd = build_model(unet, case['model'], case.get('arch', 'unet'))
        x = fx.make_input(case['x'])
        out = dict(keys=json.dumps(shapes), x_shape=np.array(case['x']))
        xt = torch.from_numpy(x)
        if case['train']:
            model.train()
            ncls = case['model'].get('out_channels', 2)
            logits = model(xt)
            logits.retain_grad()
            # (VALID nets: the target has the output's extents; SAME nets: those of the input, as before)
            tgt = fx.make_target((case['x'][0], 1) + tuple(logits.shape[2:]), ncls)
            crit = loss_mod.DiceLoss(apply_softmax=True)
            loss = crit(logits, torch.from_numpy(tgt))
            loss.backward()
            out['logits'] = logits.detach().numpy()
            out['dlogits'] = logits.grad.numpy()
            out['loss'] = np.array(float(loss))
            for k, p in model.named_parameters():
                out['grad_digest/' + k] = fx.digest(p.grad.numpy())
                if p.numel() <= 4096:
                    out['grad/' + k] = p.grad.numpy()
            for k, b in model.named_buffers():     # BN running stats after one train step
                out['buf/' + k] = b.numpy()
        else:
            model.eval()
            with torch.no_grad():
                out['logits'] = model(xt).numpy()
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
        print(name, out['logits'].shape, 'ok')

    for name, case in fx.PRED_CASES.items():
        if only and name not in only:
            continue
        model, shapes, sd = build_model(unet, case['model'])
        vol = fx.make_input(case['vol'], kind='neuro')
        out = dict(keys=json.dumps(shapes))
        oc = case['out_channels']
        for tag, kw, oshape in (('softmax', dict(apply_softmax=True), (oc, *case['vol'][2:])),
                                ('argmax', dict(apply_softmax=True, apply_argmax=True), (1, *case['vol'][2:]))):
            pred = inference.Predictor(model, device='cpu', tile_shape=case['tile'],
                                       overlap_shape=case['overlap'], offset=(0, 0, 0),
                                       out_shape=oshape, **kw)
            out[tag] = pred.predict(vol).numpy()
        # non-divisible volume -> _ensure_matching_shapes zero-padding (inference.py:645-687)
        odd = tuple(case['vol'][:2]) + tuple(s - 3 for s in case['vol'][2:])
        vol2 = fx.make_input(odd, kind='neuro')
        pred = inference.Predictor(model, device='cpu', tile_shape=case['tile'], overlap_shape=case['overlap'],
                                   offset=(0, 0, 0), out_shape=(oc, *odd[2:]), apply_softmax=True)
        out['softmax_odd'] = pred.predict(vol2).numpy()
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
        print(name, out['softmax'].shape, out['argmax'].dtype, 'ok')


if __name__ == '__main__':
    main()
