"""Per-parameter gradient error of one shape-matrix case (ours / cuDNN-TF32 / TF32-operand emulation vs fp32) for a few seeds."""
import copy, sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
import elektronn3_b200 as e3
from oracle import torch_ref

def run(seed, kw, shape):
    torch.manual_seed(seed)
    m = e3.UNet(**kw).cuda().train()
    m0 = copy.deepcopy(m)
    x = torch.randn(shape, device='cuda')
    out = m(x)
    g = torch.randn_like(out)
    out.backward(g)
    ours = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    _, g32 = torch_ref.grads_with(m0, x, g, 'fp32')
    _, gtf = torch_ref.grads_with(m0, x, g, 'tf32')
    _, gem = torch_ref.grads_with(m0, x, g, 'emulate')
    gmax = max(v.abs().max().item() for v in g32.values())
    bad = []
    for k, ref in g32.items():
        sc = max(ref.abs().max().item(), 1e-2 * gmax)
        eo = ((ours[k] - ref).abs().max() / sc).item(); et = ((gtf[k] - ref).abs().max() / sc).item(); ee = ((gem[k] - ref).abs().max() / sc).item()
        if eo > 3 * max(et, ee) + 5e-3:
            bad.append((k, round(eo, 4), round(et, 4), round(ee, 4)))
    print(seed, kw.get('planar_blocks'), shape, 'bad:', bad)

for seed in range(6):
    run(seed, dict(n_blocks=2, planar_blocks=(0, 1)), (2, 1, 1, 4, 4))
    run(seed, dict(n_blocks=2, dim=2), (2, 1, 4, 4))
    run(seed, dict(n_blocks=3, planar_blocks=(0, 1, 2)), (2, 1, 1, 8, 8))
    run(seed, dict(n_blocks=2, planar_blocks=(0, 1), normalization='group'), (2, 1, 1, 4, 4))
