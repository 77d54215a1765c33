"""Per-parameter gradient error of one case (ours / cuDNN-TF32 / TF32-operand emulation vs fp32) for a few seeds and
both norm-backward paths.   usage: python scripts/debug_case.py"""
import copy, sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
import elektronn3_b200 as e3
from oracle import torch_ref

def run(seed, kw, shape, verbose=False):
    torch.manual_seed(seed)
    m = e3.UNet(**kw).cuda().train()
    m0 = copy.deepcopy(m)
    x = torch.randn(shape, device='cuda')
    out = m(x)
    g = torch.randn_like(out)
    out.backward(g)
    ours = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    o32, g32 = torch_ref.grads_with(m0, x, g, 'fp32')
    _, gtf = torch_ref.grads_with(m0, x, g, 'tf32')
    _, gem = torch_ref.grads_with(m0, x, g, 'emulate')
    gmax = max(v.abs().max().item() for v in g32.values())
    bad = []
    for k, ref in g32.items():
        sc = max(ref.abs().max().item(), 1e-2 * gmax)
        eo = ((ours[k] - ref).abs().max() / sc).item(); et = ((gtf[k] - ref).abs().max() / sc).item(); ee = ((gem[k] - ref).abs().max() / sc).item()
        if verbose:
            print('   %-32s ours %.4f tf32 %.4f emu %.4f  scale %.3e' % (k, eo, et, ee, ref.abs().max().item()))
        if eo > 3 * max(et, ee) + 5e-3:
            bad.append((k, round(eo, 4), round(et, 4), round(ee, 4)))
    print(seed, os.environ.get('E3B_NORM_BWD'), kw, shape, 'logit err %.2e' % ((out - o32).abs().max() / o32.abs().max()).item(), 'bad:', bad)

for path in ('fused', 'split'):
    os.environ['E3B_NORM_BWD'] = path
    for seed in range(3):
        run(seed, dict(n_blocks=2, start_filts=8, normalization='group', merge_mode='add', conv_mode='valid'), (1, 1, 20, 20, 20), verbose=seed == 0)
        run(seed, dict(n_blocks=2, start_filts=8, normalization='group', merge_mode='concat', conv_mode='valid'), (1, 1, 20, 20, 20))
        run(seed, dict(n_blocks=2, start_filts=8, normalization='batch', merge_mode='add', conv_mode='valid'), (2, 1, 20, 20, 20))
        run(seed, dict(n_blocks=2, start_filts=16, normalization='group', merge_mode='add', conv_mode='valid'), (1, 1, 20, 20, 20))
