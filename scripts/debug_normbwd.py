"""Noise of the two norm-backward paths on the small dice test net: run-to-run, fused vs split, and each against the fp32
torch reference of the same network (oracle twin)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elektronn3_b200 as e3


def grads(m, x, t, path):
    os.environ['E3B_NORM_BWD'] = path
    m.zero_grad()
    loss = e3.DiceLoss()(m(x), t)
    loss.backward()
    return [p.grad.clone() for p in m.parameters()]


torch.manual_seed(10)
m = e3.UNet(n_blocks=2, start_filts=8, normalization='group').cuda().train()
x = torch.randn(2, 1, 16, 16, 16, device='cuda')
t = torch.randint(0, 2, (2, 16, 16, 16), device='cuda')
twin = m.torch_twin()
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
m.zero_grad()
from oracle import torch_ref
l = torch_ref.dice_loss(twin(x), t)
l.backward()
ref = [p.grad.clone() for p in m.parameters()]
f1, f2 = grads(m, x, t, 'fused'), grads(m, x, t, 'fused')
s1, s2 = grads(m, x, t, 'split'), grads(m, x, t, 'split')


def rel(a, b):
    return max(float((u - v).abs().max() / v.abs().max().clamp_min(1e-30)) for u, v in zip(a, b))


print('fused run-to-run %.3e   split run-to-run %.3e   fused vs split %.3e' % (rel(f1, f2), rel(s1, s2), rel(f1, s1)))
print('fused vs fp32 ref %.3e   split vs fp32 ref %.3e' % (rel(f1, ref), rel(s1, ref)))
for (name, _), a, b, r in zip(m.named_parameters(), f1, s1, ref):
    sc = float(r.abs().max())
    print('%-40s fused-ref %.2e split-ref %.2e fused-split %.2e' % (name, float((a - r).abs().max()) / sc, float((b - r).abs().max()) / sc,
                                                                   float((a - b).abs().max()) / sc))
