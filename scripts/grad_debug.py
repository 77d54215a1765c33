"""Per-parameter gradient error of elektronn3_b200.UNet vs the plain-torch restatement (tests/torch_ref.py)
in fp32 and cuDNN-TF32, for a fixtures case: python scripts/grad_debug.py <case name>"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
import elektronn3_b200 as e3
from oracle import fixtures as fx
from conftest import load_golden
import torch_ref

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2_sf8_train'
case = fx.CASES[name]
g, sd = load_golden(name)
m = e3.UNet(**case['model'])
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
m = m.cuda().train()
x = torch.from_numpy(fx.make_input(case['x'])).cuda()
dl = torch.from_numpy(g['dlogits']).cuda()
out32, g32 = torch_ref.grads_with(m, x, dl, 'fp32')
outtf, gtf = torch_ref.grads_with(m, x, dl, 'tf32')
outem, gem = torch_ref.grads_with(m, x, dl, 'emulate')
import copy
m2 = copy.deepcopy(m)
out = m2(x); out.backward(dl)
ours = {k: p.grad for k, p in m2.named_parameters()}
print('logits: ours vs emulation %.3e' % ((out.detach() - outem).abs().max() / outem.abs().max()).item())
print('logits rel err: tf32 %.3e ours %.3e' % (((outtf - out32).abs().max() / out32.abs().max()).item(), ((out.detach() - out32).abs().max() / out32.abs().max()).item()))
gmax = max(v.abs().max().item() for v in g32.values())
for k, ref in g32.items():
    sc = max(ref.abs().max().item(), 1e-2 * gmax)
    e_t = ((gtf[k] - ref).abs().max() / sc).item(); e_o = ((ours[k] - ref).abs().max() / sc).item()
    e_e = ((ours[k] - gem[k]).abs().max() / sc).item()
    print(f'{k:32s} tf32 {e_t:.2e}  ours {e_o:.2e}  ours-vs-emulation {e_e:.2e}' + ('  <<<<' if e_e > 1e-2 else ''))
