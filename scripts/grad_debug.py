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
print('logits rel err: emulation %.3e tf32 %.3e ours %.3e' % (((outem - out32).abs().max() / out32.abs().max()).item(), ((outtf - out32).abs().max() / out32.abs().max()).item(), ((out.detach() - out32).abs().max() / out32.abs().max()).item()))
def rms(a, b):
    return ((a - b).double().pow(2).sum().sqrt() / b.double().pow(2).sum().sqrt()).item()
print('logits RMS: ours-vs-emulation %.3e   tf32-vs-fp32 %.3e   ours-vs-fp32 %.3e' % (rms(out.detach(), outem), rms(outtf, out32), rms(out.detach(), out32)))
worst = max((rms(ours[k], gem[k]) / max(rms(gtf[k], g32[k]), 1e-12), k) for k in g32 if g32[k].abs().max() > 1e-3 * max(v.abs().max().item() for v in g32.values()))
print('grads: worst RMS(ours-vs-emulation)/RMS(tf32 noise) = %.3f at %s' % worst)
for k in list(g32)[:6] + list(g32)[-6:]:
    print(f'   {k:32s} rms ours-vs-emu {rms(ours[k], gem[k]):.2e}  tf32 noise {rms(gtf[k], g32[k]):.2e}  ours-vs-fp32 {rms(ours[k], g32[k]):.2e}')
gmax = max(v.abs().max().item() for v in g32.values())
for k, ref in g32.items():
    sc = max(ref.abs().max().item(), 1e-2 * gmax)
    e_t = ((gtf[k] - ref).abs().max() / sc).item(); e_o = ((ours[k] - ref).abs().max() / sc).item()
    e_e = ((ours[k] - gem[k]).abs().max() / sc).item()
    e_m = ((gem[k] - ref).abs().max() / sc).item()
    print(f'{k:32s} tf32 {e_t:.2e}  emulation {e_m:.2e}  ours {e_o:.2e}  ours-vs-emulation {e_e:.2e}' + ('  <<<<' if e_o > 3 * max(e_t, e_m) + 5e-3 else ''))
