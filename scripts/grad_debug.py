"""Per-parameter gradient error of elektronn3_b200.UNet vs a plain-torch functional restatement of the
same network (fp32 and TF32 cuDNN) on the GPU box: separates TF32 noise from bugs."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import torch.nn.functional as F
import elektronn3_b200 as e3
from oracle import fixtures as fx

def torch_forward(m, x):
    """functional re-statement of UNet.forward using the module's own parameters (3D, GN/BN/none)"""
    def norm(n, t):
        return t if isinstance(n, torch.nn.Identity) else n(t)
    enc = []
    for b in m.down_convs:
        y = F.relu(norm(b.norm0, F.conv3d(x, b.conv1.weight, b.conv1.bias, padding=b.conv1.padding)))
        y = F.relu(norm(b.norm1, F.conv3d(y, b.conv2.weight, b.conv2.bias, padding=b.conv2.padding)))
        enc.append(y)
        x = F.max_pool3d(y, b.pool.kernel_size, ceil_mode=True) if b.pooling else y
    for i, b in enumerate(m.up_convs):
        e = enc[-(i + 2)]
        u = F.conv_transpose3d(x, b.upconv.weight, b.upconv.bias, stride=b.upconv.stride)
        u = u[:, :, :e.shape[2], :e.shape[3], :e.shape[4]]
        u = F.relu(norm(b.norm0, u))
        y = F.relu(norm(b.norm1, F.conv3d(torch.cat((u, e), 1), b.conv1.weight, b.conv1.bias, padding=b.conv1.padding)))
        x = F.relu(norm(b.norm2, F.conv3d(y, b.conv2.weight, b.conv2.bias, padding=b.conv2.padding)))
    return F.conv3d(x, m.conv_final.weight, m.conv_final.bias)

kw = dict(n_blocks=3, start_filts=16, normalization=sys.argv[1] if len(sys.argv) > 1 else 'group')
shape = (2, 1, 24, 24, 24)
m = e3.UNet(**kw)
sd = fx.make_state([(k, tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()], seed=7)
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
m = m.cuda().train()
x = torch.from_numpy(fx.make_input(shape, seed=5)).cuda()
dl = torch.from_numpy(fx.make_input((2, 2, 24, 24, 24), seed=6)).cuda() * 1e-2
res = {}
for tag in ('fp32', 'tf32', 'ours'):
    m.zero_grad()
    torch.backends.cudnn.allow_tf32 = tag == 'tf32'
    out = m(x) if tag == 'ours' else torch_forward(m, x)
    out.backward(dl)
    res[tag] = (out.detach().clone(), {k: p.grad.detach().clone() for k, p in m.named_parameters()})
ref_out, ref_g = res['fp32']
print('logits rel err: tf32 %.3e ours %.3e' % tuple(((res[t][0] - ref_out).abs().max() / ref_out.abs().max()).item() for t in ('tf32', 'ours')))
gmax = max(g.abs().max().item() for g in ref_g.values())
for k in ref_g:
    sc = max(ref_g[k].abs().max().item(), 1e-2 * gmax)
    e_t = ((res['tf32'][1][k] - ref_g[k]).abs().max() / sc).item()
    e_o = ((res['ours'][1][k] - ref_g[k]).abs().max() / sc).item()
    flag = '  <<<<' if e_o > 5 * max(e_t, 2e-3) else ''
    print(f'{k:32s} tf32 {e_t:.2e}  ours {e_o:.2e}{flag}')
