"""Layer-by-layer comparison of the forward pass: libe3b vs the TF32-emulating torch restatement."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch, torch.nn as nn, torch.nn.functional as F
import elektronn3_b200 as e3
from elektronn3_b200 import engine as eng
from oracle import fixtures as fx
from conftest import load_golden
from torch_ref import tf32_round
from test_ops_gpu import from_qp_ref
torch.backends.cudnn.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2_sf8_train'
case = fx.CASES[name]; g, sd = load_golden(name)
m = e3.UNet(**case['model']); m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()}); m = m.cuda().train()
x = torch.from_numpy(fx.make_input(case['x'])).cuda()
with torch.no_grad():
    feat, tape = eng.forward_features(m._net(), x, True, True)
    def cmp(tag, q, ref):
        got = from_qp_ref(q.t, q.C)
        if ref.dim() == 4: ref = ref.unsqueeze(2)
        print(f'{tag:28s} rel err {((got - ref).abs().max() / ref.abs().max()).item():.3e}')
    conv = F.conv3d if m.dim == 3 else F.conv2d
    pool = F.max_pool3d if m.dim == 3 else F.max_pool2d
    convT = F.conv_transpose3d if m.dim == 3 else F.conv_transpose2d
    xx = tf32_round(x); enc = []
    for i, b in enumerate(m.down_convs):
        u1, u2 = tape.down[i]
        y = conv(xx, tf32_round(b.conv1.weight), b.conv1.bias, padding=b.conv1.padding); cmp(f'down{i}.conv1 y', u1.y, y)
        a = tf32_round(F.relu(b.norm0(y))); cmp(f'down{i}.a1', u1.a, a)
        y = conv(a, tf32_round(b.conv2.weight), b.conv2.bias, padding=b.conv2.padding); cmp(f'down{i}.conv2 y', u2.y, y)
        a = tf32_round(F.relu(b.norm1(y))); cmp(f'down{i}.a2', u2.a, a)
        enc.append(a)
        xx = pool(a, b.pool.kernel_size, ceil_mode=True) if b.pooling else a
        if b.pooling: cmp(f'down{i}.pooled', u2.pooled, xx)
    for i, b in enumerate(m.up_convs):
        u0, u1, u2, _ = tape.up[i]
        e = enc[-(i + 2)]
        y = convT(xx, tf32_round(b.upconv.weight), b.upconv.bias, stride=b.upconv.stride); cmp(f'up{i}.upconv y', u0.y, y)
        u = tf32_round(F.relu(b.norm0(y))); cmp(f'up{i}.u', u0.a, u)
        y = conv(torch.cat((u, e), 1), tf32_round(b.conv1.weight), b.conv1.bias, padding=b.conv1.padding); cmp(f'up{i}.conv1 y', u1.y, y)
        a = tf32_round(F.relu(b.norm1(y))); cmp(f'up{i}.a1', u1.a, a)
        y = conv(a, tf32_round(b.conv2.weight), b.conv2.bias, padding=b.conv2.padding); cmp(f'up{i}.conv2 y', u2.y, y)
        xx = tf32_round(F.relu(b.norm2(y))); cmp(f'up{i}.a2', u2.a, xx)

# --- detail of the first norm: how often do the TF32 roundings differ, and how far apart are the pre-rounding values?
with torch.no_grad():
    b = m.down_convs[0]; u1 = tape.down[0][0]
    y = from_qp_ref(u1.y.t, u1.y.C)
    if m.dim == 2: y = y.squeeze(2)
    ref_pre = F.relu(b.norm0(y))
    N, C = y.shape[:2]
    sc = u1.nstate.scale[:, :C].reshape(N, C, *([1] * (y.dim() - 2))); sh = u1.nstate.shift[:, :C].reshape(N, C, *([1] * (y.dim() - 2)))
    ours_pre = F.relu(torch.addcmul(sh, y, sc))
    a_ours = from_qp_ref(u1.a.t, u1.a.C)
    if m.dim == 2: a_ours = a_ours.squeeze(2)
    print('pre-rounding |ours - torch| max rel', ((ours_pre - ref_pre).abs().max() / ref_pre.abs().max()).item())
    print('fraction of a1 elements whose rounding differs', (a_ours != tf32_round(ref_pre)).float().mean().item(),
          ' (kernel a vs round(kernel-formula)):', (a_ours != tf32_round(ours_pre)).float().mean().item())
    print('is kernel a exactly tf32?', bool(((a_ours.view(torch.int32) & 0x1FFF) == 0).all()))
