"""Does tcgen05.mma kind::tf32 accumulate like fp32?  conv of TF32-pre-rounded random operands vs float64."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import torch, torch.nn.functional as F
from elektronn3_b200 import engine as eng
from test_ops_gpu import to_qp_ref, from_qp_ref
from torch_ref import tf32_round
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
for (C0, Co, S) in [(8, 16, 16), (32, 32, 16), (128, 128, 8)]:
    x = tf32_round(torch.randn(1, C0, S, S, S, device='cuda'))
    w = tf32_round(torch.randn(Co, C0, 3, 3, 3, device='cuda') * 0.05)
    ref = F.conv3d(x.double(), w.double(), padding=1)
    f32 = F.conv3d(x, w, padding=1)
    wpk = eng.pack_weights(0, w, None, C0, 0, Co, (3, 3, 3))
    q = eng.QP(to_qp_ref(x), 1, C0, S, S, S)
    y, _, _ = eng.conv_forward(q, wpk, eng.cpad16(Co), Co, (3, 3, 3), (1, 1, 1))
    got = from_qp_ref(y.t, Co)
    sc = ref.abs().max()
    print(f'C {C0}->{Co}: ours vs f64 {((got.double()-ref).abs().max()/sc).item():.3e}   torch fp32 vs f64 {((f32.double()-ref).abs().max()/sc).item():.3e}'
          f'   mean signed err ours {((got.double()-ref).mean()/sc).item():.3e}')
    # unrounded operands: what the tensor core does with raw fp32 (truncation?)
    xr = torch.randn(1, C0, S, S, S, device='cuda'); wr = torch.randn(Co, C0, 3, 3, 3, device='cuda') * 0.05
    wpk2 = eng.pack_weights(0, wr, None, C0, 0, Co, (3, 3, 3))   # (pack rounds the weights)
    q2 = eng.QP(to_qp_ref(xr), 1, C0, S, S, S)
    y2, _, _ = eng.conv_forward(q2, wpk2, eng.cpad16(Co), Co, (3, 3, 3), (1, 1, 1))
    ref_rn = F.conv3d(tf32_round(xr).double(), tf32_round(wr).double(), padding=1)
    ref_tr = F.conv3d((xr.view(torch.int32) & ~0x1FFF).view(torch.float32).double(), tf32_round(wr).double(), padding=1)
    g2 = from_qp_ref(y2.t, Co).double()
    print(f'      raw fp32 activations: vs RN-rounded ref {((g2-ref_rn).abs().max()/sc).item():.3e}  vs truncated ref {((g2-ref_tr).abs().max()/sc).item():.3e}')
