"""Reference-on-GPU yardstick (SURVEY.md 8d "the >=5x denominator"): the same UNet evaluated by
torch/cuDNN (TF32 convolutions, cudnn.benchmark, NCDHW, the reference module's op sequence restated in
tests/torch_ref.py) on the bench workload, timed with the same CUDA-event harness as bench.py.

  python scripts/ref_gpu_bench.py [train|pred]    -> one JSON line
"""
import json
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, 'tests'))
import torch

import bench
import elektronn3_b200 as e3
import torch_ref

torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = True
dev = torch.device('cuda')
mode = sys.argv[1] if len(sys.argv) > 1 else 'train'


def timed(fn, k):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


if mode == 'train':
    m = e3.UNet(**bench.MODEL_KW).to(dev).train()
    opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9)
    x = torch.randn(bench.BATCH, device=dev)
    t = torch.randint(0, 2, (bench.BATCH[0],) + bench.BATCH[2:], device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = bench.dice_loss(torch_ref.unet_forward(m, x), t)
        loss.backward()
        opt.step()
    for _ in range(5):
        step()
    ms = timed(step, 10)
    vox = bench.BATCH[0] * 64 ** 3
    print(json.dumps(dict(what='torch/cuDNN TF32 train step, ' + bench.WORKLOAD, ms_per_step=ms, voxels_per_s=vox / ms * 1e3)))
else:
    m = e3.UNet(n_blocks=4, start_filts=32).to(dev).eval()
    x = torch.randn(8, 1, 80, 80, 80, device=dev)
    with torch.no_grad():
        def fwd():
            return torch_ref.unet_forward(m, x).softmax(1)
        for _ in range(5):
            fwd()
        ms = timed(fwd, 10)
    print(json.dumps(dict(what='torch/cuDNN TF32 forward+softmax of UNet(n_blocks=4) on 8 tiles of 80^3 (cfg 4 tile batch)',
                          ms_per_batch=ms, out_voxels_per_s=8 * 64 ** 3 / ms * 1e3)))
