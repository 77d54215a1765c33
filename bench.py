"""bench.py -- voxels/s of the UNet train step (BASELINE.json configs[1]) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one synthetic batch: forward of
UNet(n_blocks=3,start_filts=32,normalization='group') on (4,1,64,64,64) fp32, Dice loss
(the reference's DiceLoss formula, modules/loss.py:165-233, in plain torch: a boundary consumer that
stays torch), backward, SGD step.  The launches of a step are captured once in a CUDA graph and replayed
(elektronn3_b200.GraphedTrainStep; E3B_BENCH_GRAPH=0 launches them eagerly).  N>1 (torchrun): one process per GPU,
weak scaling; the graph ends after backward, then one flat NCCL all-reduce of the gradients and the optimizer step
(E3B_BENCH_GRAPH=0: stock DistributedDataParallel).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = through the public
module API with pinned HOST input/target copied in and the loss read back every step.
`--impl reference`: the CPU oracle port (oracle/) of the reference path timed on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL_KW = dict(n_blocks=3, start_filts=32, normalization='group')
BATCH = (4, 1, 64, 64, 64)
WORKLOAD = 'UNet(n_blocks=3,start_filts=32,dim=3,norm=GN) train step (fwd+DiceLoss+bwd+SGD) on synthetic (4,1,64,64,64) fp32'
FWD_BWD_GFLOP = 1175.75          # algorithmic, SURVEY.md 8(d) / BASELINE.md 2.3 (batch of 4)
# dominant kernel for the roofline: conv_tc_kernel on up_convs.1.conv1 (virtual concat 64 -> 32 at 64^3)
DOM = dict(N=4, C0=32, C1=32, Co=32, S=64)
DOM_GFLOP = 2 * 4 * 64 ** 3 * 32 * (64 * 27) / 1e9     # 115.96
DOM_TRAFFIC = 372.551e6          # dram__bytes_read + write of that launch, ncu --set full (profiles/r01_ncu_full_step_zs.csv)


def dice_loss(logits, target, eps=1e-4):
    """DiceLoss(apply_softmax=True) of the reference (modules/loss.py:165-233) restated in torch"""
    import torch
    prob = logits.softmax(1)
    onehot = torch.zeros_like(prob).scatter_(1, target.unsqueeze(1), 1.0)
    dims = (0, 2, 3, 4)
    num = 2 * (prob * onehot).sum(dims)
    den = prob.sum(dims) + onehot.sum(dims) + eps
    return (1 - num / den).mean()


class ClockSampler:
    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx or None, reasons=sorted(reasons),
                    samples=len(sm))


def cpu_baseline_run(steps=1, warmup=0):
    """The oracle port of the reference path on the host cores: fwd + bwd of the bench model on a bounded
    sample (1,1,64,64,64) of the workload."""
    import numpy as np
    from oracle import fixtures as fx
    from oracle import oracle as orc
    import elektronn3_b200 as e3
    orc.build()
    shape = (1, 1, 64, 64, 64)
    m = e3.UNet(**MODEL_KW)
    sd = fx.make_state([(k, tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()], seed=1)
    net = orc.UNetOracle(sd, training=True, **MODEL_KW)
    x = fx.make_input(shape, seed=2)
    dl = fx.make_input((1, 2, 64, 64, 64), seed=3) * 1e-3
    times = []
    for i in range(warmup + steps):
        t0 = time.time()
        net.forward(x)
        net.backward(dl)
        if i >= warmup:
            times.append(time.time() - t0)
    dt = sum(times) / len(times)
    return dict(value=64 ** 3 / dt, unit='voxels/s', cores=os.cpu_count(), kind='port',
                sample=f'fwd+bwd of the bench model on one (1,1,64,64,64) sample, C/OpenMP oracle, {dt:.2f} s/step',
                seconds_per_step=dt)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile-steps', type=int, default=0,
                    help='run only this many train steps after one warm-up and exit (for ncu; prints no bench line)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))

    if args.impl == 'reference':
        if rank != 0:
            return
        steps = max(1, min(args.steps, 3))
        cb = cpu_baseline_run(steps=steps, warmup=min(args.warmup, 1))
        line = dict(metric='voxels/s', value=cb['value'], unit='voxels/s', n_gpus=args.gpus, steps=steps,
                    warmup=min(args.warmup, 1), ms_per_step=cb['seconds_per_step'] * 1e3, higher_is_better=True,
                    scaling='weak', vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                    config=dict(workload=WORKLOAD, sample=cb['sample']),
                    cpu_baseline={k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
                    e2e=dict(value=cb['value'], unit='voxels/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import elektronn3_b200 as e3
    from elektronn3_b200 import _lib, engine

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU path in the product)'
    # a lost rank must not leave the job hanging in a collective until the caller's limit: give up loudly instead
    limit = float(os.environ.get('E3B_BENCH_TIMEOUT', '1200'))

    def give_up():
        if rank == 0:
            print(json.dumps(dict(metric='voxels/s', n_gpus=world, error=f'bench.py did not finish within {limit:.0f} s')),
                  flush=True)
        os._exit(3)
    watchdog = threading.Timer(limit, give_up)
    watchdog.daemon = True
    watchdog.start()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(1234 + rank)
    model = e3.UNet(**MODEL_KW).to(dev).train()
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9)
    voxels = BATCH[0] * BATCH[2] * BATCH[3] * BATCH[4]
    # N = 1: the whole step is replayed as one CUDA graph (elektronn3_b200/graph.py).  N > 1: the graph ends after
    # backward; the data-parallel gradient average (one flat NCCL all-reduce: what DDP's buckets compute) and the
    # optimizer step run eagerly, so that no NCCL call is captured (capturing it hung on the 2-GPU box).
    # E3B_BENCH_GRAPH=0: eager launches, N > 1 around stock DistributedDataParallel.
    use_graph = os.environ.get('E3B_BENCH_GRAPH', '1') != '0'
    gstep = None
    step_model = model
    params = [p for p in model.parameters()]
    if use_graph:
        if world > 1:                          # same initial weights on every rank (DDP broadcasts them at construction)
            for p in params:
                dist.broadcast(p.data, 0)
        gstep = e3.GraphedTrainStep(model, dice_loss, opt if world == 1 else None, BATCH, (BATCH[0],) + BATCH[2:])
    elif world > 1:
        step_model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])

    def grad_sync():
        grads = [p.grad for p in params]
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat)
        flat.div_(world)
        torch._foreach_copy_(grads, torch._utils._unflatten_dense_tensors(flat, grads))

    x_dev = torch.randn(BATCH, device=dev)
    t_dev = torch.randint(0, 2, (BATCH[0],) + BATCH[2:], device=dev)
    x_host = torch.randn(BATCH).pin_memory()
    t_host = torch.randint(0, 2, (BATCH[0],) + BATCH[2:]).pin_memory()

    def step(x, t):
        if gstep is not None:                  # the same launches, replayed as one CUDA graph (elektronn3_b200/graph.py)
            loss = gstep(x, t)[0]
            if world > 1:
                grad_sync()
                opt.step()
            return loss
        opt.zero_grad(set_to_none=True)
        loss = dice_loss(step_model(x), t)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    if args.profile_steps:
        for _ in range(1 + args.profile_steps):
            step(x_dev, t_dev)
        torch.cuda.synchronize()
        return
    for _ in range(max(args.warmup, 3)):
        step(x_dev, t_dev)
    with ClockSampler(local_rank) as clocks:
        l0 = _lib.launch_count()
        ms = timed(lambda: step(x_dev, t_dev), args.steps)
        launches = _lib.launch_count() - l0
        if gstep is not None:                  # kernels inside the replayed graph are not re-counted by the library
            launches = gstep.launches_per_step * args.steps

        def e2e_step():
            if gstep is not None:              # pinned host batch -> the graph's static buffers (H2D), replay, loss read
                return float(step(x_host, t_host))
            x = x_host.to(dev, non_blocking=True)
            t = t_host.to(dev, non_blocking=True)
            return float(step(x, t))             # D2H read of the loss, like trainer.py:575
        for _ in range(2):
            e2e_step()
        ms_e2e = timed(e2e_step, args.steps)

        # dominant kernel alone: conv_tc_kernel on the up_convs.1.conv1 shape, CUDA events on the launch stream
        d = DOM
        q0 = engine.QP.empty_half(d['N'], d['C0'], d['S'], d['S'], d['S'], dev)
        q1 = engine.QP.empty_half(d['N'], d['C1'], d['S'], d['S'], d['S'], dev)
        q0.t.normal_(), q1.t.normal_()
        w = torch.randn(d['Co'], d['C0'] + d['C1'], 3, 3, 3, device=dev) * 0.05
        dom_var = engine.conv_variant(d['C0'], d['C1'], 32, (3, 3, 3))       # the kernel the network itself uses
        wpk = engine.pack_weights(4 if dom_var else 0, w, None, d['C0'], d['C1'], d['Co'], (3, 3, 3))

        def dom():
            engine.conv_forward(q0, wpk, 32, d['Co'], (3, 3, 3), (1, 1, 1), src1=q1, stats_channels=d['Co'], variant=dom_var)
        for _ in range(3):
            dom()
        reps = 10
        ms_dom = timed(dom, reps) / reps
        # fp16 dense peak the way MEASURED_PEAKS.json measured bf16: cuBLAS matmul burst
        a = torch.randn(8192, 8192, device=dev, dtype=torch.float16)
        b = torch.randn(8192, 8192, device=dev, dtype=torch.float16)
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(5):
            best = min(best, timed(lambda: a @ b, 1))
        f16_peak = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
        del a, b
    total_voxels = voxels * world
    value = total_voxels / (ms / args.steps * 1e-3)
    e2e_value = total_voxels / (ms_e2e / args.steps * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    bf16_peak = peaks.get('bf16_tflops', 1590.0)
    achieved = DOM_GFLOP / ms_dom            # GFLOP / ms == TFLOP/s
    roofline = dict(bound='tensor', kernel=('conv_zs_kernel' if dom_var else 'conv_tc_kernel') + ' on up_convs.1.conv1 (virtual concat 32+32 -> 32 @ 4x64^3)',
                    achieved=achieved, peak=bf16_peak, unit='TFLOP/s', frac=achieved / bf16_peak,
                    traffic=DOM_TRAFFIC, ms_per_launch=ms_dom,
                    peak_note=('tcgen05 kind::f16 (fp16 operands, fp32 accumulate) runs at the bf16 rate: peak = '
                               'MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)'
                               if peaks else 'fallback 1.59 PF dense bf16 (B200_PROFILING.md)'),
                    f16_cublas_tflops_measured_here=f16_peak,
                    step_tensor_frac=FWD_BWD_GFLOP / (ms / args.steps) / bf16_peak)
    line = dict(metric='voxels/s', value=value, unit='voxels/s', n_gpus=world, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f16 operands / f32 accumulate (TF32-equivalent mantissa), f32 storage', data='synthetic',
                config=dict(workload=WORKLOAD, global_batch=BATCH[0] * world,
                            parallelism=f'dp{world}' if world > 1 else 'single',
                            launch=('eager launches' if gstep is None else
                                    'whole step replayed as one CUDA graph (GraphedTrainStep)' if world == 1 else
                                    'forward+loss+backward replayed as one CUDA graph, flat NCCL all-reduce and '
                                    'optimizer step eager'),
                            l2='per-step working set (>3 GB of fp32 activations) exceeds the 126 MB L2; no flush needed'),
                e2e=dict(value=e2e_value, unit='voxels/s', ms_per_step=ms_e2e / args.steps,
                         h2d_bytes_per_step=(x_host.numel() * 4 + t_host.numel() * 8), d2h_bytes_per_step=4),
                gpu_launches=int(launches), clocks=clocks.summary(), roofline=roofline)
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline_run(steps=1, warmup=0)
        line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
