"""bench.py -- voxels/s of the UNet train step (BASELINE.json configs[1]) and of the Predictor (configs[3]) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-cpu-baseline] [--no-ref-gpu] [--no-predictor]

Headline record (top level of the ONE JSON line rank 0 prints): the train step.  A "step" is one pass of the hot path
over one synthetic batch: forward of UNet(n_blocks=3,start_filts=32,normalization='group') on (4,1,64,64,64) fp32, Dice
loss (the reference's DiceLoss formula, modules/loss.py:165-233, in plain torch: a boundary consumer that stays torch),
backward, SGD step.  The launches of a step are captured once in a CUDA graph and replayed
(elektronn3_b200.GraphedTrainStep; E3B_BENCH_GRAPH=0 launches them eagerly).  N>1 (torchrun): one process per GPU, weak
scaling; the graph ends after backward, then one flat NCCL all-reduce of the gradients and the optimizer step.

`value` = device-resident throughput; `e2e` = through the public module API with pinned HOST input/target copied in and
the loss read back every step.  `roofline`: the dominant kernel timed alone.  `ref_gpu`: the reference's own torch/cuDNN
TF32 path (restated in oracle/torch_ref.py) timed in this process with the same harness -- the ">= 5x cuDNN" yardstick.
`cpu_baseline`: the same restatement on the host cores (torch / oneDNN: the reference's real CPU arithmetic).

`predictor` sub-object: BASELINE cfg 4, Predictor(tile 64^3, overlap 8) of UNet(n_blocks=4) over a 512x512x256 volume
(67.1 M output voxels); under --gpus N the tile grid is sharded over the ranks (strong scaling) and rank 0 assembles the
result.  It carries its own value / e2e / roofline / ref_gpu / cpu_baseline.

`variants` list (N = 1): the same train step for the rows SURVEY.md section 8(f) marks "next" -- the residual U-Net
(elektronn3.models.resunet) and two option sets of the plain one (resize-conv up-sampling + leaky ReLU; merge_mode='add' +
SiLU) -- each next to the reference's torch/cuDNN TF32 call sequence for that model on the same GPU (--no-variants skips it).

`--impl reference`: the CPU arm alone (rank 0 only under torchrun), all host threads, bounded samples.
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
# dominant kernel for the roofline: the conv of up_convs.1.conv1 (virtual concat 32+32 -> 32 at 4 x 64^3)
DOM = dict(N=4, C0=32, C1=32, Co=32, S=64)
DOM_GFLOP = 2 * 4 * 64 ** 3 * 32 * (64 * 27) / 1e9     # 115.96
# torch's fused multi-tensor SGD (one kernel for all parameters) in BOTH GPU arms; E3B_BENCH_FUSED_OPT=0: the foreach default
FUSED_OPT = os.environ.get('E3B_BENCH_FUSED_OPT', '1') != '0'
DOM_TRAFFIC = 302.0e6            # dram__bytes_read + write of that launch: ncu --set full, profiles/r02_ncu_zs_concat_late.csv (cold L2; 268 MB
                                 # algorithmic; 507 MB before the 64-byte L2 promotion and the paired chains)

PRED_MODEL_KW = dict(n_blocks=4, start_filts=32)
PRED_VOL = (512, 512, 256)
PRED_TILE, PRED_OVL = (64, 64, 64), (8, 8, 8)
PRED_WORKLOAD = ('Predictor(tile 64^3, overlap 8^3, softmax fp32 out) of UNet(n_blocks=4,start_filts=32, BN eval) over a '
                 'synthetic 512x512x256 fp32 volume: 256 tiles of 80^3, host volume in -> host result out')
PRED_TILE_GFLOP = 218.73         # forward of one 80^3 tile, SURVEY.md appendix A cfg 4
PRED_DOM_GFLOP_PER_TILE = 2 * 80 ** 3 * 32 * (64 * 27) / 1e9      # up_convs.2.conv1 on one tile: 56.62


def dice_loss(logits, target, eps=1e-4):
    """DiceLoss(apply_softmax=True) of the reference (modules/loss.py:165-233) restated in torch (E3B_BENCH_TORCH_LOSS=1:
    the boundary consumer stays torch, as with the reference's own class); by default the step uses the drop-in
    elektronn3_b200.DiceLoss (two fused CUDA passes over the logits)."""
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


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU arithmetic (torch ATen / oneDNN through oracle/torch_ref.py) on the host cores
# ------------------------------------------------------------------------------------------------------------------
def _cpu_threads():
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(n)          # torchrun exports OMP_NUM_THREADS=1: set the intra-op pool explicitly
    return n


def cpu_train_baseline(steps=2, warmup=1, budget_s=150.0):
    """fwd + DiceLoss + bwd + SGD of the bench model on the CPU (torch/oneDNN fp32): the full (4,1,64^3) batch when
    a probe with one sample says it fits the budget, else one sample.  -> dict with voxels/s"""
    import torch
    import elektronn3_b200 as e3
    from oracle import torch_ref
    cores = _cpu_threads()
    torch.manual_seed(0)
    m = e3.UNet(**MODEL_KW).train()                      # parameter container only: the forward below is plain torch
    opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9)

    def run(shape, k):
        x = torch.randn(shape)
        t = torch.randint(0, 2, (shape[0],) + shape[2:])
        ts = []
        for _ in range(k):
            t0 = time.perf_counter()
            opt.zero_grad(set_to_none=True)
            loss = torch_ref.dice_loss(torch_ref.unet_forward(m, x), t)
            loss.backward()
            opt.step()
            ts.append(time.perf_counter() - t0)
        return ts
    probe = run((1,) + BATCH[1:], 2)[-1]
    shape = BATCH if probe * BATCH[0] * (steps + warmup) < budget_s else (1,) + BATCH[1:]
    ts = run(shape, steps + warmup)[warmup:]
    dt = sum(ts) / len(ts)
    vox = shape[0] * shape[2] * shape[3] * shape[4]
    return dict(value=vox / dt, unit='voxels/s', cores=cores, kind='port',
                sample=f'{len(ts)} train steps (fwd + DiceLoss + bwd + SGD) on {tuple(shape)} with torch {torch.__version__} CPU '
                       f'kernels (oneDNN fp32, {cores} threads): the reference\'s own ATen call sequence restated in '
                       f'oracle/torch_ref.py; {dt:.2f} s/step',
                seconds_per_step=dt)


def cpu_predictor_baseline(tiles=2):
    """the reference's tiled loop (tiled_apply: batch-1 forward + softmax per 80^3 tile) on the CPU for `tiles` tiles"""
    import torch
    import elektronn3_b200 as e3
    from oracle import torch_ref
    cores = _cpu_threads()
    torch.manual_seed(0)
    m = e3.UNet(**PRED_MODEL_KW).eval()
    in_tile = tuple(t + 2 * o for t, o in zip(PRED_TILE, PRED_OVL))
    x = torch.randn((1, 1) + in_tile)
    with torch.no_grad():
        torch_ref.unet_forward(m, x).softmax(1)          # warm-up (oneDNN primitive creation)
        t0 = time.perf_counter()
        for _ in range(tiles):
            torch_ref.unet_forward(m, x).softmax(1)[:, :, 8:-8, 8:-8, 8:-8].contiguous()
        dt = (time.perf_counter() - t0) / tiles
    vox = PRED_TILE[0] * PRED_TILE[1] * PRED_TILE[2]
    return dict(value=vox / dt, unit='voxels/s', cores=cores, kind='port',
                sample=f'{tiles} of the 256 tiles (80^3 in, 64^3 out) through the reference\'s per-tile sequence on torch CPU '
                       f'kernels ({cores} threads), {dt:.2f} s/tile; output voxels/s',
                seconds_per_tile=dt)


def reference_arm(args):
    steps = max(1, min(args.steps, 3))
    warm = 1
    cb = cpu_train_baseline(steps=steps, warmup=warm)
    line = dict(metric='voxels/s', value=cb['value'], unit='voxels/s', n_gpus=args.gpus, steps=steps, warmup=warm,
                ms_per_step=cb['seconds_per_step'] * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic', impl='reference', config=dict(workload=WORKLOAD, sample=cb['sample']),
                cpu_baseline={k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
                e2e=dict(value=cb['value'], unit='voxels/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    if not args.no_predictor:
        pb = cpu_predictor_baseline()
        line['predictor'] = dict(metric='voxels/s', value=pb['value'], unit='voxels/s', impl='reference',
                                 config=dict(workload=PRED_WORKLOAD, sample=pb['sample']),
                                 cpu_baseline={k: pb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
                                 e2e=dict(value=pb['value'], unit='voxels/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-ref-gpu', action='store_true')
    ap.add_argument('--no-predictor', action='store_true')
    ap.add_argument('--no-variants', action='store_true', help='skip the resunet / option variants of the train step')
    ap.add_argument('--profile-steps', type=int, default=0,
                    help='run only this many train steps after one warm-up and exit (for ncu; prints no bench line)')
    ap.add_argument('--profile-predictor', action='store_true', help='one warm Predictor pass on a 128^3 volume and exit (for ncu)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))

    if args.impl == 'reference':
        if rank == 0:
            reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import elektronn3_b200 as e3
    from elektronn3_b200 import _lib, engine

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU path in the product)'
    # a lost rank must not leave the job hanging in a collective until the caller's limit: give up loudly instead
    limit = float(os.environ.get('E3B_BENCH_TIMEOUT', '1500'))

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        """k calls of fn between two CUDA events on the current stream, barrier + synchronize on both sides; max over ranks (ms)"""
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

    if args.profile_predictor:
        m = e3.UNet(**PRED_MODEL_KW).to(dev).eval()
        vol = tuple(int(v) for v in os.environ.get('E3B_PROFILE_VOL', '128,128,128').split(','))     # '512,128,256': rows of 8 tiles
        p = e3.Predictor(m, device=dev, tile_shape=PRED_TILE, overlap_shape=PRED_OVL, offset=(0, 0, 0), out_shape=(2,) + vol)
        x = torch.randn((1, 1) + vol).pin_memory()
        p.predict(x), p.predict(x)
        return

    torch.manual_seed(1234 + rank)
    model = e3.UNet(**MODEL_KW).to(dev).train()
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, fused=FUSED_OPT)
    voxels = BATCH[0] * BATCH[2] * BATCH[3] * BATCH[4]
    # N = 1: the whole step is replayed as one CUDA graph (elektronn3_b200/graph.py).  N > 1: the graph ends after
    # backward; the data-parallel gradient average (one flat NCCL all-reduce: what DDP's buckets compute) and the
    # optimizer step run eagerly, so that no NCCL call is captured (capturing it hung on the 2-GPU box in round 1).
    # E3B_BENCH_GRAPH=0: eager launches, N > 1 around stock DistributedDataParallel.
    use_graph = os.environ.get('E3B_BENCH_GRAPH', '1') != '0'
    criterion = dice_loss if os.environ.get('E3B_BENCH_TORCH_LOSS') == '1' else e3.DiceLoss(apply_softmax=True).to(dev)
    gstep = None
    step_model = model
    params = [p for p in model.parameters()]
    if use_graph:
        if world > 1:                          # same initial weights on every rank (DDP broadcasts them at construction)
            for p in params:
                dist.broadcast(p.data, 0)
        gstep = e3.GraphedTrainStep(model, criterion, opt if world == 1 else None, BATCH, (BATCH[0],) + BATCH[2:])
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
        loss = criterion(step_model(x), t)
        loss.backward()
        opt.step()
        return loss

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

        def e2e_pipelined(k):
            # K steps, each with the H2D copy of ITS pinned host batch and a D2H read of its loss (trainer.py:575), all K
            # copies inside the timed region.  The copy of batch i+1 is started on the copy stream right after replay i
            # is enqueued (GraphedTrainStep.prefetch), so only the first copy is exposed.
            # The loss of every step is copied D2H right behind its replay and READ one step later (step_async): the host
            # enqueues step i + 1 before it blocks on the loss of step i, so the device never waits for a launch.
            gstep.prefetch(x_host, t_host)
            prev = None
            for i in range(k):
                fut = gstep.step_async()
                if i + 1 < k:
                    gstep.prefetch(x_host, t_host)
                if world > 1:                  # the graph ends after backward: gradient average + optimizer step, eagerly
                    grad_sync()
                    opt.step()
                if prev is not None:
                    prev.result()
                prev = fut
            prev.result()
        if gstep is not None:
            e2e_pipelined(3)
            ms_e2e = timed(lambda: e2e_pipelined(args.steps), 1)
            e2e_mode = ('H2D of batch i+1 on a copy stream behind the kernels of step i (GraphedTrainStep.prefetch); the loss of '
                        'every step is copied D2H behind its replay and read by the host one step later (step_async)')
        else:
            for _ in range(2):
                e2e_step()
            ms_e2e = timed(e2e_step, args.steps)
            e2e_mode = 'H2D, step and loss read in sequence on one stream'

        # dominant kernel alone: the conv of up_convs.1.conv1, CUDA events on the launch stream
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
        del q0, q1
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

    # ---- the reference's own torch/cuDNN TF32 path on the same GPU, same harness (rank-local; N = 1 yardstick)
    ref_gpu = None
    if not args.no_ref_gpu and world == 1:
        ref_gpu = ref_gpu_train(torch, dev, timed)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    bf16_peak = peaks.get('bf16_tflops', 1590.0)
    peak_note = ('tcgen05 kind::f16 (fp16 operands, fp32 accumulate) runs at the bf16 rate: peak = MEASURED_PEAKS.json '
                 'bf16_tflops (burst: kernel timed alone)' if peaks else 'fallback 1.59 PF dense bf16 (B200_PROFILING.md)')

    predictor = None
    if not args.no_predictor:
        del gstep, model, opt, x_dev, t_dev
        torch.cuda.empty_cache()
        predictor = predictor_bench(torch, dist, e3, engine, _lib, dev, world, rank, timed, bf16_peak, peak_note, args)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    achieved = DOM_GFLOP / ms_dom            # GFLOP / ms == TFLOP/s
    roofline = dict(bound='tensor', kernel=('conv_zs_kernel' if dom_var else 'conv_tc_kernel') + ' on up_convs.1.conv1 (virtual concat 32+32 -> 32 @ 4x64^3)',
                    achieved=achieved, peak=bf16_peak, unit='TFLOP/s', frac=achieved / bf16_peak,
                    traffic=DOM_TRAFFIC, ms_per_launch=ms_dom, peak_note=peak_note,
                    f16_cublas_tflops_measured_here=f16_peak,
                    step_tensor_frac=FWD_BWD_GFLOP / (ms / args.steps) / bf16_peak)
    line = dict(metric='voxels/s', value=value, unit='voxels/s', n_gpus=world, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f16 operands / f32 accumulate (TF32-equivalent mantissa), f32 storage', data='synthetic',
                config=dict(workload=WORKLOAD, global_batch=BATCH[0] * world,
                            parallelism=f'dp{world}' if world > 1 else 'single',
                            launch=('eager launches' if not use_graph else
                                    'whole step replayed as one CUDA graph (GraphedTrainStep)' if world == 1 else
                                    'forward+loss+backward replayed as one CUDA graph, flat NCCL all-reduce and '
                                    'optimizer step eager'),
                            l2='per-step working set (>3 GB of fp32 activations) exceeds the 126 MB L2; no flush needed'),
                e2e=dict(value=e2e_value, unit='voxels/s', ms_per_step=ms_e2e / args.steps,
                         h2d_bytes_per_step=(x_host.numel() * 4 + t_host.numel() * 8), d2h_bytes_per_step=4, mode=e2e_mode),
                gpu_launches=int(launches), clocks=clocks.summary(), roofline=roofline)
    if ref_gpu is not None:
        ref_gpu['speedup_device'] = ref_gpu['ms_per_step'] / (ms / args.steps)
        ref_gpu['speedup_e2e'] = ref_gpu['ms_per_step_e2e'] / (ms_e2e / args.steps)
        line['ref_gpu'] = ref_gpu
    if predictor is not None:
        line['predictor'] = predictor
    if world == 1 and not args.no_ref_gpu and not args.no_variants:
        line['variants'] = variants_bench(torch, e3, dev, timed)
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_train_baseline(steps=1, warmup=1)
        line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
        if predictor is not None:
            pb = cpu_predictor_baseline()
            predictor['cpu_baseline'] = {k: pb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


VARIANTS = [
    ('resunet.UNet(n_blocks=3,start_filts=32,norm=GN,enc_res_blocks=1,dec_res_blocks=1)', 'resunet',
     dict(n_blocks=3, start_filts=32, normalization='group', enc_res_blocks=1, dec_res_blocks=1)),
    ("UNet(n_blocks=3,start_filts=32,norm=GN,up_mode='resizeconv_nearest',activation='leaky')", 'unet',
     dict(n_blocks=3, start_filts=32, normalization='group', up_mode='resizeconv_nearest', activation='leaky')),
    ("UNet(n_blocks=3,start_filts=32,norm=GN,merge_mode='add',activation='silu')", 'unet',
     dict(n_blocks=3, start_filts=32, normalization='group', merge_mode='add', activation='silu')),
]


def variants_bench(torch, e3, dev, timed):
    """The rows SURVEY.md section 8(f) marks "next", measured like the headline: the cfg-2 train step (fwd + DiceLoss + bwd + SGD on
    (4,1,64^3), device-resident batch) of the residual U-Net and of two option sets, replayed as one CUDA graph, next to the
    reference's torch/cuDNN TF32 call sequence for the same model on the same GPU (eager, cudnn.benchmark)."""
    from oracle import torch_ref
    out = []
    vox = BATCH[0] * BATCH[2] * BATCH[3] * BATCH[4]
    tshape = (BATCH[0],) + BATCH[2:]
    x = torch.randn(BATCH, device=dev)
    t = torch.randint(0, 2, tshape, device=dev)
    for name, arch, kw in VARIANTS:
        cls = e3.resunet.UNet if arch == 'resunet' else e3.UNet
        torch.manual_seed(99)
        m = cls(**kw).to(dev).train()
        opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9, fused=FUSED_OPT)
        gstep = e3.GraphedTrainStep(m, e3.DiceLoss(apply_softmax=True).to(dev), opt, BATCH, tshape)
        for _ in range(3):
            gstep(x, t)
        k = 10
        ms = timed(lambda: gstep(x, t), k) / k
        del gstep, opt
        old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32)
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32 = True, True
        try:
            opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9, fused=FUSED_OPT)

            def step():
                opt.zero_grad(set_to_none=True)
                loss = torch_ref.dice_loss(torch_ref.unet_forward(m, x), t)
                loss.backward()
                opt.step()
            for _ in range(4):
                step()
            k = 5
            ms_ref = timed(step, k) / k
        finally:
            torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32 = old
        out.append(dict(workload=f'{name} train step (fwd+DiceLoss+bwd+SGD) on synthetic {tuple(BATCH)} fp32',
                        value=vox / ms * 1e3, unit='voxels/s', ms_per_step=ms,
                        ref_gpu_ms_per_step=ms_ref, speedup_device=ms_ref / ms))
        del m, opt
        torch.cuda.empty_cache()
    return out


def ref_gpu_train(torch, dev, timed):
    """The reference's GPU path for the same step: its ATen call sequence (oracle/torch_ref.py) on cuDNN with TF32
    convolutions (the default of this torch build) and cudnn.benchmark, like benchmark/train_benchmark.py."""
    import elektronn3_b200 as e3
    from oracle import torch_ref
    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32)
    torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32 = True, True
    try:
        m = e3.UNet(**MODEL_KW).to(dev).train()          # parameter container; the forward below is plain torch / cuDNN
        opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9, fused=FUSED_OPT)
        x = torch.randn(BATCH, device=dev)
        t = torch.randint(0, 2, (BATCH[0],) + BATCH[2:], device=dev)
        xh, th = torch.randn(BATCH).pin_memory(), torch.randint(0, 2, (BATCH[0],) + BATCH[2:]).pin_memory()

        def step(xx, tt):
            opt.zero_grad(set_to_none=True)
            loss = torch_ref.dice_loss(torch_ref.unet_forward(m, xx), tt)
            loss.backward()
            opt.step()
            return loss
        for _ in range(5):
            step(x, t)
        k = 10
        ms = timed(lambda: step(x, t), k) / k
        ms_e2e = timed(lambda: float(step(xh.to(dev, non_blocking=True), th.to(dev, non_blocking=True))), k) / k
        vox = BATCH[0] * BATCH[2] * BATCH[3] * BATCH[4]
        return dict(what='the reference\'s torch/cuDNN path (TF32 convolutions, cudnn.benchmark, NCDHW fp32) on this GPU, same '
                         'workload and timing harness, eager launches like Trainer._train_step',
                    value=vox / ms * 1e3, unit='voxels/s', ms_per_step=ms, ms_per_step_e2e=ms_e2e, e2e_value=vox / ms_e2e * 1e3)
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32 = old
        torch.cuda.empty_cache()


def predictor_bench(torch, dist, e3, engine, _lib, dev, world, rank, timed, bf16_peak, peak_note, args):
    """BASELINE cfg 4 through elektronn3_b200.Predictor; sharded over the ranks when world > 1 (strong scaling)."""
    torch.manual_seed(7)
    m = e3.UNet(**PRED_MODEL_KW).to(dev).eval()
    if world > 1:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, 0)
    vol_host = torch.randn((1, 1) + PRED_VOL).pin_memory()
    out_vox = PRED_VOL[0] * PRED_VOL[1] * PRED_VOL[2]
    kw = dict(device=dev, tile_shape=PRED_TILE, overlap_shape=PRED_OVL, offset=(0, 0, 0), out_shape=(2,) + PRED_VOL,
              apply_softmax=True)
    p_e2e = e3.Predictor(m, **kw)
    p_dev = e3.Predictor(m, return_device=True, **kw)
    vol_dev = vol_host.to(dev)
    reps = 3
    p_dev.predict(vol_dev), p_e2e.predict(vol_host)                # warm-up: kernels loaded, pinned result buffer cached
    l0 = _lib.launch_count()
    ms_dev = timed(lambda: p_dev.predict(vol_dev), reps) / reps
    launches = (_lib.launch_count() - l0) // reps
    ms_e2e = timed(lambda: p_e2e.predict(vol_host), reps) / reps
    st = dict(p_e2e.last_stats)
    # dominant kernel alone: the eval-mode conv of up_convs.2.conv1 on one tile batch (B x 80^3, 32+32 -> 32, folded BN + ReLU)
    S = PRED_TILE[0] + 2 * PRED_OVL[0]
    B = p_dev.tile_batch or p_dev._auto_tile_batch(32, 1, (S, S, S))        # what one forward pass of the run above held
    q0 = engine.QP.empty_half(B, 32, S, S, S, dev)
    q1 = engine.QP.empty_half(B, 32, S, S, S, dev)
    q0.t.normal_(), q1.t.normal_()
    w = torch.randn(32, 64, 3, 3, 3, device=dev) * 0.05
    bias = torch.zeros(32, device=dev)
    var = engine.conv_variant(32, 32, 32, (3, 3, 3))
    wpk = engine.pack_weights(4 if var else 0, w, None, 32, 32, 32, (3, 3, 3))

    def dom():
        engine.conv_forward(q0, wpk, 32, 32, (3, 3, 3), (1, 1, 1), src1=q1, bias=bias, relu=True, half_out=True, variant=var)
    for _ in range(3):
        dom()
    ms_dom = timed(dom, 10) / 10
    del q0, q1
    ref_gpu = None
    if not args.no_ref_gpu and world == 1:
        ref_gpu = ref_gpu_predictor(torch, dev, m, timed)
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    achieved = B * PRED_DOM_GFLOP_PER_TILE / ms_dom
    pred = dict(metric='voxels/s', value=out_vox / ms_dev * 1e3, unit='voxels/s', n_gpus=world, steps=reps, seconds_per_volume=ms_dev * 1e-3,
                higher_is_better=True, scaling='strong', dtype='f16 operands / f32 accumulate', data='synthetic',
                config=dict(workload=PRED_WORKLOAD, tile_batch=int(B), tiles=256,
                            parallelism=(f'tile rows of the 8x8x4 tile grid sharded over {world} ranks, one NCCL gather of the '
                                         f'output slabs to rank 0') if world > 1 else 'single',
                            value_is='volume resident in HBM, result left in HBM', l2='268 MB volume + 537 MB result exceed the 126 MB L2'),
                e2e=dict(value=out_vox / ms_e2e * 1e3, unit='voxels/s', seconds_per_volume=ms_e2e * 1e-3,
                         h2d_bytes_per_step=st.get('h2d_bytes'), d2h_bytes_per_step=st.get('d2h_bytes')),
                gpu_launches=int(launches),
                roofline=dict(bound='tensor', kernel=('conv_zs_kernel' if var else 'conv_tc_kernel') + f' on up_convs.2.conv1, one tile batch (32+32 -> 32 @ {B}x80^3, folded BN + ReLU, fp16 out)',
                              achieved=achieved, peak=bf16_peak, unit='TFLOP/s', frac=achieved / bf16_peak, traffic=None,
                              ms_per_launch=ms_dom, peak_note=peak_note,
                              volume_tensor_frac=256 * PRED_TILE_GFLOP / ms_dev / bf16_peak / world))
    if ref_gpu is not None:
        ref_gpu['speedup_e2e'] = ref_gpu['seconds_per_volume_extrapolated'] / (ms_e2e * 1e-3)
        pred['ref_gpu'] = ref_gpu
    return pred


def ref_gpu_predictor(torch, dev, m, timed, tiles=32):
    """The reference Predictor's per-tile sequence (inference.py:179-197: host slice -> H2D -> forward + softmax at batch
    1 on cuDNN TF32 -> crop -> D2H into the host output) for `tiles` of the 256 tiles, extrapolated to the volume."""
    from oracle import torch_ref
    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32)
    torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32 = True, True
    try:
        S = PRED_TILE[0] + 2 * PRED_OVL[0]
        padded = torch.randn(1, 1, S + 64, S + 64, S + 64)          # pageable host memory like tiled_apply's padded input
        out = torch.empty(1, 2, 128, 128, 128)

        def one(i):
            a = (i % 2) * 64
            tile = padded[:, :, a:a + S, a:a + S, a:a + S].contiguous()
            with torch.no_grad():
                r = torch_ref.unet_forward(m, tile.to(dev)).softmax(1)[:, :, 8:-8, 8:-8, 8:-8]
            out[:, :, a:a + 64, a:a + 64, a:a + 64] = r      # implicit-sync D2H like inference.py:197
        for i in range(4):
            one(i)
        ms = timed(lambda: [one(i) for i in range(tiles)], 1) / tiles
        vox = PRED_TILE[0] * PRED_TILE[1] * PRED_TILE[2]
        return dict(what=f'the reference Predictor\'s per-tile sequence on torch/cuDNN TF32 on this GPU for {tiles} of the 256 '
                         'tiles (batch 1, pageable H2D, implicit-sync D2H per tile), extrapolated',
                    value=vox / ms * 1e3, unit='voxels/s', ms_per_tile=ms, seconds_per_volume_extrapolated=ms * 256 * 1e-3)
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32 = old


if __name__ == '__main__':
    main()
