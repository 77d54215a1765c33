"""The core training step captured ONCE in a CUDA graph and replayed.

What is captured is exactly the device work of ``Trainer._train_step`` (training/trainer.py:509-543):
forward, criterion, backward, optimizer step, ``zero_grad(set_to_none=True)`` -- about 170 kernel launches
and 30 memset nodes per step for BASELINE cfg 2.  Launched one by one from Python they leave the GPU idle
for ~8 % of the step (more when the host synchronises on the loss every step, as ``Trainer`` does at
trainer.py:575); replayed as one graph they do not.  The kernels, their order and their arithmetic are those
of the eager path (tests/test_unet_gpu.py::test_graphed_train_step_matches_eager).

    step = GraphedTrainStep(model, criterion, optimizer, inp_shape, target_shape)
    dloss, dout = step(batch['inp'], batch['target'])      # host (pinned) or device tensors

    step.prefetch(next_batch['inp'], next_batch['target']) # optional: H2D of the next batch behind this step's kernels
    dloss, dout = step()

This is an optional accelerator next to the ``nn.Module`` drop-in: ``Trainer`` keeps working unchanged with
the plain module (a maintainer would replace the body of ``_train_step`` by the call above).
"""
import torch


class GraphedTrainStep:
    """forward + criterion + backward + optimizer.step of a fixed-shape batch as one CUDA graph.

    ``model``: an ``elektronn3_b200.UNet`` in train mode on a CUDA device; ``criterion(dout, dtarget)`` any
    capturable torch loss (the reference's ``DiceLoss`` / ``CombinedLoss`` are: pure tensor ops);
    ``optimizer``: a capturable torch optimizer (``SGD``; ``Adam(capturable=True)``), or ``None``: the graph then ends
    after backward, the gradients stay in the parameters' static ``.grad`` tensors and the caller all-reduces them
    and steps the optimizer eagerly (the multi-GPU form: no NCCL call is captured)."""

    def __init__(self, model, criterion, optimizer, inp_shape, target_shape, target_dtype=torch.int64, warmup=3,
                 grad_sync=None):
        """grad_sync (optional, EXPERIMENTAL -- not validated on hardware yet): callable(list of parameters) run between
        backward and optimizer.step inside the graph, e.g. a data-parallel gradient all-reduce.  With it the capture
        uses the thread-local error mode, because NCCL's watchdog thread issues CUDA calls of its own."""
        dev = next(model.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('elektronn3_b200: GraphedTrainStep needs the model on a CUDA device')
        self.model, self.criterion, self.optimizer, self.grad_sync = model, criterion, optimizer, grad_sync
        from . import _lib
        self._copy_stream, self._stage, self._has_staged = None, None, False
        self._loss_slots, self._loss_next = None, 0
        self.inp = torch.zeros(inp_shape, dtype=torch.float32, device=dev)
        self.target = torch.zeros(target_shape, dtype=target_dtype, device=dev)
        # Warm-up on a side stream: lazy initialisation of kernels, allocator and optimizer state happens here, not
        # under capture.  The warm-up steps run on an all-zero batch and must leave NO trace: parameters, BatchNorm
        # statistics / counters and optimizer state are snapshotted and restored IN PLACE afterwards (the graph keeps
        # the addresses), so the first replay starts from exactly the state the caller handed in.
        snap = self._snapshot()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        for p in model.parameters():                   # gradients are then allocated from the graph's pool
            p.grad = None
        self._invalidate()                             # the weight re-packing kernels must be part of the graph
        l0 = _lib.launch_count()
        mode = {} if grad_sync is None else {'capture_error_mode': 'thread_local'}
        with torch.cuda.graph(self.graph, **mode):
            self.dout = model(self.inp)
            self.dloss = criterion(self.dout, self.target)
            self.dloss.backward()
            if grad_sync is not None:
                grad_sync([p for p in model.parameters() if p.grad is not None])
            if optimizer is not None:
                optimizer.step()
        self.launches_per_step = _lib.launch_count() - l0      # libe3b.so kernels inside one replay
        self._restore(snap)
        self._invalidate()

    def _snapshot(self):
        model_t = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        opt_t, opt_py = {}, {}
        if self.optimizer is not None:
            for p, st in self.optimizer.state.items():
                for k, v in st.items():
                    if torch.is_tensor(v):
                        opt_t[(id(p), k)] = v.detach().clone()
                    else:
                        opt_py[(id(p), k)] = v
        return model_t, opt_t, opt_py

    @torch.no_grad()
    def _restore(self, snap):
        model_t, opt_t, opt_py = snap
        for k, v in self.model.state_dict().items():
            v.copy_(model_t[k])
        if self.optimizer is not None:
            for p, st in self.optimizer.state.items():
                for k, v in list(st.items()):
                    if torch.is_tensor(v):
                        # state created by the warm-up (momentum buffers, Adam moments / step) is zeroed in place: for
                        # SGD and Adam a zero state is the state of a fresh optimizer
                        old = opt_t.get((id(p), k))
                        v.copy_(old) if old is not None else v.zero_()
                    elif (id(p), k) in opt_py:
                        st[k] = opt_py[(id(p), k)]
                    elif isinstance(v, (int, float)):
                        st[k] = type(v)(0)

    def _eager_step(self):
        for p in self.model.parameters():
            p.grad = None
        loss = self.criterion(self.model(self.inp), self.target)
        loss.backward()
        if self.grad_sync is not None:
            self.grad_sync([p for p in self.model.parameters() if p.grad is not None])
        if self.optimizer is not None:
            self.optimizer.step()

    def _invalidate(self):
        # the replay rewrites the parameters without bumping their Python-side version counters: drop the packed
        # weight images keyed on them, so that an eager call of the module (validation) re-packs
        self.model.invalidate_weight_cache()

    def prefetch(self, inp, target):
        """Start the H2D copy of the NEXT batch (pinned host tensors) on a copy stream, into staging buffers, while the
        current replay is still running; the next ``step()`` (called without arguments) consumes it with a device-side
        copy.  This is what a ``DataLoader(pin_memory=True)`` + ``.to(device, non_blocking=True)`` pipeline cannot do
        on a single stream (training/trainer.py:515-517): the copy hides behind the previous step's kernels."""
        dev = self.inp.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = (torch.empty_like(self.inp), torch.empty_like(self.target))
            self._staged, self._consumed = torch.cuda.Event(), torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream(dev))
        self._copy_stream.wait_event(self._consumed)          # the staging buffers are free again
        with torch.cuda.stream(self._copy_stream):
            self._stage[0].copy_(inp, non_blocking=True)
            self._stage[1].copy_(target, non_blocking=True)
            self._staged.record(self._copy_stream)
        self._has_staged = True

    def __call__(self, inp=None, target=None):
        """copy the batch into the graph's static buffers (H2D when they are host tensors), replay.
        Without arguments: take the batch staged by ``prefetch``.
        Returns (dloss, dout): static device tensors, overwritten by the next call."""
        if inp is None:
            if not self._has_staged:
                raise RuntimeError('GraphedTrainStep(): no batch given and none staged by prefetch()')
            cur = torch.cuda.current_stream(self.inp.device)
            cur.wait_event(self._staged)
            self.inp.copy_(self._stage[0], non_blocking=True)
            self.target.copy_(self._stage[1], non_blocking=True)
            self._consumed.record(cur)
            self._has_staged = False
        else:
            self.inp.copy_(inp, non_blocking=True)
            self.target.copy_(target, non_blocking=True)
        self.graph.replay()
        self._invalidate()
        return self.dloss, self.dout

    def step_async(self, inp=None, target=None):
        """``__call__`` + an asynchronous D2H copy of the step's loss into a pinned host slot: returns a ``LossFuture`` whose
        ``result()`` waits for THAT copy only.  A loop that enqueues step i + 1 before it asks for the loss of step i keeps
        the device busy while the host reads (``float(dloss)`` right after every step, training/trainer.py:575, idles the
        device for a launch latency per step).  Two slots: at most two futures may be outstanding."""
        self(inp, target)
        if self._loss_slots is None:
            self._loss_slots = [(torch.empty((), dtype=self.dloss.dtype).pin_memory(), torch.cuda.Event()) for _ in range(2)]
            self._loss_next = 0
        host, ev = self._loss_slots[self._loss_next]
        self._loss_next ^= 1
        host.copy_(self.dloss.detach(), non_blocking=True)
        ev.record(torch.cuda.current_stream(self.inp.device))
        return LossFuture(host, ev)


class LossFuture:
    """the loss of one ``GraphedTrainStep.step_async`` call, on its way to the host"""
    __slots__ = ('_host', '_event')

    def __init__(self, host, event):
        self._host, self._event = host, event

    def result(self):
        self._event.synchronize()
        return float(self._host)
