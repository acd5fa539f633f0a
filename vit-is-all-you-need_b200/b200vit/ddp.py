"""Data-parallel training across the GPUs of one box: one process per GPU, parameters replicated, the batch
sharded by rank, gradients averaged by a bucketed all-reduce (NCCL over NVLink 5 / NVSwitch) that overlaps the
rest of backward.

The reference is single-process (no torch.distributed anywhere); this layer is what §8(e) adds.

Mechanics
  * every trainable parameter owns a slot in a flat fp32 bucket (buckets are filled in reverse parameter order,
    i.e. in the order backward produces gradients);
  * the fused transformer backward asks `grad_slot(param)` for its slot and lets the wgrad GEMM / bias
    reductions write straight into bucket memory, then calls `grad_ready(param)`;
  * all other parameters (patch embedding, heads, codebooks ...) are copied into their slots from a
    post-accumulate-grad hook;
  * when the last slot of a bucket is ready an event is recorded on the compute stream and the bucket's
    all-reduce is enqueued on a communication stream;
  * an autograd end-of-backward callback waits for all buckets and points `param.grad` at the (averaged)
    bucket views, so optimisers / GradScaler / clip_grad_norm_ run unchanged.
Gradient accumulation: micro-steps run inside `no_sync()` (gradients accumulate in `param.grad` through autograd, nothing is
reduced); on the following synchronised step every parameter whose `.grad` already holds something is NOT offered a bucket
slot (`grad_slot` -> None), so its new gradient is ADDED to the accumulated one by autograd and the post-accumulate hook
copies the sum into the bucket -- fused and hook-delivered parameters reduce the same quantity.
"""
import contextlib

import torch
import torch.distributed as dist

from . import functional as Fn


class _Bucket:
    def __init__(self, numel, device, dtype):
        self.flat = torch.zeros(numel, device=device, dtype=dtype)
        self.flat16 = None     # bf16 staging of the compressed all-reduce (allocated on first use)
        self.done = None       # event on the communication stream after the cast back (compressed path)
        self.params = []
        self.pending = 0
        self.work = None
        self.ready = set()


class DataParallel(torch.nn.Module):
    def __init__(self, module: torch.nn.Module, bucket_mb: float = 32.0, process_group=None, compress_bf16=None):
        """compress_bf16 (default: environment B200VIT_DDP_BF16=1): each bucket crosses NVLink as bf16 -- cast on the
        communication stream, all-reduced (average), cast back into the fp32 bucket -- i.e. half the bytes for the
        collective, like torch's bf16_compress_hook; the optimizer still sees fp32 gradients."""
        super().__init__()
        import os
        self.compress_bf16 = (os.environ.get("B200VIT_DDP_BF16", "0") == "1") if compress_bf16 is None else bool(compress_bf16)
        # diagnosis only (tools/ddp_ab2.sh): no collective at all, to separate the all-reduce's cost from rank-to-rank skew
        self._skip_comm = os.environ.get("B200VIT_DDP_DIAG_SKIP_COMM", "0") == "1"
        self._avg_op = os.environ.get("B200VIT_DDP_AVG", "0") == "1"
        self.module = module
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.backend = dist.get_backend(process_group) if dist.is_initialized() else "none"
        self._sync = True
        self._in_backward = False
        self._slots = {}
        self.buckets = []
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("DataParallel: module has no trainable parameters")
        device, dtype = params[0].device, params[0].dtype
        cap = int(bucket_mb * 1024 * 1024 / params[0].element_size())
        groups, cur, cur_n = [], [], 0
        for p in reversed(params):
            if cur and cur_n + p.numel() > cap:
                groups.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            groups.append(cur)
        for g in groups:
            b = _Bucket(sum(p.numel() for p in g), device, dtype)
            off = 0
            for p in g:
                self._slots[p] = (b, b.flat[off:off + p.numel()].view(p.shape))
                b.params.append(p)
                off += p.numel()
            self.buckets.append(b)
        self.comm_stream = torch.cuda.Stream(device=device) if device.type == "cuda" else None
        for p in params:
            p.register_post_accumulate_grad_hook(self._hook)
        self.broadcast_parameters()

    # ---------------------------------------------------------------------------------------------- plumbing
    def broadcast_parameters(self):
        """Rank 0's parameters / buffers to every rank.  Writes through `.data` do not bump Tensor._version, so the cached
        bf16 GEMM operands of the parameters are dropped explicitly (call this again after loading a checkpoint on rank 0)."""
        if self.world > 1:
            for t in list(self.module.parameters()) + list(self.module.buffers()):
                dist.broadcast(t.data, src=0, group=self.pg)
        Fn.invalidate_weight_cache(self.module)

    def forward(self, *args, **kwargs):
        # the fused autograd nodes created during this forward capture the sink; other models in the process
        # (a second wrapper, an un-wrapped teacher ...) are unaffected
        if self._in_backward and torch.is_grad_enabled():
            # a backward that raised never reached _finalize (the engine drops its callbacks): without this reset no later
            # backward would re-arm the buckets and the ranks would silently diverge
            self._reset_backward_state()
        prev = Fn.set_grad_sink(self if (self._sync and torch.is_grad_enabled()) else None)
        try:
            return self.module(*args, **kwargs)
        finally:
            Fn.set_grad_sink(prev)

    @contextlib.contextmanager
    def no_sync(self):
        old, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = old

    # ---------------------------------------------------------------------------------- gradient sink protocol
    def grad_slot(self, param):
        """Bucket view the producer may OVERWRITE with this parameter's gradient, or None when the gradient has to go
        through autograd: unknown parameter; `.grad` already populated (accumulation after no_sync() micro-steps, or a
        caller that did not zero the gradients); or the parameter was already delivered in this backward (weight sharing /
        a wrapped stack that ran twice in one forward -- its bucket may be in flight)."""
        s = self._slots.get(param)
        if s is None or not self._sync or param.grad is not None:
            return None
        b, view = s
        if self._in_backward and param in b.ready:
            return None
        return view

    def grad_ready(self, param):
        if param in self._slots:
            self._mark_ready(param)

    def _hook(self, param):
        if not self._sync:
            return
        b, view = self._slots[param]
        if self._in_backward and param in b.ready:
            if param.grad is None or param.grad.data_ptr() == view.data_ptr():
                return  # delivered through the gradient sink: torch fires this hook even for the None autograd received
            # a second gradient for a parameter whose bucket slot has been handed over already (the all-reduce may be in
            # flight): it cannot be folded in any more -- fail instead of training on a partial gradient
            raise RuntimeError("b200vit.ddp.DataParallel: a parameter received a second gradient in the same backward after "
                               "its all-reduce bucket was released (shared weights / a fused stack used twice in one forward "
                               "are not supported under DataParallel)")
        if param.grad is not None and param.grad.data_ptr() != view.data_ptr():
            view.copy_(param.grad)
        self._mark_ready(param)

    def _reset_backward_state(self):
        self._in_backward = False
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()
            b.pending = len(b.params)
            b.work = None
            b.ready = set()

    def _begin_backward(self):
        self._in_backward = True
        for b in self.buckets:
            b.pending = len(b.params)
            b.work = None
            b.ready = set()
        torch.autograd.Variable._execution_engine.queue_callback(self._finalize)

    def _mark_ready(self, param):
        if not self._sync:
            return
        if not self._in_backward:
            self._begin_backward()
        b, _ = self._slots[param]
        if param in b.ready:
            return
        b.ready.add(param)
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b):
        if self.world == 1 or self._skip_comm:
            return
        if self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                if self.compress_bf16 and self.backend == "nccl":
                    from . import ops
                    if b.flat16 is None:
                        b.flat16 = torch.empty(b.flat.numel(), device=b.flat.device, dtype=torch.bfloat16)
                    ops.cast_bf16(b.flat, out=b.flat16)
                    # SUM (not AVG): NCCL's in-switch NVLS reduction has no pre-multiplied average for bf16; the 1 / world
                    # factor is applied for free by the cast back into the fp32 bucket
                    if self._avg_op:      # A/B knob B200VIT_DDP_AVG=1: NCCL's own average (measured: RING_LL instead of NVLS)
                        dist.all_reduce(b.flat16, op=dist.ReduceOp.AVG, group=self.pg, async_op=True).wait()
                        ops.cast_f32_from_bf16(b.flat16, b.flat)
                    else:
                        dist.all_reduce(b.flat16, op=dist.ReduceOp.SUM, group=self.pg, async_op=True).wait()   # comm stream waits
                        ops.cast_f32_from_bf16(b.flat16, b.flat, scale=1.0 / self.world)
                    b.done = torch.cuda.Event()
                    b.done.record(self.comm_stream)
                else:
                    b.work = self._all_reduce(b.flat)
        else:
            b.work = self._all_reduce(b.flat)

    def _all_reduce(self, flat):
        if self.backend == "nccl":
            return dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.pg, async_op=True)
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        work.wait()
        flat.div_(self.world)
        return None

    def _finalize(self):
        self._in_backward = False
        for b in self.buckets:
            if b.pending != 0:
                # parameters that did not take part in this backward: reduce what is there (zeros for them)
                for p in b.params:
                    if p not in b.ready:
                        self._slots[p][1].zero_()
                self._launch(b)
            if b.work is not None:
                b.work.wait()  # compute stream waits for the NCCL stream
                b.work = None
            if b.done is not None:
                torch.cuda.current_stream().wait_event(b.done)
                b.done = None
        for p, (b, view) in self._slots.items():
            if p in b.ready:
                p.grad = view
