"""Fused AdamW over all parameters of a group in ONE kernel launch, which also refreshes the bf16 GEMM operands.

Drop-in for the optimiser the reference scripts build (`torch.optim.AdamW(model.parameters(), lr=..., weight_decay=...)`,
train_vit.py:82, train_titok.py:134, train_vit_vqgan.py:131, train_videogpt.py:107) and drive through
`GradScaler.step(optim)` (train_vit.py:105): same constructor arguments, `param_groups`, `state` keys (`step`, `exp_avg`,
`exp_avg_sq`), `state_dict()` layout, LR-scheduler behaviour (lr is read from the group at every step) and GradScaler
protocol (`_step_supports_amp_scaling`: `grad_scale` / `found_inf` are consumed on the device, a step with an inf is skipped
without a host synchronisation).  SURVEY.md §8f-2.

Arithmetic: torch/optim/adam.py `_single_tensor_adam` (decoupled weight decay) in fp32, see csrc/optim.cu.  After the
update the kernel writes the bf16 copy of every GEMM weight (the tensors functional.bf16_of caches), so the next forward
finds its operands ready: no separate cast pass, and -- unlike torch's fused optimisers, which do not bump
`Tensor._version` -- no stale operands.

No CPU path: CPU parameters raise.
"""
import numpy as np
import torch

from . import _cabi
from . import functional as Fn
from . import ops

_REC = 6  # int64 fields per tensor record: p, g, m, v, w16, n  (struct AdamWTensor in csrc/optim.cu)


class AdamW(torch.optim.Optimizer):
    _step_supports_amp_scaling = True   # GradScaler hands us grad_scale / found_inf instead of syncing on found_inf
    _b200_refreshes_bf16 = True         # functional._optimizer_stepped: this optimiser re-keys the operand caches itself

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, *,
                 maximize=False, foreach=None, capturable=False, differentiable=False, fused=None):
        if isinstance(lr, torch.Tensor):
            raise ValueError("b200vit.optim.AdamW: lr must be a Python number (LR schedulers write floats)")
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if amsgrad or maximize or differentiable:
            raise NotImplementedError("b200vit.optim.AdamW implements amsgrad=False, maximize=False, differentiable=False "
                                      "(what every reference script uses)")
        # foreach / fused are accepted for signature compatibility; there is exactly one (fused) implementation.
        # capturable=True keeps the step counter on the device so that step() can be captured in a CUDA graph.
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=bool(capturable), differentiable=False, fused=True)
        super().__init__(params, defaults)
        self._tables = {}   # group index -> cached device tables

    # ------------------------------------------------------------------------------------------------ state
    def _init_state(self, p, capturable):
        st = self.state[p]
        if len(st) == 0:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("b200vit.optim.AdamW: run at least one eager step before capturing step() in a CUDA graph "
                                   "(state created during capture would be re-zeroed by every replay)")
            # same keys / dtypes as torch.optim.AdamW: `step` is a float32 scalar tensor (on the device when capturable)
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device if capturable else "cpu")
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _group_tables(self, gi, group, params):
        """Device tables for one group: tensor records [T, 6] int64 and the chunk list [C, 2] int32.  The chunk list
        depends only on the sizes; the records are re-uploaded when any pointer changed (fresh .grad tensors)."""
        chunk = _cabi.load().b200vit_adamw_chunk_elems()
        sizes = tuple(p.numel() for p in params)
        tab = self._tables.get(gi)
        if tab is None or tab["sizes"] != sizes:
            dev = params[0].device
            pairs = []
            for ti, n in enumerate(sizes):
                k = (n + chunk - 1) // chunk
                pairs.append(np.stack([np.full(k, ti, dtype=np.int32), np.arange(k, dtype=np.int32)], axis=1))
            chunks = np.ascontiguousarray(np.concatenate(pairs, axis=0))
            tab = {
                "sizes": sizes,
                "chunks": torch.from_numpy(chunks).to(dev),
                "n_chunks": int(chunks.shape[0]),
                "dev": torch.zeros(len(sizes), _REC, dtype=torch.int64, device=dev),
                "last": None,
                "shared_step": None,
            }
            self._tables[gi] = tab
        return tab

    # ------------------------------------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale = getattr(self, "grad_scale", None)   # set by GradScaler.step around this call
        found_inf = getattr(self, "found_inf", None)
        if torch.cuda.is_current_stream_capturing():
            # like torch's _cuda_graph_capture_health_check: a host-side step count / bias correction captured into a graph
            # would be replayed unchanged for ever (state["step"] never advances, every replay applies step-1 corrections)
            for group in self.param_groups:
                if not group["capturable"] and found_inf is None and any(p.grad is not None for p in group["params"]):
                    raise RuntimeError("b200vit.optim.AdamW: capturing step() in a CUDA graph needs capturable=True (the step "
                                       "counter and bias corrections must live on the device)")
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            for p in params:
                if not p.is_cuda:
                    raise _cabi.B200VitError("b200vit.optim.AdamW needs CUDA parameters on a B200; there is no CPU path")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError("b200vit.optim.AdamW: parameters and gradients must be contiguous fp32 tensors")
                if p.grad.is_sparse:
                    raise RuntimeError("AdamW does not support sparse gradients")
            device_step = bool(group["capturable"]) or found_inf is not None
            if device_step:
                subsets = [params]          # one device counter per group (documented: the group steps together)
            else:
                # torch keeps a step count per parameter: parameters that joined later (their .grad was None before) have
                # their own bias corrections -> one launch per distinct step count (a single launch in the usual case)
                by_step = {}
                for p in params:
                    st = self.state[p]
                    by_step.setdefault(int(st["step"].item()) if len(st) else 0, []).append(p)
                subsets = [by_step[k] for k in sorted(by_step)]
            for si, subset in enumerate(subsets):
                self._step_subset((gi, si), group, subset, device_step, grad_scale, found_inf)
        return loss

    def _step_subset(self, gi, group, params, device_step, grad_scale, found_inf):
        tab = self._group_tables(gi, group, params)
        rec = np.empty((len(params), _REC), dtype=np.int64)
        caches = []
        for i, p in enumerate(params):
            st = self._init_state(p, device_step)
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            cached = getattr(p, "_b200_bf16", None)
            w16 = cached[1] if cached is not None and cached[1].shape == p.shape and cached[1].is_contiguous() else None
            if w16 is not None:
                caches.append((p, w16))
            rec[i] = (p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                      0 if w16 is None else w16.data_ptr(), p.numel())
            if g is not p.grad:
                caches.append((None, g))  # keep the contiguous copy alive until the launch below
        key = rec.tobytes()
        if tab["last"] != key:
            # a fresh pinned staging block per upload: the caching host allocator does not hand it out again before
            # the asynchronous copy has run, so a later step cannot overwrite records an earlier copy still needs
            staging = torch.from_numpy(rec).pin_memory()
            tab["dev"].copy_(staging, non_blocking=True)
            tab["last"] = key
            if torch.cuda.is_current_stream_capturing():
                # the copy became a graph node that re-reads this host block at every replay: keep it alive
                tab.setdefault("captured_staging", []).append(staging)
        # step counter: the parameters of a group always step together -> one counter; every state["step"] of
        # the group tracks it (host tensors are rewritten below, device tensors alias the shared one)
        step_host, step_dev = 0, None
        if device_step:
            shared = tab["shared_step"]
            if shared is None:
                first = self.state[params[0]]["step"]
                shared = first.to(params[0].device, torch.float32).clone()
                tab["shared_step"] = shared
            for p in params:
                self.state[p]["step"] = shared
            step_dev = shared
        else:
            step_host = int(self.state[params[0]]["step"].item()) + 1
        beta1, beta2 = group["betas"]
        nbytes = sum(p.numel() * (28 + (2 if r[4] else 0)) for p, r in zip(params, rec))
        P = _cabi.ptr
        ops._call("b200vit_adamw_step", params[0], P(tab["dev"]), P(tab["chunks"]), tab["n_chunks"],
                  float(group["lr"]), float(beta1), float(beta2), float(group["eps"]), float(group["weight_decay"]),
                  step_host, P(step_dev), P(grad_scale), P(found_inf), ops.stream_ptr(), hbm_bytes=float(nbytes))
        if not device_step:
            for p in params:
                self.state[p]["step"].fill_(step_host)
        for p, w16 in caches:
            if p is not None:
                p._b200_bf16 = (Fn.bf16_key(p), w16)   # the operand is current again

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}
        # Optimizer.load_state_dict moves `step` to the parameter's device because defaults say fused=True; the host-step
        # path would then pay one .item() sync per parameter and step: keep host counters on the host
        for group in self.param_groups:
            if not group["capturable"]:
                for p in group["params"]:
                    st = self.state.get(p)
                    if st and isinstance(st.get("step"), torch.Tensor) and st["step"].device.type != "cpu":
                        st["step"] = st["step"].detach().to("cpu", torch.float32)


__all__ = ["AdamW"]
