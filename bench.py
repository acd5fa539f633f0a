#!/usr/bin/env python
"""ViT-B/16 224px bf16 training throughput on 1..8 B200s (BASELINE.json configs[1]) through the drop-in modules.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B]        # our arm (the headline: --workload vit_b)
    python bench.py --workload {vit_l,vit_ti,titok_s,tatitok_s,videogpt_b,vq} ...   # the other configs (bench_workloads.py)
    python bench.py --impl reference ...      # the reference's own CPU path (unmodified modules from baseline/_ref)

One "step" = forward + cross-entropy + backward + AdamW on a per-GPU batch of synthetic images (weak scaling).
`value` is measured with the inputs already resident in HBM; `e2e` runs the same step from pinned HOST buffers
(images + labels copied host->device and the loss read back device->host inside the timed region, every step).
Prints ONE JSON line on rank 0.  See DESIGN.md §Measurement for the roofline arithmetic.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "vit-is-all-you-need_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "vit_b16_224_train_images_per_sec"
UNIT = "images/s"
IMAGE, PATCH, CLASSES = 224, 16, 1000
D, LAYERS, HEADS, NTOK = 768, 12, 12, 197


def train_flops_per_image():
    """3 x (layers x F_layer + F_head) + 2 x F_patch, 2mnk per GEMM, no recompute counted."""
    f_layer = 2 * NTOK * D * 3 * D + 4 * NTOK * NTOK * D + 16 * NTOK * D * D
    f_patch = 2 * (NTOK - 1) * (3 * PATCH * PATCH) * D
    f_head = 2 * D * CLASSES
    return 3 * (LAYERS * f_layer + f_head) + 2 * f_patch


def load_gemm_traffic():
    """Per-launch DRAM bytes (read + write) of the tcgen05 GEMM family from the committed `ncu --set full` capture
    (profiles/r1_gemm_traffic.json, written by tools/ncu_step_summary.py); None when no capture is committed."""
    for name in ("r2_gemm_traffic.json", "r1_gemm_traffic.json"):   # the newest committed capture
        try:
            return float(json.load(open(os.path.join(ROOT, "profiles", name)))["dram_bytes_per_launch"])
        except Exception:
            continue
    return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p.get("bf16_tflops"), "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p.get("hbm_gbs"), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            # the timed region was shorter than one sampling period: fall back to a single query taken right
            # after it (the GPU is still warm; this under-reports the load clocks and is marked as such)
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=10).stdout
                f = [x.strip() for x in out.strip().split(",")]
                rs = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                        f[5:9]) if v.lower().startswith("active")]
                return {"sm_mhz": float(f[1]), "sm_max_mhz": float(f[2]), "reasons": rs, "samples": 0,
                        "note": "single query right after the timed region"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU baseline
def numpy_vit_b_params(rng):
    import numpy as np

    def w(*s):
        return (rng.standard_normal(s) * 0.02).astype(np.float32)
    layers = [{"qkv_w": w(3 * D, D), "qkv_b": w(3 * D), "fc1_w": w(4 * D, D), "fc1_b": w(4 * D),
               "fc2_w": w(D, 4 * D), "fc2_b": w(D)} for _ in range(LAYERS)]
    return {"conv_w": w(D, 3, PATCH, PATCH), "conv_b": w(D), "pos_emb": w(NTOK - 1, D), "extra_emb": w(1, D),
            "layers": layers, "head_w": w(CLASSES, D), "head_b": w(CLASSES)}


def cpu_reference_step_rate(batch, steps, warmup):
    """The reference's CPU implementation of the path == the numpy oracle port (fwd + CE + bwd + AdamW, fp32, all host
    threads through the BLAS numpy links): the same step the GPU arm times.  Returns (images/s, seconds per step)."""
    import numpy as np
    from oracle import adamw_oracle as A
    from oracle import vit_oracle as O
    rng = np.random.default_rng(0)
    P = numpy_vit_b_params(rng)
    x = rng.standard_normal((batch, 3, IMAGE, IMAGE)).astype(np.float32)
    labels = rng.integers(0, CLASSES, size=(batch,))

    def leaves(tree):   # (container, key) of every parameter array, in a fixed order
        out = [(tree, k) for k in ("conv_w", "conv_b", "pos_emb", "extra_emb", "head_w", "head_b")]
        for layer in tree["layers"]:
            out += [(layer, k) for k in ("qkv_w", "qkv_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")]
        return out

    state = {}
    n_done = [0]

    def one_step():
        _, _, g = O.vit_classifier_loss_and_grads(x, labels, P, HEADS)
        n_done[0] += 1
        for i, ((pc, k), (gc, _)) in enumerate(zip(leaves(P), leaves(g))):
            m, v = state.get(i, (None, None))
            if m is None:
                m, v = np.zeros_like(pc[k]), np.zeros_like(pc[k])
            grad = np.asarray(gc[k], dtype=np.float32).reshape(pc[k].shape)
            pc[k], m, v, _ = A.adamw_step(pc[k], grad, m, v, n_done[0], 1e-4, weight_decay=1e-2)
            state[i] = (m, v)

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt


def _reference_vit_b(torch, device):
    """The UNMODIFIED reference model of configs[1] (train_vit.py:30-53 from baseline/_ref) with its own default init."""
    from baseline import loader
    ref = loader.load(("transformer", "train_vit"))
    torch.manual_seed(0)
    cfg = ref.train_vit.ViTConfig(IMAGE, 3, PATCH, "B", 1, 0.0)
    return ref.train_vit.ViTClassifier(cfg, num_classes=CLASSES).to(device)


def cpu_true_reference_step_rate(batch, steps, warmup):
    """The reference's own CPU path: train_vit.py:99-105 (forward under autocast("cuda") -- which disables itself without a
    GPU, i.e. fp32 --, CrossEntropyLoss, backward, torch.optim.AdamW) through the unmodified modules, all host threads
    (torchrun exports OMP_NUM_THREADS=1: the thread count is set explicitly).  Returns (images/s, s/step, threads)."""
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = _reference_vit_b(torch, "cpu")
    optim = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-2)
    loss_fn = torch.nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, 3, IMAGE, IMAGE, generator=g)
    y = torch.randint(0, CLASSES, (batch,), generator=g)

    def one_step():
        optim.zero_grad()
        loss = loss_fn(model(x), y)
        loss.backward()
        optim.step()
        return float(loss.detach())

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt, torch.get_num_threads()


def gpu_eager_reference_step_rate(torch, device, batch, steps, warmup):
    """Same-box GPU comparator (SURVEY.md §8d, §2b "the kernel to beat"): the UNMODIFIED reference modules on this B200 through
    PyTorch eager -- cuBLASLt GEMMs, cuDNN conv, ATen LayerNorm / GELU / SDPA -- under torch.autocast("cuda", torch.bfloat16),
    torch.optim.AdamW(fused=True); the step of train_vit.py:99-105 at the bench batch, device-resident inputs, CUDA events."""
    model = _reference_vit_b(torch, device)
    optim = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-2, fused=True)
    loss_fn = torch.nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, 3, IMAGE, IMAGE, generator=g).to(device)
    y = torch.randint(0, CLASSES, (batch,), generator=g).to(device)

    def one_step():
        optim.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = loss_fn(model(x), y)
        loss.backward()
        optim.step()

    for _ in range(warmup):
        one_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model, optim
    torch.cuda.empty_cache()
    return batch / (ms / 1e3), ms


def cpu_baseline_sample(batch, steps, warmup):
    """(images/s, s/step, cores, kind, description): the unmodified reference when baseline/_ref is present, else the
    numpy oracle port."""
    from baseline import loader
    if loader.available():
        ips, dt, threads = cpu_true_reference_step_rate(batch, steps, warmup)
        return ips, dt, threads, "reference", (f"{steps} step(s) of batch {batch} after {warmup} warm-up (ViT-B/16 224, fwd+CE+bwd+AdamW, "
                                               f"unmodified reference modules train_vit.py:30-53 on CPU fp32, torch {threads} threads)")
    ips, dt = cpu_reference_step_rate(batch, steps, warmup)
    return ips, dt, os.cpu_count(), "port", f"{steps} step(s) of batch {batch} (ViT-B/16 224, fwd+CE+bwd+AdamW, fp32 numpy oracle port)"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    batch = args.cpu_batch
    steps = max(1, min(args.steps, 4))
    warm = 1 if args.warmup > 0 else 0
    ips, dt, cores, kind, sample = cpu_baseline_sample(batch, steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        # the SAME workload identity as the GPU arm (its config keys), timed as a bounded sample on the host cores
        "config": {"workload": "ViT-B/16 224px ImageNet-shape train step (fwd + CE + bwd + AdamW), configs[1]",
                   "per_gpu_batch": args.batch, "global_batch": args.batch * max(args.gpus, 1), "seq_len": NTOK,
                   "parallelism": f"dp{max(args.gpus, 1)}", "dropout": 0.0,
                   "optimizer": "torch.optim.AdamW" if kind == "reference" else "numpy AdamW oracle",
                   "train_gflop_per_image": train_flops_per_image() / 1e9},
        "reference_arm": {"what": ("the reference's own CPU path (unmodified modules from baseline/_ref)" if kind == "reference"
                                   else "numpy oracle port of the reference (baseline/_ref absent)"),
                          "sample_batch": batch, "runs_on": "host cores of rank 0"},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
def build_model(torch, M, device, dropout=0.0):
    torch.manual_seed(0)
    cfg = M.ViTConfig(IMAGE, 3, PATCH, "B", 1, dropout)
    model = M.ViTClassifier(cfg, num_classes=CLASSES).to(device)
    return model


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from b200vit import ddp as b200_ddp
    from b200vit import modules as M
    from b200vit import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    B = args.batch
    model = build_model(torch, M, device, args.dropout)
    bucket_mb = args.bucket_mb if args.bucket_mb > 0 else 1e9
    wrapped = (b200_ddp.DataParallel(model, bucket_mb=bucket_mb, compress_bf16=(args.grad_compress == "bf16"))
               if world > 1 else model)
    from b200vit import optim as b200_optim
    if args.optimizer == "torch-fused":
        optim = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-2, fused=True)
    else:   # one fused multi-tensor launch that also refreshes the bf16 GEMM operands (csrc/optim.cu)
        optim = b200_optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-2)
    loss_fn = M.CrossEntropyLoss()   # drop-in for nn.CrossEntropyLoss() (train_vit.py:81): fused kernels, bf16 logits in

    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    n_host = 2
    host_images = [torch.randn(B, 3, IMAGE, IMAGE, generator=g).pin_memory() for _ in range(n_host)]
    host_labels = [torch.randint(0, CLASSES, (B,), generator=g).pin_memory() for _ in range(n_host)]
    dev_images = [h.to(device) for h in host_images]
    dev_labels = [h.to(device) for h in host_labels]

    def step(images, labels, after_forward=None):
        optim.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pred = wrapped(images)
            loss = loss_fn(pred, labels)
        if after_forward is not None:
            after_forward(loss)
        loss.backward()
        optim.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident arm
    def dev_step(i):
        step(dev_images[i % n_host], dev_labels[i % n_host])

    for i in range(max(args.warmup, 3)):
        dev_step(i)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ops.launch_count = 0
    ms = timed(dev_step, args.steps)
    launches = ops.launch_count
    clk = clocks.stop() if rank == 0 else None
    ips = world * B * args.steps / (ms / 1e3)

    # ---- end-to-end arm: pinned host buffers, H2D prefetch on a copy stream, loss read back every step
    copy_stream = torch.cuda.Stream(device=device)
    state = {}

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            im = host_images[i % n_host].to(device, non_blocking=True)
            lb = host_labels[i % n_host].to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        state["next"] = (im, lb, ev)

    losses = []
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
    loss_ev = torch.cuda.Event()

    def read_back(loss):
        # device -> host copy of THIS step's loss, enqueued as soon as the forward has produced it
        loss_host.copy_(loss.detach(), non_blocking=True)
        loss_ev.record()

    def e2e_step(i):
        if "next" not in state:
            prefetch(i)
        im, lb, ev = state.pop("next")
        torch.cuda.current_stream().wait_event(ev)
        im.record_stream(torch.cuda.current_stream()); lb.record_stream(torch.cuda.current_stream())
        prefetch(i + 1)
        step(im, lb, after_forward=read_back)
        loss_ev.synchronize()            # the host reads the step's loss every step (the copy finished mid-step,
        losses.append(float(loss_host))  # so the host does not drain the GPU's queue at the step boundary)

    for i in range(3):
        e2e_step(i)
    state.clear()
    ms_e2e = timed(e2e_step, args.steps)
    ips_e2e = world * B * args.steps / (ms_e2e / 1e3)
    h2d = B * 3 * IMAGE * IMAGE * 4 + B * 8
    d2h = 4

    # ---- roofline of the dominant kernel (the tcgen05 GEMM): per-launch CUDA-event timing over extra steps
    peaks = load_peaks()
    gemm_ms, gemm_flops, gemm_calls = ops.profile_gemms(lambda: dev_step(0), steps=2)
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    # the memory-bound pieces (LayerNorm fwd/bwd with their fused residual adds): algorithmic bytes / event time
    ln_ms, ln_bytes, ln_calls, ln_detail = ops.profile_kernels(lambda: dev_step(0), steps=2, kind="bytes")
    peak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
    step_tflops = ips / world * train_flops_per_image() / 1e12

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return 0

    line = {
        "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "ViT-B/16 224px ImageNet-shape train step (fwd + CE + bwd + AdamW), configs[1]",
                   "per_gpu_batch": B, "global_batch": B * world, "seq_len": NTOK, "parallelism": f"dp{world}", "dropout": args.dropout,
                   "grad_exchange": None if world == 1 else (("one bucket reduced after backward" if args.bucket_mb <= 0 else f"{args.bucket_mb:g} MB buckets overlapped with backward")
                                                             + (", bf16-compressed NCCL all-reduce" if args.grad_compress == "bf16" else ", fp32 NCCL all-reduce")),
                   "optimizer": "torch.optim.AdamW(fused=True)" if args.optimizer == "torch-fused" else "b200vit.optim.AdamW (fused multi-tensor + bf16 operand refresh)", "l2": "per-step working set >> 126 MB L2 (no flush needed)",
                   "train_gflop_per_image": train_flops_per_image() / 1e9},
        "e2e": {"value": ips_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all Linear fwd/dgrad/wgrad launches of a step)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
                     "traffic": load_gemm_traffic(), "traffic_unit": "bytes/launch (dram read+write, ncu)",
                     "flops_per_launch": gemm_flops / gemm_calls if gemm_calls else None, "launches_timed": gemm_calls,
                     "step_tflops_per_gpu": step_tflops, "step_frac_of_peak": step_tflops / peak if peak else None,
                     "step_frac_of_nominal_2250": step_tflops / 2250.0},
        "hbm_kernels": {"what": "LayerNorm fwd (+residual add) / bwd (+residual-gradient add) and the fused AdamW (+bf16 operand refresh), algorithmic bytes / CUDA-event time",
                        "achieved": ln_bytes / (ln_ms / 1e3) / 1e9 if ln_ms > 0 else None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": (ln_bytes / (ln_ms / 1e3) / 1e9) / peaks["hbm_gbs"] if ln_ms > 0 and peaks["hbm_gbs"] else None,
                        "launches_timed": ln_calls, "ms_per_step": ln_ms / 2,
                        "note": "part of each input is still L2-resident from its producer, so the fraction can exceed 1",
                        "detail": {k.replace("b200vit_", ""): {"GBps": v["rate"] / 1e9, "launches": v["launches"], "ms_per_step": v["ms"] / 2} for k, v in ln_detail.items()}},
        "final_loss": losses[-1] if losses else None,
    }
    if world == 1 and not args.no_gpu_eager_baseline:
        from baseline import loader as _ref_loader
        if _ref_loader.available():
            del model, wrapped, optim
            torch.cuda.empty_cache()
            eips, ems = gpu_eager_reference_step_rate(torch, device, B, min(args.steps, 10), 3)
            line["gpu_eager_baseline"] = {
                "value": eips, "unit": UNIT, "ms_per_step": ems, "speedup_of_this_repo": ips / eips,
                "what": "unmodified reference modules (baseline/_ref train_vit.ViTClassifier) on the same B200: PyTorch eager under "
                        "torch.autocast('cuda', bfloat16) + torch.optim.AdamW(fused=True), same batch, device-resident inputs"}
        else:
            line["gpu_eager_baseline"] = {"unavailable": "baseline/_ref absent"}
    if world == 1 and not args.no_cpu_baseline:
        cips, cdt, ccores, ckind, csample = cpu_baseline_sample(args.cpu_batch, 3, 1)
        line["cpu_baseline"] = {"value": cips, "unit": UNIT, "cores": ccores, "kind": ckind, "sample": csample}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (weak scaling); default: the workload's (256 for vit_b)")
    ap.add_argument("--workload", type=str, default="vit_b", choices=["vit_b", "vit_l", "vit_ti", "titok_s", "tatitok_s", "videogpt_b", "vq"],
                    help="vit_b (default) = BASELINE.json configs[1], the headline; the others are the remaining configs through the "
                         "same JSON contract (bench_workloads.py)")
    ap.add_argument("--bucket-mb", type=float, default=0.0,
                    help="gradient bucket size for N > 1; 0 (default) = ONE bucket reduced right after backward.  Measured at 8 GPUs "
                         "(profiles/r2_ddp_8gpu_ab.md): overlapped 32 MB buckets 31.98 ms/step, one deferred bucket 30.99 (fp32) / "
                         "30.73 (bf16): NCCL's CTAs cannot share an SM with the persistent GEMM CTAs (the GEMM owns the whole register "
                         "file), so every overlapped bucket displaces GEMM CTAs of a statically partitioned grid")
    ap.add_argument("--grad-compress", type=str, default="none", choices=["none", "bf16"],
                    help="none (default): fp32 all-reduce of the fp32 gradient bucket; bf16: the bucket crosses NVLink as bf16 (cast - "
                         "all-reduce(sum) - cast back with 1/world into the fp32 bucket), like torch's bf16_compress_hook -- measured "
                         "0.27 ms/step faster at 8 GPUs (30.73 vs 30.99 ms), gradients differ by 2.5e-3 rel-L2; not the default so that the "
                         "headline scaling numbers carry no reduced-precision step")
    ap.add_argument("--impl", type=str, default="b200vit", choices=["b200vit", "reference"])
    ap.add_argument("--cpu-batch", type=int, default=16, help="batch of the bounded CPU-baseline sample (a few seconds per step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.0,
                    help="ViTConfig.dropout (SDPA dropout_p + nn.Dropout after mlp[2]); train_vit.py's own default is 0.15. "
                         "The headline number is measured at 0.0 (BASELINE configs[1]); other values are for throughput only")
    ap.add_argument("--optimizer", type=str, default="b200vit", choices=["b200vit", "torch-fused"],
                    help="AdamW implementation inside the step (default: the fused kernel of this library)")
    args = ap.parse_args()
    args.batch_given = args.batch is not None
    if args.batch is None:
        args.batch = 256
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload != "vit_b":
        import bench_workloads
        return bench_workloads.run(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
