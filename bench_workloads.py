"""The other BASELINE.json configurations through the same bench contract as bench.py (which dispatches here for
`--workload` other than the default `vit_b`):

    python bench.py --workload vit_l      [--gpus N] [--steps K] [--warmup W] [--batch B]   # configs[3] ViT-L/16 224, batch 256/GPU
    python bench.py --workload titok_s    ...   # configs[2] TiTok-S 256 px, 32 latent tokens, 4096 x 12 (train_titok.py step)
    python bench.py --workload tatitok_s  ...   # config 3'  blocks.py TiTokEncoder + VectorQuantizer + TiTokDecoder (small)
    python bench.py --workload videogpt_b ...   # configs[4] VideoGPT-B, 16 x 64 = 1 024 causal tokens (train_videogpt.py step)
    python bench.py --workload vit_ti     ...   # configs[0] ViT-Ti/4 32 px batch 32 (whole step as one CUDA graph)
    python bench.py --workload vq         ...   # configs[2] the VQ lookup alone: rows = batch * 32, K = 4096, D = 12

Same JSON line: `value` with device-resident inputs, `e2e` from pinned host buffers (H2D of the step's inputs and D2H of its
result inside the timed region), `roofline` (tensor: every tcgen05 GEMM launch timed with CUDA events; vq: algorithmic HBM
bytes + the fp32-FMA fraction that actually bounds it), `cpu_baseline` (the unmodified reference modules on the host cores,
bounded sample) and `gpu_eager_baseline` (the unmodified reference modules on the same B200 under bf16 autocast)."""
import json
import os
import time

import bench as _b

ROOT = os.path.dirname(os.path.abspath(__file__))


def _layer_flops(N, d, causal=False):
    att = 4 * N * N * d
    return 2 * N * d * 3 * d + (att // 2 if causal else att) + 16 * N * d * d


class _TiTokCfgLike:
    """train_titok.TiTokConfig (train_titok.py:18-32) over the drop-in ViTConfig (used when baseline/_ref is absent)."""

    def __init__(self, M, image_size, patch_size, latent_tokens, codebook_size, latent_dim, transformer):
        self.image_size, self.patch_size, self.latent_tokens = image_size, patch_size, latent_tokens
        self.codebook_size, self.latent_dim, self.transformer = codebook_size, latent_dim, transformer
        self.patch_dim = image_size // patch_size
        self.n_patches = self.patch_dim ** 2
        self.enc_vit_config = M.ViTConfig(image_size, 3, patch_size, transformer, latent_tokens, 0.0)
        self.n_embd = self.enc_vit_config.trans_config.n_embd
        self.dec_vit_config = M.ViTConfig(latent_tokens, self.n_embd, 1, transformer, self.n_patches, 0.0)
        self.dec_vit_config.n_patches = latent_tokens


class _VideoGPTCfgLike:
    def __init__(self, M, frame_size, codebook_size, transformer, max_frames, dropout):
        self.frame_size, self.codebook_size, self.transformer = frame_size, codebook_size, transformer
        self.max_frames, self.dropout = max_frames, dropout
        self.max_tokens = max_frames * frame_size
        self.trans_config = M.transformer_configs[transformer](block_size=self.max_tokens, dropout=dropout, causal=True)
        self.n_embd = self.trans_config.n_embd


class _BlocksCfg:
    def __init__(self):
        self.image_size, self.patch_size, self.transformer, self.latent_tokens, self.latent_dim = 256, 16, "small", 32, 12


class _TATiTok:
    """encoder -> VectorQuantizer -> decoder with a learned latent-token parameter, as train_tatitok.TiTok wires blocks.py
    (train_tatitok.py:36-41,62-75); `ns` is either b200vit.modules (drop-ins) or the reference's blocks module."""

    @staticmethod
    def build(torch, enc_cls, dec_cls, vq_cls):
        class Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                cfg = _BlocksCfg()
                self.encoder, self.decoder = enc_cls(cfg), dec_cls(cfg)
                self.latent_tokens = torch.nn.Parameter(512 ** -0.5 * torch.randn(32, 512))
                self.quantize = vq_cls(4096, 12, 0.25, use_l2_norm=True)

            def forward(self, x):
                z = self.encoder(x, self.latent_tokens)
                zq, info = self.quantize(z)
                return self.decoder(zq), info["quantizer_loss"]
        return Net()


def _workload(name, torch, M, ref):
    """Returns a dict describing one workload.  `ref` is the reference namespace (baseline.loader) or None."""
    F = torch.nn.functional
    if name in ("vit_l", "vit_ti"):
        if name == "vit_l":
            size, patch, preset, classes, batch, cpu_batch = 224, 16, "L", 1000, 256, 4
            cfg_tuple = (24, 1024)
            label = "ViT-L/16 224px ImageNet-shape train step (fwd + CE + bwd + AdamW), configs[3]"
        else:
            M.transformer_configs.setdefault("Ti", lambda **kw: M.TransformerConfig(12, 3, 192, **kw))
            size, patch, preset, classes, batch, cpu_batch = 32, 4, "Ti", 10, 32, 32
            cfg_tuple = (12, 192)
            label = "ViT-Ti/4 32px CIFAR-shape train step (fwd + CE + bwd + AdamW), configs[0]"
        N = (size // patch) ** 2 + 1
        L, d = cfg_tuple
        flops = 3 * (L * _layer_flops(N, d) + 2 * d * classes) + 2 * 2 * (N - 1) * (3 * patch * patch) * d

        def inputs(B, g):
            return torch.randn(B, 3, size, size, generator=g), torch.randint(0, classes, (B,), generator=g)

        def make(mods, ce):
            model = mods.ViTClassifier(mods.ViTConfig(size, 3, patch, preset, 1, 0.0), num_classes=classes)
            return model, (lambda m, x, y: ce(m(x), y))
        return dict(metric=f"{name}_train_images_per_sec", unit="images/s", label=label, batch=batch, cpu_batch=cpu_batch,
                    flops_per_unit=flops, units=lambda B: B, inputs=inputs, seq_len=N,
                    ours=lambda: make(M, M.CrossEntropyLoss()),
                    reference=(lambda: make(ref.train_vit, torch.nn.CrossEntropyLoss())) if ref else None,
                    graph=(name == "vit_ti"))
    if name == "titok_s":
        cfg = ref.train_titok.TiTokConfig(256, 16, 32, 4096, 12, "S") if ref else _TiTokCfgLike(M, 256, 16, 32, 4096, 12, "S")
        flops = 3 * 12 * _layer_flops(288, 512) + 2 * 2 * 256 * 768 * 512 + 3 * 2 * 256 * 512 * 768

        def inputs(B, g):
            return (torch.rand(B, 3, 256, 256, generator=g),)

        def loss(m, x):     # train_titok.py:154-158 without the ConvNeXt perceptual term (out of scope)
            recon, _idx, qloss = m(x)
            return (recon.float() - x).pow(2).mean() + qloss
        return dict(metric="titok_s_256_train_images_per_sec", unit="images/s", batch=256, cpu_batch=4, seq_len=288,
                    label="TiTok-S 256px tokenizer train step (enc + VQ 4096x12 + dec, MSE + VQ loss, bwd, AdamW), configs[2]",
                    flops_per_unit=flops, units=lambda B: B, inputs=inputs,
                    ours=lambda: (M.TiTok(cfg), loss), reference=(lambda: (ref.train_titok.TiTok(cfg), loss)) if ref else None)
    if name == "tatitok_s":
        flops = 3 * 16 * (_layer_flops(289, 512) + 2 * 289 * 512 * 512) + 2 * 2 * 256 * 768 * 512 + 3 * 2 * 256 * 512 * 768

        def inputs(B, g):
            return (torch.rand(B, 3, 256, 256, generator=g),)

        def loss(m, x):
            recon, qloss = m(x)
            return (recon.float() - x).pow(2).mean() + qloss
        return dict(metric="tatitok_small_256_train_images_per_sec", unit="images/s", batch=128, cpu_batch=2, seq_len=289,
                    label="blocks.py TiTokEncoder + VectorQuantizer(4096x12, l2) + TiTokDecoder, small, 256px, 32 latent tokens "
                          "(train_tatitok.py wiring; MSE + VQ loss, bwd, AdamW), config 3'",
                    flops_per_unit=flops, units=lambda B: B, inputs=inputs,
                    ours=lambda: (_TATiTok.build(torch, M.BlocksTiTokEncoder, M.BlocksTiTokDecoder, M.VectorQuantizer), loss),
                    reference=(lambda: (_TATiTok.build(torch, ref.blocks.TiTokEncoder, ref.blocks.TiTokDecoder,
                                                       ref.blocks.VectorQuantizer), loss)) if ref else None)
    if name == "videogpt_b":
        cfg = ref.train_videogpt.VideoGPTConfig(64, 1024, "B", 16, 0.0) if ref else _VideoGPTCfgLike(M, 64, 1024, "B", 16, 0.0)
        flops = 3 * (12 * _layer_flops(1024, 768, True) + 2 * 1024 * 768 * 1024)

        def inputs(B, g):
            return (torch.randint(0, 1024, (B, 16, 64), generator=g),)

        def loss(m, tok):
            return m(tok)[1]
        return dict(metric="videogpt_b_1024_train_sequences_per_sec", unit="sequences/s", batch=16, cpu_batch=1, seq_len=1024,
                    label="VideoGPT-B train step (embed + causal stack N=1024 + vocab proj + CE + bwd + AdamW), configs[4]",
                    flops_per_unit=flops, units=lambda B: B, inputs=inputs,
                    ours=lambda: (M.VideoGPT(cfg), loss), reference=(lambda: (ref.train_videogpt.VideoGPT(cfg), loss)) if ref else None)
    raise SystemExit(f"unknown workload {name}")


def _train_step_factory(torch, model, loss_fn, optim):
    def step(*inp, after_forward=None):
        optim.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = loss_fn(model, *inp)
        if after_forward is not None:
            after_forward(loss)
        loss.backward()
        optim.step()
        return loss
    return step


def _time_device(torch, fn, steps, warmup):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def _cpu_reference_sample(torch, w, steps=2, warmup=1):
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model, loss_fn = w["reference"]()
    optim = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-2)
    g = torch.Generator().manual_seed(0)
    inp = w["inputs"](w["cpu_batch"], g)

    def one():
        optim.zero_grad()
        loss = loss_fn(model, *inp)
        loss.backward()
        optim.step()
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    return w["units"](w["cpu_batch"]) / dt, dt, threads


def run(args):
    import torch
    import torch.distributed as dist
    from b200vit import ddp as b200_ddp
    from b200vit import modules as M
    from b200vit import ops
    from b200vit import optim as b200_optim

    from baseline import loader
    if args.workload == "vq":
        return run_vq(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    ref = loader.load(("transformer", "train_vit", "train_titok", "train_videogpt", "blocks")) if loader.available() else None
    w = _workload(args.workload, torch, M, ref)
    B = args.batch if args.batch_given else w["batch"]
    torch.manual_seed(0)
    model, loss_fn = w["ours"]()
    model = model.to(device)
    wrapped = (b200_ddp.DataParallel(model, bucket_mb=args.bucket_mb if args.bucket_mb > 0 else 1e9,
                                     compress_bf16=(args.grad_compress == "bf16")) if world > 1 else model)
    use_graph = bool(w.get("graph")) and world == 1
    optim = b200_optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-2, capturable=use_graph)
    step = _train_step_factory(torch, wrapped, loss_fn, optim)

    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    n_host = 2
    host = [tuple(t.pin_memory() for t in w["inputs"](B, g)) for _ in range(n_host)]
    dev = [tuple(t.to(device) for t in h) for h in host]

    graphed = None
    if use_graph:      # launch-bound configuration: the whole step is captured once (b200vit.graph) and replayed
        static = tuple(t.clone() for t in dev[0])
        for _ in range(3):
            step(*static)
        torch.cuda.synchronize()
        for p in model.parameters():
            if hasattr(p, "_b200_bf16"):
                del p._b200_bf16
        graphed = torch.cuda.CUDAGraph()
        optim.zero_grad(set_to_none=True)
        with torch.cuda.graph(graphed):
            static_loss = step(*static).detach()

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

    def dev_step(i):
        if graphed is not None:
            for s, t in zip(static, dev[i % n_host]):
                s.copy_(t, non_blocking=True)
            graphed.replay()
        else:
            step(*dev[i % n_host])

    for i in range(max(args.warmup, 3)):
        dev_step(i)
    clocks = _b.ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ops.launch_count = 0
    ms = timed(dev_step, args.steps)
    launches = ops.launch_count
    clk = clocks.stop() if rank == 0 else None
    units = w["units"](B)
    rate = world * units * args.steps / (ms / 1e3)

    # end to end: pinned host inputs copied on a side stream, the loss read back every step
    copy_stream = torch.cuda.Stream(device=device)
    state = {}
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
    loss_ev = torch.cuda.Event()
    losses = []

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            t = tuple(h.to(device, non_blocking=True) for h in host[i % n_host])
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        state["next"] = (t, ev)

    def read_back(loss):
        loss_host.copy_(loss.detach(), non_blocking=True)
        loss_ev.record()

    def e2e_step(i):
        if "next" not in state:
            prefetch(i)
        t, ev = state.pop("next")
        torch.cuda.current_stream().wait_event(ev)
        for x in t:
            x.record_stream(torch.cuda.current_stream())
        prefetch(i + 1)
        if graphed is not None:
            for s, x in zip(static, t):
                s.copy_(x, non_blocking=True)
            graphed.replay()
            read_back(static_loss)
        else:
            step(*t, after_forward=read_back)
        loss_ev.synchronize()
        losses.append(float(loss_host))

    for i in range(3):
        e2e_step(i)
    state.clear()
    ms_e2e = timed(e2e_step, args.steps)
    rate_e2e = world * units * args.steps / (ms_e2e / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    peaks = _b.load_peaks()
    peak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
    gemm_ms, gemm_flops, gemm_calls = ops.profile_gemms(lambda: step(*dev[0]), steps=2)
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    step_tflops = rate / world * w["flops_per_unit"] / 1e12
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return 0
    line = {
        "metric": w["metric"], "value": rate, "unit": w["unit"], "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": w["label"], "per_gpu_batch": B, "global_batch": B * world, "seq_len": w["seq_len"],
                   "parallelism": f"dp{world}", "optimizer": "b200vit.optim.AdamW", "cuda_graph": graphed is not None,
                   "l2": "per-step working set >> 126 MB L2 (no flush needed)" if not use_graph else
                         "launch-bound configuration: working set fits L2 by construction (32 images of 32x32)",
                   "train_gflop_per_unit": w["flops_per_unit"] / 1e9},
        "e2e": {"value": rate_e2e, "unit": w["unit"], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches if graphed is None else f"{gemm_calls // 2} GEMM launches + the rest of the step, replayed as one CUDA graph",
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all Linear fwd/dgrad/wgrad launches of a step)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})", "traffic": None,
                     "launches_timed": gemm_calls, "step_tflops_per_gpu": step_tflops,
                     "step_frac_of_peak": step_tflops / peak if peak else None, "step_frac_of_nominal_2250": step_tflops / 2250.0},
        "final_loss": losses[-1] if losses else None,
    }
    if world == 1 and w["reference"] is not None and not args.no_gpu_eager_baseline:
        del model, wrapped, optim, step
        torch.cuda.empty_cache()
        torch.manual_seed(0)
        rmodel, rloss = w["reference"]()
        rmodel = rmodel.to(device)
        roptim = torch.optim.AdamW(rmodel.parameters(), lr=1e-4, weight_decay=1e-2, fused=True)
        rstep = _train_step_factory(torch, rmodel, rloss, roptim)
        try:
            ems = _time_device(torch, lambda i: rstep(*dev[i % n_host]), min(args.steps, 10), 3) / min(args.steps, 10)
            line["gpu_eager_baseline"] = {"value": units / (ems / 1e3), "unit": w["unit"], "ms_per_step": ems,
                                          "speedup_of_this_repo": rate / (units / (ems / 1e3)),
                                          "what": "unmodified reference modules on the same B200: PyTorch eager under bf16 autocast + "
                                                  "torch.optim.AdamW(fused=True), same batch, device-resident inputs"}
        except Exception as e:     # e.g. the eager path running out of memory at this batch
            line["gpu_eager_baseline"] = {"unavailable": f"{type(e).__name__}: {str(e)[:120]}"}
        del rmodel, roptim, rstep
        torch.cuda.empty_cache()
    if world == 1 and w["reference"] is not None and not args.no_cpu_baseline:
        cv, cdt, threads = _cpu_reference_sample(torch, w)
        line["cpu_baseline"] = {"value": cv, "unit": w["unit"], "cores": threads, "kind": "reference",
                                "sample": f"2 steps of batch {w['cpu_batch']} after 1 warm-up, unmodified reference modules on CPU fp32 ({cdt:.1f} s/step)"}
    print(json.dumps(line), flush=True)
    return 0


def run_vq(args):
    """configs[2]'s VQ lookup alone through the module API (b200vit.modules.Quantizer.forward): rows = batch * 32 latents
    of dim 12 against a 4096 x 12 codebook.  Inputs rotate over more buffers than fit L2."""
    import numpy as np
    import torch
    from b200vit import modules as M
    from b200vit import ops

    from baseline import loader
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    if int(os.environ.get("RANK", "0")) != 0:
        return 0     # replicas only: rank 0 reports one replica
    B = args.batch if args.batch_given else 256
    R, K, D = B * 32, 4096, 12

    class Cfg:
        codebook_size, latent_dim = K, D
    torch.manual_seed(0)
    q = M.Quantizer(Cfg()).to(device)
    q.codebook.weight.data.normal_()        # "trained-like" codebook
    n_buf = max(4, int(160e6 // (R * D * 4)) + 1)       # > 126 MB of distinct inputs
    g = torch.Generator().manual_seed(1)
    xs = [torch.randn(B, 32, D, generator=g).to(device) for _ in range(min(n_buf, 512))]
    host = [torch.randn(B, 32, D, generator=g).pin_memory() for _ in range(2)]

    def dev_step(i):
        with torch.no_grad():
            return q(xs[i % len(xs)])
    # One lookup is ~30 us of GPU work behind ~70 us of Python (allocation of the outputs, ctypes marshalling): the device-resident
    # arm replays a CUDA graph of one call per input buffer, so `value` is the kernels' rate; `e2e` below is the eager public call
    for i in range(3):
        dev_step(i)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        keep = [dev_step(i) for i in range(len(xs))]
    steps = len(xs) * max(1, (max(args.steps, 50) + len(xs) - 1) // len(xs))
    clocks = _b.ClockSampler(device.index)
    clocks.start()
    ops.launch_count = 0
    ms = _time_device(torch, lambda i: graph.replay(), steps // len(xs), max(args.warmup, 3))
    launches = 3 * steps
    clk = clocks.stop()
    rate = R * steps / (ms / 1e3)
    idx_host = torch.empty(B, 32, dtype=torch.int64).pin_memory()

    def e2e_step(i):
        x = host[i % 2].to(device, non_blocking=True)
        with torch.no_grad():
            _, idx, _ = q(x)
        idx_host.copy_(idx, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e2e_steps = max(args.steps, 50)
    ms_e2e = _time_device(torch, e2e_step, e2e_steps, 3)
    peaks = _b.load_peaks()
    us = ms / steps * 1e3
    alg_bytes = R * 104 + K * D * 4
    fma_tflops = 2.0 * R * K * D / (us * 1e-6) / 1e12
    fma_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    line = {"metric": "vq_lookup_rows_per_sec", "value": rate, "unit": "rows/s", "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "TiTok VQ nearest-codebook lookup (normalise, distances, argmin, gather, losses, straight-through), "
                                   "configs[2] shape", "rows": R, "codebook": [K, D],
                       "l2": f"inputs rotate over {len(xs)} buffers ({len(xs) * R * D * 4 / 1e6:.0f} MB > 126 MB L2)",
                       "cuda_graph": f"value: one graph of {len(xs)} lookups replayed (kernel rate); e2e: eager calls"},
            "e2e": {"value": R * e2e_steps / (ms_e2e / 1e3), "unit": "rows/s", "h2d_bytes_per_step": R * D * 4, "d2h_bytes_per_step": R * 8,
                    "ms_per_step": ms_e2e / e2e_steps, "note": "eager public call per step (Python + allocation + 3 launches), host-bound"},
            "gpu_launches": launches, "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "vq_fwd_kernel (+ codebook normalisation and loss-finalise launches)",
                         "achieved": alg_bytes / (us * 1e-6) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": alg_bytes / (us * 1e-6) / 1e9 / peaks["hbm_gbs"], "traffic": None,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "104 B/row of HBM traffic against 98 304 fp32 FLOP/row: the kernel is bounded by the fp32 FMA pipe, "
                                 "not by HBM (distances never leave the SM; bit-exact fp32 contract rules out tensor cores)",
                         "fp32_fma": {"achieved_tflops": fma_tflops, "peak_tflops_nominal": fma_peak, "frac": fma_tflops / fma_peak}}}
    if loader.available():
        ref = loader.load(("transformer", "train_vit", "train_titok"))
        rq = ref.train_titok.Quantizer(Cfg()).to(device)
        rq.load_state_dict(q.state_dict())
        with torch.no_grad():
            a = rq(xs[0])
            b = q(xs[0])
        line["indices_bit_exact_vs_reference_on_gpu"] = bool(torch.equal(a[1], b[1]))
        ems = _time_device(torch, lambda i: rq(xs[i % len(xs)]), steps, 3) / steps

        line["gpu_eager_baseline"] = {"value": R / (ems / 1e3), "unit": "rows/s", "ms_per_step": ems, "speedup_of_this_repo": rate / (R / (ems / 1e3)),
                                      "what": "train_titok.Quantizer.forward (F.normalize x2, torch.cdist, argmin, gather, losses) on the same B200"}
        if not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            cq = ref.train_titok.Quantizer(Cfg())
            cq.load_state_dict({k: v.cpu() for k, v in q.state_dict().items()})
            xc = xs[0].cpu()
            with torch.no_grad():
                cq(xc)
                t0 = time.perf_counter()
                for _ in range(5):
                    out = cq(xc)
                dt = (time.perf_counter() - t0) / 5
            line["cpu_baseline"] = {"value": R / dt, "unit": "rows/s", "cores": torch.get_num_threads(), "kind": "reference",
                                    "sample": f"5 calls of train_titok.Quantizer.forward on {R} rows (CPU fp32)"}
            line["indices_bit_exact_vs_reference_on_cpu"] = bool(np.array_equal(out[1].numpy(), b[1].cpu().numpy()))
    print(json.dumps(line), flush=True)
    return 0
