"""Decode-path micro-benchmarks: skinny (M = batch) GEMMs and VideoGPT.generate.  B200VIT_DEBUG="6=1" forces the tile kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-is-all-you-need_b200")); sys.path.insert(0, ROOT)
import torch
from b200vit import modules as M, ops
dev = "cuda:0"


def timeit(fn, iters=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


for (m, n, k) in ((16, 2304, 768), (16, 3072, 768), (16, 768, 3072), (16, 1024, 768), (64, 2304, 768)):
    x = torch.randn(m, k, device=dev).bfloat16(); w = torch.randn(n, k, device=dev).bfloat16(); b = torch.randn(n, device=dev)
    r = torch.randn(m, n, device=dev)
    us = timeit(lambda: ops.gemm_bias(x, w, b))
    us2 = timeit(lambda: ops.gemm_bias_residual(x, w, b, r))
    print(f"gemm_bias M={m} N={n} K={k}: {us:.1f} us ({n * k * 2 / us / 1e3:.0f} GB/s of weights)   +residual fp32: {us2:.1f} us")


class Cfg:
    def __init__(self):
        self.frame_size, self.codebook_size, self.transformer, self.max_frames, self.dropout = 64, 1024, "B", 16, 0.0
        self.max_tokens = 1024
        self.trans_config = M.transformer_configs["B"](block_size=1024, dropout=0.0, causal=True)
        self.n_embd = 768


torch.manual_seed(0)
net = M.VideoGPT(Cfg()).to(dev).eval()
prompt = torch.randint(0, 1024, (16, 512), device=dev)
for graph_min in (8, 10 ** 9):
    M.VideoGPT.GRAPH_MIN_STEPS = graph_min
    net.generate(prompt, 12)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    net.generate(prompt, 64)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    net.generate(prompt, 1)
    torch.cuda.synchronize(); tp = time.perf_counter() - t0
    print(f"generate 512 + 64, batch 16, {'graph' if graph_min == 8 else 'eager'}: {t * 1e3:.1f} ms total, prefill + first token {tp * 1e3:.1f} ms, "
          f"{(t - tp) / 63 * 1e3:.3f} ms per further token")

if "--profile" in sys.argv:
    from torch.profiler import ProfilerActivity, profile
    M.VideoGPT.GRAPH_MIN_STEPS = 10 ** 9
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net.generate(prompt, 17)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
