"""fc1 + GELU forward and its backward twin at the ViT-B/16 B=256 shapes: GELU' carried as bf16 vs as the 8-bit code.
    python tools/bench_gelu_q8.py      (CUDA events, buffers rotated so that nothing stays in L2)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-is-all-you-need_b200"))
import torch  # noqa: E402

from b200vit import ops  # noqa: E402

dev = "cuda:0"
M = int(os.environ.get("M", 50432))


def timeit(fn, iters=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for d in (768, 1024, 512):
    K, N = d, 4 * d
    xs = [(torch.randn(M, K, device=dev)).to(torch.bfloat16) for _ in range(3)]
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    w2 = (torch.randn(K, N, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    dys = [(torch.randn(M, K, device=dev) * 0.1).to(torch.bfloat16) for _ in range(3)]
    fl = 2.0 * M * N * K
    rows = []
    for q8 in (False, True):
        t_f = timeit(lambda i: ops.gemm_bias_gelu(xs[i % 3], w, b, q8=q8))
        gps = [ops.gemm_bias_gelu(x, w, b, q8=q8)[1] for x in xs]
        t_b = timeit(lambda i: ops.gemm_dgrad_dgelu(dys[i % 3], w2, gps[i % 3]))
        rows.append((q8, t_f, t_b))
    t_plain = timeit(lambda i: ops.gemm_bias(xs[i % 3], w, b))
    t_dg = timeit(lambda i: ops.gemm_dgrad(dys[i % 3], w2))
    print(f"d={d} M={M}: plain fwd {t_plain:.1f} us ({fl / t_plain / 1e6:.0f} TF), plain dgrad {t_dg:.1f} us | " +
          " | ".join(f"{'q8  ' if q else 'bf16'}: fwd+gelu {tf:.1f} us ({fl / tf / 1e6:.0f} TF), dgrad*gelu' {tb:.1f} us ({fl / tb / 1e6:.0f} TF)" for q, tf, tb in rows), flush=True)
    del xs, dys, gps
    torch.cuda.empty_cache()
