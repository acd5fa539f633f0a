"""GEMM micro-benchmark at the ViT-B/16 B=256 shapes: 1-CTA tiles vs CTA pairs (debug knob 5), CUDA events.
    python tools/bench_gemm.py [--workers N]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-is-all-you-need_b200"))
import torch  # noqa: E402

from b200vit import _cabi, ops  # noqa: E402

dev = "cuda:0"
lib = _cabi.ensure_device(0)
print("max co-resident CTA pairs:", lib.b200vit_debug_max_clusters())
M = int(os.environ.get("M", 50432))
iters = int(os.environ.get("ITERS", 10))


def timeit(fn, iters=iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def rnd(*s):
    return (torch.randn(*s, device=dev) * 0.1).to(torch.bfloat16)


shapes = {"qkv": (768, 2304), "fc1": (768, 3072), "fc2": (3072, 768)}
modes = [int(a) for a in os.environ.get("MODES", "1,2").split(",")]
for name, (K, N) in shapes.items():
    x, w, dy = rnd(M, K), rnd(N, K), rnd(M, N)
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev)
    gp = rnd(M, K)
    fl = 2.0 * M * N * K
    for mode in modes:
        lib.b200vit_debug_set(5, mode)
        if "--workers" in sys.argv:
            lib.b200vit_debug_set(3, int(sys.argv[sys.argv.index("--workers") + 1]))
        rows = []
        rows.append(("fwd bias", timeit(lambda: ops.gemm_bias(x, w, bias))))
        rows.append(("fwd gelu", timeit(lambda: ops.gemm_bias_gelu(x, w, bias))))
        rows.append(("fwd resid", timeit(lambda: ops.gemm_bias_residual(x, w, bias, res))))
        rows.append(("dgrad", timeit(lambda: ops.gemm_dgrad(dy, w))))
        rows.append(("dgrad*gp", timeit(lambda: ops.gemm_dgrad_dgelu(dy, w, gp))))
        rows.append(("wgrad", timeit(lambda: ops.gemm_wgrad(dy, x))))
        rows.append(("wgrad+db", timeit(lambda: ops.gemm_wgrad(dy, x, want_bias=True))))
        print(f"{name} [M={M} N={N} K={K}] ncta={mode}: " + "  ".join(f"{n} {t:6.1f}us {fl / t / 1e6:6.0f}TF" for n, t in rows), flush=True)
    t = timeit(lambda: torch.matmul(x, w.t()))
    print(f"{name} cuBLAS fwd: {t:6.1f}us {fl / t / 1e6:6.0f}TF", flush=True)
lib.b200vit_debug_set(5, 0)
