"""GPU bring-up probe for the tcgen05 GEMM family: correctness vs torch fp32 matmul on the same bf16
inputs, descriptor sweeps for the MN-major operand modes, and CUDA-event timings at ViT-B shapes.

Run on a B200:  python tools/probe_gemm.py [--quick]
"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-is-all-you-need_b200"))

import torch  # noqa: E402

from b200vit import _cabi  # noqa: E402

c_int = ctypes.c_int
dev = torch.device("cuda:0")
lib = _cabi.ensure_device(0) if "--phase" in sys.argv else None


def sp():
    return _cabi.stream_ptr()


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev, dtype=torch.float32) * scale).to(torch.bfloat16)


def report(name, got, ref, tol):
    got = got.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-6
    maxerr = err.max().item()
    rel = maxerr / denom
    nbad = (err > tol * denom).sum().item()
    ok = nbad == 0 and bool(torch.isfinite(got).all())
    print(f"  [{'OK ' if ok else 'BAD'}] {name}: max_abs_err={maxerr:.4e} rel={rel:.3e} bad={nbad}/{got.numel()}",
          flush=True)
    if not ok:
        bad = (err > tol * denom)
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        print(f"        bad rows: n={rows.numel()} first={rows[:8].tolist()} last={rows[-4:].tolist()}")
        print(f"        bad cols: n={cols.numel()} first={cols[:8].tolist()} last={cols[-4:].tolist()}")
        r0 = rows[0].item()
        c0 = cols[0].item()
        print(f"        got[{r0},{c0}:{c0+4}]={got[r0, c0:c0+4].tolist()} ref={ref[r0, c0:c0+4].tolist()}")
    return ok


def run_fwd(M, N, K):
    x, w = rnd(M, K), rnd(N, K, scale=0.05)
    bias = torch.randn(N, device=dev)
    ref = x.float() @ w.float().t() + bias
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    _cabi.check(lib.b200vit_gemm_bias(_cabi.ptr(x), _cabi.ptr(w), _cabi.ptr(bias), _cabi.ptr(y), c_int(M), c_int(N), c_int(K), sp()))
    torch.cuda.synchronize()
    ok = report(f"gemm_bias      M={M} N={N} K={K}", y, ref, 1e-2)
    # f32 + residual epilogues
    res = torch.randn(M, N, device=dev)
    out = torch.empty(M, N, device=dev)
    _cabi.check(lib.b200vit_gemm_bias_residual(_cabi.ptr(x), _cabi.ptr(w), _cabi.ptr(bias), _cabi.ptr(res), _cabi.ptr(out), c_int(M), c_int(N), c_int(K), sp()))
    torch.cuda.synchronize()
    ok &= report(f"gemm_bias_resid M={M} N={N} K={K}", out, ref + res, 2e-3)
    g = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    u = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    _cabi.check(lib.b200vit_gemm_bias_gelu(_cabi.ptr(x), _cabi.ptr(w), _cabi.ptr(bias), _cabi.ptr(g), _cabi.ptr(u), c_int(M), c_int(N), c_int(K), sp()))
    torch.cuda.synchronize()
    r64 = ref.double().requires_grad_(True)
    gref = torch.nn.functional.gelu(r64)
    gpref, = torch.autograd.grad(gref.sum(), r64)
    ok &= report(f"gemm_bias_gelu.gp M={M} N={N} K={K}", u, gpref.float(), 1e-2)
    ok &= report(f"gemm_bias_gelu.g M={M} N={N} K={K}", g, gref.detach().float(), 1e-2)
    return ok


def run_dgrad(M, N, K):
    dy, w = rnd(M, N), rnd(N, K, scale=0.05)
    ref = dy.float() @ w.float()
    dx = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
    _cabi.check(lib.b200vit_gemm_dgrad(_cabi.ptr(dy), _cabi.ptr(w), _cabi.ptr(dx), c_int(M), c_int(N), c_int(K), sp()))
    torch.cuda.synchronize()
    return report(f"gemm_dgrad     M={M} N={N} K={K}", dx, ref, 1e-2)


def run_wgrad(M, N, K):
    dy, x = rnd(M, N, scale=0.1), rnd(M, K, scale=0.1)
    ref = dy.float().t() @ x.float()
    dw = torch.empty(N, K, device=dev)
    _cabi.check(lib.b200vit_gemm_wgrad(_cabi.ptr(dy), _cabi.ptr(x), _cabi.ptr(dw), c_int(M), c_int(N), c_int(K), c_int(0), sp()))
    torch.cuda.synchronize()
    return report(f"gemm_wgrad     M={M} N={N} K={K}", dw, ref, 2e-3)


def guarded(fn, *a):
    try:
        return fn(*a)
    except Exception as e:  # noqa: BLE001
        print(f"  [EXC] {fn.__name__}{a}: {e}", flush=True)
        return False


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


FWD_SHAPES = [(128, 256, 64), (128, 256, 128), (256, 256, 768), (384, 768, 768), (1000, 2304, 768),
              (197 * 4, 192, 192), (130, 128, 48), (2080, 576, 192)]
DGRAD_SHAPES = [(128, 64, 256), (128, 128, 256), (256, 768, 768), (1000, 2304, 768), (788, 3072, 768), (520, 576, 192)]
WGRAD_SHAPES = [(64, 128, 256), (128, 128, 256), (1024, 768, 768), (197 * 8, 2304, 768), (1000, 768, 3072), (520, 576, 192)]


def phase_fwd():
    ok = True
    for shp in FWD_SHAPES:
        ok &= guarded(run_fwd, *shp)
    return ok


def phase_dgrad():
    ok = True
    for shp in DGRAD_SHAPES:
        ok &= guarded(run_dgrad, *shp)
    return ok


def phase_wgrad():
    ok = True
    for shp in WGRAD_SHAPES:
        ok &= guarded(run_wgrad, *shp)
    return ok


def phase_sweep():
    found = False
    for lbo in (8192, 1024, 128, 16, 4096, 2048):
        for sbo in (1024, 8192, 128, 2048):
            for kadv in (2048, 32, 256, 1024, 4096):
                lib.b200vit_debug_set(0, lbo); lib.b200vit_debug_set(1, sbo); lib.b200vit_debug_set(2, kadv)
                print(f" lbo={lbo} sbo={sbo} kadv={kadv}")
                a = guarded(run_dgrad, 256, 256, 256)
                b = guarded(run_wgrad, 256, 128, 256)
                if a and b:
                    found = True
                    print(f" >>> WORKING MN descriptor: lbo={lbo} sbo={sbo} kadv={kadv}")
    return found


def phase_time():
    M = 50432
    for name, N, K in [("qkv", 2304, 768), ("fc1", 3072, 768), ("fc2", 768, 3072)]:
        x, w = rnd(M, K), rnd(N, K, scale=0.05)
        bias = torch.randn(N, device=dev)
        y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        dyy = rnd(M, N)
        dx = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
        dw = torch.empty(N, K, device=dev)
        fl = 2.0 * M * N * K
        for bn in (256, 128):
            lib.b200vit_debug_set(4, bn)
            t = timeit(lambda: lib.b200vit_gemm_bias(_cabi.ptr(x), _cabi.ptr(w), _cabi.ptr(bias), _cabi.ptr(y), c_int(M), c_int(N), c_int(K), sp()))
            print(f"  {name} fwd   BN={bn}: {t*1e3:8.1f} us  {fl/t/1e9:8.1f} TFLOP/s", flush=True)
            t = timeit(lambda: lib.b200vit_gemm_dgrad(_cabi.ptr(dyy), _cabi.ptr(w), _cabi.ptr(dx), c_int(M), c_int(N), c_int(K), sp()))
            print(f"  {name} dgrad BN={bn}: {t*1e3:8.1f} us  {fl/t/1e9:8.1f} TFLOP/s", flush=True)
            t = timeit(lambda: lib.b200vit_gemm_wgrad(_cabi.ptr(dyy), _cabi.ptr(x), _cabi.ptr(dw), c_int(M), c_int(N), c_int(K), c_int(0), sp()))
            print(f"  {name} wgrad BN={bn}: {t*1e3:8.1f} us  {fl/t/1e9:8.1f} TFLOP/s", flush=True)
        lib.b200vit_debug_set(4, 0)
        t = timeit(lambda: torch.matmul(x, w.t()))
        print(f"  {name} torch.matmul (cuBLAS) fwd: {t*1e3:8.1f} us  {fl/t/1e9:8.1f} TFLOP/s", flush=True)
    return True


PHASES = {"fwd": phase_fwd, "dgrad": phase_dgrad, "wgrad": phase_wgrad, "sweep": phase_sweep, "time": phase_time}


def main():
    import subprocess
    if "--phase" in sys.argv:
        name = sys.argv[sys.argv.index("--phase") + 1]
        print(f"== phase {name} on {torch.cuda.get_device_name(0)} ==", flush=True)
        ok = PHASES[name]()
        print(f"PHASE {name}: {'OK' if ok else 'FAILED'}", flush=True)
        return 0 if ok else 1
    # each phase in its own process: a trapped kernel poisons the CUDA context of that process only
    results = {}
    for name in ("fwd", "dgrad", "wgrad"):
        r = subprocess.run([sys.executable, __file__, "--phase", name], timeout=600)
        results[name] = r.returncode
    if results["fwd"] == 0 and (results["dgrad"] != 0 or results["wgrad"] != 0):
        r = subprocess.run([sys.executable, __file__, "--phase", "sweep"], timeout=900)
        results["sweep"] = r.returncode
    if results["fwd"] == 0 and "--quick" not in sys.argv:
        r = subprocess.run([sys.executable, __file__, "--phase", "time"], timeout=600)
        results["time"] = r.returncode
    print("RESULTS", results, flush=True)
    return 0 if all(v == 0 for v in results.values()) else 1


if __name__ == "__main__":
    sys.exit(main())
