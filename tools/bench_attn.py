"""Attention / LayerNorm micro-benchmarks at the ViT-B/16 B=256 shapes (CUDA events)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-is-all-you-need_b200"))
import torch  # noqa: E402

from b200vit import ops  # noqa: E402

dev = "cuda:0"
B, N, H = int(os.environ.get("B", 256)), int(os.environ.get("N", 197)), int(os.environ.get("H", 12))
d = H * 64
iters = int(os.environ.get("ITERS", 5))
causal = os.environ.get("CAUSAL", "0") == "1"


def timeit(fn, iters=iters, warm=2):
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


qkv = torch.randn(B, N, 3 * d, device=dev).to(torch.bfloat16)
do = torch.randn(B, N, d, device=dev).to(torch.bfloat16)
o, lse = ops.flash_attn_fwd(qkv, B, N, H, causal)
fl = 4.0 * N * N * d * B * (0.5 if causal else 1.0)
t = timeit(lambda: ops.flash_attn_fwd(qkv, B, N, H, causal))
print(f"attn fwd  B={B} N={N} H={H} causal={causal}: {t:8.1f} us  {fl/t/1e6:7.1f} TFLOP/s (algorithmic 4*N^2*d*B)")
t = timeit(lambda: ops.flash_attn_bwd(qkv, o, do, lse, B, N, H, causal))
print(f"attn bwd  B={B} N={N} H={H}: {t:8.1f} us  {2.5*fl/t/1e6:7.1f} TFLOP/s (2.5x fwd)")
if "--attn-only" in sys.argv:
    sys.exit(0)
if "--sdpa" in sys.argv:
    # the same-box library comparator: F.scaled_dot_product_attention exactly as transformer.py:27-28 calls it (strided q/k/v views
    # of the fused projection, additive -inf mask when causal), forward and backward
    q, k, v = (qkv.view(B, N, 3, H, 64)[:, :, i].transpose(1, 2).detach().requires_grad_(True) for i in range(3))
    mask = None
    if causal:
        mask = torch.triu(torch.ones(N, N, device=dev), diagonal=1)
        mask = mask.masked_fill(mask == 1, float("-inf")).to(torch.bfloat16)
    sd = lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask)  # noqa: E731
    t = timeit(sd)
    print(f"torch SDPA fwd: {t:8.1f} us  {fl/t/1e6:7.1f} TFLOP/s")
    out = sd()
    go = do.view(B, N, H, 64).transpose(1, 2)
    tb = timeit(lambda: torch.autograd.grad(out, (q, k, v), go, retain_graph=True))
    print(f"torch SDPA bwd: {tb:8.1f} us  {2.5*fl/tb/1e6:7.1f} TFLOP/s")
    sys.exit(0)
M = B * N
x = torch.randn(M, d, device=dev)
add = torch.randn(M, d, device=dev).to(torch.bfloat16)
dy = torch.randn(M, d, device=dev).to(torch.bfloat16)
y, _, mean, rstd, _ = ops.layernorm_fwd(x)
t = timeit(lambda: ops.layernorm_fwd(x))
print(f"ln fwd           : {t:8.1f} us  {M*d*6/t/1e3:7.1f} GB/s (6 B/elt)")
t = timeit(lambda: ops.layernorm_fwd(x, add=add, want_x_out=True))
print(f"ln fwd + add     : {t:8.1f} us  {M*d*12/t/1e3:7.1f} GB/s (12 B/elt)")
t = timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, dres=x, want_bf16=True))
print(f"ln bwd           : {t:8.1f} us  {M*d*16/t/1e3:7.1f} GB/s (16 B/elt)")
for Ncol in (768, 2304, 3072):
    a = torch.randn(M, Ncol, device=dev).to(torch.bfloat16)
    t = timeit(lambda: ops.colsum_bf16(a))
    print(f"colsum [{M},{Ncol}] : {t:8.1f} us  {M*Ncol*2/t/1e3:7.1f} GB/s")
