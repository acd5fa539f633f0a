import sys
sys.path.insert(0, "/root/repo/vit-is-all-you-need_b200")
import torch
from b200vit import ops
dev = "cuda:0"
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for R, K in ((256 * 32, 4096), (256 * 256, 2048), (16 * 16 * 64, 1024), (32 * 32, 4096)):
    x = torch.randn(R, 12, device=dev); cb = torch.randn(K, 12, device=dev)
    t = timeit(lambda: ops.vq_fwd(x, cb))
    print(f"VQ fwd rows={R} K={K}: {t:.1f} us  {2.0*R*K*12/t/1e6:.2f} TFLOP/s")
