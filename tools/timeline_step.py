"""In-step kernel timeline of one ViT training step (warm caches, real launch order) via torch.profiler / CUPTI.

    python tools/timeline_step.py [--batch 256] [--model B] [--steps 3] > profiles/rNN_timeline.md

Prints per-kernel totals per step, the GPU busy time and the idle gaps between kernels.  nsys is not in the
image; CUPTI activity records (what torch.profiler collects) see every kernel of the process, including the
ones launched through the C ABI.
"""
import argparse
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-is-all-you-need_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from b200vit import modules as M  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--model", type=str, default="B")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--dropout", type=float, default=0.0)
ap.add_argument("--bucket-mb", type=float, default=0.0, help="torchrun only: gradient bucket size; 0 = one bucket reduced after backward (bench.py default)")
ap.add_argument("--grad-compress", type=str, default="none", choices=["none", "bf16"])
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
torch.manual_seed(0)
model = M.ViTClassifier(M.ViTConfig(224, 3, 16, args.model, 1, args.dropout), num_classes=1000).to(dev)
net = model
if world > 1:  # torchrun: the same step through the data-parallel wrapper (rank 0 reports)
    import torch.distributed as dist
    from b200vit import ddp
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    net = ddp.DataParallel(model, bucket_mb=args.bucket_mb if args.bucket_mb > 0 else 1e9, compress_bf16=(args.grad_compress == "bf16"))
from b200vit.optim import AdamW  # noqa: E402
optim = AdamW(model.parameters(), lr=1e-4, weight_decay=1e-2)
x = torch.randn(args.batch, 3, 224, 224, device=dev)
y = torch.randint(0, 1000, (args.batch,), device=dev)


def step():
    optim.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = M.CrossEntropyLoss()(net(x), y)
    loss.backward()
    optim.step()
    return loss


for _ in range(4):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    step()
e1.record()
torch.cuda.synchronize()
plain_ms = e0.elapsed_time(e1) / args.steps

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()

if rank != 0:
    dist.destroy_process_group()
    sys.exit(0)
evs = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        evs.append((e.time_range.start, e.time_range.end, e.name))
evs.sort()
t_first, t_last = evs[0][0], max(e[1] for e in evs)
busy = 0.0
cur_s, cur_e = evs[0][0], evs[0][1]
gaps = []
for s, e, n in evs[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, n))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
agg = defaultdict(lambda: [0.0, 0])
for s, e, n in evs:
    short = re.sub(r"\(.*", "", n)
    short = re.sub(r"^void ", "", short)
    agg[short][0] += e - s
    agg[short][1] += 1
k = args.steps
span = t_last - t_first
print(f"# in-step kernel timeline: ViT-{args.model}/16 224 batch {args.batch}, {k} steps under torch.profiler (CUPTI)"
      + (f", rank 0 of {world} GPUs, gradient exchange: " + ("one bucket after backward" if args.bucket_mb <= 0 else f"{args.bucket_mb:g} MB buckets overlapped")
         + f", {args.grad_compress}" if world > 1 else "") + "\n")
if world > 1:
    nccl = [(s_, e_, n_) for s_, e_, n_ in evs if "nccl" in n_.lower()]
    adam = [(s_, e_) for s_, e_, n_ in evs if "adamw_kernel" in n_]
    if nccl and adam:
        per = len(nccl) // k
        print(f"* NCCL kernels per step: {per}; total {sum(e_ - s_ for s_, e_, _ in nccl) / k:.0f} us/step; names: "
              + ", ".join(sorted({re.sub(r'[(<].*', '', n_) for _, _, n_ in nccl})))
        # exposed tail: from the end of the last non-NCCL kernel before each AdamW launch to the AdamW launch
        tails = []
        for a_s, _ in adam:
            prev = max((e_ for s_, e_, n_ in evs if e_ <= a_s and "nccl" not in n_.lower() and "adamw" not in n_), default=a_s)
            tails.append(a_s - prev)
        print(f"* compute idle before AdamW (waiting for the last all-reduce): {sum(tails) / len(tails):.0f} us/step")
print(f"* un-profiled step time (CUDA events): {plain_ms:.2f} ms")
print(f"* profiled span per step: {span / k / 1e3:.2f} ms; GPU busy per step: {busy / k / 1e3:.2f} ms; "
      f"idle (gaps between kernels) per step: {(span - busy) / k / 1e3:.2f} ms over {len(gaps) / k:.0f} gaps")
big = sorted(gaps, reverse=True)[:8]
print("* largest gaps (us, next kernel): " + "; ".join(f"{g:.0f} before {re.sub(r'[(<].*', '', n)[:40]}" for g, n in big))
print("\n| kernel | launches/step | us/step | share of busy | avg us |")
print("|---|---:|---:|---:|---:|")
tot = sum(v[0] for v in agg.values())
for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"| `{name[:100]}` | {n / k:.1f} | {us / k:.1f} | {100 * us / tot:.1f}% | {us / n:.1f} |")
