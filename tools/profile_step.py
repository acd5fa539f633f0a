"""One ViT-B/16 training step bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`.
    python tools/profile_step.py [--batch 256] [--model B]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-is-all-you-need_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from b200vit import modules as M  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--model", type=str, default="B")
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = M.ViTClassifier(M.ViTConfig(224, 3, 16, args.model, 1, 0.0), num_classes=1000).to(dev)
from b200vit.optim import AdamW  # noqa: E402
optim = AdamW(model.parameters(), lr=1e-4, weight_decay=1e-2)
x = torch.randn(args.batch, 3, 224, 224, device=dev)
y = torch.randint(0, 1000, (args.batch,), device=dev)


def step():
    optim.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = M.CrossEntropyLoss()(model(x), y)
    loss.backward()
    optim.step()
    return loss


for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", loss.item())
