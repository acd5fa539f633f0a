"""A/B: ViT-B/16 224 batch 256 training step, eager launches vs the whole step replayed as one CUDA graph (same box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-is-all-you-need_b200")); sys.path.insert(0, ROOT)
import torch
from b200vit import modules as M
from b200vit.graph import GraphedTrainStep
from b200vit.optim import AdamW
dev = "cuda:0"


def make(capturable):
    torch.manual_seed(0)
    net = M.ViTClassifier(M.ViTConfig(224, 3, 16, "B", 1, 0.0), num_classes=1000).to(dev)
    return net, AdamW(net.parameters(), lr=1e-4, weight_decay=1e-2, capturable=capturable)


x = torch.randn(256, 3, 224, 224, device=dev)
y = torch.randint(0, 1000, (256,), device=dev)
loss_fn = M.CrossEntropyLoss()


def timeit(fn, steps=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


net, opt = make(False)
def eager():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = loss_fn(net(x), y)
    loss.backward(); opt.step()
te = timeit(eager)
del net, opt
torch.cuda.empty_cache()
net2, opt2 = make(True)
step = GraphedTrainStep(net2, opt2, loss_fn, x, y)
tg = timeit(lambda: step(x, y))
te2 = None
print(f"eager {te:.2f} ms/step ({256 / te * 1e3:.0f} img/s)   graphed {tg:.2f} ms/step ({256 / tg * 1e3:.0f} img/s)")
