"""BASELINE configs[0] (ViT-Ti, 32 px, batch 32) is launch-bound: eager vs CUDA-graph replay of the whole training step
(b200vit.graph.GraphedTrainStep), same trajectory check + timing.  lr 1e-4: at 1e-3 this model's loss trajectory is
chaotic, so eager and graphed runs (atomics order differs) cannot be compared step by step."""
import sys, time
sys.path.insert(0, "/root/repo/vit-is-all-you-need_b200")
import torch
from b200vit import modules as M
from b200vit.graph import GraphedTrainStep
from b200vit.optim import AdamW
dev = "cuda:0"
M.transformer_configs.setdefault("Ti", lambda **kw: M.TransformerConfig(12, 3, 192, **kw))
def make(capturable):
    torch.manual_seed(0)
    net = M.ViTClassifier(M.ViTConfig(32, 3, 4, "Ti", 1, 0.0), num_classes=10).to(dev)
    opt = AdamW(net.parameters(), lr=1e-4, capturable=capturable)
    return net, opt
g = torch.Generator().manual_seed(1)
xs = [torch.randn(32, 3, 32, 32, generator=g).to(dev) for _ in range(8)]
ys = [torch.randint(0, 10, (32,), generator=g).to(dev) for _ in range(8)]
loss_fn = torch.nn.CrossEntropyLoss()
net, opt = make(False)
def estep(x, y):
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = loss_fn(net(x).float(), y)
    loss.backward(); opt.step(); return loss.detach()
for i in range(3): estep(xs[0], ys[0])   # same warm-up as the graphed run
eager = [float(estep(xs[i], ys[i])) for i in range(8)]
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(50): estep(xs[i % 8], ys[i % 8])
torch.cuda.synchronize(); te = (time.perf_counter() - t0) / 50
net2, opt2 = make(True)
step = GraphedTrainStep(net2, opt2, loss_fn, xs[0], ys[0])
graphed = [float(step(xs[i], ys[i])) for i in range(8)]
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(200): step(xs[i % 8], ys[i % 8])
torch.cuda.synchronize(); tg = (time.perf_counter() - t0) / 200
print("eager  losses", [round(v, 4) for v in eager])
print("graph  losses", [round(v, 4) for v in graphed])
print(f"eager {te*1e3:.3f} ms/step ({32/te:.0f} img/s)   graphed {tg*1e3:.3f} ms/step ({32/tg:.0f} img/s)   speed-up {te/tg:.1f}x")
