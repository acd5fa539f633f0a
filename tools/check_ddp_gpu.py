"""torchrun --nproc-per-node N tools/check_ddp_gpu.py: gradients of the data-parallel wrapper (bucketed all-reduce, wgrad
kernels writing straight into the buckets) on a sharded batch == gradients of the plain model on the whole batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-is-all-you-need_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from b200vit import ddp  # noqa: E402
from b200vit import modules as M  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
cfg = M.ViTConfig(64, 3, 8, "S", 1, 0.0)
ref = M.ViTClassifier(cfg, num_classes=10).to(dev)
par = M.ViTClassifier(cfg, num_classes=10).to(dev)
par.load_state_dict(ref.state_dict())
wrapped = ddp.DataParallel(par, bucket_mb=4.0)
g = torch.Generator(device="cpu").manual_seed(1)
B = 8 * world
x = torch.randn(B, 3, 64, 64, generator=g).to(dev)
y = torch.randint(0, 10, (B,), generator=g).to(dev)
loss_fn = torch.nn.CrossEntropyLoss()
for it in range(2):
    ref.zero_grad(set_to_none=True)
    par.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss_fn(ref(x).float(), y).backward()
        sl = slice(rank * 8, rank * 8 + 8)
        loss_fn(wrapped(x[sl]).float(), y[sl]).backward()
    torch.cuda.synchronize()
    worst = 0.0
    for (n, a), b in zip(ref.named_parameters(), par.parameters()):
        assert b.grad is not None, n
        err = ((a.grad - b.grad).norm() / (a.grad.norm() + 1e-12)).item()
        worst = max(worst, err)
        assert err < 2e-2, (n, err)
    if rank == 0:
        print(f"iteration {it}: max rel-L2 gradient difference over {len(list(ref.parameters()))} parameters = {worst:.3e}")
dist.destroy_process_group()
if rank == 0:
    print("DDP gradient check OK")
