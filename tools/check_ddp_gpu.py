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


# ---- the other drop-in model families through the same wrapper (their extra parameters take the hook path) ----------------
class _TiTokCfg:   # train_titok.TiTokConfig (train_titok.py:18-32)
    def __init__(self, image_size, patch_size, latent_tokens, codebook_size, latent_dim, transformer):
        self.image_size, self.patch_size, self.latent_tokens = image_size, patch_size, latent_tokens
        self.codebook_size, self.latent_dim, self.transformer = codebook_size, latent_dim, transformer
        self.patch_dim = image_size // patch_size
        self.n_patches = self.patch_dim ** 2
        self.enc_vit_config = M.ViTConfig(image_size, 3, patch_size, transformer, latent_tokens, 0.0)
        self.n_embd = self.enc_vit_config.trans_config.n_embd
        self.dec_vit_config = M.ViTConfig(latent_tokens, self.n_embd, 1, transformer, self.n_patches, 0.0)
        self.dec_vit_config.n_patches = latent_tokens


class _GPTCfg:     # train_videogpt.VideoGPTConfig (train_videogpt.py:18-28)
    def __init__(self):
        self.frame_size, self.codebook_size, self.transformer, self.max_frames, self.dropout = 16, 64, "XS2", 4, 0.0
        self.max_tokens = 64
        self.trans_config = M.transformer_configs["XS2"](block_size=64, dropout=0.0, causal=True)
        self.n_embd = self.trans_config.n_embd


M.transformer_configs["XS2"] = lambda **kw: M.TransformerConfig(n_layers=2, n_heads=2, n_embd=128, **kw)


def check(name, make, batch, loss_of, tol=3e-2):
    torch.manual_seed(0)
    ref, par = make().to(dev), make().to(dev)
    par.load_state_dict(ref.state_dict())
    wrapped = ddp.DataParallel(par, bucket_mb=1.0)
    xs = batch()
    n = xs.shape[0] // world
    ref.zero_grad(set_to_none=True); par.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss_of(ref, xs).backward()
        loss_of(wrapped, xs[rank * n:(rank + 1) * n]).backward()
    torch.cuda.synchronize()
    worst, wn = 0.0, None
    for (k, a), b in zip(ref.named_parameters(), par.parameters()):
        assert b.grad is not None, k
        err = ((a.grad - b.grad).norm() / (a.grad.norm() + 1e-12)).item()
        if err > worst:
            worst, wn = err, k
    if rank == 0:
        print(f"{name}: max rel-L2 gradient difference = {worst:.3e} ({wn})")
    assert worst < tol, (name, wn, worst)


gg = torch.Generator(device="cpu").manual_seed(2)
check("TiTok (enc + VQ + dec)", lambda: M.TiTok(_TiTokCfg(32, 4, 8, 64, 12, "XS2")),
      lambda: torch.rand(4 * world, 3, 32, 32, generator=gg).to(dev),
      lambda m, xb: (lambda out: torch.nn.functional.mse_loss(out[0], xb) + out[2])(m(xb)))
check("VideoGPT", lambda: M.VideoGPT(_GPTCfg()),
      lambda: torch.randint(0, 64, (2 * world, 4, 16), generator=gg).to(dev),
      lambda m, xb: m(xb)[1])


# ---- blocks.py flavour: the ResidualAttentionBlock stack delivers its (LayerNorm-folded) gradients through the bucket sink ----
class _BlocksCfg:
    image_size, patch_size, transformer, latent_tokens, latent_dim = 64, 16, "small", 8, 12


class _TA(torch.nn.Module):     # encoder -> VectorQuantizer -> decoder as train_tatitok.TiTok wires them (train_tatitok.py:36-41)
    def __init__(self):
        super().__init__()
        M._BLOCKS_SIZES["small"] = (128, 2, 2)     # a miniature "small" for this check only
        self.encoder, self.decoder = M.BlocksTiTokEncoder(_BlocksCfg()), M.BlocksTiTokDecoder(_BlocksCfg())
        self.latent_tokens = torch.nn.Parameter(128 ** -0.5 * torch.randn(8, 128))
        self.quantize = M.VectorQuantizer(64, 12, 0.25, use_l2_norm=True)

    def forward(self, x):
        zq, info = self.quantize(self.encoder(x, self.latent_tokens))
        return self.decoder(zq), info["quantizer_loss"]


_small = M._BLOCKS_SIZES["small"]
check("blocks.py TiTokEncoder + VectorQuantizer + TiTokDecoder", _TA,
      lambda: torch.rand(4 * world, 3, 64, 64, generator=gg).to(dev),
      lambda m, xb: (lambda out: torch.nn.functional.mse_loss(out[0], xb) + out[1])(m(xb)))
M._BLOCKS_SIZES["small"] = _small

# ---- gradient accumulation: a no_sync() micro-step followed by a synchronised one reduces the sum of both ----
torch.manual_seed(0)
ref = M.ViTClassifier(cfg, num_classes=10).to(dev)
par = M.ViTClassifier(cfg, num_classes=10).to(dev)
par.load_state_dict(ref.state_dict())
wrapped = ddp.DataParallel(par, bucket_mb=4.0)
x2 = torch.randn(B, 3, 64, 64, generator=g).to(dev)
y2 = torch.randint(0, 10, (B,), generator=g).to(dev)
with torch.autocast("cuda", dtype=torch.bfloat16):
    (loss_fn(ref(x).float(), y) + loss_fn(ref(x2).float(), y2)).backward()
    sl = slice(rank * 8, rank * 8 + 8)
    with wrapped.no_sync():
        loss_fn(wrapped(x[sl]).float(), y[sl]).backward()
    loss_fn(wrapped(x2[sl]).float(), y2[sl]).backward()
torch.cuda.synchronize()
worst = max(((a.grad - b.grad).norm() / (a.grad.norm() + 1e-12)).item() for a, b in zip(ref.parameters(), par.parameters()))
if rank == 0:
    print(f"gradient accumulation (no_sync micro-step + synchronised step): max rel-L2 difference = {worst:.3e}")
assert worst < 2e-2, worst

# ---- bf16-compressed buckets: same check, tolerance = the bf16 rounding of the averaged gradient (2^-9 relative per element) ----
torch.manual_seed(0)
par = M.ViTClassifier(cfg, num_classes=10).to(dev)
par.load_state_dict(ref.state_dict())
wrapped = ddp.DataParallel(par, bucket_mb=4.0, compress_bf16=True)
ref.zero_grad(set_to_none=True)
with torch.autocast("cuda", dtype=torch.bfloat16):
    loss_fn(ref(x).float(), y).backward()
    loss_fn(wrapped(x[sl]).float(), y[sl]).backward()
torch.cuda.synchronize()
worst = max(((a.grad - b.grad).norm() / (a.grad.norm() + 1e-12)).item() for a, b in zip(ref.parameters(), par.parameters()))
if rank == 0:
    print(f"bf16-compressed all-reduce: max rel-L2 gradient difference = {worst:.3e}")
assert worst < 2.5e-2, worst
dist.destroy_process_group()
if rank == 0:
    print("DDP gradient check OK")
