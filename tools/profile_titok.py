"""CUPTI kernel table of one TiTok-S tokenizer training step (BASELINE configs[2]: 256 px, 32 latent tokens, K = 4096, batch 256):
encoder ViT -> proj -> VQ lookup -> quant_proj -> decoder ViT -> de-patchify GEMM; MSE + L1 + VQ loss; AdamW."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-is-all-you-need_b200")); sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile
from b200vit import modules as M
from b200vit.optim import AdamW
dev = "cuda:0"


class Cfg:   # train_titok.TiTokConfig (train_titok.py:18-32)
    def __init__(self, image_size, patch_size, latent_tokens, codebook_size, latent_dim, transformer):
        self.image_size, self.patch_size, self.latent_tokens = image_size, patch_size, latent_tokens
        self.codebook_size, self.latent_dim, self.transformer = codebook_size, latent_dim, transformer
        self.patch_dim = image_size // patch_size
        self.n_patches = self.patch_dim ** 2
        self.enc_vit_config = M.ViTConfig(image_size, 3, patch_size, transformer, latent_tokens, 0.0)
        self.n_embd = self.enc_vit_config.trans_config.n_embd
        self.dec_vit_config = M.ViTConfig(latent_tokens, self.n_embd, 1, transformer, self.n_patches, 0.0)
        self.dec_vit_config.n_patches = latent_tokens


torch.manual_seed(0)
net = M.TiTok(Cfg(256, 16, 32, 4096, 12, "S")).to(dev)
opt = AdamW(net.parameters(), lr=1e-4)
x = torch.rand(256, 3, 256, 256, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        recon, idx, qloss = net(x)
        loss = torch.nn.functional.mse_loss(recon, x) + torch.nn.functional.l1_loss(recon, x) + qloss
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record(); torch.cuda.synchronize()
print(f"# TiTok-S tokenizer step, batch 256: {e0.elapsed_time(e1) / 5:.2f} ms/step (CUDA events, un-profiled)\n")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
rows = sorted(((k.key, k.device_time_total / 2, k.count // 2) for k in prof.key_averages() if k.device_time_total > 0), key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print("| kernel | launches/step | us/step | share |\n|---|---:|---:|---:|")
for name, us, n in rows[:28]:
    print(f"| `{name[:100]}` | {n} | {us:.1f} | {100 * us / tot:.1f}% |")
print(f"\nGPU kernel time per step: {tot / 1e3:.2f} ms")
