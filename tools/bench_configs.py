"""Kernel / step timings for the other BASELINE.json configs (CUDA events), written as a markdown table.
    python tools/bench_configs.py > profiles/rNN_other_configs.md
configs[0] ViT-Ti 32px bs32, configs[2] TiTok-S 256px (encoder stack N=288 + VQ 4096x12), configs[3] ViT-L/16 224 bs256,
configs[4] VideoGPT-B causal N=1024 (transformer stack only)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-is-all-you-need_b200"))
import torch  # noqa: E402

from b200vit import modules as M  # noqa: E402
from b200vit import ops  # noqa: E402

dev = "cuda:0"
peaks_hbm = 6525.2


def timeit(fn, iters=5, warm=2):
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


def layer_flops(N, d, causal=False):
    att = 4 * N * N * d
    return 2 * N * d * 3 * d + (att // 2 if causal else att) + 16 * N * d * d


rows = []
M.transformer_configs.setdefault("Ti", lambda **kw: M.TransformerConfig(12, 3, 192, **kw))


def vit_step(name, size, patch, model, B, classes):
    torch.manual_seed(0)
    net = M.ViTClassifier(M.ViTConfig(size, 3, patch, model, 1, 0.0), num_classes=classes).to(dev)
    from b200vit.optim import AdamW
    opt = AdamW(net.parameters(), lr=1e-4)
    x = torch.randn(B, 3, size, size, device=dev)
    y = torch.randint(0, classes, (B,), device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = M.CrossEntropyLoss()(net(x), y)
        loss.backward()
        opt.step()
    ms = timeit(step)
    cfg = net.vit.transformer
    N = (size // patch) ** 2 + 1
    fl = 3 * cfg.n_layers * layer_flops(N, cfg.n_embd) * B
    rows.append((name, f"{ms:.2f} ms/step", f"{B / ms * 1e3:.0f} img/s", f"{fl / ms / 1e9:.0f} TFLOP/s (transformer layers only)"))
    del net, opt
    torch.cuda.empty_cache()


def stack_step(name, cfg, B, N, causal):
    torch.manual_seed(0)
    net = M.Transformer(cfg).to(dev)
    x = torch.randn(B, N, cfg.n_embd, device=dev, requires_grad=True)

    def step():
        net.zero_grad(set_to_none=True)
        net(x).float().square().mean().backward()
    ms = timeit(step)
    fl = 3 * cfg.n_layers * layer_flops(N, cfg.n_embd, causal) * B
    rows.append((name, f"{ms:.2f} ms fwd+bwd", f"{B / ms * 1e3:.0f} seq/s", f"{fl / ms / 1e9:.0f} TFLOP/s"))
    del net
    torch.cuda.empty_cache()


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


def titok_step(name, B):
    """configs[2]: train_titok.py:152-160 without the LPIPS/ConvNeXt perceptual term (out of scope): recon = MSE + L1 + VQ loss."""
    from b200vit.optim import AdamW
    torch.manual_seed(0)
    cfg = _TiTokCfg(256, 16, 32, 4096, 12, "S")
    net = M.TiTok(cfg).to(dev)
    opt = AdamW(net.parameters(), lr=1e-4)
    x = torch.rand(B, 3, 256, 256, device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            recon, idx, qloss = net(x)
            loss = torch.nn.functional.mse_loss(recon, x) + torch.nn.functional.l1_loss(recon, x) + qloss
        loss.backward()
        opt.step()
    ms = timeit(step)
    N = 288
    fl = 3 * 12 * layer_flops(N, 512) * B
    rows.append((name, f"{ms:.2f} ms/step", f"{B / ms * 1e3:.0f} img/s", f"{fl / ms / 1e9:.0f} TFLOP/s (transformer layers only)"))
    del net, opt
    torch.cuda.empty_cache()


vit_step("configs[0] ViT-Ti/4 32px, batch 32 (fwd+CE+bwd+AdamW)", 32, 4, "Ti", 32, 10)
vit_step("configs[0] shape at batch 4096", 32, 4, "Ti", 4096, 10)
vit_step("configs[3] ViT-L/16 224px, batch 256 (fwd+CE+bwd+AdamW)", 224, 16, "L", 256, 1000)
titok_step("configs[2] TiTok-S 256px tokenizer, 32 latent tokens, K=4096: enc + VQ + dec, fwd+bwd+AdamW, batch 256", 256)
stack_step("configs[2] TiTok-S encoder stack, N=288 (32 latent + 256 patches), batch 256", M.S(block_size=288), 256, 288, False)
stack_step("configs[4] VideoGPT-B causal stack, N=1024, batch 16", M.B(block_size=1024, causal=True), 16, 1024, True)

class _VideoGPTCfg:   # train_videogpt.VideoGPTConfig (train_videogpt.py:18-28)
    def __init__(self, frame_size, codebook_size, transformer, max_frames, dropout):
        self.frame_size, self.codebook_size, self.transformer = frame_size, codebook_size, transformer
        self.max_frames, self.dropout = max_frames, dropout
        self.max_tokens = max_frames * frame_size
        self.trans_config = M.transformer_configs[transformer](block_size=self.max_tokens, dropout=dropout, causal=True)
        self.n_embd = self.trans_config.n_embd


def videogpt_cases(B):
    """configs[4]: train_videogpt.py:130-134 training step, and generate() (train_videogpt.py:56-65): KV-cached vs the
    reference's algorithm (whole stack re-run per token) through the same kernels."""
    import time
    from b200vit.optim import AdamW
    torch.manual_seed(0)
    net = M.VideoGPT(_VideoGPTCfg(64, 1024, "B", 16, 0.0)).to(dev)
    opt = AdamW(net.parameters(), lr=1e-4)
    x = torch.randint(0, 1024, (B, 16, 64), device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            _, loss = net(x)
        loss.backward()
        opt.step()
    ms = timeit(step)
    fl = 3 * (12 * layer_flops(1024, 768, True) + 2 * 1024 * 768 * 1024) * B
    rows.append((f"configs[4] VideoGPT-B training step (embed + causal stack + vocab proj + CE + bwd + AdamW), N=1024, batch {B}",
                 f"{ms:.2f} ms/step", f"{B / ms * 1e3:.0f} seq/s", f"{fl / ms / 1e9:.0f} TFLOP/s"))
    net.eval()
    prompt = torch.randint(0, 1024, (B, 512), device=dev)
    n_new = 64
    net.generate(prompt, 4)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = net.generate(prompt, n_new)
    torch.cuda.synchronize(); t_kv = time.perf_counter() - t0

    def full_recompute(tokens, n):          # the reference's generate(): re-run everything for every token
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            for _ in range(n):
                sos = torch.full((tokens.shape[0], 1), 1024, device=dev, dtype=torch.long)
                xx = torch.cat([sos, tokens], dim=-1)
                hh = net.tok_embed(xx) + net.pos_embed(torch.arange(xx.shape[1], device=dev))
                lg = net.proj(net.transformer(hh)[:, -1])
                tokens = torch.cat([tokens, lg.argmax(-1, keepdim=True)], dim=-1)
        return tokens
    full_recompute(prompt, 2)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    full_recompute(prompt, n_new)
    torch.cuda.synchronize(); t_full = time.perf_counter() - t0
    rows.append((f"configs[4] VideoGPT-B generate: prompt 512 + {n_new} new tokens, batch {B}, KV-cached",
                 f"{t_kv * 1e3:.1f} ms", f"{B * n_new / t_kv:.0f} tok/s",
                 f"full re-computation per token (the reference's algorithm, same kernels): {t_full * 1e3:.1f} ms = {t_full / t_kv:.1f}x slower"))
    del net, opt
    torch.cuda.empty_cache()


videogpt_cases(16)

# VQ lookup: configs[2] (rows = B*32, K = 4096, D = 12) and the repo default (B*256 rows, K = 2048)
for R, K in ((256 * 32, 4096), (256 * 256, 2048), (16 * 16 * 64, 1024)):
    x = torch.randn(R, 12, device=dev)
    cb = torch.randn(K, 12, device=dev)
    ms = timeit(lambda: ops.vq_fwd(x, cb), iters=20)
    flops = 2.0 * R * K * 12
    bytes_ = R * 104 + K * 48
    rows.append((f"VQ lookup fwd rows={R} K={K} D=12 (bit-exact indices)", f"{ms * 1e3:.1f} us", f"{flops / ms / 1e9:.2f} TFLOP/s fp32 FMA",
                 f"{bytes_ / ms / 1e6:.1f} GB/s of {peaks_hbm:.0f} (104 B/row algorithmic; FMA-bound, see DESIGN.md)"))

print("# Other BASELINE.json configs on one B200 (CUDA events, synthetic data; parity for these shapes is in tests/)\n")
print("| case | time | rate | note |")
print("|---|---:|---:|---|")
for r in rows:
    print("| " + " | ".join(r) + " |")
