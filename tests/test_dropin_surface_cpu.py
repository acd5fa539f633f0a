"""Drop-in boundary checks that run on CPU: constructors, public attributes and state_dict keys/shapes of the
drop-in modules equal the reference's (compared live when /root/reference is present, and always against the key
lists frozen in the golden fixtures)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vit-is-all-you-need_b200")
REF = "/root/reference"


def test_transformer_surface_matches_fixture_keys(golden_dir):
    from b200vit import modules as M
    g = np.load(os.path.join(golden_dir, "transformer.npz"))
    cfg = M.TransformerConfig(n_layers=2, n_heads=1, n_embd=64, block_size=16, causal=True, dropout=0.0)
    assert cfg.head_dim == 64
    m = M.Transformer(cfg)
    for attr in ("n_layers", "n_heads", "n_embd", "block_size", "causal", "dropout", "head_dim"):
        assert hasattr(m, attr) and hasattr(m.layers[0], attr) and hasattr(m.layers[0].multi_attn, attr)
    ref_keys = sorted(k[len("b_w_"):] for k in g.files if k.startswith("b_w_"))
    ours = sorted(k for k in m.state_dict() if not k.endswith("mask"))
    assert ours == ref_keys
    for k in ours:
        assert tuple(m.state_dict()[k].shape) == g["b_w_" + k].shape
    assert m.state_dict()["layers.0.multi_attn.mask"].shape == (16, 16)
    assert set(M.transformer_configs) >= {"S", "B", "L"}
    assert (M.B(block_size=197).n_embd, M.L(block_size=197).n_layers, M.S(block_size=1).n_heads) == (768, 24, 8)


def test_resblock_and_vq_surface(golden_dir):
    from b200vit import modules as M
    g = np.load(os.path.join(golden_dir, "resblock.npz"))
    blk = M.ResidualAttentionBlock(128, 2)
    assert sorted(blk.state_dict()) == sorted(k[2:] for k in g.files if k.startswith("w_"))
    assert "mlp.c_fc.weight" in blk.state_dict() and "attn.in_proj_weight" in blk.state_dict()
    blk0 = M.ResidualAttentionBlock(128, 2, mlp_ratio=0)
    assert not any(k.startswith("mlp") or k.startswith("ln_2") for k in blk0.state_dict())
    vq = M.VectorQuantizer(codebook_size=32, token_size=12, use_l2_norm=True)
    assert list(vq.state_dict()) == ["embedding.weight"] and vq.embedding.weight.abs().max() <= 1 / 32
    with pytest.raises(NotImplementedError):
        M.VectorQuantizer(clustering_vq=True)
    e = vq.get_codebook_entry(torch.tensor([0, 5]))
    torch.testing.assert_close(e.norm(dim=-1), torch.ones(2))


def test_unsupported_configs_fail_loudly():
    from b200vit import modules as M
    with pytest.raises(NotImplementedError):
        M.Transformer(M.TransformerConfig(n_layers=1, n_heads=4, n_embd=128, block_size=8))  # head_dim 32
    with pytest.raises(ValueError):
        M.Transformer(M.TransformerConfig(n_layers=1, n_heads=2, n_embd=128, block_size=8, dropout=1.0))
    md = M.Transformer(M.TransformerConfig(n_layers=1, n_heads=2, n_embd=128, block_size=8, dropout=0.15))
    assert M._dropout_pair(md.layers[0]) == (0.15, 0.15)          # training: attention + MLP dropout
    assert M._dropout_pair(md.layers[0].eval()) == (0.15, 0.0)    # eval: SDPA dropout stays on (transformer.py:28)
    m = M.Transformer(M.TransformerConfig(n_layers=1, n_heads=2, n_embd=128, block_size=8))
    with pytest.raises(Exception):
        m(torch.zeros(1, 8, 128))  # CPU tensors: no fallback


_LIVE = r"""
import json, os, sys, types
mode, ref, pkg = sys.argv[1], sys.argv[2], sys.argv[3]
os.environ["WANDB_MODE"] = "disabled"
if mode == "dropin":
    sys.path.insert(0, pkg)
    from b200vit import launch
    launch.install_import_shims(ref)
    launch.install_class_swap(main_only=False)
else:
    sys.path.insert(0, ref)
    sys.modules["lpips"] = types.ModuleType("lpips")
    m = types.ModuleType("vector_quantize_pytorch"); m.FSQ = object; sys.modules["vector_quantize_pytorch"] = m
import torch
import transformer, train_vit, train_titok, train_videogpt, train_vit_vqgan
out = {}
transformer.transformer_configs.setdefault("Ti", lambda **kw: transformer.TransformerConfig(12, 3, 192, **kw))
vit = train_vit.ViTClassifier(train_vit.ViTConfig(32, 3, 4, "Ti", 1, 0.0), num_classes=10)
out["vit"] = {k: list(v.shape) for k, v in vit.state_dict().items()}
titok = train_titok.TiTok(train_titok.TiTokConfig(256, 16, 32, 4096, 12, "S"))
out["titok"] = {k: list(v.shape) for k, v in titok.state_dict().items()}
transformer.transformer_configs.setdefault("XS", lambda **kw: transformer.TransformerConfig(2, 1, 64, **kw))
gpt = train_videogpt.VideoGPT(train_videogpt.VideoGPTConfig(8, 32, "XS", 4, 0.0))
out["videogpt"] = {k: list(v.shape) for k, v in gpt.state_dict().items()}
vqgan = train_vit_vqgan.ViTVQGAN(train_vit_vqgan.ViTVQGANConfig(32, 4, 64, 12, "XS"))
out["vqgan"] = {k: list(v.shape) for k, v in vqgan.state_dict().items()}
import blocks
class _BCfg:
    image_size, patch_size, transformer, latent_tokens, latent_dim = 64, 16, "small", 8, 12
benc, bdec = blocks.TiTokEncoder(_BCfg()), blocks.TiTokDecoder(_BCfg())
out["blocks_enc"] = {k: list(v.shape) for k, v in benc.state_dict().items()}
out["blocks_dec"] = {k: list(v.shape) for k, v in bdec.state_dict().items()}
out["blocks_param_order"] = [n for n, _ in benc.named_parameters()] + [n for n, _ in bdec.named_parameters()]
out["blocks_classes"] = [type(benc).__module__, type(bdec).__module__, type(benc.transformer[0]).__module__,
                         blocks.VectorQuantizer.__module__, blocks.TATiTokDecoder.__module__]
out["classes"] = [type(titok.enc.vit).__module__, type(titok.quant).__module__, type(vit.vit.transformer).__module__,
                  type(titok.enc).__module__, type(titok.dec).__module__, type(gpt).__module__, type(vqgan.encoder).__module__,
                  type(vqgan.decoder).__module__]
print("RESULT" + json.dumps(out))
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_launcher_swaps_classes_and_keeps_state_dict_contract():
    def run(mode):
        r = subprocess.run([sys.executable, "-c", _LIVE, mode, REF, PKG], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][-1]
        return json.loads(line[len("RESULT"):])
    ref, ours = run("reference"), run("dropin")
    assert ours["vit"] == ref["vit"], "ViTClassifier state_dict keys/shapes must equal the reference's"
    assert ours["titok"] == ref["titok"], "TiTok (train_titok.py) state_dict keys/shapes must equal the reference's"
    assert ours["videogpt"] == ref["videogpt"], "VideoGPT (train_videogpt.py) state_dict keys/shapes must equal the reference's"
    assert ours["vqgan"] == ref["vqgan"], "ViTVQGAN (train_vit_vqgan.py) state_dict keys/shapes must equal the reference's"
    assert all(c.startswith("b200vit") for c in ours["classes"]), ours["classes"]
    assert not any(c.startswith("b200vit") for c in ref["classes"])
    # blocks.py: TiTokEncoder / TiTokDecoder / ResidualAttentionBlock / VectorQuantizer are the drop-ins, TATiTokDecoder (out of
    # scope) is still the reference's own class, and keys / shapes / parameter order are the reference's
    assert ours["blocks_enc"] == ref["blocks_enc"] and ours["blocks_dec"] == ref["blocks_dec"]
    assert ours["blocks_param_order"] == ref["blocks_param_order"]
    assert all(c.startswith("b200vit") for c in ours["blocks_classes"][:4]), ours["blocks_classes"]
    assert not ours["blocks_classes"][4].startswith("b200vit")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_unchanged_reference_script_runs_to_the_kernel_boundary(tmp_path):
    """`python -m b200vit.launch <reference>/train_vit.py ...` with the script UNCHANGED: imports resolve to the drop-ins,
    the model / optimizer / loss are built, the first forward reaches the C ABI -- and, on this GPU-less container, fails
    loudly there instead of falling back to PyTorch."""
    env = dict(os.environ, B200VIT_SYNTHETIC="1", B200VIT_SYNTHETIC_SAMPLES="4", PYTHONPATH=PKG, WANDB_MODE="disabled")
    r = subprocess.run([sys.executable, "-m", "b200vit.launch", os.path.join(REF, "train_vit.py"), "--transformer", "S",
                        "--image_size", "32", "--patch_size", "4", "--bs", "2", "--epochs", "1"],
                       capture_output=True, text=True, timeout=600, cwd=str(tmp_path), env=env)
    out = r.stdout + r.stderr
    assert "STATS: params=" in out, out[-1500:]                 # the script got as far as building everything
    if not torch.cuda.is_available():
        assert r.returncode != 0 and "no CPU fallback" in out, out[-1500:]


def test_operand_cache_epoch_and_launcher_swaps_on_cpu():
    """Host logic that needs no GPU: (1) ANY optimizer step invalidates the bf16 operand caches (torch's fused optimisers do
    not bump Tensor._version); (2) the launcher's loss / optimizer swaps install the drop-ins and leave unsupported uses to
    torch (CPU tensors here)."""
    import torch.nn.functional as F
    from b200vit import functional as Fn
    from b200vit import launch, optim
    p = torch.nn.Parameter(torch.randn(8))
    p.grad = torch.randn(8)
    k0 = Fn.bf16_key(p)
    torch.optim.SGD([p], lr=0.1).step()
    k1 = Fn.bf16_key(p)
    assert k1 != k0 and k1[0] == k0[0] + 1                       # the epoch advanced (and the version, for SGD)
    orig_ce, orig_adamw = F.cross_entropy, torch.optim.AdamW
    try:
        launch.install_loss_swap()
        launch.install_optimizer_swap()
        assert torch.optim.AdamW is optim.AdamW
        x, y = torch.randn(5, 7), torch.randint(0, 7, (5,))
        torch.testing.assert_close(F.cross_entropy(x, y), orig_ce(x, y))                       # CPU: torch's path
        torch.testing.assert_close(torch.nn.CrossEntropyLoss(label_smoothing=0.1)(x, y), orig_ce(x, y, label_smoothing=0.1))
    finally:
        F.cross_entropy, torch.optim.AdamW = orig_ce, orig_adamw
