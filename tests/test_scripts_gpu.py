"""Script-level parity on a B200 (SURVEY.md §4 "script-level"): the reference's training scripts run UNCHANGED, once through
`python -m b200vit.launch` (drop-in modules: hand-written kernels) and once plain (B200VIT_PLAIN=1: the reference's own
modules on PyTorch's library kernels), from the same seed and the same synthetic data, exactly as written -- including
`autocast("cuda")` (= fp16 for the plain run; the drop-ins compute in bf16 whatever the autocast dtype) and
`GradScaler` (scaled loss through the fused cross-entropy / VQ backward, `found_inf` consumed on the device by the fused
AdamW) -- and the loss trajectories the scripts log through wandb.log are compared step by step.

    train_vit.py:92-111      ViTClassifier, CrossEntropyLoss, AdamW, LR schedule, GradScaler, validation pass, torch.save
    train_titok.py:144-172   TiTok (encoder + quantiser + decoder), MSE + perceptual (synthetic stand-in) + VQ losses,
                             clip_grad_norm_ after the step, codebook-usage bookkeeping, checkpoint

The scripts live in baseline/_ref (tools/install_reference.py); skipped when that copy is absent."""
import json
import os
import subprocess
import sys

import pytest

from baseline import loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vit-is-all-you-need_b200")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not loader.available(), reason="baseline/_ref (reference copy) not present")]

# measured on a B200 (profiles/r2_script_trajectories.json): the two trajectories agree to 1e-4 .. 3e-4 relative at every step
FIRST_TOL = 5e-3     # first logged loss (no optimizer step in between, or one): relative
TRAJ_TOL = 2e-2      # every later logged loss: relative (fp16-autocast eager vs bf16 kernels drift apart slowly)


def _run(script, args, workdir, plain, samples, extra_env=None):
    log = os.path.join(workdir, "plain.jsonl" if plain else "dropin.jsonl")
    env = dict(os.environ, B200VIT_SYNTHETIC="1", B200VIT_SYNTHETIC_SAMPLES=str(samples), B200VIT_SEED="7",
               B200VIT_LOG_JSONL=log, PYTHONPATH=PKG, WANDB_MODE="disabled", CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0"))
    if plain:
        env["B200VIT_PLAIN"] = "1"
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, "-m", "b200vit.launch", os.path.join(loader.reference_dir(), script)] + args,
                       capture_output=True, text=True, timeout=900, cwd=workdir, env=env)
    assert r.returncode == 0, f"{script} ({'plain' if plain else 'drop-in'}) failed:\n{(r.stdout + r.stderr)[-3000:]}"
    rows = [json.loads(ln) for ln in open(log)] if os.path.exists(log) else []
    return rows, r.stdout + r.stderr


def _save(name, plain, ours):
    out = os.environ.get("B200VIT_SCRIPT_REPORT_DIR")
    if out:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, name + ".json"), "w") as f:
            json.dump({"plain": plain, "dropin": ours}, f, indent=1)


def _check(key, plain, ours):
    a = [r[key] for r in plain if key in r]
    b = [r[key] for r in ours if key in r]
    assert len(a) == len(b) and len(a) >= 3, (key, len(a), len(b))
    for i, (x, y) in enumerate(zip(a, b)):
        tol = FIRST_TOL if i == 0 else TRAJ_TOL
        assert abs(x - y) <= tol * abs(x) + 1e-6, f"{key} step {i}: plain {x} vs drop-in {y} (all: {a} vs {b})"


def test_train_vit_script_unchanged(tmp_path):
    args = ["--transformer", "S", "--image_size", "64", "--patch_size", "8", "--bs", "16", "--epochs", "1", "--dropout", "0.0",
            "--warmup_steps", "4", "--lr", "1e-4"]
    ours, out_o = _run("train_vit.py", args, str(tmp_path), False, 16 * 12)
    plain, out_p = _run("train_vit.py", args, str(tmp_path), True, 16 * 12)
    _save("train_vit", plain, ours)
    assert "STATS: params=" in out_o and os.path.exists(tmp_path / "vit.pth")      # validation pass + checkpoint reached
    _check("train/loss", [r for r in plain if "train/loss" in r], [r for r in ours if "train/loss" in r])
    va = [r["valid/acc"] for r in ours if "valid/acc" in r]
    assert len(va) == 1
    # the checkpoint the accelerated run wrote loads into the reference's own class (state_dict contract)
    import torch
    ref = loader.load()
    sd = torch.load(tmp_path / "vit.pth", map_location="cpu")
    model = ref.train_vit.ViTClassifier(ref.train_vit.ViTConfig(64, 3, 8, "S", 1, 0.0))
    model.load_state_dict(sd, strict=True)


@pytest.mark.parametrize("script", ["train_titok.py", "train_vit_vqgan.py"])
def test_train_titok_script_unchanged(tmp_path, script):
    """train_titok.py (TiTok: 16 latent tokens) and its all-tokens twin train_vit_vqgan.py (ViT-VQGAN: one latent per patch, L1
    reconstruction loss, train_vit_vqgan.py:150-157): same harness, same comparison."""
    os.makedirs(tmp_path / "titok_models", exist_ok=True)      # train_titok.py:172 saves there without creating it
    args = ["--image_size", "64", "--patch_size", "8", "--codebook_size", "512", "--latent_dim", "12",
            "--transformer", "S", "--bs", "8", "--epochs", "5", "--warmup_steps", "2", "--lr", "1e-4"]
    if script == "train_titok.py":
        args += ["--latent_tokens", "16"]
    ours, out_o = _run(script, args, str(tmp_path), False, 16)
    plain, out_p = _run(script, args, str(tmp_path), True, 16)
    _save(script[:-3], plain, ours)
    assert "STATS: enc_params=" in out_o
    for key in ("train/loss", "train/l1_loss", "train/quant_loss"):
        _check(key, plain, ours)
    assert os.listdir(tmp_path / "titok_models"), "train_titok.py:170-172 checkpoint was not written"
