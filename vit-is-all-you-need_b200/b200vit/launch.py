"""Runs one of the reference's training scripts UNCHANGED on top of the drop-in modules:

    python -m b200vit.launch /path/to/reference/train_vit.py --transformer B --dropout 0.0 --epochs 1 ...
    B200VIT_SYNTHETIC=1 python -m b200vit.launch /path/to/reference/train_titok.py ...

What it does:
  * puts ../shim ahead of the reference directory on sys.path, so `from transformer import ...`,
    `from train_vit import ViTConfig, ViT` and `import blocks` resolve to the sm_100a-backed classes;
  * classes the executed script defines itself (`ViT` in train_vit.py:30, `Quantizer` in train_titok.py:45 /
    train_vit_vqgan.py:45) are swapped at class-creation time through builtins.__build_class__;
  * torch.nn.Conv2d becomes modules.PatchConv2d, a subclass that routes patchify-shaped convolutions (kernel == stride)
    through the im2col + tcgen05 GEMM path and is otherwise nn.Conv2d;
  * torch.optim.AdamW becomes b200vit.optim.AdamW (one fused multi-tensor launch that also refreshes the bf16 GEMM operands;
    same constructor, state_dict layout and GradScaler protocol; B200VIT_TORCH_ADAMW=1 keeps torch's);
  * F.cross_entropy (and with it nn.CrossEntropyLoss) takes the fused cross-entropy kernels for 2-D CUDA logits;
  * stubs two imports the scripts never use (lpips, vector_quantize_pytorch.FSQ) when they are not installed,
    disables wandb, and (B200VIT_SYNTHETIC=1) replaces the hard-coded ImageNet loaders with synthetic ones and
    perceptual_loss.PerceptualLoss (torchvision ConvNeXt-S whose ImageNet weights are a network download,
    perceptual_loss.py:41) with a small frozen, seeded feature network of the same interface.

Harness knobs (none of them changes what the script computes):
    B200VIT_SEED=<int>        torch.manual_seed before the script starts (the scripts never seed)
    B200VIT_LOG_JSONL=<path>  every wandb.log({...}) call of the script also appends its scalar entries as one JSON line
    B200VIT_PLAIN=1           install only the harness pieces (stubs, synthetic data, log, seed) and NOT the drop-ins: the
                              un-accelerated comparison run of tests/test_scripts_gpu.py
"""
import builtins
import os
import runpy
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(os.path.dirname(HERE), "shim")

# name defined by the script -> name in b200vit.modules (train_vit.py:30; train_titok.py:34,45,61 == train_vit_vqgan.py)
_SWAP = {"ViT": "ViT", "Quantizer": "Quantizer", "TiTokEncoder": "TiTokEncoder", "TiTokDecoder": "TiTokDecoder",
         "ViTVQGANEncoder": "TiTokEncoder", "ViTVQGANDecoder": "TiTokDecoder", "VideoGPT": "VideoGPT"}   # train_vit_vqgan.py:34,61: same code, all tokens


def install_import_shims(reference_dir):
    for p in (reference_dir, SHIM, os.path.dirname(HERE)):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, reference_dir)
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, SHIM)  # shim first: bare `transformer`, `train_vit`, `blocks` resolve here
    os.environ.setdefault("WANDB_MODE", "disabled")
    _stub_unused_imports()


def _stub_unused_imports():
    try:
        import lpips  # noqa: F401
    except Exception:
        sys.modules["lpips"] = types.ModuleType("lpips")
    try:
        import vector_quantize_pytorch  # noqa: F401
    except Exception:
        m = types.ModuleType("vector_quantize_pytorch")
        m.FSQ = object
        sys.modules["vector_quantize_pytorch"] = m


def install_class_swap(main_only=True):
    """Classes named in _SWAP that the script defines as nn.Module subclasses are replaced by the drop-ins."""
    import torch.nn as nn

    from . import modules
    orig = builtins.__build_class__

    def build_class(func, name, *bases, **kw):
        if name in _SWAP and any(isinstance(b, type) and issubclass(b, nn.Module) for b in bases):
            mod = func.__globals__.get("__name__", "")
            if not main_only or mod == "__main__":
                return getattr(modules, _SWAP[name])
        return orig(func, name, *bases, **kw)

    builtins.__build_class__ = build_class
    return orig


def install_conv_swap():
    """torch.nn.Conv2d -> PatchConv2d: patchify-shaped convolutions that the reference builds itself
    (blocks.TiTokEncoder.patch_embed blocks.py:235, the decoders' 1x1 convolutions) take the im2col + tcgen05 GEMM path;
    all other convolutions behave exactly as before (PatchConv2d falls through to nn.Conv2d)."""
    import torch.nn as nn

    from . import modules
    orig = nn.Conv2d
    nn.Conv2d = modules.PatchConv2d
    nn.modules.conv.Conv2d = modules.PatchConv2d
    return orig


def install_optimizer_swap():
    """torch.optim.AdamW -> b200vit.optim.AdamW (train_vit.py:82, train_titok.py:134, train_vit_vqgan.py:131,
    train_videogpt.py:107 all call torch.optim.AdamW(model.parameters(), lr=..., weight_decay=...))."""
    import torch

    from . import optim
    orig = torch.optim.AdamW
    torch.optim.AdamW = optim.AdamW
    return orig


def install_loss_swap():
    """nn.CrossEntropyLoss (train_vit.py:81) and F.cross_entropy (train_videogpt.py:54) -> the fused kernels of
    csrc/head_ce.cu for what those call sites pass (2-D CUDA logits, int64 labels, default arguments); any other use keeps
    torch's implementation, like PatchConv2d does for non-patchify convolutions."""
    import torch
    import torch.nn.functional as F

    from . import functional as Fn
    orig_fn = F.cross_entropy

    def cross_entropy(input, target, weight=None, size_average=None, ignore_index=-100, reduce=None, reduction="mean",
                      label_smoothing=0.0):
        if (input.is_cuda and input.dim() == 2 and target.dtype == torch.int64 and target.dim() == 1 and weight is None
                and size_average is None and reduce is None and reduction == "mean" and label_smoothing == 0.0
                and input.dtype in (torch.float32, torch.bfloat16)):
            return Fn.CrossEntropyFn.apply(input, target, ignore_index)
        return orig_fn(input, target, weight=weight, size_average=size_average, ignore_index=ignore_index, reduce=reduce,
                       reduction=reduction, label_smoothing=label_smoothing)

    F.cross_entropy = cross_entropy   # nn.CrossEntropyLoss.forward calls F.cross_entropy, so both entry points are covered
    return orig_fn


def install_synthetic_loaders():
    """datasets.get_imagenet_loaders has a hard-coded dataset root (datasets.py:7,23); benchmarks and smoke runs
    use synthetic tensors of the same shapes instead."""
    import datasets
    import torch

    class _Synthetic(torch.utils.data.Dataset):
        def __init__(self, n, size):
            self.n, self.size = n, size

        def __len__(self):
            return self.n

        def __getitem__(self, i):
            g = torch.Generator().manual_seed(i)
            return torch.rand(3, self.size, self.size, generator=g), int(torch.randint(0, 1000, (1,), generator=g))

    def get_imagenet_loaders(image_size, bs, *a, **k):
        n = int(os.environ.get("B200VIT_SYNTHETIC_SAMPLES", 64 * bs))
        mk = lambda m: torch.utils.data.DataLoader(_Synthetic(m, image_size), batch_size=bs, shuffle=False, num_workers=0)  # noqa: E731
        return mk(n), mk(max(bs, n // 8))

    datasets.get_imagenet_loaders = get_imagenet_loaders


def install_perceptual_loss_stub():
    """perceptual_loss.PerceptualLoss (perceptual_loss.py:27-70) needs torchvision's ConvNeXt-S ImageNet weights, a network
    download (perceptual_loss.py:41).  Synthetic runs get a module of the same interface -- frozen, eval-mode, inputs in
    [0, 1] resized to a fixed square, feature MSE -- over a small seeded convolutional network (out of scope for the hot
    path; plain torch, identical in the accelerated and the plain run)."""
    import torch

    class PerceptualLoss(torch.nn.Module):
        def __init__(self, model_name: str = "convnext_s"):
            super().__init__()
            if "convnext_s" not in model_name:
                raise ValueError(f"Unsupported Perceptual Loss model name {model_name}")
            conv = torch.nn.Conv2d     # padded 3x3 / 4x4 convolutions: PatchConv2d (if installed) passes them through
            with torch.random.fork_rng(devices=[]):
                torch.manual_seed(20260101)
                self.features = torch.nn.Sequential(
                    conv(3, 32, 4, stride=4, padding=1), torch.nn.GELU(),
                    conv(32, 64, 3, stride=2, padding=1), torch.nn.GELU(),
                    conv(64, 128, 3, stride=2, padding=1), torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(),
                    torch.nn.Linear(128, 1000))
            self.register_buffer("imagenet_mean", torch.tensor([0.485, 0.456, 0.406])[None, :, None, None])
            self.register_buffer("imagenet_std", torch.tensor([0.229, 0.224, 0.225])[None, :, None, None])
            for p in self.parameters():
                p.requires_grad = False

        def forward(self, input, target):
            self.eval()
            F = torch.nn.functional
            input = F.interpolate(input.float(), size=64, mode="bilinear", align_corners=False, antialias=True)
            target = F.interpolate(target.float(), size=64, mode="bilinear", align_corners=False, antialias=True)
            a = self.features((input - self.imagenet_mean) / self.imagenet_std)
            b = self.features((target - self.imagenet_mean) / self.imagenet_std)
            return F.mse_loss(a, b, reduction="mean")

    m = types.ModuleType("perceptual_loss")
    m.PerceptualLoss = PerceptualLoss
    m.__doc__ = "synthetic stand-in installed by b200vit.launch (B200VIT_SYNTHETIC=1)"
    sys.modules["perceptual_loss"] = m


def install_log_hook(path):
    """wandb.log is the scripts' only record of the loss (train_vit.py:109, train_titok.py:167); mirror its scalars to a file."""
    import json

    import wandb

    def log(data=None, *a, **k):
        try:
            row = {}
            for key, v in (data or {}).items():
                if hasattr(v, "item") and getattr(v, "numel", lambda: 1)() == 1:
                    v = v.item()
                if isinstance(v, (int, float)):
                    row[key] = v
            if row:
                with open(path, "a") as f:
                    f.write(json.dumps(row) + "\n")
        except Exception:   # pragma: no cover  (logging must never break the run)
            pass
        return log.inner(data, *a, **k)

    log.inner = wandb.log
    orig_init = wandb.init

    def init(*a, **k):      # wandb.init() re-binds the module-level wandb.log to the new run: wrap it again afterwards
        run = orig_init(*a, **k)
        if wandb.log is not log:
            log.inner = wandb.log
            wandb.log = log
        return run

    wandb.init = init
    wandb.log = log


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    plain = os.environ.get("B200VIT_PLAIN", "0") == "1"
    if plain:
        ref_dir = os.path.dirname(script)
        if ref_dir in sys.path:
            sys.path.remove(ref_dir)
        sys.path.insert(0, ref_dir)
        os.environ.setdefault("WANDB_MODE", "disabled")
        _stub_unused_imports()
    else:
        install_import_shims(os.path.dirname(script))
    if os.environ.get("B200VIT_SYNTHETIC", "0") == "1":
        install_synthetic_loaders()
        install_perceptual_loss_stub()
    if os.environ.get("B200VIT_LOG_JSONL"):
        install_log_hook(os.environ["B200VIT_LOG_JSONL"])
    orig = builtins.__build_class__
    if not plain:
        orig = install_class_swap()
        install_conv_swap()
        if os.environ.get("B200VIT_TORCH_ADAMW", "0") != "1":
            install_optimizer_swap()
        install_loss_swap()
    if os.environ.get("B200VIT_SEED"):
        import torch
        torch.manual_seed(int(os.environ["B200VIT_SEED"]))
    sys.argv = [script] + argv[1:]
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        builtins.__build_class__ = orig


if __name__ == "__main__":
    main()
