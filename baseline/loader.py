"""Imports the UNMODIFIED reference modules (baseline/_ref, installed by tools/install_reference.py; /root/reference in the
build container as a fallback) without leaving their bare top-level names (`transformer`, `train_vit`, `blocks`, `utils`,
`datasets` ...) in sys.modules / sys.path, so they cannot shadow the drop-in shims or third-party packages of the same name.

Used only by the checkers and baselines: tests/, bench.py's cpu_baseline / --impl reference / gpu_eager_baseline legs.  No
module under vit-is-all-you-need_b200/ imports this.

Two third-party imports the reference never uses are stubbed when missing (`lpips` train_titok.py:1,
`vector_quantize_pytorch.FSQ` train_titok.py:10), and wandb is disabled.
"""
import importlib
import os
import sys
import types
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("B200VIT_REFERENCE", "/root/reference"))
_BARE = ("transformer", "train_vit", "train_titok", "train_vit_vqgan", "train_videogpt", "train_tatitok", "blocks", "utils",
         "datasets", "perceptual_loss", "lpips", "vector_quantize_pytorch")
_cache = {}


def reference_dir():
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "transformer.py")):
            return c
    return None


def available():
    return reference_dir() is not None


class _Namespace(types.SimpleNamespace):
    pass


def load(names=("transformer", "train_vit", "train_titok", "train_vit_vqgan", "train_videogpt", "blocks")):
    """Returns a namespace with the requested reference modules as attributes (e.g. ref.transformer.Transformer,
    ref.train_vit.ViTClassifier) plus `ref.dir`.  Raises FileNotFoundError when no reference copy is present."""
    key = tuple(names)
    if key in _cache:
        return _cache[key]
    d = reference_dir()
    if d is None:
        raise FileNotFoundError("no reference copy: run tools/install_reference.py in the build container (baseline/_ref)")
    os.environ.setdefault("WANDB_MODE", "disabled")
    stash = {n: sys.modules.pop(n) for n in _BARE if n in sys.modules}
    sys.path.insert(0, d)
    try:
        try:
            importlib.import_module("lpips")
        except Exception:
            sys.modules["lpips"] = types.ModuleType("lpips")
        try:
            importlib.import_module("vector_quantize_pytorch")
        except Exception:
            m = types.ModuleType("vector_quantize_pytorch")
            m.FSQ = object
            sys.modules["vector_quantize_pytorch"] = m
        ns = _Namespace(dir=d)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):      # blocks.py prints "attention mode is ..." at import
                for n in names:
                    setattr(ns, n, importlib.import_module(n))
    finally:
        sys.path.remove(d)
        for n in _BARE:
            sys.modules.pop(n, None)
        sys.modules.update(stash)
    # BASELINE.json configs[0] names a "Tiny" preset the reference does not define
    if hasattr(ns, "transformer"):
        tr = ns.transformer
        tr.transformer_configs.setdefault("Ti", lambda **kw: tr.TransformerConfig(12, 3, 192, **kw))
    _cache[key] = ns
    return ns
