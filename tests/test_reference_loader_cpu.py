"""baseline/loader.py and tools/install_reference.py (CPU): the unmodified reference is importable for the checkers without
leaving its bare top-level module names behind (they would shadow the drop-in shims and third-party packages), and the
installer is a byte-for-byte, idempotent copy that is never part of the committed tree."""
import filecmp
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import loader  # noqa: E402

needs_ref = pytest.mark.skipif(not loader.available(), reason="no reference copy (baseline/_ref or /root/reference)")


@needs_ref
def test_loader_isolates_the_reference_modules():
    before_path = list(sys.path)
    before_mods = {n: sys.modules.get(n) for n in ("transformer", "blocks", "utils", "datasets", "train_vit")}
    ref = loader.load(("transformer", "train_vit", "blocks"))
    assert sys.path == before_path
    for n, m in before_mods.items():
        assert sys.modules.get(n) is m, f"{n} leaked into sys.modules"
    assert ref.transformer.Transformer.__module__ == "transformer" and ref.dir == loader.reference_dir()
    assert "Ti" in ref.transformer.transformer_configs                      # BASELINE configs[0] preset registered
    cfg = ref.train_vit.ViTConfig(32, 3, 4, "Ti", 1, 0.0)
    assert cfg.trans_config.n_embd == 192 and cfg.trans_config.head_dim == 64
    assert loader.load(("transformer", "train_vit", "blocks")) is ref       # cached
    # the drop-in package never imports it
    pkg = os.path.join(ROOT, "vit-is-all-you-need_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "baseline" not in src or "baseline/_ref" in src or "BASELINE" in src, os.path.join(dirpath, f)
                assert "from baseline" not in src and "import baseline" not in src, os.path.join(dirpath, f)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="build container only")
def test_install_reference_is_an_exact_idempotent_copy(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import install_reference
    dest = str(tmp_path / "_ref")
    assert install_reference.install("/root/reference", dest) == dest
    for f in ("transformer.py", "blocks.py", "train_vit.py", "train_titok.py", os.path.join("scripts", "vit_sweep.yaml")):
        assert filecmp.cmp(os.path.join("/root/reference", f), os.path.join(dest, f), shallow=False)
    mt = os.path.getmtime(os.path.join(dest, "blocks.py"))
    install_reference.install("/root/reference", dest)                      # second run copies nothing
    assert os.path.getmtime(os.path.join(dest, "blocks.py")) == mt
    # and the copy is not tracked by git
    out = subprocess.run(["git", "ls-files", "baseline/_ref"], capture_output=True, text=True, cwd=ROOT).stdout
    assert out.strip() == ""
