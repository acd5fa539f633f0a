"""Places the UNMODIFIED reference checkout under baseline/_ref/ (git-ignored, NOT gpurun-ignored), so that it travels to
the GPU box with the repository snapshot and can be used there as

  * the parity anchor of tests/test_reference_parity_gpu.py and tests/test_scripts_gpu.py (the real modules / scripts),
  * bench.py's `cpu_baseline` (kind "reference"), `--impl reference` arm and `gpu_eager_baseline`.

    python tools/install_reference.py [--src /root/reference]

The base contract's `pip install --target baseline/_ref /root/reference` cannot work: the reference has neither setup.py nor
pyproject.toml ("Directory ... is not installable", recorded in DESIGN.md §4); it is a flat directory of top-level modules
that import each other by bare name, so a byte-for-byte copy of that directory is the install.  Nothing from it is ever
committed (baseline/_ref/ is in .gitignore) and no product module imports it.
"""
import argparse
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
DEFAULT_SRC = os.environ.get("B200VIT_REFERENCE", "/root/reference")


def install(src: str = DEFAULT_SRC, dest: str = DEST, verbose: bool = False) -> str:
    """Copies src -> dest (only files that differ).  Returns dest, or "" when src does not exist (GPU box: the copy made in
    the build container is already there)."""
    if not os.path.isfile(os.path.join(src, "transformer.py")):
        return dest if os.path.isfile(os.path.join(dest, "transformer.py")) else ""
    n = 0
    for dirpath, dirnames, filenames in os.walk(src):
        dirnames[:] = [d for d in dirnames if d not in (".git", "__pycache__")]
        rel = os.path.relpath(dirpath, src)
        out = os.path.join(dest, rel) if rel != "." else dest
        os.makedirs(out, exist_ok=True)
        for f in filenames:
            if f.endswith(".pyc"):
                continue
            s, d = os.path.join(dirpath, f), os.path.join(out, f)
            if not os.path.exists(d) or not filecmp.cmp(s, d, shallow=False):
                shutil.copyfile(s, d)
                n += 1
    if verbose:
        print(f"baseline/_ref: {n} file(s) copied from {src}", file=sys.stderr)
    return dest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=DEFAULT_SRC)
    a = ap.parse_args()
    print(install(a.src, verbose=True) or "reference not found")
