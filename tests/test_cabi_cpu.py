"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/b200vit.h declares; the binding table covers all of them; the product path refuses CPU tensors."""
import ctypes

import pytest
import torch

from b200vit import _cabi


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    declared = _cabi.declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"libb200vit.so does not export: {missing}"
    assert lib.b200vit_version() == 100


def test_binding_table_matches_header():
    declared = set(_cabi.declared_symbols())
    bound = set(_cabi._SIGNATURES) | {"b200vit_last_error"}
    assert declared == bound, f"header-only: {declared - bound}; binding-only: {bound - declared}"


def test_no_cpu_fallback():
    from b200vit import ops
    x = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(_cabi.B200VitError):
        ops.gemm_bias(x, x, None)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a CPU-only box")
def test_init_fails_loudly_without_gpu():
    lib = _cabi.load()
    assert lib.b200vit_init(ctypes.c_int(0)) != 0
    assert len(lib.b200vit_last_error()) > 0


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under vit-is-all-you-need_b200/ may import or execute it."""
    import glob
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-is-all-you-need_b200")
    offenders = []
    for path in glob.glob(os.path.join(root, "**", "*.py"), recursive=True):
        for n, line in enumerate(open(path), 1):
            if re.search(r"^\s*(from|import)\s+oracle\b", line) or "oracle/" in line and "subprocess" in line:
                offenders.append(f"{path}:{n}")
    assert not offenders, offenders


def test_header_is_plain_c(tmp_path):
    """include/b200vit.h is a C ABI: it must compile as C99 with nothing but the standard headers."""
    import os
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hdr.c"
    src.write_text('#include "b200vit.h"\nint (*probe)(void) = b200vit_version;\nint main(void) { return probe == 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-c", str(src), "-I", os.path.join(root, "include"), "-o",
                        str(tmp_path / "hdr.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
