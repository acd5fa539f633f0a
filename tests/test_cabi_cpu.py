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


def test_adamw_chunk_table_covers_every_element_once():
    """Host logic of b200vit.optim.AdamW: the (tensor, chunk) list handed to the kernel covers every element of every tensor
    exactly once, whatever the sizes (the kernel processes b200vit_adamw_chunk_elems() elements per CTA)."""
    import numpy as np
    import torch
    from b200vit import _cabi
    from b200vit.optim import AdamW
    chunk = _cabi.load().b200vit_adamw_chunk_elems()
    assert chunk > 0 and chunk % 4 == 0
    sizes = [1, 3, chunk - 1, chunk, chunk + 1, 3 * chunk + 17, 37]
    params = [torch.nn.Parameter(torch.zeros(n)) for n in sizes]
    opt = AdamW(params, lr=1e-3)
    assert opt._step_supports_amp_scaling and opt.defaults["fused"] is True
    # the table builder only touches the device for the final upload: feed it CPU stand-ins and inspect the chunk list
    class _P:   # minimal stand-in exposing what _group_tables reads
        def __init__(self, n): self._n = n; self.device = torch.device("cpu")
        def numel(self): return self._n
    tab = opt._group_tables((0, 0), opt.param_groups[0], [_P(n) for n in sizes])
    chunks = tab["chunks"].numpy()
    assert tab["n_chunks"] == chunks.shape[0] == sum((n + chunk - 1) // chunk for n in sizes)
    covered = [np.zeros(n, dtype=np.int32) for n in sizes]
    for t, c in chunks:
        lo, hi = c * chunk, min((c + 1) * chunk, sizes[t])
        assert lo < sizes[t]
        covered[t][lo:hi] += 1
    assert all((cv == 1).all() for cv in covered)
    # signature compatibility with torch.optim.AdamW and loud failures for what is not implemented
    import pytest
    with pytest.raises(NotImplementedError):
        AdamW(params, amsgrad=True)
    with pytest.raises(ValueError):
        AdamW(params, lr=-1.0)
    sd = opt.state_dict()
    assert {"lr", "betas", "eps", "weight_decay", "capturable"} <= set(sd["param_groups"][0])
