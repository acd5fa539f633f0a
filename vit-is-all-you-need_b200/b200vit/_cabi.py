"""ctypes binding of libb200vit.so (the C ABI in include/b200vit.h).

There is deliberately no fallback: if the shared library is missing or the device is not sm_100,
every op raises.  Tensors are passed as raw device pointers on torch's *current* stream.
"""
import ctypes
import os
import re
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200VIT_LIB") or os.path.join(_HERE, "libb200vit.so")  # B200VIT_LIB: experiment builds
HEADER_PATH = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "b200vit.h")

_lib = None
_lock = threading.Lock()
_inited_devices = set()

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int


class B200VitError(RuntimeError):
    pass


def declared_symbols(header_path: str = HEADER_PATH):
    """Names of every function include/b200vit.h declares (used by the export test)."""
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200vit_[a-z0-9_]+)\s*\(", text)))


_P, _I, _L, _F, _Z, _D = c_void_p, c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_size_t, ctypes.c_double

# argument types of every entry point in include/b200vit.h (ctypes would otherwise guess c_int for ints)
_SIGNATURES = {
    "b200vit_init": (_I, [_I]),
    "b200vit_version": (_I, []),
    "b200vit_debug_set": (_I, [_I, _I]),
    "b200vit_debug_max_clusters": (_I, []),
    "b200vit_debug_tmap_cache_stats": (_I, [_P]),
    "b200vit_gemm_bias": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "b200vit_gemm_bias_gelu": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "b200vit_gemm_bias_gelu_q8": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "b200vit_gemm_dgrad_dgelu_q8": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "b200vit_gelu_grad_code_lo": (_F, []),
    "b200vit_gelu_grad_code_step": (_F, []),
    "b200vit_gemm_bias_residual": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "b200vit_gemm_bias_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "b200vit_gemm_dgrad": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "b200vit_gemm_dgrad_dgelu": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "b200vit_gemm_wgrad": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "b200vit_gemm_wgrad_bias": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "b200vit_flash_attn_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "b200vit_flash_attn_bwd_workspace_size": (_Z, [_I, _I, _I]),
    "b200vit_flash_attn_fwd_dropout": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _F, ctypes.c_uint, _P]),
    "b200vit_flash_attn_bwd_dropout": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, ctypes.c_uint, _P, _Z, _P]),
    "b200vit_gemm_bias_dropout_residual": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _F, ctypes.c_uint, _P]),
    "b200vit_dropout_cast_bf16": (_I, [_P, _P, _L, _I, _F, ctypes.c_uint, _P]),
    "b200vit_dropout_mask_rows": (_I, [_P, _L, _I, _F, ctypes.c_uint, _P]),
    "b200vit_dropout_mask_attn": (_I, [_P, _I, _I, _I, _F, ctypes.c_uint, _P]),
    "b200vit_flash_attn_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "b200vit_layernorm_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P]),
    "b200vit_layernorm_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "b200vit_layernorm_bwd_xhat": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "b200vit_colsum_bf16": (_I, [_P, _P, _I, _I, _I, _P]),
    "b200vit_colsum_f32": (_I, [_P, _P, _I, _I, _P]),
    "b200vit_cast_f32_bf16": (_I, [_P, _P, _L, _P]),
    "b200vit_cast_bf16_f32": (_I, [_P, _P, _L, _F, _P]),
    "b200vit_patch_embed_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "b200vit_patch_embed_bwd_reduce": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "b200vit_tokens_assemble_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "b200vit_tokens_assemble_bwd_reduce": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "b200vit_gather_tokens_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "b200vit_affine_fold": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "b200vit_affine_unfold_workspace_size": (_Z, [_I]),
    "b200vit_affine_unfold_grads": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _Z, _P]),
    "b200vit_im2col_bf16": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "b200vit_col2im_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "b200vit_gather_tokens_bf16": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "b200vit_scatter_tokens": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    "b200vit_depatchify_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "b200vit_cross_entropy_fwd": (_I, [_P, _I, _L, _P, _P, _P, _P, _I, _I, _L, _P]),
    "b200vit_cross_entropy_bwd": (_I, [_P, _I, _L, _P, _P, _P, _P, _P, _L, _I, _I, _L, _P]),
    "b200vit_embed_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _I, _P]),
    "b200vit_embed_bwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "b200vit_attn_decode": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "b200vit_kv_append": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "b200vit_kv_fill": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "b200vit_advance_counter": (_I, [_P, _I, _P]),
    "b200vit_adamw_chunk_elems": (_I, []),
    "b200vit_adamw_step": (_I, [_P, _P, _I, _D, _D, _D, _D, _D, _L, _P, _P, _P, _P]),
    "b200vit_vq_workspace_size": (_Z, [_L, _I, _I]),
    "b200vit_vq_fwd": (_I, [_P, _P, _L, _I, _I, _L, _L, _L, _I, _F, _P, _P, _P, _P, _Z, _P]),
    "b200vit_vq_bwd": (_I, [_P, _P, _P, _P, _P, _L, _I, _I, _L, _L, _L, _I, _P, _P, _P]),
}


def _declare(lib):
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args


def load():
    """Loads the library once; raises loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise B200VitError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`."
                " There is no CPU / PyTorch fallback for the b200vit hot path.")
        lib = ctypes.CDLL(LIB_PATH)
        lib.b200vit_last_error.restype = ctypes.c_char_p
        _declare(lib)
        _lib = lib
    return _lib


def ensure_device(device_index: int):
    lib = load()
    if device_index not in _inited_devices:
        rc = lib.b200vit_init(c_int(device_index))
        if rc != 0:
            raise B200VitError(lib.b200vit_last_error().decode())
        # bring-up knobs for A/B measurements, e.g. B200VIT_DEBUG="10=1" (programmatic dependent launch off)
        for kv in os.environ.get("B200VIT_DEBUG", "").split(","):
            if "=" in kv:
                k, v = kv.split("=")
                lib.b200vit_debug_set(int(k), int(v))
        _inited_devices.add(device_index)
    return lib


def check(rc: int):
    if rc != 0:
        raise B200VitError(f"libb200vit: {load().b200vit_last_error().decode()} (rc={rc})")


def ptr(t):
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def lib_for(t: torch.Tensor):
    if not t.is_cuda:
        raise B200VitError("b200vit ops need CUDA tensors on a B200 (sm_100a); there is no CPU fallback")
    return ensure_device(t.device.index if t.device.index is not None else torch.cuda.current_device())


def call(name: str, *args):
    """Calls lib.<name>(*args) and raises on a non-zero return code."""
    fn = getattr(load(), name)
    check(fn(*args))
