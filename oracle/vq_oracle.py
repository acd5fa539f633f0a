"""ORACLE — test infrastructure only.  numpy/ctypes front-end of oracle/vq_oracle.c (bit-exact fp32 VQ spec)."""
import ctypes

import numpy as np

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _layout(x, channels_first):
    """Returns (R, D, inner, elem_stride, outer_stride) for [.., D] rows or [b, c, h, w] (blocks.py) input."""
    if channels_first:
        b, c, h, w = x.shape
        return b * h * w, c, h * w, h * w, c * h * w
    D = x.shape[-1]
    return x.size // D, D, 1, 1, D


def vq_fwd(x, codebook, l2=True, gather_normalized=False, channels_first=False, want_dist=False):
    x = np.ascontiguousarray(x, dtype=np.float32)
    cb = np.ascontiguousarray(codebook, dtype=np.float32)
    R, D, inner, es, os_ = _layout(x, channels_first)
    idx = np.empty(R, dtype=np.int64)
    q = np.empty_like(x)
    s = np.zeros(1, dtype=np.float64)
    dist = np.empty(R, dtype=np.float32)
    flags = (1 if l2 else 0) | (2 if gather_normalized else 0)
    _load().vq_oracle_fwd(_p(x), _p(cb), ctypes.c_long(R), ctypes.c_int(D), ctypes.c_long(cb.shape[0]),
                          ctypes.c_long(inner), ctypes.c_long(es), ctypes.c_long(os_), ctypes.c_int(flags),
                          _p(idx), _p(q), _p(s), _p(dist))
    if channels_first:
        idx = idx.reshape(x.shape[0], x.shape[2], x.shape[3])
    else:
        idx = idx.reshape(x.shape[:-1])
    mse = s[0] / x.size
    return (q, idx, mse, dist) if want_dist else (q, idx, mse)


def vq_bwd(x, codebook, idx, g, a_commit, a_code, l2=True, gather_normalized=False, channels_first=False):
    x = np.ascontiguousarray(x, dtype=np.float32)
    cb = np.ascontiguousarray(codebook, dtype=np.float32)
    g = np.ascontiguousarray(g, dtype=np.float32)
    idx = np.ascontiguousarray(idx.reshape(-1), dtype=np.int64)
    R, D, inner, es, os_ = _layout(x, channels_first)
    dx = np.empty_like(x)
    dC = np.zeros_like(cb)
    flags = (1 if l2 else 0) | (2 if gather_normalized else 0)
    _load().vq_oracle_bwd(_p(x), _p(cb), _p(idx), _p(g), ctypes.c_long(R), ctypes.c_int(D),
                          ctypes.c_long(cb.shape[0]), ctypes.c_long(inner), ctypes.c_long(es), ctypes.c_long(os_),
                          ctypes.c_int(flags), ctypes.c_float(a_commit), ctypes.c_float(a_code), _p(dx), _p(dC))
    return dx, dC
