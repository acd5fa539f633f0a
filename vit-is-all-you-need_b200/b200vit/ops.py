"""Thin tensor-level wrappers over the C ABI (include/b200vit.h).  Every function allocates its outputs with
torch (caching allocator), passes raw device pointers + torch's current stream, and raises on any error.
No op here has a PyTorch/CPU fallback: a missing library or a non-sm_100 device is an error."""
import os

import torch

from . import _cabi
from ._cabi import ptr

_STREAM = object()  # placeholder replaced by torch's current stream once the device check has passed


def stream_ptr():
    return _STREAM

BF16 = torch.bfloat16
F32 = torch.float32

launch_count = 0  # number of C-ABI compute calls issued (bench.py reports it as gpu_launches evidence)


# kernels launched per C-ABI call (memsets not counted); everything else launches exactly one kernel
_KERNELS_PER_CALL = {"b200vit_vq_fwd": 3, "b200vit_patch_embed_fwd": 3, "b200vit_cross_entropy_fwd": 2, "b200vit_embed_bwd": 2,
                     "b200vit_tokens_assemble_fwd": 2, "b200vit_affine_unfold_grads": 2}

_prof = None  # list of (start_event, end_event, flops) while profile_gemms() is active


def _call(name, t, *args, flops=None, hbm_bytes=None):
    global launch_count
    lib = _cabi.lib_for(t)  # raises for CPU tensors / missing library / non-sm_100 devices
    launch_count += _KERNELS_PER_CALL.get(name, 1)
    args = tuple(_cabi.stream_ptr() if a is _STREAM else a for a in args)
    work = flops if flops is not None else hbm_bytes
    if _prof is not None and work is not None and (flops is not None) == (_prof_kind == "flops"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _cabi.check(getattr(lib, name)(*args))
        e1.record()
        _prof.append((e0, e1, work, name))
        return
    _cabi.check(getattr(lib, name)(*args))


_prof_kind = "flops"


def profile_kernels(fn, steps=1, kind="flops"):
    """Runs fn() `steps` times with a CUDA-event pair (on the launching stream) around every C-ABI launch that
    declares its algorithmic work: kind="flops" -> the tcgen05 GEMMs (2MNK), kind="bytes" -> the HBM-bound kernels
    (LayerNorm: algorithmic bytes).  Returns (total ms, total work, launches, per-entry-point detail)."""
    global _prof, _prof_kind
    _prof, _prof_kind = [], kind
    try:
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        ms = sum(e0.elapsed_time(e1) for e0, e1, _, _ in _prof)
        work = float(sum(f for _, _, f, _ in _prof))
        n = len(_prof)
        detail = {}
        for e0, e1, f, name in _prof:
            d = detail.setdefault(name, [0.0, 0.0, 0])
            d[0] += e0.elapsed_time(e1); d[1] += f; d[2] += 1
        detail = {k: {"ms": v[0], "rate": v[1] / (v[0] / 1e3) if v[0] > 0 else 0.0, "launches": v[2]} for k, v in detail.items()}
    finally:
        _prof, _prof_kind = None, "flops"
    return ms, work, n, detail


def profile_gemms(fn, steps=1):
    """(total GEMM milliseconds, total algorithmic FLOPs, number of launches) -- see profile_kernels."""
    ms, fl, n, detail = profile_kernels(fn, steps, "flops")
    profile_gemms.last_detail = {k: {"ms": v["ms"], "tflops": v["rate"] / 1e12, "launches": v["launches"]} for k, v in detail.items()}
    return ms, fl, n


def _chk(t, dtype, name):
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return t


# ---------------------------------------------------------------- GEMMs
def gemm_bias(x, w, bias=None):
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty(M, N, device=x.device, dtype=BF16)
    _call("b200vit_gemm_bias", x, ptr(_chk(x, BF16, "x")), ptr(_chk(w, BF16, "w")), ptr(bias), ptr(y), M, N, K, stream_ptr(), flops=2.0 * M * N * K)
    return y


GELU_GRAD_Q8 = os.environ.get("B200VIT_GELU_GRAD_BF16", "0") != "1"   # B200VIT_GELU_GRAD_BF16=1: keep GELU' in bf16


def gemm_bias_gelu(x, w, bias=None, q8=None):
    """g = GELU(x w^T + bias) (bf16) and gprime = GELU'(x w^T + bias) -- what backward needs instead of the pre-activation.
    gprime is the 8-bit fixed-point code of csrc/gemm_tcgen05.cuh (uint8; absolute error <= 0.0025 on a value in
    [-0.13, 1.13]) whenever the 256-wide tile kernel applies (N a multiple of 256, more than 64 rows), bf16 otherwise."""
    M, K = x.shape
    N = w.shape[0]
    if q8 is None:
        q8 = GELU_GRAD_Q8 and N % 256 == 0 and M > 64
    g = torch.empty(M, N, device=x.device, dtype=BF16)
    if q8:
        gp = torch.empty(M, N, device=x.device, dtype=torch.uint8)
        _call("b200vit_gemm_bias_gelu_q8", x, ptr(_chk(x, BF16, "x")), ptr(_chk(w, BF16, "w")), ptr(bias), ptr(g), ptr(gp), M, N, K, stream_ptr(), flops=2.0 * M * N * K)
        return g, gp
    gp = torch.empty(M, N, device=x.device, dtype=BF16)
    _call("b200vit_gemm_bias_gelu", x, ptr(_chk(x, BF16, "x")), ptr(_chk(w, BF16, "w")), ptr(bias), ptr(g), ptr(gp), M, N, K, stream_ptr(), flops=2.0 * M * N * K)
    return g, gp


def gemm_bias_residual(x, w, bias, resid):
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=x.device, dtype=F32)
    _call("b200vit_gemm_bias_residual", x, ptr(_chk(x, BF16, "x")), ptr(_chk(w, BF16, "w")), ptr(bias), ptr(_chk(resid, F32, "resid")), ptr(out), M, N, K, stream_ptr(), flops=2.0 * M * N * K)
    return out


def gemm_bias_dropout_residual(x, w, bias, resid, p, seed):
    """out = resid + dropout_p(x w^T + bias) (fp32); the keep mask is a function of (seed, row, column)."""
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=x.device, dtype=F32)
    _call("b200vit_gemm_bias_dropout_residual", x, ptr(_chk(x, BF16, "x")), ptr(_chk(w, BF16, "w")), ptr(bias), ptr(_chk(resid, F32, "resid")), ptr(out),
          M, N, K, float(p), int(seed) & 0xFFFFFFFF, stream_ptr(), flops=2.0 * M * N * K)
    return out


def dropout_cast_bf16(x, p, seed):
    """bf16(x * keep / (1 - p)) for x [M, d] fp32: the backward of the dropout above, fused with the bf16 cast."""
    d = x.shape[-1]
    M = x.numel() // d
    out = torch.empty(x.shape, device=x.device, dtype=BF16)
    _call("b200vit_dropout_cast_bf16", x, ptr(_chk(x, F32, "x")), ptr(out), M, d, float(p), int(seed) & 0xFFFFFFFF, stream_ptr())
    return out


def dropout_mask_rows(M, d, p, seed, device):
    out = torch.empty(M, d, device=device, dtype=torch.uint8)
    _call("b200vit_dropout_mask_rows", out, ptr(out), M, d, float(p), int(seed) & 0xFFFFFFFF, stream_ptr())
    return out


def dropout_mask_attn(B, H, N, p, seed, device):
    out = torch.empty(B, H, N, N, device=device, dtype=torch.uint8)
    _call("b200vit_dropout_mask_attn", out, ptr(out), B, H, N, float(p), int(seed) & 0xFFFFFFFF, stream_ptr())
    return out


def gemm_bias_f32(x, w, bias=None):
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=x.device, dtype=F32)
    _call("b200vit_gemm_bias_f32", x, ptr(_chk(x, BF16, "x")), ptr(_chk(w, BF16, "w")), ptr(bias), ptr(out), M, N, K, stream_ptr(), flops=2.0 * M * N * K)
    return out


def gemm_dgrad(dy, w):
    M, N = dy.shape
    K = w.shape[1]
    dx = torch.empty(M, K, device=dy.device, dtype=BF16)
    _call("b200vit_gemm_dgrad", dy, ptr(_chk(dy, BF16, "dy")), ptr(_chk(w, BF16, "w")), ptr(dx), M, N, K, stream_ptr(), flops=2.0 * M * N * K)
    return dx


def gemm_dgrad_dgelu(dy, w, gprime):
    u = gprime
    M, N = dy.shape
    K = w.shape[1]
    dx = torch.empty(M, K, device=dy.device, dtype=BF16)
    if u.dtype == torch.uint8:      # the 8-bit code written by gemm_bias_gelu(q8=True)
        _call("b200vit_gemm_dgrad_dgelu_q8", dy, ptr(_chk(dy, BF16, "dy")), ptr(_chk(w, BF16, "w")), ptr(_chk(u, torch.uint8, "u")), ptr(dx), M, N, K, stream_ptr(), flops=2.0 * M * N * K)
        return dx
    _call("b200vit_gemm_dgrad_dgelu", dy, ptr(_chk(dy, BF16, "dy")), ptr(_chk(w, BF16, "w")), ptr(_chk(u, BF16, "u")), ptr(dx), M, N, K, stream_ptr(), flops=2.0 * M * N * K)
    return dx


def gemm_wgrad(dy, x, out=None, accumulate=False, bias_out=None, want_bias=False):
    """dw[N,K] = dy^T x (fp32).  With want_bias (or bias_out) also returns db[N] = column sums of dy, summed inside
    the same kernel from the shared-memory dy tiles (no second pass over dy)."""
    M, N = dy.shape
    K = x.shape[1]
    if out is None:
        out = torch.empty(N, K, device=dy.device, dtype=F32)
        accumulate = False
    elif tuple(out.shape) != (N, K):
        out = out.view(N, K)
    if bias_out is None and want_bias:
        bias_out = torch.empty(N, device=dy.device, dtype=F32)
    if bias_out is None:
        _call("b200vit_gemm_wgrad", dy, ptr(_chk(dy, BF16, "dy")), ptr(_chk(x, BF16, "x")), ptr(_chk(out, F32, "dw")), M, N, K, 1 if accumulate else 0, stream_ptr(), flops=2.0 * M * N * K)
        return out
    _call("b200vit_gemm_wgrad_bias", dy, ptr(_chk(dy, BF16, "dy")), ptr(_chk(x, BF16, "x")), ptr(_chk(out, F32, "dw")), ptr(_chk(bias_out, F32, "db")), M, N, K, 1 if accumulate else 0, stream_ptr(), flops=2.0 * M * N * K)
    return out, bias_out


# ---------------------------------------------------------------- attention
def flash_attn_fwd(qkv, B, N, H, causal=False, want_lse=True, seq_first=False, dropout_p=0.0, seed=0):
    d = H * 64
    assert qkv.numel() == B * N * 3 * d
    o = torch.empty((N, B, d) if seq_first else (B, N, d), device=qkv.device, dtype=BF16)
    lse = torch.empty(B, H, N, device=qkv.device, dtype=F32) if want_lse else None
    _call("b200vit_flash_attn_fwd_dropout", qkv, ptr(_chk(qkv, BF16, "qkv")), ptr(o), ptr(lse), B, N, H, 1 if causal else 0, 1 if seq_first else 0,
          float(dropout_p), int(seed) & 0xFFFFFFFF, stream_ptr())
    return o, lse


def flash_attn_bwd(qkv, o, d_o, lse, B, N, H, causal=False, seq_first=False, dropout_p=0.0, seed=0):
    d = H * 64
    dqkv = torch.empty((N, B, 3 * d) if seq_first else (B, N, 3 * d), device=qkv.device, dtype=BF16)
    ws_bytes = _cabi.lib_for(qkv).b200vit_flash_attn_bwd_workspace_size(B, N, H)
    ws = torch.empty(max(ws_bytes // 4, 1), device=qkv.device, dtype=F32)
    _call("b200vit_flash_attn_bwd_dropout", qkv, ptr(_chk(qkv, BF16, "qkv")), ptr(_chk(o, BF16, "o")), ptr(_chk(d_o, BF16, "d_o")), ptr(_chk(lse, F32, "lse")),
          ptr(dqkv), B, N, H, 1 if causal else 0, 1 if seq_first else 0, float(dropout_p), int(seed) & 0xFFFFFFFF,
          ptr(ws), ws.numel() * 4, stream_ptr())
    return dqkv


# ---------------------------------------------------------------- LayerNorm & small element-wise helpers
def layernorm_fwd(x, add=None, gamma=None, beta=None, want_x_out=False, out_bf16=True, out_f32=False, eps=1e-5):
    d = x.shape[-1]
    M = x.numel() // d
    y = torch.empty(x.shape, device=x.device, dtype=BF16) if out_bf16 else None
    y32 = torch.empty(x.shape, device=x.device, dtype=F32) if out_f32 else None
    mean = torch.empty(M, device=x.device, dtype=F32)
    rstd = torch.empty(M, device=x.device, dtype=F32)
    x_out = torch.empty_like(x) if want_x_out else None
    nbytes = M * d * (4 + (2 if add is not None else 0) + (4 if want_x_out else 0) + (2 if out_bf16 else 0) + (4 if out_f32 else 0)) + 8 * M
    _call("b200vit_layernorm_fwd", x, ptr(_chk(x, F32, "x")), ptr(add), ptr(x_out), ptr(gamma), ptr(beta), ptr(y), ptr(y32), ptr(mean), ptr(rstd), M, d, eps, stream_ptr(),
          hbm_bytes=float(nbytes))
    return y, y32, mean, rstd, x_out


def layernorm_bwd(dy, x, mean, rstd, gamma=None, dres=None, want_bf16=True, affine_grads=False):
    d = x.shape[-1]
    M = x.numel() // d
    dx = torch.empty(x.shape, device=x.device, dtype=F32)
    dxb = torch.empty(x.shape, device=x.device, dtype=BF16) if want_bf16 else None
    dg = torch.empty(d, device=x.device, dtype=F32) if affine_grads else None
    db = torch.empty(d, device=x.device, dtype=F32) if affine_grads else None
    dy16 = dy if dy.dtype == BF16 else None
    dy32 = dy if dy.dtype == F32 else None
    nbytes = M * d * ((2 if dy16 is not None else 4) + 4 + (4 if dres is not None else 0) + 4 + (2 if want_bf16 else 0)) + 8 * M
    _call("b200vit_layernorm_bwd", x, ptr(dy16), ptr(dy32), ptr(_chk(x, F32, "x")), ptr(mean), ptr(rstd), ptr(gamma), ptr(dres), ptr(dx), ptr(dxb), ptr(dg), ptr(db), M, d, stream_ptr(),
          hbm_bytes=float(nbytes))
    return dx, dxb, dg, db


def layernorm_bwd_xhat(dy, xhat, rstd, dres=None, want_bf16=True):
    """Backward of the affine-free LayerNorm from its saved bf16 output (xhat) and rstd; returns (dx fp32, dx bf16)."""
    d = xhat.shape[-1]
    M = xhat.numel() // d
    dx = torch.empty(xhat.shape, device=xhat.device, dtype=F32)
    dxb = torch.empty(xhat.shape, device=xhat.device, dtype=BF16) if want_bf16 else None
    nbytes = M * d * (2 + 2 + (4 if dres is not None else 0) + 4 + (2 if want_bf16 else 0)) + 4 * M
    _call("b200vit_layernorm_bwd_xhat", xhat, ptr(_chk(dy, BF16, "dy")), ptr(_chk(xhat, BF16, "xhat")), ptr(_chk(rstd, F32, "rstd")),
          ptr(dres), ptr(dx), ptr(dxb), M, d, stream_ptr(), hbm_bytes=float(nbytes))
    return dx, dxb


def colsum_bf16(a, out=None, accumulate=False):
    M, N = a.shape
    if out is None:
        out = torch.empty(N, device=a.device, dtype=F32)
        accumulate = False
    _call("b200vit_colsum_bf16", a, ptr(_chk(a, BF16, "a")), ptr(out), M, N, 1 if accumulate else 0, stream_ptr())
    return out


def colsum_f32(a):
    rows, n = a.shape
    out = torch.empty(n, device=a.device, dtype=F32)
    _call("b200vit_colsum_f32", a, ptr(_chk(a, F32, "a")), ptr(out), rows, n, stream_ptr())
    return out


def cast_bf16(t, out=None):
    t = _chk(t, F32, "t")
    if out is None:
        out = torch.empty(t.shape, device=t.device, dtype=BF16)
    _call("b200vit_cast_f32_bf16", t, ptr(t), ptr(out), t.numel(), stream_ptr())     # any length: vector body + scalar tail
    return out


def cast_f32_from_bf16(t, out, scale=1.0):
    """out (fp32, contiguous) <- scale * t (bf16, contiguous)."""
    _call("b200vit_cast_bf16_f32", t, ptr(_chk(t, BF16, "t")), ptr(_chk(out, F32, "out")), t.numel(), float(scale), stream_ptr())
    return out


# ---------------------------------------------------------------- KV-cached decode
def kv_cache_alloc(B, H, n_max, device):
    """K / V planes [2, B, H, n_max, 64] bf16: one head's rows are contiguous (what the decode kernel streams)."""
    return torch.empty(2, B, H, n_max, 64, device=device, dtype=BF16)


def kv_fill(qkv, kv_cache, B, S):
    """kv_cache[:, :, :, :S] <- K / V of qkv [B*S, 3*H*64] (the fused QKV projection of the prompt)."""
    _, Bc, H, n_max, _ = kv_cache.shape
    _call("b200vit_kv_fill", qkv, ptr(_chk(qkv, BF16, "qkv")), ptr(_chk(kv_cache, BF16, "kv_cache")), B, S, n_max, H, stream_ptr())


def kv_append(qkv_rows, kv_cache, pos_dev):
    """kv_cache[:, :, :, *pos_dev] <- K / V of qkv_rows [B, 3*H*64] (the fused QKV projection of the new token)."""
    _, B, H, n_max, _ = kv_cache.shape
    if qkv_rows.numel() != B * 3 * H * 64:
        raise ValueError("kv_append: qkv_rows must be [B, 3*H*64]")
    _call("b200vit_kv_append", qkv_rows, ptr(_chk(qkv_rows, BF16, "qkv_rows")), ptr(_chk(kv_cache, BF16, "kv_cache")), B, n_max, H,
          ptr(_chk(pos_dev, torch.int32, "pos")), stream_ptr())


def attn_decode(qkv_rows, kv_cache, pos_dev):
    """One query per (batch, head) -- slot 0 of qkv_rows [B, 3*H*64] -- against the cached rows 0..*pos_dev -> o [B, H*64] bf16.
    pos_dev: int32 device tensor with one element (the position is read on the device: graph-capturable)."""
    two, B, H, n_max, hd = kv_cache.shape
    if two != 2 or hd != 64:
        raise ValueError("attn_decode: cache must be [2, B, H, Nmax, 64]")
    out = torch.empty(B, H * 64, device=kv_cache.device, dtype=BF16)
    _call("b200vit_attn_decode", kv_cache, ptr(_chk(qkv_rows, BF16, "qkv_rows")), ptr(_chk(kv_cache, BF16, "kv_cache")), ptr(out), B, n_max, H,
          ptr(_chk(pos_dev, torch.int32, "pos")), stream_ptr(), hbm_bytes=float(B * H * n_max * 256))
    return out


def advance_counter(counter, by=1):
    _call("b200vit_advance_counter", counter, ptr(_chk(counter, torch.int32, "counter")), by, stream_ptr())


# ---------------------------------------------------------------- classifier head + cross-entropy
def gather_tokens_bf16(x, t0=0, cnt=1):
    """x [B, N, d] fp32 -> bf16 [B*cnt, d] rows of tokens t0 .. t0+cnt-1 (operand of the head / proj / de-patchify GEMMs)."""
    B, N, d = x.shape
    out = torch.empty(B * cnt, d, device=x.device, dtype=BF16)
    _call("b200vit_gather_tokens_bf16", x, ptr(_chk(x, F32, "x")), ptr(out), B, N, d, t0, cnt, stream_ptr())
    return out


def scatter_tokens(dy, B, N, t0=0, cnt=1, want_bf16=True):
    """Gradient of x[:, t0:t0+cnt]: fp32 [B, N, d] that is zero outside those tokens, plus its bf16 twin."""
    d = dy.shape[-1]
    if dy.dtype not in (BF16, F32) or not dy.is_contiguous() or dy.numel() != B * cnt * d:
        raise TypeError("scatter_tokens: dy must be a contiguous bf16 or fp32 tensor of B*cnt rows")
    dx = torch.empty(B, N, d, device=dy.device, dtype=F32)
    dxb = torch.empty(B, N, d, device=dy.device, dtype=BF16) if want_bf16 else None
    _call("b200vit_scatter_tokens", dy, ptr(dy), 1 if dy.dtype == BF16 else 0, ptr(dx), ptr(dxb), B, N, d, t0, cnt, stream_ptr(),
          hbm_bytes=float(B * N * d * (4 + (2 if want_bf16 else 0))))
    return dx, dxb


def depatchify_fwd(rows, w_cmajor, bias_cmajor, B, Ht, Wt, p, C):
    """rows bf16 [B*Ht*Wt, d] x w_cmajor bf16 [C*p*p, d] (+ bias) -> image fp32 [B, C, Ht*p, Wt*p] (pixel shuffle in the epilogue)."""
    d = rows.shape[1]
    img = torch.empty(B, C, Ht * p, Wt * p, device=rows.device, dtype=F32)
    _call("b200vit_depatchify_fwd", rows, ptr(_chk(rows, BF16, "rows")), ptr(_chk(w_cmajor, BF16, "w")), ptr(bias_cmajor), ptr(img),
          B, Ht, Wt, p, C, d, stream_ptr(), flops=2.0 * rows.shape[0] * C * p * p * d)
    return img


def embed_fwd(idx, tok_embed, pos_embed, pos0=0, pos0_dev=None):
    """idx [B, S] int64 -> fp32 [B, S, d] = tok_embed[idx] + pos_embed[pos0 : pos0 + S] (train_videogpt.py:50).
    pos0_dev (int32 device tensor) overrides pos0 with a position read on the device (graph-captured decode)."""
    B, S = idx.shape
    V, d = tok_embed.shape
    if pos0_dev is None and pos0 + S > pos_embed.shape[0]:
        raise ValueError(f"embed_fwd: positions {pos0}..{pos0 + S - 1} exceed pos_embed ({pos_embed.shape[0]} rows)")
    out = torch.empty(B, S, d, device=tok_embed.device, dtype=F32)
    _call("b200vit_embed_fwd", tok_embed, ptr(_chk(idx, torch.int64, "idx")), ptr(_chk(tok_embed, F32, "tok_embed")),
          ptr(_chk(pos_embed, F32, "pos_embed")), ptr(out), B, S, d, pos0, ptr(pos0_dev), pos_embed.shape[0], V, stream_ptr())
    return out


def embed_bwd(idx, dy, vocab, n_pos):
    """(d tok_embed [vocab, d], d pos_embed [n_pos, d]) for embed_fwd with pos0 = 0."""
    B, S, d = dy.shape
    dtok = torch.empty(vocab, d, device=dy.device, dtype=F32)
    dpos = torch.zeros(n_pos, d, device=dy.device, dtype=F32) if n_pos != S else torch.empty(S, d, device=dy.device, dtype=F32)
    _call("b200vit_embed_bwd", dy, ptr(_chk(idx, torch.int64, "idx")), ptr(_chk(dy, F32, "dy")), ptr(dtok), ptr(dpos), B, S, d, vocab,
          stream_ptr())
    return dtok, dpos


def cross_entropy_fwd(logits, labels, ignore_index=-100):
    """logits [R, C] (bf16 or fp32, last dim contiguous), labels [R] int64 -> (loss2 = [mean loss, 1/n_valid], lse [R])."""
    R, C = logits.shape
    if logits.dtype not in (BF16, F32) or logits.stride(1) != 1:
        raise TypeError("cross_entropy_fwd: logits must be bf16 or fp32 with a contiguous class dimension")
    labels = _chk(labels, torch.int64, "labels")
    loss = torch.empty(2, device=logits.device, dtype=F32)
    lse = torch.empty(R, device=logits.device, dtype=F32)
    scratch = torch.empty(R, device=logits.device, dtype=F32)
    _call("b200vit_cross_entropy_fwd", logits, ptr(logits), 1 if logits.dtype == BF16 else 0, logits.stride(0), ptr(labels),
          ptr(loss), ptr(lse), ptr(scratch), R, C, ignore_index, stream_ptr())
    return loss, lse


def cross_entropy_bwd(logits, labels, lse, loss, dloss, ignore_index=-100):
    R, C = logits.shape
    dlogits = torch.empty(R, C, device=logits.device, dtype=logits.dtype)
    _call("b200vit_cross_entropy_bwd", logits, ptr(logits), 1 if logits.dtype == BF16 else 0, logits.stride(0), ptr(labels),
          ptr(lse), ptr(loss), ptr(_chk(dloss, F32, "dloss")), ptr(dlogits), C, R, C, ignore_index, stream_ptr())
    return dlogits


# ---------------------------------------------------------------- patch embedding
def patch_embed_fwd(x, w_bf16, bias, pos_emb, extra_emb, p):
    B, C, H, W = x.shape
    d = w_bf16.shape[0]
    P = (H // p) * (W // p)
    extra = 0 if extra_emb is None else extra_emb.shape[0]
    tokens = torch.empty(B, P + extra, d, device=x.device, dtype=F32)
    cols = torch.empty(B * P, C * p * p, device=x.device, dtype=BF16)
    _call("b200vit_patch_embed_fwd", x, ptr(_chk(x, F32, "x")), ptr(_chk(w_bf16.view(d, -1), BF16, "w")), ptr(bias), ptr(_chk(pos_emb, F32, "pos_emb")),
          ptr(extra_emb), ptr(tokens), ptr(cols), B, C, H, W, p, d, extra, stream_ptr())
    return tokens, cols


def patch_embed_bwd_reduce(dtokens, extra):
    B, T, d = dtokens.shape
    dsum = torch.empty(T, d, device=dtokens.device, dtype=F32)
    dpe = torch.empty(B * (T - extra), d, device=dtokens.device, dtype=BF16)
    _call("b200vit_patch_embed_bwd_reduce", dtokens, ptr(_chk(dtokens, F32, "dtokens")), ptr(dsum), ptr(dpe), B, T, extra, d, stream_ptr())
    return dsum, dpe


def tokens_assemble_fwd(cols, w_bf16, bias, pos, head0, head1, head_pos, tail_a, tail_b, B, P, extra, tail):
    """tokens[B, extra + P + tail, d] fp32 of blocks.TiTokEncoder / TiTokDecoder (blocks.py:254-268, 337-352): GEMM rows
    cols[B*P, K] x w[d, K]^T + bias + pos[P, d] at rows [extra, extra+P), broadcast head / tail rows around them."""
    K = cols.shape[1]
    d = w_bf16.shape[0]
    tokens = torch.empty(B, extra + P + tail, d, device=cols.device, dtype=F32)
    _call("b200vit_tokens_assemble_fwd", cols, ptr(_chk(cols, BF16, "cols")), ptr(_chk(w_bf16, BF16, "w")), ptr(bias), ptr(_chk(pos, F32, "pos")),
          ptr(head0), ptr(head1), ptr(head_pos), ptr(tail_a), ptr(tail_b), ptr(tokens), B, P, K, d, extra, tail, stream_ptr(),
          flops=2.0 * B * P * K * d)
    return tokens


def tokens_assemble_bwd_reduce(dtokens, extra, tail):
    B, T, d = dtokens.shape
    P = T - extra - tail
    dsum = torch.empty(T, d, device=dtokens.device, dtype=F32)
    dpe = torch.empty(B * P, d, device=dtokens.device, dtype=BF16)
    _call("b200vit_tokens_assemble_bwd_reduce", dtokens, ptr(_chk(dtokens, F32, "dtokens")), ptr(dsum), ptr(dpe), B, T, extra, tail, d,
          stream_ptr())
    return dsum, dpe


def gather_tokens_f32(x, t0=0, cnt=1):
    """x [B, N, d] fp32 -> fp32 [B, cnt, d] copy of tokens t0 .. t0+cnt-1."""
    B, N, d = x.shape
    out = torch.empty(B, cnt, d, device=x.device, dtype=F32)
    _call("b200vit_gather_tokens_f32", x, ptr(_chk(x, F32, "x")), ptr(out), B, N, d, t0, cnt, stream_ptr())
    return out


def affine_fold(W, bias, gamma, beta, out=None):
    """(bf16(W * gamma) [N, K], bias + W beta [N]): the affine LayerNorm in front of a Linear folded into its weights."""
    N, K = W.shape
    if out is None:
        out = (torch.empty(N, K, device=W.device, dtype=BF16), torch.empty(N, device=W.device, dtype=F32))
    _call("b200vit_affine_fold", W, ptr(_chk(W, F32, "W")), ptr(bias), ptr(_chk(gamma, F32, "gamma")), ptr(_chk(beta, F32, "beta")),
          ptr(out[0]), ptr(out[1]), N, K, stream_ptr())
    return out


def affine_unfold_grads(dW, W, gamma, beta, dbias, dgamma=None, dbeta=None, accumulate=False):
    """In place dW <- dW * gamma + dbias (x) beta (dW, dbias = wgrad / bias gradient on xhat); returns (dgamma, dbeta)."""
    N, K = W.shape
    if dgamma is None:
        dgamma = torch.empty(K, device=W.device, dtype=F32)
        dbeta = torch.empty(K, device=W.device, dtype=F32)
        accumulate = False
    lib = _cabi.lib_for(W)
    ws_bytes = lib.b200vit_affine_unfold_workspace_size(K)
    ws = torch.empty(ws_bytes // 4, device=W.device, dtype=F32)
    _call("b200vit_affine_unfold_grads", W, ptr(_chk(dW.view(N, K), F32, "dW")), ptr(_chk(W, F32, "W")), ptr(_chk(gamma, F32, "gamma")),
          ptr(_chk(beta, F32, "beta")), ptr(_chk(dbias, F32, "dbias")), ptr(dgamma), ptr(dbeta), N, K, 1 if accumulate else 0, ptr(ws), ws_bytes, stream_ptr())
    return dgamma, dbeta


def im2col_bf16(x, p):
    """x [B, C, H, W] fp32 -> bf16 patch rows [B * (H/p) * (W/p), C * p * p] in (c, i, j) order (the conv weight's)."""
    B, C, H, W = x.shape
    cols = torch.empty(B * (H // p) * (W // p), C * p * p, device=x.device, dtype=BF16)
    _call("b200vit_im2col_bf16", x, ptr(_chk(x, F32, "x")), ptr(cols), B, C, H, W, p, stream_ptr())
    return cols


def col2im(dcols, B, C, H, W, p):
    dx = torch.empty(B, C, H, W, device=dcols.device, dtype=F32)
    _call("b200vit_col2im_f32", dcols, ptr(_chk(dcols, BF16, "dcols")), ptr(dx), B, C, H, W, p, stream_ptr())
    return dx


# ---------------------------------------------------------------- VQ
def _vq_layout(x, channels_first):
    if channels_first:
        b, c, h, w = x.shape
        return b * h * w, c, h * w, h * w, c * h * w
    D = x.shape[-1]
    return x.numel() // D, D, 1, 1, D


def vq_fwd(x, codebook, l2=True, gather_normalized=False, channels_first=False, commitment_cost=0.25):
    x = _chk(x, F32, "x")
    codebook = _chk(codebook, F32, "codebook")
    R, D, inner, es, os_ = _vq_layout(x, channels_first)
    K = codebook.shape[0]
    lib = _cabi.lib_for(x)
    ws_bytes = lib.b200vit_vq_workspace_size(R, D, K)
    ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
    idx = torch.empty(R, device=x.device, dtype=torch.int64)
    q = torch.empty_like(x)
    losses = torch.empty(3, device=x.device, dtype=F32)
    flags = (1 if l2 else 0) | (2 if gather_normalized else 0)
    _call("b200vit_vq_fwd", x, ptr(x), ptr(codebook), R, D, K, inner, es, os_, flags, commitment_cost, ptr(idx), ptr(q), ptr(losses), ptr(ws), ws_bytes, stream_ptr())
    return q, idx, losses


def vq_bwd(x, codebook, idx, grad_q, coef, l2=True, gather_normalized=False, channels_first=False):
    R, D, inner, es, os_ = _vq_layout(x, channels_first)
    K = codebook.shape[0]
    dx = torch.empty_like(x)
    dC = torch.empty_like(codebook)
    flags = (1 if l2 else 0) | (2 if gather_normalized else 0)
    _call("b200vit_vq_bwd", x, ptr(x), ptr(codebook), ptr(idx), ptr(_chk(grad_q, F32, "grad_q")), ptr(_chk(coef, F32, "coef")), R, D, K, inner, es, os_, flags, ptr(dx), ptr(dC), stream_ptr())
    return dx, dC
