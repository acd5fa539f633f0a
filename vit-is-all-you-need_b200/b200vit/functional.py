"""torch.autograd.Functions that string the C-ABI kernels into the reference's blocks.

The save-for-backward policy lives here.  Residual stream and parameters are fp32 (as in the reference under
autocast); GEMM operands, attention I/O and saved activations are bf16; weight gradients are
produced in fp32 by the wgrad GEMMs.
"""
import torch

from . import ops

BF16 = torch.bfloat16
F32 = torch.float32


# ------------------------------------------------------------------------------------------------------------
# bf16 operand cache for fp32 parameters (refreshed when the optimiser has touched the parameter)
# ------------------------------------------------------------------------------------------------------------
# torch.optim.*(fused=True) updates parameters through torch._fused_*_ ops that do NOT bump Tensor._version (checked on
# torch 2.11: AdamW(fused=True).step() leaves p._version unchanged, foreach / single-tensor add 2), so the version
# counter alone would leave the bf16 operands stale for ever.  A global optimizer-step post-hook therefore advances an
# epoch that is part of the cache key: after ANY optimizer step every cached operand is re-cast on next use.
# b200vit.optim.AdamW writes the refreshed bf16 operands itself and is exempt (it re-keys the caches it has updated).
_WEIGHT_EPOCH = 0
_STEP_EPOCH = 0     # advances on EVERY optimizer step: key of the LayerNorm-folded operands (folded_of), which no optimizer refreshes


def _optimizer_stepped(optimizer, *_args, **_kwargs):
    global _WEIGHT_EPOCH, _STEP_EPOCH
    _STEP_EPOCH += 1
    if not getattr(optimizer, "_b200_refreshes_bf16", False):
        _WEIGHT_EPOCH += 1


from torch.optim.optimizer import register_optimizer_step_post_hook as _register_step_post_hook  # noqa: E402

_register_step_post_hook(_optimizer_stepped)


def bf16_key(p: torch.Tensor):
    return (_WEIGHT_EPOCH, p._version, p.data_ptr())


def bf16_of(p: torch.Tensor) -> torch.Tensor:
    key = bf16_key(p)
    cached = getattr(p, "_b200_bf16", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    src = p.detach()
    if src.dtype != F32:
        src = src.float()
    src = src.contiguous()
    if cached is not None and cached[1].shape == src.shape and not torch.cuda.is_current_stream_capturing():
        t = ops.cast_bf16(src, out=cached[1])   # refresh in place: the operand keeps its address from step to step
    else:
        t = ops.cast_bf16(src)
    try:
        p._b200_bf16 = (key, t)
    except Exception:  # pragma: no cover  (tensors that refuse attributes)
        pass
    return t


def invalidate_weight_cache(module_or_params=None):
    """Drops the cached bf16 GEMM operands.  Needed after parameter writes that neither bump Tensor._version nor go through
    an optimizer step: `p.data.copy_(...)`, `dist.broadcast(p.data)`, EMA updates through `.data`.  With no argument every
    cache in the process is invalidated (the epoch in the cache key advances)."""
    global _WEIGHT_EPOCH, _STEP_EPOCH
    if module_or_params is None:
        _WEIGHT_EPOCH += 1
        _STEP_EPOCH += 1
        return
    params = module_or_params.parameters() if hasattr(module_or_params, "parameters") else module_or_params
    for p in params:
        for attr in ("_b200_bf16", "_b200_folded", "_b200_depatch"):
            if hasattr(p, attr):
                delattr(p, attr)


def folded_of(W, bias, gamma, beta):
    """(bf16(W diag(gamma)), bias + W beta): the operands of a Linear with the affine LayerNorm in front of it folded in
    (csrc/affine_fold.cu), cached on W until any of the four parameters changes or an optimizer steps."""
    key = (_STEP_EPOCH, W._version, W.data_ptr(), gamma._version, gamma.data_ptr(), beta._version, beta.data_ptr(),
           None if bias is None else (bias._version, bias.data_ptr()))
    cached = getattr(W, "_b200_folded", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    out = cached[1] if (cached is not None and not torch.cuda.is_current_stream_capturing()) else None   # refresh in place
    out = ops.affine_fold(_f32c(W), _f32c(bias), _f32c(gamma), _f32c(beta), out=out)
    try:
        W._b200_folded = (key, out)
    except Exception:  # pragma: no cover
        pass
    return out


# ------------------------------------------------------------------------------------------------------------
# Gradient sink (set by ddp.DataParallel): lets the wgrad kernels write straight into all-reduce buckets and
# start a bucket's all-reduce while the rest of the fused backward is still running.
# ------------------------------------------------------------------------------------------------------------
_GRAD_SINK = None   # the sink of the DataParallel wrapper whose forward is running (captured per autograd node)


def set_grad_sink(sink):
    """Returns the previous sink so that the caller can restore it."""
    global _GRAD_SINK
    prev, _GRAD_SINK = _GRAD_SINK, sink
    return prev


def _slot(sink, param):
    return None if sink is None else sink.grad_slot(param)


def _ready(sink, *params):
    if sink is not None:
        for p in params:
            sink.grad_ready(p)


# ------------------------------------------------------------------------------------------------------------
# Dropout seeds: every dropout site of every forward call draws a fresh 32-bit seed; the masks themselves are pure
# functions of (seed, coordinates) inside the kernels (csrc/dropout.cuh), so backward only needs the seed.
# ------------------------------------------------------------------------------------------------------------
_drop_calls = 0


def next_dropout_seed() -> int:
    global _drop_calls
    _drop_calls += 1
    rank = int(__import__("os").environ.get("RANK", "0"))
    x = (torch.initial_seed() * 0x9E3779B97F4A7C15 + _drop_calls * 0xBF58476D1CE4E5B9 + rank * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 31
    return (x * 0xD6E8FEB86659FD93 >> 32) & 0xFFFFFFFF


def _f32c(t):
    if t is None:
        return None
    t = t.detach()
    if t.dtype != F32:
        t = t.float()
    return t.contiguous()


def _as_rows_f32(x):
    x = x.detach()
    if x.dtype != F32:
        x = x.float()
    return x.contiguous()


# ------------------------------------------------------------------------------------------------------------
# transformer.TransformerLayer (transformer.py:31-45): x + attn(LN(x)); x + mlp(LN(x)); affine-free LN, no out-proj
# ------------------------------------------------------------------------------------------------------------
class LayerParams:
    __slots__ = ("qkv_w", "qkv_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")

    def __init__(self, qkv_w, qkv_b, fc1_w, fc1_b, fc2_w, fc2_b):
        self.qkv_w, self.qkv_b, self.fc1_w, self.fc1_b, self.fc2_w, self.fc2_b = qkv_w, qkv_b, fc1_w, fc1_b, fc2_w, fc2_b


def layer_forward(x0, P: LayerParams, B, N, H, causal, save, dropout=(0.0, 0.0)):
    """x0: [B*N, d] fp32 contiguous.  Returns x2 [B*N, d] fp32 and (when save) the tensors backward needs.
    dropout = (p_attn, p_mlp): dropout_p on the attention probabilities (transformer.py:28, active in eval mode
    too, as in the reference) and nn.Dropout after mlp[2] (transformer.py:40, training mode only), each with its
    own fresh seed."""
    p_attn, p_mlp = dropout
    seeds = (next_dropout_seed() if p_attn > 0.0 else 0, next_dropout_seed() if p_mlp > 0.0 else 0)
    a, _, mean1, rstd1, _ = ops.layernorm_fwd(x0)
    qkv = ops.gemm_bias(a, bf16_of(P.qkv_w), _f32c(P.qkv_b))
    o, lse = ops.flash_attn_fwd(qkv, B, N, H, causal, want_lse=save, dropout_p=p_attn, seed=seeds[0])
    b, _, mean2, rstd2, x1 = ops.layernorm_fwd(x0, add=o.view(B * N, -1), want_x_out=True)
    g, u = ops.gemm_bias_gelu(b, bf16_of(P.fc1_w), _f32c(P.fc1_b))  # u := GELU'(pre-activation)
    if p_mlp > 0.0:
        x2 = ops.gemm_bias_dropout_residual(g, bf16_of(P.fc2_w), _f32c(P.fc2_b), x1, p_mlp, seeds[1])
    else:
        x2 = ops.gemm_bias_residual(g, bf16_of(P.fc2_w), _f32c(P.fc2_b), x1)
    # backward rebuilds LN' from the saved bf16 outputs a / b (= x-hat: the LN is affine-free) and rstd, so the fp32
    # residual rows x0 / x1 and the means are not kept
    saved = (rstd1, a, qkv, o, lse, rstd2, b, u, g, seeds) if save else None
    return x2, saved


def layer_backward(dx2, dx2_bf16, saved, P: LayerParams, B, N, H, causal, need_dx=True, dropout=(0.0, 0.0), sink=None):
    """dx2: [B*N, d] fp32 (dx2_bf16: optional bf16 copy).  Returns (dx0, dx0_bf16, grads in LayerParams order)."""
    rstd1, a, qkv, o, lse, rstd2, b, u, g, seeds = saved
    p_attn, p_mlp = dropout
    # all-reduce bucket slots, asked for ONCE per parameter: the sink answers None when the gradient has to go through
    # autograd instead (no sink; p.grad already holds micro-batch gradients from no_sync() steps; parameter delivered before)
    params = (P.qkv_w, P.qkv_b, P.fc1_w, P.fc1_b, P.fc2_w, P.fc2_b)
    s_qkv_w, s_qkv_b, s_fc1_w, s_fc1_b, s_fc2_w, s_fc2_b = slots = tuple(_slot(sink, q) for q in params)

    def delivered(*pairs):      # gradients that were written straight into their slots are complete: start their bucket
        _ready(sink, *(q for q, s_ in pairs if s_ is not None))

    if p_mlp > 0.0:
        dv = ops.dropout_cast_bf16(dx2, p_mlp, seeds[1])   # gradient through nn.Dropout, same mask as forward
    else:
        dv = dx2_bf16 if dx2_bf16 is not None else ops.cast_bf16(dx2)
    d_fc2_w, d_fc2_b = ops.gemm_wgrad(dv, g, out=s_fc2_w, bias_out=s_fc2_b, want_bias=True)
    delivered((P.fc2_w, s_fc2_w), (P.fc2_b, s_fc2_b))
    du = ops.gemm_dgrad_dgelu(dv, bf16_of(P.fc2_w), u)
    d_fc1_w, d_fc1_b = ops.gemm_wgrad(du, b, out=s_fc1_w, bias_out=s_fc1_b, want_bias=True)
    delivered((P.fc1_w, s_fc1_w), (P.fc1_b, s_fc1_b))
    db = ops.gemm_dgrad(du, bf16_of(P.fc1_w))
    dx1, dx1_bf16 = ops.layernorm_bwd_xhat(db, b, rstd2, dres=dx2, want_bf16=True)
    dqkv = ops.flash_attn_bwd(qkv, o, dx1_bf16.view(B, N, -1), lse, B, N, H, causal, dropout_p=p_attn, seed=seeds[0]).view(B * N, -1)
    d_qkv_w, d_qkv_b = ops.gemm_wgrad(dqkv, a, out=s_qkv_w, bias_out=s_qkv_b, want_bias=True)
    delivered((P.qkv_w, s_qkv_w), (P.qkv_b, s_qkv_b))
    dx0 = dx0_bf16 = None
    if need_dx:
        da = ops.gemm_dgrad(dqkv, bf16_of(P.qkv_w))
        dx0, dx0_bf16 = ops.layernorm_bwd_xhat(da, a, rstd1, dres=dx1, want_bf16=True)
    grads = (d_qkv_w, d_qkv_b, d_fc1_w, d_fc1_b, d_fc2_w, d_fc2_b)
    # gradients that already sit in their bucket slots have been marked ready: hand autograd None for those so that
    # AccumulateGrad does not clone them into a second buffer (the sink points param.grad at the slots)
    grads = tuple(None if s_ is not None else g_ for s_, g_ in zip(slots, grads))
    return dx0, dx0_bf16, grads


class TransformerStackFn(torch.autograd.Function):
    """A stack of transformer.TransformerLayer (1 layer == TransformerLayer.forward, n == Transformer.forward,
    transformer.py:52-54).  Keeping the stack in one node lets backward hand the bf16 copy of the residual
    gradient from layer to layer without an extra cast pass."""

    @staticmethod
    def forward(ctx, x, n_heads, causal, dropout, *params):  # dropout = (p_attention, p_mlp)
        B, N, d = x.shape
        n_layers = len(params) // 6
        layers = [LayerParams(*params[6 * i:6 * i + 6]) for i in range(n_layers)]
        need_grad = any(ctx.needs_input_grad)  # grad mode is off inside forward(); this reflects the caller's
        h = _as_rows_f32(x).view(B * N, d)
        saved_all = []
        for P in layers:
            h, saved = layer_forward(h, P, B, N, n_heads, causal, need_grad, dropout)
            saved_all.append(saved)
        ctx.layers = layers
        ctx.saved_all = saved_all
        ctx.dims = (B, N, d, n_heads, causal)
        ctx.dropout = dropout
        ctx.sink = _GRAD_SINK   # the data-parallel wrapper (if any) this forward ran under
        ctx.x_needs_grad = x.requires_grad
        return h.view(B, N, d)

    @staticmethod
    def backward(ctx, dy):
        B, N, d, H, causal = ctx.dims
        dx = _as_rows_f32(dy).view(B * N, d)
        twin = getattr(dy, "_b200_bf16_twin", None)   # TokenLinearFn / DepatchifyFn backward hand the bf16 copy along
        dx_bf16 = None
        if twin is not None and twin[1] == dy._version and dy.dtype == F32 and dy.is_contiguous() and twin[0].shape == dy.shape:
            dx_bf16 = twin[0].view(B * N, d)
        grads = []
        n = len(ctx.layers)
        for i in range(n - 1, -1, -1):
            need_dx = i > 0 or ctx.x_needs_grad
            dx, dx_bf16, g = layer_backward(dx, dx_bf16, ctx.saved_all[i], ctx.layers[i], B, N, H, causal, need_dx, ctx.dropout, ctx.sink)
            ctx.saved_all[i] = None  # free activations as we go
            grads.append(g)
        flat = []
        for g in reversed(grads):
            flat.extend(g)
        ctx.saved_all = None
        return (dx.view(B, N, d) if dx is not None else None, None, None, None, *flat)


class AttentionFn(torch.autograd.Function):
    """transformer.Attention.forward on its own (transformer.py:26-29): qkv Linear + SDPA, no out-proj."""

    @staticmethod
    def forward(ctx, x, qkv_w, qkv_b, n_heads, causal, dropout=0.0):
        B, N, d = x.shape
        seed = next_dropout_seed() if dropout > 0.0 else 0
        a = ops.cast_bf16(_as_rows_f32(x).view(B * N, d))
        qkv = ops.gemm_bias(a, bf16_of(qkv_w), _f32c(qkv_b))
        o, lse = ops.flash_attn_fwd(qkv, B, N, n_heads, causal, dropout_p=dropout, seed=seed)
        ctx.saved = (a, qkv, o, lse, qkv_w)
        ctx.dims = (B, N, d, n_heads, causal)
        ctx.drop = (dropout, seed)
        return o.float()

    @staticmethod
    def backward(ctx, do):
        a, qkv, o, lse, qkv_w = ctx.saved
        B, N, d, H, causal = ctx.dims
        do16 = ops.cast_bf16(_as_rows_f32(do).view(B * N, d)).view(B, N, d)
        dqkv = ops.flash_attn_bwd(qkv, o, do16, lse, B, N, H, causal, dropout_p=ctx.drop[0], seed=ctx.drop[1]).view(B * N, -1)
        dw, db = ops.gemm_wgrad(dqkv, a, want_bias=True)
        dx = ops.gemm_dgrad(dqkv, bf16_of(qkv_w)).float().view(B, N, d)
        return dx, dw, db, None, None, None


# ------------------------------------------------------------------------------------------------------------
# Patch embedding of ViT (train_vit.py:34-36,38-45; blocks.py:235-237,257-267)
# ------------------------------------------------------------------------------------------------------------
class PatchEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, conv_w, conv_b, pos_emb, extra_emb, patch):
        B, C, H, W = x.shape
        d = conv_w.shape[0]
        extra = 0 if extra_emb is None else extra_emb.shape[0]
        w16 = bf16_of(conv_w).view(d, -1)
        tokens, cols = ops.patch_embed_fwd(_as_rows_f32(x), w16, _f32c(conv_b), _f32c(pos_emb),
                                           _f32c(extra_emb) if extra > 0 else None, patch)
        ctx.saved = (cols, conv_w)
        ctx.dims = (B, C, H, W, patch, d, extra)
        ctx.x_needs_grad = x.requires_grad
        ctx.has_bias = conv_b is not None
        ctx.has_extra = extra_emb is not None
        return tokens

    @staticmethod
    def backward(ctx, dtokens):
        cols, conv_w = ctx.saved
        B, C, H, W, p, d, extra = ctx.dims
        dsum, dpe = ops.patch_embed_bwd_reduce(_as_rows_f32(dtokens), extra)
        dW = ops.gemm_wgrad(dpe, cols).view(conv_w.shape)
        dpos = dsum[extra:]
        db = ops.colsum_f32(dpos.contiguous()) if ctx.has_bias else None
        dextra = dsum[:extra] if ctx.has_extra else None
        dx = None
        if ctx.x_needs_grad:
            dcols = ops.gemm_dgrad(dpe, bf16_of(conv_w).view(d, -1))
            dx = ops.col2im(dcols, B, C, H, W, p)
        return dx, dW, db, dpos, dextra, None


class PatchConvFn(torch.autograd.Function):
    """nn.Conv2d with kernel_size == stride (no padding / dilation / groups) as im2col + tcgen05 GEMM: the patch
    embedding of blocks.TiTokEncoder (blocks.py:235-237,257) and the 1x1 convolutions of the tokenizer decoders
    (blocks.py:329-333, train_titok.py:67).  Output is NCHW-shaped (a channels-last view of the GEMM result)."""

    @staticmethod
    def forward(ctx, x, weight, bias, p):
        B, C, H, W = x.shape
        d = weight.shape[0]
        cols = ops.im2col_bf16(_as_rows_f32(x), p)
        y = ops.gemm_bias_f32(cols, bf16_of(weight).view(d, -1), _f32c(bias))
        ctx.saved = (cols, weight)
        ctx.dims = (B, C, H, W, p, d)
        ctx.x_needs_grad = x.requires_grad
        ctx.has_bias = bias is not None
        return y.view(B, H // p, W // p, d).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        cols, weight = ctx.saved
        B, C, H, W, p, d = ctx.dims
        dy2 = dy.permute(0, 2, 3, 1).reshape(-1, d).to(BF16).contiguous()
        dW, db = ops.gemm_wgrad(dy2, cols, want_bias=True)
        dx = None
        if ctx.x_needs_grad:
            dcols = ops.gemm_dgrad(dy2, bf16_of(weight).view(d, -1))
            dx = ops.col2im(dcols, B, C, H, W, p)
        return dx, dW.view(weight.shape), (db if ctx.has_bias else None), None


# ------------------------------------------------------------------------------------------------------------
# Classifier head and loss: head(vit(x)[:, 0]) (train_vit.py:51-53), nn.CrossEntropyLoss (train_vit.py:81,102)
# ------------------------------------------------------------------------------------------------------------
def _pad_rows(t, rows):
    if t is None or t.shape[0] == rows:
        return t
    out = t.new_zeros((rows,) + tuple(t.shape[1:]))
    out[: t.shape[0]] = t
    return out


def _pad_cols(t, cols):
    if t.shape[-1] == cols:
        return t
    return torch.nn.functional.pad(t, (0, cols - t.shape[-1]))


class TokenLinearFn(torch.autograd.Function):
    """y[B*cnt, C] = x[:, t0:t0+cnt] @ W^T + b for x [B, N, d] fp32: the classifier head on token 0 (train_vit.py:53) and
    TiTokEncoder.proj on the latent tokens (train_titok.py:41-42).  The token rows are gathered straight into the bf16
    GEMM operand; backward writes the [B, N, d] gradient (zero outside those tokens) and its bf16 twin in one pass of
    stores.  The output width is padded to a multiple of 8 inside (TMA row pitch): 10 classes, 12 latent dims."""

    @staticmethod
    def forward(ctx, x, weight, bias, t0, cnt, out_bf16):
        B, N, d = x.shape
        C = weight.shape[0]
        Cp = (C + 7) // 8 * 8
        a = ops.gather_tokens_bf16(_as_rows_f32(x), t0, cnt)
        w16 = _pad_rows(bf16_of(weight), Cp)
        b32 = _pad_rows(_f32c(bias), Cp)
        y = ops.gemm_bias(a, w16, b32) if out_bf16 else ops.gemm_bias_f32(a, w16, b32)
        ctx.saved = (a, w16, weight)
        ctx.dims = (B, N, d, C, Cp, t0, cnt)
        ctx.has_bias = bias is not None
        ctx.x_needs_grad = x.requires_grad
        return y[:, :C] if Cp != C else y

    @staticmethod
    def backward(ctx, dy):
        a, w16, weight = ctx.saved
        B, N, d, C, Cp, t0, cnt = ctx.dims
        dy16 = _pad_cols(dy.to(BF16), Cp).contiguous()
        dW, db = ops.gemm_wgrad(dy16, a, want_bias=True)
        dx = None
        if ctx.x_needs_grad:
            da = ops.gemm_dgrad(dy16, w16)
            dx, dx16 = ops.scatter_tokens(da, B, N, t0, cnt, want_bf16=True)
            # TransformerStackFn.backward picks the twin up instead of re-casting 155 MB; the version guards against
            # autograd accumulating another gradient into dx in place on the way there
            dx._b200_bf16_twin = (dx16, dx._version)
        ctx.saved = None
        return dx, dW[:C], (db[:C] if ctx.has_bias else None), None, None, None


class LinearFn(torch.autograd.Function):
    """nn.Linear on rows [M, K] through the tcgen05 GEMMs, for the skinny projections either side of the quantiser
    (TiTokDecoder.quant_proj 12 -> d, train_titok.py:66,70).  K and N are zero-padded to multiples of 8 inside."""

    @staticmethod
    def forward(ctx, x, weight, bias, out_bf16):
        K, N = weight.shape[1], weight.shape[0]
        Kp, Np = (K + 7) // 8 * 8, (N + 7) // 8 * 8
        x2 = x.reshape(-1, K)
        a = _pad_cols(x2.detach().to(BF16), Kp).contiguous()
        w16 = _pad_rows(_pad_cols(bf16_of(weight), Kp), Np).contiguous()
        b32 = _pad_rows(_f32c(bias), Np)
        y = ops.gemm_bias(a, w16, b32) if out_bf16 else ops.gemm_bias_f32(a, w16, b32)
        ctx.saved = (a, w16)
        ctx.dims = (tuple(x.shape), x.dtype, K, N, Kp, Np)
        ctx.has_bias = bias is not None
        ctx.x_needs_grad = x.requires_grad
        return (y[:, :N] if Np != N else y).reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        a, w16 = ctx.saved
        xshape, xdtype, K, N, Kp, Np = ctx.dims
        dy16 = _pad_cols(dy.reshape(-1, N).to(BF16), Np).contiguous()
        dW, db = ops.gemm_wgrad(dy16, a, want_bias=True)
        dx = None
        if ctx.x_needs_grad:
            dx = ops.gemm_dgrad(dy16, w16)[:, :K].to(xdtype).reshape(xshape)
        ctx.saved = None
        return dx, dW[:N, :K], (db[:N] if ctx.has_bias else None), None


_DEPATCH_PERM = {}


def _depatch_perm(C, p, device):
    """index of output channel (c p1 p2) in the reference's (p1 p2 c) order (train_titok.py:74) and its inverse."""
    key = (C, p, str(device))
    if key not in _DEPATCH_PERM:
        perm = torch.arange(C * p * p).view(p, p, C).permute(2, 0, 1).reshape(-1)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(C * p * p)
        _DEPATCH_PERM[key] = (perm.to(device), inv.to(device))
    return _DEPATCH_PERM[key]


def _depatch_operands(weight, bias, Cpp, d, perm):
    """Channel-major bf16 weight / fp32 bias of the de-patchify GEMM, cached on the weight until an optimizer steps or the
    parameters change (the (p1 p2 c) -> (c p1 p2) row permutation used to run as two index_select kernels per forward)."""
    key = (_STEP_EPOCH, weight._version, weight.data_ptr(), None if bias is None else (bias._version, bias.data_ptr()))
    cached = getattr(weight, "_b200_depatch", None)
    if cached is not None and cached[0] == key:
        return cached[1], cached[2]
    w_c = bf16_of(weight).view(Cpp, d).index_select(0, perm)
    b_c = None if bias is None else _f32c(bias).index_select(0, perm)
    try:
        weight._b200_depatch = (key, w_c, b_c)
    except Exception:  # pragma: no cover
        pass
    return w_c, b_c


class DepatchifyFn(torch.autograd.Function):
    """De-patchify tail of the tokenizer decoders (train_titok.py:67,71-74): tokens[:, :P] -> 'b (h w) c -> b c h w' ->
    Conv2d(d, C*p*p, 1) -> 'b (p1 p2 c) h w -> b c (h p1) (w p2)', as ONE tcgen05 GEMM over the gathered bf16 token rows
    whose epilogue stores straight into the NCHW fp32 image (the conv's output channels are re-ordered to (c p1 p2) so that
    a thread's consecutive accumulator columns are horizontally adjacent pixels).  Backward: im2col of the image gradient
    (the patch-embedding kernel) feeds the wgrad / dgrad GEMMs; the token gradient is scattered with its bf16 twin."""

    @staticmethod
    def forward(ctx, tokens, weight, bias, Ht, Wt, p):
        B, N, d = tokens.shape
        Cpp = weight.shape[0]
        C = Cpp // (p * p)
        P = Ht * Wt
        perm, inv = _depatch_perm(C, p, tokens.device)
        rows = ops.gather_tokens_bf16(_as_rows_f32(tokens), 0, P)
        w_c, b_c = _depatch_operands(weight, bias, Cpp, d, perm)
        img = ops.depatchify_fwd(rows, w_c, b_c, B, Ht, Wt, p, C)
        ctx.saved = (rows, w_c, inv)
        ctx.dims = (B, N, d, C, P, p, tuple(weight.shape))
        ctx.has_bias = bias is not None
        ctx.x_needs_grad = tokens.requires_grad
        return img

    @staticmethod
    def backward(ctx, dimg):
        rows, w_c, inv = ctx.saved
        B, N, d, C, P, p, wshape = ctx.dims
        dcols = ops.im2col_bf16(_as_rows_f32(dimg), p)             # [B*P, (c p1 p2)] bf16
        dW_c, db_c = ops.gemm_wgrad(dcols, rows, want_bias=True)
        dx = None
        if ctx.x_needs_grad:
            drows = ops.gemm_dgrad(dcols, w_c)
            dx, dx16 = ops.scatter_tokens(drows, B, N, 0, P, want_bf16=True)
            dx._b200_bf16_twin = (dx16, dx._version)
        ctx.saved = None
        return dx, dW_c.index_select(0, inv).view(wshape), (db_c.index_select(0, inv) if ctx.has_bias else None), None, None, None


# ------------------------------------------------------------------------------------------------------------
# Inference-only helpers for VideoGPT.generate (train_videogpt.py:56-65): prefill that keeps every layer's fused QKV rows
# as the KV cache, and the single-token step over that cache.  dropout must be 0 (the reference's generate() is otherwise
# stochastic even in eval mode, transformer.py:28).
# ------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def stack_prefill(x, layers, n_heads, n_max):
    """x [B, S, d] fp32 through the causal stack; returns (h [B, S, d] fp32, caches: one [2, B, H, n_max, 64] bf16 per layer)."""
    B, S, d = x.shape
    h = _as_rows_f32(x).view(B * S, d)
    caches = []
    for P in layers:
        h, saved = layer_forward(h, P, B, S, n_heads, True, True)
        qkv = saved[2]                                             # [B*S, 3d] bf16 == [B, S, 3, H, 64]
        cache = ops.kv_cache_alloc(B, n_heads, n_max, x.device)
        ops.kv_fill(qkv, cache, B, S)
        caches.append(cache)
    return h.view(B, S, d), caches


@torch.no_grad()
def stack_decode_step(x, layers, caches, pos_dev):
    """One new token per sequence: x [B, d] fp32 at position *pos_dev (int32 device scalar); appends its q|k|v row to the
    caches, returns [B, d] fp32.  Nothing here depends on the position on the host, so the step can be captured once in a
    CUDA graph and replayed for every generated token."""
    h = _as_rows_f32(x)
    for P, cache in zip(layers, caches):
        a, _, _, _, _ = ops.layernorm_fwd(h)
        qkv = ops.gemm_bias(a, bf16_of(P.qkv_w), _f32c(P.qkv_b))   # [B, 3d]
        ops.kv_append(qkv, cache, pos_dev)
        o = ops.attn_decode(qkv, cache, pos_dev)
        b, _, _, _, x1 = ops.layernorm_fwd(h, add=o, want_x_out=True)
        g, _ = ops.gemm_bias_gelu(b, bf16_of(P.fc1_w), _f32c(P.fc1_b))
        h = ops.gemm_bias_residual(g, bf16_of(P.fc2_w), _f32c(P.fc2_b), x1)
    return h


class EmbedFn(torch.autograd.Function):
    """tok_embed(idx) + pos_embed(arange(S)) (train_videogpt.py:50) -> fp32 [B, S, d], the stack's input."""

    @staticmethod
    def forward(ctx, idx, tok_embed, pos_embed):
        idx = idx.contiguous()
        ctx.saved = (idx,)
        ctx.dims = (tok_embed.shape[0], pos_embed.shape[0])
        return ops.embed_fwd(idx, _f32c(tok_embed), _f32c(pos_embed), 0)

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved
        vocab, n_pos = ctx.dims
        dtok, dpos = ops.embed_bwd(idx, _as_rows_f32(dy), vocab, n_pos)
        return None, dtok, dpos


class CrossEntropyFn(torch.autograd.Function):
    """F.cross_entropy(logits, labels) with reduction='mean' and ignore_index (train_vit.py:81,102, train_videogpt.py:54)."""

    @staticmethod
    def forward(ctx, logits, labels, ignore_index):
        C = logits.shape[-1]
        x2 = logits.reshape(-1, C)
        if x2.dtype not in (BF16, F32):
            x2 = x2.float()
        if x2.stride(1) != 1:
            x2 = x2.contiguous()
        y = labels.reshape(-1).contiguous()
        loss, lse = ops.cross_entropy_fwd(x2, y, ignore_index)
        ctx.saved = (x2, y, lse, loss)
        ctx.meta = (logits.shape, logits.dtype, ignore_index)
        return loss[0]

    @staticmethod
    def backward(ctx, dloss):
        x2, y, lse, loss = ctx.saved
        shape, dtype, ignore_index = ctx.meta
        d = ops.cross_entropy_bwd(x2, y, lse, loss, dloss.detach().float().reshape(1).contiguous(), ignore_index)
        ctx.saved = None
        return d.to(dtype).view(shape), None, None


# ------------------------------------------------------------------------------------------------------------
# blocks.ResidualAttentionBlock (blocks.py:32-70): affine LN, MHA (in_proj + out_proj), [L, B, d] layout
# blocks.UViTBlock (blocks.py:174-201): the same block batch-first ([B, L, d]), QKV bias optional
#
# The two affine LayerNorms are folded into the Linear that follows each (in_proj, c_fc): see folded_of / csrc/affine_fold.cu.
# The block then runs on the affine-free LayerNorm + GEMM kernels of transformer.TransformerLayer, keeps only the bf16 x-hat
# (no fp32 residual rows, no means) for backward, and a stack of blocks is ONE autograd node that hands the bf16 twin of
# the residual gradient from layer to layer and delivers its weight gradients into the data-parallel buckets.
# ------------------------------------------------------------------------------------------------------------
class RABParams:
    __slots__ = ("ln1_w", "ln1_b", "in_w", "in_b", "out_w", "out_b", "ln2_w", "ln2_b", "fc_w", "fc_b", "proj_w", "proj_b")
    N_ATTN, N_ALL = 6, 12

    def __init__(self, *p):
        p = tuple(p) + (None,) * (self.N_ALL - len(p))
        (self.ln1_w, self.ln1_b, self.in_w, self.in_b, self.out_w, self.out_b,
         self.ln2_w, self.ln2_b, self.fc_w, self.fc_b, self.proj_w, self.proj_b) = p

    def tensors(self, has_mlp):
        t = (self.ln1_w, self.ln1_b, self.in_w, self.in_b, self.out_w, self.out_b)
        return t + ((self.ln2_w, self.ln2_b, self.fc_w, self.fc_b, self.proj_w, self.proj_b) if has_mlp else ())


def rab_layer_forward(x0, P: RABParams, B, L, H, seq_first, has_mlp, save):
    """x0 [rows, d] fp32 (rows ordered (b, l), or (l, b) when seq_first).  Returns (x_out fp32, saved)."""
    M, d = x0.shape
    xh1, _, _, rstd1, _ = ops.layernorm_fwd(x0)                               # affine-free x-hat; gamma / beta live in w_in / b_in
    w_in, b_in = folded_of(P.in_w, P.in_b, P.ln1_w, P.ln1_b)
    qkv = ops.gemm_bias(xh1, w_in, b_in)
    o, lse = ops.flash_attn_fwd(qkv, B, L, H, False, want_lse=save, seq_first=seq_first)
    o2 = o.view(M, d)
    x1 = ops.gemm_bias_residual(o2, bf16_of(P.out_w), _f32c(P.out_b), x0)     # out_proj + residual (blocks.py:60,67)
    if not has_mlp:
        return x1, ((rstd1, xh1, qkv, o, lse) if save else None)
    xh2, _, _, rstd2, _ = ops.layernorm_fwd(x1)
    w_fc, b_fc = folded_of(P.fc_w, P.fc_b, P.ln2_w, P.ln2_b)
    g, u = ops.gemm_bias_gelu(xh2, w_fc, b_fc)                                # u := GELU'(pre-activation)
    x2 = ops.gemm_bias_residual(g, bf16_of(P.proj_w), _f32c(P.proj_b), x1)
    return x2, ((rstd1, xh1, qkv, o, lse, rstd2, xh2, u, g) if save else None)


def _folded_linear_backward(dy16, xhat, W, bias, gamma, beta, sink):
    """Gradients of Linear(affine LN(.)) with the LN folded into the weights: wgrad on x-hat -> dW', db'; then
    dW = dW' diag(gamma) + db' (x) beta, dgamma = colsum(dW' * W), dbeta = W^T db' (affine_unfold_grads).  Returns the four gradients in
    the order (gamma, beta, W, bias); entries that were written into their data-parallel bucket slots come back as None."""
    s_w, s_b, s_g, s_be = _slot(sink, W), (_slot(sink, bias) if bias is not None else None), _slot(sink, gamma), _slot(sink, beta)
    dW, db = ops.gemm_wgrad(dy16, xhat, out=s_w, bias_out=s_b, want_bias=True)
    dg, dbe = ops.affine_unfold_grads(dW, _f32c(W), _f32c(gamma), _f32c(beta), db, dgamma=s_g if s_be is not None else None,
                                      dbeta=s_be if s_g is not None else None)
    both = s_g is not None and s_be is not None     # the pair goes through the sink together or not at all
    _ready(sink, *(q for q, s_ in ((W, s_w), (bias, s_b), (gamma, s_g if both else None), (beta, s_be if both else None)) if s_ is not None))
    return (None if both else dg, None if both else dbe, None if s_w is not None else dW,
            None if (s_b is not None or bias is None) else db)


def rab_layer_backward(dx2, dx2_bf16, saved, P: RABParams, B, L, H, seq_first, has_mlp, need_dx, sink):
    """Returns (dx0 fp32, dx0 bf16, gradients in RABParams.tensors(has_mlp) order)."""
    M, d = dx2.shape
    mlp_grads = ()
    if has_mlp:
        rstd1, xh1, qkv, o, lse, rstd2, xh2, u, g = saved
        dv = dx2_bf16 if dx2_bf16 is not None else ops.cast_bf16(dx2)
        s_pw, s_pb = _slot(sink, P.proj_w), _slot(sink, P.proj_b)
        d_proj_w, d_proj_b = ops.gemm_wgrad(dv, g, out=s_pw, bias_out=s_pb, want_bias=True)
        _ready(sink, *(q for q, s_ in ((P.proj_w, s_pw), (P.proj_b, s_pb)) if s_ is not None))
        du = ops.gemm_dgrad_dgelu(dv, bf16_of(P.proj_w), u)
        d_ln2_w, d_ln2_b, d_fc_w, d_fc_b = _folded_linear_backward(du, xh2, P.fc_w, P.fc_b, P.ln2_w, P.ln2_b, sink)
        db = ops.gemm_dgrad(du, folded_of(P.fc_w, P.fc_b, P.ln2_w, P.ln2_b)[0])
        dx1, dx1_16 = ops.layernorm_bwd_xhat(db, xh2, rstd2, dres=dx2, want_bf16=True)
        mlp_grads = (d_ln2_w, d_ln2_b, d_fc_w, d_fc_b, None if s_pw is not None else d_proj_w, None if s_pb is not None else d_proj_b)
    else:
        rstd1, xh1, qkv, o, lse = saved
        dx1, dx1_16 = dx2, (dx2_bf16 if dx2_bf16 is not None else ops.cast_bf16(dx2))
    s_ow, s_ob = _slot(sink, P.out_w), _slot(sink, P.out_b)
    d_out_w, d_out_b = ops.gemm_wgrad(dx1_16, o.view(M, d), out=s_ow, bias_out=s_ob, want_bias=True)
    _ready(sink, *(q for q, s_ in ((P.out_w, s_ow), (P.out_b, s_ob)) if s_ is not None))
    do = ops.gemm_dgrad(dx1_16, bf16_of(P.out_w))
    dqkv = ops.flash_attn_bwd(qkv, o, do.view(o.shape), lse, B, L, H, False, seq_first=seq_first).view(M, -1)
    d_ln1_w, d_ln1_b, d_in_w, d_in_b = _folded_linear_backward(dqkv, xh1, P.in_w, P.in_b, P.ln1_w, P.ln1_b, sink)
    dx0 = dx0_16 = None
    if need_dx:
        da = ops.gemm_dgrad(dqkv, folded_of(P.in_w, P.in_b, P.ln1_w, P.ln1_b)[0])
        dx0, dx0_16 = ops.layernorm_bwd_xhat(da, xh1, rstd1, dres=dx1, want_bf16=True)
    grads = (d_ln1_w, d_ln1_b, d_in_w, d_in_b, None if s_ow is not None else d_out_w, None if s_ob is not None else d_out_b) + mlp_grads
    return dx0, dx0_16, grads


class ResidualAttentionStackFn(torch.autograd.Function):
    """n x blocks.ResidualAttentionBlock (n = 1: the module on its own; n = num_layers: the transformer of
    blocks.TiTokEncoder / TiTokDecoder, blocks.py:246-251,271-272) as one autograd node.  x is [L, B, d] (the reference's
    LND layout) or, with batch_first, [B, L, d] (UViTBlock; the internal layout of the encoder / decoder drop-ins)."""

    @staticmethod
    def forward(ctx, x, n_heads, has_mlp, batch_first, *params):
        if batch_first:
            B, L, d = x.shape
        else:
            L, B, d = x.shape
        per = RABParams.N_ALL if has_mlp else RABParams.N_ATTN
        layers = [RABParams(*params[per * i:per * (i + 1)]) for i in range(len(params) // per)]
        need_grad = any(ctx.needs_input_grad)
        h = _as_rows_f32(x).view(L * B, d)
        saved_all = []
        for P in layers:
            h, saved = rab_layer_forward(h, P, B, L, n_heads, not batch_first, has_mlp, need_grad)
            saved_all.append(saved)
        ctx.layers, ctx.saved_all = layers, saved_all
        ctx.dims = (B, L, d, n_heads, has_mlp, batch_first, tuple(x.shape))
        ctx.sink = _GRAD_SINK
        ctx.x_needs_grad = x.requires_grad
        return h.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        B, L, d, H, has_mlp, batch_first, xshape = ctx.dims
        dx = _as_rows_f32(dy).view(L * B, d)
        twin = getattr(dy, "_b200_bf16_twin", None)
        dx16 = None
        if twin is not None and twin[1] == dy._version and dy.dtype == F32 and dy.is_contiguous() and twin[0].shape == dy.shape:
            dx16 = twin[0].view(L * B, d)
        grads = []
        for i in range(len(ctx.layers) - 1, -1, -1):
            need_dx = i > 0 or ctx.x_needs_grad
            dx, dx16, g = rab_layer_backward(dx, dx16, ctx.saved_all[i], ctx.layers[i], B, L, H, not batch_first, has_mlp,
                                             need_dx, ctx.sink)
            ctx.saved_all[i] = None
            grads.append(g)
        flat = []
        for g in reversed(grads):
            flat.extend(g)
        ctx.saved_all = None
        return (dx.view(xshape) if dx is not None else None, None, None, None, *flat)


# ------------------------------------------------------------------------------------------------------------
# Token-sequence assembly + the LayerNorms around the stack of blocks.TiTokEncoder / TiTokDecoder (blocks.py:254-282, 337-361)
# ------------------------------------------------------------------------------------------------------------
class TokensAssembleFn(torch.autograd.Function):
    """tokens[B, extra + P + tail, d] fp32 (batch-first) of the blocks.py encoder / decoder front ends in one GEMM + one pass:

      encoder (blocks.py:257-267): src = image [B, C, H, W], patch > 0: rows [1, 1+P) = patch_embed(image) + positional_embedding[1:],
               row 0 = class_embedding + positional_embedding[0], tail = latent_tokens + latent_token_positional_embedding
      decoder (blocks.py:340-352): src = latents [B, P, K], patch == 0: rows [extra, extra+P) = decoder_embed(latents) +
               latent_token_positional_embedding, head rows = [class_embedding; mask_token x grid^2] + positional_embedding

    weight is the Conv2d [d, C, p, p] / Linear [d, K] weight.  Every broadcast row and positional table gets its gradient
    from one batch reduction of the token gradient (tokens_assemble_bwd_reduce)."""

    @staticmethod
    def forward(ctx, src, weight, bias, pos, head0, head1, head_pos, tail_a, tail_b, patch):
        d = weight.shape[0]
        if patch > 0:
            B, C, H, W = src.shape
            cols = ops.im2col_bf16(_as_rows_f32(src), patch)
            P = (H // patch) * (W // patch)
            K = Kp = C * patch * patch
            w16 = bf16_of(weight).view(d, K)
        else:
            B, P, K = src.shape
            Kp = (K + 7) // 8 * 8
            cols = _pad_cols(src.detach().reshape(B * P, K).to(BF16), Kp).contiguous()
            w16 = _pad_cols(bf16_of(weight), Kp).contiguous()
        extra = 0 if head_pos is None else head_pos.shape[0]
        tail = 0 if tail_a is None else tail_a.shape[0]
        tokens = ops.tokens_assemble_fwd(cols, w16, _f32c(bias), _f32c(pos), _f32c(head0), _f32c(head1), _f32c(head_pos),
                                         _f32c(tail_a), _f32c(tail_b), B, P, extra, tail)
        ctx.saved = (cols, w16)
        ctx.dims = (tuple(src.shape), patch, P, K, Kp, d, extra, tail, tuple(weight.shape))
        ctx.flags = (src.requires_grad, bias is not None, head1 is not None, tail_b is not None)
        return tokens

    @staticmethod
    def backward(ctx, dtokens):
        cols, w16 = ctx.saved
        sshape, patch, P, K, Kp, d, extra, tail, wshape = ctx.dims
        src_grad, has_bias, has_head1, has_tail_b = ctx.flags
        dsum, dpe = ops.tokens_assemble_bwd_reduce(_as_rows_f32(dtokens), extra, tail)
        dW = ops.gemm_wgrad(dpe, cols)
        dW = (dW[:, :K] if Kp != K else dW).reshape(wshape)
        dpos = dsum[extra:extra + P]
        db = ops.colsum_f32(dpos.contiguous()) if has_bias else None
        d_head0 = dsum[0:1] if extra > 0 else None
        d_head1 = ops.colsum_f32(dsum[1:extra].contiguous()).view(1, d) if (has_head1 and extra > 1) else None
        d_head_pos = dsum[:extra] if extra > 0 else None
        d_tail = dsum[extra + P:] if tail > 0 else None
        dsrc = None
        if src_grad:
            dcols = ops.gemm_dgrad(dpe, w16)
            if patch > 0:
                B, C, H, W = sshape
                dsrc = ops.col2im(dcols, B, C, H, W, patch)
            else:
                dsrc = dcols[:, :K].float().reshape(sshape)
        ctx.saved = None
        return dsrc, dW, db, dpos, d_head0, d_head1, d_head_pos, d_tail, (d_tail if has_tail_b else None), None


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm (affine) fp32 -> fp32 on a token range x[:, t0:t0+cnt] of x [B, N, d]: ln_pre over the whole sequence
    (t0 = 0, cnt = N; blocks.py:269,353) and ln_post over the latent / patch tokens only (blocks.py:275-276, 358-359) -- the
    range is gathered by the kernel's caller and its gradient scattered back with the bf16 twin the stack's backward wants."""

    @staticmethod
    def forward(ctx, x, weight, bias, t0, cnt, eps):
        B, N, d = x.shape
        xf = _as_rows_f32(x)
        rows = xf if (t0 == 0 and cnt == N) else ops.gather_tokens_f32(xf, t0, cnt)
        _, y, mean, rstd, _ = ops.layernorm_fwd(rows.view(B * cnt, d), gamma=_f32c(weight), beta=_f32c(bias), out_bf16=False,
                                                out_f32=True, eps=eps)
        ctx.saved = (rows, mean, rstd, weight)
        ctx.dims = (B, N, d, t0, cnt)
        ctx.x_needs_grad = x.requires_grad
        return y.view(B, cnt, d)

    @staticmethod
    def backward(ctx, dy):
        rows, mean, rstd, weight = ctx.saved
        B, N, d, t0, cnt = ctx.dims
        drows, _, dg, db = ops.layernorm_bwd(_as_rows_f32(dy).view(B * cnt, d), rows.view(B * cnt, d), mean, rstd,
                                             gamma=_f32c(weight), want_bf16=False, affine_grads=True)
        dx = None
        if ctx.x_needs_grad:
            if t0 == 0 and cnt == N:
                dx = drows.view(B, N, d)
            else:
                dx, dx16 = ops.scatter_tokens(drows, B, N, t0, cnt, want_bf16=True)
                dx._b200_bf16_twin = (dx16, dx._version)
        ctx.saved = None
        return dx, dg, db, None, None, None


# ------------------------------------------------------------------------------------------------------------
# VQ lookups: train_titok.Quantizer (train_titok.py:45-59) and blocks.VectorQuantizer (blocks.py:405-505)
# ------------------------------------------------------------------------------------------------------------
class VQFn(torch.autograd.Function):
    """Returns (quantized, indices, mse, commitment_cost*mse, (1+cc)*mse).  `gather_normalized`/`channels_first`
    select the blocks.py flavour.  Gradients: straight-through to x plus the two MSE terms."""

    @staticmethod
    def forward(ctx, x, codebook, l2, gather_normalized, channels_first, commitment_cost):
        xf = _as_rows_f32(x)
        cb = _f32c(codebook)
        q, idx, losses = ops.vq_fwd(xf, cb, l2=l2, gather_normalized=gather_normalized,
                                    channels_first=channels_first, commitment_cost=commitment_cost)
        ctx.saved = (xf, cb, idx)
        ctx.cfg = (l2, gather_normalized, channels_first, commitment_cost)
        ctx.mark_non_differentiable(idx)
        return q, idx, losses[0], losses[1], losses[2]

    @staticmethod
    def backward(ctx, dq, _didx, d_mse, d_commit, d_total):
        xf, cb, idx = ctx.saved
        l2, gn, cf, cc = ctx.cfg
        # d/d(codebook_loss = mse) -> code rows ; d/d(commitment = cc*mse) -> x ; total = both
        zero = torch.zeros((), device=xf.device, dtype=F32)
        d_mse = zero if d_mse is None else d_mse.float()
        d_commit = zero if d_commit is None else d_commit.float()
        d_total = zero if d_total is None else d_total.float()
        coef = torch.stack([cc * (d_commit + d_total), d_mse + d_total]).contiguous()
        gq = torch.zeros_like(xf) if dq is None else _as_rows_f32(dq)
        dx, dC = ops.vq_bwd(xf, cb, idx, gq, coef, l2=l2, gather_normalized=gn, channels_first=cf)
        return dx, dC, None, None, None, None
