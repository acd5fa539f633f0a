"""ORACLE — test infrastructure only (never imported by the product path).

A CPU (numpy) restatement of the reference's ViT-encoder hot path, forward AND backward, written from
the reference's module definitions.  Every function cites the reference lines it restates (paths are
relative to the upstream repo SnakeOnex/vit-is-all-you-need).  All arithmetic runs in the dtype of the
inputs (float32 to mimic the reference's CPU path, float64 for error budgeting).

Pinning: the reference ships no tests or golden vectors, so this oracle is pinned
against outputs of the reference modules themselves, generated in the build container by
tests/golden/make_golden.py (which imports /root/reference) and committed under tests/golden/*.npz;
tests/test_oracle_golden.py checks every function here against those fixtures.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import math

import numpy as np

try:  # scipy is part of the image; keep a pure-numpy erf so the oracle never silently changes meaning
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

LN_EPS = 1e-5  # F.layer_norm / nn.LayerNorm default eps (transformer.py:43-44, blocks.py:43,48)


# ------------------------------------------------------------------------------------------------
# LayerNorm — F.layer_norm(x, (d,)) affine-free (transformer.py:43-44) / nn.LayerNorm affine (blocks.py:43,48)
# ------------------------------------------------------------------------------------------------
def layer_norm_fwd(x, weight=None, bias=None, eps=LN_EPS):
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)  # biased variance
    rstd = 1.0 / np.sqrt(var + x.dtype.type(eps))
    xhat = (x - mean) * rstd
    y = xhat
    if weight is not None:
        y = y * weight
    if bias is not None:
        y = y + bias
    return y, (xhat, rstd, weight)


def layer_norm_bwd(dy, cache):
    xhat, rstd, weight = cache
    dw = db = None
    g = dy
    if weight is not None:
        dw = (dy * xhat).reshape(-1, xhat.shape[-1]).sum(axis=0)
        db = dy.reshape(-1, xhat.shape[-1]).sum(axis=0)
        g = dy * weight
    m1 = g.mean(axis=-1, keepdims=True)
    m2 = (g * xhat).mean(axis=-1, keepdims=True)
    dx = (g - m1 - xhat * m2) * rstd
    return dx, dw, db


# ------------------------------------------------------------------------------------------------
# nn.Linear (transformer.py:21,37,39 ; blocks.py:51,53 ; MHA in/out projections blocks.py:44)
# ------------------------------------------------------------------------------------------------
def linear_fwd(x, w, b=None):
    y = x @ w.T
    if b is not None:
        y = y + b
    return y


def linear_bwd(dy, x, w):
    dx = dy @ w
    dw = dy.reshape(-1, dy.shape[-1]).T @ x.reshape(-1, x.shape[-1])
    db = dy.reshape(-1, dy.shape[-1]).sum(axis=0)
    return dx, dw, db


# ------------------------------------------------------------------------------------------------
# nn.GELU() — exact erf form (transformer.py:38, blocks.py:52)
# ------------------------------------------------------------------------------------------------
def gelu_fwd(u):
    return (0.5 * u * (1.0 + _erf(u / math.sqrt(2.0)))).astype(u.dtype)


def gelu_bwd(dg, u):
    cdf = 0.5 * (1.0 + _erf(u / math.sqrt(2.0)))
    pdf = np.exp(-0.5 * u * u) / math.sqrt(2.0 * math.pi)
    return (dg * (cdf + u * pdf)).astype(u.dtype)


# The CUDA path keeps GELU'(u) for backward as an 8-bit fixed-point code (csrc/gemm_tcgen05.cuh GP_LO / GP_STEP): the
# derivative of nn.GELU() (transformer.py:38) lies in [-0.129, 1.129] for every u.  Restated here so that tests can check the
# kernel's codes and bound the error the code adds (<= GELU_GRAD_STEP / 2 absolute).
GELU_GRAD_STEP = 1.27 / 255.0
GELU_GRAD_LO = -27.0 * GELU_GRAD_STEP


def gelu_grad_code(u):
    gp = gelu_bwd(np.ones_like(u), u)
    return np.clip(np.rint((gp - GELU_GRAD_LO) / GELU_GRAD_STEP), 0, 255).astype(np.uint8)


def gelu_grad_decode(code):
    return GELU_GRAD_LO + GELU_GRAD_STEP * code.astype(np.float64)


# ------------------------------------------------------------------------------------------------
# F.scaled_dot_product_attention(q,k,v, attn_mask=-inf upper triangle or None) (transformer.py:22-28)
# q,k,v: [B,h,N,hd]; scale 1/sqrt(hd); dropout_p = 0 (parity is only defined at p = 0)
# ------------------------------------------------------------------------------------------------
def sdpa_fwd(q, k, v, causal=False, keep=None, dropout_p=0.0):
    """F.scaled_dot_product_attention (transformer.py:28).  `keep` (bool [.., n, n]) with `dropout_p` restates its
    dropout: the softmax probabilities are multiplied by keep / (1 - p) before they weight v."""
    hd = q.shape[-1]
    s = (q @ np.swapaxes(k, -1, -2)) * q.dtype.type(1.0 / math.sqrt(hd))
    if causal:
        n = q.shape[-2]
        mask = np.triu(np.ones((n, n), dtype=bool), k=1)
        s = np.where(mask, -np.inf, s)
    m = s.max(axis=-1, keepdims=True)
    e = np.exp(s - m)
    p = e / e.sum(axis=-1, keepdims=True)
    pd = p if keep is None else p * keep * q.dtype.type(1.0 / (1.0 - dropout_p))
    o = pd @ v
    return o.astype(q.dtype), (q, k, v, p.astype(q.dtype), keep, dropout_p)


def sdpa_bwd(do, cache):
    q, k, v, p, keep, dropout_p = cache
    hd = q.shape[-1]
    r = q.dtype.type(1.0 if keep is None else 1.0 / (1.0 - dropout_p))
    pd = p if keep is None else p * keep * r
    dv = np.swapaxes(pd, -1, -2) @ do
    dpd = do @ np.swapaxes(v, -1, -2)
    dp = dpd if keep is None else dpd * keep * r
    ds = p * (dp - (dp * p).sum(axis=-1, keepdims=True))
    ds = ds * q.dtype.type(1.0 / math.sqrt(hd))
    dq = ds @ k
    dk = np.swapaxes(ds, -1, -2) @ q
    return dq, dk, dv


# ------------------------------------------------------------------------------------------------
# transformer.Attention (transformer.py:16-29): fused qkv Linear, split "(qkv h d)", SDPA, NO out-proj
# ------------------------------------------------------------------------------------------------
def attention_fwd(x, wqkv, bqkv, n_heads, causal=False):
    B, N, d = x.shape
    hd = d // n_heads
    qkv = linear_fwd(x, wqkv, bqkv).reshape(B, N, 3, n_heads, hd)  # "b n (qkv h d)"
    q, k, v = (np.transpose(qkv[:, :, i], (0, 2, 1, 3)) for i in range(3))  # -> b h n d
    o, c = sdpa_fwd(q, k, v, causal)
    out = np.transpose(o, (0, 2, 1, 3)).reshape(B, N, d)  # "b h n d -> b n (h d)"
    return out, (x, wqkv, c, n_heads)


def attention_bwd(dout, cache):
    x, wqkv, c, n_heads = cache
    B, N, d = x.shape
    hd = d // n_heads
    do = np.transpose(dout.reshape(B, N, n_heads, hd), (0, 2, 1, 3))
    dq, dk, dv = sdpa_bwd(do, c)
    dqkv = np.stack([np.transpose(t, (0, 2, 1, 3)) for t in (dq, dk, dv)], axis=2).reshape(B, N, 3 * d)
    dx, dw, db = linear_bwd(dqkv, x, wqkv)
    return dx, dw, db


# ------------------------------------------------------------------------------------------------
# transformer.TransformerLayer (transformer.py:31-45): x + attn(LN(x)); x + mlp(LN(x)), affine-free LN,
# mlp = Linear(d,4d) -> GELU -> Linear(4d,d) -> Dropout(p=0)
# params: dict with qkv_w[3d,d] qkv_b fc1_w[4d,d] fc1_b fc2_w[d,4d] fc2_b
# ------------------------------------------------------------------------------------------------
def transformer_layer_fwd(x, p, n_heads, causal=False):
    a, c_ln1 = layer_norm_fwd(x)
    att, c_att = attention_fwd(a, p["qkv_w"], p["qkv_b"], n_heads, causal)
    x1 = x + att
    b, c_ln2 = layer_norm_fwd(x1)
    u = linear_fwd(b, p["fc1_w"], p["fc1_b"])
    g = gelu_fwd(u)
    v = linear_fwd(g, p["fc2_w"], p["fc2_b"])
    x2 = x1 + v
    return x2, (c_ln1, c_att, c_ln2, b, u, g, p)


def transformer_layer_bwd(dx2, cache):
    c_ln1, c_att, c_ln2, b, u, g, p = cache
    grads = {}
    dg, grads["fc2_w"], grads["fc2_b"] = linear_bwd(dx2, g, p["fc2_w"])
    du = gelu_bwd(dg, u)
    db, grads["fc1_w"], grads["fc1_b"] = linear_bwd(du, b, p["fc1_w"])
    dx1 = dx2 + layer_norm_bwd(db, c_ln2)[0]
    da, grads["qkv_w"], grads["qkv_b"] = attention_bwd(dx1, c_att)
    dx0 = dx1 + layer_norm_bwd(da, c_ln1)[0]
    return dx0, grads


def transformer_fwd(x, layers, n_heads, causal=False):
    """transformer.Transformer (transformer.py:47-54): plain stack, no final norm."""
    caches = []
    for p in layers:
        x, c = transformer_layer_fwd(x, p, n_heads, causal)
        caches.append(c)
    return x, caches


def transformer_bwd(dx, caches):
    grads = []
    for c in reversed(caches):
        dx, g = transformer_layer_bwd(dx, c)
        grads.append(g)
    return dx, grads[::-1]


# ------------------------------------------------------------------------------------------------
# blocks.ResidualAttentionBlock (blocks.py:32-70): affine LN, nn.MultiheadAttention (in_proj + out_proj),
# sequence-first [N,B,d] layout, MLP c_fc -> GELU -> c_proj (no dropout)
# params: ln1_w ln1_b in_w[3d,d] in_b out_w[d,d] out_b ln2_w ln2_b fc_w[4d,d] fc_b proj_w[d,4d] proj_b
# ------------------------------------------------------------------------------------------------
def residual_attention_block_fwd(x_lnd, p, n_heads):
    L, B, d = x_lnd.shape
    hd = d // n_heads
    x = np.transpose(x_lnd, (1, 0, 2))  # work batch-first internally; LN/Linear are per-token
    a, c_ln1 = layer_norm_fwd(x, p["ln1_w"], p["ln1_b"])
    qkv = linear_fwd(a, p["in_w"], p["in_b"]).reshape(B, L, 3, n_heads, hd)  # MHA packs [q;k;v] rows
    q, k, v = (np.transpose(qkv[:, :, i], (0, 2, 1, 3)) for i in range(3))
    o, c_sdpa = sdpa_fwd(q, k, v, False)
    o2 = np.transpose(o, (0, 2, 1, 3)).reshape(B, L, d)
    att = linear_fwd(o2, p["out_w"], p["out_b"])
    x1 = x + att
    b, c_ln2 = layer_norm_fwd(x1, p["ln2_w"], p["ln2_b"])
    u = linear_fwd(b, p["fc_w"], p["fc_b"])
    g = gelu_fwd(u)
    x2 = x1 + linear_fwd(g, p["proj_w"], p["proj_b"])
    cache = (c_ln1, a, c_sdpa, o2, c_ln2, b, u, g, p, n_heads)
    return np.transpose(x2, (1, 0, 2)), cache


def residual_attention_block_bwd(dy_lnd, cache):
    c_ln1, a, c_sdpa, o2, c_ln2, b, u, g, p, n_heads = cache
    dx2 = np.transpose(dy_lnd, (1, 0, 2))
    B, L, d = dx2.shape
    hd = d // n_heads
    gr = {}
    dg, gr["proj_w"], gr["proj_b"] = linear_bwd(dx2, g, p["proj_w"])
    du = gelu_bwd(dg, u)
    db, gr["fc_w"], gr["fc_b"] = linear_bwd(du, b, p["fc_w"])
    dln2, gr["ln2_w"], gr["ln2_b"] = layer_norm_bwd(db, c_ln2)
    dx1 = dx2 + dln2
    do2, gr["out_w"], gr["out_b"] = linear_bwd(dx1, o2, p["out_w"])
    do = np.transpose(do2.reshape(B, L, n_heads, hd), (0, 2, 1, 3))
    dq, dk, dv = sdpa_bwd(do, c_sdpa)
    dqkv = np.stack([np.transpose(t, (0, 2, 1, 3)) for t in (dq, dk, dv)], axis=2).reshape(B, L, 3 * d)
    da, gr["in_w"], gr["in_b"] = linear_bwd(dqkv, a, p["in_w"])
    dln1, gr["ln1_w"], gr["ln1_b"] = layer_norm_bwd(da, c_ln1)
    dx0 = dx1 + dln1
    return np.transpose(dx0, (1, 0, 2)), gr


# ------------------------------------------------------------------------------------------------
# blocks.TiTokEncoder / TiTokDecoder (blocks.py:208-361): token-sequence assembly, ln_pre, ResidualAttentionBlock stack,
# ln_post, conv_out (encoder) / ffn de-patchify (decoder; the final 3x3 conv_out of the decoder, blocks.py:334,361, is outside
# the hot path and NOT restated: `blocks_titok_decoder_fwd` returns the image that enters it).
# P: patch_w [d,3,p,p] patch_b cls [1,d] pos [1+G,d] latent_pos [L,d] ln_pre_w/b blocks [list of RAB dicts] ln_post_w/b
#    encoder: conv_out_w [ts,d,1,1] conv_out_b ; decoder: embed_w [d,ts] embed_b mask [1,1,d] ffn_w [3pp,d,1,1] ffn_b
# ------------------------------------------------------------------------------------------------
def blocks_tokens_encoder(x, latent_tokens, P):
    """blocks.py:257-267: [cls + pos[0] | patch_embed(x) + pos[1:] | latent_tokens + latent_pos] -> [B, 1 + G + L, d]."""
    d, C, p, _ = P["patch_w"].shape
    cols = im2col(x, p)
    patches = cols @ P["patch_w"].reshape(d, C * p * p).T + P["patch_b"] + P["pos"][1:]
    B = x.shape[0]
    head = np.broadcast_to((P["cls"] + P["pos"][:1])[None], (B, 1, d))
    tail = np.broadcast_to((latent_tokens + P["latent_pos"])[None], (B,) + latent_tokens.shape)
    return np.concatenate([head, patches, tail], axis=1).astype(x.dtype), cols


def _rab_stack_fwd(h, blocks, n_heads):
    caches = []
    h = np.transpose(h, (1, 0, 2))                 # the reference runs the blocks in LND (blocks.py:270)
    for bp in blocks:
        h, c = residual_attention_block_fwd(h, bp, n_heads)
        caches.append(c)
    return np.transpose(h, (1, 0, 2)), caches


def _rab_stack_bwd(dh, caches):
    dh = np.transpose(dh, (1, 0, 2))
    grads = []
    for c in reversed(caches):
        dh, g = residual_attention_block_bwd(dh, c)
        grads.append(g)
    return np.transpose(dh, (1, 0, 2)), grads[::-1]


def blocks_titok_encoder_fwd(x, latent_tokens, P, n_heads):
    """blocks.TiTokEncoder.forward (blocks.py:254-282) -> z [B, token_size, 1, L]."""
    tokens, cols = blocks_tokens_encoder(x, latent_tokens, P)
    G, L = P["pos"].shape[0] - 1, latent_tokens.shape[0]
    h0, c_pre = layer_norm_fwd(tokens, P["ln_pre_w"], P["ln_pre_b"])
    h, c_blocks = _rab_stack_fwd(h0, P["blocks"], n_heads)
    lat, c_post = layer_norm_fwd(h[:, 1 + G:], P["ln_post_w"], P["ln_post_b"])
    ts = P["conv_out_w"].shape[0]
    y = lat @ P["conv_out_w"].reshape(ts, -1).T + P["conv_out_b"]          # 1x1 conv over [B, d, L, 1]
    z = np.transpose(y, (0, 2, 1))[:, :, None, :]
    return z, (cols, c_pre, c_blocks, c_post, lat, h.shape, G, L, P)


def blocks_titok_encoder_bwd(dz, cache):
    """Gradients of blocks.TiTokEncoder wrt its parameters and latent_tokens (the image is a leaf)."""
    cols, c_pre, c_blocks, c_post, lat, hshape, G, L, P = cache
    d = hshape[-1]
    ts = P["conv_out_w"].shape[0]
    dy = np.transpose(dz[:, :, 0, :], (0, 2, 1))                            # [B, L, ts]
    g = {"conv_out_w": (dy.reshape(-1, ts).T @ lat.reshape(-1, d)).reshape(P["conv_out_w"].shape), "conv_out_b": dy.reshape(-1, ts).sum(0)}
    dlat = dy @ P["conv_out_w"].reshape(ts, d)
    dl, g["ln_post_w"], g["ln_post_b"] = layer_norm_bwd(dlat, c_post)
    dh = np.zeros(hshape, dtype=dz.dtype)
    dh[:, 1 + G:] = dl
    dh0, g["blocks"] = _rab_stack_bwd(dh, c_blocks)
    dtok, g["ln_pre_w"], g["ln_pre_b"] = layer_norm_bwd(dh0, c_pre)
    dsum = dtok.sum(axis=0)                                                 # [T, d]
    g["cls"] = dsum[:1]
    g["pos"] = dsum[:1 + G]
    g["latent_pos"] = dsum[1 + G:]
    g["latent_tokens"] = dsum[1 + G:]
    dpe = dtok[:, 1:1 + G].reshape(-1, d)
    g["patch_b"] = dpe.sum(axis=0)
    g["patch_w"] = (dpe.T @ cols.reshape(-1, cols.shape[-1])).reshape(P["patch_w"].shape)
    return g


def blocks_titok_decoder_fwd(zq, P, n_heads, grid, p):
    """blocks.TiTokDecoder.forward up to (not including) the final 3x3 conv_out (blocks.py:337-360): zq [B, ts, 1, L] ->
    image [B, 3, grid p, grid p]."""
    B, ts, _, L = zq.shape
    d = P["embed_w"].shape[0]
    x = np.transpose(zq.reshape(B, ts, L), (0, 2, 1)) @ P["embed_w"].T + P["embed_b"] + P["latent_pos"][:L]
    G = grid * grid
    head = np.concatenate([P["cls"], np.broadcast_to(P["mask"].reshape(1, d), (G, d))], axis=0) + P["pos"]
    tokens = np.concatenate([np.broadcast_to(head[None], (B, 1 + G, d)), x], axis=1).astype(zq.dtype)
    h0, _ = layer_norm_fwd(tokens, P["ln_pre_w"], P["ln_pre_b"])
    h, _ = _rab_stack_fwd(h0, P["blocks"], n_heads)
    grid_tokens, _ = layer_norm_fwd(h[:, 1:1 + G], P["ln_post_w"], P["ln_post_b"])
    img, _ = depatchify_fwd(grid_tokens, P["ffn_w"], P["ffn_b"], grid, grid, p)
    return img, tokens


def affine_fold(W, b, gamma, beta):
    """csrc/affine_fold.cu: Linear(gamma * xhat + beta) == xhat @ (W diag(gamma))^T + (b + W beta)."""
    return W * gamma[None, :], (0.0 if b is None else b) + W @ beta


def affine_unfold_grads(dWf, dbf, W, gamma, beta):
    """Gradients of (W, gamma, beta) from the gradients of the folded pair (dWf = d/dW', dbf = d/db')."""
    return dWf * gamma[None, :] + np.outer(dbf, beta), (dWf * W).sum(axis=0), W.T @ dbf


# ------------------------------------------------------------------------------------------------
# Patch embedding as used by ViT (train_vit.py:34-36,38-45): Conv2d(k = s = p) == im2col GEMM,
# flatten (h w) row-major, + pos_emb, prepend extra_emb (extra tokens FIRST, no pos-emb on them)
# ------------------------------------------------------------------------------------------------
def im2col(x, p):
    B, C, H, W = x.shape
    gh, gw = H // p, W // p
    cols = x.reshape(B, C, gh, p, gw, p).transpose(0, 2, 4, 1, 3, 5)  # b gh gw c i j
    return cols.reshape(B, gh * gw, C * p * p)


def patch_embed_fwd(x, conv_w, conv_b, pos_emb, extra_emb):
    d, C, p, _ = conv_w.shape
    cols = im2col(x, p)
    pe = cols @ conv_w.reshape(d, C * p * p).T + conv_b + pos_emb
    B = x.shape[0]
    ext = np.broadcast_to(extra_emb[None], (B,) + extra_emb.shape)
    return np.concatenate([ext, pe], axis=1).astype(x.dtype), (cols, conv_w.shape, extra_emb.shape[0])


def patch_embed_bwd(demb, cache):
    cols, wshape, extra = cache
    d = wshape[0]
    dpe = demb[:, extra:]
    g = {
        "extra_emb": demb[:, :extra].sum(axis=0),
        "pos_emb": dpe.sum(axis=0),
        "conv_b": dpe.reshape(-1, d).sum(axis=0),
        "conv_w": (dpe.reshape(-1, d).T @ cols.reshape(-1, cols.shape[-1])).reshape(wshape),
    }
    dcols = dpe @ np.zeros((d, cols.shape[-1]), dtype=demb.dtype) if False else None  # image is a leaf
    return dcols, g


# ------------------------------------------------------------------------------------------------
# ViT / ViTClassifier (train_vit.py:30-53) + CrossEntropyLoss (train_vit.py:81,102)
# params: conv_w conv_b pos_emb extra_emb layers[list of dict] head_w head_b
# ------------------------------------------------------------------------------------------------
def vit_fwd(x, P, n_heads):
    emb, c_pe = patch_embed_fwd(x, P["conv_w"], P["conv_b"], P["pos_emb"], P["extra_emb"])
    out, c_tr = transformer_fwd(emb, P["layers"], n_heads)
    return out, (c_pe, c_tr)


def vit_bwd(dout, cache):
    c_pe, c_tr = cache
    demb, layer_grads = transformer_bwd(dout, c_tr)
    _, g = patch_embed_bwd(demb, c_pe)
    g["layers"] = layer_grads
    return g


def cross_entropy_fwd(logits, labels, ignore_index=-100):
    """nn.CrossEntropyLoss() / F.cross_entropy defaults (train_vit.py:81,102; train_videogpt.py:54): mean over the rows
    whose label is not ignore_index of logsumexp(x) - x[label]."""
    m = logits.max(axis=-1, keepdims=True)
    lse = m + np.log(np.exp(logits - m).sum(axis=-1, keepdims=True))
    logp = logits - lse
    valid = labels != ignore_index
    safe = np.where(valid, labels, 0)
    rows = -logp[np.arange(logits.shape[0]), safe] * valid
    loss = rows.sum() / valid.sum()
    return loss.astype(logits.dtype), (np.exp(logp), labels, valid)


def cross_entropy_bwd(cache):
    p, labels, valid = cache
    d = p.copy()
    d[np.arange(p.shape[0]), np.where(valid, labels, 0)] -= 1.0
    d *= valid[:, None]
    return (d / valid.sum()).astype(p.dtype)


def vit_classifier_loss_and_grads(x, labels, P, n_heads):
    """One training step's forward + backward of train_vit.py:100-104 (no optimiser)."""
    out, c = vit_fwd(x, P, n_heads)
    cls = out[:, 0]
    logits = linear_fwd(cls, P["head_w"], P["head_b"])
    loss, c_ce = cross_entropy_fwd(logits, labels)
    dlogits = cross_entropy_bwd(c_ce)
    dcls, dhw, dhb = linear_bwd(dlogits, cls, P["head_w"])
    dout = np.zeros_like(out)
    dout[:, 0] = dcls
    g = vit_bwd(dout, c)
    g["head_w"], g["head_b"] = dhw, dhb
    return loss, logits, g


# ------------------------------------------------------------------------------------------------
# Steps either side of the stack: token-range Linear (classifier head train_vit.py:53, TiTokEncoder.proj
# train_titok.py:41-42, VideoGPT.proj train_videogpt.py:52), de-patchify tail (train_titok.py:67,71-74), token + positional
# embedding (train_videogpt.py:45-50), VideoGPT forward / greedy generate (train_videogpt.py:44-65)
# ------------------------------------------------------------------------------------------------
def token_linear_fwd(x, w, b, t0, cnt):
    """y[B, cnt, C] = x[:, t0:t0+cnt] @ w^T + b."""
    rows = x[:, t0:t0 + cnt]
    return rows @ w.T + b, (rows, w, x.shape, t0, cnt)


def token_linear_bwd(dy, cache):
    rows, w, xshape, t0, cnt = cache
    d = rows.shape[-1]
    dw = dy.reshape(-1, dy.shape[-1]).T @ rows.reshape(-1, d)
    db = dy.reshape(-1, dy.shape[-1]).sum(axis=0)
    dx = np.zeros(xshape, dtype=dy.dtype)
    dx[:, t0:t0 + cnt] = dy @ w
    return dx, dw, db


def depatchify_fwd(tokens, conv_w, conv_b, Ht, Wt, p):
    """tokens[:, :Ht*Wt] -> 'b (h w) c -> b c h w' -> Conv2d(d, C*p*p, 1) -> 'b (p1 p2 c) h w -> b c (h p1) (w p2)'."""
    B, N, d = tokens.shape
    P = Ht * Wt
    Cpp = conv_w.shape[0]
    C = Cpp // (p * p)
    rows = tokens[:, :P]
    y = rows @ conv_w.reshape(Cpp, d).T + conv_b                       # [B, P, (p1 p2 c)]
    img = y.reshape(B, Ht, Wt, p, p, C).transpose(0, 5, 1, 3, 2, 4).reshape(B, C, Ht * p, Wt * p)
    return img, (rows, conv_w, tokens.shape, Ht, Wt, p, C)


def depatchify_bwd(dimg, cache):
    rows, conv_w, tshape, Ht, Wt, p, C = cache
    B, d = rows.shape[0], rows.shape[-1]
    dy = dimg.reshape(B, C, Ht, p, Wt, p).transpose(0, 2, 4, 3, 5, 1).reshape(B, Ht * Wt, p * p * C)   # (p1 p2 c) order
    dw = (dy.reshape(-1, dy.shape[-1]).T @ rows.reshape(-1, d)).reshape(conv_w.shape)
    db = dy.reshape(-1, dy.shape[-1]).sum(axis=0)
    dtok = np.zeros(tshape, dtype=dimg.dtype)
    dtok[:, :Ht * Wt] = dy @ conv_w.reshape(-1, d)
    return dtok, dw, db


def embed_fwd(idx, tok_embed, pos_embed, pos0=0):
    S = idx.shape[1]
    return tok_embed[idx] + pos_embed[pos0:pos0 + S][None], (idx, tok_embed.shape, pos_embed.shape, pos0)


def embed_bwd(dy, cache):
    idx, tshape, pshape, pos0 = cache
    dtok = np.zeros(tshape, dtype=dy.dtype)
    np.add.at(dtok, idx.reshape(-1), dy.reshape(-1, dy.shape[-1]))
    dpos = np.zeros(pshape, dtype=dy.dtype)
    dpos[pos0:pos0 + dy.shape[1]] = dy.sum(axis=0)
    return dtok, dpos


def titok_encoder_fwd(x, P, n_heads, latent_tokens):
    """train_titok.TiTokEncoder.forward (train_titok.py:40-43); P = vit params + proj_w / proj_b."""
    tokens, _ = vit_fwd(x, P, n_heads)
    lat, c = token_linear_fwd(tokens, P["proj_w"], P["proj_b"], 0, latent_tokens)
    return lat, tokens, c


def titok_decoder_fwd(z, P, n_heads, patch_dim, p):
    """train_titok.TiTokDecoder.forward (train_titok.py:69-76); P = vit params + quant_proj_w/b + embd_proj_w/b."""
    h = z @ P["quant_proj_w"].T + P["quant_proj_b"]                    # [B, L, d]
    zimg = h.transpose(0, 2, 1)[..., None]                             # 'b h c -> b c h 1'
    tokens, _ = vit_fwd(zimg, P, n_heads)
    img, c = depatchify_fwd(tokens, P["embd_proj_w"], P["embd_proj_b"], patch_dim, patch_dim, p)
    return img, tokens, c


def videogpt_fwd(x, P, n_heads, codebook_size):
    """train_videogpt.VideoGPT.forward (train_videogpt.py:44-55): x [B, T, N] int tokens -> (logits, loss, caches)."""
    B = x.shape[0]
    y = x.reshape(B, -1)
    sos = np.full((B, 1), codebook_size, dtype=y.dtype)
    inp = np.concatenate([sos, y[:, :-1]], axis=1)
    h0, c_emb = embed_fwd(inp, P["tok_embed"], P["pos_embed"])
    h, c_tr = transformer_fwd(h0, P["layers"], n_heads, True)
    logits, c_proj = token_linear_fwd(h, P["proj_w"], P["proj_b"], 0, h.shape[1])
    loss, c_ce = cross_entropy_fwd(logits.reshape(-1, logits.shape[-1]), y.reshape(-1))
    return logits, loss, (c_emb, c_tr, c_proj, c_ce, logits.shape)


def videogpt_bwd(caches):
    c_emb, c_tr, c_proj, c_ce, lshape = caches
    dlogits = cross_entropy_bwd(c_ce).reshape(lshape)
    dh, dpw, dpb = token_linear_bwd(dlogits, c_proj)
    dh0, layer_grads = transformer_bwd(dh, c_tr)
    dtok, dpos = embed_bwd(dh0, c_emb)
    return {"tok_embed": dtok, "pos_embed": dpos, "proj_w": dpw, "proj_b": dpb, "layers": layer_grads}


def videogpt_generate(tokens, n, P, n_heads, codebook_size):
    """train_videogpt.VideoGPT.generate (train_videogpt.py:56-65): the whole stack is re-run for every new token.
    Returns (tokens [B, T0 + n], top-2 logit margins [B, n])."""
    margins = []
    for _ in range(n):
        B = tokens.shape[0]
        sos = np.full((B, 1), codebook_size, dtype=tokens.dtype)
        h0, _ = embed_fwd(np.concatenate([sos, tokens], axis=1), P["tok_embed"], P["pos_embed"])
        h, _ = transformer_fwd(h0, P["layers"], n_heads, True)
        lg = h[:, -1] @ P["proj_w"].T + P["proj_b"]
        top2 = np.sort(lg, axis=-1)[:, -2:]
        margins.append(top2[:, 1] - top2[:, 0])
        tokens = np.concatenate([tokens, lg.argmax(axis=-1)[:, None].astype(tokens.dtype)], axis=1)
    return tokens, np.stack(margins, axis=1)


# ------------------------------------------------------------------------------------------------
# VQ lookups
# ------------------------------------------------------------------------------------------------
def _normalize(x, eps=1e-12):
    """F.normalize(x, dim=-1): x / max(||x||_2, eps)."""
    n = np.sqrt((x * x).sum(axis=-1, keepdims=True))
    return x / np.maximum(n, x.dtype.type(eps)), n


def quantizer_fwd(x, codebook):
    """train_titok.Quantizer.forward (train_titok.py:50-59 == train_vit_vqgan.py:50-59).
    indices from l2-normalised x and codebook; gathers RAW codebook rows."""
    xh, xn = _normalize(x)
    eh, _ = _normalize(codebook)
    flat = xh.reshape(-1, xh.shape[-1])
    # cdist -> argmin; squared distance has the same argmin as the euclidean distance
    d = (flat * flat).sum(-1, keepdims=True) + (eh * eh).sum(-1)[None, :] - 2.0 * (flat @ eh.T)
    idx = d.argmin(axis=-1).reshape(x.shape[:-1])
    c = codebook[idx]
    codebook_loss = ((c - xh) ** 2).mean()
    commitment_loss = 0.25 * ((c - xh) ** 2).mean()
    loss = codebook_loss + commitment_loss
    quantized = xh + (c - xh)  # straight-through value, two roundings as in the reference
    return quantized.astype(x.dtype), idx.astype(np.int64), loss.astype(x.dtype), (x, xh, xn, c, idx, codebook.shape)


def quantizer_bwd(dq, dloss, cache):
    """Gradients of (quantized, loss) wrt x and the codebook (SURVEY.md §8a formulas)."""
    x, xh, xn, c, idx, cshape = cache
    nel = x.size
    dxh = dq + dloss * 0.5 * (xh - c) / nel  # straight-through + commitment (2 * 0.25)
    dcode_rows = dloss * 2.0 * (c - xh) / nel  # codebook loss, on the raw rows
    denom = np.maximum(xn, x.dtype.type(1e-12))
    dx = (dxh - xh * (xh * dxh).sum(-1, keepdims=True)) / denom
    dC = np.zeros(cshape, dtype=x.dtype)
    np.add.at(dC, idx.reshape(-1), dcode_rows.reshape(-1, cshape[1]))
    return dx.astype(x.dtype), dC


def vector_quantizer_fwd(z, embedding, commitment_cost=0.25, use_l2_norm=True):
    """blocks.VectorQuantizer.forward (blocks.py:428-494), clustering_vq=False.
    z: [b,c,h,w]; returns z_q[b,c,h,w] and (quantizer_loss, commitment_loss, codebook_loss, indices[b,h,w])."""
    b, cdim, h, w = z.shape
    zt = np.transpose(z, (0, 2, 3, 1))
    flat = zt.reshape(-1, cdim)
    if use_l2_norm:
        flat_n, zn = _normalize(flat)
        emb, en = _normalize(embedding)
    else:
        flat_n, zn, emb, en = flat, None, embedding, None
    d = (flat_n ** 2).sum(axis=1, keepdims=True) + (emb ** 2).sum(axis=1)[None, :] - 2.0 * (flat_n @ emb.T)
    idx = d.argmin(axis=1)
    zq = emb[idx].reshape(zt.shape)  # get_codebook_entry: normalised rows when use_l2_norm
    zc = flat_n.reshape(zt.shape)
    commitment_loss = commitment_cost * ((zq - zc) ** 2).mean()
    codebook_loss = ((zq - zc) ** 2).mean()
    loss = commitment_loss + codebook_loss
    out = zc + (zq - zc)
    out = np.transpose(out, (0, 3, 1, 2))
    cache = (z, zc, zn, zq, idx, embedding, en, commitment_cost, use_l2_norm)
    return (out.astype(z.dtype), loss.astype(z.dtype), commitment_loss.astype(z.dtype),
            codebook_loss.astype(z.dtype), idx.reshape(b, h, w).astype(np.int64), cache)


def vector_quantizer_bwd(dout, dloss, cache):
    z, zc, zn, zq, idx, embedding, en, cc, use_l2 = cache
    b, cdim, h, w = z.shape
    nel = z.size
    dzc = np.transpose(dout, (0, 2, 3, 1)) + dloss * cc * 2.0 * (zc - zq) / nel
    dzq = dloss * 2.0 * (zq - zc) / nel
    if use_l2:
        denom = np.maximum(zn.reshape(zc.shape[:-1] + (1,)), z.dtype.type(1e-12))
        dzt = (dzc - zc * (zc * dzc).sum(-1, keepdims=True)) / denom
        rows = dzq.reshape(-1, cdim)
        c = zq.reshape(-1, cdim)
        en_rows = np.maximum(en[idx], z.dtype.type(1e-12))
        drows = (rows - c * (c * rows).sum(-1, keepdims=True)) / en_rows
    else:
        dzt = dzc
        drows = dzq.reshape(-1, cdim)
    dE = np.zeros_like(embedding)
    np.add.at(dE, idx, drows)
    return np.transpose(dzt, (0, 3, 1, 2)).astype(z.dtype), dE
