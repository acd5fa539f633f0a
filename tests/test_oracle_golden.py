"""Pins the numpy oracle (oracle/vit_oracle.py) against fixtures produced by the reference modules
themselves (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import vit_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _close(a, b, rtol=2e-4, atol=2e-5):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def _layer_params(g, prefix, i):
    k = f"{prefix}layers.{i}."
    return {
        "qkv_w": g[k + "multi_attn.qkv.weight"], "qkv_b": g[k + "multi_attn.qkv.bias"],
        "fc1_w": g[k + "mlp.0.weight"], "fc1_b": g[k + "mlp.0.bias"],
        "fc2_w": g[k + "mlp.2.weight"], "fc2_b": g[k + "mlp.2.bias"],
    }


@pytest.mark.parametrize("tag", ["a", "b"])
def test_transformer_fwd_bwd(golden_dir, tag):
    g = _load(golden_dir, "transformer.npz")
    L, h, d, N, B, causal = (int(v) for v in g[f"{tag}_cfg"])
    layers = [_layer_params(g, f"{tag}_w_", i) for i in range(L)]
    y, caches = O.transformer_fwd(g[f"{tag}_x"], layers, h, bool(causal))
    _close(y, g[f"{tag}_y"])
    dx, grads = O.transformer_bwd(g[f"{tag}_dy"], caches)
    _close(dx, g[f"{tag}_dx"])
    names = {"qkv_w": "multi_attn.qkv.weight", "qkv_b": "multi_attn.qkv.bias", "fc1_w": "mlp.0.weight",
             "fc1_b": "mlp.0.bias", "fc2_w": "mlp.2.weight", "fc2_b": "mlp.2.bias"}
    for i in range(L):
        for k, ref in names.items():
            _close(grads[i][k], g[f"{tag}_g_layers.{i}.{ref}"], rtol=5e-4, atol=5e-5)


def _vit_params(g, prefix, n_layers):
    return {
        "conv_w": g[prefix + "vit.patch_proj.weight"], "conv_b": g[prefix + "vit.patch_proj.bias"],
        "pos_emb": g[prefix + "vit.pos_emb.weight"], "extra_emb": g[prefix + "vit.extra_emb.weight"],
        "layers": [_layer_params(g, prefix + "vit.transformer.", i) for i in range(n_layers)],
        "head_w": g[prefix + "head.weight"], "head_b": g[prefix + "head.bias"],
    }


def test_vit_classifier_step_xs(golden_dir):
    g = _load(golden_dir, "vit.npz")
    P = _vit_params(g, "xs_w_", 2)
    tokens, _ = O.vit_fwd(g["xs_x"], P, n_heads=1)
    _close(tokens, g["xs_tokens"])
    loss, logits, grads = O.vit_classifier_loss_and_grads(g["xs_x"], g["xs_labels"], P, n_heads=1)
    _close(logits, g["xs_logits"])
    _close(loss, g["xs_loss"])
    _close(grads["conv_w"], g["xs_g_vit.patch_proj.weight"], rtol=1e-3, atol=1e-5)
    _close(grads["conv_b"], g["xs_g_vit.patch_proj.bias"], rtol=1e-3, atol=1e-5)
    _close(grads["pos_emb"], g["xs_g_vit.pos_emb.weight"], rtol=1e-3, atol=1e-5)
    _close(grads["extra_emb"], g["xs_g_vit.extra_emb.weight"], rtol=1e-3, atol=1e-5)
    _close(grads["head_w"], g["xs_g_head.weight"], rtol=1e-3, atol=1e-5)
    _close(grads["layers"][0]["qkv_w"], g["xs_g_vit.transformer.layers.0.multi_attn.qkv.weight"], rtol=1e-3, atol=1e-5)
    _close(grads["layers"][1]["fc2_w"], g["xs_g_vit.transformer.layers.1.mlp.2.weight"], rtol=1e-3, atol=1e-5)


def det_weights_np(shapes, seed, scale):
    """Same recipe as tests/golden/make_golden.py:det_weights (numpy Generator in state_dict order)."""
    rng = np.random.default_rng(seed)
    return {k: rng.standard_normal(s).astype(np.float32) * scale for k, s in shapes}


def vit_ti_shapes(n_classes=10, d=192, L=12, C=3, p=4, P=64, extra=1):
    shapes = [("vit.patch_proj.weight", (d, C, p, p)), ("vit.patch_proj.bias", (d,)),
              ("vit.pos_emb.weight", (P, d)), ("vit.extra_emb.weight", (extra, d))]
    for i in range(L):
        k = f"vit.transformer.layers.{i}."
        shapes += [(k + "multi_attn.qkv.weight", (3 * d, d)), (k + "multi_attn.qkv.bias", (3 * d,)),
                   (k + "mlp.0.weight", (4 * d, d)), (k + "mlp.0.bias", (4 * d,)),
                   (k + "mlp.2.weight", (d, 4 * d)), (k + "mlp.2.bias", (d,))]
    shapes += [("head.weight", (n_classes, d)), ("head.bias", (n_classes,))]
    return shapes


def test_vit_tiny_baseline_config1(golden_dir):
    """BASELINE.json configs[0]: ViT-Ti (192/12/3), patch 4, 32x32, batch 32, fwd + CE + bwd on CPU."""
    g = _load(golden_dir, "vit.npz")
    w = det_weights_np(vit_ti_shapes(), seed=1, scale=0.03)
    P = _vit_params({("ti_w_" + k): v for k, v in w.items()}, "ti_w_", 12)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((32, 3, 32, 32)).astype(np.float32)
    labels = rng.integers(0, 10, size=(32,))
    loss, logits, grads = O.vit_classifier_loss_and_grads(x, labels, P, n_heads=3)
    _close(logits, g["ti_logits"], rtol=1e-3, atol=1e-4)
    _close(loss, g["ti_loss"], rtol=1e-4)
    flat = {"vit.patch_proj.weight": grads["conv_w"], "vit.patch_proj.bias": grads["conv_b"],
            "vit.pos_emb.weight": grads["pos_emb"], "vit.extra_emb.weight": grads["extra_emb"],
            "head.weight": grads["head_w"], "head.bias": grads["head_b"]}
    m = {"qkv_w": "multi_attn.qkv.weight", "qkv_b": "multi_attn.qkv.bias", "fc1_w": "mlp.0.weight",
         "fc1_b": "mlp.0.bias", "fc2_w": "mlp.2.weight", "fc2_b": "mlp.2.bias"}
    for i in range(12):
        for k, ref in m.items():
            flat[f"vit.transformer.layers.{i}.{ref}"] = grads["layers"][i][k]
    for name, norm, sample in zip(g["ti_grad_names"], g["ti_grad_norms"], g["ti_grad_samples"]):
        got = flat[str(name)]
        np.testing.assert_allclose(np.linalg.norm(got.astype(np.float64)), norm, rtol=2e-3)
        np.testing.assert_allclose(got.reshape(-1)[:8], sample, rtol=5e-3, atol=1e-6)


@pytest.mark.parametrize("tag", ["default", "trained", "small"])
def test_quantizer(golden_dir, tag):
    g = _load(golden_dir, "quantizer.npz")
    q, idx, loss, cache = O.quantizer_fwd(g[f"{tag}_x"], g[f"{tag}_codebook"])
    assert np.array_equal(idx, g[f"{tag}_indices"]), "VQ indices must be bit-exact"
    _close(q, g[f"{tag}_quantized"], rtol=0, atol=3e-7)
    _close(loss, g[f"{tag}_loss"], rtol=1e-5, atol=0)
    dx, dC = O.quantizer_bwd(g[f"{tag}_dq"], float(g[f"{tag}_dloss"]), cache)
    _close(dx, g[f"{tag}_dx"], rtol=1e-4, atol=1e-6)
    _close(dC, g[f"{tag}_dC"], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("tag,l2", [("l2", True), ("plain", False)])
def test_vector_quantizer(golden_dir, tag, l2):
    g = _load(golden_dir, "vector_quantizer.npz")
    zq, loss, closs, cbloss, idx, cache = O.vector_quantizer_fwd(g[f"{tag}_z"], g[f"{tag}_embedding"], 0.25, l2)
    assert np.array_equal(idx, g[f"{tag}_indices"]), "VQ indices must be bit-exact"
    _close(zq, g[f"{tag}_zq"], rtol=0, atol=5e-7)
    _close(loss, g[f"{tag}_loss"], rtol=1e-5)
    _close(closs, g[f"{tag}_commitment_loss"], rtol=1e-5)
    _close(cbloss, g[f"{tag}_codebook_loss"], rtol=1e-5)
    dz, dE = O.vector_quantizer_bwd(g[f"{tag}_dz"], 2.0, cache)
    _close(dz, g[f"{tag}_dzin"], rtol=1e-4, atol=1e-6)
    _close(dE, g[f"{tag}_dE"], rtol=1e-4, atol=1e-7)


def test_residual_attention_block(golden_dir):
    g = _load(golden_dir, "resblock.npz")
    d, h, L, B = (int(v) for v in g["cfg"])
    p = {"ln1_w": g["w_ln_1.weight"], "ln1_b": g["w_ln_1.bias"], "in_w": g["w_attn.in_proj_weight"],
         "in_b": g["w_attn.in_proj_bias"], "out_w": g["w_attn.out_proj.weight"], "out_b": g["w_attn.out_proj.bias"],
         "ln2_w": g["w_ln_2.weight"], "ln2_b": g["w_ln_2.bias"], "fc_w": g["w_mlp.c_fc.weight"],
         "fc_b": g["w_mlp.c_fc.bias"], "proj_w": g["w_mlp.c_proj.weight"], "proj_b": g["w_mlp.c_proj.bias"]}
    y, cache = O.residual_attention_block_fwd(g["x"], p, h)
    _close(y, g["y"])
    dx, gr = O.residual_attention_block_bwd(g["dy"], cache)
    _close(dx, g["dx"], rtol=5e-4, atol=5e-5)
    ref = {"ln1_w": "ln_1.weight", "ln1_b": "ln_1.bias", "in_w": "attn.in_proj_weight", "in_b": "attn.in_proj_bias",
           "out_w": "attn.out_proj.weight", "out_b": "attn.out_proj.bias", "ln2_w": "ln_2.weight",
           "ln2_b": "ln_2.bias", "fc_w": "mlp.c_fc.weight", "fc_b": "mlp.c_fc.bias",
           "proj_w": "mlp.c_proj.weight", "proj_b": "mlp.c_proj.bias"}
    for k, r in ref.items():
        _close(gr[k], g[f"g_{r}"], rtol=1e-3, atol=5e-5)


# ---- the C oracle (bit-exact fp32 spec of the VQ lookup) against the same reference fixtures ----
from oracle import vq_oracle as VQ  # noqa: E402


@pytest.mark.parametrize("tag", ["default", "trained", "small"])
def test_c_oracle_quantizer(golden_dir, tag):
    g = _load(golden_dir, "quantizer.npz")
    q, idx, mse = VQ.vq_fwd(g[f"{tag}_x"], g[f"{tag}_codebook"], l2=True, gather_normalized=False)
    assert np.array_equal(idx, g[f"{tag}_indices"]), "C oracle indices must equal the reference's argmin"
    _close(q, g[f"{tag}_quantized"], rtol=0, atol=3e-7)
    _close(np.float32(1.25 * mse), g[f"{tag}_loss"], rtol=1e-5, atol=0)
    dL = float(g[f"{tag}_dloss"])
    dx, dC = VQ.vq_bwd(g[f"{tag}_x"], g[f"{tag}_codebook"], idx, g[f"{tag}_dq"], 0.25 * dL, dL)
    _close(dx, g[f"{tag}_dx"], rtol=1e-4, atol=1e-6)
    _close(dC, g[f"{tag}_dC"], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("tag,l2", [("l2", True), ("plain", False)])
def test_c_oracle_vector_quantizer(golden_dir, tag, l2):
    g = _load(golden_dir, "vector_quantizer.npz")
    q, idx, mse = VQ.vq_fwd(g[f"{tag}_z"], g[f"{tag}_embedding"], l2=l2, gather_normalized=l2, channels_first=True)
    assert np.array_equal(idx, g[f"{tag}_indices"])
    _close(q, g[f"{tag}_zq"], rtol=0, atol=5e-7)
    _close(np.float32(1.25 * mse), g[f"{tag}_loss"], rtol=1e-5)
    dz, dE = VQ.vq_bwd(g[f"{tag}_z"], g[f"{tag}_embedding"], idx, g[f"{tag}_dz"], 0.25 * 2.0, 2.0, l2=l2,
                       gather_normalized=l2, channels_first=True)
    _close(dz, g[f"{tag}_dzin"], rtol=1e-4, atol=1e-6)
    _close(dE, g[f"{tag}_dE"], rtol=1e-4, atol=1e-7)


def test_adamw_oracle(golden_dir):
    """oracle/adamw_oracle.py against trajectories of torch.optim.AdamW itself (tests/golden/make_golden_adamw.py)."""
    from oracle import adamw_oracle as A
    g = _load(golden_dir, "adamw.npz")
    b1, b2 = (float(v) for v in g["betas"])
    for i in range(int(g["n_tensors"])):
        p = g[f"p0_{i}"].copy()
        m = np.zeros_like(p)
        v = np.zeros_like(p)
        for s in range(int(g["steps"])):
            p, m, v, took = A.adamw_step(p, g[f"g{s}_{i}"], m, v, s + 1, float(g["lrs"][s]), b1, b2, float(g["eps"]),
                                         float(g["weight_decay"]))
            assert took
            # fp32 op-for-op restatement: a few ulp of the parameter scale (torch's CPU kernels contract some mul+add)
            np.testing.assert_allclose(p, g[f"p{s + 1}_{i}"], rtol=2e-6, atol=2e-8)
        m_ref, v_ref = g[f"m{int(g['steps'])}_{i}"], g[f"v{int(g['steps'])}_{i}"]
        np.testing.assert_allclose(m, m_ref, rtol=2e-6, atol=5e-7 * np.abs(m_ref).max())  # lerp cancels near zero
        np.testing.assert_allclose(v, v_ref, rtol=2e-6, atol=5e-7 * np.abs(v_ref).max())
    # GradScaler protocol: found_inf skips everything; grad_scale divides the gradient
    p0, g0 = g["p0_1"], g["g0_1"]
    z = np.zeros_like(p0)
    p1, m1, v1, took = A.adamw_step(p0, g0, z, z, 1, 1e-3, found_inf=1.0)
    assert not took and np.array_equal(p1, p0) and not m1.any() and not v1.any()
    pa, ma, va, _ = A.adamw_step(p0, g0 * np.float32(1024.0), z, z, 1, 1e-3, grad_scale=1024.0)
    pb, mb, vb, _ = A.adamw_step(p0, g0, z, z, 1, 1e-3)
    np.testing.assert_allclose(pa, pb, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_cross_entropy_oracle(golden_dir, tag):
    """oracle cross-entropy (with ignore_index) against torch.nn.functional.cross_entropy (tests/golden/make_golden_ce.py)."""
    g = _load(golden_dir, "cross_entropy.npz")
    loss, cache = O.cross_entropy_fwd(g[f"{tag}_x"], g[f"{tag}_y"])
    np.testing.assert_allclose(loss, g[f"{tag}_loss"], rtol=1e-5)
    np.testing.assert_allclose(O.cross_entropy_bwd(cache), g[f"{tag}_dx"], rtol=1e-4, atol=1e-7)


# ------------------------------------------------------------------------------------------------ §8f rows
def _np_weights(shapes_in_state_dict_order, seed, scale):
    """tests/golden/make_golden.py:det_weights: numpy Generator values in state_dict order (masks excluded by the caller)."""
    rng = np.random.default_rng(seed)
    return {k: rng.standard_normal(s).astype(np.float32) * scale for k, s in shapes_in_state_dict_order}


def _vit_shapes(prefix, d, L, C, p, n_pos, extra):
    out = [(prefix + "patch_proj.weight", (d, C, p, p)), (prefix + "patch_proj.bias", (d,)),
           (prefix + "pos_emb.weight", (n_pos, d)), (prefix + "extra_emb.weight", (extra, d))]
    for i in range(L):
        k = f"{prefix}transformer.layers.{i}."
        out += [(k + "multi_attn.qkv.weight", (3 * d, d)), (k + "multi_attn.qkv.bias", (3 * d,)),
                (k + "mlp.0.weight", (4 * d, d)), (k + "mlp.0.bias", (4 * d,)),
                (k + "mlp.2.weight", (d, 4 * d)), (k + "mlp.2.bias", (d,))]
    return out


def _vit_P(w, prefix, L):
    return {"conv_w": w[prefix + "patch_proj.weight"], "conv_b": w[prefix + "patch_proj.bias"],
            "pos_emb": w[prefix + "pos_emb.weight"], "extra_emb": w[prefix + "extra_emb.weight"],
            "layers": [{"qkv_w": w[f"{prefix}transformer.layers.{i}.multi_attn.qkv.weight"],
                        "qkv_b": w[f"{prefix}transformer.layers.{i}.multi_attn.qkv.bias"],
                        "fc1_w": w[f"{prefix}transformer.layers.{i}.mlp.0.weight"], "fc1_b": w[f"{prefix}transformer.layers.{i}.mlp.0.bias"],
                        "fc2_w": w[f"{prefix}transformer.layers.{i}.mlp.2.weight"], "fc2_b": w[f"{prefix}transformer.layers.{i}.mlp.2.bias"]}
                       for i in range(L)]}


def test_titok_encoder_decoder_oracle(golden_dir):
    """oracle titok_encoder_fwd / titok_decoder_fwd / token_linear_bwd / depatchify_bwd against the reference's
    train_titok.TiTokEncoder / TiTokDecoder (tests/golden/titok.npz, make_golden.py titok)."""
    g = _load(golden_dir, "titok.npz")
    img_size, p, latent, K, D = (int(v) for v in g["cfg"])
    d, L, P_ = 64, 2, (img_size // p) ** 2
    shapes = _vit_shapes("vit.", d, L, 3, p, P_, latent) + [("proj.weight", (D, d)), ("proj.bias", (D,))]
    assert [k for k, _ in shapes] == [str(k) for k in g["enc_keys"]]
    w = _np_weights(shapes, 41, 0.05)
    P = _vit_P(w, "vit.", L)
    P["proj_w"], P["proj_b"] = w["proj.weight"], w["proj.bias"]
    lat, tokens, c = O.titok_encoder_fwd(g["enc_x"], P, 1, latent)
    _close(lat, g["enc_lat"])
    _, dw, db = O.token_linear_bwd(g["enc_dlat"], c)
    _close(dw, g["enc_g_proj.weight"], rtol=5e-4, atol=5e-5)
    _close(db, g["enc_g_proj.bias"], rtol=5e-4, atol=5e-5)
    # decoder: the "image" is [B, d, latent, 1], patch 1, n_patches mask tokens prepended
    shapes = _vit_shapes("vit.", d, L, d, 1, latent, P_) + [("quant_proj.weight", (d, D)), ("quant_proj.bias", (d,)),
                                                             ("embd_proj.weight", (3 * p * p, d, 1, 1)), ("embd_proj.bias", (3 * p * p,))]
    assert [k for k, _ in shapes] == [str(k) for k in g["dec_keys"]]
    w = _np_weights(shapes, 42, 0.05)
    P = _vit_P(w, "vit.", L)
    P.update(quant_proj_w=w["quant_proj.weight"], quant_proj_b=w["quant_proj.bias"],
             embd_proj_w=w["embd_proj.weight"], embd_proj_b=w["embd_proj.bias"])
    img, tokens, c = O.titok_decoder_fwd(g["dec_z"], P, 1, img_size // p, p)
    _close(img, g["dec_img"])
    _, dw, db = O.depatchify_bwd(g["dec_dimg"], c)
    _close(dw, g["dec_g_embd_proj.weight"], rtol=5e-4, atol=5e-5)
    _close(db, g["dec_g_embd_proj.bias"], rtol=5e-4, atol=5e-5)


def _videogpt_P(seed, scale, keys):
    d, L, V, S = 64, 2, 32, 32
    shapes = [("tok_embed.weight", (V + 1, d)), ("pos_embed.weight", (S, d))]
    for i in range(L):
        k = f"transformer.layers.{i}."
        shapes += [(k + "multi_attn.mask", (S, S)), (k + "multi_attn.qkv.weight", (3 * d, d)), (k + "multi_attn.qkv.bias", (3 * d,)),
                   (k + "mlp.0.weight", (4 * d, d)), (k + "mlp.0.bias", (4 * d,)), (k + "mlp.2.weight", (d, 4 * d)), (k + "mlp.2.bias", (d,))]
    shapes += [("proj.weight", (V, d)), ("proj.bias", (V,))]
    assert [k for k, _ in shapes] == [str(k) for k in keys]
    w = _np_weights([(k, s) for k, s in shapes if not k.endswith("mask")], seed, scale)   # det_weights skips the masks
    P = {"tok_embed": w["tok_embed.weight"], "pos_embed": w["pos_embed.weight"], "proj_w": w["proj.weight"], "proj_b": w["proj.bias"],
         "layers": [{"qkv_w": w[f"transformer.layers.{i}.multi_attn.qkv.weight"], "qkv_b": w[f"transformer.layers.{i}.multi_attn.qkv.bias"],
                     "fc1_w": w[f"transformer.layers.{i}.mlp.0.weight"], "fc1_b": w[f"transformer.layers.{i}.mlp.0.bias"],
                     "fc2_w": w[f"transformer.layers.{i}.mlp.2.weight"], "fc2_b": w[f"transformer.layers.{i}.mlp.2.bias"]} for i in range(L)]}
    return P


def test_videogpt_oracle(golden_dir):
    """oracle videogpt_fwd / videogpt_bwd / videogpt_generate against train_videogpt.VideoGPT (tests/golden/videogpt.npz)."""
    g = _load(golden_dir, "videogpt.npz")
    P = _videogpt_P(51, 0.1, g["keys"])
    logits, loss, caches = O.videogpt_fwd(g["x"], P, 1, 32)
    _close(logits, g["logits"], rtol=1e-3, atol=1e-4)
    _close(loss, g["loss"], rtol=1e-4)
    grads = O.videogpt_bwd(caches)
    for name, key in (("tok_embed", "tok_embed.weight"), ("pos_embed", "pos_embed.weight"), ("proj_w", "proj.weight"), ("proj_b", "proj.bias")):
        _close(grads[name], g[f"g_{key}"], rtol=2e-3, atol=2e-5)
    for i in range(2):
        ref = float(g[f"gn_transformer.layers.{i}.mlp.0.weight"])
        assert abs(np.linalg.norm(grads["layers"][i]["fc1_w"].astype(np.float64)) - ref) < 2e-3 * ref
    # greedy generation: identical tokens up to the first near-tie of the reference (fp32 both sides: margin 1e-3 is ample)
    P = _videogpt_P(53, 0.3, g["keys"])
    toks, margins = O.videogpt_generate(g["prompt"], 12, P, 1, 32)
    ref, ref_m = g["generated"], g["margins"]
    for b in range(ref.shape[0]):
        for j in range(12):
            if ref_m[b, j] < 1e-2:
                break
            assert toks[b, 5 + j] == ref[b, 5 + j], (b, j)
            assert abs(margins[b, j] - ref_m[b, j]) < 2e-2 * max(1.0, ref_m[b, j])


# ------------------------------------------------------------------------------------------------ blocks.py encoder / decoder
def blocks_titok_weights(g, tag):
    """The weights tests/golden/make_golden.py:gen_blocks_titok gave the reference blocks.TiTokEncoder / TiTokDecoder
    (det_weights recipe: numpy Generator in state_dict order, scale 0.03, LayerNorm weights recentred at 1)."""
    rng = np.random.default_rng({"enc": 71, "dec": 72}[tag])
    w = {}
    for k, shp in zip(g[f"{tag}_keys"], g[f"{tag}_shapes"]):
        k = str(k)
        shape = tuple(int(v) for v in str(shp).split(",")) if str(shp) else ()
        a = rng.standard_normal(shape).astype(np.float32) * 0.03
        if k.endswith("ln_1.weight") or k.endswith("ln_2.weight") or k in ("ln_pre.weight", "ln_post.weight"):
            a = a + 1.0
        w[k] = a
    return w


def _blocks_P(w, n_layers):
    P = {"cls": w["class_embedding"], "pos": w["positional_embedding"], "latent_pos": w["latent_token_positional_embedding"],
         "ln_pre_w": w["ln_pre.weight"], "ln_pre_b": w["ln_pre.bias"], "ln_post_w": w["ln_post.weight"], "ln_post_b": w["ln_post.bias"]}
    P["blocks"] = [{"ln1_w": w[f"transformer.{i}.ln_1.weight"], "ln1_b": w[f"transformer.{i}.ln_1.bias"],
                    "in_w": w[f"transformer.{i}.attn.in_proj_weight"], "in_b": w[f"transformer.{i}.attn.in_proj_bias"],
                    "out_w": w[f"transformer.{i}.attn.out_proj.weight"], "out_b": w[f"transformer.{i}.attn.out_proj.bias"],
                    "ln2_w": w[f"transformer.{i}.ln_2.weight"], "ln2_b": w[f"transformer.{i}.ln_2.bias"],
                    "fc_w": w[f"transformer.{i}.mlp.c_fc.weight"], "fc_b": w[f"transformer.{i}.mlp.c_fc.bias"],
                    "proj_w": w[f"transformer.{i}.mlp.c_proj.weight"], "proj_b": w[f"transformer.{i}.mlp.c_proj.bias"]}
                   for i in range(n_layers)]
    return P


def test_blocks_titok_encoder_decoder_oracle(golden_dir):
    """oracle blocks_titok_encoder_fwd / _bwd and blocks_titok_decoder_fwd against the reference's own blocks.TiTokEncoder /
    TiTokDecoder (tests/golden/blocks_titok.npz, make_golden.py blocks_titok): the token-sequence assembly (class / mask
    tokens, both positional tables, latent tokens), ln_pre, 8 ResidualAttentionBlocks, ln_post, conv_out / ffn."""
    g = _load(golden_dir, "blocks_titok.npz")
    w = blocks_titok_weights(g, "enc")
    P = _blocks_P(w, 8)
    P.update(patch_w=w["patch_embed.weight"], patch_b=w["patch_embed.bias"], conv_out_w=w["conv_out.weight"], conv_out_b=w["conv_out.bias"])
    z, cache = O.blocks_titok_encoder_fwd(g["enc_x"].astype(np.float64), g["enc_latent_tokens"].astype(np.float64),
                                          {k: (v.astype(np.float64) if isinstance(v, np.ndarray) else [{a: b.astype(np.float64) for a, b in d.items()} for d in v])
                                           for k, v in P.items()}, 8)
    _close(z, g["enc_z"], rtol=1e-3, atol=1e-4)
    gr = O.blocks_titok_encoder_bwd(g["enc_dz"].astype(np.float64), cache)
    _close(gr["latent_tokens"], g["enc_dlatent_tokens"], rtol=2e-3, atol=2e-4)
    for name, key in (("class_embedding", "cls"), ("positional_embedding", "pos"), ("latent_token_positional_embedding", "latent_pos"),
                      ("ln_pre.weight", "ln_pre_w"), ("ln_pre.bias", "ln_pre_b"), ("ln_post.weight", "ln_post_w"),
                      ("ln_post.bias", "ln_post_b"), ("conv_out.bias", "conv_out_b"), ("patch_embed.bias", "patch_b")):
        _close(gr[key].reshape(g[f"enc_g_{name}"].shape), g[f"enc_g_{name}"], rtol=2e-3, atol=2e-4)
    norms = dict(zip((str(n) for n in g["enc_grad_names"]), g["enc_grad_norms"]))
    assert abs(np.linalg.norm(gr["patch_w"]) - norms["patch_embed.weight"]) < 2e-3 * norms["patch_embed.weight"]
    for i in (0, 7):
        for ours, theirs in (("ln1_w", "ln_1.weight"), ("in_w", "attn.in_proj_weight"), ("fc_w", "mlp.c_fc.weight"), ("proj_b", "mlp.c_proj.bias")):
            n = norms[f"transformer.{i}.{theirs}"]
            assert abs(np.linalg.norm(gr["blocks"][i][ours]) - n) < 2e-3 * n, (i, ours)
    # decoder (up to the 3x3 conv_out, which is outside the path)
    w = blocks_titok_weights(g, "dec")
    P = _blocks_P(w, 8)
    P.update(embed_w=w["decoder_embed.weight"], embed_b=w["decoder_embed.bias"], mask=w["mask_token"], ffn_w=w["ffn.0.weight"], ffn_b=w["ffn.0.bias"])
    img, _ = O.blocks_titok_decoder_fwd(g["dec_zq"], P, 8, 4, 16)
    _close(img, g["dec_img_before_conv_out"], rtol=2e-3, atol=2e-4)


def test_affine_fold_identity():
    """The algebra behind csrc/affine_fold.cu, in fp64: folding an affine LayerNorm into the next Linear changes neither the
    output nor (after unfolding) any gradient."""
    rng = np.random.default_rng(3)
    M, K, N = 37, 24, 40
    x = rng.standard_normal((M, K)); W = rng.standard_normal((N, K)); b = rng.standard_normal(N)
    gamma = 1 + 0.3 * rng.standard_normal(K); beta = 0.2 * rng.standard_normal(K); dy = rng.standard_normal((M, N))
    a, c = O.layer_norm_fwd(x, gamma, beta)
    y = O.linear_fwd(a, W, b)
    xhat, _ = O.layer_norm_fwd(x)
    Wf, bf = O.affine_fold(W, b, gamma, beta)
    np.testing.assert_allclose(O.linear_fwd(xhat, Wf, bf), y, rtol=1e-12, atol=1e-12)
    da, dW, db = O.linear_bwd(dy, a, W)
    _, dgamma, dbeta = O.layer_norm_bwd(da, c)
    _, dWf, dbf = O.linear_bwd(dy, xhat, Wf)
    dW2, dgamma2, dbeta2 = O.affine_unfold_grads(dWf, dbf, W, gamma, beta)
    np.testing.assert_allclose(dW2, dW, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(dgamma2, dgamma, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(dbeta2, dbeta, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(dbf, db, rtol=1e-12)
