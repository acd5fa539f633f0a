"""Drop-in nn.Modules with the reference's constructors, forward signatures, attributes and state_dict keys.

    transformer.py:5-59   TransformerConfig, Attention, TransformerLayer, Transformer, S, B, L, transformer_configs
    train_vit.py:16-53    ViTConfig, ViT, ViTClassifier
    train_titok.py:45-59  Quantizer
    blocks.py:32-70       ResidualAttentionBlock
    blocks.py:405-505     VectorQuantizer

Parameters stay ordinary fp32 nn.Parameters (AdamW / GradScaler / clip_grad_norm_ / torch.save work unchanged);
all arithmetic on the hot path runs in the hand-written sm_100a kernels behind the C ABI.  There is no CPU or
eager-PyTorch fallback: CPU tensors raise.
"""
from collections import OrderedDict
from dataclasses import dataclass

import torch
import torch.nn as nn

from . import functional as Fn


# ------------------------------------------------------------------------------------------------ transformer.py
@dataclass
class TransformerConfig:
    n_layers: int
    n_heads: int
    n_embd: int
    block_size: int
    causal: bool = False
    dropout: float = 0.0

    def __post_init__(self):
        self.head_dim = self.n_embd // self.n_heads


def _copy_config(module, config):
    if "causal" not in config.__dict__:
        config.causal = False  # same backwards-compat hack as transformer.py:19
    for k, v in config.__dict__.items():
        setattr(module, k, v)


def _check_supported(module):
    if module.head_dim != 64:
        raise NotImplementedError(f"b200vit attention kernels are built for head_dim 64 (got {module.head_dim}); "
                                  "every shipped reference config uses 64 (transformer.py:56-58)")
    if not (0.0 <= float(module.dropout) < 1.0):
        raise ValueError(f"dropout must be in [0, 1), got {module.dropout}")


def _dropout_pair(module):
    """(p_attention, p_mlp): SDPA's dropout_p is applied whether or not the module is training (transformer.py:28
    passes self.dropout unconditionally); nn.Dropout after mlp[2] (transformer.py:40) only in training mode."""
    p = float(module.dropout)
    return (p, p if module.training else 0.0)


class Attention(nn.Module):
    """transformer.Attention (transformer.py:16-29): fused QKV Linear + SDPA, no output projection."""

    def __init__(self, config: TransformerConfig):
        super().__init__()
        _copy_config(self, config)
        self.qkv = nn.Linear(self.n_embd, self.n_embd * 3)
        if self.causal:
            # kept only so that state_dict() carries the same `mask` key/shape as the reference
            mask = torch.triu(torch.ones(config.block_size, config.block_size), diagonal=1)
            mask = mask.masked_fill(mask == 1, float("-inf"))
            self.register_buffer("mask", mask)
        _check_supported(self)

    def forward(self, x):
        # like the reference (transformer.py:28) the attention dropout is active regardless of self.training
        return Fn.AttentionFn.apply(x, self.qkv.weight, self.qkv.bias, self.n_heads, bool(self.causal), float(self.dropout))


class TransformerLayer(nn.Module):
    """transformer.TransformerLayer (transformer.py:31-45)."""

    def __init__(self, config: TransformerConfig):
        super().__init__()
        _copy_config(self, config)
        self.multi_attn = Attention(config)
        self.mlp = nn.Sequential(
            nn.Linear(self.n_embd, 4 * self.n_embd),
            nn.GELU(),
            nn.Linear(4 * self.n_embd, self.n_embd),
            nn.Dropout(self.dropout),
        )

    def _params(self):
        return (self.multi_attn.qkv.weight, self.multi_attn.qkv.bias, self.mlp[0].weight, self.mlp[0].bias,
                self.mlp[2].weight, self.mlp[2].bias)

    def forward(self, x):
        return Fn.TransformerStackFn.apply(x, self.n_heads, bool(self.causal), _dropout_pair(self), *self._params())


class Transformer(nn.Module):
    """transformer.Transformer (transformer.py:47-54): layer stack, no final norm."""

    def __init__(self, config: TransformerConfig):
        super().__init__()
        _copy_config(self, config)
        self.layers = nn.ModuleList([TransformerLayer(config) for _ in range(config.n_layers)])

    def forward(self, x):
        params = []
        for layer in self.layers:
            params.extend(layer._params())
        return Fn.TransformerStackFn.apply(x, self.n_heads, bool(self.causal), _dropout_pair(self), *params)


def S(**kwargs): return TransformerConfig(n_layers=6, n_heads=8, n_embd=512, **kwargs)
def B(**kwargs): return TransformerConfig(n_layers=12, n_heads=12, n_embd=768, **kwargs)
def L(**kwargs): return TransformerConfig(n_layers=24, n_heads=16, n_embd=1024, **kwargs)


transformer_configs = {"S": S, "B": B, "L": L}


# ------------------------------------------------------------------------------------------------ train_vit.py
@dataclass
class ViTConfig:
    image_size: int
    in_channels: int
    patch_size: int
    transformer: str
    extra_tokens: int
    dropout: float

    def __post_init__(self):
        self.n_patches = (self.image_size // self.patch_size) ** 2
        self.patch_dim = 3 * self.patch_size ** 2
        self.trans_config = transformer_configs[self.transformer](block_size=self.n_patches + self.extra_tokens,
                                                                   dropout=self.dropout)


class ViT(nn.Module):
    """train_vit.ViT (train_vit.py:30-45): patchify conv + pos_emb + prepended extra tokens + Transformer."""

    def __init__(self, args: ViTConfig):
        super().__init__()
        self.config = args
        self.patch_proj = nn.Conv2d(in_channels=args.in_channels, out_channels=args.trans_config.n_embd,
                                    kernel_size=args.patch_size, stride=args.patch_size)
        self.pos_emb = nn.Embedding(args.n_patches, args.trans_config.n_embd)
        self.extra_emb = nn.Embedding(args.extra_tokens, args.trans_config.n_embd)
        self.transformer = Transformer(args.trans_config)

    def forward(self, x):
        extra = self.extra_emb.weight if self.config.extra_tokens > 0 else None
        emb = Fn.PatchEmbedFn.apply(x, self.patch_proj.weight, self.patch_proj.bias,
                                    self.pos_emb.weight[: self.config.n_patches], extra, self.config.patch_size)
        return self.transformer(emb)


class ViTClassifier(nn.Module):
    """train_vit.ViTClassifier (train_vit.py:47-53): head(vit(x)[:, 0]).  `head` is an ordinary nn.Linear parameter
    container (same state_dict keys); the token gather, the tcgen05 GEMM and the backward scatter run in TokenLinearFn.
    Like nn.Linear under autocast the logits are bf16 when autocast is on and fp32 otherwise."""

    def __init__(self, vit_config: ViTConfig, num_classes=1000):
        super().__init__()
        self.vit = ViT(vit_config)
        self.head = nn.Linear(vit_config.trans_config.n_embd, num_classes)

    def forward(self, x):
        return Fn.TokenLinearFn.apply(self.vit(x), self.head.weight, self.head.bias, 0, 1, torch.is_autocast_enabled())


class CrossEntropyLoss(nn.Module):
    """nn.CrossEntropyLoss() as the reference scripts build it (train_vit.py:81, mean reduction, no class weights, no label
    smoothing) over the fused kernels of csrc/head_ce.cu; logits [R, C] in bf16 or fp32, int64 class labels [R].
    b200vit.launch installs it as torch.nn.CrossEntropyLoss; other configurations raise rather than fall back."""

    def __init__(self, weight=None, size_average=None, ignore_index=-100, reduce=None, reduction="mean", label_smoothing=0.0):
        super().__init__()
        if weight is not None or reduction != "mean" or label_smoothing != 0.0 or size_average is not None or reduce is not None:
            raise NotImplementedError("b200vit.CrossEntropyLoss implements reduction='mean' without class weights or label "
                                      "smoothing (what the reference scripts use)")
        self.ignore_index = ignore_index
        self.reduction = reduction

    def forward(self, input, target):
        if input.dim() != 2 or target.dtype != torch.int64 or target.shape != input.shape[:1]:
            raise NotImplementedError("b200vit.CrossEntropyLoss takes logits [R, C] and int64 class indices [R] "
                                      "(train_vit.py:102, train_videogpt.py:54)")
        return Fn.CrossEntropyFn.apply(input, target, self.ignore_index)


class PatchConv2d(nn.Conv2d):
    """nn.Conv2d whose patchify-shaped instances (kernel_size == stride, no padding / dilation / groups: the patch
    embedding of blocks.TiTokEncoder blocks.py:235-237, the 1x1 convolutions of the decoders blocks.py:329-333 /
    train_titok.py:67) run as im2col + tcgen05 GEMM on CUDA under autocast; every other use falls through to
    nn.Conv2d unchanged.  Same parameters and state_dict keys.  b200vit.launch installs it as torch.nn.Conv2d."""

    def _patchify_shaped(self, x):
        k, s = self.kernel_size, self.stride
        return (x.is_cuda and x.dim() == 4 and torch.is_autocast_enabled() and k[0] == k[1] == s[0] == s[1]
                and self.padding == (0, 0) and self.dilation == (1, 1) and self.groups == 1
                and self.padding_mode == "zeros" and x.shape[2] % k[0] == 0 and x.shape[3] % k[0] == 0
                and (self.in_channels * k[0] * k[0]) % 8 == 0 and self.out_channels % 8 == 0
                and (k[0] == 1 or x.shape[3] % 4 == 0))

    def forward(self, x):
        if self._patchify_shaped(x):
            return Fn.PatchConvFn.apply(x, self.weight, self.bias, self.kernel_size[0])
        return super().forward(x)


# ------------------------------------------------------------------------------------------------ train_titok.py
class Quantizer(nn.Module):
    """train_titok.Quantizer (train_titok.py:45-59 == train_vit_vqgan.py:45-59): indices from l2-normalised
    latents/codes, RAW codebook rows gathered, codebook + 0.25 commitment loss, straight-through output."""

    def __init__(self, titok_config):
        super().__init__()
        self.codebook = nn.Embedding(titok_config.codebook_size, titok_config.latent_dim)
        self.codebook.weight.data.uniform_(-1.0 / titok_config.codebook_size, 1.0 / titok_config.codebook_size)

    def forward(self, x):
        quantized, indices, _mse, _commit, total = Fn.VQFn.apply(x, self.codebook.weight, True, False, False, 0.25)
        return quantized, indices.view(x.shape[:-1]), total


class TiTokEncoder(nn.Module):
    """train_titok.TiTokEncoder (train_titok.py:34-43): proj(vit(x)[:, :latent_tokens]); the token slice is gathered
    straight into the GEMM operand (TokenLinearFn).  `titok_config` is the script's own TiTokConfig."""

    def __init__(self, titok_config):
        super().__init__()
        self.latent_tokens = titok_config.latent_tokens
        self.vit = ViT(titok_config.enc_vit_config)
        self.proj = nn.Linear(titok_config.n_embd, titok_config.latent_dim)

    def forward(self, x):
        y = Fn.TokenLinearFn.apply(self.vit(x), self.proj.weight, self.proj.bias, 0, self.latent_tokens,
                                   torch.is_autocast_enabled())
        return y.reshape(x.shape[0], self.latent_tokens, -1)


class TiTokDecoder(nn.Module):
    """train_titok.TiTokDecoder (train_titok.py:61-76): quant_proj, the decoder ViT over the latent "image" [B, d, L, 1]
    with n_patches mask tokens prepended, and the de-patchify tail (1x1 conv + pixel shuffle) as one GEMM whose epilogue
    writes the image (DepatchifyFn).  Same parameters / state_dict keys as the reference (`embd_proj` stays an nn.Conv2d)."""

    def __init__(self, titok_config):
        super().__init__()
        self.config = titok_config
        self.vit = ViT(titok_config.dec_vit_config)
        self.quant_proj = nn.Linear(titok_config.latent_dim, titok_config.n_embd)
        self.embd_proj = nn.Conv2d(titok_config.n_embd, 3 * titok_config.patch_size ** 2, kernel_size=1)

    def forward(self, z):
        z = Fn.LinearFn.apply(z, self.quant_proj.weight, self.quant_proj.bias, torch.is_autocast_enabled())
        z = z.permute(0, 2, 1).unsqueeze(-1)                                  # 'b h c -> b c h 1'
        tokens = self.vit(z)
        pd = self.config.patch_dim
        return Fn.DepatchifyFn.apply(tokens, self.embd_proj.weight, self.embd_proj.bias, pd, pd, self.config.patch_size)


class TiTok(nn.Module):
    """train_titok.TiTok (train_titok.py:78-92)."""

    def __init__(self, titok_config):
        super().__init__()
        self.config = titok_config
        self.enc = TiTokEncoder(titok_config)
        self.quant = Quantizer(titok_config)
        self.dec = TiTokDecoder(titok_config)

    def encode(self, z): return self.quant(self.enc(z))[1]
    def decode(self, z_quant): return self.dec(z_quant)
    def decode_indices(self, indices): return self.dec(self.quant.codebook(indices))

    def forward(self, x):
        latent_embs = self.enc(x)
        quantized, indices, quantize_loss = self.quant(latent_embs)
        image_recon = self.dec(quantized)
        return image_recon, indices, quantize_loss


# ------------------------------------------------------------------------------------------------ train_videogpt.py
class VideoGPT(nn.Module):
    """train_videogpt.VideoGPT (train_videogpt.py:38-69): token + positional embedding, causal Transformer, vocabulary
    projection, cross-entropy -- and generate() with a KV cache instead of the reference's full re-computation per token
    (same greedy tokens).  `config` is the script's own VideoGPTConfig (codebook_size, n_embd, max_tokens, trans_config,
    frame_size)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.tok_embed = nn.Embedding(config.codebook_size + 1, config.n_embd)
        self.pos_embed = nn.Embedding(config.max_tokens, config.n_embd)
        self.transformer = Transformer(config.trans_config)
        self.proj = nn.Linear(config.n_embd, config.codebook_size)

    def _layers(self):
        return [Fn.LayerParams(*layer._params()) for layer in self.transformer.layers]

    def forward(self, x):
        B, T, N = x.shape
        S = T * N
        y = x.reshape(B, S)
        sos = torch.full((B, 1), self.config.codebook_size, device=x.device, dtype=torch.long)
        inp = torch.cat([sos, y[:, :-1]], dim=-1)
        h = Fn.EmbedFn.apply(inp, self.tok_embed.weight, self.pos_embed.weight[:S])
        h = self.transformer(h)
        logits = Fn.TokenLinearFn.apply(h, self.proj.weight, self.proj.bias, 0, S, torch.is_autocast_enabled())
        loss = Fn.CrossEntropyFn.apply(logits, y.reshape(-1), -100)
        return logits.reshape(B, S, -1), loss

    @torch.no_grad()
    def generate(self, tokens, n=1):
        """Greedy continuation of tokens [B, T0] by n tokens (train_videogpt.py:56-65), KV-cached: one prefill over
        [sos, tokens], then one single-token step per generated token."""
        if n <= 0:
            return tokens
        if float(self.transformer.dropout) > 0.0:
            raise NotImplementedError("generate() with dropout > 0 is stochastic in the reference (SDPA dropout_p is applied "
                                      "in eval mode, transformer.py:28); the KV-cached path implements dropout = 0")
        B, T0 = tokens.shape
        total = T0 + n            # the last step of the reference embeds T0 + n positions
        if total > self.pos_embed.weight.shape[0]:
            raise ValueError(f"generate: {total} positions exceed max_tokens={self.pos_embed.weight.shape[0]}")
        layers, H = self._layers(), self.transformer.n_heads
        wproj, bproj = Fn.bf16_of(self.proj.weight), Fn._f32c(self.proj.bias)

        def next_token(h_last):   # h_last [B, d] fp32 -> greedy token [B, 1]
            logits = Fn.ops.gemm_bias_f32(Fn.ops.cast_bf16(h_last.contiguous()), wproj, bproj)
            return torch.argmax(logits, dim=-1, keepdim=True)

        sos = torch.full((B, 1), self.config.codebook_size, device=tokens.device, dtype=torch.long)
        inp = torch.cat([sos, tokens], dim=-1)                                     # [B, T0 + 1]
        tok_w, pos_w = Fn._f32c(self.tok_embed.weight), Fn._f32c(self.pos_embed.weight)
        h = Fn.ops.embed_fwd(inp.contiguous(), tok_w, pos_w, 0)
        h, caches = Fn.stack_prefill(h, layers, H, total)
        out = torch.empty(B, total, device=tokens.device, dtype=torch.long)
        out[:, :T0] = tokens
        new = next_token(h[:, -1]).contiguous()                                    # static buffer: the token fed to the next step
        out[:, T0] = new.view(-1)
        if n == 1:
            return out
        pos_dev = torch.full((1,), T0 + 1, device=tokens.device, dtype=torch.int32)  # position of the token generated last

        def step():   # embeds `new` at *pos_dev, runs the cached stack, leaves the next greedy token in `new`, advances
            x = Fn.ops.embed_fwd(new, tok_w, pos_w, 0, pos0_dev=pos_dev)
            h1 = Fn.stack_decode_step(x.view(B, -1), layers, caches, pos_dev)
            new.copy_(next_token(h1))
            Fn.ops.advance_counter(pos_dev, 1)

        step()                                                                      # j = 1, eager (also the graph's warm-up)
        out[:, T0 + 1] = new.view(-1)
        graph = None
        if n - 2 >= self.GRAPH_MIN_STEPS:
            # the step has no host-side dependence on the position: capture it once, replay it per token
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
        for j in range(2, n):
            if graph is not None:
                graph.replay()
            else:
                step()
            out[:, T0 + j] = new.view(-1)
        return out

    GRAPH_MIN_STEPS = 8   # below this many remaining tokens capturing a graph does not pay

    def generate_frames(self, video_tokens, n=1):
        tokens = video_tokens.reshape(video_tokens.shape[0], -1)
        return self.generate(tokens, n * self.config.frame_size)


# ------------------------------------------------------------------------------------------------ blocks.py
class ResidualAttentionBlock(nn.Module):
    """blocks.ResidualAttentionBlock (blocks.py:32-70); x is sequence-first [L, B, d]."""

    def __init__(self, d_model, n_head, mlp_ratio=4.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        if act_layer is not nn.GELU or norm_layer is not nn.LayerNorm:
            raise NotImplementedError("b200vit ResidualAttentionBlock implements nn.GELU + nn.LayerNorm only")
        if d_model // n_head != 64:
            raise NotImplementedError("b200vit attention kernels are built for head_dim 64")
        self.n_head = n_head
        self.ln_1 = norm_layer(d_model)
        self.attn = nn.MultiheadAttention(d_model, n_head)  # parameter container: in_proj_*, out_proj.*
        self.mlp_ratio = mlp_ratio
        if mlp_ratio > 0:
            self.ln_2 = norm_layer(d_model)
            mlp_width = int(d_model * mlp_ratio)
            self.mlp = nn.Sequential(OrderedDict([
                ("c_fc", nn.Linear(d_model, mlp_width)),
                ("gelu", act_layer()),
                ("c_proj", nn.Linear(mlp_width, d_model)),
            ]))

    def _params(self):
        p = (self.ln_1.weight, self.ln_1.bias, self.attn.in_proj_weight, self.attn.in_proj_bias,
             self.attn.out_proj.weight, self.attn.out_proj.bias)
        if self.mlp_ratio > 0:
            p += (self.ln_2.weight, self.ln_2.bias, self.mlp.c_fc.weight, self.mlp.c_fc.bias,
                  self.mlp.c_proj.weight, self.mlp.c_proj.bias)
        return p

    def forward(self, x):
        return Fn.ResidualAttentionStackFn.apply(x, self.n_head, self.mlp_ratio > 0, False, *self._params())


class _UViTAttention(nn.Module):
    """Parameter container with the names of blocks.Attention (blocks.py:84-93): qkv (bias optional), proj."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        if qk_scale is not None or attn_drop != 0.0 or proj_drop != 0.0:
            raise NotImplementedError("b200vit UViTBlock implements qk_scale=None and zero attention / projection dropout")
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)


class _UViTMlp(nn.Module):
    """Parameter container with the names of blocks.Mlp (blocks.py:157-171): fc1, act, fc2, drop."""

    def __init__(self, in_features, hidden_features, drop=0.0):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, in_features)
        self.drop = nn.Dropout(drop)


class UViTBlock(nn.Module):
    """blocks.UViTBlock (blocks.py:174-201): batch-first pre-norm block x + attn(norm1(x)); x + mlp(norm2(x)) with an
    optional skip_linear over cat([x, skip]).  Same parameter names / state_dict keys (norm1, attn.qkv, attn.proj, norm2,
    mlp.fc1, mlp.fc2, skip_linear); runs on the kernels of ResidualAttentionBlock with the batch-first attention layout.
    Dropout, drop-path and activation checkpointing are not implemented (the reference never instantiates this block)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, skip=False, use_checkpoint=False):
        super().__init__()
        if act_layer is not nn.GELU or norm_layer is not nn.LayerNorm:
            raise NotImplementedError("b200vit UViTBlock implements nn.GELU + nn.LayerNorm only")
        if drop != 0.0 or drop_path != 0.0 or use_checkpoint:
            raise NotImplementedError("b200vit UViTBlock implements drop = drop_path = 0 and use_checkpoint=False")
        if dim // num_heads != 64:
            raise NotImplementedError("b200vit attention kernels are built for head_dim 64")
        self.norm1 = norm_layer(dim)
        self.attn = _UViTAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                                   proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = _UViTMlp(dim, int(dim * mlp_ratio), drop=drop)
        self.skip_linear = nn.Linear(2 * dim, dim) if skip else None
        self.use_checkpoint = use_checkpoint

    def forward(self, x, skip=None):
        if self.skip_linear is not None:
            x = Fn.LinearFn.apply(torch.cat([x, skip], dim=-1), self.skip_linear.weight, self.skip_linear.bias, False)
        return Fn.ResidualAttentionStackFn.apply(x, self.attn.num_heads, True, True, self.norm1.weight, self.norm1.bias,
                                                 self.attn.qkv.weight, self.attn.qkv.bias, self.attn.proj.weight,
                                                 self.attn.proj.bias, self.norm2.weight, self.norm2.bias,
                                                 self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias)


_BLOCKS_SIZES = {"small": (512, 8, 8), "base": (768, 12, 12), "large": (1024, 24, 16)}   # width, layers, heads (blocks.py:219-233)


def _run_rab_stack(blocks, x, n_heads):
    """x [B, N, d] fp32 (batch-first: the drop-ins never materialise the reference's NLD -> LND permute, blocks.py:270,273)."""
    params = []
    for blk in blocks:
        params.extend(blk._params())
    return Fn.ResidualAttentionStackFn.apply(x, n_heads, True, True, *params)


class BlocksTiTokEncoder(nn.Module):
    """blocks.TiTokEncoder (blocks.py:208-282) -- exported as `blocks.TiTokEncoder` by shim/blocks.py.  Same constructor
    (`config` with image_size, patch_size, transformer in {small, base, large}, latent_tokens, latent_dim), same parameters
    and state_dict keys, same forward(pixel_values, latent_tokens) -> [B, latent_dim, 1, latent_tokens].

    The whole front end of blocks.py:257-270 -- patch_embed conv, reshape / permute, class-token cat, positional add,
    latent tokens + their positions, cat -- is one im2col + one tcgen05 GEMM whose epilogue writes the patch rows of the
    [B, 1 + grid^2 + latent, d] sequence and one pass for the broadcast rows (TokensAssembleFn); ln_pre, the block stack (one
    autograd node, batch-first) and ln_post on the latent tokens only follow; conv_out (1x1) is a skinny GEMM."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.image_size = config.image_size
        self.patch_size = config.patch_size
        self.grid_size = self.image_size // self.patch_size
        self.model_size = config.transformer
        self.num_latent_tokens = config.latent_tokens
        self.token_size = config.latent_dim
        self.width, self.num_layers, self.num_heads = _BLOCKS_SIZES[self.model_size]
        self.patch_embed = nn.Conv2d(in_channels=3, out_channels=self.width, kernel_size=self.patch_size,
                                     stride=self.patch_size, bias=True)
        scale = self.width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(1, self.width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(self.grid_size ** 2 + 1, self.width))
        self.latent_token_positional_embedding = nn.Parameter(scale * torch.randn(self.num_latent_tokens, self.width))
        self.ln_pre = nn.LayerNorm(self.width)
        self.transformer = nn.ModuleList([ResidualAttentionBlock(self.width, self.num_heads, mlp_ratio=4.0)
                                          for _ in range(self.num_layers)])
        self.ln_post = nn.LayerNorm(self.width)
        self.conv_out = nn.Conv2d(self.width, self.token_size, kernel_size=1, bias=True)

    def forward(self, pixel_values, latent_tokens):
        B, P, L = pixel_values.shape[0], self.grid_size ** 2, self.num_latent_tokens
        pos = self.positional_embedding
        tokens = Fn.TokensAssembleFn.apply(pixel_values, self.patch_embed.weight, self.patch_embed.bias, pos[1:], self.class_embedding,
                                           None, pos[:1], latent_tokens, self.latent_token_positional_embedding, self.patch_size)
        x = Fn.LayerNormFn.apply(tokens, self.ln_pre.weight, self.ln_pre.bias, 0, 1 + P + L, self.ln_pre.eps)
        x = _run_rab_stack(self.transformer, x, self.num_heads)
        lat = Fn.LayerNormFn.apply(x, self.ln_post.weight, self.ln_post.bias, 1 + P, L, self.ln_post.eps)     # [B, L, d]
        y = Fn.LinearFn.apply(lat, self.conv_out.weight.view(self.token_size, self.width), self.conv_out.bias, False)
        return y.permute(0, 2, 1).unsqueeze(2)                                    # [B, token_size, 1, L] (blocks.py:281)


class BlocksTiTokDecoder(nn.Module):
    """blocks.TiTokDecoder (blocks.py:285-361) -- exported as `blocks.TiTokDecoder` by shim/blocks.py.  decoder_embed + the
    mask-token / class-token / positional assembly of blocks.py:341-352 is one GEMM + one pass (TokensAssembleFn), ln_pre, the
    block stack, ln_post on the grid tokens, and `ffn` (1x1 conv + pixel-shuffle Rearrange) as the de-patchify GEMM whose
    epilogue stores the NCHW image.  The final 3x3 `conv_out` (blocks.py:334,361) is an ordinary padded convolution outside
    the §8 path and stays nn.Conv2d."""

    def __init__(self, config):
        super().__init__()
        from einops.layers.torch import Rearrange
        self.config = config
        self.image_size = config.image_size
        self.patch_size = config.patch_size
        self.grid_size = self.image_size // self.patch_size
        self.model_size = config.transformer
        self.num_latent_tokens = config.latent_tokens
        self.token_size = config.latent_dim
        self.width, self.num_layers, self.num_heads = _BLOCKS_SIZES[self.model_size]
        self.decoder_embed = nn.Linear(self.token_size, self.width, bias=True)
        scale = self.width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(1, self.width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(self.grid_size ** 2 + 1, self.width))
        self.mask_token = nn.Parameter(scale * torch.randn(1, 1, self.width))
        self.latent_token_positional_embedding = nn.Parameter(scale * torch.randn(self.num_latent_tokens, self.width))
        self.ln_pre = nn.LayerNorm(self.width)
        self.transformer = nn.ModuleList([ResidualAttentionBlock(self.width, self.num_heads, mlp_ratio=4.0)
                                          for _ in range(self.num_layers)])
        self.ln_post = nn.LayerNorm(self.width)
        self.ffn = nn.Sequential(                                            # parameter container: ffn.0.{weight,bias}
            nn.Conv2d(self.width, self.patch_size * self.patch_size * 3, 1, padding=0, bias=True),
            Rearrange("b (p1 p2 c) h w -> b c (h p1) (w p2)", p1=self.patch_size, p2=self.patch_size))
        self.conv_out = nn.Conv2d(3, 3, 3, padding=1, bias=True)

    def forward(self, z_quantized):
        N, C, H, W = z_quantized.shape
        assert H == 1 and W == self.num_latent_tokens, f"{H}, {W}, {self.num_latent_tokens}"
        P = self.grid_size ** 2
        lat = z_quantized.reshape(N, C * H, W).permute(0, 2, 1)                    # [B, L, token_size]
        tokens = Fn.TokensAssembleFn.apply(lat, self.decoder_embed.weight, self.decoder_embed.bias,
                                           self.latent_token_positional_embedding[:W], self.class_embedding,
                                           self.mask_token.view(1, self.width), self.positional_embedding, None, None, 0)
        x = Fn.LayerNormFn.apply(tokens, self.ln_pre.weight, self.ln_pre.bias, 0, 1 + P + W, self.ln_pre.eps)
        x = _run_rab_stack(self.transformer, x, self.num_heads)
        x = Fn.LayerNormFn.apply(x, self.ln_post.weight, self.ln_post.bias, 1, P, self.ln_post.eps)            # [B, P, d]
        img = Fn.DepatchifyFn.apply(x, self.ffn[0].weight, self.ffn[0].bias, self.grid_size, self.grid_size, self.patch_size)
        return self.conv_out(img)


class VectorQuantizer(nn.Module):
    """blocks.VectorQuantizer (blocks.py:405-505).  The clustering_vq branch is dead code in the reference
    (undefined `gather`, blocks.py:457) and is rejected here."""

    def __init__(self, codebook_size: int = 1024, token_size: int = 256, commitment_cost: float = 0.25,
                 use_l2_norm: bool = False, clustering_vq: bool = False):
        super().__init__()
        if clustering_vq:
            raise NotImplementedError("clustering_vq is broken in the reference (blocks.py:457) and not implemented")
        self.codebook_size = codebook_size
        self.token_size = token_size
        self.commitment_cost = commitment_cost
        self.embedding = nn.Embedding(codebook_size, token_size)
        self.embedding.weight.data.uniform_(-1.0 / codebook_size, 1.0 / codebook_size)
        self.use_l2_norm = use_l2_norm
        self.clustering_vq = clustering_vq

    def forward(self, z):
        z_q, idx, mse, commit, total = Fn.VQFn.apply(z, self.embedding.weight, self.use_l2_norm, self.use_l2_norm,
                                                     True, float(self.commitment_cost))
        result = dict(quantizer_loss=total, commitment_loss=commit, codebook_loss=mse,
                      min_encoding_indices=idx.view(z.shape[0], z.shape[2], z.shape[3]))
        return z_q, result

    def get_codebook_entry(self, indices):
        if len(indices.shape) == 1:
            z_quantized = self.embedding(indices)
        elif len(indices.shape) == 2:
            z_quantized = torch.einsum("bd,dn->bn", indices, self.embedding.weight)
        else:
            raise NotImplementedError
        if self.use_l2_norm:
            z_quantized = torch.nn.functional.normalize(z_quantized, dim=-1)
        return z_quantized
