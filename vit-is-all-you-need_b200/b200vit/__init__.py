"""b200vit — B200-native (sm_100a) ViT-encoder hot path behind the reference's nn.Module surface.

Public surface mirrors the reference (SnakeOnex/vit-is-all-you-need): see modules.py.  The drop-in shims in
../shim/ (transformer.py, blocks.py, train_vit.py) let the reference's training scripts import these classes
by their original bare module names.
"""
from .modules import (Attention, B, BlocksTiTokDecoder, BlocksTiTokEncoder, CrossEntropyLoss, L, PatchConv2d, Quantizer, ResidualAttentionBlock, S,  # noqa: F401
                      Transformer, TransformerConfig,
                      TiTok, TiTokDecoder, TiTokEncoder, TransformerLayer, UViTBlock, VectorQuantizer, VideoGPT, ViT, ViTClassifier, ViTConfig, transformer_configs)

__version__ = "0.1.0"
