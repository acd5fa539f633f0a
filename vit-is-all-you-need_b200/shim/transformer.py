"""Drop-in for the reference's top-level module `transformer` (transformer.py:1-59): same public names, backed
by the sm_100a kernels.  Put this directory ahead of the reference on sys.path (b200vit.launch does)."""
from b200vit.modules import (Attention, B, L, S, Transformer, TransformerConfig, TransformerLayer,  # noqa: F401
                             transformer_configs)
