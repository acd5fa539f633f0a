"""Drop-in for the hot-path classes of the reference's `blocks` module (blocks.py:32-70, 405-505)."""
from b200vit.modules import ResidualAttentionBlock, VectorQuantizer  # noqa: F401
