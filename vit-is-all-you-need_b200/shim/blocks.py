"""Drop-in for the reference's top-level module `blocks`.

The hot-path classes (ResidualAttentionBlock blocks.py:32-70, UViTBlock blocks.py:174-201, TiTokEncoder / TiTokDecoder
blocks.py:208-361 with the fused token-sequence assembly, VectorQuantizer blocks.py:405-505) come from b200vit.modules;
everything else the reference's blocks.py defines (TATiTokDecoder, the dead UViT helpers ...) is taken from the reference's own
file, executed in this module's namespace, so that `from blocks import TiTokEncoder, TiTokDecoder, VectorQuantizer`
(train_tatitok.py:13) keeps working."""
import os
import sys

from b200vit.modules import BlocksTiTokDecoder as _DEC
from b200vit.modules import BlocksTiTokEncoder as _ENC
from b200vit.modules import ResidualAttentionBlock as _RAB
from b200vit.modules import UViTBlock as _UVB
from b200vit.modules import VectorQuantizer as _VQ

_here = os.path.dirname(os.path.abspath(__file__))
for _p in sys.path:
    _cand = os.path.join(_p or ".", "blocks.py")
    if os.path.isfile(_cand) and os.path.abspath(os.path.dirname(_cand)) != _here:
        with open(_cand) as _f:
            exec(compile(_f.read(), _cand, "exec"), globals())  # the reference's own definitions, unmodified
        break

ResidualAttentionBlock = _RAB
TiTokEncoder = _ENC
TiTokDecoder = _DEC
VectorQuantizer = _VQ
UViTBlock = _UVB
