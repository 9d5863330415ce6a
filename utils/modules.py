"""Import-path shim for the reference's notebook variant ``utils/modules.py`` -> paella_b200.modules."""
from paella_b200.modules import (Attention2D, AttnBlock, FeedForwardBlock, GlobalResponseNorm, LayerNorm2d, Paella,  # noqa: F401
                                 ResBlock, TimestepBlock)
