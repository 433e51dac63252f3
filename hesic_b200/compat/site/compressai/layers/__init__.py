from .gdn import *  # noqa: F401,F403
from .layers import *  # noqa: F401,F403

__all__ = [
    "GDN",
    "GDN1",
    "AttentionBlock",
    "MaskedConv2d",
    "ResidualBlock",
    "ResidualBlockUpsample",
    "ResidualBlockWithStride",
    "conv3x3",
    "subpel_conv3x3",
]
