"""Drop-in for ywz/mywork/.trash/newnet9.py -- the module ywz/mywork/test3real.py:12 imports:
HESIC without the "twiceLeft" re-encode."""
from _star_names import *  # noqa: F401,F403
from hesic_b200.stereo import (AverageMeter, CompressionModel, Decoder1, Decoder2, Encoder1, Encoder2,  # noqa: F401
                               Enhancement, Enhancement_Block, Independent_EN, RateDistortionLoss, encode_hyper,
                               gmm_hyper_y1, gmm_hyper_y2, spatial_pool2d)
from hesic_b200.stereo import HSIC_NoTwiceLeft as HSIC  # noqa: F401
