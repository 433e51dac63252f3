"""Drop-in for ywz/mywork/newnet1_joint.py (HESIC+)."""
from _star_names import *  # noqa: F401,F403
from compressai.models.priors import SCALES_LEVELS, SCALES_MAX, SCALES_MIN, get_scale_table  # noqa: F401
from hesic_b200.stereo import (AverageMeter, CompressionModel, Decoder1, Decoder2, Encoder1, Encoder2,  # noqa: F401
                               Enhancement, Enhancement_Block, Independent_EN, RateDistortionLoss, encode_hyper,
                               gmm_hyper_y1, gmm_hyper_y2, spatial_pool2d)
from hesic_b200.stereo import HSIC_Joint as HSIC  # noqa: F401
