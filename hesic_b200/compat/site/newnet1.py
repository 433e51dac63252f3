"""Drop-in for ywz/mywork/newnet1.py (HESIC): ``from newnet1 import *`` gives the same names."""
from _star_names import *  # noqa: F401,F403
from hesic_b200.stereo import (AverageMeter, CompressionModel, Decoder1, Decoder2, Encoder1, Encoder2,  # noqa: F401
                               Enhancement, Enhancement_Block, GMM_together, HSIC, Independent_EN,
                               RateDistortionLoss, encode_hyper, gmm_hyper_y1, gmm_hyper_y2, spatial_pool2d)
