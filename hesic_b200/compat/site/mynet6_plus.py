"""Drop-in for ywz/DSIC/mynet6_plus.py (DSIC): ``from mynet6_plus import *`` gives the same names."""
from _star_names import *  # noqa: F401,F403
from hesic_b200.dsic import (DSIC, DSIC_plus, Enhancement, Enhancement_Block, Independent_EN, cost_volume,  # noqa: F401
                             dense_warp, global_context)
from hesic_b200.stereo import (AverageMeter, CompressionModel, Decoder1, Encoder1, RateDistortionLoss,  # noqa: F401
                               encode_hyper, gmm_hyper_y1, gmm_hyper_y2, spatial_pool2d)
