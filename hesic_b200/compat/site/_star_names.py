"""Names the reference's model files leak through ``from newnet1 import *`` and that the drivers
rely on: standard modules, torch namespaces, compressai symbols."""
import argparse  # noqa: F401
import math  # noqa: F401
import os  # noqa: F401
import random  # noqa: F401
import shutil  # noqa: F401
import sys  # noqa: F401
import time  # noqa: F401

import numpy as np  # noqa: F401
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import torch.optim as optim  # noqa: F401
from torch.utils.data import DataLoader  # noqa: F401

try:
    from torchvision import transforms  # noqa: F401
except Exception:  # pragma: no cover
    transforms = None
try:
    from PIL import Image  # noqa: F401
except Exception:  # pragma: no cover
    Image = None

import kornia  # noqa: F401  (real package if installed, else hesic_b200/compat/shims/kornia)
from compressai.ans import BufferedRansEncoder, RansDecoder  # noqa: F401
from compressai.datasets import ImageFolder  # noqa: F401
from compressai.entropy_models import EntropyBottleneck, GaussianConditional, GaussianMixtureConditional  # noqa: F401
from compressai.layers import *  # noqa: F401,F403
from compressai.layers import GDN, MaskedConv2d  # noqa: F401
from compressai.models.utils import conv, deconv, update_registered_buffers  # noqa: F401

try:
    from range_coder import RangeDecoder, RangeEncoder, prob_to_cum_freq  # noqa: F401
except Exception:  # pragma: no cover
    RangeEncoder = RangeDecoder = prob_to_cum_freq = None
