"""Drop-in for ywz/mywork/model.py (== udh/udh/model.py): the homography network the drivers build as
``HomographyModel`` (test3real.py:46-51)."""
import os  # noqa: F401

import kornia  # noqa: F401
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F  # noqa: F401

from hesic_b200.homography import Block, Flatten, Net, photometric_loss, save_pic  # noqa: F401
