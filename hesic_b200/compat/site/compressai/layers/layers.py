"""compressai/layers/layers.py on the hesic_b200 conv kernels.  MaskedConv2d, conv3x3 and
ResidualBlock are on the HESIC / HESIC+ / Independent_EN paths; the remaining blocks exist for the
operator surface (Cheng2020 zoo models, out of scope) and are thin compositions of the same convs."""
import torch
import torch.nn as nn

from hesic_b200 import _capi as _C
from hesic_b200 import functional as _F
from hesic_b200.modules import Conv2d

from .gdn import GDN

__all__ = ["AttentionBlock", "MaskedConv2d", "ResidualBlock", "ResidualBlockUpsample", "ResidualBlockWithStride",
           "conv3x3", "subpel_conv3x3", "conv1x1"]


class MaskedConv2d(Conv2d):
    """PixelCNN-style causal convolution (layers.py:21-45): mask type 'A' also hides the centre tap."""

    def __init__(self, *args, mask_type="A", **kwargs):
        super().__init__(*args, **kwargs)
        if mask_type not in ("A", "B"):
            raise ValueError(f'Invalid "mask_type" value "{mask_type}"')
        self.register_buffer("mask", torch.ones_like(self.weight.data))
        _, _, h, w = self.mask.size()
        self.mask[:, :, h // 2, w // 2 + (mask_type == "B"):] = 0
        self.mask[:, :, h // 2 + 1:] = 0

    def forward(self, x):
        _C.require_cuda(x)
        plan = self._hesic_plan
        key = _F.ConvPlan._ver(self.weight, self.bias, self.mask)
        if plan is None or plan._key != key:
            # the reference masks the stored weights in place on every call (layers.py:44); do it when
            # the weights changed, then pack (the pack kernel applies the mask as well)
            self.weight.data *= self.mask
        return _F.conv2d(x, self.hesic_plan())


def conv3x3(in_ch, out_ch, stride=1):
    return Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)


def subpel_conv3x3(in_ch, out_ch, r=1):
    return nn.Sequential(Conv2d(in_ch, out_ch * r ** 2, kernel_size=3, padding=1), nn.PixelShuffle(r))


def conv1x1(in_ch, out_ch, stride=1):
    return Conv2d(in_ch, out_ch, kernel_size=1, stride=stride)


class ResidualBlockWithStride(nn.Module):
    def __init__(self, in_ch, out_ch, stride=2):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch, stride=stride)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.gdn = GDN(out_ch)
        self.downsample = conv1x1(in_ch, out_ch, stride=stride) if stride != 1 else None

    def forward(self, x):
        out = self.gdn(self.conv2(self.leaky_relu(self.conv1(x))))
        identity = x if self.downsample is None else self.downsample(x)
        return out + identity


class ResidualBlockUpsample(nn.Module):
    def __init__(self, in_ch, out_ch, upsample=2):
        super().__init__()
        self.subpel_conv = subpel_conv3x3(in_ch, out_ch, upsample)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv = conv3x3(out_ch, out_ch)
        self.igdn = GDN(out_ch, inverse=True)
        self.upsample = subpel_conv3x3(in_ch, out_ch, upsample)

    def forward(self, x):
        out = self.igdn(self.conv(self.leaky_relu(self.subpel_conv(x))))
        return out + self.upsample(x)


class ResidualBlock(nn.Module):
    """Two 3x3 convolutions with LeakyReLU and a skip (layers.py:125-147)."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)

    def forward(self, x):
        _C.require_cuda(x)
        out = _F.conv2d(x, self.conv1.hesic_plan(), act=_C.ACT_LEAKY)
        out = _F.conv2d(out, self.conv2.hesic_plan(), act=_C.ACT_LEAKY)
        return out + x


class AttentionBlock(nn.Module):
    def __init__(self, N):
        super().__init__()

        class ResidualUnit(nn.Module):
            def __init__(self):
                super().__init__()
                self.conv = nn.Sequential(conv1x1(N, N // 2), nn.ReLU(inplace=True), conv3x3(N // 2, N // 2),
                                          nn.ReLU(inplace=True), conv1x1(N // 2, N))
                self.relu = nn.ReLU(inplace=True)

            def forward(self, x):
                return self.relu(self.conv(x) + x)

        self.conv_a = nn.Sequential(ResidualUnit(), ResidualUnit(), ResidualUnit())
        self.conv_b = nn.Sequential(ResidualUnit(), ResidualUnit(), ResidualUnit(), conv1x1(N, N))

    def forward(self, x):
        return self.conv_a(x) * torch.sigmoid(self.conv_b(x)) + x
