"""Unsupervised deep-homography front-end (ywz/mywork/model.py == udh/udh/model.py): the
network that produces ``h_matrix`` for ``HSIC.forward`` in ywz/mywork/test3real.py:171-181.

Same class tree, constructor arguments and ``state_dict`` keys as the reference file
(``cnn.{0-3}.layers.{0,2}.{weight,bias}``, ``fc.{2,5}.{weight,bias}``).  Operator level: the eight 3x3
convolutions and the two fully-connected layers (as 1x1 convolutions over a 1x1 image) run on the hesic_b200
conv kernels with the ReLU fused; max-pooling is ``hesic_max_pool2x2`` and ``get_h`` (4-point DLT + 3x3 inverse) ONE
kernel launch (``hesic_perspective_transform``; r01/r02 left these in eager torch: 1.7 ms of the 25 ms driver flow) --
this is the step BEFORE the hot path (2.6 GF per pair against HSIC's 155.7).
"""
import torch
import torch.nn as nn

from . import _capi as C
from . import functional as F
from .modules import Conv2d


class _ConvReLU(Conv2d):
    """nn.Conv2d followed by the nn.ReLU that comes next in the reference's nn.Sequential (fused into the conv
    epilogue; the nn.ReLU module stays in place as a no-op on the non-negative result, keeping the key layout)."""

    def forward(self, x):
        C.require_cuda(x)
        return F.conv2d(x, self.hesic_plan(), act=C.ACT_RELU)


class Linear(nn.Linear):
    """nn.Linear evaluated as a 1x1 convolution over a 1x1 image on the same kernels."""
    _plan = None
    fused_relu = False

    def forward(self, x):
        C.require_cuda(x)
        if self._plan is None:
            object.__setattr__(self, "_plan", F.ConvPlan(self.in_features, self.out_features, 1, 1, 0))
        w = self.weight.detach().reshape(self.out_features, self.in_features, 1, 1)
        self._plan.load(w, self.bias)
        y = F.conv2d(x.reshape(x.shape[0], self.in_features, 1, 1), self._plan,
                     act=C.ACT_RELU if self.fused_relu else C.ACT_NONE)
        return y.reshape(x.shape[0], self.out_features)

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        if self._plan is not None:
            self._plan._key = None
        return r


def save_pic(data, path):
    """model.py:9-16."""
    from torchvision import transforms
    reimage = data.cpu().clone() if data.device.type != "cpu" else data
    transforms.ToPILImage()(reimage.squeeze(0)).save(path)


def photometric_loss(delta, img_a, patch_b, corners):
    """model.py:18-45."""
    import kornia
    corners_hat = corners + delta
    corners = corners - corners[:, 0].view(-1, 1, 2)
    h = kornia.get_perspective_transform(corners, corners_hat)
    h_inv = torch.inverse(h)
    patch_b_hat = kornia.warp_perspective(img_a, h_inv, (patch_b.shape[-2], patch_b.shape[-1]))
    return torch.nn.functional.l1_loss(patch_b_hat, patch_b)


class MaxPool2d(nn.MaxPool2d):
    """nn.MaxPool2d(2, 2) of model.py:66 on the library's kernel (CUDA fp32 NCHW); any other configuration or input is
    torch's."""

    def forward(self, x):
        k = self.kernel_size if isinstance(self.kernel_size, int) else None
        s = self.stride if isinstance(self.stride, int) else None
        if (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and k == 2 and s == 2 and self.padding == 0 and
                self.dilation == 1 and not self.ceil_mode and not self.return_indices and not x.requires_grad):
            return F.max_pool2x2(x)
        return super().forward(x)


class Flatten(nn.Module):
    def forward(self, x):
        return x.view(x.size(0), -1)


class Block(nn.Module):
    """model.py:53-70: conv3x3, ReLU, [BN], conv3x3, ReLU, [BN], [MaxPool2d(2, 2)]."""

    def __init__(self, inchannels, outchannels, batch_norm=False, pool=True):
        super().__init__()
        layers = [_ConvReLU(inchannels, outchannels, kernel_size=3, padding=1), nn.ReLU()]
        if batch_norm:
            layers.append(nn.BatchNorm2d(outchannels))
        layers += [_ConvReLU(outchannels, outchannels, kernel_size=3, padding=1), nn.ReLU()]
        if batch_norm:
            layers.append(nn.BatchNorm2d(outchannels))
        if pool:
            layers.append(MaxPool2d(2, 2))
        self.layers = nn.Sequential(*layers)

    def forward(self, x):
        return self.layers(x)


class Net(nn.Module):
    """model.py:73-111."""

    def __init__(self, batch_norm=False, patch_size=128):
        super().__init__()
        self.cnn = nn.Sequential(Block(2, 64, batch_norm), Block(64, 64, batch_norm), Block(64, 128, batch_norm),
                                 Block(128, 128, batch_norm, pool=False))
        fc1 = Linear(128 * (patch_size // 8) * (patch_size // 8), 1024)
        fc1.fused_relu = True
        self.fc = nn.Sequential(Flatten(), nn.Dropout(p=0.5), fc1, nn.ReLU(), nn.Dropout(p=0.5), Linear(1024, 4 * 2))

    def _delta(self, a, b):
        x = torch.cat((a, b), dim=1)
        x = self.cnn(x)
        return self.fc(x.contiguous()).view(-1, 4, 2)

    def forward(self, a, b):
        return self._delta(a, b)

    def get_h(self, a, b, corners):
        corners_hat = corners + self._delta(a, b)
        if corners.is_cuda and corners.dtype == torch.float32:
            return F.perspective_transform(corners, corners_hat, invert=True)   # DLT + inverse, one launch
        import kornia
        return torch.inverse(kornia.get_perspective_transform(corners, corners_hat))
