"""GDN / IGDN (compressai/layers/gdn.py:22-97) on the hesic_b200 kernels.

y[i] = x[i] * (beta[i] + sum_j gamma[i,j] x[j]^2) ^ (-1/2)   (inverse: ^ (+1/2))

Parameters are stored exactly as the reference stores them (pre-reparametrisation ``beta``,
``gamma`` plus the ``*_reparam`` buffers) so checkpoints load strictly; the reparametrisation is
applied on the device.  Inside HSIC.forward the GDN is fused into the preceding convolution."""
import torch
import torch.nn as nn

from compressai.ops.parametrizers import NonNegativeParametrizer
from hesic_b200 import _capi as _C
from hesic_b200 import functional as _F

__all__ = ["GDN", "GDN1"]


class GDN(nn.Module):
    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.beta_min = float(beta_min)
        gamma_init = float(gamma_init)
        self.inverse = bool(inverse)

        self.beta_reparam = NonNegativeParametrizer(minimum=self.beta_min)
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))

        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(gamma_init * torch.eye(in_channels)))

    def forward(self, x):
        _C.require_cuda(x)
        if x.size(1) != self.beta.numel():
            raise ValueError(f"GDN: expected {self.beta.numel()} channels, got {x.size(1)}")
        return _F.gdn(x, self.beta, self.gamma, self.inverse, self.beta_min)


class GDN1(GDN):
    """Simplified GDN (gdn.py:73-97): y = x / (beta + sum_j gamma |x_j|).  Not on the HESIC path
    (SURVEY.md section 2, row 6); kept for the operator surface, evaluated through the GDN kernel's
    1x1 contraction on |x|."""

    def forward(self, x):
        _C.require_cuda(x)
        beta = self.beta_reparam(self.beta)
        gamma = self.gamma_reparam(self.gamma)
        # contraction of |x| with gamma through the conv path (weights = gamma, bias = beta)
        Cn = x.size(1)
        plan = _F.ConvPlan(Cn, Cn, 1, 1, 0).load(gamma.reshape(Cn, Cn, 1, 1).contiguous(), beta.contiguous())
        norm = _F.conv2d(torch.abs(x), plan, path=_C.PATH_SIMT)
        if not self.inverse:
            norm = 1.0 / norm
        return x * norm
