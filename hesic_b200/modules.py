"""nn.Module leaves whose forward runs on libhesic_b200.so.

They subclass the torch modules the reference instantiates (nn.Conv2d,
nn.ConvTranspose2d) so parameters, ``state_dict`` keys, initialisation and
``.to(device)`` behave identically; only ``forward`` is replaced.  CPU tensors
raise -- there is no CPU fallback on this path.
"""
import torch
import torch.nn as nn

from . import _capi as C
from . import functional as F


def _one(v, what):
    if isinstance(v, (tuple, list)):
        if len(set(v)) != 1:
            raise NotImplementedError(f"hesic_b200 conv: non-square {what} {v}")
        return int(v[0])
    return int(v)


class _PlanMixin:
    _hesic_plan = None

    def _geometry_ok(self):
        if _one(self.dilation, "dilation") != 1 or self.groups != 1:
            raise NotImplementedError("hesic_b200 conv: dilation/groups are not used by the HESIC path")
        if getattr(self, "padding_mode", "zeros") != "zeros":
            raise NotImplementedError("hesic_b200 conv: only zero padding")

    def hesic_plan(self):
        """Lazily created ConvPlan with the current weights packed (re-packed when they change) and NO fused GDN:
        an engine that fuses the GDN that follows re-attaches it (ConvPlan.set_gdn) after fetching the plan, so a
        stand-alone call of this layer after an engine run is again a plain convolution."""
        if self._hesic_plan is None:
            self._geometry_ok()
            tr = isinstance(self, nn.ConvTranspose2d)
            self._hesic_plan = F.ConvPlan(self.in_channels, self.out_channels, tuple(int(k) for k in self.kernel_size),
                                          _one(self.stride, "stride"), _one(self.padding, "padding"), tr,
                                          _one(self.output_padding, "output_padding") if tr else 0)
        mask = getattr(self, "mask", None)
        self._hesic_plan.load(self.weight, self.bias, mask)
        self._hesic_plan.set_gdn(None, None, False)
        return self._hesic_plan

    def _apply(self, fn, *a, **k):
        # moving devices invalidates packed operands
        r = super()._apply(fn, *a, **k)
        if self._hesic_plan is not None:
            self._hesic_plan._key = None
            self._hesic_plan._gdn_key = None
        return r

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == "_hesic_plan" else copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_hesic_plan"] = None
        return d


class Conv2d(_PlanMixin, nn.Conv2d):
    """nn.Conv2d as built by compressai.models.utils.conv (models/utils.py:104-109)."""

    def forward(self, x):
        C.require_cuda(x)
        return F.conv2d(x, self.hesic_plan())


class ConvTranspose2d(_PlanMixin, nn.ConvTranspose2d):
    """nn.ConvTranspose2d as built by compressai.models.utils.deconv (models/utils.py:112-118)."""

    def forward(self, x, output_size=None):
        if output_size is not None:
            raise NotImplementedError("hesic_b200 deconv: output_size is not supported")
        C.require_cuda(x)
        return F.conv2d(x, self.hesic_plan())
