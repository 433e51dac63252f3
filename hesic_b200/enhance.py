"""Whole-forward orchestration of Independent_EN, the cross-quality enhancement that follows HSIC in
ywz/mywork/test3real.py:186 (newnet1.py:272-311, 1278-1300).

Host glue only.  Per view: ``cat(x, x_other_warp)`` is packed once into the NHWC_HILO activation format (32
channel slots, 128 B per pixel), the 20 convolutions run as ``en_conv_kernel`` launches (enhance.cu) that keep
that format end to end, and the LeakyReLU / identity additions of ResidualBlock, Enhancement_Block and
Enhancement live in the conv epilogues -- 22 kernels per view instead of ~75 operator-level launches with
layout conversions on both sides of every conv.
"""
import torch

from . import _capi as C
from . import functional as F
from .engine import EngineBase

_lib = C.lib


class EnConvPlan:
    """Owns a ``hesic_en_conv`` handle: the packed [Wl | Wh] tap tiles of one conv3x3 of the enhancement network."""

    def __init__(self, Cin, Cout):
        self.geom = (Cin, Cout)
        self.h = _lib.hesic_en_conv_create(Cin, Cout)
        if not self.h:
            raise ValueError(C.last_error())
        self._key = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.hesic_en_conv_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def __deepcopy__(self, memo):       # one owner per C handle: a copy is a fresh, unloaded plan (see ConvPlan)
        return type(self)(*self.geom)

    def __reduce__(self):
        return (type(self), tuple(self.geom))

    def invalidate(self):
        self._key = None

    @C.device_guard
    def load(self, weight, bias=None):
        key = F.ConvPlan._ver(weight, bias)
        if key != self._key:
            w = F._f32(weight.detach())
            b = F._f32(bias.detach()) if bias is not None else None
            C.check(_lib.hesic_en_conv_load(self.h, C.ptr(w), C.ptr(b), C.stream()))
            self._key = key
        return self

    def run(self, x_desc, y_desc, act=C.ACT_NONE, res1=None, res2=None):
        C.check(_lib.hesic_en_conv_forward(self.h, C.ref(x_desc), C.ref(y_desc), act,
                                           C.ref(res1) if res1 is not None else None,
                                           C.ref(res2) if res2 is not None else None, C.stream()))


def en_plan(conv_mod):
    """Lazily created EnConvPlan of a 3x3 / stride 1 / padding 1 nn.Conv2d with the current weights packed."""
    plan = getattr(conv_mod, "_hesic_en_plan", None)
    if plan is None:
        if tuple(conv_mod.kernel_size) != (3, 3) or tuple(conv_mod.stride) != (1, 1) or tuple(conv_mod.padding) != (1, 1):
            raise NotImplementedError("hesic_b200 enhancement conv: 3x3, stride 1, padding 1 only")
        plan = EnConvPlan(conv_mod.in_channels, conv_mod.out_channels)
        object.__setattr__(conv_mod, "_hesic_en_plan", plan)
    return plan.load(conv_mod.weight, conv_mod.bias)


class EnhanceEngine(EngineBase):
    def __init__(self, model, align_corners=True):
        self.m = model
        self.align_corners = align_corners
        self._bufs = {}

    def _buf(self, name, B, H, W, dev):
        key = (name, B, H, W, str(dev))
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty((B, H, W, 64), device=dev, dtype=torch.bfloat16)
            self._bufs = {k: v for k, v in self._bufs.items() if k[1:] == key[1:]}   # drop buffers of other shapes
            self._bufs[key] = t
        return t

    def _enhancement(self, eh, x, x_other_warp, out):
        """Enhancement.forward (newnet1.py:296-311) for one view; x / x_other_warp / out: NCHW fp32.  x_other_warp
        None: the DSIC variant without a cross-view input (mynet6_plus.py:57-78)."""
        B, _, H, W = x.shape
        dev = x.device
        a, b, c, t = (C.hilo(self._buf(n, B, H, W, dev)) for n in "abct")
        C.check(_lib.hesic_en_pack_input(C.ref(C.nchw(x)), C.ref(C.nchw(x_other_warp)) if x_other_warp is not None else None,
                                         C.ref(t), C.stream()))
        en_plan(eh.conv1).run(t, a)
        for eb in (eh.EB1, eh.EB2, eh.EB3):
            # Enhancement_Block (newnet1.py:272-287): RB3(RB2(RB1(a))) + a, ResidualBlock = lrelu(conv2(lrelu(conv1(x)))) + x
            en_plan(eb.RB1.conv1).run(a, t, C.ACT_LEAKY)
            en_plan(eb.RB1.conv2).run(t, b, C.ACT_LEAKY, res1=a)
            en_plan(eb.RB2.conv1).run(b, t, C.ACT_LEAKY)
            en_plan(eb.RB2.conv2).run(t, c, C.ACT_LEAKY, res1=b)
            en_plan(eb.RB3.conv1).run(c, t, C.ACT_LEAKY)
            en_plan(eb.RB3.conv2).run(t, b, C.ACT_LEAKY, res1=c, res2=a)
            a, b = b, a
        en_plan(eh.conv2).run(a, C.nchw(out), C.ACT_NONE, res1=C.nchw(x))
        return out

    @C.device_guard
    def forward_mono(self, x1_hat, x2_hat):
        """mynet6_plus.Independent_EN.forward: each view enhanced on its own."""
        C.require_cuda(x1_hat, x2_hat)
        x1_hat, x2_hat = F._f32(x1_hat), F._f32(x2_hat)
        if x1_hat.dim() != 4 or x1_hat.shape[1] != 3 or x2_hat.shape != x1_hat.shape:
            raise ValueError("Independent_EN: x1_hat and x2_hat must be [B,3,H,W] tensors of the same shape")
        return {"x1_hat": self._enhancement(self.m.EH1, x1_hat, None, torch.empty_like(x1_hat)),
                "x2_hat": self._enhancement(self.m.EH2, x2_hat, None, torch.empty_like(x2_hat))}

    @C.device_guard
    def forward(self, x1_hat, x2_hat, h_matrix):
        C.require_cuda(x1_hat, x2_hat, h_matrix)
        x1_hat, x2_hat = F._f32(x1_hat), F._f32(x2_hat)
        if x1_hat.shape != x2_hat.shape or x1_hat.dim() != 4 or x1_hat.shape[1] != 3:
            raise ValueError("Independent_EN: x1_hat and x2_hat must be [B,3,H,W] tensors of the same shape")
        size = (x1_hat.size(-2), x1_hat.size(-1))
        x1_hat_warp = F.warp_perspective(x1_hat, h_matrix, size, self.align_corners)
        x2_hat_warp = F.warp_perspective(x2_hat, torch.inverse(h_matrix), size, self.align_corners)
        o1 = self._enhancement(self.m.EH1, x1_hat, x2_hat_warp, torch.empty_like(x1_hat))
        o2 = self._enhancement(self.m.EH2, x2_hat, x1_hat_warp, torch.empty_like(x2_hat))
        return {"x1_hat": o1, "x2_hat": o2}
