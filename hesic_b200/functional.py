"""Operator-level host API over the C-ABI: the same operations the reference's
``compressai.layers`` / ``compressai.entropy_models`` / ``kornia`` calls perform,
taking and returning NCHW fp32 CUDA tensors.  Plumbing only -- every function
allocates its outputs with torch and launches hand-written sm_100a kernels
through ``_capi``.
"""
import ctypes

import numpy as np
import torch

from . import _capi as C

_lib = C.lib


def _f32(t):
    C.require_cuda(t)
    if t.dtype != torch.float32:
        raise TypeError(f"hesic_b200: expected float32, got {t.dtype}")
    return t.contiguous()


class ConvPlan:
    """Owns a ``hesic_conv`` handle: packed weights (+ optional fused GDN) of one
    nn.Conv2d / nn.ConvTranspose2d as built by compressai/models/utils.py:104-118."""

    def __init__(self, Cin, Cout, k, stride, pad, transposed=False, output_padding=0):
        kh, kw = (k, k) if isinstance(k, int) else k
        self.geom = (Cin, Cout, kh, kw, stride, pad, int(transposed), output_padding)
        self.h = _lib.hesic_conv_create(*self.geom)
        if not self.h:
            raise ValueError(C.last_error())
        self._key = None
        self._gdn_key = None
        self._gdn_on = False

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.hesic_conv_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # The C handle is owned by exactly one Python object: a copy (copy.deepcopy of a module that cached a plan) or an
    # unpickled plan (torch.save(model)) is a fresh, unloaded plan of the same geometry, never a second owner.
    def __deepcopy__(self, memo):
        Cin, Cout, kh, kw, stride, pad, tr, op = self.geom
        return type(self)(Cin, Cout, (kh, kw), stride, pad, bool(tr), op)

    def __reduce__(self):
        Cin, Cout, kh, kw, stride, pad, tr, op = self.geom
        return (type(self), (Cin, Cout, (kh, kw), stride, pad, bool(tr), op))

    def invalidate(self):
        """Force a re-pack at the next load() / set_gdn().  Needed after weight surgery through ``.data`` (``w.data.copy_()``,
        EMA swaps): that does not bump the parameter's version counter, which is what load() keys on."""
        self._key = None
        self._gdn_key = None

    @staticmethod
    def _ver(*ts):
        return tuple((t.data_ptr(), t._version, t.device.index) if t is not None else None for t in ts)

    @C.device_guard
    def load(self, weight, bias=None, mask=None):
        key = self._ver(weight, bias, mask)
        if key == self._key:
            return self
        w = _f32(weight.detach())
        b = _f32(bias.detach()) if bias is not None else None
        m = _f32(mask.detach()) if mask is not None else None
        C.check(_lib.hesic_conv_load(self.h, C.ptr(w), C.ptr(b), C.ptr(m), C.stream()))
        self._key = key
        return self

    @C.device_guard
    def set_gdn(self, beta, gamma, inverse, beta_min=1e-6):
        """Attach (beta, gamma given) or detach (beta None) the fused GDN.  The packed operands are kept across a
        detach, so a layer shared by a fused engine and stand-alone operator calls toggles without re-packing."""
        if beta is None:
            if self._gdn_on:
                C.check(_lib.hesic_conv_enable_gdn(self.h, 0))
                self._gdn_on = False
            return self
        key = self._ver(beta, gamma) + (inverse,)
        if key != self._gdn_key:
            C.check(_lib.hesic_conv_set_gdn(self.h, C.ptr(_f32(beta.detach())), C.ptr(_f32(gamma.detach())), int(inverse),
                                            float(beta_min), C.stream()))
            self._gdn_key = key
        elif not self._gdn_on:
            C.check(_lib.hesic_conv_enable_gdn(self.h, 1))
        self._gdn_on = True
        return self

    @C.device_guard
    def detect_kband(self, weight):
        """Block-banded weights: let the tensor-core path skip the all-zero (N tile, K chunk) blocks.  ``weight`` is the
        tensor last passed to load().  Synchronises the stream once (a handful of flags are read back)."""
        C.check(_lib.hesic_conv_detect_kband(self.h, C.ptr(_f32(weight.detach())), C.stream()))
        return self

    def out_hw(self, H, W):
        Cin, Cout, kh, kw, s, p, tr, op = self.geom
        if not tr:
            return (H + 2 * p - kh) // s + 1, (W + 2 * p - kw) // s + 1
        return (H - 1) * s - 2 * p + kh + op, (W - 1) * s - 2 * p + kw + op

    def run(self, x_desc, y_desc, act=C.ACT_NONE, path=C.PATH_AUTO, xb_desc=None, sse=None):
        """xb_desc: second part of a channel concatenation (torch.cat((x, xb), 1) on the reference side).
        sse = (target_desc, acc): also add sum((y - target)^2) to the fp64 device scalar ``acc`` (hesic_conv_forward_sse)."""
        if sse is not None:
            C.check(_lib.hesic_conv_forward_sse(self.h, C.ref(x_desc), C.ref(xb_desc) if xb_desc is not None else None,
                                                C.ref(y_desc), act, path, C.ref(sse[0]), C.ptr(sse[1]), C.stream()))
        elif xb_desc is not None:
            C.check(_lib.hesic_conv_forward_cat(self.h, C.ref(x_desc), C.ref(xb_desc), C.ref(y_desc), act, path, C.stream()))
        else:
            C.check(_lib.hesic_conv_forward(self.h, C.ref(x_desc), C.ref(y_desc), act, path, C.stream()))


@C.device_guard
def conv2d(x, plan, act=C.ACT_NONE, path=C.PATH_AUTO):
    """NCHW fp32 in -> NCHW fp32 out through a loaded ConvPlan."""
    x = _f32(x)
    B, Cin, H, W = x.shape
    Ho, Wo = plan.out_hw(H, W)
    Cout = plan.geom[1]
    y = torch.empty((B, Cout, Ho, Wo), device=x.device, dtype=torch.float32)
    if path == C.PATH_SIMT or (path == C.PATH_AUTO and Cin <= 8 and Cout <= 4 and plan.geom[4] == 1):
        plan.run(C.nchw(x), C.nchw(y), act, path)     # CUDA-core paths read and write NCHW fp32 directly
        return y
    # the tensor-core path consumes bf16 (hi, lo) planes: ROWPAD8 for the <= 8-channel edge layers,
    # channels-last otherwise
    if Cin <= 8:
        slots = C.rowpad_slots(Cin, plan.geom[4], plan.geom[6])
        xin = torch.zeros((2, B, H + C.ROWPAD_Y, W + C.ROWPAD_X, slots), device=x.device, dtype=torch.bfloat16)
        xd = C.rowpad(xin, Cin)
    else:
        xin = torch.empty((2, B, H, W, Cin), device=x.device, dtype=torch.bfloat16)
        xd = C.split(xin)
    C.check(_lib.hesic_convert(C.ref(C.nchw(x)), C.ref(xd), C.OP_COPY, C.stream()))
    if Cout <= 4:
        plan.run(xd, C.nchw(y), act, path)     # planar epilogue writes NCHW directly
        return y
    yn = torch.empty((B, Ho, Wo, Cout), device=x.device, dtype=torch.float32)
    plan.run(xd, C.nhwc(yn), act, path)
    C.check(_lib.hesic_convert(C.ref(C.nhwc(yn)), C.ref(C.nchw(y)), C.OP_COPY, C.stream()))
    return y


@C.device_guard
def gdn(x, beta, gamma, inverse=False, beta_min=1e-6):
    """compressai/layers/gdn.py:55-70 with the raw (un-reparametrised) parameters."""
    x = _f32(x)
    C.require_cuda(beta, gamma)
    y = torch.empty_like(x)
    C.check(_lib.hesic_gdn(C.ref(C.nchw(x)), C.ref(C.nchw(y)), C.ptr(_f32(beta.detach())), C.ptr(_f32(gamma.detach())),
                           int(inverse), float(beta_min), C.stream()))
    return y


@C.device_guard
def warp_perspective(src, M, dsize, align_corners=True, out=None):
    """kornia.warp_perspective(src, M, dsize) -- bilinear, zero padding."""
    src = _f32(src)
    M = _f32(M)
    if M.shape != (src.shape[0], 3, 3):
        raise ValueError(f"warp_perspective: M must be [B,3,3], got {tuple(M.shape)}")
    if out is None:
        out = torch.empty((src.shape[0], src.shape[1], int(dsize[0]), int(dsize[1])), device=src.device, dtype=torch.float32)
    C.check(_lib.hesic_warp_perspective(C.ref(C.nchw(src)), C.ptr(M), C.ref(C.nchw(out)), None, int(bool(align_corners)), C.stream()))
    return out


@C.device_guard
def perspective_transform(src, dst, invert=False):
    """kornia.get_perspective_transform(src, dst) ([B,4,2] corners -> [B,3,3]); invert=True also applies torch.inverse."""
    src, dst = _f32(src), _f32(dst)
    if src.shape != dst.shape or src.dim() != 3 or tuple(src.shape[1:]) != (4, 2):
        raise ValueError(f"perspective_transform: src and dst must be [B,4,2], got {tuple(src.shape)} and {tuple(dst.shape)}")
    H = torch.empty((src.shape[0], 3, 3), device=src.device, dtype=torch.float32)
    C.check(_lib.hesic_perspective_transform(C.ptr(src), C.ptr(dst), src.shape[0], int(bool(invert)), C.ptr(H), C.stream()))
    return H


@C.device_guard
def max_pool2x2(x):
    """nn.MaxPool2d(2, 2) on NCHW fp32."""
    x = _f32(x)
    y = torch.empty((x.shape[0], x.shape[1], x.shape[2] // 2, x.shape[3] // 2), device=x.device, dtype=torch.float32)
    C.check(_lib.hesic_max_pool2x2(C.ref(C.nchw(x)), C.ref(C.nchw(y)), C.stream()))
    return y


@C.device_guard
def eb_pack(matrices, biases, factors, quantiles):
    """Pre-activate the EntropyBottleneck parameters into the 60-float-per-channel device table."""
    Cn = quantiles.shape[0]
    ts = [_f32(t.detach()) for t in list(matrices) + list(biases) + list(factors)]
    q = _f32(quantiles.detach())
    out = torch.empty((Cn, C.EB_PARAMS), device=q.device, dtype=torch.float32)
    arr = lambda xs: (ctypes.c_void_p * len(xs))(*[t.data_ptr() for t in xs])
    C.check(_lib.hesic_eb_pack(arr(ts[0:5]), arr(ts[5:10]), arr(ts[10:14]), C.ptr(q), Cn, C.ptr(out), C.stream()))
    return out


@C.device_guard
def entropy_bottleneck(z, params, likelihood_bound=1e-9, log2_acc=None):
    """EntropyBottleneck.forward (eval): returns (z_hat, likelihood), NCHW fp32."""
    z = _f32(z)
    z_hat = torch.empty_like(z)
    lik = torch.empty_like(z)
    C.check(_lib.hesic_entropy_bottleneck(C.ref(C.nchw(z)), C.ptr(params), float(likelihood_bound), C.ref(C.nchw(z_hat)),
                                          C.ref(C.nchw(lik)), C.ptr(log2_acc), C.stream()))
    return z_hat, lik


@C.device_guard
def gaussian_mixture_conditional(y, scales, means, weights, K, scale_bound=0.11, likelihood_bound=1e-9, log2_acc=None):
    """GaussianMixtureConditional.forward (eval).  weights: [B, K*M, 1, 1]."""
    y, scales, means = _f32(y), _f32(scales), _f32(means)
    w = _f32(weights).reshape(y.shape[0], -1)
    if w.shape[1] != K * y.shape[1]:
        raise ValueError("weights must have K*M channels")
    y_hat = torch.empty_like(y)
    lik = torch.empty_like(y)
    C.check(_lib.hesic_gaussian_conditional(C.ref(C.nchw(y)), C.ref(C.nchw(scales)), C.ref(C.nchw(means)), C.ptr(w), K, 1,
                                            float(scale_bound), float(likelihood_bound), C.ref(C.nchw(y_hat)),
                                            C.ref(C.nchw(lik)), None, C.ptr(log2_acc), C.stream()))
    return y_hat, lik


@C.device_guard
def gaussian_conditional(y, scales, means=None, scale_bound=0.11, likelihood_bound=1e-9, log2_acc=None):
    """GaussianConditional.forward (eval)."""
    y, scales = _f32(y), _f32(scales)
    mu = C.nchw(_f32(means)) if means is not None else C.null()
    y_hat = torch.empty_like(y)
    lik = torch.empty_like(y)
    C.check(_lib.hesic_gaussian_conditional(C.ref(C.nchw(y)), C.ref(C.nchw(scales)), C.ref(mu), None, 1, 0,
                                            float(scale_bound), float(likelihood_bound), C.ref(C.nchw(y_hat)),
                                            C.ref(C.nchw(lik)), None, C.ptr(log2_acc), C.stream()))
    return y_hat, lik


@C.device_guard
def spatial_max(x):
    """spatial_pool2d (newnet1.py:441-453): [B,C,H,W] -> [B,C,1,1] fp32."""
    x = _f32(x)
    out = torch.empty((x.shape[0], x.shape[1]), device=x.device, dtype=torch.float32)
    C.check(_lib.hesic_spatial_max(C.ref(C.nchw(x)), C.ptr(out), C.stream()))
    return out.reshape(x.shape[0], x.shape[1], 1, 1)


@C.device_guard
def mixture_weights(pooled, w1x1, bias, K, M):
    """LeakyReLU -> conv1x1 -> softmax over the K components: [B,K*M(,1,1)] -> [B,K*M,1,1]."""
    B = pooled.shape[0]
    p = _f32(pooled).reshape(B, K * M)
    out = torch.empty((B, K * M), device=p.device, dtype=torch.float32)
    C.check(_lib.hesic_mixture_weights(C.ptr(p), C.ptr(_f32(w1x1.detach()).reshape(K * M, K * M)),
                                       C.ptr(_f32(bias.detach()) if bias is not None else None), B, K, M, C.ptr(out), C.stream()))
    return out.reshape(B, K * M, 1, 1)


@C.device_guard
def upsample_bilinear(x, scale):
    x = _f32(x)
    y = torch.empty((x.shape[0], x.shape[1], x.shape[2] * scale, x.shape[3] * scale), device=x.device, dtype=torch.float32)
    C.check(_lib.hesic_upsample_bilinear(C.ref(C.nchw(x)), C.ref(C.nchw(y)), int(scale), C.stream()))
    return y


@C.device_guard
def images_from_uint8(u8, out=None):
    """[B,H,W,C] uint8 (interleaved, as cv2 / PIL hold an image) -> [B,C,H,W] float32 = u8 / 255: what
    ``transforms.ToTensor()`` does per image on the host in the reference's loader (compressai/datasets/utils.py:101-102),
    done on the device so that only 1 byte per sample crosses PCIe."""
    C.require_cuda(u8)
    if u8.dtype != torch.uint8 or u8.dim() != 4:
        raise TypeError(f"images_from_uint8: expected a [B,H,W,C] uint8 tensor, got {u8.dtype} {tuple(u8.shape)}")
    u8 = u8.contiguous()
    B, H, W, Cn = u8.shape
    if out is None:
        out = torch.empty((B, Cn, H, W), device=u8.device, dtype=torch.float32)
    C.check(_lib.hesic_images_from_u8(C.ptr(u8), B, H, W, Cn, C.ref(C.nchw(out)), C.stream()))
    return out


@C.device_guard
def round_half_even(x):
    x = _f32(x)
    y = torch.empty_like(x)
    C.check(_lib.hesic_convert(C.ref(C.nchw(x)), C.ref(C.nchw(y)), C.OP_ROUND, C.stream()))
    return y


@C.device_guard
def prepare_symbols(x, means=None):
    """int32 symbols = round_half_even(x - means), [B, C*H*W] (EntropyModel.compress prep).
    ``means``: None, a [1,C,1,1] / [C] per-channel tensor, or a full tensor like x."""
    x = _f32(x)
    B = x.shape[0]
    out = torch.empty((B, x[0].numel()), device=x.device, dtype=torch.int32)
    cm, full = None, None
    if means is not None:
        if means.numel() == x.shape[1]:
            cm = _f32(means.detach()).reshape(-1)
        elif means.shape == x.shape:
            full = C.nchw(_f32(means))
        else:
            raise ValueError("Invalid means parameters")
    C.check(_lib.hesic_prepare_symbols(C.ref(C.nchw(x)), C.ptr(cm), C.ref(full) if full is not None else None, C.ptr(out),
                                       C.stream()))
    return out


def build_indexes_channel(size, device):
    B, Cn, H, W = size
    out = torch.empty((B, Cn, H, W), device=device, dtype=torch.int32)
    C.check(_lib.hesic_build_indexes_channel(B, Cn, H, W, C.ptr(out), C.stream()))
    return out


@C.device_guard
def build_indexes_scale(scales, table, scale_bound=0.11):
    scales = _f32(scales)
    table = _f32(table)
    out = torch.empty(scales.shape, device=scales.device, dtype=torch.int32)
    C.check(_lib.hesic_build_indexes_scale(C.ref(C.nchw(scales)), C.ptr(table), table.numel(), float(scale_bound),
                                           C.ptr(out), C.stream()))
    return out


@C.device_guard
def sum_squared_error(a, b, acc):
    a, b = _f32(a), _f32(b)
    C.check(_lib.hesic_sum_squared_error(C.ref(C.nchw(a)), C.ref(C.nchw(b)), C.ptr(acc), C.stream()))


# ---------------------------------------------------------------------------------------------
# host-side coder (numpy arrays; no device involved)
def pmf_to_quantized_cdf(pmf, precision=16):
    pmf = np.ascontiguousarray(pmf, dtype=np.float32)
    cdf = np.zeros(pmf.size + 1, dtype=np.uint32)
    C.check(_lib.hesic_pmf_to_quantized_cdf(pmf.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), pmf.size, int(precision),
                                            cdf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))))
    return cdf


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class _OwnedHandle:
    """A Python object that owns one C handle (destroyed in ``__del__``).  A copy or an unpickled instance is a FRESH
    coder object, never a second owner of the same handle: entropy models keep their coder as an attribute, so
    ``copy.deepcopy(model)`` / ``torch.save(model)`` reach these (a shared handle was a double free)."""

    def _ctor_args(self):
        return ()

    def __deepcopy__(self, memo):
        return type(self)(*self._ctor_args())

    def __reduce__(self):
        return (type(self), self._ctor_args())


class RansEncoderHandle(_OwnedHandle):
    def __init__(self):
        self.h = _lib.hesic_rans_encoder_create()

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.hesic_rans_encoder_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def push(self, symbols, indexes, cdfs, cdf_sizes, offsets):
        s, i, c, z, o = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1), _i32(cdfs), _i32(cdf_sizes), _i32(offsets)
        if s.size != i.size:
            raise ValueError("symbols and indexes must have the same length")
        if c.ndim != 2 or z.size != c.shape[0] or o.size != c.shape[0]:
            raise ValueError("cdfs must be [n_cdfs, pitch] with matching sizes/offsets")
        C.check(_lib.hesic_rans_encoder_push(self.h, s.ctypes.data, i.ctypes.data, s.size, c.ctypes.data, c.shape[0],
                                             c.shape[1], z.ctypes.data, o.ctypes.data))

    def flush(self):
        n = _lib.hesic_rans_encoder_flush(self.h, None, 0)
        if n < 0:
            C.check(int(n))
        buf = np.empty(n, dtype=np.uint8)
        n2 = _lib.hesic_rans_encoder_flush(self.h, buf.ctypes.data, n)
        assert n2 == n
        return buf.tobytes()


class RansDecoderHandle(_OwnedHandle):
    def __init__(self):
        self.h = _lib.hesic_rans_decoder_create()
        self._buf = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.hesic_rans_decoder_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_stream(self, data):
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        C.check(_lib.hesic_rans_decoder_set_stream(self.h, buf.ctypes.data, buf.size))

    def decode(self, indexes, cdfs, cdf_sizes, offsets):
        i, c, z, o = _i32(indexes).reshape(-1), _i32(cdfs), _i32(cdf_sizes), _i32(offsets)
        out = np.empty(i.size, dtype=np.int32)
        C.check(_lib.hesic_rans_decoder_decode(self.h, i.ctypes.data, i.size, c.ctypes.data, c.shape[0], c.shape[1],
                                               z.ctypes.data, o.ctypes.data, out.ctypes.data))
        return out


# ---------------------------------------------------------------------------------------------
# operators of the DSIC variant (ywz/DSIC/mynet6_plus.py)
@C.device_guard
def group_norm(x, groups, weight=None, bias=None, eps=1e-5, relu=False):
    """nn.GroupNorm(groups, C) (+ the ReLU that follows it everywhere in mynet6_plus.py:224-290)."""
    x = _f32(x)
    y = torch.empty_like(x)
    w = _f32(weight.detach()) if weight is not None else None
    b = _f32(bias.detach()) if bias is not None else None
    C.check(_lib.hesic_group_norm(C.ref(C.nchw(x)), C.ref(C.nchw(y)), int(groups), C.ptr(w), C.ptr(b), float(eps), int(relu),
                                  C.stream()))
    return y


@C.device_guard
def softmax_channels(x):
    """nn.functional.softmax(x, dim=-3) (mynet6_plus.py:311)."""
    x = _f32(x)
    y = torch.empty_like(x)
    C.check(_lib.hesic_softmax_channels(C.ref(C.nchw(x)), C.ref(C.nchw(y)), C.stream()))
    return y


@C.device_guard
def dense_warp(h1, cost):
    """dense_warp.forward (mynet6_plus.py:316-345): sum_d cost[:, d] * (h1 shifted left by d pixels)."""
    h1, cost = _f32(h1), _f32(cost)
    out = torch.empty_like(h1)
    C.check(_lib.hesic_dense_warp(C.ref(C.nchw(h1)), C.ref(C.nchw(cost)), C.ref(C.nchw(out)), C.stream()))
    return out


# ---------------------------------------------------------------------------------------------
# file codec of the stereo models (newnet1.py:823-1273)
@C.device_guard
def gmm_cdf_tables(scales, means, weights, K, channels, minmax, scale_bound=0.11):
    """Per-element cumulative-frequency rows of the K-component mixture (newnet1.py:934-978) for the given channels of a
    [1, K*M, H, W] parameter set: int32 [len(channels)*H*W, 2*minmax+2] on the device, rows in (channel, h, w) order."""
    scales, means = _f32(scales), _f32(means)
    w = _f32(weights).reshape(-1)
    M = scales.shape[1] // K
    if scales.shape[0] != 1 or w.numel() != K * M:
        raise ValueError("gmm_cdf_tables: one image at a time, weights with K*M entries")
    ch = torch.as_tensor(np.asarray(channels, dtype=np.int32).reshape(-1), device=scales.device)
    n = int(ch.numel()) * scales.shape[2] * scales.shape[3]
    out = torch.empty((n, 2 * int(minmax) + 2), device=scales.device, dtype=torch.int32)
    C.check(_lib.hesic_gmm_cdf_tables(C.ref(C.nchw(scales)), C.ref(C.nchw(means)), C.ptr(w), int(K), int(M), C.ptr(ch),
                                      int(ch.numel()), int(minmax), float(scale_bound), C.ptr(out), C.stream()))
    return out


class RangeEncoderHandle(_OwnedHandle):
    """Host range coder (csrc/coder.cpp): one symbol per cumulative-frequency row."""

    def __init__(self):
        self.h = _lib.hesic_range_encoder_create()

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.hesic_range_encoder_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def push(self, symbols, cdfs):
        s, c = _i32(symbols).reshape(-1), _i32(cdfs)
        if c.ndim != 2 or c.shape[0] != s.size:
            raise ValueError("range coder: one cumulative row per symbol")
        C.check(_lib.hesic_range_encoder_push(self.h, s.ctypes.data, s.size, c.ctypes.data, c.shape[1], c.shape[1]))

    def finish(self):
        n = _lib.hesic_range_encoder_finish(self.h, None, 0)
        if n < 0:
            C.check(int(n))
        buf = np.empty(n, dtype=np.uint8)
        assert _lib.hesic_range_encoder_finish(self.h, buf.ctypes.data, n) == n
        return buf.tobytes()


class RangeDecoderHandle(_OwnedHandle):
    def _ctor_args(self):
        return (self._buf.tobytes(),)

    def __init__(self, data):
        self._buf = np.frombuffer(bytes(data), dtype=np.uint8)
        self.h = _lib.hesic_range_decoder_create(self._buf.ctypes.data if self._buf.size else None, self._buf.size)
        if not self.h:
            raise ValueError(C.last_error())

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.hesic_range_decoder_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def decode(self, cdfs):
        c = _i32(cdfs)
        out = np.empty(c.shape[0], dtype=np.int32)
        C.check(_lib.hesic_range_decoder_decode(self.h, c.shape[0], c.ctypes.data, c.shape[1], c.shape[1], out.ctypes.data))
        return out
