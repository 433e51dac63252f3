"""ctypes binding of include/hesic_b200.h (the C-ABI drop-in boundary).

PyTorch is used here only as the owner of device memory and streams: every
call passes raw device pointers, sizes and the current CUDA stream handle.
There is no fallback: if libhesic_b200.so is missing or cannot be loaded the
import fails loudly, and device entry points raise when handed CPU tensors.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_uint32, c_void_p

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libhesic_b200.so")

FMT_NCHW, FMT_NHWC, FMT_SPLIT, FMT_ROWPAD, FMT_HILO = 0, 1, 2, 3, 4
ROWPAD_Y, ROWPAD_X = 4, 8
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
PATH_AUTO, PATH_SIMT, PATH_TC = 0, 1, 2
OP_COPY, OP_ABS, OP_ROUND = 0, 1, 2
EB_PARAMS = 60
GN_SLOTS = 8


class HesicError(RuntimeError):
    pass


class CTensor(ctypes.Structure):
    _fields_ = [("p0", c_void_p), ("p1", c_void_p), ("fmt", c_int32), ("B", c_int32), ("C", c_int32),
                ("H", c_int32), ("W", c_int32), ("Cs", c_int32)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"hesic_b200: {LIB_PATH} not found. Build it with `python -m hesic_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
    return ctypes.CDLL(LIB_PATH)


lib = _load()
_TP = POINTER(CTensor)
_FP = POINTER(c_float)
_sig = {
    "hesic_abi_version": ([], c_int),
    "hesic_last_error": ([], c_char_p),
    "hesic_device_check": ([c_char_p, c_int], c_int),
    "hesic_launch_count": ([c_int], c_int64),
    "hesic_conv_create": ([c_int] * 8, c_void_p),
    "hesic_conv_destroy": ([c_void_p], None),
    "hesic_conv_load": ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "hesic_conv_set_gdn": ([c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p], c_int),
    "hesic_conv_enable_gdn": ([c_void_p, c_int], c_int),
    "hesic_conv_detect_kband": ([c_void_p, c_void_p, c_void_p], c_int),
    "hesic_conv_forward": ([c_void_p, _TP, _TP, c_int, c_int, c_void_p], c_int),
    "hesic_conv_forward_cat": ([c_void_p, _TP, _TP, _TP, c_int, c_int, c_void_p], c_int),
    "hesic_conv_forward_sse": ([c_void_p, _TP, _TP, _TP, c_int, c_int, _TP, c_void_p, c_void_p], c_int),
    "hesic_en_conv_create": ([c_int, c_int], c_void_p),
    "hesic_en_conv_destroy": ([c_void_p], None),
    "hesic_en_conv_load": ([c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "hesic_en_conv_forward": ([c_void_p, _TP, _TP, c_int, _TP, _TP, c_void_p], c_int),
    "hesic_en_pack_input": ([_TP, _TP, _TP, c_void_p], c_int),
    "hesic_tc_status": ([], c_int),
    "hesic_gdn": ([_TP, _TP, c_void_p, c_void_p, c_int, c_float, c_void_p], c_int),
    "hesic_warp_perspective": ([_TP, c_void_p, _TP, _TP, c_int, c_void_p], c_int),
    "hesic_perspective_transform": ([c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p], c_int),
    "hesic_max_pool2x2": ([_TP, _TP, c_void_p], c_int),
    "hesic_eb_pack": ([POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), c_void_p, c_int, c_void_p, c_void_p], c_int),
    "hesic_entropy_bottleneck": ([_TP, c_void_p, c_float, _TP, _TP, c_void_p, c_void_p], c_int),
    "hesic_gaussian_conditional": ([_TP, _TP, _TP, c_void_p, c_int, c_int, c_float, c_float, _TP, _TP, _TP, c_void_p, c_void_p], c_int),
    "hesic_spatial_max": ([_TP, c_void_p, c_void_p], c_int),
    "hesic_mixture_weights": ([c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p], c_int),
    "hesic_upsample_bilinear": ([_TP, _TP, c_int, c_void_p], c_int),
    "hesic_convert": ([_TP, _TP, c_int, c_void_p], c_int),
    "hesic_images_from_u8": ([c_void_p, c_int, c_int, c_int, c_int, _TP, c_void_p], c_int),
    "hesic_group_norm": ([_TP, _TP, c_int, c_void_p, c_void_p, c_float, c_int, c_void_p], c_int),
    "hesic_conv_forward_gn": ([c_void_p, _TP, _TP, c_int, c_void_p, c_int, c_void_p], c_int),
    "hesic_group_norm_apply": ([_TP, _TP, c_int, c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p], c_int),
    "hesic_softmax_channels": ([_TP, _TP, c_void_p], c_int),
    "hesic_dense_warp": ([_TP, _TP, _TP, c_void_p], c_int),
    "hesic_prepare_symbols": ([_TP, c_void_p, _TP, c_void_p, c_void_p], c_int),
    "hesic_build_indexes_channel": ([c_int, c_int, c_int, c_int, c_void_p, c_void_p], c_int),
    "hesic_build_indexes_scale": ([_TP, c_void_p, c_int, c_float, c_void_p, c_void_p], c_int),
    "hesic_sum_squared_error": ([_TP, _TP, c_void_p, c_void_p], c_int),
    "hesic_gmm_cdf_tables": ([_TP, _TP, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p], c_int),
    "hesic_range_encoder_create": ([], c_void_p),
    "hesic_range_encoder_destroy": ([c_void_p], None),
    "hesic_range_encoder_push": ([c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int], c_int),
    "hesic_range_encoder_finish": ([c_void_p, c_void_p, c_int64], c_int64),
    "hesic_range_decoder_create": ([c_void_p, c_int64], c_void_p),
    "hesic_range_decoder_destroy": ([c_void_p], None),
    "hesic_range_decoder_decode": ([c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p], c_int),
    "hesic_pmf_to_quantized_cdf": ([_FP, c_int, c_int, POINTER(c_uint32)], c_int),
    "hesic_rans_encoder_create": ([], c_void_p),
    "hesic_rans_encoder_destroy": ([c_void_p], None),
    "hesic_rans_encoder_push": ([c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p], c_int),
    "hesic_rans_encoder_flush": ([c_void_p, c_void_p, c_int64], c_int64),
    "hesic_rans_decoder_create": ([], c_void_p),
    "hesic_rans_decoder_destroy": ([c_void_p], None),
    "hesic_rans_decoder_set_stream": ([c_void_p, c_void_p, c_int64], c_int),
    "hesic_rans_decoder_decode": ([c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p], c_int),
}
EXPORTS = tuple(_sig)
for _name, (_args, _res) in _sig.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = _res

if lib.hesic_abi_version() != 1:
    raise ImportError("hesic_b200: ABI version mismatch between _capi.py and libhesic_b200.so")


def last_error():
    return (lib.hesic_last_error() or b"").decode()


def check(rc):
    if rc == 0:
        return
    msg = last_error()
    if rc == -1:
        raise ValueError(msg)
    if rc == -3:
        raise NotImplementedError(msg)
    raise HesicError(msg)


def stream():
    """Current stream of the CURRENT device: callers run under ``device_guard`` / ``torch.cuda.device(t.device)``, so
    that is the device of the tensors being passed."""
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _first_cuda_tensor(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a
        elif isinstance(a, (list, tuple)):
            for b in a:
                if isinstance(b, torch.Tensor) and b.is_cuda:
                    return b
    return None


def device_guard(fn):
    """Run ``fn`` with the device of its first CUDA tensor argument current: the library allocates its packed operands
    and launches on the current device / its current stream, so a model living on cuda:1 while cuda:0 is current must
    switch for the duration of the call (as torch's own operators do)."""
    import functools

    @functools.wraps(fn)
    def inner(*args, **kwargs):
        t = _first_cuda_tensor(args, kwargs)
        if t is None or t.device.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(t.device):
            return fn(*args, **kwargs)
    return inner


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("hesic_b200: CUDA tensors required -- this path has no CPU fallback "
                               "(the CPU implementation is the reference itself)")


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def nchw(t, C=None, c0=0):
    """Descriptor for a contiguous NCHW fp32 torch tensor (optionally a channel slice [c0, c0+C))."""
    assert t.dtype == torch.float32 and t.is_contiguous(), "expected contiguous fp32"
    B, Cs, H, W = t.shape
    C = Cs - c0 if C is None else C
    return CTensor(t.data_ptr() + 4 * c0 * H * W, None, FMT_NCHW, B, C, H, W, Cs)


def nhwc(t, C=None, c0=0):
    """Descriptor for a contiguous [B,H,W,C] fp32 torch tensor (NHWC storage)."""
    assert t.dtype == torch.float32 and t.is_contiguous()
    B, H, W, Cs = t.shape
    C = Cs - c0 if C is None else C
    return CTensor(t.data_ptr() + 4 * c0, None, FMT_NHWC, B, C, H, W, Cs)


def split(t, C=None, c0=0):
    """Descriptor for a [2,B,H,W,C] bf16 torch tensor holding the (hi, lo) planes."""
    assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.shape[0] == 2
    _, B, H, W, Cs = t.shape
    C = Cs - c0 if C is None else C
    plane = B * H * W * Cs * 2
    return CTensor(t.data_ptr() + 2 * c0, t.data_ptr() + plane + 2 * c0, FMT_SPLIT, B, C, H, W, Cs)


def rowpad(t, C=None, c0=0):
    """Descriptor for a [2,B,H+4,W+8,S] bf16 torch tensor in ROWPAD split format, S = 8 or 4 channel slots (image
    at offset (2,2), zero border and zero unused channel slots): the input format of the Cin <= 8 edge layers."""
    assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.shape[0] == 2 and t.shape[-1] in (4, 8)
    _, B, Hp, Wp, S = t.shape
    C = S - c0 if C is None else C
    plane = B * Hp * Wp * S * 2
    return CTensor(t.data_ptr() + 2 * c0, t.data_ptr() + plane + 2 * c0, FMT_ROWPAD, B, C, Hp - ROWPAD_Y, Wp - ROWPAD_X, S)


def hilo(t, C=None):
    """Descriptor for a [B,H,W,2*S] bf16 torch tensor in NHWC_HILO format (per pixel S 'hi' then S 'lo' values):
    the activation format of the enhancement network (S = 32)."""
    assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.dim() == 4 and t.shape[-1] % 2 == 0
    B, H, W, S2 = t.shape
    return CTensor(t.data_ptr(), None, FMT_HILO, B, S2 // 2 if C is None else C, H, W, S2 // 2)


def rowpad_slots(Cin, stride, transposed):
    """Channel slots of the ROWPAD input a k5 layer with Cin <= 8 reads on the tensor-core path (conv.h: ROW2 / ROW)."""
    return 4 if (Cin <= 4 and stride == 2 and not transposed) else 8


def null():
    return CTensor(None, None, 0, 0, 0, 0, 0, 0)


def ref(ct):
    return ctypes.byref(ct)
