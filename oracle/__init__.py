"""Oracle package: CPU checkers for the hesic_b200 CUDA path.  TEST INFRASTRUCTURE ONLY.

Importers allowed: ``tests/``, ``__graft_entry__.smoke()``, and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  The product package
``hesic_b200`` must never import from here (tests/test_boundary.py enforces it).
"""
import ctypes
import importlib.machinery
import importlib.util
import os
import subprocess
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def build(ref=True):
    """Compile the C restatement and, when /root/reference is present, oracle/_ref."""
    targets = ["oracle"]
    if ref and os.path.isdir("/root/reference/compressai/cpp_exts"):
        targets.append("ref")
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


_coder = None


def coder():
    """ctypes handle on oracle/_build/liboracle_coder.so (coder_oracle.c)."""
    global _coder
    if _coder is None:
        path = os.path.join(HERE, "_build", "liboracle_coder.so")
        if not os.path.exists(path):
            build(ref=False)
        lib = ctypes.CDLL(path)
        i32p = ctypes.POINTER(ctypes.c_int32)
        lib.orc_pmf_to_quantized_cdf.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.c_int,
                                                 ctypes.POINTER(ctypes.c_uint32)]
        lib.orc_pmf_to_quantized_cdf.restype = ctypes.c_int
        lib.orc_rans_encode.argtypes = [i32p, i32p, ctypes.c_long, i32p, ctypes.c_int, i32p, i32p,
                                        ctypes.POINTER(ctypes.c_uint8)]
        lib.orc_rans_encode.restype = ctypes.c_long
        lib.orc_rans_decode.argtypes = [ctypes.POINTER(ctypes.c_uint8), i32p, ctypes.c_long, i32p, ctypes.c_int,
                                        i32p, i32p, i32p]
        lib.orc_rans_decode.restype = None
        _coder = lib
    return _coder


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def pmf_to_quantized_cdf(pmf, precision=16):
    pmf = np.ascontiguousarray(pmf, dtype=np.float32)
    cdf = np.zeros(pmf.size + 1, dtype=np.uint32)
    rc = coder().orc_pmf_to_quantized_cdf(pmf.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), pmf.size, precision,
                                          cdf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    if rc != 0:
        raise ValueError("pmf_to_quantized_cdf: no symbol can donate frequency")
    return cdf


def rans_encode(symbols, indexes, cdfs, cdf_sizes, offsets):
    s, sp = _i32(symbols)
    i, ip = _i32(indexes)
    c, cp = _i32(cdfs)
    z, zp = _i32(cdf_sizes)
    o, op = _i32(offsets)
    out = np.zeros(4 * (9 * s.size + 16), dtype=np.uint8)
    n = coder().orc_rans_encode(sp, ip, s.size, cp, c.shape[1], zp, op, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return out[:n].tobytes()


def rans_decode(stream, indexes, cdfs, cdf_sizes, offsets):
    i, ip = _i32(indexes)
    c, cp = _i32(cdfs)
    z, zp = _i32(cdf_sizes)
    o, op = _i32(offsets)
    buf = np.frombuffer(stream + b"\0" * 8, dtype=np.uint8).copy()
    out = np.zeros(i.size, dtype=np.int32)
    coder().orc_rans_decode(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), ip, i.size, cp, c.shape[1], zp, op,
                            out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return out


def ref_ext(name):
    """Load the reference's own compiled C++ module ('ans' or '_CXX') from oracle/_ref, or None."""
    path = os.path.join(HERE, "_ref", name + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.exists(path):
        return None
    loader = importlib.machinery.ExtensionFileLoader("compressai." + name, path)
    spec = importlib.util.spec_from_loader("compressai." + name, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod
