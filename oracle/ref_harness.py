"""Import the UNMODIFIED reference from /root/reference (build container only).

Used by tests/golden/make_golden.py to produce the committed fixtures and by
tests that are skipped when /root/reference is absent.  Nothing is written into
the reference tree; the missing third-party modules are provided as in-memory
stand-ins:

* ``compressai.ans`` / ``compressai._CXX``: the reference's own C++ compiled
  into oracle/_ref by oracle/Makefile;
* ``kornia``: ``warp_perspective`` from oracle/hesic_oracle.py (the reference
  does not vendor or pin kornia -- SURVEY.md 8c; parity for the warp is unpinned);
* ``range_coder``, ``pytorch_msssim``, ``matplotlib``, ``imageio``: empty
  stubs (only used by compress()/plotting, never by forward()).
"""
import os
import sys
import types

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "ywz", "mywork"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_done = False


def install():
    global _done
    if _done:
        return
    if not available():
        raise RuntimeError("/root/reference not present")
    import oracle
    from oracle import hesic_oracle as O

    oracle.build(ref=True)
    for ext in ("ans", "_CXX"):
        mod = oracle.ref_ext(ext)
        if mod is None:
            raise RuntimeError(f"oracle/_ref/{ext} missing")
        sys.modules["compressai." + ext] = mod

    import torch

    def get_perspective_transform(src, dst):
        raise NotImplementedError

    _stub("kornia", warp_perspective=lambda src, M, dsize, **kw: O.warp_perspective(src, M, dsize, True),
          get_perspective_transform=get_perspective_transform)

    class _NA:
        def __init__(self, *a, **k):
            raise NotImplementedError("range_coder is not available")

    _stub("range_coder", RangeEncoder=_NA, RangeDecoder=_NA, prob_to_cum_freq=None)
    _stub("pytorch_msssim", ssim=None, ms_ssim=None, SSIM=None, MS_SSIM=None)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mp = _stub("matplotlib")
            mp.pyplot = _stub("matplotlib.pyplot")
    try:
        import imageio  # noqa: F401
    except ImportError:
        _stub("imageio")
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "ywz", "mywork"))
    import compressai  # noqa: F401  (the reference's package)

    assert compressai.__file__.startswith(REF), compressai.__file__
    _done = True


def load(module_name):
    """Import a reference model file (newnet1, newnet1_joint, newnet9, mynet6_plus, model) and return the module."""
    install()
    import importlib.util

    paths = {
        "newnet1": os.path.join(REF, "ywz/mywork/newnet1.py"),
        "newnet1_joint": os.path.join(REF, "ywz/mywork/newnet1_joint.py"),
        "newnet9": os.path.join(REF, "ywz/mywork/.trash/newnet9.py"),
        "mynet6_plus": os.path.join(REF, "ywz/DSIC/mynet6_plus.py"),
        "model": os.path.join(REF, "ywz/mywork/model.py"),
    }
    spec = importlib.util.spec_from_file_location("_ref_" + module_name, paths[module_name])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
