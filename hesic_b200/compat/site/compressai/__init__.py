"""hesic_b200 stand-in for the reference's ``compressai`` package (CompressAI 1.0.0 fork):
same module tree and names (compressai/__init__.py:15-60), operators backed by the
sm_100a kernels of libhesic_b200.so."""
from compressai import datasets, entropy_models, layers, models, ops

__version__ = "1.0.0+hesic_b200"

_entropy_coder = "ans"
_available_entropy_coders = [_entropy_coder]

try:
    import range_coder  # noqa: F401

    if getattr(range_coder, "RangeEncoder", None) is not None and not getattr(range_coder, "_HESIC_STUB", False):
        _available_entropy_coders.append("rangecoder")
except ImportError:
    pass


def set_entropy_coder(entropy_coder):
    """Select the default entropy coder (compressai/__init__.py:32-47)."""
    global _entropy_coder
    if entropy_coder not in _available_entropy_coders:
        raise ValueError(f'Invalid entropy coder "{entropy_coder}", choose from'
                         f'({", ".join(_available_entropy_coders)}).')
    _entropy_coder = entropy_coder


def get_entropy_coder():
    return _entropy_coder


def available_entropy_coders():
    return _available_entropy_coders
