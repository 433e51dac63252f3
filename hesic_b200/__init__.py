"""hesic_b200 -- Blackwell-native (sm_100a) forward path of the HESIC stereo image codecs.

    import hesic_b200
    hesic_b200.install()          # make `compressai`, `newnet1`, ... resolve to this package
    from newnet1 import HSIC

The compute lives in hesic_b200/lib/libhesic_b200.so (C ABI: include/hesic_b200.h); importing
``hesic_b200._capi`` fails loudly when that library has not been built.
"""
from .compat import install  # noqa: F401

__version__ = "0.1.0"
