"""``compressai._CXX`` (compressai/cpp_exts/ops/ops.cpp:83-90) on libhesic_b200.so."""
from hesic_b200.functional import pmf_to_quantized_cdf as _impl


def pmf_to_quantized_cdf(pmf, precision):
    return _impl(pmf, precision).tolist()
