"""Placeholder for the un-vendored PyPI ``range_coder`` (only the file codec of
newnet1.py:823-1273 uses it)."""
_HESIC_STUB = True


class _Unavailable:
    def __init__(self, *a, **k):
        raise NotImplementedError("range_coder is not installed")


RangeEncoder = RangeDecoder = _Unavailable


def prob_to_cum_freq(*a, **k):
    raise NotImplementedError("range_coder is not installed")
