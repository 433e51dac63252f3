"""Placeholder for the un-vendored ``pytorch_msssim`` (caller-side metric in test3real.py:105-107)."""
_HESIC_STUB = True


def _unavailable(*a, **k):
    raise NotImplementedError("pytorch_msssim is not installed")


ssim = ms_ssim = _unavailable


class SSIM:
    def __init__(self, *a, **k):
        _unavailable()


MS_SSIM = SSIM
