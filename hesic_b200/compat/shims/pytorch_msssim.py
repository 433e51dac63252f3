"""Stand-in for the un-vendored ``pytorch_msssim`` package, used only when the real one is not installed: (MS-)SSIM of
Wang et al. (2003/2004) with that package's calling surface and defaults, because the reference's criterion calls
``ms_ssim(x_hat, target, data_range=1, size_average=False)`` on every batch (ywz/mywork/test3real.py:107-108).

Plain torch (works on CPU and CUDA tensors alike): this is the caller's quality metric, not part of the forward hot
path.  Definitions: 11-tap Gaussian window (sigma 1.5) applied separably without padding, K = (0.01, 0.03), five
scales weighted (0.0448, 0.2856, 0.3001, 0.2363, 0.1333), 2x2 average pooling between scales; contrast-structure terms
of the coarser scales and the final SSIM term are clamped at zero before the weighted product."""
import torch
import torch.nn.functional as F

_HESIC_STUB = True
_MS_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def _window(size, sigma, channels, like):
    x = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-x * x / (2.0 * sigma * sigma))
    g = (g / g.sum()).to(device=like.device, dtype=like.dtype)
    return g.reshape(1, 1, 1, size).repeat(channels, 1, 1, 1)


def _blur(t, win):
    """Separable 'valid' Gaussian filter; a spatial dim shorter than the window is left unfiltered."""
    ch, k = t.shape[1], win.shape[-1]
    if t.shape[-1] >= k:
        t = F.conv2d(t, win, groups=ch)
    if t.shape[-2] >= k:
        t = F.conv2d(t, win.transpose(-1, -2), groups=ch)
    return t


def _ssim_terms(x, y, data_range, win, K):
    c1, c2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mx, my = _blur(x, win), _blur(y, win)
    vx = _blur(x * x, win) - mx * mx
    vy = _blur(y * y, win) - my * my
    cxy = _blur(x * y, win) - mx * my
    cs = (2.0 * cxy + c2) / (vx + vy + c2)
    full = (2.0 * mx * my + c1) / (mx * mx + my * my + c1) * cs
    return full.flatten(2).mean(-1), cs.flatten(2).mean(-1)          # [B, C] each


def _check(x, y, win_size):
    if x.shape != y.shape:
        raise ValueError(f"Input images should have the same dimensions, but got {tuple(x.shape)} and {tuple(y.shape)}.")
    if x.dim() != 4:
        raise ValueError(f"Input images should be 4-d tensors (N,C,H,W), but got {tuple(x.shape)}")
    if win_size % 2 != 1:
        raise ValueError("Window size should be odd.")


def ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None, K=(0.01, 0.03),
         nonnegative_ssim=False):
    _check(X, Y, win_size)
    w = _window(win_size, win_sigma, X.shape[1], X) if win is None else win.to(X)
    val, _ = _ssim_terms(X, Y, data_range, w, K)
    if nonnegative_ssim:
        val = torch.relu(val)
    return val.mean() if size_average else val.mean(1)


def ms_ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None, weights=None, K=(0.01, 0.03)):
    _check(X, Y, win_size)
    if min(X.shape[-2:]) <= (win_size - 1) * 2 ** 4:
        raise AssertionError("Image size should be larger than %d due to the 4 downsamplings in ms-ssim" % ((win_size - 1) * 2 ** 4))
    wts = torch.as_tensor(_MS_WEIGHTS if weights is None else weights, device=X.device, dtype=X.dtype)
    w = _window(win_size, win_sigma, X.shape[1], X) if win is None else win.to(X)
    terms = []
    for level in range(wts.numel()):
        full, cs = _ssim_terms(X, Y, data_range, w, K)
        if level + 1 < wts.numel():
            terms.append(torch.relu(cs))
            pad = [s % 2 for s in X.shape[2:]]
            X, Y = F.avg_pool2d(X, 2, padding=pad), F.avg_pool2d(Y, 2, padding=pad)
        else:
            terms.append(torch.relu(full))
    val = torch.prod(torch.stack(terms, 0) ** wts.reshape(-1, 1, 1), dim=0)     # [B, C]
    return val.mean() if size_average else val.mean(1)


class SSIM(torch.nn.Module):
    def __init__(self, data_range=255, size_average=True, win_size=11, win_sigma=1.5, channel=3, spatial_dims=2, K=(0.01, 0.03),
                 nonnegative_ssim=False):
        super().__init__()
        self.kw = dict(data_range=data_range, size_average=size_average, win_size=win_size, win_sigma=win_sigma, K=K,
                       nonnegative_ssim=nonnegative_ssim)

    def forward(self, X, Y):
        return ssim(X, Y, **self.kw)


class MS_SSIM(torch.nn.Module):
    def __init__(self, data_range=255, size_average=True, win_size=11, win_sigma=1.5, channel=3, spatial_dims=2, weights=None,
                 K=(0.01, 0.03)):
        super().__init__()
        self.kw = dict(data_range=data_range, size_average=size_average, win_size=win_size, win_sigma=win_sigma, weights=weights, K=K)

    def forward(self, X, Y):
        return ms_ssim(X, Y, **self.kw)
