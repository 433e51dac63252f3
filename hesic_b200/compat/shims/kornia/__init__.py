"""Minimal stand-in for the parts of ``kornia`` the HESIC path calls, used only when the real
package is not installed.  ``warp_perspective`` runs the hesic_b200 bilinear-gather kernel with
kornia's normalise -> invert -> grid_sample arithmetic (align_corners=True convention, see
SURVEY.md section 8c); ``get_perspective_transform`` is the 4-point DLT solve used by the caller
(test3real.py:179): the library's kernel for CUDA tensors, plain torch otherwise."""
import torch

from hesic_b200 import functional as _F

_HESIC_STUB = True


def warp_perspective(src, M, dsize, flags="bilinear", border_mode="zeros", align_corners=True, **kwargs):
    if flags != "bilinear" or border_mode != "zeros":
        raise NotImplementedError("hesic_b200 warp_perspective: bilinear + zeros padding only")
    return _F.warp_perspective(src, M, dsize, align_corners=align_corners)


def get_perspective_transform(src, dst):
    """[B,4,2] point pairs -> [B,3,3] homography (direct linear transform): one kernel launch for CUDA fp32 corners
    (hesic_perspective_transform), the same system through torch.linalg.solve otherwise."""
    if src.is_cuda and src.dtype == torch.float32 and dst.dtype == torch.float32 and not (src.requires_grad or dst.requires_grad):
        return _F.perspective_transform(src, dst)
    B = src.shape[0]
    x, y = src[..., 0], src[..., 1]
    u, v = dst[..., 0], dst[..., 1]
    zeros, ones = torch.zeros_like(x), torch.ones_like(x)
    ax = torch.stack([x, y, ones, zeros, zeros, zeros, -x * u, -y * u], dim=-1)
    ay = torch.stack([zeros, zeros, zeros, x, y, ones, -x * v, -y * v], dim=-1)
    A = torch.cat([ax, ay], dim=1)
    b = torch.cat([u, v], dim=1).unsqueeze(-1)
    sol = torch.linalg.solve(A, b).squeeze(-1)
    H = torch.cat([sol, torch.ones(B, 1, dtype=sol.dtype, device=sol.device)], dim=1)
    return H.reshape(B, 3, 3)
