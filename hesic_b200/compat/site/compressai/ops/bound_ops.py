"""LowerBound (reference: compressai/ops/bound_ops.py:19-52): ``max(x, bound)``.

On the inference path this is a plain clamp (the fused GDN / likelihood kernels apply the same
clamp on the device from the raw parameter); the module exists for the ``bound`` buffer that the
reference's checkpoints carry (``*.lower_bound.bound``, ``likelihood_lower_bound.bound`` ...).
When autograd is recording, the reference's gradient rule is kept so the operator surface stays
complete: a clamped entry still receives the gradient if the step would raise it (grad < 0).
"""
import torch
import torch.nn as nn


class _ClampFromBelow(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x < bound)
        return torch.maximum(x, bound)

    @staticmethod
    def backward(ctx, grad):
        (clamped,) = ctx.saved_tensors
        # drop only what would push an already clamped entry further below the bound
        return grad.masked_fill(clamped & (grad >= 0), 0), None


LowerBoundFunction = _ClampFromBelow   # the reference's name for it


class LowerBound(nn.Module):
    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.tensor([float(bound)], dtype=torch.float32))

    def forward(self, x):
        if torch.is_grad_enabled() and x.requires_grad:
            return _ClampFromBelow.apply(x, self.bound)
        return torch.maximum(x, self.bound)
