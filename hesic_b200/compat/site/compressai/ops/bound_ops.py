"""LowerBound (compressai/ops/bound_ops.py:19-52): max(x, bound) whose gradient passes
through when x is above the bound or is being pushed up towards it.  Forward semantics are what
the inference path needs (the fused kernels apply the same clamp on the device); the custom
backward is kept so the operator surface stays complete."""
import torch
import torch.nn as nn


class LowerBoundFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input_, bound):
        ctx.save_for_backward(input_, bound)
        return torch.max(input_, bound)

    @staticmethod
    def backward(ctx, grad_output):
        input_, bound = ctx.saved_tensors
        keep = (input_ >= bound) | (grad_output < 0)
        return grad_output * keep.to(grad_output.dtype), None


class LowerBound(nn.Module):
    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    @torch.jit.unused
    def lower_bound(self, x):
        return LowerBoundFunction.apply(x, self.bound)

    def forward(self, x):
        if torch.jit.is_scripting():
            return torch.max(x, self.bound)
        return self.lower_bound(x)
