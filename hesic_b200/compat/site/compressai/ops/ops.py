import torch


def ste_round(x):
    """Round with a straight-through gradient (compressai/ops/ops.py:18-31)."""
    return x + (torch.round(x) - x).detach()
