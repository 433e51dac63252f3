"""Data-parallel sharding of stereo-pair batches: one process per GPU, pairs split
evenly, weights replicated, no data-path collective.  The only exchange is ONE all-reduce(sum) of the
rate/distortion partial sums per measured batch -- 6 fp64 values, 48 bytes:

    [sum log2 p(y1), sum log2 p(y2), sum log2 p(z1), sum log2 p(z2), SSE view 1, SSE view 2]

(the terms of RateDistortionLoss, ywz/mywork/test3real.py:90-124).  Works with any
``torch.distributed`` backend: NCCL over NVLink on the B200 box, gloo in the CPU tests.
"""
import math

import torch
import torch.distributed as dist

N_PARTIALS = 6


def shard_range(n_pairs, rank, world):
    """Contiguous [begin, end) slice of a global batch for ``rank`` (sizes differ by at most one)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_pairs, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def reduce_partials(partial, group=None):
    """In-place sum over ranks of the [6] fp64 partial-sum vector; the path's single collective."""
    if partial.numel() != N_PARTIALS or partial.dtype != torch.float64:
        raise ValueError("partials must be 6 float64 values")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def metrics_from_partials(partial, n_pairs, height, width):
    """bpp (total / per view) and PSNR per view from GLOBAL partial sums over ``n_pairs`` pairs."""
    p = [float(v) for v in partial]
    pixels = n_pairs * height * width
    mse1, mse2 = p[4] / (3 * pixels), p[5] / (3 * pixels)
    return {"bpp": -(p[0] + p[1] + p[2] + p[3]) / pixels, "bpp1": -(p[0] + p[2]) / pixels, "bpp2": -(p[1] + p[3]) / pixels,
            "mse1": mse1, "mse2": mse2,
            "psnr1": 10 * math.log10(1 / mse1) if mse1 > 0 else float("inf"),
            "psnr2": 10 * math.log10(1 / mse2) if mse2 > 0 else float("inf")}
