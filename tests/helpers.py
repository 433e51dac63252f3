"""Shared test helpers: tolerances and golden loading."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def load_json(name):
    return json.load(open(os.path.join(GOLDEN, name + ".json")))


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def close_stats(a, b, floor=None):
    """max |a-b| / max(|b|, floor); floor defaults to rms(b) (the per-op bar of SURVEY 8d:
    abs(a-b) <= 1e-4 * max(abs(b), floor))."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    if floor is None:
        floor = float(b.pow(2).mean().sqrt()) + 1e-30
    return float(((a - b).abs() / torch.clamp(b.abs(), min=floor)).max())


def assert_close(a, b, tol=1e-4, floor=None, what=""):
    err = close_stats(a, b, floor)
    assert err <= tol, f"{what}: rel err {err:.3e} > {tol:.1e}"
    return err


def mismatch_fraction(a, b):
    a = torch.as_tensor(a).cpu()
    b = torch.as_tensor(b).cpu()
    return float((a != b).double().mean())
