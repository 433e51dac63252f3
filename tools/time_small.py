#!/usr/bin/env python
"""CUDA-event time of the full-resolution 6 -> 3 stencil layers (conv_small_kernel), L2 flushed between launches, and a
checksum of the output (bit-identical builds print identical checksums).

    python tools/time_small.py [B] [H] [W] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch

import tc_check as T
from hesic_b200 import _capi as C

B, H, W, reps = ([int(v) for v in sys.argv[1:5]] + [16, 512, 512, 10][len(sys.argv) - 1:])[:4]
g = torch.Generator().manual_seed(5)
flush = torch.empty(256 << 20, device=T.DEV, dtype=torch.uint8)
for tr in (False, True):
    mod = (T.deconv if tr else T.conv)(6, 3, kernel_size=5, stride=1)
    mod.load_state_dict({"weight": torch.randn(mod.weight.shape, generator=g) * 0.1, "bias": torch.randn(3, generator=g) * 0.1})
    mod = mod.to(T.DEV)
    plan = mod.hesic_plan()
    if not tr:
        plan.set_gdn((torch.rand(3, generator=g) + 0.5).to(T.DEV), (torch.rand(3, 3, generator=g) * 0.2).to(T.DEV), False)
    xa, xb = torch.randn(B, 3, H, W, generator=g).to(T.DEV), torch.randn(B, 3, H, W, generator=g).to(T.DEV)
    out = torch.zeros((B, 3, H, W), device=T.DEV)
    ts = []
    for i in range(reps + 1):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run(C.nchw(xa), C.nchw(out), C.ACT_NONE, C.PATH_AUTO, C.nchw(xb))
        e1.record()
        torch.cuda.synchronize()
        if i:
            ts.append(e0.elapsed_time(e1))
    us = 1e3 * sum(ts) / len(ts)
    mb = (xa.numel() * 2 + out.numel()) * 4 / 1e6
    print(f"{'deconv' if tr else 'conv+gdn'} 6->3 k5 s1 {H}x{W} B={B}: {us:.1f} us, {mb / us * 1e-3:.2f} TB/s algorithmic, "
          f"checksum {float(out.double().sum()):.10e} {float(out.double().abs().sum()):.10e}")
