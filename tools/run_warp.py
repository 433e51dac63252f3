#!/usr/bin/env python
"""warp_perspective in isolation (B x 3 x 512 x 512, CUDA events, L2 flushed): python tools/run_warp.py [B] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import hesic_b200
from hesic_b200 import functional as F, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
x1, _, h = (t.cuda() for t in synth.stereo_pairs(B, 512, 512, seed=1234))
out = torch.empty_like(x1)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
ts = []
for i in range(reps + 1):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); F.warp_perspective(x1, h, (512, 512), True, out=out); e1.record(); torch.cuda.synchronize()
    if i: ts.append(e0.elapsed_time(e1))
ms = sum(ts) / len(ts)
print(f"warp_perspective {B} x 3 x 512 x 512: {ms * 1e3:.1f} us, {2 * x1.numel() * 4 / ms / 1e6:.0f} GB/s algorithmic, checksum {float(out.double().sum()):.6f}")
