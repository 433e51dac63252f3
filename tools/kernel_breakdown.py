#!/usr/bin/env python
"""Per-kernel device-time totals of one forward (torch.profiler / CUPTI), for the models that still run at
operator level.

    python tools/kernel_breakdown.py dsic|en|hesic [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

import hesic_b200
from hesic_b200 import synth

hesic_b200.install()
which = sys.argv[1] if len(sys.argv) > 1 else "dsic"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = "cuda:0"
x1, x2, h = (t.to(dev) for t in synth.stereo_pairs(B, 512, 512, seed=1234))
if which == "dsic":
    import mynet6_plus
    net = mynet6_plus.DSIC(128, 192, 21, 32, 5).eval()
    args = (x1, x2)
elif which == "en":
    import newnet1
    net = newnet1.Independent_EN().eval()
    args = (x1, x2, h)
else:
    import newnet1
    net = newnet1.HSIC(128, 192, 5).eval()
    args = (x1, x2, h)
net.load_state_dict(synth.synth_state_dict(net, seed=0))
net = net.to(dev)
for _ in range(2):
    net(*args)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    net(*args)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"{which} B={B} 512x512: {ms:.2f} ms/forward = {B / ms * 1e3:.1f} pairs/s, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    net(*args)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
