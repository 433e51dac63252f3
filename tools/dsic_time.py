#!/usr/bin/env python
"""DSIC forward timing (BASELINE config 5: batch 8 x 512x512 on one B200), fused engine (hesic_b200/dsic_engine.py).

    python tools/dsic_time.py [B] [H] [W] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import hesic_b200
from hesic_b200 import _capi as C
from hesic_b200 import synth

hesic_b200.install()
import mynet6_plus  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H = int(sys.argv[2]) if len(sys.argv) > 2 else 512
W = int(sys.argv[3]) if len(sys.argv) > 3 else 512
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = "cuda:0"
net = mynet6_plus.DSIC(128, 192, 21, 32, 5).eval()
net.load_state_dict(synth.synth_state_dict(net, seed=0))
net = net.to(dev)
x1, x2, _ = (t.to(dev) for t in synth.stereo_pairs(B, H, W, seed=1234))
for _ in range(2):
    out = net(x1, x2)
torch.cuda.synchronize()
C.check(C.lib.hesic_tc_status())
C.lib.hesic_launch_count(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = net(x1, x2)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
m = synth.rd_metrics({"x1_hat": out["x1_hat"].cpu(), "x2_hat": out["x2_hat"].cpu(),
                      "likelihoods": {k: v.cpu() for k, v in out["likelihoods"].items()}}, x1.cpu(), x2.cpu())
print(f"DSIC B={B} {H}x{W}: {ms:.2f} ms/forward = {B / ms * 1e3:.1f} pairs/s, {C.lib.hesic_launch_count(0) // reps} kernels/forward, "
      f"{1366.5 * B * (H * W) / (512 * 512) / ms:.1f} TFLOP/s algorithmic (1366.5 GF/pair), peak mem "
      f"{torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, bpp={m['bpp']:.4f} psnr1={m['psnr1']:.3f} psnr2={m['psnr2']:.3f}")
