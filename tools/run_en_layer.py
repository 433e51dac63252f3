#!/usr/bin/env python
"""One enhancement-network conv3x3 (32 -> 32, NHWC_HILO, 512x512) in isolation: CUDA-event time per epilogue mode,
or a single launch for `ncu --set full`.

    python tools/run_en_layer.py [B] [reps] [mode: all|plain|lrelu|res1|res2|first|last]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import hesic_b200
from hesic_b200 import _capi as C
from hesic_b200.enhance import EnConvPlan

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
which = sys.argv[3] if len(sys.argv) > 3 else "all"
dev = "cuda:0"
torch.manual_seed(0)
H = W = 512
mk = lambda: (torch.randn(B, H, W, 64, device=dev) * 0.5).to(torch.bfloat16)
x, r1, r2, y = mk(), mk(), mk(), mk()
img = torch.rand(B, 3, H, W, device=dev)
out = torch.empty_like(img)
p32 = EnConvPlan(32, 32).load(torch.randn(32, 32, 3, 3, device=dev) * 0.08, torch.randn(32, device=dev) * 0.1)
p6 = EnConvPlan(6, 32).load(torch.randn(32, 6, 3, 3, device=dev) * 0.2, torch.randn(32, device=dev) * 0.1)
p3 = EnConvPlan(32, 3).load(torch.randn(3, 32, 3, 3, device=dev) * 0.08, torch.randn(3, device=dev) * 0.1)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
modes = {
    "plain": lambda: p32.run(C.hilo(x), C.hilo(y)),
    "lrelu": lambda: p32.run(C.hilo(x), C.hilo(y), C.ACT_LEAKY),
    "res1": lambda: p32.run(C.hilo(x), C.hilo(y), C.ACT_LEAKY, C.hilo(r1)),
    "res2": lambda: p32.run(C.hilo(x), C.hilo(y), C.ACT_LEAKY, C.hilo(r1), C.hilo(r2)),
    "first": lambda: p6.run(C.hilo(x), C.hilo(y)),
    "last": lambda: p3.run(C.hilo(x), C.nchw(out), C.ACT_NONE, C.nchw(img)),
}
flop = 2.0 * B * H * W * 32 * 32 * 9
for name, fn in modes.items():
    if which not in ("all", name):
        continue
    ts = []
    for i in range(reps + 1):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i:
            ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    nbytes = B * H * W * 128 * (2 + (name == "res1") + 2 * (name == "res2"))
    print(f"{name:6s} {ms * 1e3:8.1f} us  {flop / ms / 1e9:7.1f} TFLOP/s alg  {nbytes / ms / 1e6:7.1f} GB/s algorithmic")
C.check(C.lib.hesic_tc_status())
