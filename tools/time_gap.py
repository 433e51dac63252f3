#!/usr/bin/env python
"""Kernel-to-kernel gap of the tcgen05 conv kernels: N back-to-back launches of one layer on one stream against N launches
with a synchronisation in between (both without an L2 flush, same cache state).

    python tools/time_gap.py Cin Cout k stride transposed H W B gdn act [N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch

import tc_check as T
from hesic_b200 import _capi as C

a = [int(v) for v in sys.argv[1:11]]
N = int(sys.argv[11]) if len(sys.argv) > 11 else 40
Cin, Cout, k, s, tr, H, W, B, gdn, act = a
mod = (T.deconv if tr else T.conv)(Cin, Cout, kernel_size=k, stride=s).to(T.DEV)
plan = mod.hesic_plan()
if gdn:
    plan.set_gdn(torch.ones(Cout, device=T.DEV), 0.1 * torch.eye(Cout, device=T.DEV) + 0.01, gdn == 2)
xd = T.to_split(torch.randn(B, Cin, H, W).to(T.DEV), s, tr)
Ho, Wo = plan.out_hw(H, W)
yt = torch.zeros((2, B, Ho, Wo, Cout), device=T.DEV, dtype=torch.bfloat16)
yd = C.split(yt)
for _ in range(3):
    plan.run(xd, yd, act, C.PATH_TC)
torch.cuda.synchronize()
single = []
for _ in range(N):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pad = torch.empty(1 << 20, device=T.DEV).zero_()      # keeps the GPU busy while the host enqueues the launch
    e0.record()
    plan.run(xd, yd, act, C.PATH_TC)
    e1.record()
    torch.cuda.synchronize()
    single.append(e0.elapsed_time(e1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
pad = torch.empty(64 << 20, device=T.DEV).zero_()
e0.record()
for _ in range(N):
    plan.run(xd, yd, act, C.PATH_TC)
e1.record()
torch.cuda.synchronize()
C.check(C.lib.hesic_tc_status())
s_us = 1e3 * sorted(single)[len(single) // 2]
b_us = 1e3 * e0.elapsed_time(e1) / N
print(f"{'deconv' if tr else 'conv'} {Cin}->{Cout} {H}x{W} B={B} gdn={gdn}: single {s_us:.1f} us, back-to-back {b_us:.1f} us per launch, gap {b_us - s_us:+.1f} us")
