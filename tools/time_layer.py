#!/usr/bin/env python
"""CUDA-event time of one conv layer shape on the tcgen05 path (L2 flushed between launches).

    python tools/time_layer.py Cin Cout k stride transposed H W B gdn act [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch

import tc_check as T
from hesic_b200 import _capi as C

a = [int(v) for v in sys.argv[1:11]]
reps = int(sys.argv[11]) if len(sys.argv) > 11 else 5
Cin, Cout, k, s, tr, H, W, B, gdn, act = a
mod = (T.deconv if tr else T.conv)(Cin, Cout, kernel_size=k, stride=s).to(T.DEV)
plan = mod.hesic_plan()
if gdn:
    plan.set_gdn(torch.ones(Cout, device=T.DEV), 0.1 * torch.eye(Cout, device=T.DEV) + 0.01, gdn == 2)
xd = T.to_split(torch.randn(B, Cin, H, W).to(T.DEV), s, tr)
Ho, Wo = plan.out_hw(H, W)
yt = torch.zeros((2, B, Ho, Wo, Cout), device=T.DEV, dtype=torch.bfloat16)
yd = C.split(yt)
flush = torch.empty(256 << 20, device=T.DEV, dtype=torch.uint8)
ts = []
for i in range(reps + 1):
    if not os.environ.get("NOFLUSH"):
        flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.run(xd, yd, act, C.PATH_TC)
    e1.record()
    torch.cuda.synchronize()
    if i:
        ts.append(e0.elapsed_time(e1))
C.check(C.lib.hesic_tc_status())
ms = sum(ts) / len(ts)
fl = 2.0 * B * Ho * Wo * Cout * Cin * k * k / (s * s if tr else 1)
print(f"{'deconv' if tr else 'conv'} {Cin}->{Cout} k{k} s{s} {H}x{W} B={B} gdn={gdn}: {ms * 1e3:.1f} us, {fl / ms / 1e9:.1f} TFLOP/s alg, "
      f"pair={os.environ.get('HESIC_TC_PAIR', '0')} checksum {float(yt.float().abs().sum()):.4e}")
