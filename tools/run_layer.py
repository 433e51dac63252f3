#!/usr/bin/env python
"""Launch one conv layer shape on the tcgen05 path a few times (target for ncu captures).

    python tools/run_layer.py Cin Cout k stride transposed H W B gdn act [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch

import tc_check as T
from hesic_b200 import _capi as C

a = [int(v) for v in sys.argv[1:11]]
reps = int(sys.argv[11]) if len(sys.argv) > 11 else 3
Cin, Cout, k, s, tr, H, W, B, gdn, act = a
mod = (T.deconv if tr else T.conv)(Cin, Cout, kernel_size=k, stride=s).to(T.DEV)
plan = mod.hesic_plan()
if gdn:
    plan.set_gdn(torch.ones(Cout, device=T.DEV), 0.1 * torch.eye(Cout, device=T.DEV) + 0.01, gdn == 2)
xd = T.to_split(torch.randn(B, Cin, H, W).to(T.DEV), s, tr)
Ho, Wo = plan.out_hw(H, W)
if Cout <= 4:
    yt = torch.zeros((B, Cout, Ho, Wo), device=T.DEV); yd = C.nchw(yt)
else:
    yt = torch.zeros((2, B, Ho, Wo, Cout), device=T.DEV, dtype=torch.bfloat16); yd = C.split(yt)
for _ in range(reps):
    plan.run(xd, yd, act, C.PATH_TC)
torch.cuda.synchronize()
C.check(C.lib.hesic_tc_status())
print("ok")
