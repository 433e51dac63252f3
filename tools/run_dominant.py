#!/usr/bin/env python
"""Launch the dominant kernel of the path in isolation (for `ncu --set full`): g_a_conv2 of Encoder1,
conv 128->128 k5 s2 + fused GDN on 16 x 256x256 split-bf16 activations (newnet1.py:585,593-596).

    python tools/run_dominant.py [reps] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import hesic_b200
from hesic_b200 import _capi as C

hesic_b200.install()
from compressai.models.utils import conv  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = "cuda:0"
torch.manual_seed(0)
layer = conv(128, 128, kernel_size=5, stride=2).to(dev)
plan = layer.hesic_plan()
plan.set_gdn(torch.ones(128, device=dev), 0.1 * torch.eye(128, device=dev) + 0.01, False)
x = torch.randn(2, B, 256, 256, 128, device=dev).to(torch.bfloat16)
y = torch.empty(2, B, 128, 128, 128, device=dev, dtype=torch.bfloat16)
for _ in range(reps):
    plan.run(C.split(x), C.split(y), C.ACT_NONE, C.PATH_TC)
torch.cuda.synchronize()
C.check(C.lib.hesic_tc_status())
print("ok", float(y[0].float().abs().mean()))
