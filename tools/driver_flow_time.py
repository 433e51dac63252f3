#!/usr/bin/env python
"""Throughput of the whole per-batch body of test_epoch in ywz/mywork/test3real.py:171-186 (homography net ->
get_perspective_transform -> inverse -> h_adjust -> HSIC (newnet9) -> Independent_EN) at batch B, 512x512.

    python tools/driver_flow_time.py [B] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import hesic_b200
from hesic_b200 import synth

hesic_b200.install()
import kornia  # noqa: E402
import newnet9  # noqa: E402
from model import Net  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = "cuda:0"
homo = Net(patch_size=128).eval()
sd = synth.synth_state_dict(homo, seed=0)
sd["fc.5.weight"] *= 0.01
sd["fc.5.bias"] *= 0.01
homo.load_state_dict(sd)
net = newnet9.HSIC(128, 192, 5).eval()
net.load_state_dict(synth.synth_state_dict(net, seed=0))
en = newnet9.Independent_EN().eval()
en.load_state_dict(synth.synth_state_dict(en, seed=0))
homo, net, en = homo.to(dev), net.to(dev), en.to(dev)
d1, d2, _ = (t.to(dev) for t in synth.stereo_pairs(B, 512, 512, seed=1234))
g1 = torch.nn.functional.interpolate(d1.mean(1, keepdim=True), size=(128, 128), mode="bilinear", align_corners=False)
g2 = torch.nn.functional.interpolate(d2.mean(1, keepdim=True), size=(128, 128), mode="bilinear", align_corners=False)
corners = torch.tensor([[[64., 64.], [191., 64.], [191., 191.], [64., 191.]]], device=dev).repeat(B, 1, 1)


def body():
    with torch.no_grad():
        c0 = corners - corners[:, 0].view(-1, 1, 2)
        delta = homo(g1, g2)
        h = torch.inverse(kornia.get_perspective_transform(c0, c0 + delta))
        a = d1.shape[-2] / 256
        h[:, 0, :] = a * h[:, 0, :]
        h[:, :, 0] = (1. / a) * h[:, :, 0]
        h[:, 1, :] = a * h[:, 1, :]
        h[:, :, 1] = (1. / a) * h[:, :, 1]
        out = net(d1, d2, h)
        return en(out["x1_hat"], out["x2_hat"], h)


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(body)
ms_h = timed(lambda: homo(g1, g2))
print(f"test3real.py per-batch body, B={B} 512x512: {ms:.2f} ms = {B / ms * 1e3:.1f} pairs/s "
      f"(homography net alone {ms_h:.2f} ms)")
