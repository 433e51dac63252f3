#!/usr/bin/env python
"""Per-launch device times of one HSIC.forward (CUDA events around every C-ABI call), grouped by layer shape.

    python tools/layer_times.py [B] [model: hesic|hesic_plus] [reps]

Events bracket each library call on the launching stream, so a row's time is that kernel's duration plus
any host-side launch gap that preceded it; the table also prints the whole-forward time measured without
the per-call events for comparison."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import hesic_b200
from hesic_b200 import _capi as C
from hesic_b200 import functional as F
from hesic_b200 import synth

hesic_b200.install()

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
model = sys.argv[2] if len(sys.argv) > 2 else "hesic"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
mod = __import__("newnet1_joint" if model == "hesic_plus" else "newnet1")
net = mod.HSIC(128, 192, 5).eval()
net.load_state_dict(synth.synth_state_dict(net, seed=0))
net = net.to(dev)
net.hesic_engine.two_streams = False     # one stream: the per-call events then bracket exactly one kernel
x1, x2, h = (t.to(dev) for t in synth.stereo_pairs(B, 512, 512, seed=1234))

for _ in range(2):
    net(x1, x2, h)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    net(x1, x2, h)
e1.record()
torch.cuda.synchronize()
plain_ms = e0.elapsed_time(e1) / reps

records = []


def wrap(name, fn, tag):
    def inner(*a):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn(*a)
        e.record()
        records.append((tag(*a) if tag else name, s, e))
        return rc
    return inner


def tdesc(ref):
    t = ref._obj
    return f"{('nchw', 'nhwc', 'split', 'rowpad')[t.fmt]}[{t.B},{t.C},{t.H},{t.W}]"


orig_run = F.ConvPlan.run


def run(self, x_desc, y_desc, act=C.ACT_NONE, path=C.PATH_AUTO, xb_desc=None, sse=None):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    orig_run(self, x_desc, y_desc, act, path, xb_desc, sse)
    e.record()
    Cin, Cout, kh, kw, st, p, tr, op = self.geom
    gdn = ("+gdn" if self._gdn_key is not None else "") + ("+sse" if sse is not None else "")
    fl = 2.0 * y_desc.B * y_desc.H * y_desc.W * Cout * Cin * kh * kw / (st * st if tr else 1)
    records.append((f"{'deconv' if tr else 'conv'} {Cin}->{Cout} k{kh} s{st}{gdn} out {y_desc.H}x{y_desc.W} "
                    f"{('nchw', 'nhwc', 'split', 'rowpad')[y_desc.fmt]}", s, e, fl))


F.ConvPlan.run = run


class LibProxy:
    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, k):
        fn = getattr(self._lib, k)
        if k in ("hesic_convert", "hesic_warp_perspective", "hesic_entropy_bottleneck", "hesic_gaussian_conditional",
                 "hesic_spatial_max", "hesic_mixture_weights", "hesic_upsample_bilinear", "hesic_sum_squared_error"):
            if k == "hesic_convert":
                return wrap(k, fn, lambda s, d, op, st: f"convert {tdesc(s)}->{tdesc(d)} op{op}")
            return wrap(k, fn, None)
        return fn


import hesic_b200.engine as E  # noqa: E402

E._lib = LibProxy(C.lib)
tot = collections.OrderedDict()
wall = []
for _ in range(reps):
    records.clear()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    net(x1, x2, h)
    s1.record()
    torch.cuda.synchronize()
    wall.append(s0.elapsed_time(s1))
    for r in records:
        name, s, e = r[0], r[1], r[2]
        fl = r[3] if len(r) > 3 else 0.0
        ent = tot.setdefault(name, [0, 0.0, 0.0])
        ent[0] += 1
        ent[1] += s.elapsed_time(e)
        ent[2] += fl

total = sum(v[1] for v in tot.values()) / reps
print(f"model={model} B={B}: plain forward {plain_ms:.3f} ms ({B / plain_ms * 1e3:.0f} pairs/s); instrumented wall "
      f"{sum(wall) / reps:.3f} ms; sum of rows {total:.3f} ms")
print(f"{'ms/fwd':>8} {'%':>5} {'n':>3} {'TF/s':>7}  name")
for name, (n, ms, fl) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    ms /= reps
    tf = fl / reps / (ms * 1e-3) / 1e12 if fl else 0.0
    print(f"{ms:8.3f} {100 * ms / total:5.1f} {n // reps:3d} {tf:7.1f}  {name}")
