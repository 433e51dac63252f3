#!/usr/bin/env python
"""Per-layer device times of one DSIC forward (CUDA events around every conv launch), grouped by layer shape.

    python tools/dsic_layer_times.py [B] [reps]"""
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
import mynet6_plus  # noqa: E402
import hesic_b200.dsic_engine as DE  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = "cuda:0"
net = mynet6_plus.DSIC(128, 192, 21, 32, 5).eval()
net.load_state_dict(synth.synth_state_dict(net, seed=0))
net = net.to(dev)
x1, x2, _ = (t.to(dev) for t in synth.stereo_pairs(B, 512, 512, seed=1234))
for _ in range(2):
    net(x1, x2)
torch.cuda.synchronize()
records = []
geom_of = {}
orig_run = F.ConvPlan.run
orig_init = F.ConvPlan.__init__


def run(self, x_desc, y_desc, act=C.ACT_NONE, path=C.PATH_AUTO, xb_desc=None, sse=None):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    orig_run(self, x_desc, y_desc, act, path, xb_desc, sse)
    e.record()
    note(self, y_desc, s, e, "")


def note(plan, y_desc, s, e, tag):
    Cin, Cout, kh, kw, st, p, tr, op = plan.geom
    fl = 2.0 * y_desc.B * y_desc.H * y_desc.W * Cout * Cin * kh * kw / (st * st if tr else 1)
    gdn = "+gdn" if plan._gdn_on else ""
    records.append((f"{'deconv' if tr else 'conv'} {Cin}->{Cout} k{kh} s{st}{gdn}{tag} out {y_desc.H}x{y_desc.W}", s, e, fl))


F.ConvPlan.run = run
plans = {}
for m in net.modules():
    p = m.__dict__.get("_hesic_plan")
    if p is not None:
        plans[p.h] = p


class LibProxy:
    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, k):
        fn = getattr(self._lib, k)
        if k == "hesic_conv_forward_gn":
            def inner(h, x, y, path, stats, groups, stream):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                rc = fn(h, x, y, path, stats, groups, stream)
                e.record()
                hv = h if isinstance(h, int) else h.value
                plan = PLANS.get(hv)
                if plan is not None:
                    note(plan, y._obj, s, e, "+gnstats")
                else:
                    records.append((f"conv+gnstats (plan {hv:#x}) out {y._obj.H}x{y._obj.W} C={y._obj.C}", s, e, 0.0))
                return rc
            return inner
        if k in ("hesic_group_norm_apply", "hesic_upsample_bilinear", "hesic_dense_warp", "hesic_softmax_channels"):
            def inner(*a):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                rc = fn(*a)
                e.record()
                records.append((k, s, e, 0.0))
                return rc
            return inner
        return fn


PLANS = {}
orig_plan_init = F.ConvPlan.__init__


def find_plans():
    import gc
    for o in gc.get_objects():
        if isinstance(o, F.ConvPlan) and getattr(o, "h", None):
            PLANS[o.h] = o


find_plans()
DE._lib = LibProxy(C.lib)
tot = collections.OrderedDict()
for _ in range(reps):
    records.clear()
    net(x1, x2)
    torch.cuda.synchronize()
    for name, s, e, fl in records:
        ent = tot.setdefault(name, [0, 0.0, 0.0])
        ent[0] += 1
        ent[1] += s.elapsed_time(e)
        ent[2] += fl
total = sum(v[1] for v in tot.values()) / reps
print(f"DSIC B={B}: sum of rows {total:.2f} ms")
print(f"{'ms/fwd':>8} {'%':>5} {'n':>3} {'TF/s':>7}  name")
for name, (n, ms, fl) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    ms /= reps
    tf = fl / reps / (ms * 1e-3) / 1e12 if fl else 0.0
    print(f"{ms:8.3f} {100 * ms / total:5.1f} {n // reps:3d} {tf:7.1f}  {name}")
