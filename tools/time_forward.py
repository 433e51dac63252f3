#!/usr/bin/env python
"""Whole-forward time with the engine's stream options switched on and off (A/B on one box, interleaved):
    python tools/time_forward.py [hesic|hesic_plus|dsic] [B] [rounds]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import hesic_b200
from hesic_b200 import synth
hesic_b200.install()
model = sys.argv[1] if len(sys.argv) > 1 else "hesic"
B = int(sys.argv[2]) if len(sys.argv) > 2 else (8 if model == "dsic" else 16)
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda", 0)
if model == "dsic":
    import mynet6_plus
    net = mynet6_plus.DSIC(128, 192, 21, 32, 5).eval()
else:
    net = __import__("newnet1_joint" if model == "hesic_plus" else "newnet1").HSIC(128, 192, 5).eval()
net.load_state_dict(synth.synth_state_dict(net, seed=0))
net = net.to(dev)
sets = [tuple(t.to(dev) for t in synth.stereo_pairs(B, 512, 512, seed=1234 + s)) for s in range(2)]
call = (lambda s: net(s[0], s[1])) if model == "dsic" else (lambda s: net(*s))
for s in sets:
    call(s)
eng = net.hesic_engine
n = 6 if model != "dsic" else 3
res = {}
configs = [(True, True), (True, False), (False, False)]
for r in range(rounds):
    for two, br in configs:
        eng.two_streams, eng.branch_streams = two, br
        time.sleep(1.0)     # every burst starts from an idle (cool, full-clock) GPU: a sustained loop is power-capped at
                            # ~1 kW and hides scheduling effects behind the clock governor
        for i in range(2):
            call(sets[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            call(sets[i % 2])
        e1.record()
        torch.cuda.synchronize()
        res.setdefault((two, br), []).append(e0.elapsed_time(e1) / n)
for k, v in res.items():
    print(f"{model} B={B} two_streams={k[0]!s:5} branch_streams={k[1]!s:5}: min {min(v):.3f} ms  median {sorted(v)[len(v) // 2]:.3f} ms  "
          f"({B / min(v) * 1e3:.0f} pairs/s)  all " + " ".join(f"{t:.3f}" for t in v))
