#!/usr/bin/env python
"""Device-side check of the tcgen05 conv path against the SIMT fp32 path on identical split inputs.

    python tools/tc_check.py [group ...]      groups: s1 s2 deconv gdn row head big edge

Prints one line per case: max |tc - simt| relative to rms(simt).  Exit code 1 on any mismatch.
Run each group in its own process (a trapped kernel poisons the CUDA context)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import hesic_b200
from hesic_b200 import _capi as C
from hesic_b200 import functional as F

hesic_b200.install()
from compressai.models.utils import conv, deconv  # noqa: E402

DEV = "cuda:0"
TOL = 1e-4

GROUPS = {
    # Cin, Cout, k, stride, transposed, H, W, B, gdn(0/1/2), act
    "s1": [(128, 128, 5, 1, False, 16, 16, 1, 0, 0), (192, 128, 5, 1, False, 8, 8, 3, 0, 1), (128, 960, 5, 1, False, 8, 8, 2, 0, 2),
           (320, 128, 5, 1, False, 32, 32, 2, 0, 0), (192, 128, 3, 1, False, 8, 8, 1, 0, 0), (768, 640, 1, 1, False, 8, 8, 1, 0, 2),
           (64, 64, 3, 1, False, 24, 40, 1, 0, 0), (128, 288, 5, 1, False, 4, 4, 5, 0, 0)],
    "s2": [(128, 128, 5, 2, False, 32, 32, 1, 0, 0), (128, 128, 5, 2, False, 32, 48, 2, 0, 0), (128, 192, 5, 2, False, 16, 16, 3, 0, 0),
           (128, 128, 5, 2, False, 4, 4, 3, 0, 1), (128, 128, 5, 2, False, 64, 64, 2, 0, 0)],
    "deconv": [(192, 128, 5, 2, True, 8, 8, 2, 0, 0), (128, 128, 5, 2, True, 16, 24, 1, 0, 1), (128, 960, 5, 2, True, 4, 4, 2, 0, 0),
               (128, 192, 5, 2, True, 2, 2, 3, 0, 2), (128, 128, 5, 2, True, 32, 32, 2, 0, 0)],
    "gdn": [(128, 128, 5, 2, False, 32, 32, 2, 1, 0), (128, 128, 5, 2, True, 16, 16, 2, 2, 0), (128, 128, 5, 2, False, 64, 96, 3, 1, 0),
            (192, 128, 5, 2, True, 8, 8, 4, 2, 0), (128, 128, 5, 2, False, 128, 128, 4, 1, 0)],
    "row": [(3, 128, 5, 2, False, 64, 64, 2, 1, 0), (3, 128, 5, 2, False, 32, 48, 1, 0, 2), (6, 3, 5, 1, False, 40, 24, 2, 1, 0),
            (6, 3, 5, 1, True, 24, 40, 1, 0, 0), (3, 3, 5, 1, False, 16, 16, 1, 2, 0), (3, 128, 5, 2, False, 512, 512, 2, 1, 0)],
    "head": [(128, 3, 5, 2, True, 16, 16, 2, 0, 0), (128, 3, 5, 2, True, 32, 24, 1, 2, 0), (192, 3, 5, 2, True, 8, 8, 3, 1, 0),
               (128, 3, 5, 2, True, 256, 256, 2, 2, 0)],
    "edge": [(3, 128, 5, 2, False, 512, 512, 16, 1, 0), (128, 3, 5, 2, True, 256, 256, 16, 2, 0), (6, 3, 5, 1, False, 512, 512, 16, 1, 0),
             (6, 3, 5, 1, True, 512, 512, 16, 0, 0)],
    "time": [(3, 128, 5, 2, False, 512, 512, 16, 1, 0), (3, 128, 5, 2, False, 512, 512, 16, 0, 0),
             (128, 128, 5, 2, False, 256, 256, 16, 1, 0), (128, 128, 5, 2, False, 256, 256, 16, 0, 0),
             (128, 128, 5, 2, True, 128, 128, 16, 2, 0), (128, 128, 5, 2, True, 128, 128, 16, 0, 0),
             (128, 128, 5, 2, True, 64, 64, 16, 2, 0), (192, 128, 5, 2, True, 32, 32, 16, 2, 0),
             (128, 128, 5, 2, False, 128, 128, 16, 1, 0), (128, 192, 5, 2, False, 64, 64, 16, 0, 0),
             (128, 960, 5, 1, False, 32, 32, 16, 0, 1), (320, 128, 5, 1, False, 32, 32, 16, 0, 1),
             (128, 3, 5, 2, True, 256, 256, 16, 2, 0), (6, 3, 5, 1, False, 512, 512, 16, 1, 0)],
    "big": [(128, 128, 5, 2, False, 256, 256, 4, 1, 0), (128, 128, 5, 2, True, 128, 128, 4, 2, 0), (320, 128, 5, 1, False, 32, 32, 16, 0, 1),
            (128, 960, 5, 1, False, 32, 32, 16, 0, 1)],
}


def to_split(x, stride=1, transposed=False):
    """-> descriptor of the bf16 (hi, lo) planes the tensor-core path reads (ROWPAD, 4 or 8 slots, for <= 8 channels)."""
    B, Cn, H, W = x.shape
    if Cn <= 8:
        slots = C.rowpad_slots(Cn, stride, transposed)
        xs = torch.zeros((2, B, H + C.ROWPAD_Y, W + C.ROWPAD_X, slots), device=x.device, dtype=torch.bfloat16)
        d = C.rowpad(xs, Cn)
    else:
        xs = torch.empty((2, B, H, W, Cn), device=x.device, dtype=torch.bfloat16)
        d = C.split(xs)
    C.check(C.lib.hesic_convert(C.ref(C.nchw(x)), C.ref(d), C.OP_COPY, C.stream()))
    d._keep = xs
    return d


def time_case(case):
    """tcgen05 path only: milliseconds per launch (split / planar output), L2 flushed between launches."""
    Cin, Cout, k, s, tr, H, W, B, gdn, act = case
    g = torch.Generator().manual_seed(1)
    mod = (deconv if tr else conv)(Cin, Cout, kernel_size=k, stride=s).to(DEV)
    plan = mod.hesic_plan()
    if gdn:
        plan.set_gdn(torch.ones(Cout, device=DEV), 0.1 * torch.eye(Cout, device=DEV) + 0.01, gdn == 2)
    xd = to_split(torch.randn(B, Cin, H, W, generator=g).to(DEV), s, tr)
    Ho, Wo = plan.out_hw(H, W)
    if Cout <= 4:
        yt = torch.zeros((B, Cout, Ho, Wo), device=DEV); yd = C.nchw(yt)
    else:
        yt = torch.zeros((2, B, Ho, Wo, Cout), device=DEV, dtype=torch.bfloat16); yd = C.split(yt)
    flush = torch.empty(256 << 20, device=DEV, dtype=torch.uint8)
    ts = []
    for i in range(6):
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
    fl = 2.0 * B * Ho * Wo * Cout * Cin * k * k / (s * s if tr else 1) + (2.0 * B * Ho * Wo * Cout * Cout if gdn else 0)
    out_b = yt.numel() * yt.element_size()
    print(f"TIME {case} {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s(alg)  out {out_b / ms / 1e6:.0f} GB/s", flush=True)
    return True


def run_case(case, timing=False):
    Cin, Cout, k, s, tr, H, W, B, gdn, act = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    mod = (deconv if tr else conv)(Cin, Cout, kernel_size=k, stride=s)
    w = torch.randn(mod.weight.shape, generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    mod.load_state_dict({"weight": w, "bias": b})
    mod = mod.to(DEV)
    plan = mod.hesic_plan()
    if gdn:
        beta = (torch.rand(Cout, generator=g) + 0.5).to(DEV)
        gamma = (torch.rand(Cout, Cout, generator=g) * 0.02 + 0.1 * torch.eye(Cout)).to(DEV)
        plan.set_gdn(beta, gamma, gdn == 2)
    x = torch.randn(B, Cin, H, W, generator=g).to(DEV)
    xd = to_split(x, s, tr)
    Ho, Wo = plan.out_hw(H, W)
    outs = {}
    planar = Cout <= 4
    for name, path in (("simt", C.PATH_SIMT), ("tc", C.PATH_TC)):
        for kind in ("nhwc", "split"):
            if planar:
                y = torch.full((B, Cout + 1, Ho, Wo), float("nan"), device=DEV)   # slice of a wider buffer
                plan.run(xd, C.nchw(y, Cout, 0 if kind == "nhwc" else 1), act, path)
                y = y[:, :Cout] if kind == "nhwc" else y[:, 1:]
            elif kind == "nhwc":
                y = torch.full((B, Ho, Wo, Cout), float("nan"), device=DEV)
                plan.run(xd, C.nhwc(y), act, path)
            else:
                ys = torch.zeros((2, B, Ho, Wo, Cout), device=DEV, dtype=torch.bfloat16)
                plan.run(xd, C.split(ys), act, path)
                y = ys[0].float() + ys[1].float()
            torch.cuda.synchronize()
            C.check(C.lib.hesic_tc_status())
            outs[(name, kind)] = y
    ref = outs[("simt", "nhwc")]
    rms = float(ref.pow(2).mean().sqrt())
    e1 = float((outs[("tc", "nhwc")] - ref).abs().max()) / rms
    e2 = float((outs[("tc", "split")] - outs[("simt", "split")]).abs().max()) / rms
    nan = bool(torch.isnan(outs[("tc", "nhwc")]).any())
    ok = (e1 < TOL) and (e2 < 4 * TOL) and not nan
    line = f"{'OK ' if ok else 'BAD'} {case} err_f32={e1:.2e} err_split={e2:.2e} nan={nan} rms={rms:.3f}"
    if timing:
        if planar:
            yt = torch.zeros((B, Cout, Ho, Wo), device=DEV)
            yd = C.nchw(yt)
        else:
            yt = torch.zeros((2, B, Ho, Wo, Cout), device=DEV, dtype=torch.bfloat16)
            yd = C.split(yt)
        for path, nm in ((C.PATH_TC, "tc"), (C.PATH_SIMT, "simt")):
            reps = 5 if nm == "tc" else 1
            e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            plan.run(xd, yd, act, path)
            e0.record()
            for _ in range(reps):
                plan.run(xd, yd, act, path)
            e1_.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1_) / reps
            taps = k * k
            fl = 2.0 * B * Ho * Wo * Cout * Cin * taps / (s * s if tr else 1) + (2.0 * B * Ho * Wo * Cout * Cout if gdn else 0)
            line += f" | {nm} {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s"
    print(line, flush=True)
    return ok


def main():
    groups = sys.argv[1:] or ["s1", "s2", "deconv", "gdn"]
    name = C.ctypes.create_string_buffer(128)
    C.check(C.lib.hesic_device_check(name, 128))
    print("device:", name.value.decode(), flush=True)
    ok = True
    for gname in groups:
        timing = gname in ("big", "edge")
        for case in GROUPS[gname]:
            t0 = time.time()
            try:
                ok &= time_case(case) if gname == "time" else run_case(case, timing)
            except Exception as e:  # noqa: BLE001
                print(f"EXC {case}: {type(e).__name__}: {e}", flush=True)
                return 1
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
