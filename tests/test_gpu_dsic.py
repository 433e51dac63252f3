"""DSIC (ywz/DSIC/mynet6_plus.py, BASELINE config 5 / SURVEY 8a row 15) on the B200: the three DSIC-only kernels
and the Conv3d-as-banded-conv2d against the oracle, then the whole forward against the oracle and the fixture the
unmodified reference produced."""
import math

import numpy as np
import pytest
import torch

from hesic_b200 import compat, synth
from oracle import hesic_oracle as O
from tests.helpers import T, assert_close, load_json, load_npz, mismatch_fraction

compat.install()
pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand(shape, seed, scale=1.0):
    return torch.from_numpy((np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32))


@pytest.mark.parametrize("shape,groups", [((2, 128, 16, 24), 4), ((1, 672, 4, 16), 21), ((3, 224, 8, 8), 1), ((2, 12, 5, 7), 3)])
def test_group_norm_relu(shape, groups):
    from hesic_b200 import functional as F
    x = _rand(shape, 1, 2.0) + 0.3
    w, b = 1 + _rand((shape[1],), 2, 0.1), _rand((shape[1],), 3, 0.1)
    ref = torch.nn.functional.group_norm(x, groups, w, b, 1e-5)
    assert_close(F.group_norm(x.to(DEV), groups, w.to(DEV), b.to(DEV)), ref, 1e-5, what="group_norm")
    assert_close(F.group_norm(x.to(DEV), groups, w.to(DEV), b.to(DEV), relu=True), torch.relu(ref), 1e-5, what="group_norm+relu")
    assert_close(F.group_norm(x.to(DEV), groups), torch.nn.functional.group_norm(x, groups), 1e-5, what="group_norm no affine")


def test_softmax_channels_and_dense_warp():
    from hesic_b200 import functional as F
    x = _rand((2, 32, 9, 40), 4, 3.0)
    ref = torch.softmax(x, dim=-3)
    got = F.softmax_channels(x.to(DEV))
    assert_close(got, ref, 1e-5, floor=1e-6, what="softmax over channels")
    assert float((got.sum(1) - 1).abs().max()) < 1e-5
    for (B, C, H, W, D) in ((2, 128, 8, 64, 32), (1, 5, 3, 40, 32), (1, 33, 2, 100, 7), (1, 4, 2, 16, 32)):
        h1 = _rand((B, C, H, W), 5)
        cost = torch.softmax(_rand((B, D, H, W), 6, 2.0), dim=1)
        if W >= D:
            ref = O.dsic_dense_warp(h1, cost)
        else:   # the reference's slicing breaks for W < D (mynet6_plus.py:342); the definition does not
            ref = torch.zeros_like(h1)
            for d in range(min(D, W)):
                ref[..., :W - d] += cost[:, d:d + 1, :, :W - d] * h1[..., d:]
        assert_close(F.dense_warp(h1.to(DEV), cost.to(DEV)), ref, 1e-5, what=f"dense_warp {B, C, H, W, D}")
    # one-hot cost = pure shift (size-independent property, full-size row)
    h1 = _rand((1, 128, 2, 256), 7)
    cost = torch.zeros(1, 32, 2, 256)
    cost[:, 5] = 1
    out = F.dense_warp(h1.to(DEV), cost.to(DEV)).cpu()
    assert torch.equal(out[..., :251], h1[..., 5:]) and float(out[..., 251:].abs().max()) == 0


def test_conv3d_as_banded_conv2d():
    from hesic_b200.dsic import Conv3dAs2d
    m = Conv3dAs2d(7, 7, kernel_size=5, stride=1, padding=2)
    w, b = _rand(tuple(m.weight.shape), 8, (2.0 / 875) ** 0.5), _rand((7,), 9, 0.1)
    m.load_state_dict({"weight": w, "bias": b})
    x = _rand((2, 7, 32, 16, 24), 10)
    ref = torch.nn.functional.conv3d(x, w, b, padding=2)
    assert_close(m.to(DEV)(x.to(DEV)), ref, 1e-4, what="Conv3d on the tensor-core path")


@pytest.fixture(scope="module")
def dsic():
    import mynet6_plus
    net = mynet6_plus.DSIC(128, 192, 21, 32, 5).eval()
    sd = synth.synth_state_dict(net, seed=0)
    net.load_state_dict(sd)
    return net.to(DEV), sd


def test_cost_volume_stages_vs_oracle(dsic):
    """Every stage of one cost volume (mynet6_plus.py:249-313) fed the ORACLE's input of that stage, each held at the
    path's 1e-4 bar: conv+GroupNorm+ReLU (x2), bilinear x8, Conv3d+GroupNorm+ReLU (x2, as banded 2-D convs), conv+GN+ReLU
    (x2), the logits conv, the disparity softmax, dense_warp."""
    from hesic_b200 import functional as F
    net, sd = dsic
    x1, x2, _ = synth.stereo_pairs(1, 64, 256, seed=1234)
    taps = {}
    with torch.no_grad():
        O.dsic_forward(sd, x1, x2, taps=taps)
    t = taps["cv1"]
    cv = net._cost_volume1
    d = lambda v: v.to(DEV)
    got = cv.model1[0:3](d(torch.cat((taps["g1_1"], taps["a1"]), 1)))
    assert_close(got, t["m1a"], 1e-4, what="model1 conv+GN+ReLU #1")
    assert_close(cv.model1[3:6](d(t["m1a"])), t["h_out"], 1e-4, what="model1 conv+GN+ReLU #2")
    ctx0 = taps["ctx0"]
    d_in = ctx0.reshape(-1, ctx0.size(-3), ctx0.size(-2), ctx0.size(-1))
    assert_close(F.upsample_bilinear(d(d_in).contiguous(), cv.scale_factor), t["d_up"], 1e-5, what="bilinear x8 (align_corners)")
    v0 = t["d_up"].reshape(-1, cv.F0, cv.C, t["d_up"].size(-2), t["d_up"].size(-1))
    assert_close(cv.model2[0:3](d(v0)), t["v1"], 1e-4, what="Conv3d+GN+ReLU #1")
    d_out = cv.model2[3:6](d(t["v1"]))
    assert_close(d_out.reshape(t["d_out"].shape), t["d_out"], 1e-4, what="Conv3d+GN+ReLU #2")
    assert_close(cv.model3[0:3](d(torch.cat((t["h_out"], t["d_out"]), 1))), t["m3a"], 1e-4, what="model3 conv+GN+ReLU #1")
    assert_close(cv.model3[3:6](d(t["m3a"])), t["m3b"], 1e-4, what="model3 conv+GN+ReLU #2")
    assert_close(cv.model3[6](d(t["m3b"])), t["logits"], 1e-4, what="disparity logits")
    assert_close(F.softmax_channels(d(t["logits"])), taps["cost1"], 1e-5, floor=1e-6, what="softmax over disparities")
    assert_close(net._warp1(d(taps["g1_1"]), d(taps["cost1"])), taps["warp1"], 1e-5, what="dense warp on the oracle's cost")


def test_cost_volume_chain_vs_oracle(dsic):
    """The whole cost volume in one go (nine layers back to back, nothing re-fed from the oracle).  Where the end-to-end
    error comes from: the stages are each inside 1e-4 (test above) and the chain's LOGITS are still within 3e-4 of the
    oracle's rms -- but softmax turns an ABSOLUTE logit error e into a RELATIVE probability error ~e, and the logits here
    have rms ~ several units, so the probabilities carry |logit|*1e-4-sized relative errors.  The softmax output is therefore
    held to exp(2*max|logit error|) - 1, with the logit error measured in the same run, not to a fixed 1e-4."""
    net, sd = dsic
    x1, x2, _ = synth.stereo_pairs(1, 64, 256, seed=1234)
    taps = {}
    with torch.no_grad():
        O.dsic_forward(sd, x1, x2, taps=taps)
    t = taps["cv1"]
    cv = net._cost_volume1
    h1, h2, ctx0 = taps["g1_1"].to(DEV), taps["a1"].to(DEV), taps["ctx0"].to(DEV)
    from hesic_b200 import functional as F
    h_out = cv.model1(torch.cat((h1, h2), 1))
    d_in = ctx0.reshape(-1, ctx0.size(-3), ctx0.size(-2), ctx0.size(-1))
    d_up = F.upsample_bilinear(d_in.contiguous(), cv.scale_factor)
    d_out = cv.model2(d_up.reshape(-1, cv.F0, cv.C, d_up.size(-2), d_up.size(-1)))
    logits = cv.model3(torch.cat((h_out, d_out.reshape(-1, cv.F0 * cv.C, d_out.size(-2), d_out.size(-1))), 1))
    assert_close(logits, t["logits"], 3e-4, what="logits after the nine-layer chain")
    e = float((logits.cpu().double() - t["logits"].double()).abs().max())
    got = net._cost_volume1(h1, h2, ctx0)
    ref = taps["cost1"]
    rel = float(((got.cpu().double() - ref.double()).abs() / ref.double().clamp(min=1e-6)).max())
    assert rel <= math.expm1(2 * e) + 1e-5, (rel, e)
    assert rel < 2e-3


def test_dsic_forward_vs_oracle_and_reference_fixture(dsic):
    from hesic_b200 import _capi as C
    net, sd = dsic
    g, meta = load_npz("dsic"), load_json("dsic")
    x1, x2, _ = synth.stereo_pairs(1, meta["H"], meta["W"], seed=1234)
    with torch.no_grad():
        ref = O.dsic_forward(sd, x1, x2)
    out = net(x1.to(DEV), x2.to(DEV))
    C.check(C.lib.hesic_tc_status())
    cpu = {"x1_hat": out["x1_hat"].cpu(), "x2_hat": out["x2_hat"].cpu(),
           "likelihoods": {k: v.cpu() for k, v in out["likelihoods"].items()}}
    # symbols may flip at exact x.5 ties (fp32 summation order), so images are compared in the L2 sense
    for k in ("x1_hat", "x2_hat"):
        a, b = cpu[k].double(), T(g[k]).double()
        rel = float((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt())
        assert rel < 5e-3, (k, rel)
    assert mismatch_fraction(cpu["likelihoods"]["y1"] > 0.5, T(g["lik_y1"]) > 0.5) < 2e-3
    assert_close(cpu["likelihoods"]["z1"], g["lik_z1"], 1e-3, floor=1e-9)
    m, r = synth.rd_metrics(cpu, x1, x2), synth.rd_metrics(ref, x1, x2)
    for k in ("bpp", "bpp1", "bpp2"):
        assert abs(m[k] - r[k]) <= 5e-3 * r[k], (k, m[k], r[k])
        assert abs(m[k] - meta["metrics"][k]) <= 5e-3 * meta["metrics"][k], (k, m[k], meta["metrics"][k])
    for k in ("psnr1", "psnr2"):
        assert abs(m[k] - r[k]) <= 0.05, (k, m[k], r[k])


def _to_split(x, Cs=None, c0=0):
    """NCHW fp32 -> channel slice [c0, c0+C) of a [2,B,H,W,Cs] SPLIT buffer (rest NaN-free zeros)."""
    from hesic_b200 import _capi as C
    B, Cn, H, W = x.shape
    Cs = Cn if Cs is None else Cs
    t = torch.zeros((2, B, H, W, Cs), device=DEV, dtype=torch.bfloat16)
    xd = x.to(DEV).contiguous()
    C.check(C.lib.hesic_convert(C.ref(C.nchw(xd)), C.ref(C.split(t, Cn, c0)), C.OP_COPY, C.stream()))
    torch.cuda.synchronize()
    return t


def _from_split(t, Cn=None, c0=0):
    v = t[0].float() + t[1].float()
    Cn = v.shape[-1] - c0 if Cn is None else Cn
    return v[..., c0:c0 + Cn].permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("shape,groups", [((2, 128, 16, 24), 4), ((1, 672, 4, 16), 21), ((3, 224, 8, 8), 1), ((2, 32, 5, 7), 2)])
def test_group_norm_channels_last(shape, groups):
    """The fused engine's GroupNorm: NHWC fp32 in (the conv output), SPLIT channel slice out (the next conv's input)."""
    from hesic_b200 import _capi as C
    B, Cn, H, W = shape
    x = _rand(shape, 1, 2.0) + 0.3
    w, b = 1 + _rand((Cn,), 2, 0.1), _rand((Cn,), 3, 0.1)
    ref = torch.relu(torch.nn.functional.group_norm(x, groups, w, b, 1e-5))
    xn = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    wd, bd = w.to(DEV), b.to(DEV)
    out = torch.zeros((2, B, H, W, Cn + 40), device=DEV, dtype=torch.bfloat16)
    C.check(C.lib.hesic_group_norm(C.ref(C.nhwc(xn)), C.ref(C.split(out, Cn, 8)), groups, C.ptr(wd), C.ptr(bd), 1e-5, 1, C.stream()))
    assert_close(_from_split(out, Cn, 8), ref, 2e-5, what="group_norm NHWC -> SPLIT slice")
    assert float(out[..., :8].float().abs().max()) == 0 and float(out[..., Cn + 8:].float().abs().max()) == 0
    on = torch.empty((B, H, W, Cn), device=DEV)
    C.check(C.lib.hesic_group_norm(C.ref(C.nhwc(xn)), C.ref(C.nhwc(on)), groups, None, None, 1e-5, 0, C.stream()))
    assert_close(on.permute(0, 3, 1, 2), torch.nn.functional.group_norm(x, groups), 1e-5, what="group_norm NHWC no affine")


def test_softmax_and_dense_warp_channels_last():
    from hesic_b200 import _capi as C
    x = _rand((2, 32, 9, 40), 4, 3.0)
    xn = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    yn = torch.empty_like(xn)
    C.check(C.lib.hesic_softmax_channels(C.ref(C.nhwc(xn)), C.ref(C.nhwc(yn)), C.stream()))
    assert_close(yn.permute(0, 3, 1, 2), torch.softmax(x, dim=-3), 1e-5, floor=1e-6, what="softmax NHWC")
    for (B, Cn, H, W, D) in ((2, 128, 8, 64, 32), (1, 128, 3, 100, 32), (1, 8, 2, 40, 7)):
        h1 = _rand((B, Cn, H, W), 5)
        cost = torch.softmax(_rand((B, D, H, W), 6, 2.0), dim=1)
        buf = _to_split(h1, 3 * Cn, 2 * Cn)                      # [w | a | g] level buffer of the engine
        hq = _from_split(buf, Cn, 2 * Cn).cpu()
        ref = O.dsic_dense_warp(hq, cost)
        cn = cost.permute(0, 2, 3, 1).contiguous().to(DEV)
        C.check(C.lib.hesic_dense_warp(C.ref(C.split(buf, Cn, 2 * Cn)), C.ref(C.nhwc(cn)), C.ref(C.split(buf, Cn, 0)), C.stream()))
        assert_close(_from_split(buf, Cn, 0), ref, 2e-5, what=f"dense_warp channels-last {B, Cn, H, W, D}")
        assert torch.equal(_from_split(buf, Cn, 2 * Cn).cpu(), hq) and float(buf[..., Cn:2 * Cn].float().abs().max()) == 0


def test_dsic_engine_and_operator_level_second_size(dsic):
    """A second size and batch 2: the fused engine and the operator-level composition of the same kernels both match
    the oracle (they differ from each other only through x.5 rounding ties of the latents)."""
    net, sd = dsic
    x1, x2, _ = synth.stereo_pairs(2, 128, 320, seed=21)   # W / 8 >= 32 disparities (the reference breaks below)
    with torch.no_grad():
        ref = O.dsic_forward(sd, x1, x2)
    a = net(x1.to(DEV), x2.to(DEV))
    b = net.forward_operator_level(x1.to(DEV), x2.to(DEV))
    from hesic_b200 import _capi as C
    C.check(C.lib.hesic_tc_status())
    r = synth.rd_metrics(ref, x1, x2)
    for name, out in (("engine", a), ("operator level", b)):
        cpu = {"x1_hat": out["x1_hat"].cpu(), "x2_hat": out["x2_hat"].cpu(),
               "likelihoods": {k: v.cpu() for k, v in out["likelihoods"].items()}}
        for k in ("x1_hat", "x2_hat"):
            rel = float((cpu[k].double() - ref[k].double()).pow(2).sum().sqrt() / ref[k].double().pow(2).sum().sqrt())
            assert rel < 2e-2, (name, k, rel)
        assert_close(cpu["likelihoods"]["z1"], ref["likelihoods"]["z1"], 1e-3, floor=1e-9, what=name + " z1 likelihood")
        m = synth.rd_metrics(cpu, x1, x2)
        for k in ("bpp", "bpp1", "bpp2"):
            assert abs(m[k] - r[k]) <= 5e-3 * r[k], (name, k, m[k], r[k])
        for k in ("psnr1", "psnr2"):
            assert abs(m[k] - r[k]) <= 0.1, (name, k, m[k], r[k])


def test_dsic_plus_enhancement_vs_reference_fixture():
    """mynet6_plus.Independent_EN (second stage of DSIC_plus) on the fused enhancement kernels."""
    import mynet6_plus
    en = mynet6_plus.Independent_EN().eval()
    en.load_state_dict(synth.synth_state_dict(en, seed=0))
    x1, x2, _ = synth.stereo_pairs(1, 64, 64, seed=98)
    out = en.to(DEV)(x1.to(DEV), x2.to(DEV))
    g = load_npz("dsic_independent_en")
    assert_close(out["x1_hat"], g["x1_hat"], 1e-4, what="DSIC EN x1")
    assert_close(out["x2_hat"], g["x2_hat"], 1e-4, what="DSIC EN x2")


@pytest.mark.parametrize("scale", [2, 4, 8])
def test_upsample_channels_last(scale):
    """nn.UpsamplingBilinear2d (align_corners=True) in the engines' layouts: an NHWC fp32 channel slice (a context volume
    of the global-context tensor, mynet6_plus.py:269) or SPLIT planes (z_hat, newnet1.py:564) into a SPLIT slice."""
    from hesic_b200 import _capi as C
    x = _rand((2, 224, 4, 6), 61)
    ref = torch.nn.functional.interpolate(x, scale_factor=scale, mode="bilinear", align_corners=True)
    wide = torch.zeros((2, 4, 6, 672), device=DEV)
    wide[..., 224:448] = x.permute(0, 2, 3, 1).to(DEV)
    out = torch.zeros((2, 2, 4 * scale, 6 * scale, 224 + 16), device=DEV, dtype=torch.bfloat16)
    C.check(C.lib.hesic_upsample_bilinear(C.ref(C.nhwc(wide, 224, 224)), C.ref(C.split(out, 224, 8)), scale, C.stream()))
    assert_close(_from_split(out, 224, 8), ref, 2e-5, what="upsample NHWC slice -> SPLIT slice")
    assert float(out[..., :8].float().abs().max()) == 0 and float(out[..., 232:].float().abs().max()) == 0
    xs = _to_split(x)
    xq = _from_split(xs).cpu()
    out2 = torch.zeros((2, 2, 4 * scale, 6 * scale, 224), device=DEV, dtype=torch.bfloat16)
    C.check(C.lib.hesic_upsample_bilinear(C.ref(C.split(xs)), C.ref(C.split(out2)), scale, C.stream()))
    assert_close(_from_split(out2), torch.nn.functional.interpolate(xq, scale_factor=scale, mode="bilinear", align_corners=True), 2e-5,
                 what="upsample SPLIT -> SPLIT")


def test_conv3d_depth_major_banded_plan():
    """The engine's form of the cost-volume Conv3d: channels stacked (depth, feature), the all-zero (N tile, K chunk)
    blocks of the banded 2-D weight skipped by the tensor-core kernel -- same result as nn.Conv3d."""
    from hesic_b200 import _capi as C
    from hesic_b200.dsic import Conv3dAs2d
    m = Conv3dAs2d(7, 7, kernel_size=5, stride=1, padding=2)
    w, b = _rand(tuple(m.weight.shape), 8, (2.0 / 875) ** 0.5), _rand((7,), 9, 0.1)
    m.load_state_dict({"weight": w, "bias": b})
    x = _rand((2, 7, 32, 16, 24), 10)
    ref = torch.nn.functional.conv3d(x, w, b, padding=2)                      # [B, F, D, H, W]
    m = m.to(DEV)
    plan = m._plan_for(32, True)
    xs = _to_split(x.permute(0, 2, 1, 3, 4).reshape(2, 224, 16, 24))          # (d, f)-ordered channels
    y = torch.empty((2, 16, 24, 224), device=DEV)
    plan.run(C.split(xs), C.nhwc(y), C.ACT_NONE, C.PATH_TC)
    C.check(C.lib.hesic_tc_status())
    got = y.permute(0, 3, 1, 2).reshape(2, 32, 7, 16, 24).permute(0, 2, 1, 3, 4)
    xq = _from_split(xs).cpu().reshape(2, 32, 7, 16, 24).permute(0, 2, 1, 3, 4)
    assert_close(got, torch.nn.functional.conv3d(xq, w, b, padding=2), 1e-4, what="depth-major banded Conv3d")
    assert_close(got, ref, 1e-4, what="depth-major banded Conv3d vs fp32 input")


@pytest.mark.parametrize("Cin,Cout,H,W,groups", [(64, 128, 16, 24, 4), (64, 128, 8, 8, 4), (32, 224, 16, 16, 1), (64, 672, 16, 16, 21),
                                                  (64, 128, 40, 56, 4)])
def test_conv_with_fused_group_norm_statistics(Cin, Cout, H, W, groups):
    """conv -> GroupNorm -> ReLU with the statistics accumulated by the conv epilogue (hesic_conv_forward_gn +
    hesic_group_norm_apply) against torch; includes tiles that straddle images (statistics-kernel route), a one-group norm
    over 224 channels, 21 groups of 32, and an image whose edge tiles are partly outside."""
    from hesic_b200 import _capi as C
    from hesic_b200 import functional as F
    B = 3
    x = _rand((B, Cin, H, W), 71)
    w = _rand((Cout, Cin, 5, 5), 72, (2.0 / (Cin * 25)) ** 0.5)
    b = _rand((Cout,), 73, 0.1)
    gw, gb = 1 + _rand((Cout,), 74, 0.1), _rand((Cout,), 75, 0.1)
    plan = F.ConvPlan(Cin, Cout, 5, 1, 2)
    wd, bd = w.to(DEV), b.to(DEV)
    plan.load(wd, bd)
    xs = _to_split(x)
    xq = _from_split(xs).cpu()
    conv_ref = torch.nn.functional.conv2d(xq, w, b, padding=2)
    ref = torch.relu(torch.nn.functional.group_norm(conv_ref, groups, gw, gb, 1e-5))
    y = torch.empty((B, H, W, Cout), device=DEV)
    stats = torch.full((B, groups, C.GN_SLOTS, 2), float("nan"), device=DEV, dtype=torch.float64)   # the call zeroes it
    C.check(C.lib.hesic_conv_forward_gn(plan.h, C.ref(C.split(xs)), C.ref(C.nhwc(y)), C.PATH_TC, C.ptr(stats), groups, C.stream()))
    C.check(C.lib.hesic_tc_status())
    assert_close(y.permute(0, 3, 1, 2), conv_ref, 1e-4, what="conv output")
    # the statistics are those of the tensor the kernel wrote (fp32 partial sums of 32 values, then fp64)
    s = stats.sum(2).cpu()
    cg = y.permute(0, 3, 1, 2).double().reshape(B, groups, -1).cpu()
    scale = cg.abs().sum(-1)
    assert ((s[..., 0] - cg.sum(-1)).abs() <= 2e-6 * scale).all(), "group sums"
    assert ((s[..., 1] - cg.pow(2).sum(-1)).abs() <= 2e-6 * cg.pow(2).sum(-1)).all(), "group sums of squares"
    out = torch.zeros((2, B, H, W, Cout + 16), device=DEV, dtype=torch.bfloat16)
    gwd, gbd = gw.to(DEV), gb.to(DEV)
    C.check(C.lib.hesic_group_norm_apply(C.ref(C.nhwc(y)), C.ref(C.split(out, Cout, 8)), groups, C.ptr(gwd), C.ptr(gbd), 1e-5, 1,
                                         C.ptr(stats), C.stream()))
    assert_close(_from_split(out, Cout, 8), ref, 1e-4, what="conv -> GroupNorm -> ReLU")
