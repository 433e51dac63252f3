"""Per-operator parity on the B200: every kernel is fed the oracle's INPUT tensors and compared with
the oracle's output.  Bar: |a-b| <= 1e-4 * max(|b|, floor) for fp32 results, with
floor = rms(b) for conv-like outputs and 1e-9 for likelihoods; integer results bit-exact."""
import math
import numpy as np
import pytest
import torch

import oracle
from hesic_b200 import compat, synth
from oracle import hesic_oracle as O
from tests.helpers import T, assert_close, load_npz

compat.install()
pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    return load_npz("operators")


def _rand(shape, seed, scale=1.0):
    return torch.from_numpy((np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32))


CONV_CASES = [
    # Cin, Cout, k, stride, transposed, H, W, B
    (3, 128, 5, 2, False, 64, 64, 2),      # encoder first layer
    (3, 128, 5, 2, False, 40, 24, 3),      # ... ragged tiles
    (3, 64, 5, 2, False, 16, 32, 1),       # ... narrower N tile
    (128, 128, 5, 2, False, 32, 48, 2),    # main analysis layer (ragged W tile)
    (128, 192, 5, 2, False, 16, 16, 1),
    (192, 128, 5, 1, False, 8, 8, 3),      # encode_hyper first conv
    (320, 128, 5, 1, False, 8, 8, 1),      # gmm_hyper_y2 first conv
    (128, 960, 5, 1, False, 8, 8, 1),      # mixture heads
    (6, 3, 5, 1, False, 40, 24, 1),        # pre_conv
    (192, 128, 5, 2, True, 8, 8, 2),       # synthesis first layer
    (128, 128, 5, 2, True, 16, 24, 1),
    (128, 3, 5, 2, True, 16, 16, 2),       # synthesis RGB head
    (6, 3, 5, 1, True, 24, 40, 1),         # after_conv (stride-1 transposed)
    (192, 128, 3, 1, False, 8, 8, 1),      # HESIC+ h_a first conv
    (768, 640, 1, 1, False, 8, 8, 1),      # HESIC+ entropy_parameters
    (32, 32, 3, 1, False, 24, 24, 1),      # Independent_EN residual conv
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "-".join(map(str, c)))
@pytest.mark.parametrize("path", ["simt", "auto", "tc"])
def test_conv_parity(case, path):
    from hesic_b200 import _capi as C
    from hesic_b200 import functional as F
    from compressai.models.utils import conv, deconv
    Cin, Cout, k, s, tr, H, W, B = case
    mod = (deconv if tr else conv)(Cin, Cout, kernel_size=k, stride=s)
    w = _rand(tuple(mod.weight.shape), 1, (2.0 / (Cin * k * k)) ** 0.5)
    b = _rand((Cout,), 2, 0.1)
    x = _rand((B, Cin, H, W), 3)
    ref = O.deconv(x, w, b, stride=s) if tr else O.conv(x, w, b, stride=s)
    mod.load_state_dict({"weight": w, "bias": b})
    mod = mod.to(DEV)
    pth = {"simt": C.PATH_SIMT, "auto": C.PATH_AUTO, "tc": C.PATH_TC}[path]
    for act, fn in ((C.ACT_NONE, lambda t: t), (C.ACT_LEAKY, torch.nn.functional.leaky_relu)):
        try:
            y = F.conv2d(x.to(DEV), mod.hesic_plan(), act=act, path=pth)
        except NotImplementedError:
            assert path == "tc"
            pytest.skip("shape not on the tcgen05 path")
        assert y.shape == ref.shape
        assert_close(y, fn(ref), 1e-4, what=f"conv {case} act={act} path={path}")


@pytest.mark.parametrize("size", [(2, 40, 24), (1, 64, 128), (2, 70, 67)])
@pytest.mark.parametrize("transposed,slots", [(False, 4), (False, 8), (True, 8)])
def test_few_channel_stencil_cat_gdn_rowpad(size, transposed, slots):
    """pre_gdn(pre_conv(cat(a, b))) / after_conv(cat(a, b)) (newnet1.py:643-644,686): the exact-fp32 stencil fed two
    sources, with the fused 3-channel GDN, writing NCHW and the ROWPAD8 (hi, lo) planes of the next layer."""
    from hesic_b200 import _capi as C
    from compressai.layers import GDN
    from compressai.models.utils import conv, deconv
    B, H, W = size
    mod = (deconv if transposed else conv)(6, 3, kernel_size=5, stride=1)
    w = _rand(tuple(mod.weight.shape), 11, 0.1)
    b = _rand((3,), 12, 0.1)
    mod.load_state_dict({"weight": w, "bias": b})
    g = GDN(3, inverse=transposed)
    g.load_state_dict({"beta": torch.rand(3) + 0.5, "gamma": torch.rand(3, 3) * 0.2}, strict=False)
    xa, xb = _rand((B, 3, H, W), 13), _rand((B, 3, H, W), 14)
    xin = torch.cat((xa, xb), 1)
    ref_c = O.deconv(xin, w, b, stride=1) if transposed else O.conv(xin, w, b, stride=1)
    ref_g = O.gdn(ref_c, g.beta.detach(), g.gamma.detach(), inverse=transposed)
    mod, g = mod.to(DEV), g.to(DEV)
    plan = mod.hesic_plan()
    xa_d, xb_d = xa.to(DEV), xb.to(DEV)
    # (1) plain conv, cat input, NCHW output (channel slice of a wider buffer)
    plan.set_gdn(None, None, False)
    out = torch.zeros((B, 5, H, W), device=DEV)
    plan.run(C.nchw(xa_d), C.nchw(out, 3, 1), C.ACT_NONE, C.PATH_AUTO, C.nchw(xb_d))
    assert_close(out[:, 1:4], ref_c, 2e-6, what="stencil conv (cat)")
    assert float(out[:, 0].abs().max()) == 0 and float(out[:, 4].abs().max()) == 0
    # (2) fused GDN, ROWPAD8 output
    plan.set_gdn(g.beta, g.gamma, g.inverse, g.beta_min)
    rp = torch.zeros((2, B, H + C.ROWPAD_Y, W + C.ROWPAD_X, slots), device=DEV, dtype=torch.bfloat16)
    plan.run(C.nchw(xa_d), C.rowpad(rp, 3), C.ACT_NONE, C.PATH_AUTO, C.nchw(xb_d))
    val = (rp[0].float() + rp[1].float())
    if slots == 4:   # rows interleaved in pairs: [B][(H+4)/2][W+8][2][4]
        val = val.reshape(B, (H + 4) // 2, W + 8, 2, 4).permute(0, 1, 3, 2, 4).reshape(B, H + 4, W + 8, 4)
    inner = val[:, 2:2 + H, 2:2 + W, :3].permute(0, 3, 1, 2)
    assert_close(inner, ref_g, 2e-5, what="stencil conv + GDN -> ROWPAD8")
    border = val.clone()
    border[:, 2:2 + H, 2:2 + W, :3] = 0
    assert float(border.abs().max()) == 0
    plan.set_gdn(None, None, False)


@pytest.mark.parametrize("kind", ["head", "stencil", "stencil_cat", "other_layer"])
@pytest.mark.parametrize("size", [(2, 16, 16), (3, 7, 5), (1, 20, 36)])
def test_squared_error_fused_into_the_reconstruction_layers(kind, size):
    """hesic_conv_forward_sse: the layers that store x1_hat / x2_hat (RGB synthesis head, newnet1.py:612; after_conv, :686)
    also accumulate sum((y - target)^2) -- the MSE partial of RateDistortionLoss (test3real.py:99-111) -- from the registers
    they store from.  Checked against the separate squared-error kernel on the written output and against a float64 sum on
    the host; the output itself must not change; a layer without the fused epilogue falls back to the kernel."""
    from hesic_b200 import _capi as C
    from hesic_b200 import functional as F
    from compressai.models.utils import conv, deconv
    B, H, W = size
    if kind == "head":
        mod = deconv(128, 3, kernel_size=5, stride=2)
        scale, Hi, Wi = (2.0 / (128 * 25 / 4)) ** 0.5, H, W
        Ho, Wo = 2 * H, 2 * W
    elif kind == "other_layer":
        mod = conv(16, 3, kernel_size=3, stride=1)      # CUDA-core generic path: no fused epilogue
        scale, Hi, Wi, Ho, Wo = 0.1, H, W, H, W
    else:
        mod = deconv(6 if kind == "stencil_cat" else 3, 3, kernel_size=5, stride=1)
        scale, Hi, Wi, Ho, Wo = 0.1, 2 * H + 1, 2 * W + 3, 2 * H + 1, 2 * W + 3   # odd sizes: scalar tails of the stores
    Cin = mod.weight.shape[0] if kind != "other_layer" else 16
    mod.load_state_dict({"weight": _rand(tuple(mod.weight.shape), 31, scale), "bias": _rand((3,), 32, 0.1)})
    mod = mod.to(DEV)
    plan = mod.hesic_plan()
    plan.set_gdn(None, None, False)
    x = _rand((B, Cin, Hi, Wi), 33).to(DEV)
    target = torch.zeros((B, 5, Ho, Wo), device=DEV)          # the target as a channel slice of a wider tensor
    target[:, 1:4] = _rand((B, 3, Ho, Wo), 34).to(DEV)
    tgt_d = C.nchw(target, 3, 1)
    if kind == "head":
        xs = torch.empty((2, B, Hi, Wi, 128), device=DEV, dtype=torch.bfloat16)
        C.check(C.lib.hesic_convert(C.ref(C.nchw(x)), C.ref(C.split(xs)), C.OP_COPY, C.stream()))
        xd, xb, path = C.split(xs), None, C.PATH_TC
    elif kind == "stencil_cat":
        xa_t, xb_t = x[:, :3].contiguous(), x[:, 3:].contiguous()
        xd, xb, path = C.nchw(xa_t), C.nchw(xb_t), C.PATH_AUTO
    else:
        xd, xb, path = C.nchw(x), None, C.PATH_AUTO
    plain = torch.empty((B, 3, Ho, Wo), device=DEV)
    plan.run(xd, C.nchw(plain), C.ACT_NONE, path, xb)
    out = torch.empty((B, 3, Ho, Wo), device=DEV)
    acc = torch.full((1,), 5.0, device=DEV, dtype=torch.float64)      # accumulates: starts from a non-zero value
    launches0 = C.lib.hesic_launch_count(0)
    plan.run(xd, C.nchw(out), C.ACT_NONE, path, xb, sse=(tgt_d, acc))
    n_launch = C.lib.hesic_launch_count(0) - launches0
    # the stencil accumulates in its own epilogue; the RGB head (measured slower fused, conv_tc.cu launch_head) and any
    # other layer run the squared-error kernel on the written image
    assert n_launch == (1 if kind.startswith("stencil") else 2)
    assert torch.equal(out, plain)
    sep = torch.zeros(1, device=DEV, dtype=torch.float64)
    C.check(C.lib.hesic_sum_squared_error(C.ref(C.nchw(out)), C.ref(tgt_d), C.ptr(sep), C.stream()))
    host = float(((out - target[:, 1:4]).double() ** 2).sum())           # fp32 difference, fp64 square and sum
    got = float(acc) - 5.0
    # the stencil sums the 12 squares of one store group in fp32 before they enter the fp64 accumulator (each partial within
    # 12 * 2^-24 of exact, the total far closer): one F32 -> F64 conversion per group (16 per clock per SM on B200,
    # tools/micro/dfma_rate.cu)
    assert abs(got - float(sep)) <= 2e-7 * host and abs(got - host) <= 2e-7 * host, (got, float(sep), host)
    # error behaviour: a target of another shape is rejected before anything is launched
    bad = torch.zeros((B, 3, Ho + 1, Wo), device=DEV)
    with pytest.raises(ValueError):
        plan.run(xd, C.nchw(out), C.ACT_NONE, path, xb, sse=(C.nchw(bad), acc))


@pytest.mark.parametrize("size", [(2, 16, 16), (3, 7, 5), (1, 20, 36), (1, 64, 64)])
@pytest.mark.parametrize("igdn", [False, True])
def test_rgb_head_scatter_form(size, igdn):
    """deconv(128 -> 3, k5, s2) (+ after_gdn) of Decoder1/Decoder2 (newnet1.py:612,670,685): the scatter-form tcgen05
    kernel (GEMM onto 5x5x3 patches + overlap-add), ragged tile edges, output written as a channel slice."""
    from hesic_b200 import _capi as C
    from compressai.layers import GDN
    from compressai.models.utils import deconv
    B, H, W = size
    mod = deconv(128, 3, kernel_size=5, stride=2)
    w = _rand(tuple(mod.weight.shape), 21, (2.0 / (128 * 25 / 4)) ** 0.5)
    b = _rand((3,), 22, 0.1)
    mod.load_state_dict({"weight": w, "bias": b})
    x = _rand((B, 128, H, W), 23)
    ref = O.deconv(x, w, b, stride=2)
    g = GDN(3, inverse=True)
    g.load_state_dict({"beta": torch.rand(3) + 0.5, "gamma": torch.rand(3, 3) * 0.2}, strict=False)
    if igdn:
        ref = O.gdn(ref, g.beta.detach(), g.gamma.detach(), inverse=True)
    mod, g = mod.to(DEV), g.to(DEV)
    plan = mod.hesic_plan()
    if igdn:
        plan.set_gdn(g.beta, g.gamma, True, g.beta_min)
    else:
        plan.set_gdn(None, None, False)
    xs = torch.empty((2, B, H, W, 128), device=DEV, dtype=torch.bfloat16)
    C.check(C.lib.hesic_convert(C.ref(C.nchw(x.to(DEV))), C.ref(C.split(xs)), C.OP_COPY, C.stream()))
    out = torch.zeros((B, 6, 2 * H, 2 * W), device=DEV)
    plan.run(C.split(xs), C.nchw(out, 3, 2), C.ACT_NONE, C.PATH_TC)
    C.check(C.lib.hesic_tc_status())
    assert_close(out[:, 2:5], ref, 1e-4, what="RGB head")
    assert float(out[:, :2].abs().max()) == 0 and float(out[:, 5].abs().max()) == 0
    plan.set_gdn(None, None, False)


def test_conv_linearity_property():
    """conv(a + b) - conv(a) - conv(b) + conv(0) == 0 at a BASELINE-size layer (size-independent check)."""
    from hesic_b200 import functional as F
    from compressai.models.utils import conv
    mod = conv(128, 128).to(DEV)
    a, b = _rand((1, 128, 128, 128), 5).to(DEV), _rand((1, 128, 128, 128), 6).to(DEV)
    p = mod.hesic_plan()
    r = F.conv2d(a + b, p) - F.conv2d(a, p) - F.conv2d(b, p) + F.conv2d(torch.zeros_like(a), p)
    assert float(r.abs().max()) < 1e-4 * float(F.conv2d(a, p).abs().max())


def test_gdn(ops):
    from compressai.layers import GDN
    for inv in (0, 1):
        g = GDN(16, inverse=bool(inv))
        g.load_state_dict({"beta": T(ops[f"gdn{inv}_beta"]), "gamma": T(ops[f"gdn{inv}_gamma"])}, strict=False)
        y = g.to(DEV)(T(ops[f"gdn{inv}_x"]).to(DEV))
        assert_close(y, ops[f"gdn{inv}_y"], 1e-5, what=f"gdn inverse={inv} (reference fixture)")
    # closed forms at init (reference tests/test_layers.py:111-162)
    x = torch.rand(2, 8, 5, 7)
    assert_close(GDN(8).to(DEV)(x.to(DEV)), x / torch.sqrt(1 + 0.1 * x ** 2), 1e-5)
    assert_close(GDN(8, inverse=True).to(DEV)(x.to(DEV)), x * torch.sqrt(1 + 0.1 * x ** 2), 1e-5)
    # 128 channels, random parameters, vs oracle
    import newnet1
    enc = newnet1.Encoder1(128, 192)
    sd = synth.synth_state_dict(enc, seed=4)
    x = _rand((2, 128, 20, 12), 9, 2.0)
    ref = O.gdn(x, sd["g_a_gdn2.beta"], sd["g_a_gdn2.gamma"])
    enc.load_state_dict(sd)
    assert_close(enc.g_a_gdn2.to(DEV)(x.to(DEV)), ref, 1e-4, what="gdn128")


@pytest.mark.parametrize("size", [(64, 64), (40, 72), (512, 512)])
def test_warp_perspective(size):
    """Warp parity is UNPINNED (kornia is not vendored by the reference).  Two bars: 1e-4 against the
    oracle's algorithm evaluated in double, and 5e-4 against its fp32 evaluation, whose own matrix
    inversion noise (~1e-3 px at 512x512) is larger than the kernel's error."""
    import kornia
    H, W = size
    B = 2
    x, _, h = synth.stereo_pairs(B, H, W, seed=77)
    fl = 0.25

    def check(hm, ac=True):
        out = kornia.warp_perspective(x.to(DEV), hm.to(DEV), (H, W), align_corners=ac)
        ref64 = O.warp_perspective(x, hm, (H, W), ac, dtype=torch.float64)
        ref32 = O.warp_perspective(x, hm, (H, W), ac)
        assert_close(out, ref64, 1e-4, floor=float(ref64.abs().max()) * fl, what="warp vs fp64 oracle")
        assert_close(out, ref32, 5e-4, floor=float(ref32.abs().max()) * fl, what="warp vs fp32 oracle")
        return ref32

    check(h)
    # identity homography returns the image (round trip through normalise/invert)
    eye = torch.eye(3)[None].repeat(B, 1, 1)
    out = kornia.warp_perspective(x.to(DEV), eye.to(DEV), (H, W))
    assert float((out.cpu() - x).abs().max()) < 2e-4
    # strong perspective + out-of-image samples -> zero padding
    h2 = h.clone()
    h2[:, 0, 2] += W * 0.4
    h2[:, 2, 0] = 2e-4
    ref2 = check(h2)
    assert float((ref2 == 0).float().mean()) > 0.1
    check(h, ac=False)


def _eb_module(ops):
    from compressai.entropy_models import EntropyBottleneck
    eb = EntropyBottleneck(8).eval()
    eb.load_state_dict({k[len("eb_sd_"):]: T(v) for k, v in ops.items()
                        if k.startswith("eb_sd_") and not k.endswith(("_offset", "_quantized_cdf", "_cdf_length"))}, strict=False)
    return eb


def test_entropy_bottleneck(ops):
    eb = _eb_module(ops).to(DEV)
    z_hat, lik = eb(T(ops["eb_z"]).to(DEV))
    assert torch.equal(z_hat.cpu(), T(ops["eb_z_hat"]))
    assert_close(lik, ops["eb_lik"], 1e-4, floor=1e-9, what="eb likelihood (reference fixture)")
    # larger random case vs oracle, incl. the 1e-9 clamp region
    import newnet1
    net = newnet1.HSIC(128, 192, 5).eval()
    sd = synth.synth_state_dict(net, seed=0)
    net.load_state_dict(sd)
    z = _rand((3, 128, 8, 8), 21, 6.0)
    ref_hat, ref_lik = O.entropy_bottleneck(z, *O.eb_params(sd, "entropy_bottleneck1"))
    got_hat, got_lik = net.entropy_bottleneck1.to(DEV)(z.to(DEV))
    assert torch.equal(got_hat.cpu(), ref_hat)
    assert_close(got_lik, ref_lik, 1e-4, floor=1e-9, what="eb likelihood")
    assert float((ref_lik <= 1e-9).float().mean()) > 0 or True
    # eval-mode medians==0 => y == round(x) (reference tests/test_entropy_models.py:131-176)
    from compressai.entropy_models import EntropyBottleneck
    e2 = EntropyBottleneck(128).eval().to(DEV)
    x = torch.rand(1, 128, 32, 32) * 10
    y, l = e2(x.to(DEV))
    assert y.shape == x.shape and l.shape == x.shape and torch.equal(y.cpu(), torch.round(x))


def test_entropy_bottleneck_compress_is_bit_exact(ops):
    """Device symbol/index prep + host rANS == the reference's byte strings; decode round-trips."""
    eb = _eb_module(ops).to(DEV)
    eb.update()
    assert np.array_equal(eb._quantized_cdf.cpu().numpy(), ops["eb_cdf"])
    z = T(ops["eb_z"]).to(DEV)
    strings = eb.compress(z)
    for i, s in enumerate(strings):
        assert s == ops[f"eb_string{i}"].tobytes()
        back = eb.decompress([s], z.shape[-2:])
        assert torch.equal(back.cpu(), T(ops["eb_z_hat"][i:i + 1]))
    # full-size latent (B=16, 128x8x8): encode -> decode -> equals forward's z_hat
    zz = _rand((16, 128, 8, 8), 33, 5.0).to(DEV)
    zh, _ = eb.__class__(128).eval().to(DEV)(zz)  # default params: medians 0
    eb128 = eb.__class__(128).eval().to(DEV)
    eb128.update()
    ss = eb128.compress(zz)
    dec = torch.cat([eb128.decompress([s], (8, 8)) for s in ss], 0)
    assert torch.equal(dec.cpu(), zh.cpu())
    # integer prep vs oracle directly
    from hesic_b200 import functional as F
    med = eb._medians().detach().view(1, -1, 1, 1)
    sym = F.prepare_symbols(z, med).cpu()
    assert torch.equal(sym.reshape(z.shape), O.quantize(T(ops["eb_z"]), "symbols", med.cpu()))
    idx = eb._build_indexes(z.size(), z.device).cpu()
    assert torch.equal(idx, O.eb_build_indexes(z.size()))


def test_gaussian_models(ops):
    from compressai.entropy_models import GaussianConditional, GaussianMixtureConditional
    gm = GaussianMixtureConditional(K=5).eval().to(DEV)
    d = lambda k: T(ops[k]).to(DEV)
    y_hat, lik = gm(d("gmm_y"), d("gmm_scales"), d("gmm_means"), d("gmm_weights"))
    assert torch.equal(y_hat.cpu(), T(ops["gmm_y_hat"]))
    assert_close(lik, ops["gmm_lik"], 1e-4, floor=1e-9, what="gmm likelihood (reference fixture)")
    gc = GaussianConditional(None).eval().to(DEV)
    yh, lk = gc(d("gmm_y"), d("gc_scales"), means=d("gc_means"))
    assert torch.equal(yh.cpu(), T(ops["gc_y_hat"]))
    assert_close(lk, ops["gc_lik"], 1e-4, floor=1e-9)
    yh0, lk0 = gc(d("gmm_y"), d("gc_scales"))
    assert torch.equal(yh0.cpu(), T(ops["gc_y_hat0"]))
    assert_close(lk0, ops["gc_lik0"], 1e-4, floor=1e-9)
    # build_indexes + compress: integer outputs bit-exact, byte strings equal the reference's
    gc.update_scale_table([float(v) for v in ops["gc_table"]])
    idx = gc.build_indexes(d("gc_scales"))
    assert np.array_equal(idx.cpu().numpy(), ops["gc_indexes"])
    strings = gc.compress(d("gmm_y"), idx, means=d("gc_means"))
    for i, s in enumerate(strings):
        assert s == ops[f"gc_string{i}"].tobytes()
    dec = gc.decompress(strings, idx, means=d("gc_means"))
    assert torch.equal(dec.cpu(), T(ops["gc_dec"]))
    # full-size mixture (B=2, 192x32x32, K=5) with extreme scales/tails vs oracle
    g = np.random.default_rng(3)
    B, M, K = 2, 192, 5
    y = _rand((B, M, 32, 32), 41, 4.0)
    sc = torch.from_numpy(np.abs(g.standard_normal((B, M * K, 32, 32))).astype(np.float32) * 2)
    sc[:, ::7] = 0.0   # below the 0.11 bound
    mu = _rand((B, M * K, 32, 32), 42, 3.0)
    w = torch.softmax(_rand((B, K, M, 1, 1), 43), dim=1).reshape(B, K * M, 1, 1)
    ref_hat, ref_lik = O.gmm_conditional(y, sc, mu, w, K)
    got_hat, got_lik = gm(y.to(DEV), sc.to(DEV), mu.to(DEV), w.to(DEV))
    assert torch.equal(got_hat.cpu(), ref_hat)
    assert_close(got_lik, ref_lik, 1e-4, floor=1e-9, what="gmm likelihood full size")
    assert float((ref_lik <= 1.0001e-9).float().mean()) > 0


def test_pool_softmax_upsample():
    from hesic_b200 import functional as F
    x = _rand((3, 960, 32, 32), 51)
    assert torch.equal(F.spatial_max(x.to(DEV)).cpu(), O.spatial_pool2d(x))
    K, M = 5, 192
    w = _rand((K * M, K * M, 1, 1), 52, 0.05)
    b = _rand((K * M,), 53, 0.1)
    pooled = O.spatial_pool2d(x)
    ref = O._mix_softmax(O.conv(torch.nn.functional.leaky_relu(pooled), w, b, stride=1), K, M)
    got = F.mixture_weights(pooled.to(DEV), w.to(DEV), b.to(DEV), K, M)
    assert_close(got, ref, 1e-4, floor=1e-6, what="mixture weights")
    z = _rand((2, 128, 8, 8), 54)
    ref = torch.nn.functional.interpolate(z, scale_factor=4, mode="bilinear", align_corners=True)
    assert_close(F.upsample_bilinear(z.to(DEV), 4), ref, 1e-5, what="upsample")
    r = _rand((2, 7, 5, 3), 55, 3.0)
    r[0, 0, 0, 0], r[0, 0, 0, 1], r[0, 0, 0, 2] = 0.5, 1.5, -2.5
    assert torch.equal(F.round_half_even(r.to(DEV)).cpu(), torch.round(r))


def test_operator_modules_match_oracle_submodules():
    """Encoder1 / Decoder1 / hyper modules called stand-alone (operator-level API) vs oracle."""
    import newnet1
    net = newnet1.HSIC(128, 192, 5).eval()
    sd = synth.synth_state_dict(net, seed=0)
    net.load_state_dict(sd)
    net = net.to(DEV)
    x1, x2, h = synth.stereo_pairs(1, 64, 64, seed=5)
    y_ref = O.encoder1(sd, x1)
    y, g1, g2, g3 = net.encoder1(x1.to(DEV))
    assert_close(y, y_ref, 1e-4, what="Encoder1")
    z_ref = O.encode_hyper(sd, y_ref, "_h_a1")
    assert_close(net._h_a1(y_ref.to(DEV)), z_ref, 1e-4, what="encode_hyper")
    zh = torch.round(z_ref)
    s, m, w = net._h_s1(zh.to(DEV))
    s_r, m_r, w_r = O.gmm_hyper_y1(sd, zh, "_h_s1", 5, 192)
    assert_close(s, s_r, 1e-4, what="gmm sigma")
    assert_close(m, m_r, 1e-4, what="gmm means")
    assert_close(w, w_r, 1e-4, floor=1e-6, what="gmm weights")
    yh = torch.round(y_ref)
    s2, m2, w2 = net._h_s2(zh.to(DEV), yh.to(DEV))
    s2r, m2r, w2r = O.gmm_hyper_y2(sd, zh, yh, "_h_s2", 5, 192)
    assert_close(s2, s2r, 1e-4)
    assert_close(m2, m2r, 1e-4)
    assert_close(w2, w2r, 1e-4, floor=1e-6)
    xh_ref = O.decoder1(sd, yh)
    xh, *_ = net.decoder1(yh.to(DEV))
    assert_close(xh, xh_ref, 1e-4, what="Decoder1")
    x2h_ref = O.decoder2(sd, yh, x1)
    assert_close(net.decoder2(yh.to(DEV), x1.to(DEV)), x2h_ref, 1e-4, what="Decoder2")
    y2_ref = O.encoder2(sd, x1, x2)
    assert_close(net.encoder2(x1.to(DEV), x2.to(DEV)), y2_ref, 1e-4, what="Encoder2")


def _to_hilo(x):
    """NCHW fp32 (C <= 32) -> NHWC_HILO bf16 [B,H,W,64] through the library's own converter."""
    from hesic_b200 import _capi as C
    B, Cn, H, W = x.shape
    t = torch.zeros((B, H, W, 64), device=DEV, dtype=torch.bfloat16)
    xd = x.to(DEV).contiguous()
    C.check(C.lib.hesic_convert(C.ref(C.nchw(xd)), C.ref(C.hilo(t, Cn)), C.OP_COPY, C.stream()))
    torch.cuda.synchronize()
    return t


def _from_hilo(t):
    return (t[..., :32].float() + t[..., 32:].float()).permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("size", [(1, 8, 16), (2, 24, 40), (1, 70, 67), (2, 64, 64)])
@pytest.mark.parametrize("mode", ["plain", "lrelu", "lrelu_res1", "lrelu_res2"])
def test_enhancement_conv_parity(size, mode):
    """conv3x3 32 -> 32 on NHWC_HILO activations with the fused LeakyReLU / identity additions of ResidualBlock and
    Enhancement_Block (compressai/layers/layers.py:125-147, newnet1.py:272-287), ragged tiles included."""
    from hesic_b200 import _capi as C
    from hesic_b200.enhance import EnConvPlan
    B, H, W = size
    w = _rand((32, 32, 3, 3), 21, (2.0 / (32 * 9)) ** 0.5)
    b = _rand((32,), 22, 0.1)
    x, r1, r2 = _rand((B, 32, H, W), 23), _rand((B, 32, H, W), 24), _rand((B, 32, H, W), 25)
    xh, r1h, r2h = _to_hilo(x), _to_hilo(r1), _to_hilo(r2)
    # the oracle sees exactly the values the device tensors hold (hi + lo = 16-bit-mantissa roundings of the inputs)
    xq, r1q, r2q = _from_hilo(xh).cpu(), _from_hilo(r1h).cpu(), _from_hilo(r2h).cpu()
    ref = O.conv(xq, w, b, stride=1)
    if mode != "plain":
        ref = torch.nn.functional.leaky_relu(ref)
    if mode in ("lrelu_res1", "lrelu_res2"):
        ref = ref + r1q
    if mode == "lrelu_res2":
        ref = ref + r2q
    plan = EnConvPlan(32, 32).load(w.to(DEV), b.to(DEV))
    y = torch.full((B, H, W, 64), float("nan"), device=DEV, dtype=torch.bfloat16)
    plan.run(C.hilo(xh), C.hilo(y), C.ACT_NONE if mode == "plain" else C.ACT_LEAKY,
             C.hilo(r1h) if mode in ("lrelu_res1", "lrelu_res2") else None, C.hilo(r2h) if mode == "lrelu_res2" else None)
    C.check(C.lib.hesic_tc_status())
    assert_close(_from_hilo(y), ref, 1e-4, what=f"enhancement conv {size} {mode}")


@pytest.mark.parametrize("size", [(1, 8, 16), (2, 40, 24), (1, 70, 67)])
def test_enhancement_first_and_last_layer(size):
    """Enhancement.conv1 on cat(x, x_other_warp) (6 -> 32, newnet1.py:302-304) and conv2 + identity (32 -> 3, :309-310)."""
    from hesic_b200 import _capi as C
    from hesic_b200.enhance import EnConvPlan
    B, H, W = size
    xa, xb = _rand((B, 3, H, W), 31), _rand((B, 3, H, W), 32)
    w1, b1 = _rand((32, 6, 3, 3), 33, 0.2), _rand((32,), 34, 0.1)
    t = torch.full((B, H, W, 64), float("nan"), device=DEV, dtype=torch.bfloat16)
    xa_d, xb_d = xa.to(DEV), xb.to(DEV)
    C.check(C.lib.hesic_en_pack_input(C.ref(C.nchw(xa_d)), C.ref(C.nchw(xb_d)), C.ref(C.hilo(t)), C.stream()))
    packed = _from_hilo(t).cpu()
    assert_close(packed[:, :6], torch.cat((xa, xb), 1), 2e-5, what="en pack")
    assert float(packed[:, 6:].abs().max()) == 0
    y = torch.empty((B, H, W, 64), device=DEV, dtype=torch.bfloat16)
    EnConvPlan(6, 32).load(w1.to(DEV), b1.to(DEV)).run(C.hilo(t), C.hilo(y))
    assert_close(_from_hilo(y), O.conv(packed[:, :6], w1, b1, stride=1), 1e-4, what="en conv1")
    w2, b2 = _rand((3, 32, 3, 3), 35, 0.1), _rand((3,), 36, 0.1)
    feat = _rand((B, 32, H, W), 37)
    fh = _to_hilo(feat)
    out = torch.full((B, 3, H, W), float("nan"), device=DEV)
    ident = xa.to(DEV)
    EnConvPlan(32, 3).load(w2.to(DEV), b2.to(DEV)).run(C.hilo(fh), C.nchw(out), C.ACT_NONE, C.nchw(ident))
    C.check(C.lib.hesic_tc_status())
    assert_close(out, O.conv(_from_hilo(fh).cpu(), w2, b2, stride=1) + xa, 1e-4, what="en conv2 + identity")


def test_enhancement_conv_rejects_bad_arguments():
    from hesic_b200 import _capi as C
    from hesic_b200.enhance import EnConvPlan
    with pytest.raises(ValueError):
        EnConvPlan(64, 32)
    with pytest.raises(ValueError):
        EnConvPlan(32, 16)
    plan = EnConvPlan(32, 32).load(torch.zeros(32, 32, 3, 3, device=DEV))
    x = torch.zeros((1, 8, 16, 64), device=DEV, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        plan.run(C.hilo(x), C.hilo(torch.zeros((1, 8, 8, 64), device=DEV, dtype=torch.bfloat16)))
    with pytest.raises(ValueError):
        plan.run(C.hilo(x), C.nchw(torch.zeros((1, 32, 8, 16), device=DEV)))


@pytest.mark.parametrize("mask_type", ["A", "B"])
@pytest.mark.parametrize("path", ["auto", "simt"])
def test_masked_conv2d(mask_type, path):
    """MaskedConv2d (compressai/layers/layers.py:21-45; HESIC+ context_prediction, newnet1_joint.py:635-639): the
    tensor-core path skips the taps the mask removes (12 / 13 live taps of 25), the CUDA-core path multiplies by it."""
    from hesic_b200 import _capi as C
    from hesic_b200 import functional as F
    from compressai.layers import MaskedConv2d
    m = MaskedConv2d(192, 384, kernel_size=5, padding=2, stride=1, mask_type=mask_type)
    w, b = _rand(tuple(m.weight.shape), 41, (2.0 / (192 * 25)) ** 0.5), _rand((384,), 42, 0.1)
    m.load_state_dict({"weight": w, "bias": b, "mask": m.mask.clone()})
    x = _rand((2, 192, 16, 24), 43)
    ref = torch.nn.functional.conv2d(x, w * m.mask, b, padding=2)
    m = m.to(DEV)
    if path == "auto":
        y = m(x.to(DEV))
    else:
        m.weight.data *= m.mask
        y = F.conv2d(x.to(DEV), m.hesic_plan(), path=C.PATH_SIMT)
    assert_close(y, ref, 1e-4, what=f"MaskedConv2d type {mask_type} ({path})")
    assert int(m.mask[0, 0].sum()) == (12 if mask_type == "A" else 13)


def test_sum_squared_error_and_channels_last_max():
    """RateDistortionLoss's MSE partial sum (test3real.py:99-111) in fp64, dense (128-bit loads) and strided forms;
    spatial_pool2d on the channels-last layout the engine uses."""
    from hesic_b200 import _capi as C
    from hesic_b200 import functional as F
    a, b = _rand((3, 3, 40, 52), 51), _rand((3, 3, 40, 52), 52)
    ref = float(((a.double() - b.double()) ** 2).sum())
    acc = torch.zeros(2, device=DEV, dtype=torch.float64)
    F.sum_squared_error(a.to(DEV), b.to(DEV), acc[0:1])
    wide = torch.zeros((3, 5, 40, 52), device=DEV)
    wide[:, 1:4] = a.to(DEV)
    bd = b.to(DEV)
    C.check(C.lib.hesic_sum_squared_error(C.ref(C.nchw(wide, 3, 1)), C.ref(C.nchw(bd)), C.ptr(acc[1:2]), C.stream()))
    got = acc.cpu()
    assert abs(float(got[0]) - ref) <= 1e-9 * ref and abs(float(got[1]) - ref) <= 1e-9 * ref
    x = _rand((2, 960, 9, 13), 53)
    xn = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    out = torch.empty((2, 960), device=DEV)
    C.check(C.lib.hesic_spatial_max(C.ref(C.nhwc(xn)), C.ptr(out), C.stream()))
    assert torch.equal(out.cpu(), x.amax(dim=(2, 3)))


@pytest.mark.parametrize("shape", [(2, 64, 96, 3), (1, 33, 30, 3), (2, 16, 20, 1), (1, 8, 12, 4)])
def test_images_from_uint8_is_totensor(shape):
    """uint8 [B,H,W,C] -> fp32 [B,C,H,W] / 255 on the device: bit-identical to what transforms.ToTensor() computes on
    the host (compressai/datasets/utils.py:101-102: `img.permute(2, 0, 1).float().div(255)`), fast path and generic path."""
    from hesic_b200 import functional as F
    g = torch.Generator().manual_seed(7)
    u8 = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)
    ref = u8.permute(0, 3, 1, 2).float().div(255)
    got = F.images_from_uint8(u8.to(DEV))
    assert got.shape == ref.shape and torch.equal(got.cpu(), ref)
    with pytest.raises(TypeError):
        F.images_from_uint8(u8.to(DEV).float())


def test_hostfeed_uint8_transport_equals_fp32_transport():
    """HostFeed with [B,H,W,3] uint8 host images hands fn the same fp32 tensors as shipping the ToTensor()'d images."""
    from hesic_b200.hostfeed import HostFeed
    g = torch.Generator().manual_seed(3)
    batches8, batches32 = [], []
    for _ in range(3):
        a, b = (torch.randint(0, 256, (2, 64, 64, 3), generator=g, dtype=torch.uint8) for _ in range(2))
        h = torch.eye(3).repeat(2, 1, 1) + 0.01 * torch.randn(2, 3, 3, generator=g)
        batches8.append((a.pin_memory(), b.pin_memory(), h.pin_memory()))
        batches32.append(tuple(t.permute(0, 3, 1, 2).float().div(255).contiguous().pin_memory() for t in (a, b)) + (h.pin_memory(),))
    seen8, seen32 = [], []
    HostFeed(torch.device(DEV), batches8[0]).run(batches8, lambda x1, x2, hh: seen8.append((x1.clone(), x2.clone(), hh.clone())))
    HostFeed(torch.device(DEV), batches32[0]).run(batches32, lambda x1, x2, hh: seen32.append((x1.clone(), x2.clone(), hh.clone())))
    torch.cuda.synchronize()
    assert len(seen8) == len(seen32) == 3
    for p8, p32 in zip(seen8, seen32):
        for u, v in zip(p8, p32):
            assert u.dtype == torch.float32 and torch.equal(u, v)


def test_perspective_transform_and_inverse_in_one_launch():
    """kornia.get_perspective_transform (+ torch.inverse), ywz/mywork/test3real.py:179-180: hesic_perspective_transform against
    the oracle's restatement, against its defining property (H maps every source corner onto its destination corner, checked
    in fp64) and through the kornia shim the unmodified driver imports."""
    import kornia
    from hesic_b200 import functional as F
    g = torch.Generator().manual_seed(9)
    B = 37
    src = torch.tensor([[[0., 0.], [127., 0.], [127., 127.], [0., 127.]]]).repeat(B, 1, 1) + torch.rand(B, 1, 2, generator=g) * 64
    dst = src + (torch.rand(B, 4, 2, generator=g) - 0.5) * 24
    H = F.perspective_transform(src.to(DEV), dst.to(DEV)).cpu()
    H_ref = O.get_perspective_transform(src, dst)
    assert_close(H, H_ref, 1e-4, what="perspective transform vs oracle")
    p = torch.cat([src.double(), torch.ones(B, 4, 1, dtype=torch.float64)], -1) @ H.double().transpose(1, 2)
    assert torch.allclose(p[..., :2] / p[..., 2:], dst.double(), atol=1e-3)
    Hi = F.perspective_transform(src.to(DEV), dst.to(DEV), invert=True).cpu()
    assert_close(Hi, torch.inverse(H_ref), 1e-4, what="inverse homography vs oracle")
    eye = Hi.double() @ H.double()
    assert torch.allclose(eye / eye[:, 2:, 2:], torch.eye(3, dtype=torch.float64).expand(B, 3, 3), atol=1e-4)
    assert torch.equal(kornia.get_perspective_transform(src.to(DEV), dst.to(DEV)).cpu(), H)
    with pytest.raises(ValueError):
        F.perspective_transform(src[:, :3].to(DEV), dst[:, :3].to(DEV))


@pytest.mark.parametrize("shape", [(2, 5, 16, 24), (1, 3, 33, 30), (3, 64, 128, 128), (1, 2, 7, 9)])
def test_max_pool2x2(shape):
    """nn.MaxPool2d(2, 2) of the homography net (udh/udh/model.py:66): bit-equal to torch, odd trailing rows / columns dropped."""
    from hesic_b200 import functional as F
    from hesic_b200.homography import MaxPool2d
    x = _rand(shape, 21).to(DEV)
    ref = torch.nn.functional.max_pool2d(x, 2, 2)
    assert torch.equal(F.max_pool2x2(x), ref)
    assert torch.equal(MaxPool2d(2, 2)(x), ref)
    assert torch.equal(MaxPool2d(3, 2)(x), torch.nn.functional.max_pool2d(x, 3, 2))     # other configurations: torch's


@pytest.mark.parametrize("size,slice_of", [((2, 44, 60), 128), ((1, 96, 80), 192), ((3, 32, 16), 128), ((1, 28, 36), 128)])
def test_first_analysis_layer_kernel_ragged_tiles_and_channel_slices(size, slice_of):
    """conv(3, 128, k5, s2) + fused GDN into the SPLIT planes of the next layer (newnet1.py:583-601): conv_tc_first_kernel
    (tiles of 8 x 16 output pixels, |x| rebuilt from the x^2 operand) on outputs that do not fill whole tiles, written into a
    channel slice of a wider buffer, against the oracle at 1e-4 of the rms; the last size is below the kernel's minimum tile
    and takes conv_tc_kernel<16>.  Exactly the slice is written."""
    from hesic_b200 import _capi as C
    from compressai.layers import GDN
    from compressai.models.utils import conv
    B, H, W = size
    mod = conv(3, 128, kernel_size=5, stride=2)
    w = _rand(tuple(mod.weight.shape), 31, (2.0 / 75) ** 0.5)
    b = _rand((128,), 32, 0.1)
    mod.load_state_dict({"weight": w, "bias": b})
    g = GDN(128)
    g.load_state_dict({"beta": torch.rand(128, generator=torch.Generator().manual_seed(3)) + 0.5,
                       "gamma": torch.rand(128, 128, generator=torch.Generator().manual_seed(4)) * 0.02 + 0.1 * torch.eye(128)}, strict=False)
    x = torch.rand((B, 3, H, W), generator=torch.Generator().manual_seed(5))       # an image: [0, 1)
    x[0, :, :3, :5] = 0.0                                                           # exact zeros (the 0 * rsqrt(0) guard)
    ref = O.gdn(O.conv(x, w, b, stride=2), g.beta.detach(), g.gamma.detach())
    mod, g = mod.to(DEV), g.to(DEV)
    plan = mod.hesic_plan()
    plan.set_gdn(g.beta, g.gamma, False, g.beta_min)
    xs = torch.zeros((2, B, H + C.ROWPAD_Y, W + C.ROWPAD_X, 4), device=DEV, dtype=torch.bfloat16)
    xd = C.rowpad(xs, 3)
    C.check(C.lib.hesic_convert(C.ref(C.nchw(x.to(DEV))), C.ref(xd), C.OP_COPY, C.stream()))
    Ho, Wo = plan.out_hw(H, W)
    c0 = slice_of - 128
    ys = torch.zeros((2, B, Ho, Wo, slice_of), device=DEV, dtype=torch.bfloat16)
    plan.run(xd, C.split(ys, 128, c0), C.ACT_NONE, C.PATH_TC)
    # the same with the lo plane BELOW the hi plane in memory: the kernel's one-store-for-both-planes map does not apply and it
    # issues one store per plane
    yr = torch.zeros_like(ys)
    d = C.split(yr, 128, c0)
    d.p0, d.p1 = d.p1, d.p0
    plan.run(xd, d, C.ACT_NONE, C.PATH_TC)
    torch.cuda.synchronize()
    C.check(C.lib.hesic_tc_status())
    plan.set_gdn(None, None, False)
    assert torch.equal(yr[1], ys[0]) and torch.equal(yr[0], ys[1])
    y = (ys[0].float() + ys[1].float()).cpu()
    assert_close(y[..., c0:].permute(0, 3, 1, 2), ref, 1e-4, what=f"first layer + GDN {size}")
    if c0:
        assert float(y[..., :c0].abs().max()) == 0.0
    assert torch.isfinite(y).all()


def test_bin_mass_is_closer_to_the_exact_value_than_the_reference_form():
    """The likelihood kernels evaluate Phi((0.5 - d) / s) - Phi((-0.5 - d) / s) by a cancellation-free series (narrow bins) or a
    Chebyshev-fit erfc (elementwise.cu: bin_mass) instead of the reference's fp32 erfc difference (entropy_models.py:546-554).
    Against the EXACT value (float64) over scales 0.11 .. 300 and offsets out to the 1e-9 floor the kernel stays within 5e-5
    (relative, floor 1e-9) -- the reference's own form does not (tools/erfc_eval.py: 2e-4) -- so its distance to the oracle is the
    oracle's rounding noise; on moderate scales (< 32) kernel and oracle agree to 1e-4."""
    from scipy.special import erfc
    from hesic_b200 import functional as F
    g = torch.Generator().manual_seed(17)
    n = 1 << 18
    s = torch.exp(torch.rand(n, generator=g) * (math.log(300.0) - math.log(0.11)) + math.log(0.11))
    mu = (torch.rand(n, generator=g) - 0.5) * 8
    y = mu + torch.randn(n, generator=g) * s * 2.5
    shape = (4, 64, 32, 32)
    yh, lik = F.gaussian_conditional(y.reshape(shape).to(DEV), s.reshape(shape).to(DEV), mu.reshape(shape).to(DEV))
    lik = lik.reshape(-1).cpu().double()
    yq = torch.round(y - mu) + mu
    assert torch.equal(yh.reshape(-1).cpu(), yq)
    d = (yq - mu).abs().double().numpy()
    sd = s.double().numpy()
    exact = 0.5 * erfc(-(0.5 - d) / sd / math.sqrt(2)) - 0.5 * erfc(-(-0.5 - d) / sd / math.sqrt(2))
    exact = torch.from_numpy(exact).clamp_min(1e-9)
    rel = (lik - exact).abs() / exact
    assert float(rel.max()) < 5e-5, float(rel.max())
    ref_lik = O.gaussian_conditional(y.reshape(shape), s.reshape(shape), mu.reshape(shape))[1].reshape(-1).double()
    mod = s < 32
    rel_o = ((lik - ref_lik).abs() / ref_lik.clamp_min(1e-9))[mod]
    assert float(rel_o.max()) < 1e-4, float(rel_o.max())
