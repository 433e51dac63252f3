"""The oracle (CPU restatement) against the fixtures the UNMODIFIED reference produced
(tests/golden/make_golden.py).  Runs without a GPU."""
import math

import numpy as np
import pytest
import torch

import oracle
from hesic_b200 import synth
from oracle import hesic_oracle as O
from tests.helpers import T, assert_close, load_json, load_npz, mismatch_fraction


@pytest.fixture(scope="module")
def ops():
    return load_npz("operators")


def test_gdn(ops):
    for inv in (0, 1):
        y = O.gdn(T(ops[f"gdn{inv}_x"]), T(ops[f"gdn{inv}_beta"]), T(ops[f"gdn{inv}_gamma"]), bool(inv))
        assert_close(y, ops[f"gdn{inv}_y"], 1e-6, what=f"gdn inverse={inv}")


def _eb(ops):
    mats = [T(ops[f"eb_sd__matrices.{i}"]) for i in range(5)]
    bias = [T(ops[f"eb_sd__biases.{i}"]) for i in range(5)]
    fac = [T(ops[f"eb_sd__factors.{i}"]) for i in range(4)]
    return mats, bias, fac, T(ops["eb_sd_quantiles"])


def test_entropy_bottleneck_forward(ops):
    z_hat, lik = O.entropy_bottleneck(T(ops["eb_z"]), *_eb(ops))
    assert torch.equal(z_hat, T(ops["eb_z_hat"]))
    assert_close(lik, ops["eb_lik"], 1e-5, floor=1e-9, what="eb likelihood")


def test_entropy_bottleneck_tables_and_stream(ops):
    pmf, tail, length, offset = O.eb_update_pmf(*_eb(ops))
    assert np.array_equal(offset.numpy(), ops["eb_offset"])
    assert np.array_equal((length + 2).numpy(), ops["eb_cdf_length"])
    cdf = np.zeros_like(ops["eb_cdf"])
    for c in range(pmf.shape[0]):
        p = torch.cat((pmf[c, : length[c]], tail[c]), dim=0).numpy()
        q = oracle.pmf_to_quantized_cdf(p, 16)
        cdf[c, : q.size] = q
    assert np.array_equal(cdf, ops["eb_cdf"])
    z = T(ops["eb_z"])
    med = T(ops["eb_sd_quantiles"])[:, :, 1:2].reshape(1, -1, 1, 1)
    sym = O.quantize(z, "symbols", med)
    idx = O.eb_build_indexes(z.shape)
    for i in range(z.shape[0]):
        s = oracle.rans_encode(sym[i].reshape(-1).numpy(), idx[i].reshape(-1).numpy(), ops["eb_cdf"], ops["eb_cdf_length"],
                               ops["eb_offset"])
        assert s == ops[f"eb_string{i}"].tobytes()
        dec = oracle.rans_decode(s, idx[i].reshape(-1).numpy(), ops["eb_cdf"], ops["eb_cdf_length"], ops["eb_offset"])
        assert np.array_equal(dec, sym[i].reshape(-1).numpy())


def test_gmm_conditional(ops):
    y_hat, lik = O.gmm_conditional(T(ops["gmm_y"]), T(ops["gmm_scales"]), T(ops["gmm_means"]), T(ops["gmm_weights"]), 5)
    assert torch.equal(y_hat, T(ops["gmm_y_hat"]))
    assert_close(lik, ops["gmm_lik"], 1e-5, floor=1e-9, what="gmm likelihood")


def test_gaussian_conditional(ops):
    y = T(ops["gmm_y"])
    yh, lk = O.gaussian_conditional(y, T(ops["gc_scales"]), T(ops["gc_means"]))
    assert torch.equal(yh, T(ops["gc_y_hat"]))
    assert_close(lk, ops["gc_lik"], 1e-5, floor=1e-9)
    yh0, lk0 = O.gaussian_conditional(y, T(ops["gc_scales"]))
    assert torch.equal(yh0, T(ops["gc_y_hat0"]))
    assert_close(lk0, ops["gc_lik0"], 1e-5, floor=1e-9)
    tab = O.scale_table()
    assert np.allclose(tab.numpy(), ops["gc_table"], rtol=1e-6)
    idx = O.build_indexes(T(ops["gc_scales"]), T(ops["gc_table"]))
    assert np.array_equal(idx.numpy(), ops["gc_indexes"])
    sym = O.quantize(y, "symbols", T(ops["gc_means"]))
    for i in range(y.shape[0]):
        s = oracle.rans_encode(sym[i].reshape(-1).numpy(), idx[i].reshape(-1).numpy(), ops["gc_cdf"], ops["gc_cdf_length"],
                               ops["gc_offset"])
        assert s == ops[f"gc_string{i}"].tobytes()


def test_masked_conv(ops):
    sd = {"m.weight": T(ops["mc_w"]), "m.bias": T(ops["mc_b"]), "m.mask": T(ops["mc_mask"])}
    assert_close(O.masked_conv(sd, "m", T(ops["mc_x"])), ops["mc_y"], 1e-6)
    mask = ops["mc_mask"]
    assert mask[:, :, 2, 2:].sum() == 0 and mask[:, :, 3:].sum() == 0 and mask[:, :, :2].min() == 1


def test_pmf_to_quantized_cdf(ops):
    for i in range(4):
        assert np.array_equal(oracle.pmf_to_quantized_cdf(ops[f"pmf{i}"], 16), ops[f"pmf{i}_cdf"])


def test_rans_known_answer(ops):
    s = oracle.rans_encode(ops["rans_symbols"], ops["rans_indexes"], ops["rans_cdfs"], ops["rans_sizes"], ops["rans_offsets"])
    assert s == ops["rans_stream"].tobytes()
    dec = oracle.rans_decode(s, ops["rans_indexes"], ops["rans_cdfs"], ops["rans_sizes"], ops["rans_offsets"])
    assert np.array_equal(dec, ops["rans_symbols"])


def test_coder_oracle_vs_reference_build(ops):
    """The C restatement against the reference's own compiled C++ (oracle/_ref), when it travelled."""
    ans = oracle.ref_ext("ans")
    cxx = oracle.ref_ext("_CXX")
    if ans is None or cxx is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    g = np.random.default_rng(11)
    for n in (2, 5, 33, 200):
        p = g.dirichlet(np.ones(n) * 0.3).astype(np.float32)
        assert np.array_equal(oracle.pmf_to_quantized_cdf(p, 16), np.asarray(cxx.pmf_to_quantized_cdf(p.tolist(), 16)))
    cdfs, sizes, offs = ops["rans_cdfs"], ops["rans_sizes"], ops["rans_offsets"]
    for n in (1, 2, 7, 1000):
        idx = g.integers(0, 3, n).astype(np.int32)
        sym = np.array([g.integers(offs[i] - 40, offs[i] + sizes[i] + 40) for i in idx], dtype=np.int32)
        ref = ans.RansEncoder().encode_with_indexes(sym.tolist(), idx.tolist(), cdfs.tolist(), sizes.tolist(), offs.tolist())
        assert oracle.rans_encode(sym, idx, cdfs, sizes, offs) == ref


def _model_sd(ctor):
    net = ctor().eval()
    return net, synth.synth_state_dict(net, seed=0)


@pytest.mark.parametrize("name", ["hsic_newnet1", "hsic_newnet9", "hsic_joint"])
def test_full_forward_vs_reference(name):
    """Oracle forward vs the reference's stored outputs.  x.5 rounding may flip a symbol on a
    different host CPU (fp32 summation order), so symbols are compared as a mismatch fraction
    and images by relative L2."""
    from hesic_b200 import compat
    compat.install()
    meta, gold = load_json(name), load_npz(name)
    if name == "hsic_joint":
        import newnet1_joint as mod
    elif name == "hsic_newnet9":
        import newnet9 as mod
    else:
        import newnet1 as mod
    net, sd = _model_sd(lambda: mod.HSIC(128, 192, 5))
    x1, x2, h = synth.stereo_pairs(meta["B"], meta["H"], meta["W"], seed=1234)
    with torch.no_grad():
        if name == "hsic_joint":
            out = O.hsic_joint_forward(sd, x1, x2, h)
        else:
            out = O.hsic_forward(sd, x1, x2, h, twice_left=(name == "hsic_newnet1"))
    if "y1_hat" in gold:
        assert mismatch_fraction(out["y1_hat"], gold["y1_hat"].astype(np.float32)) < 2e-3
        assert mismatch_fraction(out["y2_hat"], gold["y2_hat"].astype(np.float32)) < 2e-3
    for k in ("x1_hat", "x2_hat"):
        d = (out[k].double() - T(gold[k]).double()).pow(2).sum().sqrt() / T(gold[k]).double().pow(2).sum().sqrt()
        assert float(d) < 1e-2, (k, float(d))
    m = synth.rd_metrics(out, x1, x2)
    for k, v in meta["metrics"].items():
        assert math.isclose(m[k], v, rel_tol=2e-3, abs_tol=2e-3), (k, m[k], v)


def test_dsic_forward_matches_reference_fixture():
    """Oracle restatement of DSIC.forward (ywz/DSIC/mynet6_plus.py:675-761) vs the unmodified reference at
    64x256, including two of the cost volumes the reference keeps as module attributes."""
    import mynet6_plus
    g, meta = load_npz("dsic"), load_json("dsic")
    net = mynet6_plus.DSIC(128, 192, 21, 32, 5)
    sd = synth.synth_state_dict(net, seed=0)
    x1, x2, _ = synth.stereo_pairs(1, meta["H"], meta["W"], seed=1234)
    taps = {}
    with torch.no_grad():
        out = O.dsic_forward(sd, x1, x2, taps=taps)
    assert_close(taps["cost1"], g["cost1"], 1e-4, floor=1e-6, what="cost volume 1")
    assert_close(out["x1_hat"], g["x1_hat"], 1e-5, what="x1_hat")
    assert_close(out["x2_hat"], g["x2_hat"], 1e-4, what="x2_hat")
    for k in ("y1", "z1", "z2"):
        assert_close(out["likelihoods"][k], g["lik_" + k], 1e-4, floor=1e-9, what="likelihood " + k)
    assert mismatch_fraction(out["likelihoods"]["y2"] > 0.5, T(g["lik_y2"]) > 0.5) < 1e-3
    m = synth.rd_metrics(out, x1, x2)
    for k, v in meta["metrics"].items():
        assert abs(m[k] - v) <= 2e-4 * abs(v), (k, m[k], v)


def test_dsic_dense_warp_and_conv3d_identities():
    """dense_warp with a one-hot cost volume is a pure shift; a Conv3d equals the block-banded 2-D convolution
    the CUDA path evaluates (hesic_b200/dsic.py: Conv3dAs2d)."""
    h1 = torch.randn(2, 5, 4, 40)
    cost = torch.zeros(2, 32, 4, 40)
    cost[:, 3] = 1.0
    out = O.dsic_dense_warp(h1, cost)
    assert torch.equal(out[..., :37], h1[..., 3:]) and float(out[..., 37:].abs().max()) == 0
    w3, b3 = torch.randn(7, 7, 5, 5, 5) * 0.05, torch.randn(7) * 0.1
    x = torch.randn(1, 7, 8, 6, 9)
    ref = torch.nn.functional.conv3d(x, w3, b3, padding=2)
    D = 8
    w2 = torch.zeros(7, D, 7, D, 5, 5)
    for dd in range(D):
        lo, hi = max(0, dd - 2), min(D, dd + 3)
        w2[:, dd, :, lo:hi] = w3[:, :, lo - dd + 2:hi - dd + 2]
    y2 = torch.nn.functional.conv2d(x.reshape(1, 7 * D, 6, 9), w2.reshape(7 * D, 7 * D, 5, 5), b3.repeat_interleave(D), padding=2)
    assert_close(y2.reshape(ref.shape), ref, 1e-5, what="conv3d as banded conv2d")


def test_homography_net_oracle_vs_reference_fixture():
    """ywz/mywork/model.py Net: oracle restatement against the delta the unmodified reference produced."""
    from hesic_b200 import compat
    compat.install()
    import model
    net = model.Net(patch_size=128).eval()
    gold_sd = load_json("homography_net")["state_dict_init"]
    sd0 = net.state_dict()
    assert set(sd0) == set(gold_sd) and all(list(sd0[k].shape) == gold_sd[k]["shape"] for k in gold_sd)
    sd = synth.synth_state_dict(net, seed=0)
    x1, x2, _ = synth.stereo_pairs(2, 128, 128, seed=55)
    with torch.no_grad():
        delta = O.homography_net_forward(sd, x1.mean(1, keepdim=True), x2.mean(1, keepdim=True))
    assert_close(delta, load_npz("homography_net")["delta"], 1e-4, what="homography delta")


def test_perspective_transform_maps_the_corners():
    """get_perspective_transform (kornia, un-vendored: parity unpinned) -- checked by its defining property:
    H maps each source corner onto its destination corner; h_adjust is the caller's frame rescale."""
    g = torch.Generator().manual_seed(5)
    src = torch.tensor([[[0., 0.], [127., 0.], [127., 127.], [0., 127.]]]).repeat(3, 1, 1)
    dst = src + (torch.rand(3, 4, 2, generator=g) - 0.5) * 16
    H = O.get_perspective_transform(src, dst).double()
    p = torch.cat([src.double(), torch.ones(3, 4, 1, dtype=torch.float64)], -1) @ H.transpose(1, 2)
    assert torch.allclose(p[..., :2] / p[..., 2:], dst.double(), atol=1e-3)
    Hs = O.h_adjust(512, 512, 256, 256, H.float())
    q = torch.cat([2 * src.double(), torch.ones(3, 4, 1, dtype=torch.float64)], -1) @ Hs.double().transpose(1, 2)
    assert torch.allclose(q[..., :2] / q[..., 2:], 2 * dst.double(), atol=5e-3)


def test_warp_perspective_agrees_with_the_convention_the_reference_computes_H_in():
    """kornia is un-vendored (warp parity unpinned), but the reference's own data pipeline fixes the convention:
    its H comes from cv2.findHomography on keypoints in PIXEL coordinates (compressai/datasets/utils.py:52-57) and
    is handed unchanged to kornia.warp_perspective (newnet1.py:746).  The oracle's warp must therefore agree with
    cv2.warpPerspective for the same H: up to cv2's 5-bit interpolation weights with align_corners=True, while the
    align_corners=False variant of kornia 0.4.x is off by half-pixel shifts (an order of magnitude more)."""
    cv2 = pytest.importorskip("cv2")
    x1, _, h = synth.stereo_pairs(2, 128, 160, seed=7)

    def against_cv2(align):
        w = O.warp_perspective(x1, h, (128, 160), align_corners=align)
        worst, mean = 0.0, 0.0
        for b in range(2):
            ref = cv2.warpPerspective(x1[b].permute(1, 2, 0).numpy(), h[b].numpy().astype(np.float64), (160, 128),
                                      flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
            d = np.abs(w[b].permute(1, 2, 0).numpy() - ref)
            worst, mean = max(worst, float(d.max())), max(mean, float(d.mean()))
        return worst, mean

    worst, mean = against_cv2(True)
    assert worst < 2e-2 and mean < 1.5e-3, (worst, mean)        # 1/32-pixel weight quantisation of cv2
    worst_f, mean_f = against_cv2(False)
    assert worst_f > 5 * worst and mean_f > 10 * mean, (worst_f, mean_f)
    # identity and a whole-pixel translation are exact resamplings
    assert (O.warp_perspective(x1[:1], torch.eye(3)[None], (128, 160)) - x1[:1]).abs().max() < 1e-4
    shift = torch.tensor([[[1., 0, 5], [0, 1, -3], [0, 0, 1]]])
    ref = torch.zeros_like(x1[:1])
    ref[:, :, :-3, 5:] = x1[:1, :, 3:, :-5]
    assert (O.warp_perspective(x1[:1], shift, (128, 160)) - ref).abs().max() < 1e-4


def test_dsic_independent_en_oracle_vs_reference_fixture():
    from hesic_b200 import compat
    compat.install()
    import mynet6_plus
    en = mynet6_plus.Independent_EN().eval()
    sd = synth.synth_state_dict(en, seed=0)
    x1, x2, _ = synth.stereo_pairs(1, 64, 64, seed=98)
    with torch.no_grad():
        o = O.dsic_independent_en_forward(sd, x1, x2)
    g = load_npz("dsic_independent_en")
    assert_close(o["x1_hat"], g["x1_hat"], 1e-5, what="DSIC EN x1")
    assert_close(o["x2_hat"], g["x2_hat"], 1e-5, what="DSIC EN x2")


@pytest.mark.parametrize("name", ["hsic_newnet1", "hsic_newnet9", "hsic_joint", "dsic", "independent_en"])
def test_default_state_matches_the_reference_constructors(name):
    """oracle/default_state.py rebuilds the constructor-fixed tensors (bounds, pedestals, bottleneck target / matrices,
    masks) from the key table alone; each one is SHA-1-equal to what the unmodified reference constructed, and the
    seeded weights derived from it equal the ones the fixtures were generated with.  bench.py's CPU arm relies on this
    to build its weights without instantiating a model of this repository."""
    import hashlib
    from hesic_b200 import synth
    from oracle import default_state as D
    meta = load_json(name)
    tab = meta["state_dict_init"]
    sd = D.initial_state_dict(tab)
    sha = lambda t: hashlib.sha1(t.detach().contiguous().numpy().tobytes()).hexdigest() if t.numel() else ""
    assert set(sd) == set(tab)
    for k, g in tab.items():
        assert tuple(sd[k].shape) == tuple(g["shape"]) and str(sd[k].dtype).replace("torch.", "") == g["dtype"], k
        if D.rule_governed(k) or sd[k].numel() == 0:
            assert sha(sd[k]) == g["sha1"], f"constructor value of {k} differs from the reference"
    if "synth_sha1" in meta:
        ss = synth.synth_state_dict(sd, seed=0)
        for k, h in meta["synth_sha1"].items():
            assert sha(ss[k]) == h, f"seeded value of {k} differs from the fixture's"
