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
