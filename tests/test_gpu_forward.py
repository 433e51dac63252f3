"""End-to-end parity of HSIC.forward on the B200 against the oracle and the reference fixtures."""
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


def _model(modname):
    mod = __import__(modname)
    net = mod.HSIC(128, 192, 5).eval()
    sd = synth.synth_state_dict(net, seed=0)
    net.load_state_dict(sd)
    return net.to(DEV), sd


def _check_against(out, ref, x1, x2, sym_tol=2e-3, what=""):
    """Symbols may flip at exact x.5 ties (fp32 summation order); everything else is tight."""
    for k in ("y1_hat", "y2_hat"):
        if k in out and k in ref:
            assert mismatch_fraction(out[k], ref[k]) < sym_tol, (what, k)
    for k in ("x1_hat", "x2_hat"):
        a, b = out[k].cpu().double(), torch.as_tensor(ref[k]).double()
        rel = float((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt())
        assert rel < 5e-3, (what, k, rel)
    m = synth.rd_metrics({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in out.items() if k != "likelihoods"}
                         | {"likelihoods": {k: v.cpu() for k, v in out["likelihoods"].items()}}, x1, x2)
    return m


@pytest.mark.parametrize("name,modname", [("hsic_newnet1", "newnet1"), ("hsic_newnet9", "newnet9"),
                                          ("hsic_joint", "newnet1_joint")])
def test_forward_small_vs_reference_fixture(name, modname):
    net, sd = _model(modname)
    meta, gold = load_json(name), load_npz(name)
    x1, x2, h = synth.stereo_pairs(meta["B"], meta["H"], meta["W"], seed=1234)
    out = net(x1.to(DEV), x2.to(DEV), h.to(DEV))
    assert set(out) == ({"x1_hat", "x2_hat", "likelihoods"} | (set() if name == "hsic_newnet9" else {"y1_hat", "y2_hat"}))
    assert set(out["likelihoods"]) == {"y1", "y2", "z1", "z2"}
    ref = {k: (T(v.astype(np.float32)) if k.endswith("_hat") else T(v)) for k, v in gold.items()}
    m = _check_against(out, ref, x1, x2, what=name)
    for k, v in meta["metrics"].items():
        assert math.isclose(m[k], v, rel_tol=2e-3, abs_tol=2e-3), (k, m[k], v)
    # view-1 quantities do not depend on tie-flips of view 2: tight comparison
    assert_close(out["likelihoods"]["z1"], gold["lik_z1"], 1e-3, floor=1e-9, what="z1 likelihood")
    # fused bpp partial sums agree with the likelihood tensors they summarise
    sums = net.hesic_engine.log2_sums.cpu()
    for i, k in enumerate(("y1", "y2", "z1", "z2")):
        direct = float(torch.log2(out["likelihoods"][k].double()).sum())
        assert math.isclose(float(sums[i]), direct, rel_tol=1e-5, abs_tol=1e-3), k


@pytest.mark.parametrize("name,modname", [("hsic_newnet1", "newnet1"), ("hsic_joint", "newnet1_joint")])
def test_forward_full_size_metrics(name, modname):
    """512x512 (the BASELINE size): bpp / PSNR against the reference's stored scalars and the live oracle."""
    net, sd = _model(modname)
    meta = load_json(name)
    x1, x2, h = synth.stereo_pairs(1, 512, 512, seed=1234)
    out = net(x1.to(DEV), x2.to(DEV), h.to(DEV))
    cpu = {k: v.cpu() for k, v in out.items() if k != "likelihoods"}
    cpu["likelihoods"] = {k: v.cpu() for k, v in out["likelihoods"].items()}
    m = synth.rd_metrics(cpu, x1, x2)
    for k, v in meta["metrics_512"].items():
        assert math.isclose(m[k], v, rel_tol=2e-3, abs_tol=2e-3), (k, m[k], v)
    with torch.no_grad():
        ref = O.hsic_joint_forward(sd, x1, x2, h) if name == "hsic_joint" else O.hsic_forward(sd, x1, x2, h)
    _check_against(out, ref, x1, x2, what=name + " 512")
    assert out["x1_hat"].shape == (1, 3, 512, 512) and out["y1_hat"].shape == (1, 192, 32, 32)
    assert out["likelihoods"]["z1"].shape == (1, 128, 8, 8)


def test_forward_batch_is_per_sample_independent():
    """Pairs are independent end to end: a batch equals its samples run one by one."""
    net, _ = _model("newnet1")
    x1, x2, h = synth.stereo_pairs(3, 128, 128, seed=9)
    full = net(x1.to(DEV), x2.to(DEV), h.to(DEV))
    for i in range(3):
        one = net(x1[i:i + 1].to(DEV), x2[i:i + 1].to(DEV), h[i:i + 1].to(DEV))
        assert torch.equal(one["y1_hat"], full["y1_hat"][i:i + 1])
        assert torch.equal(one["x1_hat"], full["x1_hat"][i:i + 1])
        assert torch.equal(one["likelihoods"]["y2"], full["likelihoods"]["y2"][i:i + 1])


def test_forward_rejects_bad_arguments():
    net, _ = _model("newnet1")
    x = torch.rand(1, 3, 100, 128, device=DEV)
    with pytest.raises(ValueError):
        net(x, x, torch.eye(3, device=DEV)[None])
    x = torch.rand(2, 3, 128, 128, device=DEV)
    with pytest.raises(ValueError):
        net(x, x, torch.eye(3, device=DEV)[None])


def test_independent_en_vs_reference_fixture():
    import newnet1
    en = newnet1.Independent_EN().eval()
    sd = synth.synth_state_dict(en, seed=0)
    en.load_state_dict(sd)
    gold = load_npz("independent_en")
    x1, x2, h = synth.stereo_pairs(1, 64, 64, seed=99)
    out = en.to(DEV)(x1.to(DEV), x2.to(DEV), h.to(DEV))
    assert_close(out["x1_hat"], gold["x1_hat"], 1e-4, what="EN x1")
    assert_close(out["x2_hat"], gold["x2_hat"], 1e-4, what="EN x2")


def test_independent_en_full_size_vs_oracle():
    """512x512 (the BASELINE size), batch 2: the fused enhancement engine against the live oracle."""
    import newnet1
    en = newnet1.Independent_EN().eval()
    sd = synth.synth_state_dict(en, seed=0)
    en.load_state_dict(sd)
    x1, x2, h = synth.stereo_pairs(2, 512, 512, seed=77)
    with torch.no_grad():
        ref = O.independent_en_forward(sd, x1, x2, h)
    out = en.to(DEV)(x1.to(DEV), x2.to(DEV), h.to(DEV))
    from hesic_b200 import _capi as C
    C.check(C.lib.hesic_tc_status())
    assert_close(out["x1_hat"], ref["x1_hat"], 1e-4, what="EN x1 512")
    assert_close(out["x2_hat"], ref["x2_hat"], 1e-4, what="EN x2 512")
    # pairs are independent
    one = en(x1[1:].to(DEV), x2[1:].to(DEV), h[1:].to(DEV))
    assert torch.equal(one["x1_hat"], out["x1_hat"][1:])
