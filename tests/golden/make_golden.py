"""Generate tests/golden/*.npz|json by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports /root/reference through oracle/ref_harness.py (reference Python code +
its own C++ coder compiled into oracle/_ref; kornia stood in by the oracle's
warp restatement -- warp parity is therefore unpinned, see oracle/hesic_oracle.py).
Inputs and weights come from hesic_b200/synth.py (numpy PCG64), so tests can
regenerate them bit-identically anywhere.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hesic_b200 import synth  # noqa: E402
from oracle import ref_harness  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def sd_table(sd):
    tab = {}
    for k, v in sd.items():
        tab[k] = {"shape": list(v.shape), "dtype": str(v.dtype).replace("torch.", ""),
                  "sha1": hashlib.sha1(v.detach().contiguous().numpy().tobytes()).hexdigest() if v.numel() else ""}
    return tab


def npf(t):
    return t.detach().cpu().numpy()


def model_golden(modname, ctor, fname, B, H, W, with_yhat=True):
    m = ref_harness.load(modname)
    net = ctor(m).eval()
    init_tab = sd_table(net.state_dict())
    sd = synth.synth_state_dict(net, seed=0)
    net.load_state_dict(sd)
    x1, x2, h = synth.stereo_pairs(B, H, W, seed=1234)
    with torch.no_grad():
        out = net(x1, x2, h)
    arrs = {"x1_hat": npf(out["x1_hat"]), "x2_hat": npf(out["x2_hat"])}
    if with_yhat:
        arrs["y1_hat"] = npf(out["y1_hat"]).astype(np.int16)
        arrs["y2_hat"] = npf(out["y2_hat"]).astype(np.int16)
    for k, v in out["likelihoods"].items():
        arrs["lik_" + k] = npf(v)
    np.savez_compressed(os.path.join(OUT, fname + ".npz"), **arrs)
    meta = {"module": modname, "B": B, "H": H, "W": W, "metrics": synth.rd_metrics(out, x1, x2),
            "state_dict_init": init_tab, "synth_sha1": {k: v["sha1"] for k, v in sd_table(sd).items()}}
    # full-size single pair: scalars only
    x1, x2, h = synth.stereo_pairs(1, 512, 512, seed=1234)
    with torch.no_grad():
        out = net(x1, x2, h)
    meta["metrics_512"] = synth.rd_metrics(out, x1, x2)
    meta["sums_512"] = {"x1_hat": float(out["x1_hat"].double().sum()), "x2_hat": float(out["x2_hat"].double().sum())}
    json.dump(meta, open(os.path.join(OUT, fname + ".json"), "w"), indent=1, sort_keys=True)
    print(fname, meta["metrics"], meta["metrics_512"])
    return net


def operators_golden():
    ref_harness.install()
    from compressai.entropy_models import EntropyBottleneck, GaussianConditional, GaussianMixtureConditional
    from compressai.layers import GDN, MaskedConv2d
    from compressai.models.priors import get_scale_table
    import compressai.ans as ans
    import compressai._CXX as cxx

    g = np.random.default_rng(7)
    T = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    A = {}
    # --- GDN / IGDN (compressai/layers/gdn.py) on non-trivial params
    for inv in (False, True):
        gd = GDN(16, inverse=inv)
        sd = synth.synth_state_dict({"g_a_gdn1." + k: v for k, v in gd.state_dict().items()}, seed=3)
        gd.load_state_dict({k.split(".", 1)[1]: v for k, v in sd.items()})
        x = T(g.standard_normal((2, 16, 9, 7)) * 2)
        with torch.no_grad():
            A[f"gdn{int(inv)}_x"] = npf(x)
            A[f"gdn{int(inv)}_beta"] = npf(gd.beta)
            A[f"gdn{int(inv)}_gamma"] = npf(gd.gamma)
            A[f"gdn{int(inv)}_y"] = npf(gd(x))
    # --- EntropyBottleneck forward / update / compress / decompress
    eb = EntropyBottleneck(8).eval()
    sd = synth.synth_state_dict({"entropy_bottleneck1." + k: v for k, v in eb.state_dict().items()}, seed=5)
    eb.load_state_dict({k.split(".", 1)[1]: v for k, v in sd.items()})
    z = T(g.standard_normal((2, 8, 5, 6)) * 4)
    z[0, 0, 0, 0] = 40.0   # out-of-table symbols -> bypass coding
    z[1, 3, 2, 1] = -37.0
    with torch.no_grad():
        z_hat, z_lik = eb(z)
        eb.update()
        strings = eb.compress(z)
        # reference quirk: decompress() rejects batch > 1 (entropy_models.py:221-224) -> one item at a time
        z_dec = torch.cat([eb.decompress([s_], z.shape[-2:]) for s_ in strings], dim=0)
    for k, v in eb.state_dict().items():
        A["eb_sd_" + k] = npf(v)
    assert torch.equal(z_dec, z_hat)
    A.update(eb_z=npf(z), eb_z_hat=npf(z_hat), eb_lik=npf(z_lik), eb_z_dec=npf(z_dec),
             eb_cdf=npf(eb._quantized_cdf), eb_cdf_length=npf(eb._cdf_length), eb_offset=npf(eb._offset))
    for i, s in enumerate(strings):
        A[f"eb_string{i}"] = np.frombuffer(s, dtype=np.uint8)
    # --- GaussianMixtureConditional forward
    K, M = 5, 12
    gm = GaussianMixtureConditional(K=K).eval()
    y = T(g.standard_normal((2, M, 6, 5)) * 3)
    sc = T(np.abs(g.standard_normal((2, M * K, 6, 5))) * 1.5)
    mu = T(g.standard_normal((2, M * K, 6, 5)) * 2)
    w = torch.softmax(T(g.standard_normal((2, K, M, 1, 1))), dim=1).reshape(2, K * M, 1, 1)
    with torch.no_grad():
        y_hat, lik = gm(y, sc, mu, w)
    A.update(gmm_y=npf(y), gmm_scales=npf(sc), gmm_means=npf(mu), gmm_weights=npf(w), gmm_y_hat=npf(y_hat), gmm_lik=npf(lik))
    # --- GaussianConditional forward / build_indexes / update / compress
    table = get_scale_table()
    gc = GaussianConditional(None).eval()
    sc1 = T(np.abs(g.standard_normal((2, M, 6, 5))) * 2)
    mu1 = T(g.standard_normal((2, M, 6, 5)))
    with torch.no_grad():
        yh, lk = gc(y, sc1, means=mu1)
        yh0, lk0 = gc(y, sc1)
        gc.update_scale_table(table)
        idx = gc.build_indexes(sc1)
        gstr = gc.compress(y, idx, means=mu1)
        gdec = gc.decompress(gstr, idx, means=mu1)
        assert torch.equal(gdec, yh)
    A.update(gc_scales=npf(sc1), gc_means=npf(mu1), gc_y_hat=npf(yh), gc_lik=npf(lk), gc_y_hat0=npf(yh0), gc_lik0=npf(lk0),
             gc_indexes=npf(idx), gc_table=npf(table), gc_cdf=npf(gc._quantized_cdf), gc_cdf_length=npf(gc._cdf_length),
             gc_offset=npf(gc._offset), gc_dec=npf(gdec))
    for i, s in enumerate(gstr):
        A[f"gc_string{i}"] = np.frombuffer(s, dtype=np.uint8)
    # --- MaskedConv2d (layers.py:21-45)
    mc = MaskedConv2d(6, 8, kernel_size=5, padding=2, stride=1)
    xin = T(g.standard_normal((1, 6, 7, 7)))
    with torch.no_grad():
        A.update(mc_w=npf(mc.weight.clone()), mc_b=npf(mc.bias), mc_x=npf(xin), mc_y=npf(mc(xin)), mc_mask=npf(mc.mask))
    # --- pmf_to_quantized_cdf known answers (ops.cpp:24-81), incl. the "steal" loop
    pmfs = [g.dirichlet(np.ones(n)).astype(np.float32) for n in (3, 17, 64)]
    sharp = np.full(40, 1e-7, dtype=np.float32)
    sharp[20] = 1.0
    pmfs.append(sharp)
    for i, p in enumerate(pmfs):
        A[f"pmf{i}"] = p
        A[f"pmf{i}_cdf"] = np.asarray(cxx.pmf_to_quantized_cdf(p.tolist(), 16), dtype=np.uint32)
    # --- raw rANS known answers (rans_interface.cpp), incl. bypass symbols
    cdfs = np.zeros((3, 66), dtype=np.int32)
    sizes = []
    for i, p in enumerate(pmfs[:3]):
        c = A[f"pmf{i}_cdf"]
        cdfs[i, :c.size] = c
        sizes.append(c.size)
    offsets = [-1, -8, -32]
    idxs = g.integers(0, 3, 500).astype(np.int32)
    syms = np.array([g.integers(offsets[i] - 3, offsets[i] + sizes[i] + 2) for i in idxs], dtype=np.int32)
    syms[10] = 1000
    syms[11] = -70000
    s = ans.RansEncoder().encode_with_indexes(syms.tolist(), idxs.tolist(), cdfs.tolist(), sizes, offsets)
    dec = ans.RansDecoder().decode_with_indexes(s, idxs.tolist(), cdfs.tolist(), sizes, offsets)
    assert dec == syms.tolist()
    A.update(rans_symbols=syms, rans_indexes=idxs, rans_cdfs=cdfs, rans_sizes=np.array(sizes, dtype=np.int32),
             rans_offsets=np.array(offsets, dtype=np.int32), rans_stream=np.frombuffer(s, dtype=np.uint8))
    np.savez_compressed(os.path.join(OUT, "operators.npz"), **A)
    print("operators.npz", len(A), "arrays")


def dsic_golden():
    """DSIC (ywz/DSIC/mynet6_plus.py), BASELINE config 5: whole forward at 64x256 (the reference's dense_warp needs
    W/8 >= C = 32 disparities) plus cost volumes the reference keeps as module attributes, and scalar metrics
    of one 256x256 pair."""
    m = ref_harness.load("mynet6_plus")
    net = m.DSIC(128, 192, 21, 32, 5).eval()
    init_tab = sd_table(net.state_dict())
    sd = synth.synth_state_dict(net, seed=0)
    net.load_state_dict(sd)
    x1, x2, _ = synth.stereo_pairs(1, 64, 256, seed=1234)
    with torch.no_grad():
        out = net(x1, x2)
    arrs = {"x1_hat": npf(out["x1_hat"]), "x2_hat": npf(out["x2_hat"]), "cost1": npf(net._cost_volume1.cost),
            "cost3": npf(net._cost_volume3.cost)}
    for k, v in out["likelihoods"].items():
        arrs["lik_" + k] = npf(v)
    np.savez_compressed(os.path.join(OUT, "dsic.npz"), **arrs)
    meta = {"module": "mynet6_plus", "B": 1, "H": 64, "W": 256, "metrics": synth.rd_metrics(out, x1, x2),
            "state_dict_init": init_tab}
    x1, x2, _ = synth.stereo_pairs(1, 256, 256, seed=1234)
    with torch.no_grad():
        out = net(x1, x2)
    meta["metrics_256"] = synth.rd_metrics(out, x1, x2)
    meta["sums_256"] = {"x1_hat": float(out["x1_hat"].double().sum()), "x2_hat": float(out["x2_hat"].double().sum())}
    json.dump(meta, open(os.path.join(OUT, "dsic.json"), "w"), indent=1, sort_keys=True)
    print("dsic", meta["metrics"], meta["metrics_256"])


def dsic_en_golden():
    """mynet6_plus.Independent_EN (the enhancement stage of DSIC_plus, mynet6_plus.py:57-100,1352-1370) at 64x64."""
    m = ref_harness.load("mynet6_plus")
    en = m.Independent_EN().eval()
    en.load_state_dict(synth.synth_state_dict(en, seed=0))
    x1, x2, _ = synth.stereo_pairs(1, 64, 64, seed=98)
    with torch.no_grad():
        o = en(x1, x2)
    np.savez_compressed(os.path.join(OUT, "dsic_independent_en.npz"), x1_hat=npf(o["x1_hat"]), x2_hat=npf(o["x2_hat"]))
    print("dsic en", float(o["x1_hat"].abs().mean()))


def homography_golden():
    """ywz/mywork/model.py Net (the front-end that produces h_matrix) on two 128x128 gray patches."""
    m = ref_harness.load("model")
    net = m.Net(patch_size=128).eval()
    tab = sd_table(net.state_dict())
    net.load_state_dict(synth.synth_state_dict(net, seed=0))
    x1, x2, _ = synth.stereo_pairs(2, 128, 128, seed=55)
    a, b = x1.mean(1, keepdim=True), x2.mean(1, keepdim=True)
    with torch.no_grad():
        delta = net(a, b)
    np.savez_compressed(os.path.join(OUT, "homography_net.npz"), delta=npf(delta))
    json.dump({"state_dict_init": tab}, open(os.path.join(OUT, "homography_net.json"), "w"), indent=1, sort_keys=True)
    print("homography", delta.abs().mean().item())


if __name__ == "__main__":
    if "--dsic-en-only" in sys.argv:
        dsic_en_golden()
        sys.exit(0)
    if "--homography-only" in sys.argv:
        homography_golden()
        sys.exit(0)
    if "--dsic-only" in sys.argv:
        dsic_golden()
        sys.exit(0)
    operators_golden()
    model_golden("newnet1", lambda m: m.HSIC(128, 192, 5), "hsic_newnet1", 2, 128, 128)
    model_golden("newnet9", lambda m: m.HSIC(128, 192, 5), "hsic_newnet9", 2, 128, 128, with_yhat=False)
    model_golden("newnet1_joint", lambda m: m.HSIC(128, 192, 5), "hsic_joint", 2, 128, 128)
    # Independent_EN ("next")
    m = ref_harness.load("newnet1")
    en = m.Independent_EN().eval()
    tab = sd_table(en.state_dict())
    sd = synth.synth_state_dict(en, seed=0)
    en.load_state_dict(sd)
    x1, x2, h = synth.stereo_pairs(1, 64, 64, seed=99)
    with torch.no_grad():
        o = en(x1, x2, h)
    np.savez_compressed(os.path.join(OUT, "independent_en.npz"), x1_hat=npf(o["x1_hat"]), x2_hat=npf(o["x2_hat"]))
    json.dump({"state_dict_init": tab}, open(os.path.join(OUT, "independent_en.json"), "w"), indent=1, sort_keys=True)
    dsic_golden()
    homography_golden()
    dsic_en_golden()
    print("done")
