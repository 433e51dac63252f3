"""Parity at the CONFIGURED sizes (BASELINE.json configs 2-5), not only at fixture size:

* HESIC and HESIC+ at batch 16 x 512x512 (configs 2 / 3) against the live oracle on the same seeded inputs;
* the N > 1 reduction of config 4 (batch sharded over ranks, partial sums added) against the oracle's metrics of the
  whole batch -- the shards run one after another on this GPU, the reduction is hesic_b200.sharding's;
* DSIC at 512x512 (config 5's image size) against the live oracle.

Bars: symbols may flip where fp32 summation order moves a latent across x.5 (measured flip fraction printed with
every run; the bar is 10x tighter than round 1's 2e-3); everything a flip cannot reach is held at 1e-4.
Measured values of the last GPU run are appended to gpurun_out/parity_stats.jsonl when that directory exists."""
import json
import math
import os

import pytest
import torch

from hesic_b200 import compat, sharding, synth
from hesic_b200 import functional as F
from oracle import hesic_oracle as O
from tests.helpers import assert_close, mismatch_fraction

compat.install()
pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _note(name, **kv):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_stats.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **kv}) + "\n")
    print(name, kv)


def _model(modname, cls="HSIC", args=(128, 192, 5)):
    mod = __import__(modname)
    net = getattr(mod, cls)(*args).eval()
    sd = synth.synth_state_dict(net, seed=0)
    net.load_state_dict(sd)
    return net.to(DEV), sd


def _cpu(out):
    r = {k: v.cpu() for k, v in out.items() if k != "likelihoods"}
    r["likelihoods"] = {k: v.cpu() for k, v in out["likelihoods"].items()}
    return r


def _rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def _oracle_batched(fn, sd, x1, x2, h):
    """The oracle pair by pair (pairs are independent; batch 1 is also the CPU's fastest form)."""
    outs = []
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        for i in range(x1.shape[0]):
            outs.append(fn(sd, x1[i:i + 1], x2[i:i + 1], h[i:i + 1]) if h is not None else fn(sd, x1[i:i + 1], x2[i:i + 1]))
    cat = lambda key: torch.cat([o[key] for o in outs]) if key in outs[0] else None
    res = {k: cat(k) for k in ("x1_hat", "x2_hat", "y1_hat", "y2_hat") if k in outs[0]}
    res["likelihoods"] = {k: torch.cat([o["likelihoods"][k] for o in outs]) for k in outs[0]["likelihoods"]}
    return res


@pytest.mark.parametrize("name,modname,fwd", [("hesic", "newnet1", "hsic_forward"), ("hesic_plus", "newnet1_joint", "hsic_joint_forward")])
def test_configured_batch_16_at_512_vs_oracle(name, modname, fwd):
    """BASELINE configs 2 and 3: batch 16 x 512x512, seed 1234 -- the benchmarked computation itself."""
    net, sd = _model(modname)
    x1, x2, h = synth.stereo_pairs(16, 512, 512, seed=1234)
    out = _cpu(net(x1.to(DEV), x2.to(DEV), h.to(DEV)))
    sums = net.hesic_engine.log2_sums.cpu()
    ref = _oracle_batched(getattr(O, fwd), sd, x1, x2, h)
    flips = {k: mismatch_fraction(out[k], ref[k]) for k in ("y1_hat", "y2_hat")}
    l2 = {k: _rel_l2(out[k], ref[k]) for k in ("x1_hat", "x2_hat")}
    m, r = synth.rd_metrics(out, x1, x2), synth.rd_metrics(ref, x1, x2)
    rel = {k: abs(m[k] - r[k]) / abs(r[k]) for k in ("bpp", "bpp1", "bpp2")}
    dps = {k: abs(m[k] - r[k]) for k in ("psnr1", "psnr2")}
    # element-wise where a flip cannot reach: the z1 likelihoods of pairs without a z1 flip (a flipped z1 symbol moves its
    # likelihood by far more than 1e-3), and the y1 likelihoods of the symbols that did not flip in those pairs
    from tests.helpers import close_stats
    zrel = [close_stats(out["likelihoods"]["z1"][i], ref["likelihoods"]["z1"][i], floor=1e-9) for i in range(16)]
    zclean = [i for i in range(16) if zrel[i] < 1e-3]
    same = out["y1_hat"] == ref["y1_hat"]
    lik_y1 = 0.0
    for i in zclean:
        a_, b_ = out["likelihoods"]["y1"][i][same[i]], ref["likelihoods"]["y1"][i][same[i]]
        lik_y1 = max(lik_y1, close_stats(a_, b_, floor=1e-6))
    _note(f"b16_512_{name}", flips=flips, x_hat_rel_l2=l2, bpp_rel=rel, psnr_abs=dps, pairs_without_z1_flip=len(zclean),
          z1_lik_rel=max(zrel[i] for i in zclean) if zclean else None, y1_lik_rel_unflipped=lik_y1, bpp=m["bpp"], bpp_oracle=r["bpp"])
    assert flips["y1_hat"] < 2e-4 and flips["y2_hat"] < 2e-4, flips
    assert l2["x1_hat"] < 1e-3 and l2["x2_hat"] < 1e-3, l2
    assert max(rel.values()) < (2e-5 if name == "hesic" else 1e-4), rel
    assert max(dps.values()) < 5e-4, dps
    # (measured on the B200, seed 1234: flips 2.8e-5 / 4.3e-5, x_hat L2 3.5e-4 / 2.8e-4, bpp 1e-7 (HESIC) / 1.4e-5 (HESIC+),
    #  PSNR 4e-5 dB; 11 / 7 of the 16 pairs without a z1 flip, their z1 likelihoods within 5.8e-6 (HESIC))
    assert len(zclean) >= 4
    if name == "hesic":
        assert max(zrel[i] for i in zclean) < 1e-4
        # likelihood of an unflipped symbol: the mixture parameters carry the convs' 1e-4-of-rms error, which the Gaussian
        # tail amplifies by |y - mu| / sigma^2 (up to ~50x at sigma = 0.11): held at 2e-2 of max(p, 1e-6)
        assert lik_y1 < 2e-2, lik_y1
    # (HESIC+: a flipped symbol changes the context model's input of its 12 causal neighbours, so per-symbol likelihoods
    #  are only comparable in aggregate -- the bpp bars above)
    # the fused partial sums are the sums of the returned likelihoods
    for i, k in enumerate(("y1", "y2", "z1", "z2")):
        direct = float(torch.log2(out["likelihoods"][k].double()).sum())
        assert math.isclose(float(sums[i]), direct, rel_tol=1e-6), k


def test_sharded_reduction_matches_the_oracle_of_the_whole_batch():
    """BASELINE config 4 in miniature: a global batch of 8 pairs cut into 4 rank shards (hesic_b200.sharding), each shard's
    six partial sums produced on the device exactly as bench.py does, added, and turned into bpp / PSNR -- against the
    oracle's metrics of the whole batch computed the reference's way (ywz/mywork/test3real.py:90-124)."""
    net, sd = _model("newnet1")
    n, world = 8, 4
    x1, x2, h = synth.stereo_pairs(n, 512, 512, seed=77)
    total = torch.zeros(6, dtype=torch.float64)
    for rank in range(world):
        b, e = sharding.shard_range(n, rank, world)
        a1, a2, ah = x1[b:e].to(DEV), x2[b:e].to(DEV), h[b:e].to(DEV)
        out = net(a1, a2, ah)
        partial = torch.zeros(6, device=DEV, dtype=torch.float64)
        partial[:4].copy_(net.hesic_engine.log2_sums)
        partial[4:].copy_(net.hesic_engine.sse_sums)            # accumulated by the epilogues that store x1_hat / x2_hat
        check = torch.zeros(2, device=DEV, dtype=torch.float64)
        F.sum_squared_error(out["x1_hat"], a1, check[0:1])
        F.sum_squared_error(out["x2_hat"], a2, check[1:2])
        assert torch.allclose(partial[4:], check, rtol=1e-8, atol=0)
        total += sharding.reduce_partials(partial).cpu()        # world size 1 here: the sum over ranks is the loop
    got = sharding.metrics_from_partials(total, n, 512, 512)
    ref = _oracle_batched(O.hsic_forward, sd, x1, x2, h)
    r = synth.rd_metrics(ref, x1, x2)
    _note("sharded_reduction", got={k: got[k] for k in ("bpp", "bpp1", "bpp2", "psnr1", "psnr2")}, oracle=r)
    for k in ("bpp", "bpp1", "bpp2"):
        assert abs(got[k] - r[k]) <= 2e-4 * r[k], (k, got[k], r[k])
    for k in ("psnr1", "psnr2"):
        assert abs(got[k] - r[k]) <= 2e-3, (k, got[k], r[k])


def test_dsic_at_512_vs_oracle():
    """BASELINE config 5's image size (512x512), batch 1: the fused DSIC engine against the live oracle."""
    from hesic_b200 import _capi as C
    net, sd = _model("mynet6_plus", "DSIC", (128, 192, 21, 32, 5))
    x1, x2, _ = synth.stereo_pairs(1, 512, 512, seed=1234)
    torch.set_num_threads(os.cpu_count() or 1)
    taps = {}
    with torch.no_grad():
        ref = O.dsic_forward(sd, x1, x2, taps=taps)
    out = _cpu(net(x1.to(DEV), x2.to(DEV)))
    C.check(C.lib.hesic_tc_status())
    sse = net.hesic_engine.sse_sums.cpu()        # from the epilogues of the two RGB heads (128 and 256 input channels)
    for i, (k, x) in enumerate((("x1_hat", x1), ("x2_hat", x2))):
        assert math.isclose(float(sse[i]), float(((out[k] - x).double() ** 2).sum()), rel_tol=1e-8), k
    m, r = synth.rd_metrics(out, x1, x2), synth.rd_metrics(ref, x1, x2)
    l2 = {k: _rel_l2(out[k], ref[k]) for k in ("x1_hat", "x2_hat")}
    rel = {k: abs(m[k] - r[k]) / abs(r[k]) for k in ("bpp", "bpp1", "bpp2")}
    dps = {k: abs(m[k] - r[k]) for k in ("psnr1", "psnr2")}
    lik1 = mismatch_fraction(out["likelihoods"]["y1"] > 0.5, ref["likelihoods"]["y1"] > 0.5)
    _note("dsic_512", x_hat_rel_l2=l2, bpp_rel=rel, psnr_abs=dps, y1_lik_side_mismatch=lik1, bpp=m["bpp"], bpp_oracle=r["bpp"])
    # z1 likelihoods element-wise, flip-aware as in the HESIC test above: a hyper-latent that sits within the convs' error of a
    # rounding boundary may round the other way (its likelihood then moves by percents); r02 build: no flip at this seed,
    # r03 build (first-layer kernel with the x^2-reconstructed GDN pass): one of 8192
    za, zb = out["likelihoods"]["z1"], ref["likelihoods"]["z1"]
    zrel = (za - zb).abs() / zb.abs().clamp_min(1e-9)
    zflip = zrel > 1e-3
    _note("dsic_512_z1", flipped=int(zflip.sum()), of=zflip.numel(), rel_unflipped=float(zrel[~zflip].max()))
    assert int(zflip.sum()) <= 4, int(zflip.sum())
    assert float(zrel[~zflip].max()) < 1e-4
    # view 1 is plain HESIC analysis / synthesis: tight; view 2 passes six softmax-normalised cost volumes whose logits
    # carry the conv error times their magnitude (see test_gpu_dsic.py::test_cost_volume_stages_vs_oracle)
    assert l2["x1_hat"] < 5e-4 and l2["x2_hat"] < 5e-3, l2
    assert rel["bpp1"] < 2e-4 and rel["bpp"] < 1e-3 and rel["bpp2"] < 1e-3, rel
    assert dps["psnr1"] < 2e-3 and dps["psnr2"] < 2e-2, dps


def test_two_models_interleaved_do_not_share_buffers():
    """Two model instances (HESIC and HESIC+) called alternately, each on its own stream, return exactly what they
    return when run alone: the engines own their intermediates per instance (r01 kept them in one module-global list
    that every forward cleared)."""
    a, _ = _model("newnet1")
    b, _ = _model("newnet1_joint")
    x1, x2, h = (t.to(DEV) for t in synth.stereo_pairs(2, 256, 256, seed=5))
    y1, y2, g = (t.to(DEV) for t in synth.stereo_pairs(2, 256, 256, seed=6))
    ra, rb = a(x1, x2, h), b(y1, y2, g)
    torch.cuda.synchronize()
    keep = lambda o: {k: v.clone() for k, v in o.items() if k != "likelihoods"} | {"lik_" + k: v.clone() for k, v in o["likelihoods"].items()}
    ra, rb = keep(ra), keep(rb)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for _ in range(3):
        with torch.cuda.stream(sa):
            oa = a(x1, x2, h)
        with torch.cuda.stream(sb):
            ob = b(y1, y2, g)
        outs.append((oa, ob))
    torch.cuda.synchronize()
    for oa, ob in outs:
        oa, ob = keep(oa), keep(ob)
        for k in ra:
            assert torch.equal(oa[k], ra[k]), ("model A", k)
        for k in rb:
            assert torch.equal(ob[k], rb[k]), ("model B", k)


def test_cuda_graph_replay_equals_eager():
    """engine.capture(): one forward frozen into a CUDA graph; a replay on new inputs returns bit-identical tensors to
    the eager call and launches no library kernel from Python."""
    from hesic_b200 import _capi as C
    net, _ = _model("newnet1")
    x1, x2, h = (t.to(DEV) for t in synth.stereo_pairs(2, 256, 256, seed=8))
    y1, y2, g = (t.to(DEV) for t in synth.stereo_pairs(2, 256, 256, seed=9))
    cap = net.hesic_engine.capture(x1, x2, h)
    assert cap.n_launches > 40
    eager = net(y1, y2, g)
    eager = {k: v.clone() for k, v in eager.items() if k != "likelihoods"} | {k: v.clone() for k, v in eager["likelihoods"].items()}
    sums = net.hesic_engine.log2_sums.clone()
    sse = net.hesic_engine.sse_sums.clone()
    C.lib.hesic_launch_count(1)
    out = cap.replay(y1, y2, g)
    torch.cuda.synchronize()
    assert C.lib.hesic_launch_count(0) == 0
    for k in ("x1_hat", "x2_hat", "y1_hat", "y2_hat"):
        assert torch.equal(out[k], eager[k]), k
    for k in ("y1", "y2", "z1", "z2"):
        assert torch.equal(out["likelihoods"][k], eager[k]), k
    assert torch.allclose(cap.log2_sums, sums, rtol=1e-12)
    assert torch.allclose(cap.sse_sums, sse, rtol=1e-12)
    # and again on the first inputs: the graph is reusable, the engine still runs eagerly afterwards
    out = cap.replay(x1, x2, h)
    again = net(x1, x2, h)
    torch.cuda.synchronize()
    assert torch.equal(out["x2_hat"], again["x2_hat"])
    with pytest.raises(ValueError):
        cap.replay(x1[:1], x2[:1], h[:1])


def test_invalidate_after_data_surgery():
    """Weight surgery through ``.data`` does not bump the parameter version the plan caches key on;
    hesic_b200.invalidate(model) forces the re-pack."""
    import hesic_b200
    net, sd = _model("newnet1")
    x1, x2, h = (t.to(DEV) for t in synth.stereo_pairs(1, 128, 128, seed=3))
    a = net(x1, x2, h)["x1_hat"].clone()
    net.decoder1.g_s_conv4.bias.data.add_(0.25)
    hesic_b200.invalidate(net)
    b = net(x1, x2, h)["x1_hat"]
    assert float((b - a).abs().min()) > 0.2
    with torch.no_grad():
        ref = O.hsic_forward({k: v.cpu() for k, v in net.state_dict().items()}, x1.cpu(), x2.cpu(), h.cpu())
    assert_close(b, ref["x1_hat"], 1e-4, what="x1_hat after bias surgery")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_a_device_that_is_not_current():
    """A model moved to cuda:1 while cuda:0 is the current device: operands are re-packed on cuda:1 and every launch
    goes to cuda:1's stream."""
    net, sd = _model("newnet1")
    x1, x2, h = synth.stereo_pairs(1, 128, 128, seed=3)
    a = net(x1.to(DEV), x2.to(DEV), h.to(DEV))["x1_hat"].cpu()
    net = net.to("cuda:1")
    assert torch.cuda.current_device() == 0
    b = net(x1.to("cuda:1"), x2.to("cuda:1"), h.to("cuda:1"))["x1_hat"]
    assert b.device.index == 1 and torch.equal(b.cpu(), a)
    y = net.encoder1.g_a_conv1(x1.to("cuda:1"))
    assert y.device.index == 1
