"""The reference's driver scripts load UNCHANGED against the drop-in tree, with no stub supplied by this test:
``hesic_b200.install()`` alone must make every module-level import of ywz/mywork/test3real.py, test3_real.py,
test3_joint_real.py and codec-test/test2_codec.py resolve (compressai.*, newnet*, model, kornia, range_coder,
pytorch_msssim, matplotlib, imageio), the model classes they name must be this package's, and
``compressai.datasets.ImageFolder`` must be the reference's own loader (delegated, not re-implemented).
Needs the reference checkout (build container); skipped where /root/reference does not exist (the GPU box)."""
import os
import runpy
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DRIVERS = ["ywz/mywork/test3real.py", "ywz/mywork/test3_real.py", "ywz/mywork/test3_joint_real.py",
           "ywz/mywork/codec-test/test2_codec.py"]

PROBE = r"""
import os, runpy, sys, json
sys.path.insert(0, {root!r})
os.environ["HESIC_REFERENCE_ROOT"] = {ref!r}
import hesic_b200
hesic_b200.install()
ns = runpy.run_path({driver!r}, run_name="hesic_b200_probe")
out = {{}}
for name in ("HSIC", "Independent_EN", "ImageFolder", "GDN", "CompressionModel", "Net", "RateDistortionLoss", "ms_ssim", "kornia"):
    v = ns.get(name)
    out[name] = None if v is None else (getattr(v, "__module__", None) or getattr(v, "__name__", ""))
out["ImageFolder_file"] = sys.modules[ns["ImageFolder"].__module__].__file__
out["has_main"] = callable(ns.get("main")) and callable(ns.get("test_epoch"))
# the driver's own argument parser and model construction (main() up to the data loader), as main() does it
args = ns["parse_args"](["-d", "/nonexistent", "--cuda", "-1", "--patch-size", "512", "512"])
out["patch"] = list(args.patch_size)
net = ns["HSIC"](N=128, M=192, K=5)
out["n_state"] = len(net.state_dict())
out["hsic_class_module"] = type(net).__module__
try:
    ns["ImageFolder"]("/nonexistent", split="test", patch_size=args.patch_size)
    out["imagefolder_error"] = None
except Exception as e:
    out["imagefolder_error"] = type(e).__name__ + ": " + str(e)
print("PROBE=" + json.dumps(out))
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
@pytest.mark.parametrize("driver", DRIVERS)
def test_driver_script_loads_unchanged(driver):
    import json
    code = PROBE.format(root=ROOT, ref=REF, driver=os.path.join(REF, driver))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("PROBE=")][-1]
    d = json.loads(line[len("PROBE="):])
    assert d["HSIC"].startswith("hesic_b200.") and d["hsic_class_module"].startswith("hesic_b200."), d
    if "codec" not in driver:
        assert d["Independent_EN"].startswith("hesic_b200."), d
    assert d["GDN"].startswith("compressai.layers") and d["has_main"]
    # the data loader is the REFERENCE's class, loaded from the reference checkout
    assert d["ImageFolder"] == "compressai.datasets._reference_utils" and d["ImageFolder_file"].startswith(REF), d
    assert d["imagefolder_error"] == 'RuntimeError: Invalid directory "/nonexistent"', d       # utils.py:91-92 of the reference
    assert d["n_state"] in (222, 230) and d["patch"] == [512, 512]


def test_imagefolder_without_a_checkout_says_what_to_do():
    code = ("import sys; sys.path.insert(0, %r); import hesic_b200; hesic_b200.install()\n"
            "from compressai.datasets import ImageFolder\n"
            "try:\n    ImageFolder('/tmp')\nexcept RuntimeError as e:\n    print('OK' if 'HESIC_REFERENCE_ROOT' in str(e) else e)\n" % ROOT)
    env = {k: v for k, v in os.environ.items() if k != "HESIC_REFERENCE_ROOT"}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/", env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "OK", (r.stdout, r.stderr[-2000:])


def _np_ms_ssim(x, y):
    """Independent float64 restatement (numpy) of MS-SSIM with data_range 1, for one [C,H,W] pair."""
    k = np.arange(11) - 5.0
    g = np.exp(-k * k / (2 * 1.5 ** 2))
    g /= g.sum()

    def blur(a):
        a = np.apply_along_axis(lambda v: np.convolve(v, g, mode="valid"), -1, a)
        return np.apply_along_axis(lambda v: np.convolve(v, g, mode="valid"), -2, a)

    def terms(a, b):
        c1, c2 = 0.01 ** 2, 0.03 ** 2
        ma, mb = blur(a), blur(b)
        va, vb, cab = blur(a * a) - ma * ma, blur(b * b) - mb * mb, blur(a * b) - ma * mb
        cs = (2 * cab + c2) / (va + vb + c2)
        return ((2 * ma * mb + c1) / (ma * ma + mb * mb + c1) * cs).mean((-1, -2)), cs.mean((-1, -2))

    w = [0.0448, 0.2856, 0.3001, 0.2363, 0.1333]
    acc = np.ones(x.shape[0])
    for i in range(5):
        full, cs = terms(x, y)
        acc *= np.maximum(full if i == 4 else cs, 0) ** w[i]
        if i < 4:
            x = x.reshape(x.shape[0], x.shape[1] // 2, 2, x.shape[2] // 2, 2).mean((2, 4))
            y = y.reshape(y.shape[0], y.shape[1] // 2, 2, y.shape[2] // 2, 2).mean((2, 4))
    return acc.mean()


def test_ms_ssim_shim_against_a_float64_restatement():
    from hesic_b200 import compat
    compat.install()
    import pytorch_msssim
    if not getattr(pytorch_msssim, "_HESIC_STUB", False):
        pytest.skip("the real pytorch_msssim is installed")
    g = np.random.default_rng(5)
    a = g.random((2, 3, 192, 256))
    b = np.clip(a + 0.1 * g.standard_normal(a.shape), 0, 1)
    got = pytorch_msssim.ms_ssim(torch.from_numpy(a).float(), torch.from_numpy(b).float(), data_range=1, size_average=False)
    ref = np.array([_np_ms_ssim(a[i], b[i]) for i in range(2)])
    assert np.allclose(got.numpy(), ref, rtol=2e-5, atol=2e-6), (got, ref)
    assert float(pytorch_msssim.ms_ssim(torch.from_numpy(a).float(), torch.from_numpy(a).float(), data_range=1)) == pytest.approx(1.0, abs=1e-6)
    with pytest.raises(AssertionError):
        pytorch_msssim.ms_ssim(torch.rand(1, 3, 128, 128), torch.rand(1, 3, 128, 128), data_range=1)


def test_plan_and_engine_handles_are_never_shared_by_copies():
    """copy.deepcopy / pickle of a module that cached a C handle yields a fresh unloaded plan, never a second owner of the
    same handle (a double hesic_conv_destroy), and engines are dropped from copies."""
    import copy
    import pickle
    from hesic_b200 import compat
    compat.install()
    from hesic_b200 import functional as F
    from hesic_b200.enhance import EnConvPlan
    p = F.ConvPlan(8, 16, 3, 1, 1)
    q = copy.deepcopy(p)
    r = pickle.loads(pickle.dumps(p))
    assert q.h != p.h and r.h != p.h and q.geom == p.geom == r.geom and q._key is None
    e = EnConvPlan(32, 32)
    e2, e3 = copy.deepcopy(e), pickle.loads(pickle.dumps(e))
    assert e2.h != e.h and e3.h != e.h and e2.geom == e.geom
    from hesic_b200 import homography as homo
    import newnet1
    lin = homo.Linear(4, 2)
    object.__setattr__(lin, "_plan", F.ConvPlan(4, 2, 1, 1, 0))
    lin2 = copy.deepcopy(lin)
    assert lin2._plan.h != lin._plan.h
    net = newnet1.HSIC(32, 48, 2)
    eng = net.hesic_engine
    net2 = copy.deepcopy(net)
    assert net2._engine is None and net._engine is eng and net2.hesic_engine.m is net2
    # the entropy models' coder objects own C handles too (r01: a copied model freed them twice at collection time)
    net3 = pickle.loads(pickle.dumps(net))
    assert net3.entropy_bottleneck1.entropy_coder is not net.entropy_bottleneck1.entropy_coder
    del net2, net3
    import gc
    gc.collect()
    conv = net.encoder1.g_a_conv1
    object.__setattr__(conv, "_hesic_plan", F.ConvPlan(3, 32, 5, 2, 2))
    assert copy.deepcopy(conv)._hesic_plan is None
    en = newnet1.Independent_EN()
    c1 = en.EH1.conv1
    object.__setattr__(c1, "_hesic_en_plan", EnConvPlan(c1.in_channels, c1.out_channels))
    en.hesic_engine
    en2 = pickle.loads(pickle.dumps(en))
    assert en2.__dict__.get("_engine") is None and en2.EH1.conv1._hesic_en_plan.h != c1._hesic_en_plan.h
