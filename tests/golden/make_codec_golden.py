#!/usr/bin/env python
"""Writes the golden codec files tests/golden/codec_<model>.{npz,bin} (run on a B200: the tables come from the CUDA
path).  One synthetic 128x128 pair (seed 1234), seeded weights (seed 0), HESIC and HESIC+:

    python tests/golden/make_codec_golden.py [outdir]        # default: gpurun_out/codec_golden

The byte streams are this library's own format (the reference's ``range_coder`` package is un-vendored: parity of the
stream is unpinned); the fixtures pin THAT format and the table arithmetic against silent drift --
tests/test_gpu_codec.py::test_codec_files_match_the_committed_golden."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import hesic_b200  # noqa: E402
from hesic_b200 import synth  # noqa: E402

hesic_b200.install()
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "codec_golden")
os.makedirs(out, exist_ok=True)
for tag, modname in (("hsic_newnet1", "newnet1"), ("hsic_joint", "newnet1_joint")):
    mod = __import__(modname)
    net = mod.HSIC(128, 192, 5).eval()
    net.load_state_dict(synth.synth_state_dict(net, seed=0))
    net = net.to("cuda:0")
    net.entropy_bottleneck1.update(force=True)
    net.entropy_bottleneck2.update(force=True)
    x1, x2, h = (t.to("cuda:0") for t in synth.stereo_pairs(1, 128, 128, seed=1234))
    enc = net.compress(x1, x2, h, "codec_" + tag, output_path=out)
    dec = net.decompress(x1, x2, h, "codec_" + tag, output_path=out)
    assert all(torch.equal(dec[k], enc[k]) for k in ("y1_hat", "y2_hat", "z1_hat", "z2_hat"))
    print(tag, "bpp_real", enc["bpp_real"], {e: os.path.getsize(os.path.join(out, f"codec_{tag}.{e}")) for e in ("npz", "bin")})

# DSIC (mynet6_plus.py:799-1350): same tables, coder and file layout; one 64x256 pair
import mynet6_plus  # noqa: E402

net = mynet6_plus.DSIC(128, 192, 21, 32, 5).eval()
net.load_state_dict(synth.synth_state_dict(net, seed=0))
net = net.to("cuda:0")
net.entropy_bottleneck1.update(force=True)
net.entropy_bottleneck2.update(force=True)
x1, x2, _ = (t.to("cuda:0") for t in synth.stereo_pairs(1, 64, 256, seed=1234))
enc = net.compress(x1, x2, "codec_dsic", output_path=out)
dec = net.decompress("cuda:0", "codec_dsic", output_path=out)
assert all(torch.equal(dec[k], enc[k]) for k in ("y1_hat", "y2_hat", "z1_hat", "z2_hat"))
print("dsic bpp_real", enc["bpp_real"], {e: os.path.getsize(os.path.join(out, f"codec_dsic.{e}")) for e in ("npz", "bin")})
