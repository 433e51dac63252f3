"""Synthetic stereo pairs, homographies and weights for the HESIC forward path.

No dataset or checkpoint ships with the reference (Readme.md:51-76), so every
test, golden fixture and bench run uses the generators below.  They depend on
numpy's PCG64 only (bit-reproducible across machines and torch builds), never
on torch's RNG, so the container that writes ``tests/golden`` and the GPU box
that checks it see identical bytes.

Recipe follows SURVEY.md section 8(d): band-limited images, a 12-pixel
horizontal disparity, a near-identity homography, Kaiming-scaled convolutions
and the "R1" rescale that keeps the latents away from the all-zero symbol.
"""
import math
import re
import zlib

import numpy as np
import torch


def _rng(seed, name=""):
    return np.random.default_rng([int(seed), zlib.crc32(name.encode())])


def _box_blur(a, k):
    """k x k mean filter over the last two axes ('valid')."""
    c = np.cumsum(np.cumsum(np.pad(a, [(0, 0)] * (a.ndim - 2) + [(1, 0), (1, 0)]), axis=-1), axis=-2)
    return (c[..., k:, k:] - c[..., :-k, k:] - c[..., k:, :-k] + c[..., :-k, :-k]) / float(k * k)


def stereo_pairs(B, H=512, W=512, seed=1234):
    """Returns x1, x2 [B,3,H,W] float32 in [0,1] and h_matrix [B,3,3] float32."""
    g = _rng(seed, "pairs")
    base = g.random((B, 3, H + 8, W + 8), dtype=np.float64)
    base = _box_blur(base, 9)  # -> [B,3,H,W]
    lo = base.min(axis=(1, 2, 3), keepdims=True)
    hi = base.max(axis=(1, 2, 3), keepdims=True)
    base = (base - lo) / (hi - lo)
    x1 = base
    x2 = np.roll(base, -12, axis=3) + 0.01 * g.standard_normal(base.shape)
    x2 = np.clip(x2, 0.0, 1.0)
    h = np.tile(np.eye(3), (B, 1, 1))
    h[:, 0, 0] += g.uniform(-0.01, 0.01, B)
    h[:, 0, 1] += g.uniform(-0.01, 0.01, B)
    h[:, 1, 0] += g.uniform(-0.01, 0.01, B)
    h[:, 1, 1] += g.uniform(-0.01, 0.01, B)
    h[:, 0, 2] = g.uniform(-24.0, 0.0, B) * (W / 512.0)
    h[:, 1, 2] = g.uniform(-2.0, 2.0, B) * (H / 512.0)
    h[:, 2, 0] = g.uniform(-1e-5, 1e-5, B)
    h[:, 2, 1] = g.uniform(-1e-5, 1e-5, B)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.float32)))
    return t(x1), t(x2), t(h)


def _fan_in(name, shape, transposed):
    # torch.nn.init._calculate_fan_in_and_fan_out: fan_in = size(1) * receptive field
    rf = 1
    for s in shape[2:]:
        rf *= s
    return shape[1] * rf


def synth_state_dict(model, seed=0, latent_gain=4.0, synthesis_gain=0.45, image_gain=0.08):
    """Fill every floating tensor of ``model.state_dict()`` deterministically.

    * conv / deconv weights: N(0, 2/fan_in) (what ``kaiming_normal_`` draws,
      newnet1.py:64-69), biases N(0, 0.02) so the bias path is exercised;
    * GDN: beta = 1 + U(0,0.5), gamma = 0.1*I + |N(0,0.004)| stored through the
      reparametrisation of gdn.py:46-53 / parametrizers.py:38-39;
    * EntropyBottleneck: the init of entropy_models.py:276-293 plus small
      noise on matrices/factors and non-zero medians;
    * R1: last analysis conv of both encoders scaled by ``latent_gain`` and the
      sigma head biased to 1 so symbols and likelihoods are non-trivial; the
      synthesis deconvs are damped (``synthesis_gain``, ``image_gain``, bias 0.5
      on the RGB head) so the three IGDNs do not blow the reconstruction up and
      the "twiceLeft" re-encode of x1_hat sees image-like values.
    Integer buffers and constant buffers (bounds, pedestals, masks, targets)
    keep the values the constructor gave them.
    """
    sd = model.state_dict() if hasattr(model, "state_dict") else dict(model)
    out = {}
    pedestal = (2.0 ** -18) ** 2
    for name, t in sd.items():
        leaf = name.rsplit(".", 1)[-1]
        if not t.is_floating_point() or t.numel() == 0:
            out[name] = t.clone()
            continue
        g = _rng(seed, name)
        shape = tuple(t.shape)
        parent = name.rsplit(".", 1)[0] if "." in name else ""
        if leaf in ("bound", "pedestal", "target", "mask", "scale_bound", "scale_table"):
            out[name] = t.clone()
        elif "gdn" in parent.rsplit(".", 1)[-1] and leaf == "beta":
            beta = 1.0 + g.uniform(0.0, 0.5, shape)
            out[name] = torch.from_numpy(np.sqrt(np.maximum(beta + pedestal, pedestal)).astype(np.float32))
        elif "gdn" in parent.rsplit(".", 1)[-1] and leaf == "gamma":
            gamma = 0.1 * np.eye(shape[0]) + np.abs(g.standard_normal(shape)) * 0.004
            out[name] = torch.from_numpy(np.sqrt(np.maximum(gamma + pedestal, pedestal)).astype(np.float32))
        elif "_matrices" in name:
            out[name] = (t.double() + torch.from_numpy(g.standard_normal(shape) * 0.05)).float()
        elif "_biases" in name:
            out[name] = torch.from_numpy(g.uniform(-0.5, 0.5, shape).astype(np.float32))
        elif "_factors" in name:
            out[name] = torch.from_numpy((g.standard_normal(shape) * 0.1).astype(np.float32))
        elif leaf == "quantiles":
            q = np.zeros(shape)
            med = g.standard_normal(shape[0]) * 0.3
            q[:, 0, 0] = med - 10.0
            q[:, 0, 1] = med
            q[:, 0, 2] = med + 10.0
            out[name] = torch.from_numpy(q.astype(np.float32))
        elif leaf == "weight" and len(shape) == 1:
            # nn.GroupNorm scale (DSIC, ywz/DSIC/mynet6_plus.py:224-290): around its init value 1
            out[name] = torch.from_numpy((1.0 + g.standard_normal(shape) * 0.1).astype(np.float32))
        elif leaf == "weight" and len(shape) == 5:
            # nn.Conv3d of the DSIC cost volumes: kaiming scale
            out[name] = torch.from_numpy((g.standard_normal(shape) * math.sqrt(2.0 / _fan_in(name, shape, False))).astype(np.float32))
        elif leaf == "weight" and len(shape) == 4:
            std = math.sqrt(2.0 / _fan_in(name, shape, False))
            w = g.standard_normal(shape) * std
            if name.endswith("g_a_conv4.weight"):
                w = w * latent_gain
            elif re.search(r"g_s_conv[123]\.weight$", name):
                w = w * synthesis_gain
            elif name.endswith("g_s_conv4.weight"):
                w = w * image_gain
            elif name.endswith("after_conv.weight"):
                w = w * 0.3
            out[name] = torch.from_numpy(w.astype(np.float32))
        elif leaf == "bias":
            b = g.standard_normal(shape) * 0.02
            if "gmm_sigma.4" in name:
                b = b + 1.0
            if name.endswith("g_s_conv4.bias"):
                b = b + 0.5  # image-like reconstructions in ~[0,1]
            if name.startswith("entropy_parameters") and name.endswith(".4.bias"):
                b[: shape[0] // 2] += 1.0  # scales_hat chunk, newnet1_joint.py:689,724
            out[name] = torch.from_numpy(b.astype(np.float32))
        else:
            out[name] = torch.from_numpy((g.standard_normal(shape) * 0.05).astype(np.float32))
    return out


def rd_metrics(out, x1, x2):
    """bpp / PSNR exactly as the caller's criterion (ywz/mywork/test3real.py:90-124)."""
    N, _, H, W = x1.shape
    num_pixels = N * H * W
    res = {}
    lik = out["likelihoods"]
    term = {k: float(torch.log(v.double()).sum() / (-math.log(2) * num_pixels)) for k, v in lik.items()}
    res["bpp"] = sum(term.values())
    res["bpp1"] = term["y1"] + term["z1"]
    res["bpp2"] = term["y2"] + term["z2"]
    mse1 = float(((out["x1_hat"].double() - x1.double()) ** 2).mean())
    mse2 = float(((out["x2_hat"].double() - x2.double()) ** 2).mean())
    res["psnr1"] = 10 * math.log10(1 / mse1)
    res["psnr2"] = 10 * math.log10(1 / mse2)
    return res
