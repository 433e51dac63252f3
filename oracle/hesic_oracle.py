"""CPU oracle for the HESIC / HESIC+ stereo forward path.  TEST INFRASTRUCTURE ONLY.

A functional restatement, in plain torch-CPU fp32, of what the reference's
``HSIC.forward`` computes, written against a ``state_dict`` with the
reference's key names.  It is the checker for the CUDA path; nothing in the
product (``hesic_b200/``) imports it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may use it.

Pinning: ``tests/golden/make_golden.py`` imports the *reference itself* from
/root/reference (in the build container) and stores its outputs; the CPU test
suite checks every function here against those fixtures.  One input of the
path is NOT pinned by the reference: ``kornia.warp_perspective`` is an
un-vendored, un-pinned third-party dependency, so
``warp_perspective`` below restates kornia's published algorithm
(``align_corners=True`` convention) and the goldens were produced with that
same restatement standing in for kornia: **warp parity is unpinned**.

Every function cites the reference file:line it follows (paths relative to
/root/reference).
"""
import math

import torch
import torch.nn.functional as F

REPARAM_OFFSET = 2.0 ** -18


# ----------------------------------------------------------------------------
# compressai/ops/parametrizers.py:21-44, bound_ops.py:19-52
def nonneg_reparam(p, minimum=0.0):
    pedestal = REPARAM_OFFSET ** 2
    bound = (minimum + pedestal) ** 0.5
    out = torch.max(p, torch.tensor([bound], dtype=p.dtype))
    return out ** 2 - torch.tensor([pedestal], dtype=p.dtype)


# compressai/layers/gdn.py:55-70
def gdn(x, beta_p, gamma_p, inverse=False, beta_min=1e-6):
    C = x.shape[1]
    beta = nonneg_reparam(beta_p, beta_min)
    gamma = nonneg_reparam(gamma_p, 0.0).reshape(C, C, 1, 1)
    norm = F.conv2d(x ** 2, gamma, beta)
    norm = torch.sqrt(norm) if inverse else torch.rsqrt(norm)
    return x * norm


# compressai/models/utils.py:104-118
def conv(x, w, b, stride=2, k=None):
    k = w.shape[-1] if k is None else k
    return F.conv2d(x, w, b, stride=stride, padding=k // 2)


def deconv(x, w, b, stride=2):
    k = w.shape[-1]
    return F.conv_transpose2d(x, w, b, stride=stride, padding=k // 2, output_padding=stride - 1)


# ----------------------------------------------------------------------------
# kornia.warp_perspective (third-party, un-vendored; call sites newnet1.py:746,753,767)
def _normal_transform_pixel(h, w, dtype=torch.float32):
    return torch.tensor([[2.0 / (w - 1), 0.0, -1.0], [0.0, 2.0 / (h - 1), -1.0], [0.0, 0.0, 1.0]], dtype=dtype)


def warp_perspective(src, M, dsize, align_corners=True, dtype=torch.float32):
    """``dtype=torch.float64`` evaluates the same algorithm in double: kornia's fp32
    normalise -> invert -> denormalise chain carries ~1e-3 px of rounding noise at 512x512
    (SURVEY.md section 7 "Warp numerics"), which the double evaluation removes."""
    B, _, H, W = src.shape
    h_out, w_out = dsize
    src = src.to(dtype)
    src_norm = _normal_transform_pixel(H, W, dtype)[None]
    dst_norm = _normal_transform_pixel(h_out, w_out, dtype)[None]
    dst_norm_trans_src_norm = dst_norm @ (M.to(dtype) @ torch.inverse(src_norm))
    src_norm_trans_dst_norm = torch.inverse(dst_norm_trans_src_norm)
    xs = torch.linspace(-1, 1, w_out, dtype=dtype)
    ys = torch.linspace(-1, 1, h_out, dtype=dtype)
    gx, gy = torch.meshgrid(xs, ys, indexing="xy")  # [h_out, w_out]
    pts = torch.stack((gx, gy, torch.ones_like(gx)), dim=-1)  # [h,w,3]
    p = pts[None] @ src_norm_trans_dst_norm[:, None].transpose(-1, -2)  # [B,h,w,3]
    z = p[..., 2:3]
    scale = torch.where(z.abs() > 1e-8, 1.0 / z, torch.ones_like(z))
    grid = scale * p[..., :2]
    return F.grid_sample(src, grid, mode="bilinear", padding_mode="zeros", align_corners=align_corners)


# ----------------------------------------------------------------------------
# compressai/entropy_models/entropy_models.py:98-125
def quantize(x, mode, means=None):
    if mode not in ("dequantize", "symbols"):
        raise ValueError(f'Invalid quantization mode: "{mode}"')
    out = x.clone()
    if means is not None:
        out = out - means
    out = torch.round(out)
    if mode == "dequantize":
        if means is not None:
            out = out + means
        return out
    return out.int()


# entropy_models.py:350-369
def eb_logits_cumulative(v, matrices, biases, factors):
    logits = v
    for i in range(len(matrices)):
        logits = torch.matmul(F.softplus(matrices[i]), logits)
        logits = logits + biases[i]
        if i < len(factors):
            logits = logits + torch.tanh(factors[i]) * torch.tanh(logits)
    return logits


def eb_params(sd, prefix):
    mats = [sd[f"{prefix}._matrices.{i}"] for i in range(5)]
    bias = [sd[f"{prefix}._biases.{i}"] for i in range(5)]
    fac = [sd[f"{prefix}._factors.{i}"] for i in range(4)]
    return mats, bias, fac, sd[f"{prefix}.quantiles"]


# entropy_models.py:384-411 (eval mode), :371-382
def entropy_bottleneck(x, mats, bias, fac, quantiles, likelihood_bound=1e-9):
    xp = x.permute(1, 2, 3, 0).contiguous()
    shape = xp.shape
    values = xp.reshape(xp.size(0), 1, -1)
    medians = quantiles[:, :, 1:2]
    outputs = quantize(values, "dequantize", medians)
    lower = eb_logits_cumulative(outputs - 0.5, mats, bias, fac)
    upper = eb_logits_cumulative(outputs + 0.5, mats, bias, fac)
    sign = -torch.sign(lower + upper)
    lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
    if likelihood_bound > 0:
        lik = torch.max(lik, torch.tensor([likelihood_bound]))
    outputs = outputs.reshape(shape).permute(3, 0, 1, 2).contiguous()
    lik = lik.reshape(shape).permute(3, 0, 1, 2).contiguous()
    return outputs, lik


# entropy_models.py:615-620
def std_cumulative(x):
    return 0.5 * torch.erfc(float(-(2 ** -0.5)) * x)


# entropy_models.py:661-702 (GaussianMixtureConditional, eval; quantises with means=None :697)
def gmm_conditional(y, scales, means, weights, K, scale_bound=0.11, likelihood_bound=1e-9):
    y_hat = torch.round(y)
    M = y.shape[1]
    lik = None
    for k in range(K):
        sl = slice(M * k, M * (k + 1))
        values = torch.abs(y_hat - means[:, sl])
        s = torch.max(scales[:, sl], torch.tensor([scale_bound]))
        upper = std_cumulative((0.5 - values) / s)
        lower = std_cumulative((-0.5 - values) / s)
        term = (upper - lower) * weights[:, sl]
        lik = term if lik is None else lik + term
    if likelihood_bound > 0:
        lik = torch.max(lik, torch.tensor([likelihood_bound]))
    return y_hat, lik


# entropy_models.py:528-554 (GaussianConditional, eval)
def gaussian_conditional(y, scales, means=None, scale_bound=0.11, likelihood_bound=1e-9):
    y_hat = quantize(y, "dequantize", means)
    values = y_hat - means if means is not None else y_hat
    s = torch.max(scales, torch.tensor([scale_bound]))
    values = torch.abs(values)
    upper = std_cumulative((0.5 - values) / s)
    lower = std_cumulative((-0.5 - values) / s)
    lik = upper - lower
    if likelihood_bound > 0:
        lik = torch.max(lik, torch.tensor([likelihood_bound]))
    return y_hat, lik


# entropy_models.py:556-562 ; scale table compressai/models/priors.py get_scale_table
def scale_table(mn=0.11, mx=256.0, levels=64):
    return torch.exp(torch.linspace(math.log(mn), math.log(mx), levels))


def build_indexes(scales, table, scale_bound=0.11):
    s = torch.max(scales, torch.tensor([scale_bound]))
    idx = torch.full(s.shape, len(table) - 1, dtype=torch.int32)
    for t in table[:-1]:
        idx -= (s <= t).int()
    return idx


# entropy_models.py:413-418
def eb_build_indexes(size):
    N, C, H, W = size
    return torch.arange(C).view(1, -1, 1, 1).int().repeat(N, 1, H, W)


# ywz/mywork/newnet1.py:441-453 (python double loop == global max per (b,c))
def spatial_pool2d(X):
    return X.amax(dim=(2, 3), keepdim=True).float()


# ----------------------------------------------------------------------------
# model pieces, keyed by reference state_dict names
def _cw(sd, p):
    return sd[p + ".weight"], sd[p + ".bias"]


def _gdn(sd, p, x, inverse=False):
    return gdn(x, sd[p + ".beta"], sd[p + ".gamma"], inverse)


# newnet1.py:580-601
def encoder1(sd, x, p="encoder1"):
    x = _gdn(sd, p + ".g_a_gdn1", conv(x, *_cw(sd, p + ".g_a_conv1")))
    x = _gdn(sd, p + ".g_a_gdn2", conv(x, *_cw(sd, p + ".g_a_conv2")))
    x = _gdn(sd, p + ".g_a_gdn3", conv(x, *_cw(sd, p + ".g_a_conv3")))
    return conv(x, *_cw(sd, p + ".g_a_conv4"))


# newnet1.py:627-655
def encoder2(sd, x1_warp, x2, p="encoder2"):
    pre = conv(torch.cat((x1_warp, x2), dim=-3), *_cw(sd, p + ".pre_conv"), stride=1)
    pre = _gdn(sd, p + ".pre_gdn", pre)
    return encoder1(sd, pre, p)


# newnet1.py:603-624
def decoder1(sd, y_hat, p="decoder1"):
    x = _gdn(sd, p + ".g_s_gdn1", deconv(y_hat, *_cw(sd, p + ".g_s_conv1")), True)
    x = _gdn(sd, p + ".g_s_gdn2", deconv(x, *_cw(sd, p + ".g_s_conv2")), True)
    x = _gdn(sd, p + ".g_s_gdn3", deconv(x, *_cw(sd, p + ".g_s_conv3")), True)
    return deconv(x, *_cw(sd, p + ".g_s_conv4"))


# newnet1.py:657-692
def decoder2(sd, y_hat, x1_hat_warp, p="decoder2"):
    c4 = decoder1(sd, y_hat, p)
    a1 = _gdn(sd, p + ".after_gdn", c4, True)
    return deconv(torch.cat((a1, x1_hat_warp), dim=-3), *_cw(sd, p + ".after_conv"), stride=1)


# newnet1.py:420-437
def encode_hyper(sd, y, p):
    q = p + ".encode_hyper"
    x = F.relu(conv(torch.abs(y), *_cw(sd, q + ".0"), stride=1))
    x = F.relu(conv(x, *_cw(sd, q + ".2")))
    return conv(x, *_cw(sd, q + ".4"))


def _mix_softmax(t, K, M):
    t = torch.reshape(t, (-1, K, M, 1, 1))
    t = F.softmax(t, dim=-4)
    return torch.reshape(t, (-1, M * K, 1, 1))


# newnet1.py:456-514
def gmm_hyper_y1(sd, z, p, K, M):
    s = F.relu(deconv(z, *_cw(sd, p + ".gmm_sigma.0")))
    s = F.relu(deconv(s, *_cw(sd, p + ".gmm_sigma.2")))
    s = F.relu(conv(s, *_cw(sd, p + ".gmm_sigma.4"), stride=1))
    m = F.leaky_relu(deconv(z, *_cw(sd, p + ".gmm_means.0")))
    m = F.leaky_relu(deconv(m, *_cw(sd, p + ".gmm_means.2")))
    m = conv(m, *_cw(sd, p + ".gmm_means.4"), stride=1)
    w = F.leaky_relu(deconv(z, *_cw(sd, p + ".gmm_weights.0")))
    w = deconv(w, *_cw(sd, p + ".gmm_weights.2"))
    w = F.leaky_relu(spatial_pool2d(w))
    w = conv(w, *_cw(sd, p + ".gmm_weights.5"), stride=1)
    return s, m, _mix_softmax(w, K, M)


# newnet1.py:517-577 (UpsamplingBilinear2d == align_corners=True)
def gmm_hyper_y2(sd, z2, y1, p, K, M):
    up = F.interpolate(z2, scale_factor=4, mode="bilinear", align_corners=True)
    c = torch.cat((up, y1), dim=-3)
    s = F.relu(conv(c, *_cw(sd, p + ".gmm_sigma.0"), stride=1))
    s = F.relu(conv(s, *_cw(sd, p + ".gmm_sigma.2"), stride=1))
    s = F.relu(conv(s, *_cw(sd, p + ".gmm_sigma.4"), stride=1))
    m = F.leaky_relu(conv(c, *_cw(sd, p + ".gmm_means.0"), stride=1))
    m = F.leaky_relu(conv(m, *_cw(sd, p + ".gmm_means.2"), stride=1))
    m = conv(m, *_cw(sd, p + ".gmm_means.4"), stride=1)
    w = F.leaky_relu(conv(c, *_cw(sd, p + ".gmm_weights.0"), stride=1))
    w = conv(w, *_cw(sd, p + ".gmm_weights.2"), stride=1)
    w = F.leaky_relu(spatial_pool2d(w))
    w = conv(w, *_cw(sd, p + ".gmm_weights.5"), stride=1)
    return s, m, _mix_softmax(w, K, M)


def hsic_forward(sd, x1, x2, h, K=5, twice_left=True, align_corners=True, taps=None):
    """newnet1.HSIC.forward (newnet1.py:724-783); ``twice_left=False`` gives the
    .trash/newnet9.py:620-661 variant that test3real.py imports.  ``taps`` (a dict)
    collects intermediate tensors for per-op parity tests."""
    T = taps if taps is not None else {}
    M = sd["encoder1.g_a_conv4.weight"].shape[0]
    size = (x1.shape[-2], x1.shape[-1])
    y1 = encoder1(sd, x1)
    z1 = encode_hyper(sd, y1, "_h_a1")
    z1_hat, z1_lik = entropy_bottleneck(z1, *eb_params(sd, "entropy_bottleneck1"))
    s1, m1, w1 = gmm_hyper_y1(sd, z1_hat, "_h_s1", K, M)
    y1_hat, y1_lik = gmm_conditional(y1, s1, m1, w1, K)
    x1_hat = decoder1(sd, y1_hat)
    x1_warp = warp_perspective(x1, h, size, align_corners)
    y2 = encoder2(sd, x1_warp, x2)
    x1_hat_warp = warp_perspective(x1_hat, h, size, align_corners)
    if twice_left:
        y1_cond = torch.round(encoder1(sd, x1_hat_warp))
    else:
        y1_cond = y1_hat
    z2 = encode_hyper(sd, y2, "_h_a2")
    z2_hat, z2_lik = entropy_bottleneck(z2, *eb_params(sd, "entropy_bottleneck2"))
    s2, m2, w2 = gmm_hyper_y2(sd, z2_hat, y1_cond, "_h_s2", K, M)
    y2_hat, y2_lik = gmm_conditional(y2, s2, m2, w2, K)
    x2_hat = decoder2(sd, y2_hat, x1_hat_warp)
    T.update(y1=y1, z1=z1, z1_hat=z1_hat, sigma1=s1, means1=m1, weights1=w1, x1_warp=x1_warp, y2=y2,
             x1_hat_warp=x1_hat_warp, y1_cond=y1_cond, z2=z2, z2_hat=z2_hat, sigma2=s2, means2=m2, weights2=w2)
    return {"x1_hat": x1_hat, "x2_hat": x2_hat, "y1_hat": y1_hat, "y2_hat": y2_hat,
            "likelihoods": {"y1": y1_lik, "y2": y2_lik, "z1": z1_lik, "z2": z2_lik}}


def _seq(sd, p, x, spec):
    """spec: list of ('conv'|'deconv', idx, stride) / 'lrelu' entries, nn.Sequential style."""
    for item in spec:
        if item == "lrelu":
            x = F.leaky_relu(x)
        else:
            kind, idx, stride = item
            w, b = _cw(sd, f"{p}.{idx}")
            x = conv(x, w, b, stride=stride) if kind == "conv" else deconv(x, w, b, stride=stride)
    return x


_HA = [("conv", 0, 1), "lrelu", ("conv", 2, 2), "lrelu", ("conv", 4, 2)]
_HS = [("deconv", 0, 2), "lrelu", ("deconv", 2, 2), "lrelu", ("conv", 4, 1)]
_EP = [("conv", 0, 1), "lrelu", ("conv", 2, 1), "lrelu", ("conv", 4, 1)]


def masked_conv(sd, p, x):
    # compressai/layers/layers.py:21-45 (weight *= mask, then conv, padding 2)
    w = sd[p + ".weight"] * sd[p + ".mask"]
    return F.conv2d(x, w, sd[p + ".bias"], stride=1, padding=w.shape[-1] // 2)


def hsic_joint_forward(sd, x1, x2, h, align_corners=True, taps=None):
    """newnet1_joint.HSIC.forward (newnet1_joint.py:675-753), eval mode."""
    T = taps if taps is not None else {}
    size = (x1.shape[-2], x1.shape[-1])
    y1 = encoder1(sd, x1)
    z1 = _seq(sd, "h_a1", y1, _HA)
    z1_hat, z1_lik = entropy_bottleneck(z1, *eb_params(sd, "entropy_bottleneck1"))
    params1 = _seq(sd, "h_s1", z1_hat, _HS)
    y1_hat = torch.round(y1)
    ctx1 = masked_conv(sd, "context_prediction1", y1_hat)
    gp1 = _seq(sd, "entropy_parameters1", torch.cat((params1, ctx1), dim=1), _EP)
    sc1, mu1 = gp1.chunk(2, 1)
    _, y1_lik = gaussian_conditional(y1, sc1, mu1)
    x1_hat = decoder1(sd, y1_hat)
    x1_warp = warp_perspective(x1, h, size, align_corners)
    y2 = encoder2(sd, x1_warp, x2)
    z2 = _seq(sd, "h_a2", y2, _HA)
    z2_hat, z2_lik = entropy_bottleneck(z2, *eb_params(sd, "entropy_bottleneck2"))
    x1_hat_warp = warp_perspective(x1_hat, h, size, align_corners)
    y1_cond = torch.round(encoder1(sd, x1_hat_warp))
    params2 = _seq(sd, "h_s2", z2_hat, _HS)
    y2_hat = torch.round(y2)
    ctx2 = masked_conv(sd, "context_prediction2", y2_hat)
    gp2 = _seq(sd, "entropy_parameters2", torch.cat((params2, ctx2, y1_cond), dim=1), _EP)
    sc2, mu2 = gp2.chunk(2, 1)
    _, y2_lik = gaussian_conditional(y2, sc2, mu2)
    x2_hat = decoder2(sd, y2_hat, x1_hat_warp)
    T.update(y1=y1, z1=z1, z1_hat=z1_hat, scales1=sc1, means1=mu1, y2=y2, z2=z2, z2_hat=z2_hat,
             scales2=sc2, means2=mu2, x1_warp=x1_warp, x1_hat_warp=x1_hat_warp, y1_cond=y1_cond)
    return {"x1_hat": x1_hat, "x2_hat": x2_hat, "y1_hat": y1_hat, "y2_hat": y2_hat,
            "likelihoods": {"y1": y1_lik, "y2": y2_lik, "z1": z1_lik, "z2": z2_lik}}


# ----------------------------------------------------------------------------
# Independent_EN (newnet1.py:272-311, 1278-1300; layers.py:125-147) -- SURVEY 8f rank 1
def _resblock(sd, p, x):
    out = F.leaky_relu(F.conv2d(x, *_cw(sd, p + ".conv1"), padding=1))
    out = F.leaky_relu(F.conv2d(out, *_cw(sd, p + ".conv2"), padding=1))
    return out + x


def _enh_block(sd, p, x):
    out = _resblock(sd, p + ".RB1", x)
    out = _resblock(sd, p + ".RB2", out)
    out = _resblock(sd, p + ".RB3", out)
    return out + x


def enhancement(sd, p, x, other_warp):
    out = F.conv2d(torch.cat((x, other_warp), dim=-3), *_cw(sd, p + ".conv1"), padding=1)
    for eb in ("EB1", "EB2", "EB3"):
        out = _enh_block(sd, f"{p}.{eb}", out)
    out = F.conv2d(out, *_cw(sd, p + ".conv2"), padding=1)
    return out + x


def dsic_independent_en_forward(sd, x1_hat, x2_hat):
    """ywz/DSIC/mynet6_plus.py:57-100: the DSIC enhancement has no cross-view input (conv1 is 3 -> 32 on x alone)."""
    def one(p, x):
        out = F.conv2d(x, *_cw(sd, p + ".conv1"), padding=1)
        for eb in ("EB1", "EB2", "EB3"):
            out = _enh_block(sd, f"{p}.{eb}", out)
        return F.conv2d(out, *_cw(sd, p + ".conv2"), padding=1) + x
    return {"x1_hat": one("EH1", x1_hat), "x2_hat": one("EH2", x2_hat)}


def independent_en_forward(sd, x1_hat, x2_hat, h, align_corners=True):
    size = (x1_hat.shape[-2], x1_hat.shape[-1])
    x1_hat_warp = warp_perspective(x1_hat, h, size, align_corners)
    x2_hat_warp = warp_perspective(x2_hat, torch.inverse(h), size, align_corners)
    return {"x1_hat": enhancement(sd, "EH1", x1_hat, x2_hat_warp),
            "x2_hat": enhancement(sd, "EH2", x2_hat, x1_hat_warp)}


# ----------------------------------------------------------------------------
# File codec of the stereo models (newnet1.py:934-978 / 1135-1180) -- SURVEY 8f rank 2.  The reference evaluates the
# pmf with torch ops (on 'cuda:0', hard-coded) and the integer table with numpy on the host; this follows the same
# op sequence on the CPU.  `range_coder` itself is un-vendored and un-pinned: parity of the byte stream is unpinned.
def std_cumulative_cr(x):
    """``_standardized_cumulative`` (newnet1.py:795-797: 0.5 * erfc(float(-(2 ** -0.5)) * x)) with a CORRECTLY ROUNDED
    fp32 erfc: the fp32 argument is evaluated in fp64 and rounded once.  torch.erfc itself is implementation-defined in
    the last bit (SLEEF on a CPU, CUDA's erfcf on the 'cuda:0' the reference hard-codes), and the integer tables built
    from it are the code both ends of the codec must agree on; the correctly rounded value is the one result every
    implementation can reproduce.  Used for the codec tables only -- the likelihood path keeps torch.erfc."""
    t = torch.tensor(-(2 ** -0.5), dtype=torch.float32) * x
    return 0.5 * torch.erfc(t.double()).float()


def codec_cdf_tables(scales, means, weights, K, channels, minmax, scale_bound=0.11):
    """-> int array [len(channels)*H*W, 2*minmax+2]: per latent element the cumulative-frequency row the reference passes
    to RangeEncoder.encode / RangeDecoder.decode, rows in the coding order (channel, h, w).  Same op sequence as
    newnet1.py:934-978, every step one IEEE fp32 operation, erfc correctly rounded (``std_cumulative_cr``)."""
    import numpy as np
    M = scales.shape[1] // K
    H, W = scales.shape[-2:]
    S = 2 * minmax + 1
    samples = torch.arange(0, S, dtype=torch.float32).reshape(S, 1, 1).expand(S, H, W)
    w_all = weights.reshape(-1)
    rows = []
    for ch in channels:
        idx = [int(ch) + k * M for k in range(K)]
        sigma, mu = scales[0, idx], means[0, idx] + minmax
        pmf = None
        for k in range(K):
            values = torch.abs(samples - mu[k])
            sc = torch.max(sigma[k], torch.tensor([scale_bound]))
            upper = std_cumulative_cr((0.5 - values) / sc)
            lower = std_cumulative_cr((-0.5 - values) / sc)
            term = (upper - lower) * w_all[idx[k]]
            pmf = term if pmf is None else pmf + term
        pmf = pmf.numpy()
        p = np.clip(pmf.reshape(S, H * W).T.copy(), np.float32(1.0 / 65536), np.float32(1.0))    # [H*W, S] rows, fp32
        for r in range(H * W):
            q = np.round(p[r] / np.sum(p[r]) * 65536)
            rows.append(np.concatenate(([0], np.add.accumulate(q))).astype(np.int64))
    return np.asarray(rows, dtype=np.int64).reshape(-1, S + 1)


# ----------------------------------------------------------------------------
# Homography front-end (ywz/mywork/model.py:53-111; test3real.py:171-181) -- SURVEY 8f rank 3
def homography_net_forward(sd, a, b):
    """Net.forward: four Blocks (conv3x3, ReLU, conv3x3, ReLU, [MaxPool 2x2]) then Flatten, FC 1024, ReLU, FC 8
    (Dropout is the identity in eval mode) -> delta [B,4,2]."""
    x = torch.cat((a, b), dim=1)
    for i in range(4):
        x = F.relu(F.conv2d(x, *_cw(sd, f"cnn.{i}.layers.0"), padding=1))
        x = F.relu(F.conv2d(x, *_cw(sd, f"cnn.{i}.layers.2"), padding=1))
        if i < 3:
            x = F.max_pool2d(x, 2, 2)
    x = x.reshape(x.shape[0], -1)
    x = F.relu(F.linear(x, sd["fc.2.weight"], sd["fc.2.bias"]))
    return F.linear(x, sd["fc.5.weight"], sd["fc.5.bias"]).view(-1, 4, 2)


def get_perspective_transform(src, dst):
    """kornia.get_perspective_transform (un-vendored; call sites test3real.py:179, model.py:26,108): the homography
    mapping four points src[B,4,2] onto dst[B,4,2], by the direct linear transform with h33 = 1, in float64."""
    x, y = src[..., 0].double(), src[..., 1].double()
    u, v = dst[..., 0].double(), dst[..., 1].double()
    z, o = torch.zeros_like(x), torch.ones_like(x)
    A = torch.cat([torch.stack([x, y, o, z, z, z, -x * u, -y * u], -1), torch.stack([z, z, z, x, y, o, -x * v, -y * v], -1)], 1)
    rhs = torch.cat([u, v], 1).unsqueeze(-1)
    sol = torch.linalg.solve(A, rhs).squeeze(-1)
    H = torch.cat([sol, torch.ones(src.shape[0], 1, dtype=sol.dtype)], 1).reshape(-1, 3, 3)
    return H.to(src.dtype)


def h_adjust(orishapea, orishapeb, resizeshapea, resizeshapeb, h):
    """test3real.py:56-66 (the caller's in-place rescale of H from the 256-pixel frame to the image frame)."""
    a, b = orishapea / resizeshapea, orishapeb / resizeshapeb
    h = h.clone()
    h[:, 0, :] = a * h[:, 0, :]
    h[:, :, 0] = (1.0 / a) * h[:, :, 0]
    h[:, 1, :] = b * h[:, 1, :]
    h[:, :, 1] = (1.0 / b) * h[:, :, 1]
    return h


# ----------------------------------------------------------------------------
# EntropyBottleneck.update() tables (entropy_models.py:302-343) -- float part; the
# pmf -> integer CDF step is oracle/coder_oracle.c (ops.cpp:24-81)
def eb_update_pmf(mats, bias, fac, quantiles):
    medians = quantiles[:, 0, 1]
    minima = torch.clamp(torch.ceil(medians - quantiles[:, 0, 0]).int(), min=0)
    maxima = torch.clamp(torch.ceil(quantiles[:, 0, 2] - medians).int(), min=0)
    offset = -minima
    pmf_start = medians - minima
    pmf_length = maxima + minima + 1
    max_length = int(pmf_length.max())
    samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
    lower = eb_logits_cumulative(samples - 0.5, mats, bias, fac)
    upper = eb_logits_cumulative(samples + 0.5, mats, bias, fac)
    sign = -torch.sign(lower + upper)
    pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
    tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
    return pmf, tail, pmf_length, offset


# ----------------------------------------------------------------------------
# DSIC (ywz/DSIC/mynet6_plus.py) -- BASELINE config 5, SURVEY.md 8a row 15
def _gn_relu(sd, p, x, groups):
    """nn.GroupNorm(groups, C) followed by nn.ReLU (mynet6_plus.py:224-238,262-290)."""
    return F.relu(F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], 1e-5))


# mynet6_plus.py:520-530, 543-553 (the DSIC file's Encoder1/Decoder1 also return the GDN outputs)
def dsic_encoder1(sd, x, p="encoder1"):
    g1 = _gdn(sd, p + ".g_a_gdn1", conv(x, *_cw(sd, p + ".g_a_conv1")))
    g2 = _gdn(sd, p + ".g_a_gdn2", conv(g1, *_cw(sd, p + ".g_a_conv2")))
    g3 = _gdn(sd, p + ".g_a_gdn3", conv(g2, *_cw(sd, p + ".g_a_conv3")))
    return conv(g3, *_cw(sd, p + ".g_a_conv4")), g1, g2, g3


def dsic_decoder1(sd, y_hat, p="decoder1"):
    g1 = _gdn(sd, p + ".g_s_gdn1", deconv(y_hat, *_cw(sd, p + ".g_s_conv1")), True)
    g2 = _gdn(sd, p + ".g_s_gdn2", deconv(g1, *_cw(sd, p + ".g_s_conv2")), True)
    g3 = _gdn(sd, p + ".g_s_gdn3", deconv(g2, *_cw(sd, p + ".g_s_conv3")), True)
    return deconv(g3, *_cw(sd, p + ".g_s_conv4")), g1, g2, g3


# mynet6_plus.py:216-246
def dsic_global_context(sd, y1, Fn, Cn, p="_global_context.global_net"):
    t = y1
    for i, gn in ((0, 1), (3, 4), (6, 7)):
        t = _gn_relu(sd, f"{p}.{gn}", conv(t, *_cw(sd, f"{p}.{i}"), stride=1), Fn)
    t = conv(t, *_cw(sd, f"{p}.9"), stride=1)
    return torch.reshape(t, (-1, 3, Fn // 3, Cn, t.size(-2), t.size(-1))).split(1, dim=1)


# mynet6_plus.py:249-313
def dsic_cost_volume(sd, p, h1, h2, d, scale, Fn, Cn, taps=None):
    """``taps`` (optional dict) receives every stage boundary, so that a test can feed a kernel the oracle's own
    input of that stage: m1a, h_out, d_up, v1, d_out, m3a, m3b, logits."""
    T = taps if taps is not None else {}
    F0 = Fn // 3
    x = torch.cat((h1, h2), dim=1)
    x = T["m1a"] = _gn_relu(sd, p + ".model1.1", conv(x, *_cw(sd, p + ".model1.0"), stride=1), 4)
    h_out = T["h_out"] = _gn_relu(sd, p + ".model1.4", conv(x, *_cw(sd, p + ".model1.3"), stride=1), 4)
    d_in = torch.reshape(d, (-1, d.size(-3), d.size(-2), d.size(-1)))
    d_up = T["d_up"] = F.interpolate(d_in, scale_factor=scale, mode="bilinear", align_corners=True)   # nn.UpsamplingBilinear2d
    v = torch.reshape(d_up, (-1, F0, Cn, d_up.size(-2), d_up.size(-1)))
    for i, gn in ((0, 1), (3, 4)):
        v = F.conv3d(v, sd[f"{p}.model2.{i}.weight"], sd[f"{p}.model2.{i}.bias"], stride=1, padding=2)
        v = F.relu(F.group_norm(v, 1, sd[f"{p}.model2.{gn}.weight"], sd[f"{p}.model2.{gn}.bias"], 1e-5))
        if i == 0:
            T["v1"] = v
    d_out = T["d_out"] = torch.reshape(v, (-1, F0 * Cn, v.size(-2), v.size(-1)))
    x = torch.cat((h_out, d_out), dim=1)
    x = T["m3a"] = _gn_relu(sd, p + ".model3.1", conv(x, *_cw(sd, p + ".model3.0"), stride=1), 4)
    x = T["m3b"] = _gn_relu(sd, p + ".model3.4", conv(x, *_cw(sd, p + ".model3.3"), stride=1), 4)
    x = T["logits"] = conv(x, *_cw(sd, p + ".model3.6"), stride=1)
    return F.softmax(x, dim=-3)


# mynet6_plus.py:316-345
def dsic_dense_warp(h1, cost):
    g2 = torch.zeros_like(h1)
    W = cost.size(-1)
    for d in range(cost.size(-3)):
        g2[:, :, :, 0:W - d] += cost[:, d:d + 1, :, 0:W - d] * h1[:, :, :, d:W]
    return g2


# mynet6_plus.py:675-761
def dsic_forward(sd, x1, x2, K=5, Fn=21, Cn=32, taps=None):
    T = taps if taps is not None else {}
    M = sd["encoder1.g_a_conv4.weight"].shape[0]
    cat = lambda a, b: torch.cat((a, b), dim=-3)
    cv = lambda i, h1, h2, d, s: dsic_cost_volume(sd, f"_cost_volume{i}", h1, h2, d, s, Fn, Cn,
                                                  taps=T.setdefault(f"cv{i}", {}))
    y1, g1_1, g1_2, g1_3 = dsic_encoder1(sd, x1)
    z1_hat, z1_lik = entropy_bottleneck(encode_hyper(sd, y1, "_h_a1"), *eb_params(sd, "entropy_bottleneck1"))
    y1_hat, y1_lik = gmm_conditional(y1, *gmm_hyper_y1(sd, z1_hat, "_h_s1", K, M), K)
    x1_hat, g1_4, g1_5, g1_6 = dsic_decoder1(sd, y1_hat)
    ctx = dsic_global_context(sd, y1_hat, Fn, Cn)

    a1 = _gdn(sd, "pic2_g_a_gdn1", conv(x2, *_cw(sd, "pic2_g_a_conv1")))
    c1 = cv(1, g1_1, a1, ctx[0], 8)
    w1 = dsic_dense_warp(g1_1, c1)
    a2 = _gdn(sd, "pic2_g_a_gdn2", conv(cat(w1, a1), *_cw(sd, "pic2_g_a_conv2")))
    w2 = dsic_dense_warp(g1_2, cv(2, g1_2, a2, ctx[1], 4))
    a3 = _gdn(sd, "pic2_g_a_gdn3", conv(cat(w2, a2), *_cw(sd, "pic2_g_a_conv3")))
    w3 = dsic_dense_warp(g1_3, cv(3, g1_3, a3, ctx[2], 2))
    y2 = conv(cat(w3, a3), *_cw(sd, "pic2_g_a_conv4"))

    z2_hat, z2_lik = entropy_bottleneck(encode_hyper(sd, y2, "_h_a2"), *eb_params(sd, "entropy_bottleneck2"))
    y2_hat, y2_lik = gmm_conditional(y2, *gmm_hyper_y2(sd, z2_hat, y1_hat, "_h_s2", K, M), K)

    s1 = _gdn(sd, "pic2_g_s_gdn1", deconv(y2_hat, *_cw(sd, "pic2_g_s_conv1")), True)
    w4 = dsic_dense_warp(g1_4, cv(4, g1_4, s1, ctx[2], 2))
    s2 = _gdn(sd, "pic2_g_s_gdn2", deconv(cat(w4, s1), *_cw(sd, "pic2_g_s_conv2")), True)
    w5 = dsic_dense_warp(g1_5, cv(5, g1_5, s2, ctx[1], 4))
    s3 = _gdn(sd, "pic2_g_s_gdn3", deconv(cat(w5, s2), *_cw(sd, "pic2_g_s_conv3")), True)
    w6 = dsic_dense_warp(g1_6, cv(6, g1_6, s3, ctx[0], 8))
    x2_hat = deconv(cat(w6, s3), *_cw(sd, "pic2_g_s_conv4"))
    T.update(y1=y1, y1_hat=y1_hat, g1_1=g1_1, g1_2=g1_2, g1_3=g1_3, g1_4=g1_4, g1_5=g1_5, g1_6=g1_6, a1=a1, a2=a2, a3=a3,
             s1=s1, s2=s2, s3=s3, cost1=c1, warp1=w1, warp2=w2, warp3=w3, y2=y2, y2_hat=y2_hat, ctx0=ctx[0], ctx1=ctx[1],
             ctx2=ctx[2])
    return {"x1_hat": x1_hat, "x2_hat": x2_hat,
            "likelihoods": {"y1": y1_lik, "y2": y2_lik, "z1": z1_lik, "z2": z2_lik}}
