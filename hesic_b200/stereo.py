"""Stereo codec modules with the reference's interface (ywz/mywork/newnet1.py,
newnet1_joint.py, .trash/newnet9.py): same class names, constructor arguments,
attribute / ``state_dict`` names and ``forward`` return values, so the authors'
checkpoints load strictly and the drivers run unchanged.  ``HSIC.forward`` hands
the whole pass to ``HesicEngine``; the sub-modules' own ``forward`` methods work
stand-alone on NCHW fp32 CUDA tensors through the operator-level kernels.

``compressai`` here is the stand-in under ``hesic_b200/compat/site``
(``hesic_b200.compat.install()`` must have run, which the ``newnet*`` modules
in that directory guarantee by construction).
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn

from compressai.entropy_models import EntropyBottleneck, GaussianConditional, GaussianMixtureConditional
from compressai.layers import GDN, MaskedConv2d, ResidualBlock, conv3x3
from compressai.models.utils import conv, deconv

from . import _capi as C
from . import functional as F
from .engine import HesicEngine


class CompressionModel(nn.Module):
    """Two-bottleneck base class of the stereo models (newnet1.py:36-106)."""

    def __init__(self, entropy_bottleneck_channels, init_weights=True):
        super().__init__()
        self.entropy_bottleneck1 = EntropyBottleneck(entropy_bottleneck_channels)
        self.entropy_bottleneck2 = EntropyBottleneck(entropy_bottleneck_channels)
        if init_weights:
            self._initialize_weights()

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def _initialize_weights(self):
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, *args):
        raise NotImplementedError()

    def parameters(self):
        for m in self.children():
            if isinstance(m, EntropyBottleneck):
                continue
            for p in m.parameters():
                yield p

    def aux_parameters(self):
        for m in self.children():
            if isinstance(m, EntropyBottleneck):
                for p in m.parameters():
                    yield p

    def update(self, force=False):
        for m in self.children():
            if isinstance(m, EntropyBottleneck):
                m.update(force=force)


class RateDistortionLoss(nn.Module):
    """newnet1.py:110-129 (single-image form kept for name compatibility)."""

    def __init__(self, lmbda=1e-2):
        super().__init__()
        self.mse = nn.MSELoss()
        self.lmbda = lmbda

    def forward(self, output, target):
        N, _, H, W = target.size()
        num_pixels = N * H * W
        out = {}
        out["bpp_loss"] = sum((torch.log(lk).sum() / (-math.log(2) * num_pixels)) for lk in output["likelihoods"].values())
        out["mse_loss"] = self.mse(output["x_hat"], target)
        out["loss"] = self.lmbda * 255 ** 2 * out["mse_loss"] + out["bpp_loss"]
        return out


class AverageMeter:
    """Running average (newnet1.py:132-144)."""

    def __init__(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class encode_hyper(nn.Module):
    """|y| -> conv s1, ReLU, conv s2, ReLU, conv s2 (newnet1.py:420-437)."""

    def __init__(self, N, M):
        super().__init__()
        self.encode_hyper = nn.Sequential(conv(M, N, kernel_size=5, stride=1), nn.ReLU(), conv(N, N, kernel_size=5),
                                          nn.ReLU(), conv(N, N, kernel_size=5))

    def forward(self, y):
        return self.encode_hyper(torch.abs(y))


class spatial_pool2d(nn.Module):
    """Global spatial maximum per (b, c) (newnet1.py:441-453), one reduction kernel."""

    def forward(self, X):
        C.require_cuda(X)
        return F.spatial_max(X)


def _mix_softmax_head(seq, x, K, M):
    """deconv/conv, LeakyReLU, conv, global max, LeakyReLU, conv1x1, softmax over K."""
    t = seq[2](nn.functional.leaky_relu(seq[0](x)))
    pooled = F.spatial_max(t)
    return F.mixture_weights(pooled, seq[5].weight, seq[5].bias, K, M)


class gmm_hyper_y1(nn.Module):
    """z1_hat -> (sigma, means, weights) of the K-component mixture (newnet1.py:456-514)."""

    def __init__(self, N, M, K):
        super().__init__()
        self.N, self.M, self.K = N, M, K
        self.gmm_sigma = nn.Sequential(deconv(N, N, kernel_size=5), nn.ReLU(), deconv(N, N, kernel_size=5), nn.ReLU(),
                                       conv(N, M * K, kernel_size=5, stride=1), nn.ReLU())
        self.gmm_means = nn.Sequential(deconv(N, N, kernel_size=5), nn.LeakyReLU(), deconv(N, N, kernel_size=5),
                                       nn.LeakyReLU(), conv(N, M * K, kernel_size=5, stride=1))
        self.gmm_weights = nn.Sequential(deconv(N, N, kernel_size=5), nn.LeakyReLU(), deconv(N, M * K, kernel_size=5),
                                         spatial_pool2d(), nn.LeakyReLU(), conv(M * K, M * K, kernel_size=1, stride=1))

    def forward(self, z1):
        return self.gmm_sigma(z1), self.gmm_means(z1), _mix_softmax_head(self.gmm_weights, z1, self.K, self.M)


class gmm_hyper_y2(nn.Module):
    """(z2_hat upsampled x4, y1) -> (sigma, means, weights) (newnet1.py:517-577)."""

    def __init__(self, N, M, K):
        super().__init__()
        self.N, self.M, self.K = N, M, K
        self.upsample_layer = nn.UpsamplingBilinear2d(scale_factor=4)
        self.gmm_sigma = nn.Sequential(conv(N + M, N, kernel_size=5, stride=1), nn.ReLU(),
                                       conv(N, N, kernel_size=5, stride=1), nn.ReLU(),
                                       conv(N, M * K, kernel_size=5, stride=1), nn.ReLU())
        self.gmm_means = nn.Sequential(conv(N + M, N, kernel_size=5, stride=1), nn.LeakyReLU(),
                                       conv(N, N, kernel_size=5, stride=1), nn.LeakyReLU(),
                                       conv(N, M * K, kernel_size=5, stride=1))
        self.gmm_weights = nn.Sequential(conv(N + M, N, kernel_size=5, stride=1), nn.LeakyReLU(),
                                         conv(N, M * K, kernel_size=5, stride=1), spatial_pool2d(), nn.LeakyReLU(),
                                         conv(M * K, M * K, kernel_size=1, stride=1))

    def forward(self, z2, y1):
        C.require_cuda(z2, y1)
        cat_in = torch.cat((F.upsample_bilinear(z2, 4), y1), dim=-3)
        return (self.gmm_sigma(cat_in), self.gmm_means(cat_in),
                _mix_softmax_head(self.gmm_weights, cat_in, self.K, self.M))


class Encoder1(nn.Module):
    """Left-view analysis transform (newnet1.py:580-601)."""

    def __init__(self, N, M, **kwargs):
        super().__init__()
        self.g_a_conv1 = conv(3, N)
        self.g_a_gdn1 = GDN(N)
        self.g_a_conv2 = conv(N, N)
        self.g_a_gdn2 = GDN(N)
        self.g_a_conv3 = conv(N, N)
        self.g_a_gdn3 = GDN(N)
        self.g_a_conv4 = conv(N, M)

    def forward(self, x):
        g1 = self.g_a_gdn1(self.g_a_conv1(x))
        g2 = self.g_a_gdn2(self.g_a_conv2(g1))
        g3 = self.g_a_gdn3(self.g_a_conv3(g2))
        return self.g_a_conv4(g3), g1, g2, g3


class Decoder1(nn.Module):
    """Left-view synthesis transform (newnet1.py:603-624)."""

    def __init__(self, N, M, **kwargs):
        super().__init__()
        self.g_s_conv1 = deconv(M, N)
        self.g_s_gdn1 = GDN(N, inverse=True)
        self.g_s_conv2 = deconv(N, N)
        self.g_s_gdn2 = GDN(N, inverse=True)
        self.g_s_conv3 = deconv(N, N)
        self.g_s_gdn3 = GDN(N, inverse=True)
        self.g_s_conv4 = deconv(N, 3)

    def forward(self, y_hat):
        g1 = self.g_s_gdn1(self.g_s_conv1(y_hat))
        g2 = self.g_s_gdn2(self.g_s_conv2(g1))
        g3 = self.g_s_gdn3(self.g_s_conv3(g2))
        return self.g_s_conv4(g3), g1, g2, g3


class Encoder2(nn.Module):
    """Right-view analysis: cat(x1_warp, x2) -> conv(6->3) -> GDN(3) -> g_a stack (newnet1.py:627-655)."""

    def __init__(self, N, M, **kwargs):
        super().__init__()
        self.pre_conv = conv(6, 3, stride=1)
        self.pre_gdn = GDN(3)
        self.g_a_conv1 = conv(3, N)
        self.g_a_gdn1 = GDN(N)
        self.g_a_conv2 = conv(N, N)
        self.g_a_gdn2 = GDN(N)
        self.g_a_conv3 = conv(N, N)
        self.g_a_gdn3 = GDN(N)
        self.g_a_conv4 = conv(N, M)

    def forward(self, x1_warp, x2):
        x = self.pre_gdn(self.pre_conv(torch.cat((x1_warp, x2), dim=-3)))
        x = self.g_a_gdn1(self.g_a_conv1(x))
        x = self.g_a_gdn2(self.g_a_conv2(x))
        x = self.g_a_gdn3(self.g_a_conv3(x))
        return self.g_a_conv4(x)


class Decoder2(nn.Module):
    """Right-view synthesis: g_s stack -> IGDN(3) -> cat(., x1_hat_warp) -> deconv(6->3, s1)
    (newnet1.py:657-692)."""

    def __init__(self, N, M, **kwargs):
        super().__init__()
        self.g_s_conv1 = deconv(M, N)
        self.g_s_gdn1 = GDN(N, inverse=True)
        self.g_s_conv2 = deconv(N, N)
        self.g_s_gdn2 = GDN(N, inverse=True)
        self.g_s_conv3 = deconv(N, N)
        self.g_s_gdn3 = GDN(N, inverse=True)
        self.g_s_conv4 = deconv(N, 3)
        self.after_gdn = GDN(3, inverse=True)
        self.after_conv = deconv(6, 3, stride=1)

    def forward(self, y_hat, x1_hat_warp):
        x = self.g_s_gdn1(self.g_s_conv1(y_hat))
        x = self.g_s_gdn2(self.g_s_conv2(x))
        x = self.g_s_gdn3(self.g_s_conv3(x))
        x = self.after_gdn(self.g_s_conv4(x))
        return self.after_conv(torch.cat((x, x1_hat_warp), dim=-3))


# ---------------------------------------------------------------------------------------------
# file-codec helpers shared by HSIC and DSIC (newnet1.py:870-1066,1069-1273; mynet6_plus.py:799-1350)
def codec_code_view(model, coder, y_hat, gmm, minmax, channels, decode):
    """Range-code (or decode into) the non-zero channels of one view: the per-element cumulative rows come from the
    device (hesic_gmm_cdf_tables), the coder runs on the host (newnet1.py:934-985 / 1135-1181)."""
    if len(channels) == 0:
        return
    sigma, means, weights = gmm
    tables = F.gmm_cdf_tables(sigma, means, weights, model.K, channels, minmax, model.gaussian1._scale_bound_value())
    tables = tables.cpu().numpy()
    Hy, Wy = y_hat.shape[-2:]
    ch = torch.as_tensor(np.asarray(channels, dtype=np.int64), device=y_hat.device)
    if decode:
        sym = torch.from_numpy(coder.decode(tables).astype(np.float32) - minmax).reshape(len(channels), Hy, Wy)
        y_hat[0, ch] = sym.to(y_hat.device)
    else:
        coder.push((y_hat[0, ch] + minmax).to(torch.int32).reshape(-1).cpu().numpy(), tables)


def codec_write(model, shape_hw, views, output_name, output_path):
    """views = [(y_hat, gmm, z_strings)] * 2 -> ``<name>.npz`` (header, newnet1.py:877-906) and ``<name>.bin`` (range-coded
    y1 then y2); returns (path1, path2, seconds spent entropy coding)."""
    import time
    out1 = os.path.join(output_path, str(output_name) + ".npz")
    out2 = os.path.join(output_path, str(output_name) + ".bin")
    heads = []
    with open(out1, "wb") as f:
        f.write(np.array(shape_hw, dtype=np.uint16).tobytes())
        for y_hat, _, zs in views:
            yi = y_hat[0].to(torch.int64)
            flag = (yi.abs().sum(dim=(1, 2)) > 0).cpu().numpy().astype(np.uint8)
            minmax = int(max(int(yi.max().abs()), int(yi.min().abs()), 1))
            if len(zs[0]) > 65535 or minmax > 32767:
                raise ValueError("compress: z string or symbol range exceeds the reference's uint16 / table header")
            f.write(np.array([len(zs[0]), minmax], dtype=np.uint16).tobytes())
            f.write(np.packbits(flag).tobytes())
            f.write(zs[0])
            heads.append((np.flatnonzero(flag), minmax))
    start = time.time()
    enc = F.RangeEncoderHandle()
    for (y_hat, gmm, _), (channels, minmax) in zip(views, heads):
        codec_code_view(model, enc, y_hat, gmm, minmax, channels, decode=False)
    with open(out2, "wb") as f:
        f.write(enc.finish())
    return out1, out2, time.time() - start


def codec_read(model, output_name, output_path):
    """-> image size (H, W), [(z string, minmax, non-zero channels)] * 2, range decoder over ``<name>.bin``."""
    with open(os.path.join(output_path, str(output_name) + ".npz"), "rb") as f:
        x_shape = np.frombuffer(f.read(4), dtype=np.uint16)
        heads = []
        for _ in range(2):
            length, minmax = (int(v) for v in np.frombuffer(f.read(4), dtype=np.uint16))
            flag = np.unpackbits(np.frombuffer(f.read((model.M + 7) // 8), dtype=np.uint8))[:model.M]
            heads.append((f.read(length), minmax, np.flatnonzero(flag)))
    with open(os.path.join(output_path, str(output_name) + ".bin"), "rb") as f:
        dec = F.RangeDecoderHandle(f.read())
    return (int(x_shape[0]), int(x_shape[1])), heads, dec


class _HSICBase(CompressionModel):
    _variant = "newnet1"

    def __init__(self, N=128, M=192, K=5, **kwargs):
        super().__init__(entropy_bottleneck_channels=N, **kwargs)
        self.gaussian1 = GaussianMixtureConditional(K=K)
        self.gaussian2 = GaussianMixtureConditional(K=K)
        self.N, self.M, self.K = int(N), int(M), int(K)
        self.encoder1 = Encoder1(N, M)
        self.encoder2 = Encoder2(N, M)
        self.decoder1 = Decoder1(N, M)
        self.decoder2 = Decoder2(N, M)
        self._build_hyper(N, M, K)
        object.__setattr__(self, "_engine", None)

    def _build_hyper(self, N, M, K):
        self._h_a1 = encode_hyper(N=N, M=M)
        self._h_a2 = encode_hyper(N=N, M=M)
        self._h_s1 = gmm_hyper_y1(N=N, M=M, K=K)
        self._h_s2 = gmm_hyper_y2(N=N, M=M, K=K)

    @property
    def hesic_engine(self):
        if self._engine is None:
            object.__setattr__(self, "_engine", HesicEngine(self, self._variant))
        return self._engine

    def forward(self, x1, x2, h_matrix):
        """The stereo forward pass (newnet1.py:724-783), eval mode, on the B200 kernels."""
        if self.training:
            raise NotImplementedError("hesic_b200: HSIC.forward is the inference path; call .eval() first")
        return self.hesic_engine.forward(x1, x2, h_matrix)

    # EntropyModel helpers the reference also hangs on the model (newnet1.py:786-821)
    def _quantize(self, inputs, mode, means=None):
        return self.gaussian1._quantize(inputs, mode, means)

    def _standardized_cumulative(self, inputs):
        return 0.5 * torch.erfc(float(-(2 ** -0.5)) * inputs)

    # ---- file codec (newnet1.py:823-1273) -------------------------------------------------
    def _codec_front(self, x1, x2, h_matrix):
        """Everything compress() computes before entropy coding (newnet1.py:825-868), at operator level."""
        if self.training:
            raise NotImplementedError("hesic_b200: compress() is an inference path; call .eval() first")
        C.require_cuda(x1, x2, h_matrix)
        if x1.shape[0] != 1:
            raise ValueError("HSIC.compress codes one stereo pair per call, as the reference does (newnet1.py:945)")
        size = (x1.size(-2), x1.size(-1))
        y1 = self.encoder1(x1)[0]
        z1 = self._h_a1(y1)
        z1_strings = self.entropy_bottleneck1.compress(z1)
        z1_hat = self.entropy_bottleneck1.decompress(z1_strings, z1.size()[-2:])
        gmm1 = self._h_s1(z1_hat)
        y1_hat = self._quantize(y1, "dequantize", means=None)
        x1_hat = self.decoder1(y1_hat)[0]
        x1_warp = F.warp_perspective(x1, h_matrix, size)
        y2 = self.encoder2(x1_warp, x2)
        z2 = self._h_a2(y2)
        z2_strings = self.entropy_bottleneck2.compress(z2)
        z2_hat = self.entropy_bottleneck2.decompress(z2_strings, z2.size()[-2:])
        gmm2 = self._h_s2(z2_hat, self._codec_condition(x1_hat, y1_hat, h_matrix, size))
        y2_hat = self._quantize(y2, "dequantize", means=None)
        return (y1_hat, gmm1, z1_hat, z1_strings), (y2_hat, gmm2, z2_hat, z2_strings), x1_hat

    def _codec_condition(self, x1_hat, y1_hat, h_matrix, size):
        """What conditions the right view's mixture: the re-encoded warped reconstruction ("twiceLeft",
        newnet1.py:857-861), or y1_hat itself in the variants without it (.trash/newnet9.py, codec-test/newnet1.py)."""
        if self._variant == "newnet9":
            return y1_hat
        x1_warp_aftercodec = F.warp_perspective(x1_hat, h_matrix, size)
        return self.gaussian1._quantize(self.encoder1(x1_warp_aftercodec)[0], "dequantize")

    def compress(self, x1, x2, h_matrix, output_name, output_path="", device="cpu"):
        """newnet1.py:823-1066: writes ``<name>.npz`` (sizes, z strings, non-zero-channel flags, symbol range) and
        ``<name>.bin`` (range-coded y1 then y2).  File layout as in the reference; the .bin byte stream is this
        library's own range coder (the reference's ``range_coder`` package is un-vendored and un-pinned)."""
        if self._variant == "joint":
            return self._compress_joint(x1, x2, h_matrix, output_name, output_path)
        (y1_hat, gmm1, z1_hat, z1s), (y2_hat, gmm2, z2_hat, z2s), _ = self._codec_front(x1, x2, h_matrix)
        out1, out2, delta = codec_write(self, x1.shape[2:], [(y1_hat, gmm1, z1s), (y2_hat, gmm2, z2s)], output_name, output_path)
        num_pixels = x1.shape[2] * x1.shape[3] * 2
        return {"bpp_real": (os.path.getsize(out1) + os.path.getsize(out2)) * 8 / num_pixels, "enctime": delta,
                "bpp_side": os.path.getsize(out1) * 8 / num_pixels,
                "y1_hat": y1_hat, "y2_hat": y2_hat, "z1_hat": z1_hat, "z2_hat": z2_hat}

    def decompress(self, x1, x2, h_matrix, output_name, output_path="", device="cpu"):
        """newnet1.py:1069-1273 (x1 / x2 are only consulted for the device, as in the reference)."""
        import time
        if self._variant == "joint":
            return self._decompress_joint(x1, h_matrix, output_name, output_path)
        C.require_cuda(x1, h_matrix)
        dev = x1.device
        size, heads, dec = codec_read(self, output_name, output_path)
        y_shape = [v // 16 for v in size]
        z_shape = [v // 4 for v in y_shape]
        start = time.time()
        z1_hat = self.entropy_bottleneck1.decompress([heads[0][0]], z_shape)
        z2_hat = self.entropy_bottleneck2.decompress([heads[1][0]], z_shape)
        y1_hat = torch.zeros((1, self.M, *y_shape), device=dev)
        y2_hat = torch.zeros((1, self.M, *y_shape), device=dev)
        codec_code_view(self, dec, y1_hat, self._h_s1(z1_hat), heads[0][1], heads[0][2], decode=True)
        x1_hat = self.decoder1(y1_hat)[0]
        gmm2 = self._h_s2(z2_hat, self._codec_condition(x1_hat, y1_hat, h_matrix, size))
        codec_code_view(self, dec, y2_hat, gmm2, heads[1][1], heads[1][2], decode=True)
        x1_hat_warp = F.warp_perspective(x1_hat, h_matrix, size)
        x2_hat = self.decoder2(y2_hat, x1_hat_warp)
        return {"x1_hat": x1_hat, "x2_hat": x2_hat, "y1_hat": y1_hat, "y2_hat": y2_hat, "z1_hat": z1_hat, "z2_hat": z2_hat,
                "dectime": time.time() - start}


    # ---- HESIC+ file codec: autoregressive over the latent positions (newnet1_joint.py:793-1321) ---------------------
    def _joint_position_params(self, view, y_pad, params, extra, h, w):
        """Gaussian parameters of all channels at latent position (h, w) from what is already known there: the causal
        5x5 neighbourhood of y_hat through the masked context model, the hyper-decoder output and (view 2) the warped
        left latent (newnet1_joint.py:899-906 / 988-995).  The encoder and the decoder both call exactly this routine, so
        they derive identical tables."""
        ctxp = getattr(self, f"context_prediction{view}")
        ep = getattr(self, f"entropy_parameters{view}")
        ctx = ctxp(y_pad[:, :, h:h + 5, w:w + 5].contiguous())[:, :, 2:3, 2:3]
        parts = [params[:, :, h:h + 1, w:w + 1], ctx]
        if extra is not None:
            parts.append(extra[:, :, h:h + 1, w:w + 1])
        scales, means = ep(torch.cat(parts, dim=1).contiguous()).chunk(2, 1)
        return scales.contiguous(), means.contiguous()

    def _joint_code_view(self, view, coder, y_hat, params, extra, minmax, channels, decode):
        """Raster scan over the positions, the non-zero channels of a position in channel order (newnet1_joint.py:898-963).
        Encoding gathers the per-position parameters on the device first and codes everything with one table launch;
        decoding is serial by nature: one table launch, one D2H copy and one host decode per position."""
        if len(channels) == 0:
            return
        dev = y_hat.device
        Hy, Wy = y_hat.shape[-2:]
        ones = torch.ones(self.M, device=dev)
        ch = torch.as_tensor(np.asarray(channels, dtype=np.int64), device=dev)
        bound = self.gaussian1._scale_bound_value()
        y_pad = torch.zeros((1, self.M, Hy + 4, Wy + 4), device=dev)
        if not decode:
            y_pad[:, :, 2:2 + Hy, 2:2 + Wy] = y_hat
            sc, mu = torch.empty((1, self.M, Hy, Wy), device=dev), torch.empty((1, self.M, Hy, Wy), device=dev)
            for h in range(Hy):
                for w in range(Wy):
                    s_hw, m_hw = self._joint_position_params(view, y_pad, params, extra, h, w)
                    sc[:, :, h, w], mu[:, :, h, w] = s_hw[:, :, 0, 0], m_hw[:, :, 0, 0]
            tables = F.gmm_cdf_tables(sc, mu, ones, 1, channels, minmax, bound).cpu().numpy()
            tables = tables.reshape(len(channels), Hy * Wy, -1).transpose(1, 0, 2).reshape(len(channels) * Hy * Wy, -1)
            sym = (y_hat[0, ch] + minmax).to(torch.int32).reshape(len(channels), Hy * Wy).t().reshape(-1).cpu().numpy()
            coder.push(sym, np.ascontiguousarray(tables))
            return
        for h in range(Hy):
            for w in range(Wy):
                s_hw, m_hw = self._joint_position_params(view, y_pad, params, extra, h, w)
                tables = F.gmm_cdf_tables(s_hw, m_hw, ones, 1, channels, minmax, bound).cpu().numpy()
                sym = torch.from_numpy(coder.decode(tables).astype(np.float32) - minmax).to(dev)
                y_pad[0, ch, h + 2, w + 2] = sym
        y_hat.copy_(y_pad[:, :, 2:2 + Hy, 2:2 + Wy])

    def _joint_front(self, x1, x2, h_matrix):
        size = (x1.size(-2), x1.size(-1))
        y1 = self.encoder1(x1)[0]
        z1 = self.h_a1(y1)
        z1s = self.entropy_bottleneck1.compress(z1)
        z1_hat = self.entropy_bottleneck1.decompress(z1s, z1.size()[-2:])
        y1_hat = self._quantize(y1, "dequantize", means=None)
        x1_hat = self.decoder1(y1_hat)[0]
        y2 = self.encoder2(F.warp_perspective(x1, h_matrix, size), x2)
        z2 = self.h_a2(y2)
        z2s = self.entropy_bottleneck2.compress(z2)
        z2_hat = self.entropy_bottleneck2.decompress(z2s, z2.size()[-2:])
        y2_hat = self._quantize(y2, "dequantize", means=None)
        return (y1_hat, z1_hat, z1s), (y2_hat, z2_hat, z2s), x1_hat, size

    def _compress_joint(self, x1, x2, h_matrix, output_name, output_path):
        import time
        if self.training:
            raise NotImplementedError("hesic_b200: compress() is an inference path; call .eval() first")
        C.require_cuda(x1, x2, h_matrix)
        if x1.shape[0] != 1:
            raise ValueError("HSIC.compress codes one stereo pair per call, as the reference does")
        (y1_hat, z1_hat, z1s), (y2_hat, z2_hat, z2s), x1_hat, size = self._joint_front(x1, x2, h_matrix)
        out1 = os.path.join(output_path, str(output_name) + ".npz")
        out2 = os.path.join(output_path, str(output_name) + ".bin")
        heads = []
        with open(out1, "wb") as f:
            f.write(np.array(x1.shape[2:], dtype=np.uint16).tobytes())
            for y_hat, zs in ((y1_hat, z1s), (y2_hat, z2s)):
                yi = y_hat[0].to(torch.int64)
                flag = (yi.abs().sum(dim=(1, 2)) > 0).cpu().numpy().astype(np.uint8)
                minmax = int(max(int(yi.max().abs()), int(yi.min().abs()), 1))
                if len(zs[0]) > 65535 or minmax > 32767:
                    raise ValueError("compress: z string or symbol range exceeds the reference's uint16 / table header")
                f.write(np.array([len(zs[0]), minmax], dtype=np.uint16).tobytes())
                f.write(np.packbits(flag).tobytes())
                f.write(zs[0])
                heads.append((np.flatnonzero(flag), minmax))
        start = time.time()
        enc = F.RangeEncoderHandle()
        self._joint_code_view(1, enc, y1_hat, self.h_s1(z1_hat), None, heads[0][1], heads[0][0], decode=False)
        y1_hat_warpf2 = self.gaussian1._quantize(self.encoder1(F.warp_perspective(x1_hat, h_matrix, size))[0], "dequantize")
        self._joint_code_view(2, enc, y2_hat, self.h_s2(z2_hat), y1_hat_warpf2, heads[1][1], heads[1][0], decode=False)
        with open(out2, "wb") as f:
            f.write(enc.finish())
        num_pixels = x1.shape[2] * x1.shape[3] * 2
        return {"bpp_real": (os.path.getsize(out1) + os.path.getsize(out2)) * 8 / num_pixels, "enctime": time.time() - start,
                "y1_hat": y1_hat, "y2_hat": y2_hat, "z1_hat": z1_hat, "z2_hat": z2_hat}

    def _decompress_joint(self, x1, h_matrix, output_name, output_path):
        import time
        C.require_cuda(x1, h_matrix)
        dev = x1.device
        size, heads, dec = codec_read(self, output_name, output_path)
        y_shape = [v // 16 for v in size]
        z_shape = [v // 4 for v in y_shape]
        start = time.time()
        z1_hat = self.entropy_bottleneck1.decompress([heads[0][0]], z_shape)
        z2_hat = self.entropy_bottleneck2.decompress([heads[1][0]], z_shape)
        y1_hat = torch.zeros((1, self.M, *y_shape), device=dev)
        y2_hat = torch.zeros((1, self.M, *y_shape), device=dev)
        self._joint_code_view(1, dec, y1_hat, self.h_s1(z1_hat), None, heads[0][1], heads[0][2], decode=True)
        x1_hat = self.decoder1(y1_hat)[0]
        x1_hat_warp = F.warp_perspective(x1_hat, h_matrix, size)
        y1_hat_warpf2 = self.gaussian1._quantize(self.encoder1(x1_hat_warp)[0], "dequantize")
        self._joint_code_view(2, dec, y2_hat, self.h_s2(z2_hat), y1_hat_warpf2, heads[1][1], heads[1][2], decode=True)
        x2_hat = self.decoder2(y2_hat, x1_hat_warp)
        return {"x1_hat": x1_hat, "x2_hat": x2_hat, "y1_hat": y1_hat, "y2_hat": y2_hat, "z1_hat": z1_hat, "z2_hat": z2_hat,
                "dectime": time.time() - start}


class HSIC(_HSICBase):
    """HESIC (ywz/mywork/newnet1.py:698-783), with the "twiceLeft" re-encode of the warped x1_hat."""
    _variant = "newnet1"


class HSIC_NoTwiceLeft(_HSICBase):
    """ywz/mywork/.trash/newnet9.py:620-661 -- what test3real.py imports: view-2 mixture conditioned on
    y1_hat directly, and no y*_hat in the returned dict."""
    _variant = "newnet9"


class HSIC_Joint(_HSICBase):
    """HESIC+ (ywz/mywork/newnet1_joint.py:586-753): mean-scale hyperprior + masked-conv context model,
    with the warped left latent concatenated into the right view's entropy parameters."""
    _variant = "joint"

    def _build_hyper(self, N, M, K):
        def h_a():
            return nn.Sequential(conv(M, N, stride=1, kernel_size=3), nn.LeakyReLU(inplace=True),
                                 conv(N, N, stride=2, kernel_size=5), nn.LeakyReLU(inplace=True),
                                 conv(N, N, stride=2, kernel_size=5))

        def h_s():
            return nn.Sequential(deconv(N, M, stride=2, kernel_size=5), nn.LeakyReLU(inplace=True),
                                 deconv(M, M * 3 // 2, stride=2, kernel_size=5), nn.LeakyReLU(inplace=True),
                                 conv(M * 3 // 2, M * 2, stride=1, kernel_size=3))

        def ep(cin):
            from hesic_b200.modules import Conv2d
            return nn.Sequential(Conv2d(cin, M * 10 // 3, 1), nn.LeakyReLU(inplace=True),
                                 Conv2d(M * 10 // 3, M * 8 // 3, 1), nn.LeakyReLU(inplace=True),
                                 Conv2d(M * 8 // 3, M * 6 // 3, 1))

        self.h_a1 = h_a()
        self.h_s1 = h_s()
        self.entropy_parameters1 = ep(M * 12 // 3)
        self.context_prediction1 = MaskedConv2d(M, 2 * M, kernel_size=5, padding=2, stride=1)
        self.gaussian_conditional1 = GaussianConditional(None)
        self.h_a2 = h_a()
        self.h_s2 = h_s()
        self.entropy_parameters2 = ep(5 * M)
        self.context_prediction2 = MaskedConv2d(M, 2 * M, kernel_size=5, padding=2, stride=1)
        self.gaussian_conditional2 = GaussianConditional(None)


# ---------------------------------------------------------------------------------------------
# Independent_EN: cross-quality enhancement that follows HSIC in test3real.py:186
# (newnet1.py:272-311, 1278-1300) -- SURVEY.md 8f rank 1.
class Enhancement_Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.RB1 = ResidualBlock(32, 32)
        self.RB2 = ResidualBlock(32, 32)
        self.RB3 = ResidualBlock(32, 32)

    def forward(self, x):
        return self.RB3(self.RB2(self.RB1(x))) + x


class Enhancement(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = conv3x3(6, 32)
        self.EB1 = Enhancement_Block()
        self.EB2 = Enhancement_Block()
        self.EB3 = Enhancement_Block()
        self.conv2 = conv3x3(32, 3)

    def forward(self, x, x_another_warp):
        out = self.conv1(torch.cat((x, x_another_warp), dim=-3))
        out = self.EB3(self.EB2(self.EB1(out)))
        return self.conv2(out) + x


class Independent_EN(nn.Module):
    def __init__(self):
        super().__init__()
        self.EH1 = Enhancement()
        self.EH2 = Enhancement()

    @property
    def hesic_engine(self):
        if self.__dict__.get("_engine") is None:
            from .enhance import EnhanceEngine
            self.__dict__["_engine"] = EnhanceEngine(self)
        return self.__dict__["_engine"]

    def forward(self, x1_hat, x2_hat, h_matrix):
        """newnet1.py:1286-1300 on the fused enhancement kernels (hesic_b200/enhance.py); the Enhancement /
        ResidualBlock sub-modules remain callable on their own (operator level)."""
        return self.hesic_engine.forward(x1_hat, x2_hat, h_matrix)


class GMM_together(nn.Module):
    """HSIC followed by Independent_EN (newnet1.py:1304-1321)."""

    def __init__(self, N=128, M=192, K=5, **kwargs):
        super().__init__()
        self.m1 = HSIC(N, M, K)
        self.m2 = Independent_EN()

    def forward(self, x1, x2, h):
        out1 = self.m1(x1, x2, h)
        out2 = self.m2(out1["x1_hat"], out1["x2_hat"], h)
        return {"x1_hat": out2["x1_hat"], "x2_hat": out2["x2_hat"], "likelihoods": out1["likelihoods"]}
