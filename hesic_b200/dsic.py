"""DSIC (ywz/DSIC/mynet6_plus.py): the parallax-attention stereo codec that HESIC is compared against
(BASELINE config 5), on the hesic_b200 kernels.

Same class tree, constructor arguments, ``forward`` signatures and ``state_dict`` keys as the reference
file.  ``DSIC.forward`` runs on the fused engine of hesic_b200/dsic_engine.py; the modules below are also callable
on their own at operator level (every convolution a tcgen05 launch through ``functional.conv2d``, NCHW in/out).
GroupNorm+ReLU, the disparity softmax and ``dense_warp`` are the kernels
of csrc/dsic_ops.cu, and the two ``nn.Conv3d`` layers of each cost volume run as ONE 2-D convolution over
the (F0 x C) = 224 stacked channels with a block-banded weight (a 5-tap correlation along the disparity
axis is a banded channel-mixing matrix), i.e. on the same tensor-core path as every other layer.
"""
import torch
import torch.nn as nn

from compressai.entropy_models import GaussianMixtureConditional
from compressai.layers import GDN, ResidualBlock, conv3x3
from compressai.models.utils import conv, deconv

from . import _capi as C
from . import functional as F
from .modules import Conv2d
from .stereo import CompressionModel, Decoder1, Encoder1, encode_hyper, gmm_hyper_y1, gmm_hyper_y2


class _GNReLU(nn.GroupNorm):
    """nn.GroupNorm whose forward also applies the nn.ReLU that follows it in the reference's nn.Sequential
    (the nn.ReLU module stays in the Sequential as a no-op on non-negative input, keeping the key layout)."""

    def forward(self, x):
        C.require_cuda(x)
        return F.group_norm(x, self.num_groups, self.weight, self.bias, self.eps, relu=True)


class Conv3dAs2d(nn.Conv3d):
    """nn.Conv3d(F0, F0, k, padding=k//2) over [B, F0, D, H, W], evaluated as a 2-D convolution over the F0*D
    stacked channels: W2[(f', d'), (f, d), ky, kx] = W3[f', f, d - d' + k//2, ky, kx] (zero outside the band)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._plans = {}      # depth_major -> (plan, key, packed weight, packed bias)

    def _plan_for(self, D, depth_major=False):
        """depth_major: stack the channels as (d, f) instead of (f, d).  The 5-tap band along the depth axis then makes
        the 2-D weight block-banded at the granularity of the tensor-core tiles (an output tile of 128 channels = 18
        depths needs 22 input depths = 3 of the 4 64-channel K chunks), and the kernel skips the empty blocks."""
        key = (self.weight.data_ptr(), self.weight._version, self.bias.data_ptr(), self.bias._version, D,
               self.weight.device.index, bool(depth_major))
        ent = self._plans.get(bool(depth_major))
        if ent is not None and ent[1] == key:
            return ent[0]
        Fo, Fi, kd, kh, kw = self.weight.shape
        w3 = self.weight.detach()
        w2 = torch.zeros((Fo, D, Fi, D, kh, kw), device=w3.device, dtype=torch.float32)
        p = kd // 2
        for dd in range(D):          # output depth d'
            lo, hi = max(0, dd - p), min(D, dd + p + 1)
            w2[:, dd, :, lo:hi] = w3[:, :, lo - dd + p:hi - dd + p]
        plan = F.ConvPlan(Fi * D, Fo * D, (kh, kw), 1, kh // 2)
        if depth_major:
            wp = w2.permute(1, 0, 3, 2, 4, 5).reshape(D * Fo, D * Fi, kh, kw).contiguous()
            bp = self.bias.detach().repeat(D).contiguous()
        else:
            wp = w2.reshape(Fo * D, Fi * D, kh, kw).contiguous()
            bp = self.bias.detach().repeat_interleave(D).contiguous()
        plan.load(wp, bp)
        if depth_major:
            plan.detect_kband(wp)
        self._plans[bool(depth_major)] = (plan, key, wp, bp)
        return plan

    def forward(self, x):
        C.require_cuda(x)
        B, Fi, D, H, W = x.shape
        y = F.conv2d(x.reshape(B, Fi * D, H, W), self._plan_for(D))
        return y.reshape(B, self.out_channels, D, H, W)


class _GN3dReLU(nn.GroupNorm):
    """GroupNorm(num_groups=1, F0) + ReLU on a 5-D [B, F0, D, H, W] tensor: one group over everything, the affine
    parameters per F0 channel (mynet6_plus.py:277-283)."""

    def forward(self, x):
        C.require_cuda(x)
        B, Fc, D, H, W = x.shape
        w = self.weight.detach().repeat_interleave(D)
        b = self.bias.detach().repeat_interleave(D)
        y = F.group_norm(x.reshape(B, Fc * D, H, W), self.num_groups, w, b, self.eps, relu=True)
        return y.reshape(B, Fc, D, H, W)


class global_context(nn.Module):
    """y1_hat -> three [B, 1, F0, C, h, w] context volumes (mynet6_plus.py:216-246)."""

    def __init__(self, M, F_, C_):
        super().__init__()
        self.F, self.F0, self.M, self.C = F_, F_ // 3, M, C_
        FC = F_ * C_
        self.global_net = nn.Sequential(
            conv(M, FC, kernel_size=5, stride=1), _GNReLU(num_channels=FC, num_groups=F_), nn.ReLU(),
            conv(FC, FC, kernel_size=5, stride=1), _GNReLU(num_channels=FC, num_groups=F_), nn.ReLU(),
            conv(FC, FC, kernel_size=5, stride=1), _GNReLU(num_channels=FC, num_groups=F_), nn.ReLU(),
            conv(FC, FC, kernel_size=5, stride=1))

    def forward(self, y1):
        t = self.global_net(y1)
        return torch.reshape(t, (-1, 3, self.F0, self.C, t.size(-2), t.size(-1))).split(1, dim=1)


class cost_volume(nn.Module):
    """(h1, h2, context volume) -> softmax-normalised disparity cost [B, C, H, W] (mynet6_plus.py:249-313)."""

    def __init__(self, N, scale_factor, F_, C_):
        super().__init__()
        self.N, self.scale_factor, self.F, self.F0, self.C = N, scale_factor, F_, F_ // 3, C_
        self.model1 = nn.Sequential(
            conv(2 * N, N, kernel_size=5, stride=1), _GNReLU(num_channels=N, num_groups=4), nn.ReLU(),
            conv(N, N, kernel_size=5, stride=1), _GNReLU(num_channels=N, num_groups=4), nn.ReLU())
        self.upsample_layer = nn.UpsamplingBilinear2d(scale_factor=scale_factor)
        self.model2 = nn.Sequential(
            Conv3dAs2d(self.F0, self.F0, kernel_size=5, stride=1, padding=2), _GN3dReLU(num_channels=self.F0, num_groups=1),
            nn.ReLU(),
            Conv3dAs2d(self.F0, self.F0, kernel_size=5, stride=1, padding=2), _GN3dReLU(num_channels=self.F0, num_groups=1),
            nn.ReLU())
        self.model3 = nn.Sequential(
            conv(self.F0 * C_ + N, N, kernel_size=5, stride=1), _GNReLU(num_channels=N, num_groups=4), nn.ReLU(),
            conv(N, N, kernel_size=5, stride=1), _GNReLU(num_channels=N, num_groups=4), nn.ReLU(),
            conv(N, C_, kernel_size=5, stride=1))

    def forward(self, h1, h2, d):
        C.require_cuda(h1, h2, d)
        h_out = self.model1(torch.cat((h1, h2), dim=1))
        d_in = torch.reshape(d, (-1, d.size(-3), d.size(-2), d.size(-1)))          # [B*F0, C, h, w]
        d_up = F.upsample_bilinear(d_in.contiguous(), self.scale_factor)
        d_up_3d = torch.reshape(d_up, (-1, self.F0, self.C, d_up.size(-2), d_up.size(-1)))
        d_out_3d = self.model2(d_up_3d)
        d_out = torch.reshape(d_out_3d, (-1, self.F0 * self.C, d_out_3d.size(-2), d_out_3d.size(-1)))
        all_out = self.model3(torch.cat((h_out, d_out), dim=1))
        return F.softmax_channels(all_out)


class dense_warp(nn.Module):
    """g2[..., x] = sum_d cost[:, d, :, x] * h1[..., x + d] (mynet6_plus.py:316-345), one kernel instead of a
    C-iteration Python loop."""

    def forward(self, h1, cost):
        C.require_cuda(h1, cost)
        return F.dense_warp(h1, cost)


class DSIC(CompressionModel):
    """mynet6_plus.py:616-761."""

    def __init__(self, N=128, M=192, F=21, C=32, K=5, **kwargs):
        super().__init__(entropy_bottleneck_channels=N, **kwargs)
        self.gaussian1 = GaussianMixtureConditional(K=K)
        self.gaussian2 = GaussianMixtureConditional(K=K)
        self.N, self.M, self.F, self.C, self.K = int(N), int(M), F, C, K
        self.encoder1 = Encoder1(N, M)
        self.decoder1 = Decoder1(N, M)
        self.pic2_g_a_conv1 = conv(3, N)
        self.pic2_g_a_gdn1 = GDN(N)
        self.pic2_g_a_conv2 = conv(2 * N, N)
        self.pic2_g_a_gdn2 = GDN(N)
        self.pic2_g_a_conv3 = conv(2 * N, N)
        self.pic2_g_a_gdn3 = GDN(N)
        self.pic2_g_a_conv4 = conv(2 * N, M)
        self.pic2_g_s_conv1 = deconv(M, N)
        self.pic2_g_s_gdn1 = GDN(N, inverse=True)
        self.pic2_g_s_conv2 = deconv(2 * N, N)
        self.pic2_g_s_gdn2 = GDN(N, inverse=True)
        self.pic2_g_s_conv3 = deconv(2 * N, N)
        self.pic2_g_s_gdn3 = GDN(N, inverse=True)
        self.pic2_g_s_conv4 = deconv(2 * N, 3)
        self._global_context = global_context(M, F, C)
        self._cost_volume1 = cost_volume(N, 8, F, C)
        self._cost_volume2 = cost_volume(N, 4, F, C)
        self._cost_volume3 = cost_volume(N, 2, F, C)
        self._cost_volume4 = cost_volume(N, 2, F, C)
        self._cost_volume5 = cost_volume(N, 4, F, C)
        self._cost_volume6 = cost_volume(N, 8, F, C)
        for i in range(1, 7):
            setattr(self, f"_warp{i}", dense_warp())
        self._h_a1 = encode_hyper(N=N, M=M)
        self._h_a2 = encode_hyper(N=N, M=M)
        self._h_s1 = gmm_hyper_y1(N=N, M=M, K=K)
        self._h_s2 = gmm_hyper_y2(N=N, M=M, K=K)

    @property
    def hesic_engine(self):
        if self.__dict__.get("_engine") is None:
            from .dsic_engine import DsicEngine
            self.__dict__["_engine"] = DsicEngine(self)
        return self.__dict__["_engine"]

    def forward(self, x1, x2):
        """mynet6_plus.py:675-761, eval mode, on the fused engine (hesic_b200/dsic_engine.py); the operator-level
        composition of the same kernels remains available as ``forward_operator_level`` (parity cross-check)."""
        if self.training:
            raise NotImplementedError("hesic_b200: DSIC.forward is the inference path; call .eval() first")
        return self.hesic_engine.forward(x1, x2)

    def _analysis_oplevel(self, x1, x2):
        """Everything up to the quantised latents (mynet6_plus.py:675-723 / 799-845), module by module.  z is passed
        through ``code_z`` (forward: the bottleneck's likelihood path; codec: compress + decompress)."""
        cat = lambda a, b: torch.cat((a, b), dim=-3)
        y1, g1_1, g1_2, g1_3 = self.encoder1(x1)
        z1 = self._h_a1(y1)
        return y1, (g1_1, g1_2, g1_3), z1, cat

    def _view2_analysis_oplevel(self, x2, g1, ctx, cat):
        a1 = self.pic2_g_a_gdn1(self.pic2_g_a_conv1(x2))
        w1 = self._warp1(g1[0], self._cost_volume1(g1[0], a1, ctx[0]))
        a2 = self.pic2_g_a_gdn2(self.pic2_g_a_conv2(cat(w1, a1)))
        w2 = self._warp2(g1[1], self._cost_volume2(g1[1], a2, ctx[1]))
        a3 = self.pic2_g_a_gdn3(self.pic2_g_a_conv3(cat(w2, a2)))
        w3 = self._warp3(g1[2], self._cost_volume3(g1[2], a3, ctx[2]))
        return self.pic2_g_a_conv4(cat(w3, a3))

    def _view2_synthesis_oplevel(self, y2_hat, g1_dec, ctx, cat):
        g1_4, g1_5, g1_6 = g1_dec
        s1 = self.pic2_g_s_gdn1(self.pic2_g_s_conv1(y2_hat))
        w4 = self._warp4(g1_4, self._cost_volume4(g1_4, s1, ctx[2]))
        s2 = self.pic2_g_s_gdn2(self.pic2_g_s_conv2(cat(w4, s1)))
        w5 = self._warp5(g1_5, self._cost_volume5(g1_5, s2, ctx[1]))
        s3 = self.pic2_g_s_gdn3(self.pic2_g_s_conv3(cat(w5, s2)))
        w6 = self._warp6(g1_6, self._cost_volume6(g1_6, s3, ctx[0]))
        return self.pic2_g_s_conv4(cat(w6, s3))

    def forward_operator_level(self, x1, x2):
        if self.training:
            raise NotImplementedError("hesic_b200: DSIC.forward is the inference path; call .eval() first")
        C.require_cuda(x1, x2)
        with torch.no_grad():
            y1, g1, z1, cat = self._analysis_oplevel(x1, x2)
            z1_hat, z1_lik = self.entropy_bottleneck1(z1)
            y1_hat, y1_lik = self.gaussian1(y1, *self._h_s1(z1_hat))
            x1_hat, g1_4, g1_5, g1_6 = self.decoder1(y1_hat)
            ctx = self._global_context(y1_hat)
            y2 = self._view2_analysis_oplevel(x2, g1, ctx, cat)
            z2_hat, z2_lik = self.entropy_bottleneck2(self._h_a2(y2))
            y2_hat, y2_lik = self.gaussian2(y2, *self._h_s2(z2_hat, y1_hat))
            x2_hat = self._view2_synthesis_oplevel(y2_hat, (g1_4, g1_5, g1_6), ctx, cat)
        return {"x1_hat": x1_hat, "x2_hat": x2_hat,
                "likelihoods": {"y1": y1_lik, "y2": y2_lik, "z1": z1_lik, "z2": z2_lik}}

    def _quantize(self, inputs, mode, means=None):
        return self.gaussian1._quantize(inputs, mode, means)

    def compress(self, x1, x2, output_name, output_path=""):
        """mynet6_plus.py:799-1126: same ``.npz`` / ``.bin`` layout and table arithmetic as HSIC.compress;
        the .bin stream is this library's own range coder (``range_coder`` is un-vendored: parity unpinned)."""
        import os
        from .stereo import codec_write
        if self.training:
            raise NotImplementedError("hesic_b200: compress() is an inference path; call .eval() first")
        C.require_cuda(x1, x2)
        if x1.shape[0] != 1:
            raise ValueError("DSIC.compress codes one stereo pair per call, as the reference does")
        with torch.no_grad():
            y1, g1, z1, cat = self._analysis_oplevel(x1, x2)
            z1s = self.entropy_bottleneck1.compress(z1)
            z1_hat = self.entropy_bottleneck1.decompress(z1s, z1.size()[-2:])
            gmm1 = self._h_s1(z1_hat)
            y1_hat = self._quantize(y1, "dequantize", means=None)
            ctx = self._global_context(y1_hat)
            y2 = self._view2_analysis_oplevel(x2, g1, ctx, cat)
            z2 = self._h_a2(y2)
            z2s = self.entropy_bottleneck2.compress(z2)
            z2_hat = self.entropy_bottleneck2.decompress(z2s, z2.size()[-2:])
            gmm2 = self._h_s2(z2_hat, y1_hat)
            y2_hat = self._quantize(y2, "dequantize", means=None)
            out1, out2, delta = codec_write(self, x1.shape[2:], [(y1_hat, gmm1, z1s), (y2_hat, gmm2, z2s)], output_name, output_path)
        num_pixels = x1.shape[2] * x1.shape[3] * 2
        return {"bpp_real": (os.path.getsize(out1) + os.path.getsize(out2)) * 8 / num_pixels, "enctime": delta,
                "y1_hat": y1_hat, "y2_hat": y2_hat, "z1_hat": z1_hat, "z2_hat": z2_hat}

    def decompress(self, device, output_name, output_path=""):
        """mynet6_plus.py:1129-1350."""
        import time
        from .stereo import codec_code_view, codec_read
        if self.training:
            raise NotImplementedError("hesic_b200: decompress() is an inference path; call .eval() first")
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("hesic_b200: DSIC.decompress needs the model on a CUDA device (no CPU fallback)")
        with torch.no_grad():
            size, heads, dec = codec_read(self, output_name, output_path)
            y_shape = [v // 16 for v in size]
            z_shape = [v // 4 for v in y_shape]
            start = time.time()
            z1_hat = self.entropy_bottleneck1.decompress([heads[0][0]], z_shape)
            z2_hat = self.entropy_bottleneck2.decompress([heads[1][0]], z_shape)
            y1_hat = torch.zeros((1, self.M, *y_shape), device=dev)
            y2_hat = torch.zeros((1, self.M, *y_shape), device=dev)
            codec_code_view(self, dec, y1_hat, self._h_s1(z1_hat), heads[0][1], heads[0][2], decode=True)
            x1_hat, g1_4, g1_5, g1_6 = self.decoder1(y1_hat)
            codec_code_view(self, dec, y2_hat, self._h_s2(z2_hat, y1_hat), heads[1][1], heads[1][2], decode=True)
            ctx = self._global_context(y1_hat)
            x2_hat = self._view2_synthesis_oplevel(y2_hat, (g1_4, g1_5, g1_6), ctx, lambda a, b: torch.cat((a, b), dim=-3))
        return {"x1_hat": x1_hat, "x2_hat": x2_hat, "y1_hat": y1_hat, "y2_hat": y2_hat, "z1_hat": z1_hat, "z2_hat": z2_hat,
                "dectime": time.time() - start}


class Enhancement_Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.RB1 = ResidualBlock(32, 32)
        self.RB2 = ResidualBlock(32, 32)
        self.RB3 = ResidualBlock(32, 32)

    def forward(self, x):
        return self.RB3(self.RB2(self.RB1(x))) + x


class Enhancement(nn.Module):
    """mynet6_plus.py:57-78 (no cross-view input, unlike newnet1's)."""

    def __init__(self):
        super().__init__()
        self.conv1 = conv3x3(3, 32)
        self.EB1 = Enhancement_Block()
        self.EB2 = Enhancement_Block()
        self.EB3 = Enhancement_Block()
        self.conv2 = conv3x3(32, 3)

    def forward(self, x):
        return self.conv2(self.EB3(self.EB2(self.EB1(self.conv1(x))))) + x


class Independent_EN(nn.Module):
    def __init__(self):
        super().__init__()
        self.EH1 = Enhancement()
        self.EH2 = Enhancement()

    def forward(self, x1_hat, x2_hat):
        """mynet6_plus.py Independent_EN on the fused enhancement kernels (hesic_b200/enhance.py)."""
        if self.__dict__.get("_engine") is None:
            from .enhance import EnhanceEngine
            self.__dict__["_engine"] = EnhanceEngine(self)
        return self.__dict__["_engine"].forward_mono(x1_hat, x2_hat)


class DSIC_plus(nn.Module):
    """mynet6_plus.py:1352-1370."""

    def __init__(self, N=128, M=192, F=21, C=32, K=5, **kwargs):
        super().__init__()
        self.m1 = DSIC(N=N, M=M, F=F, C=C, K=K)
        self.m2 = Independent_EN()

    def forward(self, x1, x2):
        out1 = self.m1(x1, x2)
        out2 = self.m2(out1["x1_hat"], out1["x2_hat"])
        return {"x1_hat": out2["x1_hat"], "x2_hat": out2["x2_hat"], "likelihoods": out1["likelihoods"]}
