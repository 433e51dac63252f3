"""Whole-forward orchestration of DSIC (ywz/DSIC/mynet6_plus.py:675-761; BASELINE config 5).

Host glue only, in the style of ``engine.HesicEngine`` (whose helpers it reuses for the hyper path): activations
stay channels-last between kernels -- bf16 (hi, lo) SPLIT planes into every tensor-core convolution, NHWC fp32 out
of a convolution that feeds GroupNorm / softmax / the entropy model -- and no ``torch.cat`` or layout conversion is
materialised.  Per pyramid level one 384-channel SPLIT buffer ``[w | a | g]`` holds

    g  (slots 256..383)  the view-1 feature of that level (encoder1 / decoder1 output after GDN),
    a  (slots 128..255)  the view-2 feature of that level (pic2_* conv + GDN),
    w  (slots   0..127)  dense_warp(g, cost),

so that ``cat(h1, h2)`` of cost_volume.model1 is the channel slice 128..383 (weights loaded with the two input
halves swapped), ``cat(w, a)`` of the next pic2_* layer is the slice 0..255, and dense_warp reads g and writes w in
place.  GDN / IGDN are fused into the producing convolution; GroupNorm(+ReLU) reads the conv's NHWC fp32 output and
writes the next conv's SPLIT input (directly into the 352-channel ``cat(h_out, d_out)`` buffer of model3 where
needed); the two nn.Conv3d of a cost volume run as one 224 -> 224 2-D convolution with a block-banded weight
(dsic.Conv3dAs2d).  ~230 kernels per forward instead of 379 + torch glue.
"""
import torch

from . import _capi as C
from . import functional as F
from .engine import CapturedForward, HesicEngine

_lib = C.lib


class DsicEngine(HesicEngine):
    def __init__(self, model, path=C.PATH_AUTO):
        super().__init__(model, "newnet9", True, path)   # hyper path = the no-twiceLeft HESIC one (y1_hat conditions view 2)
        self._perm_plans = {}

    # ---- helpers -------------------------------------------------------------------------
    def _conv(self, plan, x_desc, B, H, W, kind, act=C.ACT_NONE, dst=None):
        """Run a loaded ConvPlan; returns (tensor, descriptor, Ho, Wo).  dst = (tensor, c0): channel slice to write."""
        Ho, Wo = plan.out_hw(H, W)
        Cout = plan.geom[1]
        if dst is not None:
            t, c0 = dst
        else:
            c0 = 0
            t = {"split": self._split, "nhwc": self._nhwc}[kind](B, Ho, Wo, Cout) if kind != "nchw" else self._nchw(B, Cout, Ho, Wo)
        d = {"split": C.split, "nhwc": C.nhwc, "nchw": C.nchw}[kind](t, Cout, c0)
        plan.run(x_desc, d, act, self.path)
        return t, d, Ho, Wo

    def _perm_plan(self, conv_mod, tag, in_perm=None, out_perm=None):
        """Plan of ``conv_mod`` with its input and / or output channels permuted (new channel i = old channel perm[i]):
        the engine's buffers store some concatenations in a different channel order than the reference's tensors."""
        key = (id(conv_mod), tag)
        ent = self._perm_plans.get(key)
        ver = F.ConvPlan._ver(conv_mod.weight, conv_mod.bias)
        if ent is None or ent[1] != ver:
            w, b = conv_mod.weight.detach(), conv_mod.bias.detach() if conv_mod.bias is not None else None
            if in_perm is not None:
                w = w.index_select(1, torch.as_tensor(in_perm, device=w.device))
            if out_perm is not None:
                op = torch.as_tensor(out_perm, device=w.device)
                w = w.index_select(0, op)
                b = b.index_select(0, op) if b is not None else None
            w = w.contiguous()
            b = b.contiguous() if b is not None else None
            plan = ent[0] if ent else F.ConvPlan(conv_mod.in_channels, conv_mod.out_channels, tuple(conv_mod.kernel_size),
                                                 conv_mod.stride[0], conv_mod.padding[0])
            plan._key = None
            plan.load(w, b)
            plan.set_gdn(None, None, False)
            ent = self._perm_plans[key] = (plan, ver, w, b)
        return ent[0]

    def _swapped_plan(self, conv_mod):
        """cost_volume.model1[0] (conv 2N -> N over cat(h1, h2)) with its two input halves swapped: the level buffer
        stores [.. | h2 | h1]."""
        half = conv_mod.in_channels // 2
        return self._perm_plan(conv_mod, "swap", in_perm=list(range(half, 2 * half)) + list(range(half)))

    @staticmethod
    def _depth_major(F0, D, blocks=1, lead=0):
        """Permutation taking (f, d)-ordered channel blocks to (d, f) order, after ``lead`` untouched channels."""
        perm = list(range(lead))
        for k in range(blocks):
            perm += [lead + k * F0 * D + f * D + d for d in range(D) for f in range(F0)]
        return perm

    def _gn(self, gn_mod, x_t, dst_t, c0, Cn, weight=None, bias=None):
        """GroupNorm + ReLU: NHWC fp32 tensor -> channel slice [c0, c0 + Cn) of a SPLIT buffer."""
        w = gn_mod.weight.detach() if weight is None else weight
        b = gn_mod.bias.detach() if bias is None else bias
        C.check(_lib.hesic_group_norm(C.ref(C.nhwc(x_t)), C.ref(C.split(dst_t, Cn, c0)), gn_mod.num_groups, C.ptr(w), C.ptr(b),
                                      float(gn_mod.eps), 1, C.stream()))

    def _conv_gn(self, conv_or_plan, gn_mod, x_desc, B, H, W, dst=None, weight=None, bias=None):
        """conv -> GroupNorm -> ReLU; returns the SPLIT descriptor of the result (a fresh buffer unless dst is given)."""
        plan = conv_or_plan if isinstance(conv_or_plan, F.ConvPlan) else self._plan(conv_or_plan)
        Ho, Wo = plan.out_hw(H, W)
        Cn = plan.geom[1]
        G = gn_mod.num_groups
        t = self._nhwc(B, Ho, Wo, Cn)
        # the conv's epilogue also accumulates the GroupNorm statistics: no separate pass over its output
        stats = self._keep(torch.empty((B, G, C.GN_SLOTS, 2), device=self.dev, dtype=torch.float64))
        C.check(_lib.hesic_conv_forward_gn(plan.h, C.ref(x_desc), C.ref(C.nhwc(t)), self.path, C.ptr(stats), G, C.stream()))
        if dst is None:
            dst = (self._split(B, Ho, Wo, Cn), 0)
        w = gn_mod.weight.detach() if weight is None else weight
        b = gn_mod.bias.detach() if bias is None else bias
        C.check(_lib.hesic_group_norm_apply(C.ref(C.nhwc(t)), C.ref(C.split(dst[0], Cn, dst[1])), G, C.ptr(w), C.ptr(b),
                                            float(gn_mod.eps), 1, C.ptr(stats), C.stream()))
        return C.split(dst[0], Cn, dst[1])

    def _cost_volume(self, cv, lvl, ctx_t, j, B, H, W):
        """cost_volume.forward (mynet6_plus.py:292-313) on level buffer ``lvl`` ([w | a | g], 384 slots), context
        volume j (channels 224j .. 224j+223 of the NHWC fp32 global-context tensor); followed by dense_warp -> w."""
        N, FC = cv.N, cv.F0 * cv.C
        m1, m2, m3 = cv.model1, cv.model2, cv.model3
        cat3 = self._split(B, H, W, N + FC)                      # cat(h_out, d_out)  mynet6_plus.py:308
        # model1 on cat(h1, h2) = slots [2N, 3N) ++ [N, 2N): read as the slice [N, 3N) with swapped weight halves
        u = self._conv_gn(self._swapped_plan(m1[0]), m1[1], C.split(lvl, 2 * N, N), B, H, W)
        self._conv_gn(m1[3], m1[4], u, B, H, W, dst=(cat3, 0))
        # context volume: bilinear upsample (align_corners=True), two Conv3d as block-banded 2-D convs over the
        # (depth, feature)-ordered channels (the global-context conv already emits that order, see forward()),
        # GroupNorm(1 group) with its per-feature affine repeated over the depths
        D, F0 = cv.C, cv.F0
        up = self._split(B, H, W, FC)
        C.check(_lib.hesic_upsample_bilinear(C.ref(C.nhwc(ctx_t, FC, FC * j)), C.ref(C.split(up)), cv.scale_factor, C.stream()))
        g1, g2 = m2[1], m2[4]
        exp = lambda v: self._keep(v.detach().repeat(D).contiguous())
        v = self._conv_gn(m2[0]._plan_for(D, True), g1, C.split(up), B, H, W, weight=exp(g1.weight), bias=exp(g1.bias))
        self._conv_gn(m2[3]._plan_for(D, True), g2, v, B, H, W, dst=(cat3, N), weight=exp(g2.weight), bias=exp(g2.bias))
        # model3 -> softmax over the disparities
        v = self._conv_gn(self._perm_plan(m3[0], "dm", in_perm=self._depth_major(F0, D, 1, N)), m3[1], C.split(cat3), B, H, W)
        v = self._conv_gn(m3[3], m3[4], v, B, H, W)
        raw, raw_d, _, _ = self._conv(self._plan(m3[6]), v, B, H, W, "nhwc")
        cost = self._nhwc(B, H, W, cv.C)
        C.check(_lib.hesic_softmax_channels(C.ref(raw_d), C.ref(C.nhwc(cost)), C.stream()))
        # dense_warp(h1 = g, cost) -> w
        C.check(_lib.hesic_dense_warp(C.ref(C.split(lvl, N, 2 * N)), C.ref(C.nhwc(cost)), C.ref(C.split(lvl, N, 0)), C.stream()))

    # ---- forward --------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x1, x2):
        m = self.m
        C.require_cuda(x1, x2)
        if x1.shape != x2.shape or x1.dim() != 4 or x1.shape[1] != 3:
            raise ValueError(f"expected two [B,3,H,W] images, got {tuple(x1.shape)} and {tuple(x2.shape)}")
        B, _, H, W = x1.shape
        if H % 64 or W % 64:
            raise ValueError("DSIC.forward needs H and W divisible by 64")
        with torch.cuda.device(x1.device):
            return self._forward(x1, x2, B, H, W)

    def capture(self, x1, x2):
        C.require_cuda(x1, x2)
        self._streams_for_capture(x1.device)
        return CapturedForward(self, (x1, x2), self.forward)

    def _forward(self, x1, x2, B, H, W):
        m = self.m
        self.dev = dev = x1.device
        self._work = B * H * W
        main = self._begin(dev)
        x1 = self._keep(x1.float().contiguous())
        x2 = self._keep(x2.float().contiguous())
        N, M, K = m.N, m.M, m.K
        acc = torch.zeros(4, device=dev, dtype=torch.float64)
        self.log2_sums = acc
        sse = torch.zeros(2, device=dev, dtype=torch.float64)   # squared errors of x1_hat / x2_hat, from the RGB heads' epilogues
        self.sse_sums = sse
        a = lambda i: acc[i:i + 1]
        lv = [None] + [self._split(B, H >> s, W >> s, 3 * N) for s in (1, 2, 3, 3, 2, 1)]   # level buffers 1..6
        g = lambda k: C.split(lv[k], N, 2 * N)
        wa = lambda k: C.split(lv[k], 2 * N, 0)

        # ---- view 1: encoder1 (features kept per level), hyper path, decoder1 -----------------------
        e1, d1 = m.encoder1, m.decoder1
        self._run(e1.g_a_conv1, self._rowpad("x1", C.nchw(x1), B, 3, H, W), B, H, W, "split", gdn=e1.g_a_gdn1, dst=(lv[1], 2 * N))
        self._run(e1.g_a_conv2, g(1), B, H >> 1, W >> 1, "split", gdn=e1.g_a_gdn2, dst=(lv[2], 2 * N))
        self._run(e1.g_a_conv3, g(2), B, H >> 2, W >> 2, "split", gdn=e1.g_a_gdn3, dst=(lv[3], 2 * N))
        y1, y1_d, Hy, Wy = self._run(e1.g_a_conv4, g(3), B, H >> 3, W >> 3, "nhwc")
        y1_abs = self._split(B, Hy, Wy, M)
        self._convert(y1_d, C.split(y1_abs), C.OP_ABS)
        R, L = C.ACT_RELU, C.ACT_LEAKY
        z1, z1_d, Hz, Wz = self._seq3(m._h_a1.encode_hyper, (0, 2, 4), (R, R, C.ACT_NONE), C.split(y1_abs), B, Hy, Wy, "nhwc")
        z1_hat, z1h_d, z1_lik = self._bottleneck(m.entropy_bottleneck1, z1, z1_d, B, Hz, Wz, a(2))
        hs = m._h_s1
        (_, s_d, _, _), (_, m_d, _, _), w1 = self._branches(
            lambda: self._seq3(hs.gmm_sigma, (0, 2, 4), (R, R, R), z1h_d, B, Hz, Wz, "nhwc"),
            lambda: self._seq3(hs.gmm_means, (0, 2, 4), (L, L, C.ACT_NONE), z1h_d, B, Hz, Wz, "nhwc"),
            lambda: self._mixture_head(hs.gmm_weights, z1h_d, B, Hz, Wz, K, M))
        y1_hat, y1_lik, y1h_d = self._gmm(m.gaussian1, y1_d, s_d, m_d, w1, B, Hy, Wy, M, K, a(0))
        self._run(d1.g_s_conv1, y1h_d, B, Hy, Wy, "split", gdn=d1.g_s_gdn1, dst=(lv[4], 2 * N))
        self._run(d1.g_s_conv2, g(4), B, H >> 3, W >> 3, "split", gdn=d1.g_s_gdn2, dst=(lv[5], 2 * N))
        self._run(d1.g_s_conv3, g(5), B, H >> 2, W >> 2, "split", gdn=d1.g_s_gdn3, dst=(lv[6], 2 * N))
        x1_hat, _, _, _ = self._run(d1.g_s_conv4, g(6), B, H >> 1, W >> 1, "nchw", sse=(C.nchw(x1), sse[0:1]))

        # ---- global context volumes from y1_hat (mynet6_plus.py:240-246) ------------------------------
        gc = m._global_context.global_net
        v = self._conv_gn(gc[0], gc[1], y1h_d, B, Hy, Wy)
        v = self._conv_gn(gc[3], gc[4], v, B, Hy, Wy)
        v = self._conv_gn(gc[6], gc[7], v, B, Hy, Wy)
        cv1 = m._cost_volume1
        gc_last = self._perm_plan(gc[9], "dm", out_perm=self._depth_major(cv1.F0, cv1.C, 3))
        ctx, _, _, _ = self._conv(gc_last, v, B, Hy, Wy, "nhwc")   # [B, Hy, Wy, 3 x (C depths x F0 features)], depth-major

        # ---- view 2 analysis -------------------------------------------------------------------------
        self._run(m.pic2_g_a_conv1, self._rowpad("x2", C.nchw(x2), B, 3, H, W), B, H, W, "split", gdn=m.pic2_g_a_gdn1, dst=(lv[1], N))
        self._cost_volume(m._cost_volume1, lv[1], ctx, 0, B, H >> 1, W >> 1)
        self._run(m.pic2_g_a_conv2, wa(1), B, H >> 1, W >> 1, "split", gdn=m.pic2_g_a_gdn2, dst=(lv[2], N))
        self._cost_volume(m._cost_volume2, lv[2], ctx, 1, B, H >> 2, W >> 2)
        self._run(m.pic2_g_a_conv3, wa(2), B, H >> 2, W >> 2, "split", gdn=m.pic2_g_a_gdn3, dst=(lv[3], N))
        self._cost_volume(m._cost_volume3, lv[3], ctx, 2, B, H >> 3, W >> 3)
        y2, y2_d, _, _ = self._run(m.pic2_g_a_conv4, wa(3), B, H >> 3, W >> 3, "nhwc")

        # ---- view 2 entropy model, conditioned on y1_hat (mynet6_plus.py:719-723) ---------------------
        y2_abs = self._split(B, Hy, Wy, M)
        self._convert(y2_d, C.split(y2_abs), C.OP_ABS)
        z2, z2_d, Hz, Wz = self._seq3(m._h_a2.encode_hyper, (0, 2, 4), (R, R, C.ACT_NONE), C.split(y2_abs), B, Hy, Wy, "nhwc")
        z2_hat, z2h_d, z2_lik = self._bottleneck(m.entropy_bottleneck2, z2, z2_d, B, Hz, Wz, a(3))
        cond = self._split(B, Hy, Wy, N + M)                                       # cat(up(z2_hat), y1_hat)
        C.check(_lib.hesic_upsample_bilinear(C.ref(z2h_d), C.ref(C.split(cond, N, 0)), 4, C.stream()))
        self._convert(y1h_d, C.split(cond, M, N))
        hs = m._h_s2
        cd = C.split(cond)
        _, s_d, _, _ = self._seq3(hs.gmm_sigma, (0, 2, 4), (R, R, R), cd, B, Hy, Wy, "nhwc")
        _, m_d, _, _ = self._seq3(hs.gmm_means, (0, 2, 4), (L, L, C.ACT_NONE), cd, B, Hy, Wy, "nhwc")
        w2 = self._mixture_head(hs.gmm_weights, cd, B, Hy, Wy, K, M)
        y2_hat, y2_lik, y2h_d = self._gmm(m.gaussian2, y2_d, s_d, m_d, w2, B, Hy, Wy, M, K, a(1))

        # ---- view 2 synthesis ------------------------------------------------------------------------
        self._run(m.pic2_g_s_conv1, y2h_d, B, Hy, Wy, "split", gdn=m.pic2_g_s_gdn1, dst=(lv[4], N))
        self._cost_volume(m._cost_volume4, lv[4], ctx, 2, B, H >> 3, W >> 3)
        self._run(m.pic2_g_s_conv2, wa(4), B, H >> 3, W >> 3, "split", gdn=m.pic2_g_s_gdn2, dst=(lv[5], N))
        self._cost_volume(m._cost_volume5, lv[5], ctx, 1, B, H >> 2, W >> 2)
        self._run(m.pic2_g_s_conv3, wa(5), B, H >> 2, W >> 2, "split", gdn=m.pic2_g_s_gdn3, dst=(lv[6], N))
        self._cost_volume(m._cost_volume6, lv[6], ctx, 0, B, H >> 1, W >> 1)
        x2_hat, _, _, _ = self._run(m.pic2_g_s_conv4, wa(6), B, H >> 1, W >> 1, "nchw", sse=(C.nchw(x2), sse[1:2]))

        self._end(main)
        return {"x1_hat": x1_hat, "x2_hat": x2_hat,
                "likelihoods": {"y1": y1_lik, "y2": y2_lik, "z1": z1_lik, "z2": z2_lik}}
