"""Stage-by-stage check of the operator-level modules at batch 2 against the oracle (debug)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import hesic_b200
from hesic_b200 import synth
hesic_b200.install()
from oracle import hesic_oracle as O
from tests.helpers import close_stats
import mynet6_plus
DEV = "cuda:0"
net = mynet6_plus.DSIC(128, 192, 21, 32, 5).eval()
sd = synth.synth_state_dict(net, seed=0); net.load_state_dict(sd); net = net.to(DEV)
for B in (1, 2):
    x1, x2, _ = synth.stereo_pairs(B, 128, 320, seed=21)
    with torch.no_grad():
        e = O._analysis(sd, "encoder1", x1) if hasattr(O, "_analysis") else None
        y1, g1, g2, g3 = net.encoder1(x1.to(DEV))
        # oracle pieces
        import torch.nn.functional as Fn
        def cw(p): return sd[p + ".weight"], sd[p + ".bias"]
        r1 = O.gdn(O.conv(x1, *cw("encoder1.g_a_conv1")), sd["encoder1.g_a_gdn1.beta"], sd["encoder1.g_a_gdn1.gamma"])
        r2 = O.gdn(O.conv(r1, *cw("encoder1.g_a_conv2")), sd["encoder1.g_a_gdn2.beta"], sd["encoder1.g_a_gdn2.gamma"])
        r3 = O.gdn(O.conv(r2, *cw("encoder1.g_a_conv3")), sd["encoder1.g_a_gdn3.beta"], sd["encoder1.g_a_gdn3.gamma"])
        ry = O.conv(r3, *cw("encoder1.g_a_conv4"))
        print(B, "g1", close_stats(g1, r1), "g2", close_stats(g2, r2), "g3", close_stats(g3, r3), "y1", close_stats(y1, ry))
        z = net._h_a1(y1)
        zh, zl = net.entropy_bottleneck1(z)
        s, mu, w = net._h_s1(zh)
        yh, yl = net.gaussian1(y1, s, mu, w)
        print(B, "y_hat per-sample abs mean", [float(yh[i].abs().mean()) for i in range(B)], "w", [float(w[i].abs().sum()) for i in range(B)])
        xh, d1, d2, d3 = net.decoder1(yh)
        # per-sample consistency: sample 0 of batch run vs alone
        if B == 2:
            y1a = net.encoder1(x1[:1].to(DEV))[0]
            print("enc batch-vs-single", float((y1a - y1[:1]).abs().max()))
            za = net._h_a1(y1a); zha, _ = net.entropy_bottleneck1(za)
            print("z", float((za - z[:1]).abs().max()), float((zha - zh[:1]).abs().max()))
            sa, mua, wa = net._h_s1(zha)
            print("sigma", float((sa - s[:1]).abs().max()), "mu", float((mua - mu[:1]).abs().max()), "w", float((wa - w[:1]).abs().max()))
            yha, yla = net.gaussian1(y1a, sa, mua, wa)
            print("yhat", float((yha - yh[:1]).abs().max()), "lik", float((yla - yl[:1]).abs().max()))
            xha = net.decoder1(yha)[0]
            print("xhat", float((xha - xh[:1]).abs().max()))
            y1b = net.encoder1(x1[1:].to(DEV))[0]
            zhb, _ = net.entropy_bottleneck1(net._h_a1(y1b))
            sb, mub, wb = net._h_s1(zhb)
            print("sample1: sigma", float((sb - s[1:]).abs().max()), "mu", float((mub - mu[1:]).abs().max()), "w", float((wb - w[1:]).abs().max()))
            yhb, ylb = net.gaussian1(y1b, sb, mub, wb)
            print("sample1: yhat", float((yhb - yh[1:]).abs().max()), "lik", float((ylb - yl[1:]).abs().max()))
            xhb = net.decoder1(yhb)[0]
            print("sample1: xhat", float((xhb - xh[1:]).abs().max()))

print("---- full forward comparisons")
for B in (1, 2):
    x1, x2, _ = synth.stereo_pairs(B, 128, 320, seed=21)
    taps = {}
    with torch.no_grad():
        ref = O.dsic_forward(sd, x1, x2, taps=taps)
    a = net(x1.to(DEV), x2.to(DEV))
    b = net.forward_operator_level(x1.to(DEV), x2.to(DEV))
    rel = lambda u, v: float((u.cpu().double() - v.double()).pow(2).sum().sqrt() / v.double().pow(2).sum().sqrt())
    print(B, "engine x1", rel(a["x1_hat"], ref["x1_hat"]), "x2", rel(a["x2_hat"], ref["x2_hat"]),
          "| oplevel x1", rel(b["x1_hat"], ref["x1_hat"]), "x2", rel(b["x2_hat"], ref["x2_hat"]),
          "| engine vs oplevel x1", rel(a["x1_hat"], b["x1_hat"].cpu()))
    y1, g1, g2, g3 = net.encoder1(x1.to(DEV))
    zh, zl = net.entropy_bottleneck1(net._h_a1(y1))
    s, mu, w = net._h_s1(zh)
    yh, yl = net.gaussian1(y1, s, mu, w)
    print(B, "taps:", sorted(taps)[:40])
    if "y1_hat" in taps:
        print(B, "oplevel y1_hat mismatch frac", float((yh.cpu() != taps["y1_hat"]).double().mean()))
    xh = net.decoder1(yh)[0]
    print(B, "oplevel decoder1(yh) vs oplevel forward x1_hat", rel(xh, b["x1_hat"].cpu()), " vs ref", rel(xh, ref["x1_hat"]))
    with torch.no_grad():
        xo = O._synthesis(sd, "decoder1", yh.cpu()) if hasattr(O, "_synthesis") else None
    if xo is not None:
        xo = xo[0] if isinstance(xo, (tuple, list)) else xo
        print(B, "oracle decoder on oplevel y_hat vs oplevel x1_hat", rel(xh, xo))
