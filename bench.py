#!/usr/bin/env python
"""Headline benchmark: stereo pairs/sec of HSIC.forward at 512x512 (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repository (B200 kernels)
    python bench.py --impl reference --steps K --warmup W    # CPU implementation of the same path

One "step" = one pass of the hot path over one batch of synthetic stereo pairs
(16 pairs per GPU: BASELINE config[1] at N=1, config[3] at N=8 -- weak scaling),
including the fused rate/distortion partial sums and, for N>1, the single NCCL
all-reduce of those scalars.  Prints ONE JSON line (rank 0).

Besides the contract's keys the line carries
  parity    rank 0's first pair of the timed batch against the CPU oracle (bpp, PSNR, symbol flips), at every N;
  sustained the headline loop run for >= 3 s with its own clock sample (steady-state power/thermal clock);
  extras    (N = 1 only) the other BASELINE configurations and the callers either side of the path, each timed the
            same way: HESIC+ B=16 (config 3), DSIC B=8 (config 5), Independent_EN B=16, the per-batch body of
            test3real.py's test_epoch, and CUDA-graph replay latency of one forward at B=1 and B=16.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.md section 2 / SURVEY.md 8(d): algorithmic FLOPs per stereo pair (2*MAC, conv + GDN contraction)
GFLOP_PER_PAIR = {"hesic": 155.66, "hesic_plus": 130.09, "dsic": 1366.5, "hesic_nt": 133.55, "en": 176.67}
MODEL_INFO = {"hesic": ("HESIC newnet1.HSIC", "hsic_newnet1", "newnet1", 2, 16),
              "hesic_plus": ("HESIC+ newnet1_joint.HSIC", "hsic_joint", "newnet1_joint", 3, 16),
              "dsic": ("DSIC mynet6_plus.DSIC", "dsic", "mynet6_plus", 5, 8)}
# g_a_conv2 + g_a_gdn2 of Encoder1 (newnet1.py:585,593-596): 13.42 + 0.54 GF per pair, runs 3x per forward
DOMINANT = dict(Cin=128, Cout=128, k=5, stride=2, H=256, W=256)
METRIC = "stereo pairs/sec @512x512 (HSIC.forward)"


def make_config(model, B, world):
    """The workload description shared verbatim by both arms (the CPU arm runs the same step on the host cores)."""
    title, _, _, cfg, _ = MODEL_INFO[model]
    return {"workload": f"{title} forward, batch {B} x 512x512 synthetic stereo pairs per GPU (BASELINE config {cfg}"
                        f"{'; config 4 at 8 GPUs' if model == 'hesic' else ''})",
            "pairs_per_gpu": B, "global_pairs": B * world,
            "l2": "two alternating resident input batches; per-step activations exceed the 126 MB L2",
            "gflop_per_pair": GFLOP_PER_PAIR[model]}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe): one long-lived
    `nvidia-smi -lms 50` process, its lines collected by this thread."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                c = [x.strip() for x in line.strip().split(",")]
                if len(c) >= 7:
                    self.rows.append(c)
        except Exception:
            pass

    def mark(self):
        """Number of samples so far (to select the ones taken inside a timed region)."""
        return len(self.rows)

    def summary(self, first=0, last=None):
        rows = self.rows[first:last] or self.rows
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        pw = sorted(float(r[2]) for r in rows if r[2].replace(".", "").isdigit())
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": float(rows[0][1]) if rows else None, "power_w": pw[len(pw) // 2] if pw else None,
                "reasons": sorted(reasons), "samples": len(rows)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=3)


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a torch-CPU fp32 port of the reference's forward, bit-identical to it: VERDICT r01) on the
# host cores.  Nothing here imports a model class of this repository or loads libhesic_b200.so: the weights come
# from the reference's key table (tests/golden/*.json) through oracle/default_state.py + hesic_b200/synth.py (numpy).
def cpu_weights(model):
    from hesic_b200 import synth            # numpy/torch only; `import hesic_b200` does not load the library
    from oracle import default_state
    return synth.synth_state_dict(default_state.initial_state_dict(default_state.spec(MODEL_INFO[model][1])), seed=0)


def cpu_forward_fn(model):
    from oracle import hesic_oracle as O
    if model == "hesic_plus":
        return O.hsic_joint_forward
    if model == "dsic":
        return lambda sd, x1, x2, h: O.dsic_forward(sd, x1, x2)
    return O.hsic_forward


def cpu_port_time(model, pairs_per_step, steps, warmup, seed=1234, max_seconds=None):
    """Each step = ``pairs_per_step`` single-pair forwards (batch 1 is the CPU's fastest batch size per pair).
    ``max_seconds`` bounds the sample: timing stops once that much CPU time has been measured (>= 2 steps)."""
    import torch
    from hesic_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    sd = cpu_weights(model)
    fwd = cpu_forward_fn(model)
    x1, x2, h = synth.stereo_pairs(pairs_per_step, 512, 512, seed=seed)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            for j in range(pairs_per_step):
                fwd(sd, x1[j:j + 1], x2[j:j + 1], h[j:j + 1])
            if i >= warmup:
                times.append(time.perf_counter() - t0)
                if max_seconds and len(times) >= 2 and sum(times) >= max_seconds:
                    break
    sec = sum(times) / len(times)
    return pairs_per_step / sec, torch.get_num_threads(), sec, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.batch or MODEL_INFO[args.model][4]
    pps, cores, sec, _ = cpu_port_time(args.model, B, max(args.steps, 1), max(args.warmup, 0))
    line = {"impl": "reference", "metric": METRIC, "value": pps, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(args.model, B, 1),
            "cpu_baseline": {"value": pps, "unit": "pairs/s", "cores": cores, "kind": "port",
                             "sample": f"every step = the whole {B}-pair batch, run as {B} single-pair forwards (the CPU's fastest "
                                       "batch size per pair) of oracle/hesic_oracle.py -- the torch CPU fp32 port of the reference's "
                                       "forward, bit-identical to it -- on all host threads; rank 0 only"},
            "e2e": {"value": pps, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import hesic_b200
    from hesic_b200 import _capi as C
    from hesic_b200 import functional as F
    from hesic_b200 import sharding, synth
    hesic_b200.install()

    # libraries (NCCL prints its version banner) write to fd 1: park stdout on stderr until the result line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    C.check(C.lib.hesic_device_check(None, 0))

    model = args.model
    B = args.batch or MODEL_INFO[model][4]

    def build(kind):
        modname = MODEL_INFO[kind][2]
        mod = __import__(modname)
        net = (mod.DSIC(128, 192, 21, 32, 5) if kind == "dsic" else mod.HSIC(128, 192, 5)).eval()
        sd = synth.synth_state_dict(net, seed=0)
        net.load_state_dict(sd)
        return net.to(dev), sd

    def call(net, kind, x1, x2, h):
        return net(x1, x2) if kind == "dsic" else net(x1, x2, h)

    net, sd_host = build(model)
    # two different resident batches, alternated, so no step re-reads the previous step's inputs from L2;
    # per-step activations (GBs) exceed the 126 MB L2 by far in any case
    sets, host = [], []
    for s in range(2):
        x1, x2, h = synth.stereo_pairs(B, 512, 512, seed=1234 + 17 * rank + s)
        host.append((x1.pin_memory(), x2.pin_memory(), h.pin_memory()))
        sets.append((x1.to(dev), x2.to(dev), h.to(dev)))
    partial = torch.zeros(6, device=dev, dtype=torch.float64)   # sum log2 p (y1,y2,z1,z2), SSE view1, SSE view2
    result_host = torch.zeros(6, dtype=torch.float64).pin_memory()

    def step(x1, x2, h):
        out = call(net, model, x1, x2, h)
        partial.zero_()
        eng = net.hesic_engine
        partial[:4].copy_(eng.log2_sums)
        if getattr(eng, "sse_sums", None) is not None:
            # HESIC / HESIC+: the squared errors come out of the epilogues that store x1_hat / x2_hat (no pass over the images)
            partial[4:].copy_(eng.sse_sums)
        else:
            F.sum_squared_error(out["x1_hat"], x1, partial[4:5])
            F.sum_squared_error(out["x2_hat"], x2, partial[5:6])
        sharding.reduce_partials(partial)     # the path's only collective: 48 bytes over NVLink (no-op at N=1)
        return out

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, whole=False, reduce=True):
        sync() if reduce else torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if whole:
            fn(steps)
        else:
            for i in range(steps):
                fn(i)
        e1.record()
        sync() if reduce else torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1 and reduce:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for i in range(args.warmup):
        step(*sets[i % 2])
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    C.lib.hesic_launch_count(1)
    mark0 = sampler.mark() if sampler else 0
    ms = timed(lambda i: step(*sets[i % 2]), args.steps)
    launches = C.lib.hesic_launch_count(0)
    last_set = (args.steps - 1) % 2
    m_dev = sharding.metrics_from_partials(partial.cpu(), B * world, 512, 512)
    m_dev = {k: m_dev[k] for k in ("bpp", "psnr1", "psnr2")}

    # end to end through the public API: pinned host inputs -> H2D -> forward -> metric partials -> D2H.
    # Every step's inputs are copied inside the timed region; hostfeed.HostFeed copies batch i+1 on a side
    # stream while batch i computes.
    from hesic_b200.hostfeed import HostFeed
    # the images travel as 8-bit samples ([B,H,W,3] uint8, what the reference's loader holds before ToTensor) and are
    # converted to fp32 by the first kernel on the device; h_matrix as fp32
    host8 = [tuple((t.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous().pin_memory() for t in hb[:2]) + (hb[2],)
             for hb in host]
    feed = HostFeed(dev, host8[0])

    def e2e_one(dx1, dx2, dh):
        step(dx1, dx2, dh)
        result_host.copy_(partial, non_blocking=True)

    def e2e_all(steps):
        feed.run([host8[i % 2] for i in range(steps)], e2e_one)

    e2e_all(2)
    ms_e2e = timed(e2e_all, args.steps, whole=True)
    mark1 = sampler.mark() if sampler else 0
    h2d = sum(t.numel() * t.element_size() for t in host8[0])
    # the same loop with fp32 images on the wire (4 bytes per sample), for comparison
    feed32 = HostFeed(dev, host[0])
    feed32.run([host[i % 2] for i in range(2)], e2e_one)
    ms_e2e32 = timed(lambda n: feed32.run([host[i % 2] for i in range(n)], e2e_one), args.steps, whole=True)
    h2d32 = sum(t.numel() * t.element_size() for t in host[0])
    del feed32

    # steady state: the same loop for >= 3 s (the 10-20-step burst above is shorter than the power/thermal time constant)
    n_sus = max(args.steps, int(math.ceil(args.sustain_s * 1e3 / (ms / args.steps))))
    ms_sus = timed(lambda i: step(*sets[i % 2]), n_sus)
    mark2 = sampler.mark() if sampler else 0

    # parity of the timed computation itself: pair 0 of this rank's last timed batch against the CPU oracle
    parity = None
    if rank == 0:
        # rank 0 alone: the forward only -- NO collective here (the other ranks have left the timed loops)
        out = call(net, model, *sets[last_set])
        torch.cuda.synchronize()
        one = {k: v[:1].cpu() for k, v in out.items() if k != "likelihoods"}
        one["likelihoods"] = {k: v[:1].cpu() for k, v in out["likelihoods"].items()}
        hx1, hx2, hh = (t[:1] for t in host[last_set])
        with torch.no_grad():
            torch.set_num_threads(os.cpu_count() or 1)
            ref = cpu_forward_fn(model)({k: v for k, v in sd_host.items()}, hx1, hx2, hh)
        mg, mr = synth.rd_metrics(one, hx1, hx2), synth.rd_metrics(ref, hx1, hx2)
        parity = {"against": "oracle (CPU fp32 port of the reference) on pair 0 of rank 0's last timed batch",
                  "bpp": mg["bpp"], "bpp_oracle": mr["bpp"], "bpp_rel": abs(mg["bpp"] - mr["bpp"]) / mr["bpp"],
                  "psnr_abs": max(abs(mg["psnr1"] - mr["psnr1"]), abs(mg["psnr2"] - mr["psnr2"])),
                  "psnr1": mg["psnr1"], "psnr2": mg["psnr2"],
                  "x_hat_rel_l2": max(float((one[k].double() - ref[k].double()).norm() / ref[k].double().norm())
                                      for k in ("x1_hat", "x2_hat"))}
        if "y1_hat" in one and "y1_hat" in ref:
            parity["flips"] = max(float((one[k] != ref[k]).double().mean()) for k in ("y1_hat", "y2_hat"))
        # a hyper-latent (z) symbol that sits within ~1e-5 of a rounding boundary may round the other way under a different
        # fp32 summation order; it moves the mixture parameters of a 4x4 latent block and with them a few hundred bits of
        # this ONE pair (the 16-pair batch agrees to 1e-7, tests/test_gpu_fullsize.py) -- visible as a likelihood that jumps
        zrel = max(float(((one["likelihoods"][k].double() - ref["likelihoods"][k].double()).abs()
                          / ref["likelihoods"][k].double().clamp(min=1e-9)).max()) for k in ("z1", "z2"))
        parity["z_likelihood_max_rel"] = zrel
        parity["z_symbol_flipped"] = bool(zrel > 1e-2)
        parity["batch_metrics_all_ranks"] = m_dev

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    clocks = sampler.summary(mark0, mark1)
    clocks_sus = sampler.summary(mark1, mark2)

    # dominant kernel (g_a_conv2 + fused g_a_gdn2: 128->128, k5, s2, 256^2 -> 128^2), timed alone with CUDA events on
    # the launching stream; profiles/dominant_kernel_traffic.json holds the ncu DRAM bytes of the SAME launch
    # (tools/run_dominant.py)
    peaks = load_peaks()
    from compressai.models.utils import conv
    d = DOMINANT
    Bk = 16
    layer = conv(d["Cin"], d["Cout"], kernel_size=d["k"], stride=d["stride"]).to(dev)
    plan = layer.hesic_plan()
    plan.set_gdn(torch.ones(d["Cout"], device=dev), 0.1 * torch.eye(d["Cout"], device=dev) + 0.01, False)
    xin = torch.randn(2, Bk, d["H"], d["W"], d["Cin"], device=dev).to(torch.bfloat16)
    Ho, Wo = plan.out_hw(d["H"], d["W"])
    yout = torch.empty(2, Bk, Ho, Wo, d["Cout"], device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    kt = []
    for i in range(8):
        flush.zero_()   # evict L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run(C.split(xin), C.split(yout), C.ACT_NONE, C.PATH_TC)
        e1.record()
        torch.cuda.synchronize()
        if i > 1:
            kt.append(e0.elapsed_time(e1))
    C.check(C.lib.hesic_tc_status())
    k_ms = sum(kt) / len(kt)
    k_flop = 2.0 * Bk * Ho * Wo * d["Cout"] * (d["Cin"] * d["k"] * d["k"] + d["Cout"])
    k_tflops = k_flop / (k_ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    del xin, yout, flush

    total_pairs = B * world * args.steps
    value = total_pairs / (ms * 1e-3)
    e2e_value = total_pairs / (ms_e2e * 1e-3)
    sus_value = B * world * n_sus / (ms_sus * 1e-3)
    gf = GFLOP_PER_PAIR[model]

    extras = None
    if world == 1 and not args.no_extras:
        extras = run_extras(args, dev, build, call, timed, C, F, synth, sets, net, model)

    cpu_pps, cores, cpu_sec, cpu_n = cpu_port_time(model, 1, args.cpu_iters, 1, max_seconds=20.0)
    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16x3 (fp32 split into bf16 hi+lo, fp32 accumulate)",
        "data": "synthetic",
        "config": make_config(model, B, world),
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 48,
                "ms_per_step": ms_e2e / args.steps,
                "how": "hesic_b200.hostfeed.HostFeed: pinned host batch, images as [B,H,W,3] uint8 (the form the reference's loader holds "
                       "them in before ToTensor) -> H2D on a copy stream (double-buffered, batch i+1 copies while batch i computes; K "
                       "copies inside the timed region) -> u8/255 on the device -> HSIC.forward -> partial sums -> D2H",
                "fp32_on_the_wire": {"value": total_pairs / (ms_e2e32 * 1e-3), "h2d_bytes_per_step": h2d32,
                                     "ms_per_step": ms_e2e32 / args.steps}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "parity": parity,
        "sustained": {"value": sus_value, "unit": "pairs/s", "steps": n_sus, "seconds": ms_sus * 1e-3,
                      "ms_per_step": ms_sus / n_sus, "clocks": clocks_sus},
        "roofline": {"bound": "tensor", "achieved": k_tflops, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                     "frac": k_tflops / peaks["bf16_burst"], "traffic": traffic,
                     "kernel": f"conv_tc_pair_kernel: g_a_conv2 + fused g_a_gdn2, 128->128 k5 s2 256^2->128^2 x{Bk} (one launch)",
                     "kernel_ms": k_ms, "algorithmic_flop_per_launch": k_flop,
                     "algorithmic_bytes_per_launch": 4.0 * Bk * (d["H"] * d["W"] * d["Cin"] + Ho * Wo * d["Cout"]),
                     "peak_source": peaks["source"] + ", burst (kernel timed alone)",
                     "note": "algorithmic FLOPs (2*MAC of the conv + the GDN contraction), not inflated by the 3 bf16 products "
                             "per MAC (fp32 parity: Ah.Wh + Ah.Wl + Al.Wh); the tensor pipe executes 3x this",
                     "mma_issued": {"achieved": 3.0 * k_tflops, "frac": 3.0 * k_tflops / peaks["bf16_burst"],
                                    "unit": "TFLOP/s of bf16 MMAs actually executed, of the same measured peak"},
                     "whole_step": {"achieved": value / world * gf / 1e3, "peak": peaks["bf16_sustained"],
                                    "frac": value / world * gf / 1e3 / peaks["bf16_sustained"],
                                    "unit": "TFLOP/s per GPU, of measured sustained bf16"}},
        "cpu_baseline": {"value": cpu_pps, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"{cpu_n} forwards of 1 synthetic 512x512 pair with the oracle (torch CPU fp32 port of "
                                   f"the reference's forward), {cpu_sec:.2f} s each"},
    }
    if extras is not None:
        line["extras"] = extras
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)     # the JSON line is the only thing this process writes to stdout
    print(json.dumps(line), flush=True)


def run_extras(args, dev, build, call, timed, C, F, synth, sets, net, model):
    """The other BASELINE configurations and the path's callers, timed like the headline (device-resident inputs,
    CUDA events, >= 3 warm-up steps).  N = 1 only."""
    import torch
    ex = {}
    K, W = max(args.steps, 5), max(args.warmup, 3)

    def measure(fn, n_pairs, gflop_pair, steps=K):
        for _ in range(W):
            fn()
        C.lib.hesic_launch_count(1)
        ms = timed(lambda i: fn(), steps, reduce=False) / steps
        n = C.lib.hesic_launch_count(0) // steps
        r = {"pairs_per_s": n_pairs / ms * 1e3, "ms_per_step": ms, "pairs_per_step": n_pairs, "launches_per_step": int(n)}
        if gflop_pair:
            r["alg_tflops"] = n_pairs / ms * gflop_pair
        return r

    def guard(name, fn):
        try:
            ex[name] = fn()
        except Exception as e:           # an extra must never cost the headline line
            ex[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

    x1, x2, h = sets[0]

    def other_model(kind):
        def go():
            n2, _ = build(kind)
            Bk = MODEL_INFO[kind][4]
            a, b, hh = (t.to(dev) for t in synth.stereo_pairs(Bk, 512, 512, seed=1234))
            r = measure(lambda: call(n2, kind, a, b, hh), Bk, GFLOP_PER_PAIR[kind], steps=K if kind != "dsic" else max(5, K // 2))
            r["workload"] = make_config(kind, Bk, 1)["workload"]
            return r
        return go

    for kind in ("hesic", "hesic_plus", "dsic"):
        if kind != model:
            guard(kind, other_model(kind))

    def en_flow():
        import newnet1
        en = newnet1.Independent_EN().eval()
        en.load_state_dict(synth.synth_state_dict(en, seed=0))
        en = en.to(dev)
        a = torch.rand(16, 3, 512, 512, device=dev)
        b = torch.rand(16, 3, 512, 512, device=dev)
        h16 = h.repeat(16 // h.shape[0] + 1, 1, 1)[:16].contiguous()
        r = measure(lambda: en(a, b, h16), 16, GFLOP_PER_PAIR["en"])
        r["workload"] = "Independent_EN forward (newnet1.py:1278-1300), batch 16 x 512x512 (SURVEY 8f rank 1)"
        return r
    guard("independent_en", en_flow)

    def epoch_flow():
        import kornia
        import newnet9
        from model import Net
        homo = Net(patch_size=128).eval()
        sdh = synth.synth_state_dict(homo, seed=0)
        sdh["fc.5.weight"] *= 0.01
        sdh["fc.5.bias"] *= 0.01
        homo.load_state_dict(sdh)
        n9 = newnet9.HSIC(128, 192, 5).eval()
        n9.load_state_dict(synth.synth_state_dict(n9, seed=0))
        en = newnet9.Independent_EN().eval()
        en.load_state_dict(synth.synth_state_dict(en, seed=0))
        homo, n9, en = homo.to(dev), n9.to(dev), en.to(dev)
        d1, d2 = (t.to(dev) for t in synth.stereo_pairs(16, 512, 512, seed=1234)[:2])
        g1 = torch.nn.functional.interpolate(d1.mean(1, keepdim=True), size=(128, 128), mode="bilinear", align_corners=False)
        g2 = torch.nn.functional.interpolate(d2.mean(1, keepdim=True), size=(128, 128), mode="bilinear", align_corners=False)
        corners = torch.tensor([[[64., 64.], [191., 64.], [191., 191.], [64., 191.]]], device=dev).repeat(16, 1, 1)

        def body():
            with torch.no_grad():
                c0 = corners - corners[:, 0].view(-1, 1, 2)
                delta = homo(g1, g2)
                hm = torch.inverse(kornia.get_perspective_transform(c0, c0 + delta))
                a = d1.shape[-2] / 256
                hm[:, 0, :] = a * hm[:, 0, :]
                hm[:, :, 0] = (1. / a) * hm[:, :, 0]
                hm[:, 1, :] = a * hm[:, 1, :]
                hm[:, :, 1] = (1. / a) * hm[:, :, 1]
                out = n9(d1, d2, hm)
                return en(out["x1_hat"], out["x2_hat"], hm)
        r = measure(body, 16, GFLOP_PER_PAIR["hesic_nt"] + GFLOP_PER_PAIR["en"] + 2.6)
        r["workload"] = ("per-batch body of test_epoch in ywz/mywork/test3real.py:171-186 (homography Net -> get_perspective_transform "
                         "-> inverse -> h_adjust -> newnet9.HSIC -> Independent_EN), batch 16 x 512x512")
        return r
    guard("test_epoch_flow", epoch_flow)

    def latency():
        """One forward replayed from a CUDA graph (engine.capture): device time of a single call, B = 1 and B = 16."""
        if not hasattr(net.hesic_engine, "capture"):
            raise NotImplementedError("engine has no graph capture")
        r = {}
        for Bl in (1, 16):
            a, b, hh = (t[:Bl].contiguous() for t in sets[0])
            g = net.hesic_engine.capture(a, b, hh)
            for _ in range(3):
                g.replay()
            ms = timed(lambda i: g.replay(), 20, reduce=False) / 20
            r[f"B{Bl}"] = {"ms": ms, "pairs_per_s": Bl / ms * 1e3, "kernels_per_forward": g.n_launches}
            # the same call issued eagerly (Python-driven launches, per-forward allocations)
            for _ in range(3):
                net(a, b, hh)
            ms_e = timed(lambda i: net(a, b, hh), 20, reduce=False) / 20
            r[f"B{Bl}"]["eager_ms"] = ms_e
        r["workload"] = f"{MODEL_INFO[model][0]} forward latency at 512x512, CUDA-graph replay vs eager launches"
        return r
    if model != "dsic":
        guard("latency", latency)

    def hbm_kernels():
        """The HBM-side kernels of one forward (everything that is not a K-heavy contraction), each timed ALONE with CUDA events
        on the launching stream, L2 flushed between launches, against the measured copy bandwidth: algorithmic bytes = the
        fp32-equivalent tensors the operator reads and writes once, at the benchmark's batch (16 x 512 x 512)."""
        from compressai.layers import GDN
        from compressai.models.utils import conv, deconv
        peak = load_peaks()["hbm"]
        Bk, H, Wd = 16, 512, 512
        g = torch.Generator(device="cpu").manual_seed(7)
        flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
        rows = {}

        def run(name, nbytes, fn, ref=None):
            ts = []
            for i in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                if i > 1:
                    ts.append(e0.elapsed_time(e1))
            us = 1e3 * sum(ts) / len(ts)
            rows[name] = {"us": us, "algorithmic_mb": nbytes / 1e6, "gb_s": nbytes / us / 1e3, "frac_of_hbm_peak": nbytes / us / 1e3 / peak}
            if ref:
                rows[name]["reference"] = ref

        img = lambda: torch.rand((Bk, 3, H, Wd), generator=g).to(dev)
        x1, x2 = img(), img()
        # first analysis layer: conv(3, 128, k5, s2) + GDN, ROWPAD4 planes in, SPLIT planes out
        l1 = conv(3, 128, kernel_size=5, stride=2).to(dev)
        p1 = l1.hesic_plan()
        p1.set_gdn(torch.ones(128, device=dev), 0.1 * torch.eye(128, device=dev) + 0.01, False)
        rp = torch.zeros((2, Bk, H + C.ROWPAD_Y, Wd + C.ROWPAD_X, 4), device=dev, dtype=torch.bfloat16)
        y1 = torch.empty((2, Bk, H // 2, Wd // 2, 128), device=dev, dtype=torch.bfloat16)
        run("x1 -> ROWPAD4 planes (rowpad4_pair_kernel)", Bk * H * Wd * (12 + 16),
            lambda: C.check(C.lib.hesic_convert(C.ref(C.nchw(x1)), C.ref(C.rowpad(rp, 3)), C.OP_COPY, C.stream())))
        run("conv 3->128 k5 s2 + GDN (conv_tc_first_kernel)", 4 * Bk * (H * Wd * 3 + (H // 2) * (Wd // 2) * 128),
            lambda: p1.run(C.rowpad(rp, 3), C.split(y1), C.ACT_NONE, C.PATH_TC), "newnet1.py:583-601")
        p1.set_gdn(None, None, False)
        # RGB head: deconv(128, 3, k5, s2) + IGDN, SPLIT planes in, NCHW out
        l2 = deconv(128, 3, kernel_size=5, stride=2).to(dev)
        p2 = l2.hesic_plan()
        p2.set_gdn(torch.ones(3, device=dev), 0.1 * torch.eye(3, device=dev) + 0.01, True)
        out3 = torch.empty((Bk, 3, H, Wd), device=dev)
        run("deconv 128->3 k5 s2 + IGDN (conv_head_kernel)", 4 * Bk * ((H // 2) * (Wd // 2) * 128 + H * Wd * 3),
            lambda: p2.run(C.split(y1), C.nchw(out3), C.ACT_NONE, C.PATH_TC), "newnet1.py:606-624,669-670")
        # hesic_conv_forward_sse on this layer (x1_hat vs x1 in the forward): the head + the squared-error kernel, two launches
        acc_f = torch.zeros(1, device=dev, dtype=torch.float64)
        run("deconv 128->3 k5 s2 + IGDN + squared error vs target (conv_head_kernel, sse_dense_kernel)",
            4 * Bk * ((H // 2) * (Wd // 2) * 128 + 2 * H * Wd * 3),
            lambda: p2.run(C.split(y1), C.nchw(out3), C.ACT_NONE, C.PATH_TC, None, (C.nchw(x2), acc_f)), "newnet1.py:612 + test3real.py:99-111")
        p2.set_gdn(None, None, False)
        # full-resolution stencil: conv(6, 3, k5, s1) on cat(a, b) + GDN
        l3 = conv(6, 3, kernel_size=5, stride=1).to(dev)
        p3 = l3.hesic_plan()
        p3.set_gdn(torch.ones(3, device=dev), 0.1 * torch.eye(3, device=dev) + 0.01, False)
        run("conv 6->3 k5 s1 on cat + GDN (conv_small_kernel)", 4 * Bk * H * Wd * 9,
            lambda: p3.run(C.nchw(x1), C.nchw(out3), C.ACT_NONE, C.PATH_AUTO, C.nchw(x2)), "newnet1.py:643-644")
        p3.set_gdn(None, None, False)
        l4 = deconv(6, 3, kernel_size=5, stride=1).to(dev)
        p4 = l4.hesic_plan()
        p4.set_gdn(None, None, False)
        x3 = img()
        run("deconv 6->3 k5 s1 on cat + squared error vs target in the epilogue (conv_small_kernel)", 4 * Bk * H * Wd * 12,
            lambda: p4.run(C.nchw(x1), C.nchw(out3), C.ACT_NONE, C.PATH_AUTO, C.nchw(x2), (C.nchw(x3), acc_f)),
            "newnet1.py:686 + test3real.py:99-111")
        # warp
        hm = sets[0][2][:Bk].contiguous()
        run("warp_perspective, 3 channels (warp_rgb_kernel)", 4 * Bk * H * Wd * 6,
            lambda: F.warp_perspective(x1, hm, (H, Wd), out=out3), "newnet1.py:746")
        # GMM likelihood, K = 5, M = 192 at 32 x 32
        M_, Kc = 192, 5
        yl = (torch.randn((Bk, 32, 32, M_), generator=g) * 3).to(dev)                 # channels-last, as the engine feeds it
        sc = (torch.rand((Bk, 32, 32, M_ * Kc), generator=g) * 2 + 0.1).to(dev)
        mu = torch.randn((Bk, 32, 32, M_ * Kc), generator=g).to(dev)
        wt = torch.softmax(torch.randn((Bk, Kc, M_), generator=g), 1).reshape(Bk, Kc * M_).contiguous().to(dev)
        yh, lk = torch.empty((Bk, M_, 32, 32), device=dev), torch.empty((Bk, M_, 32, 32), device=dev)
        ys = torch.empty((2, Bk, 32, 32, M_), device=dev, dtype=torch.bfloat16)
        lacc = torch.zeros(1, device=dev, dtype=torch.float64)
        run("GMM likelihood K=5 + quantise + y_hat planes (gaussian_tile_kernel)", 4 * Bk * 32 * 32 * M_ * (4 + 2 * Kc),
            lambda: C.check(C.lib.hesic_gaussian_conditional(C.ref(C.nhwc(yl)), C.ref(C.nhwc(sc)), C.ref(C.nhwc(mu)), C.ptr(wt), Kc, 1,
                                                             0.11, 1e-9, C.ref(C.nchw(yh)), C.ref(C.nchw(lk)), C.ref(C.split(ys)),
                                                             C.ptr(lacc), C.stream())), "entropy_models.py:693-702")
        # SSE partial
        acc = torch.zeros(1, device=dev, dtype=torch.float64)
        run("sum of squared errors (sse_dense_kernel)", 4 * Bk * H * Wd * 6, lambda: F.sum_squared_error(x1, x2, acc), "test3real.py:99-111")
        return {"peak_gb_s": peak, "kernels": rows,
                "note": "single launches, CUDA events, L2 flushed; an isolated launch carries ~10 us of launch ramp that a launch "
                        "inside the forward does not (tools/time_gap.py)"}
    if model == "hesic":
        guard("hbm_kernels", hbm_kernels)
    return ex


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="hesic", choices=["hesic", "hesic_plus", "dsic"])
    ap.add_argument("--batch", type=int, default=0, help="stereo pairs per GPU per step (default: the BASELINE config's, 16; DSIC 8)")
    ap.add_argument("--cpu-iters", type=int, default=60, help="CPU-baseline forwards (about 0.2 s each on 16 cores)")
    ap.add_argument("--sustain-s", type=float, default=3.0, help="length of the steady-state loop in seconds")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra configurations (N = 1 prints them by default)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
