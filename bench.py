#!/usr/bin/env python
"""bench.py — MoCoGAN train steps/s on B200 (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            our arm: libmcg.so kernels, CUDA-graph step
  python bench.py --impl reference --gpus N --steps K ...  reference arm: the Chainer/NumPy CPU algorithm (oracle
                                                           port; Chainer 3.1.0 itself cannot run here) on host cores

Workload (config.workload): BASELINE config 2 — normal model as train.py builds it on MUG
(ImageGenerator(50,10,6,3,64,16), ImageDiscriminator(3,1,64,True,0.2), VideoDiscriminator(3,1,64,True,0.2)),
synthetic clips (35,3,16,64,64) per GPU, one step = one Updater.update_core (G, Di, Dv each updated once).
`value` = batch-35 steps per second summed over ranks (weak scaling), inputs already in HBM, device-timed.
`e2e`   = the same through the public API (Updater.update(): iterator -> pinned host batch -> H2D -> step -> D2H of
          the three losses every step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

USEFUL_GF_PER_STEP = 1839.2     # algorithmic FLOPs of one useful step at batch 35 (normal model)
AS_EXECUTED_GF_PER_STEP = 2545.9
BATCH = 35
CLIP = (3, 16, 64, 64)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ our arm
class PinnedClipIterator(object):
    """Synthetic MUG-shaped data: a ring of pre-stacked batches in PINNED host memory (float32 (N,3,16,64,64) in
    [-1,1), labels in [0,6)) — the output contract of datasets.py:105-107 after concat_examples."""

    def __init__(self, batch, seed, ring=4):
        import torch
        rng = np.random.default_rng(seed)
        self.x, self.t = [], []
        for _ in range(ring):
            x = torch.from_numpy(rng.uniform(-1, 1, size=(batch,) + CLIP).astype(np.float32)).pin_memory()
            t = torch.from_numpy(rng.integers(0, 6, size=batch).astype(np.int32)).pin_memory()
            self.x.append(x)
            self.t.append(t)
        self.i, self.epoch, self.is_new_epoch, self.epoch_detail = 0, 0, False, 0.0
        self.batch_size = batch

    def next(self):
        k = self.i % len(self.x)
        self.i += 1
        return PreStacked(self.x[k], self.t[k])

    __next__ = next


class PreStacked(list):
    """A batch that is already concatenated; concat_examples passes it through (see build_updater)."""

    def __init__(self, x, t):
        super(PreStacked, self).__init__()
        self.x, self.t = x, t

    def __len__(self):
        return self.x.shape[0]


def make_clip_cache(batch, seed, n_videos=70, frames=40):
    """The e2e input: pre-decoded uint8 videos (MUG-shaped: 64x64x3, 40 frames each) in one pinned buffer; every step the
    cache draws a 16-frame sub-sequence per clip (datasets.py:72-88), gathers the batch into pinned staging memory and the
    updater copies it host->device as uint8; (v-128)/128 happens on the device (mocogan_chainer_b200/datasets.py)."""
    from mocogan_chainer_b200.datasets import Uint8ClipCache
    rng = np.random.default_rng(seed)
    videos = [rng.integers(0, 256, size=(frames, 64, 64, 3), dtype=np.uint8) for _ in range(n_videos)]
    labels = rng.integers(0, 6, size=n_videos)
    np.random.seed(seed)
    return Uint8ClipCache(videos, labels, batch, video_length=16, extract_speed=2, shuffle=True, pin=True)


def build_updater(batch, seed, use_graph, model="normal", no_comm=False):
    import torch
    from mocogan_chainer_b200 import chainer, parallel
    from mocogan_chainer_b200 import random as mrandom
    from mocogan_chainer_b200.model import updater as updater_mod
    from mocogan_chainer_b200.model.net import ImageDiscriminator, ImageGenerator, VideoDiscriminator
    chainer.config.compute_dtype = "bf16"
    np.random.seed(0)
    out = 7 if model == "infogan" else 1
    G = ImageGenerator(50, 10, 6, 3, 64, 16)
    Di = ImageDiscriminator(3, out, 64, True, 0.2)
    Dv = VideoDiscriminator(3, out, 64, True, 0.2)
    opts = {}
    for name, m in (("image_gen", G), ("image_dis", Di), ("video_dis", Dv)):
        o = chainer.optimizers.Adam(alpha=2e-4, beta1=5e-5)   # train.py:93-101
        o.setup(m)
        o.add_hook(chainer.optimizer.WeightDecay(1e-5), "hook_dec")
        opts[name] = o
    parallel.attach(list(opts.values()))
    if no_comm:      # timing only: the same step with the gradient collectives taken out (replicas then diverge)
        for o in opts.values():
            o.grad_transform = None
    mrandom.set_source(mrandom.DeviceRandom(seed=seed, device="cuda", video_length=16))
    it = PinnedClipIterator(batch, seed)

    def concat(b):
        return (b.x, b.t) if isinstance(b, PreStacked) else chainer.dataset.concat_examples(b)

    updater_mod.concat_examples = concat
    up = updater_mod.Updater(model=model, models=(G, Di, Dv), video_length=16, img_size=64, channel=3, dim_zl=6,
                             tensorboard_writer=None, iterator=it, optimizer=opts, device=torch.cuda.current_device(),
                             use_graph=use_graph, graph_warmup=2)
    return up, it


def csrc_sha():
    """sha256 over the CUDA sources libmcg.so is built from: ties committed ncu figures to the build they came from."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "mocogan_chainer_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def load_traffic():
    """Measured DRAM bytes per launch (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum) of the kernels
    profiled under profiles/, keyed kernel:layer, together with the csrc sha of the build that was profiled.  A file
    whose sha differs from the sources in the tree is stale: its figures are NOT reported (traffic = null)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            with open(p) as f:
                d = json.load(f)
            sha = d.get("csrc_sha")
            if sha is not None and sha == csrc_sha():
                return d, name
            return {"stale": True, "csrc_sha": sha}, name
    return {}, None


def time_conv_layers(K, torch, peaks):
    """Per-kernel roofline: each tcgen05 convolution launch of the step timed ALONE with CUDA events (burst peak)."""
    layers = [  # name, N, Cin, Cout, in_sp, k, s, p, calls per step as (fprop, dgrad, wgrad) of the conv geometry
        ("Dv.dc2", 35, 64, 128, (13, 32, 32), (4, 4, 4), (1, 2, 2), (0, 1, 1), (2, 3, 2)),
        ("Dv.dc3", 35, 128, 256, (10, 16, 16), (4, 4, 4), (1, 2, 2), (0, 1, 1), (2, 3, 2)),
        ("Dv.dc4", 35, 256, 512, (7, 8, 8), (4, 4, 4), (1, 2, 2), (0, 1, 1), (2, 3, 2)),
        ("Di.dc2", 35, 64, 128, (1, 32, 32), (1, 4, 4), (1, 2, 2), (0, 1, 1), (2, 3, 2)),
        ("Di.dc3", 35, 128, 256, (1, 16, 16), (1, 4, 4), (1, 2, 2), (0, 1, 1), (2, 3, 2)),
        ("Di.dc4", 35, 256, 512, (1, 8, 8), (1, 4, 4), (1, 2, 2), (0, 1, 1), (2, 3, 2)),
        # generator deconvs, written as the conv they are the dgrad of: deconv fwd = dgrad, bwd-data = fprop
        ("G.dc2", 560, 256, 512, (1, 8, 8), (1, 4, 4), (1, 2, 2), (0, 1, 1), (1, 1, 1)),
        ("G.dc3", 560, 128, 256, (1, 16, 16), (1, 4, 4), (1, 2, 2), (0, 1, 1), (1, 1, 1)),
        ("G.dc4", 560, 64, 128, (1, 32, 32), (1, 4, 4), (1, 2, 2), (0, 1, 1), (1, 1, 1)),
    ]
    rows = []
    for name, N, Cin, Cout, in_sp, k, s, p, calls in layers:
        g = K.make_geom(N, Cin, Cout, in_sp, k, s, p)
        x = torch.randn((N,) + in_sp + (Cin,), device="cuda").bfloat16()
        w = (torch.randn((Cout,) + k + (Cin,), device="cuda") * 0.05).bfloat16()
        gy = torch.randn((N, g.To, g.Ho, g.Wo, Cout), device="cuda").bfloat16()
        y, dx = torch.empty_like(gy), torch.empty_like(x)
        dw = torch.zeros(w.shape, device="cuda")
        flops = 2.0 * N * g.To * g.Ho * g.Wo * Cout * Cin * k[0] * k[1] * k[2]
        fns = (("fprop", lambda: K.conv_fprop(g, x, w, None, y, K.IMPL_TC)),
               ("dgrad", lambda: K.conv_dgrad(g, gy, w, None, dx, K.IMPL_TC)),
               ("wgrad", lambda: K.conv_wgrad(g, x, gy, dw, K.IMPL_TC)))
        for (kind, fn), ncalls in zip(fns, calls):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            # best of 3 batches of 10 back-to-back launches: the burst peak these are divided by is itself a best-of-10
            # (MEASURED_PEAKS.json "how"), so both sides are the clocks-up figure of a kernel running alone
            reps, best = 10, None
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                t = e0.elapsed_time(e1) / reps
                best = t if best is None or t < best else best
            ms = best
            rows.append({"kernel": "tc_conv_%s" % kind, "layer": name, "ms": ms, "gflop": flops / 1e9,
                         "tflops": flops / ms / 1e9, "calls_per_step": ncalls,
                         "frac_of_burst_peak": flops / ms / 1e9 / peaks["bf16_burst"]})
    return rows


def time_narrow_layers(K, torch, peaks):
    """The 3-channel image layers (Dv.dc1, Di.dc1, G.dc5): 1.5 % of the step's FLOPs but HBM/L2-bound (Dv.dc1 writes 29.8 M
    outputs).  Each call timed ALONE; achieved = algorithmic bytes (input + output tensors of the call, once each) / time
    against the measured copy bandwidth.  A call includes its layout passes (row interleave / weight pad / depth-to-space)."""
    layers = [  # name, N, Cin, Cout, in_sp, k, s, p, calls per step (fprop, dgrad, wgrad) of the conv geometry
        ("Dv.dc1", 35, 3, 64, (16, 64, 64), (4, 4, 4), (1, 2, 2), (0, 1, 1), (2, 1, 2)),
        ("Di.dc1", 35, 3, 64, (1, 64, 64), (1, 4, 4), (1, 2, 2), (0, 1, 1), (2, 1, 2)),
        ("G.dc5", 560, 3, 64, (1, 64, 64), (1, 4, 4), (1, 2, 2), (0, 1, 1), (1, 1, 1)),
    ]
    rows = []
    for name, N, Cin, Cout, in_sp, k, s, p, calls in layers:
        g = K.make_geom(N, Cin, Cout, in_sp, k, s, p)
        x = torch.randn((N,) + in_sp + (Cin,), device="cuda").bfloat16()
        w = (torch.randn((Cout,) + k + (Cin,), device="cuda") * 0.05).bfloat16()
        gy = torch.randn((N, g.To, g.Ho, g.Wo, Cout), device="cuda").bfloat16()
        y, dx = torch.empty_like(gy), torch.empty_like(x)
        dw = torch.zeros(w.shape, device="cuda")
        ws = K.conv_fprop(g, x, w, None, y, K.IMPL_TC)
        xb, yb = x.numel() * 2, gy.numel() * 2
        fns = (("fprop", lambda: K.conv_fprop(g, x, w, None, y, K.IMPL_TC, ws=ws), xb + yb),
               ("dgrad", lambda: K.conv_dgrad(g, gy, w, None, dx, K.IMPL_TC), xb + yb),
               ("wgrad (x layout reused from fprop)", lambda: K.conv_wgrad(g, x, gy, dw, K.IMPL_TC, ws=ws, cols_valid=True), xb + yb))
        for (kind, fn, nbytes), ncalls in zip(fns, calls):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            rows.append({"kernel": "conv_%s" % kind, "layer": name, "ms": ms, "calls_per_step": ncalls,
                         "algorithmic_mb": nbytes / 1e6, "gb_per_s": nbytes / ms / 1e6,
                         "frac_of_hbm_peak": nbytes / ms / 1e6 / peaks["hbm_gbs"]})
    return rows


def time_stream_kernels(K, torch, peaks):
    """HBM-bound passes of the step ("which roofline": K5/K8/K10/K11), each timed ALONE with CUDA events on
    tensors larger than L2 (the generator's widest BatchNorm layer, G.bn4: 35*16 frames x 32x32 pixels x 64 channels =
    73 MB in bf16; Adam on the video discriminator's 11.06 M parameters).  achieved = ALGORITHMIC bytes (every operand
    read or written once) / time, against the measured copy bandwidth."""
    M, C = 35 * 16 * 32 * 32, 64
    dev = "cuda"
    y = torch.randn((M, C), device=dev).bfloat16()
    g = torch.randn((M, C), device=dev).bfloat16()
    out = torch.empty_like(y)
    vec = lambda v: torch.full((C,), v, device=dev)
    mean, invstd, scale, shift, gamma = vec(0.1), vec(0.9), vec(0.9), vec(-0.09), vec(1.0)
    dgam, dbet = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    am, av = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    n_adam = 11_060_000 // 8 * 8
    p32, g32 = torch.randn(n_adam, device=dev), torch.randn(n_adam, device=dev) * 1e-3
    m32, v32 = torch.zeros(n_adam, device=dev), torch.zeros(n_adam, device=dev)
    pb16 = torch.empty(n_adam, device=dev, dtype=torch.bfloat16)
    t_dev = torch.ones(1, dtype=torch.int32, device=dev)
    eb = y.element_size()
    cases = [
        ("bn_stats (colreduce + finalize)", "G.bn4 fwd", M * C * eb,
         lambda: K.bn_stats(y, M, C, gamma, shift, 2e-5, 0.9, mean, invstd, scale, shift, am, av)),
        ("affine_act_noise (BN apply + ReLU)", "G.bn4 fwd", 2 * M * C * eb,
         lambda: K.affine_act_noise(y, M, C, 32 * 32, scale, shift, K.ACT_RELU, 0.2, 0.0, None, None, None, 0, out)),
        ("act_bn_bwd_reduce (colreduce + finalize)", "G.bn4 bwd", 2 * M * C * eb,
         lambda: K.act_bn_bwd_reduce(g, y, M, C, mean, invstd, scale, shift, K.ACT_RELU, 0.2, dgam, dbet, None, None)),
        ("act_bn_bwd_apply", "G.bn4 bwd", 3 * M * C * eb,
         lambda: K.act_bn_bwd_apply(g, y, M, C, mean, invstd, gamma, scale, shift, K.ACT_RELU, 0.2, 0, dgam, dbet, out)),
        ("adam_kernel (Adam + WeightDecay + bf16 copy)", "video_dis", 30 * n_adam,
         lambda: K.adam_step(p32, g32, m32, v32, pb16, 2e-4, 5e-5, 0.999, 1e-8, 1e-5, 1.0, t_dev)),
    ]
    rows = []
    for name, layer, nbytes, fn in cases:
        # 20 back-to-back launches replayed from a CUDA graph: these kernels run 25-70 us, less than the host needs to issue
        # one through ctypes, so an eager loop would time the host
        reps = 20
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(reps):
                fn()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = nbytes / ms / 1e6
        rows.append({"kernel": name, "layer": layer, "ms": ms, "algorithmic_mb": nbytes / 1e6, "gb_per_s": gbs,
                     "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]})
    return rows


GEN_GF_PER_FRAME = 116.82 / 560.0     # generator forward 116.82 GF for 560 frames of 64x64


def gen_frames_per_s(torch, peaks, rank, world, barrier, batch=256, video_len=32, iters=8, size=64):
    """BASELINE config 5 (generate_samples inference): generator only, batch 256 clips of 32 frames PER GPU, BatchNorm
    in batch-statistics mode as generate_samples.py runs it, plus the uint8 / grid post-processing; the whole batch is
    one CUDA-graph replay.  size=64 is the reference's generator (net.py:115 hard-codes 64, 64);
    size=128 is the labelled extension (model/net128.py: one more deconvolution stage, parity unpinned)."""
    from mocogan_chainer_b200 import chainer, generate_samples
    from mocogan_chainer_b200 import random as mrandom
    chainer.config.compute_dtype = "bf16"
    np.random.seed(0)
    if size == 64:
        from mocogan_chainer_b200.model.net import ImageGenerator
        G = ImageGenerator(50, 10, 6, 3, 64, video_len)
        gf_per_frame = GEN_GF_PER_FRAME
    else:
        from mocogan_chainer_b200.model.net128 import ImageGenerator128
        G = ImageGenerator128(50, 10, 6, 3, 64, video_len)
        gf_per_frame = G.forward_gflop_per_frame()
    G.arena()
    src = mrandom.set_source(mrandom.DeviceRandom(seed=99 + rank, device="cuda", video_length=video_len))
    out = {}

    def once():
        src.begin_step()
        out["u8"], out["grid"] = generate_samples.generate(G, batch, grid=True)

    torch.cuda.empty_cache()      # the training legs before this one leave multi-GB caches and graph pools behind
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    graph = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            once()
        graph = g
    except Exception as e:   # noqa: BLE001 — an uncapturable generator is reported (cuda_graph: false), not hidden
        sys.stderr.write("gen: CUDA-graph capture failed (%s); timing eager launches\n" % (e,))
        torch.cuda.synchronize()
    run = graph.replay if graph is not None else once
    for _ in range(3):
        run()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / iters
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt[0])
    assert int(out["u8"].max()) > 0
    frames = batch * video_len * world
    return {"metric": "generator frames/s (generate_samples path: G forward + uint8/grid post-processing)",
            "value": frames / ms * 1e3, "unit": "frames/s (summed over ranks)", "n_gpus": world, "ms_per_batch": ms,
            "config": {"workload": "BASELINE config 5%s: generator only, batch %d clips x %d frames of %dx%d per GPU, BatchNorm "
                                   "batch statistics, bf16" % (" at the reference's native 64x64" if size == 64 else
                                                                " as worded (128x128): EXTENSION architecture, parity unpinned",
                                                                batch, video_len, size, size),
                       "cuda_graph": graph is not None},
            "achieved_tflops_per_gpu": batch * video_len * gf_per_frame / ms,
            "frac_of_sustained_peak": batch * video_len * gf_per_frame / ms / peaks["bf16_sustained"]}


def time_steps(torch, up, x_dev, t_dev, n, barrier):
    """n replays of the captured step, bracketed by barrier + synchronize; device time in ms (this rank)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(n):
        up.step_host_inputs(x_dev, t_dev)
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def max_over_ranks(torch, world, vals):
    if world == 1:
        return [float(v) for v in vals]
    import torch.distributed as dist
    tt = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return [float(v) for v in tt]


def train_leg(torch, K, parallel, args, model, rank, world, barrier, sampler=None, no_comm=False):
    """Builds the three networks + Updater for `model`, captures the step, does W warm-up replays and times K steps."""
    up, it = build_updater(BATCH, parallel.shard_seed(1234, rank), use_graph=not args.no_graph, model=model, no_comm=no_comm)
    x_dev = it.x[0].cuda()
    t_dev = it.t[0].cuda()
    # set-up (not warm-up): one eager step to count our kernel launches per step, then graph_warmup eager steps + capture
    launches0 = K.launch_count()
    up.step_host_inputs(x_dev, t_dev)
    launches_per_step = K.launch_count() - launches0
    while not args.no_graph and up._graph is None:
        up.step_host_inputs(x_dev, t_dev)
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        up.step_host_inputs(x_dev, t_dev)
    if sampler is not None:
        sampler.start()
        time.sleep(0.3)
    ms = time_steps(torch, up, x_dev, t_dev, args.steps, barrier)
    return up, it, x_dev, t_dev, ms, launches_per_step, warmup


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mocogan_chainer_b200 import kernels as K
    from mocogan_chainer_b200 import parallel
    rank, world = parallel.init_from_env()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    peaks = load_peaks()
    K.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed: inputs resident in HBM, K steps bracketed by barrier + synchronize, max over ranks
    sampler = ClockSampler(local) if rank == 0 else None
    up, it, x_dev, t_dev, ms_dev, launches_per_step, warmup = train_leg(torch, K, parallel, args, args.model, rank, world,
                                                                        barrier, sampler)
    # ---- sustained: the same replay loop for >= 2 s (>= 5 windows of >= 100 steps), median window, its own clock record
    sustained = None
    if not args.no_sustained:
        clocks_short = sampler.stop() if rank == 0 else None
        sus_sampler = ClockSampler(local) if rank == 0 else None
        if rank == 0:
            sus_sampler.start()
        win_steps, wins = max(100, args.steps), []
        for _ in range(args.sustained_windows):
            wins.append(max_over_ranks(torch, world, [time_steps(torch, up, x_dev, t_dev, win_steps, barrier)])[0])
        sus_clocks = sus_sampler.stop() if rank == 0 else None
        med = float(np.median(wins))
        sustained = {"value": world * win_steps / (med / 1e3), "unit": "steps/s", "ms_per_step": med / win_steps,
                     "windows": len(wins), "steps_per_window": win_steps, "window_ms": wins, "seconds": sum(wins) / 1e3,
                     "stat": "median window, max over ranks per window", "clocks": sus_clocks}
        if rank == 0:
            sampler = ClockSampler(local)
            sampler.start()
    else:
        clocks_short = None
    # ---- end to end through the public API: Updater.update() pulling uint8 clips from the pinned clip cache (iterator
    # -> sub-sequence draw -> pinned staging batch -> H2D -> step), losses read back each step
    cache = make_clip_cache(BATCH, parallel.shard_seed(1234, rank))
    up._iterators["main"] = cache
    up._static = up._stage = up._copy_stream = up._graph = None     # new input dtype/layout: re-stage and re-capture
    up._eager_steps = 0
    for _ in range(4):
        up.update()
    barrier()
    t0 = time.perf_counter()
    loss_sink = 0.0
    for _ in range(args.steps):
        up.update()
        loss_sink += sum(float(v) for v in up.losses.values())   # D2H of the step's three losses (synchronises)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    if clocks_short is not None and clocks_short.get("samples"):
        clocks = clocks_short if not clocks or not clocks.get("samples") else {
            "sm_mhz": float(np.median([clocks_short["sm_mhz"], clocks["sm_mhz"]])),
            "sm_max_mhz": max(clocks_short["sm_max_mhz"], clocks["sm_max_mhz"]),
            "reasons": sorted(set(clocks_short["reasons"]) | set(clocks["reasons"])),
            "samples": clocks_short["samples"] + clocks["samples"], "device_timed": clocks_short, "e2e": clocks}
    ms_dev, ms_e2e = max_over_ranks(torch, world, [ms_dev, ms_e2e])
    assert K.tc_error_flag() == 0, "a tcgen05 kernel reported an mbarrier timeout"
    assert np.isfinite(loss_sink), "non-finite loss"
    # ---- replicas: every rank must hold bit-identical parameters after the timed loops (BatchNorm running stats are local)
    identical = parallel.replicas_identical([up.image_gen, up.image_dis, up.video_dis])
    assert identical, "data-parallel replicas diverged"
    exposed = None

    # ---- per-kernel tables (rank 0): every launch timed ALONE, before the long legs below heat the part into its power cap
    layer_rows, stream_rows, narrow_rows, dominant = None, None, None, None
    if rank == 0:
        time.sleep(0.5)
        layer_rows = time_conv_layers(K, torch, peaks)
        stream_rows = time_stream_kernels(K, torch, peaks)
        narrow_rows = time_narrow_layers(K, torch, peaks)
        tot = {}
        for r in layer_rows:
            tot[(r["kernel"], r["layer"])] = r["ms"] * r["calls_per_step"]
        dk = max(tot, key=tot.get)
        dominant = [r for r in layer_rows if (r["kernel"], r["layer"]) == dk][0]
    barrier()

    # ---- BASELINE config 4 (infogan) on the same N GPUs, same timing rules
    other = None
    if not args.no_other_model:
        om = "infogan" if args.model == "normal" else "normal"
        del up, cache
        torch.cuda.empty_cache()
        up2, _, _, _, ms2, lps2, _ = train_leg(torch, K, parallel, args, om, rank, world, barrier)
        ms2 = max_over_ranks(torch, world, [ms2])[0]
        ident2 = parallel.replicas_identical([up2.image_gen, up2.image_dis, up2.video_dis])
        other = {"metric": "MoCoGAN train steps/s (bs35, 16x3x64x64)", "model": om, "value": world * args.steps / (ms2 / 1e3),
                 "unit": "steps/s (summed over ranks)", "n_gpus": world, "ms_per_step": ms2 / args.steps, "steps": args.steps,
                 "workload": "BASELINE config %s: MoCoGAN %s model, clips (35,3,16,64,64) per GPU" % ("4" if om == "infogan" else "2", om),
                 "gpu_launches_per_step": int(lps2), "replicas_identical": bool(ident2)}
        del up2
        torch.cuda.empty_cache()
    # ---- what the collectives cost a step: the same captured step on the same GPUs with the all-reduces taken out
    if world > 1 and not args.no_other_model:
        up3, _, _, _, ms3, _, _ = train_leg(torch, K, parallel, args, args.model, rank, world, barrier, no_comm=True)
        ms3 = max_over_ranks(torch, world, [ms3])[0]
        exposed = {"ms_per_step_without_collectives": ms3 / args.steps, "exposed_collective_ms": (ms_dev - ms3) / args.steps,
                   "note": "same run, same GPUs, gradient all-reduces (and their bf16 casts) removed from the captured step; "
                           "the difference is what the three collectives add to the critical path"}
        del up3
        torch.cuda.empty_cache()
    gen = gen128 = None
    if not args.no_gen:
        gen = gen_frames_per_s(torch, peaks, rank, world, barrier)
        try:
            gen128 = gen_frames_per_s(torch, peaks, rank, world, barrier, size=128, iters=4)
        except ImportError:
            gen128 = None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_config1()
    if rank != 0:
        return
    steps_per_s = world * args.steps / (ms_dev / 1e3)
    e2e_steps_per_s = world * args.steps / (ms_e2e / 1e3)
    conv_ms = sum(r["ms"] * r["calls_per_step"] for r in layer_rows)
    traffic, traffic_file = load_traffic()
    tkey = "%s:%s" % (dominant["kernel"], dominant["layer"])
    cfg_no = "4" if args.model == "infogan" else "2"
    line = {
        "metric": "MoCoGAN train steps/s (bs35, 16x3x64x64)", "value": steps_per_s,
        "unit": "steps/s (batch-35 update_core steps, summed over ranks)", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "BASELINE config %s: MoCoGAN %s model (train.py MUG wiring), clips (35,3,16,64,64) "
                               "per GPU, one update_core (G+Di+Dv) per step" % (cfg_no, args.model) +
                               ("; data-parallel, NCCL all-reduce of the three flat gradient buffers" if world > 1 else ""),
                   "global_batch": BATCH * world, "parallelism": "dp%d" % world, "dp_overlap": parallel.describe() if world > 1 else None,
                   "cuda_graph": not args.no_graph,
                   "l2": "no explicit flush: one step streams > 1 GB of activations/weights, >> 126 MB L2",
                   "weights": "random init (GlorotNormal/LeCunNormal)", "noise": "device Philox",
                   "setup": "1 eager step (launch count) + 2 eager + graph capture before the %d warm-up replays" % warmup},
        "clocks": clocks,
        "sustained": sustained,
        "replicas_identical": bool(identical),
        "exposed_collective_ms": exposed,
        "e2e": {"value": e2e_steps_per_s, "unit": "steps/s",
                "h2d_bytes_per_step": int(BATCH * 16 * 64 * 64 * 3 + BATCH * 4),
                "d2h_bytes_per_step": 12, "ms_per_step": ms_e2e / args.steps,
                "input": "uint8 clips from a pinned host cache, normalised on the device (datasets.py:72-104 restated)"},
        "gpu_launches": int(launches_per_step * args.steps),
        "roofline": {"bound": "tensor", "kernel": dominant["kernel"], "layer": dominant["layer"],
                     "achieved": dominant["tflops"], "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                     "frac": dominant["frac_of_burst_peak"],
                     "traffic": None if traffic.get("stale") else traffic.get(tkey),
                     "traffic_source": {"file": traffic_file, "csrc_sha_now": csrc_sha(), "csrc_sha_profiled": traffic.get("csrc_sha"),
                                        "stale": bool(traffic.get("stale"))},
                     "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                     "algorithmic_flop_per_launch": dominant["gflop"] * 1e9, "peak_source": peaks["src"] + " (burst: kernel timed alone)",
                     "step": {"useful_gflop": USEFUL_GF_PER_STEP, "achieved_tflops": USEFUL_GF_PER_STEP * steps_per_s / world / 1e3,
                              "peak_sustained": peaks["bf16_sustained"],
                              "frac": USEFUL_GF_PER_STEP * steps_per_s / world / 1e3 / peaks["bf16_sustained"],
                              "tc_conv_ms_per_step_isolated": conv_ms,
                              "tc_conv_frac_of_burst_isolated": sum(r["gflop"] * r["calls_per_step"] for r in layer_rows) / conv_ms / peaks["bf16_burst"]},
                     "layers": layer_rows,
                     "narrow_layers": narrow_rows,
                     "hbm_kernels": {"peak_gb_per_s": peaks["hbm_gbs"], "peak_source": peaks["src"] + " (copy bandwidth)",
                                     "kernels": stream_rows}},
        "cpu_baseline": cpu,
        "other_model": other,
        "gen": gen,
        "gen128": gen128,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ CPU legs
def oracle_stepper(config, batch, as_executed=True, seed=0):
    """The NumPy restatement of the reference's CPU step (im2col + BLAS, float32, three full backward passes) as a
    callable: step(i) runs the i-th update_core on the same models and returns its wall time in seconds."""
    from oracle import mocogan_ref as ref
    model, G, Di, Dv = ref.build_models(config, dtype=np.float32, seed=seed)
    up = ref.Updater(model, G, Di, Dv)
    C = G.out_channels
    x = np.random.default_rng(1234).uniform(-1, 1, size=(batch, C, 16, 64, 64)).astype(np.float32)
    t_real = np.random.default_rng(5).integers(0, 6, size=batch) if G.dim_zl else None

    def step(i):
        t0 = time.perf_counter()
        # float64 randn then cast, as add_noise does (net.py:13); drawing is part of the reference's step
        r = ref.draw_step_randoms(np.random.default_rng(100 + i), np.random.default_rng(200 + i), G, Di, Dv, batch, x.shape,
                                  dtype=np.float32)
        up.update_core(x, t_real, r, as_executed=as_executed)
        return time.perf_counter() - t0

    return step


def oracle_step_time(config, batch, n_steps, as_executed=True, seed=0):
    """One warm-up step, then n_steps timed ones (n_steps = 0: the warm-up step's own time)."""
    step = oracle_stepper(config, batch, as_executed, seed)
    times = [step(i) for i in range(n_steps + 1)]
    return times[1:] if n_steps > 0 else times


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
        return int(n)
    except Exception:
        return os.cpu_count() or 1


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms are meant to use all the host threads they can, so
    the BLAS pools are set explicitly.  Returns the context manager that holds the limit (keep it alive)."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        return None


def cpu_baseline_config1(n_timed=3):
    """BASELINE config 1 (the reference's own CPU-runnable case): batch 8, (16,1,64,64), as-executed steps; median of
    `n_timed` timed steps after one warm-up step."""
    hold = use_all_host_cores()
    times = oracle_step_time("mnist_normal", 8, n_timed, as_executed=True)
    s = float(np.median(times))
    del hold
    return {"value": 1.0 / s, "unit": "steps/s (batch-8 update_core steps, config 1)", "cores": cpu_threads(),
            "kind": "port", "sample": "1 warm-up + %d timed as-executed update_core steps of BASELINE config 1 (batch 8, "
            "clips (16,1,64,64), float32 NumPy/BLAS restatement of the Chainer v3.1.0 CPU path; median %.1f s, all %s)"
            % (n_timed, s, ["%.1f" % t for t in times]), "host_cpus": os.cpu_count()}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path on the host cores (oracle port — genuine Chainer
    3.1.0 cannot be imported here), at BASELINE config 2's TRUE batch of 35 clips — no extrapolation.
    One as-executed step is ~40 s on the box's cores, so the number of timed steps is bounded by a time budget
    (>= 2 timed steps after one warm-up step) instead of taking --steps literally; `steps` reports what ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    hold = use_all_host_cores()
    model = "mug_infogan" if args.model == "infogan" else "mug_normal"
    step = oracle_stepper(model, BATCH, as_executed=True)
    t_probe = step(0)                                                  # the warm-up step, timed to size the run
    budget = float(os.environ.get("MCG_REF_BUDGET_S", "150"))
    n = max(2, min(args.steps, int(budget / max(t_probe, 1e-3))))
    times = [step(1 + i) for i in range(n)]
    s = float(np.median(times))
    v = 1.0 / s
    del hold
    line = {"impl": "reference", "metric": "MoCoGAN train steps/s (bs35, 16x3x64x64)", "value": v,
            "unit": "steps/s (batch-35 update_core steps, summed over ranks)", "n_gpus": args.gpus, "steps": n,
            "warmup": 1, "ms_per_step": 1e3 * s, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE config %s: MoCoGAN %s model (train.py MUG wiring), clips (35,3,16,64,64), "
                                   "one as-executed update_core per step, CPU" % ("4" if args.model == "infogan" else "2", args.model),
                       "global_batch": BATCH, "note": "rank 0 only; a CPU step does not shard over GPUs"},
            "cpu_baseline": {"value": v, "unit": "steps/s", "cores": cpu_threads(), "kind": "port",
                             "sample": "1 warm-up + %d timed steps at the full batch of 35 (median %.1f s, all %s); float32 "
                                       "NumPy/BLAS restatement of the Chainer v3.1.0 CPU path incl. its discarded backward work"
                                       % (n, s, ["%.1f" % t for t in times]), "host_cpus": os.cpu_count()},
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", choices=["normal", "infogan"], default="normal",
                    help="normal = BASELINE config 2 (the headline); infogan = config 4.  The other one is timed as `other_model`.")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained leg")
    ap.add_argument("--sustained-windows", type=int, default=5)
    ap.add_argument("--no-other-model", action="store_true", help="skip the second model's (config 4) throughput leg")
    ap.add_argument("--no-gen", action="store_true", help="skip the generator frames/s legs (config 5)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    main()
