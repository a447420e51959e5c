#!/usr/bin/env python
"""Prints the loss-deviation curves behind tests/test_step_gpu.py::test_losses_track_oracle_over_100_steps: strict fp32
device path and float32 oracle, each against the float64 oracle, 100 free-running steps."""
import sys, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_step_gpu import build_pair  # noqa: E402
from oracle import mocogan_ref as ref
from mocogan_chainer_b200 import random as mrandom
model, (G, Di, Dv), (oG, oI, oV), up, oup = build_pair("mug_normal", 8, "fp32")
# float32 oracle from the same start
m32, g32, i32, v32 = ref.build_models("mug_normal", dtype=np.float64, seed=3, n_filters=8)
for net, src in ((g32, oG), (i32, oI), (v32, oV)):
    net.dtype = np.float32
    net.params = {k: v.astype(np.float32) for k, v in src.params.items()}
    net.persistent = {k: v.astype(np.float32) for k, v in net.persistent.items()}
o32 = ref.Updater(m32, g32, i32, v32)
N, C = 2, oG.out_channels
for step in range(100):
    x_real = np.random.default_rng(1234 + step).uniform(-1, 1, size=(N, C, 16, 64, 64)).astype(np.float32)
    t_real = np.random.default_rng(5 + step).integers(0, 6, size=N)
    r = ref.draw_step_randoms(np.random.default_rng(100 + step), np.random.default_rng(200 + step), oG, oI, oV, N, x_real.shape, t=(7 + 3 * step) % 16, dtype=np.float32)
    mrandom.set_source(mrandom.InjectedRandom(r))
    up.step_on_device(torch.from_numpy(x_real).cuda(), torch.from_numpy(t_real).int().cuda())
    l64 = oup.update_core(x_real.astype(np.float64), t_real, r)
    l32 = o32.update_core(x_real, t_real, r)
    names = (("ImageDiscriminator", "image_dis/loss"), ("VideoDiscriminator", "video_dis/loss"), ("ImageGenerator", "image_gen/loss"))
    d_dev = max(abs(float(up.losses[a]) - l64[b]) for a, b in names)
    d_32 = max(abs(float(l32[b]) - l64[b]) for a, b in names)
    if step < 10 or step % 10 == 9:
        print(step, "dev-vs-f64 %.2e  oracle32-vs-f64 %.2e  losses %s" % (d_dev, d_32, [round(float(l64[b]), 4) for a, b in names]), flush=True)
