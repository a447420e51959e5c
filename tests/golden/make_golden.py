#!/usr/bin/env python
"""Generates tests/golden/step_*.npz: small-configuration outputs of the NumPy oracle (float64) for one update_core.

The reference cannot be imported here (chainer==3.1.0 is neither installed nor installable), so these
vectors pin the ORACLE — itself cross-checked against torch float64 autograd in tests/test_oracle_vs_torch.py — against
drift, and give the GPU tests a committed, seed-independent target.  Inputs are regenerated from the seeds recorded in
the file; only outputs are stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mocogan_ref as ref  # noqa: E402

GRAD_SAMPLE = 4096   # gradients are stored whole up to this many elements, beyond that as every k-th element


def grad_sample(g):
    """The elements of a gradient tensor the fixtures keep: all of it, or a fixed stride through the flattened array."""
    flat = np.asarray(g).reshape(-1)
    return flat[::max(1, -(-flat.size // GRAD_SAMPLE))]


CASES = {"mnist_normal": dict(config="mnist_normal", nf=8, N=2), "mug_infogan": dict(config="mug_infogan", nf=8, N=3),
         "mug_cgan": dict(config="mug_cgan", nf=8, N=2)}


def make_inputs(config, nf, N, dtype=np.float64):
    model, G, Di, Dv = ref.build_models(config, dtype=dtype, seed=3, n_filters=nf)
    prng = np.random.default_rng(11)
    for net in (G, Di, Dv):
        for k, v in net.params.items():
            if k.endswith("/b") or k.endswith("beta"):
                v += 0.05 * prng.standard_normal(v.shape)
            if k.endswith("gamma"):
                v += 0.1 * prng.standard_normal(v.shape)
        for k in net.params:  # the weights both sides load are float32-representable
            net.params[k] = net.params[k].astype(np.float32).astype(dtype)
    C = G.out_channels
    x_real = np.random.default_rng(1234).uniform(-1, 1, size=(N, C, 16, 64, 64)).astype(np.float32)
    t_real = np.random.default_rng(5).integers(0, 6, size=N) if G.dim_zl else None
    r = ref.draw_step_randoms(np.random.default_rng(100), np.random.default_rng(200), G, Di, Dv, N, x_real.shape, t=7,
                              dtype=np.float32)
    return model, G, Di, Dv, x_real, t_real, r


def run_case(config, nf, N):
    model, G, Di, Dv, x_real, t_real, r = make_inputs(config, nf, N)
    trace = {}
    losses = ref.Updater(model, G, Di, Dv).update_core(x_real.astype(np.float64), t_real, r, trace=trace)
    out = {"loss_" + k.replace("/", "_"): np.float64(v) for k, v in losses.items()}
    out["x_fake_sample"] = trace["x_fake"][::5, :, :, ::16, ::16].astype(np.float32)
    out["x_fake_mean_std"] = np.array([trace["x_fake"].mean(), trace["x_fake"].std()])
    for name, key in (("g", "grads_g"), ("di", "grads_di"), ("dv", "grads_dv")):
        for k, v in trace[key].items():
            out["gradnorm_%s_%s" % (name, k.replace("/", "_"))] = np.float64(np.linalg.norm(v))
            out["grad_%s_%s" % (name, k.replace("/", "_"))] = grad_sample(v).astype(np.float32)
    for key in ("y_real_i", "y_real_v", "y_fake_i", "y_fake_v"):
        out[key] = trace[key].astype(np.float32)
    for name, net in (("g", G), ("di", Di), ("dv", Dv)):
        out["post_%s_dc5_W" % name] = net.params["dc5/W"].astype(np.float32)
        for k, v in net.persistent.items():
            out["running_%s_%s" % (name, k.replace("/", "_"))] = v.astype(np.float32)
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name, kw in CASES.items():
        out = run_case(**kw)
        np.savez_compressed(os.path.join(here, "step_%s.npz" % name), **out)
        print(name, {k: float(v) for k, v in out.items() if k.startswith("loss_")})
