"""Whole-step parity: Updater.step_on_device (the CUDA path behind the reference's update_core) against the
NumPy oracle's update_core on identical weights and identical injected random tensors.

Checked per step: the generator's clip, all four discriminator outputs, every layer's input and convolution output
and the BatchNorm batch statistics of the five network calls (`check_forward`, against trace["cache_*"]), the three
losses, every parameter gradient of passes A, B, C (including the stale-activation / fresh-weight semantics of pass
C), the BatchNorm running statistics, and the post-step weights (Adam + WeightDecay).
Tolerances (max|a-b| / max|b| per tensor): fp32 mode 1e-5 on losses, outputs and activations, 1e-4 on gradients (which
chain ~10 fp32 kernels); bf16 mode 2e-2 on losses, outputs and activations.  bf16 whole-network gradients: bf16
storage flips (Leaky)ReLU masks and the discriminator losses back-propagate from ONE sample (updater.py:25-26), so the
element-wise bound is per parameter group (BF16_GRAD_BOUNDS: what was measured on B200 plus margin) and every tensor
must also agree in direction; the <= 2e-2 per-layer bound on identical inputs lives in test_kernels_gpu.py.
"""
import json
import os
import numpy as np
import pytest
import torch

from oracle import chainer_ops as ops
from oracle import mocogan_ref as ref

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def chain_nodes(var, stop_at=None):
    """FunctionNodes on the first-input chain that produced `var`, first executed first (optionally from the last node
    of type `stop_at` on)."""
    nodes, v = [], var
    while v is not None and v.creator_node is not None:
        n = v.creator_node
        nodes.append(n)
        if stop_at is not None and isinstance(n, stop_at):
            break
        v = n.inputs[0]
    return nodes[::-1]


def logical(t5):
    """channels-last (N,T,H,W,C) device storage -> float64 NumPy (N,C,T,H,W)."""
    return t5.float().permute(0, 4, 1, 2, 3).cpu().numpy().astype(np.float64)


# Two kinds of tensor get a stated slack over the per-layer tolerance (measured on B200, profiles/r02_parity_report.json):
#  * the four logits: an 8,192- to 32,768-term dot product of O(1) terms that cancels to O(0.1), at the END of five
#    chained layers — fp32 mode reaches 1.4e-5 of max|y| (mnist shapes) where every layer tensor stays below 7e-6;
#  * x_fake: the output of five chained bf16 layers compared on tanh's absolute scale (max = 1): 1.8e-2 .. 2.3e-2, while
#    every single layer's input and output stays below 1.2e-2.
FORWARD_SLACK = {"y_real_i": 3.0, "y_real_v": 3.0, "y_fake_i": 3.0, "y_fake_v": 3.0, "x_fake": 1.5}


def check_forward(up, trace, oG, tol, report=None):
    """x_fake, the four discriminator outputs and every layer of the five network calls against the oracle's trace:
    the input each convolution read (previous activation + add_noise), its output (pre-BatchNorm, bias included) and
    the batch statistics BatchNorm normalised with."""
    from mocogan_chainer_b200.chainer import functions as F
    fw = up.last_forward
    errs = {}

    def cmp(tag, got, want):
        want = np.asarray(want, np.float64)
        got = np.asarray(got, np.float64).reshape(want.shape)
        errs[tag] = relerr(got, want)

    C = oG.out_channels
    xf = fw["x_fake"].data.float().permute(2, 0, 1, 3, 4)[:, :, :C].cpu().numpy()     # (N,C[+zl],T,H,W) -> (T,N,C,H,W)
    cmp("x_fake", xf, trace["x_fake"])
    for key in ("y_real_i", "y_real_v", "y_fake_i", "y_fake_v"):
        cmp(key, fw[key].data.float().cpu().numpy(), trace[key])
    for key, cache in (("y_real_i", "cache_ri"), ("y_real_v", "cache_rv"), ("y_fake_i", "cache_fi"), ("y_fake_v", "cache_fv")):
        nodes = chain_nodes(fw[key], stop_at=F.PackVideo)
        convs = [n for n in nodes if isinstance(n, F.ConvolutionND)]
        bns = [n for n in nodes if isinstance(n, F.BNActNoise)]
        assert len(convs) == 5 and len(bns) == 4, (key, [type(n).__name__ for n in nodes])
        acts = trace[cache]["acts"]
        for i, cv in enumerate(convs):
            h = acts[i][0]
            got = logical(cv.xp)
            cmp("%s.dc%d.in" % (key, i + 1), got[:, :, 0] if h.ndim == 4 else got, h)
        for i, bn in enumerate(bns):
            y, stats = acts[i][1], acts[i][2]
            got = logical(bn.yp)
            cmp("%s.dc%d.out" % (key, i + 1), got[:, :, 0] if y.ndim == 4 else got, y)
            if stats is not None:
                cmp("%s.bn%d.mean" % (key, i + 1), bn.mean.cpu().numpy(), stats[0])
                cmp("%s.bn%d.std" % (key, i + 1), 1.0 / bn.invstd.cpu().numpy(), stats[1])
    nodes = chain_nodes(fw["x_fake"])
    convs = [n for n in nodes if isinstance(n, F.ConvolutionND)]
    bns = [n for n in nodes if isinstance(n, F.BNActNoise)]
    assert len(convs) == 5 and len(bns) == 5, [type(n).__name__ for n in nodes]
    acts = trace["cache_g"]["acts"]
    for i, cv in enumerate(convs):
        x = acts[i][0]
        cmp("G.dc%d.in" % (i + 1), logical(cv.xp)[:, :x.shape[1], 0], x)       # dc1: latent zero-padded 60 -> 64 channels
    for i, bn in enumerate(bns):
        y, stats = acts[i][1], acts[i][2]
        cmp("G.dc%d.out" % (i + 1), logical(bn.yp)[:, :, 0], y)
        if stats is not None:
            cmp("G.bn%d.mean" % (i + 1), bn.mean.cpu().numpy(), stats[0])
            cmp("G.bn%d.std" % (i + 1), 1.0 / bn.invstd.cpu().numpy(), stats[1])
    if report is not None:
        report.setdefault("forward", []).append(errs)
    bad = {k: v for k, v in errs.items() if not v < tol * FORWARD_SLACK.get(k, 1.0)}
    assert not bad, ("forward tensors beyond %g" % tol, bad)
    return errs


# bf16 whole-step gradients.  What was measured on B200 over the bf16 cases of this file (every tensor is listed in
# profiles/r02_parity_report.json), worst case per network as (max|g - g_ref| / max|g_ref|, L2 relative error, cosine):
#   ImageGenerator      0.49  0.24  0.972     (g0/*, dc1/W worst: the end of the longest chain, ~20 bf16 tensors deep)
#   ImageDiscriminator  0.35  0.14  0.992     (dc5/b, one element, is judged absolutely: a difference of two ~0.5/N terms)
#   VideoDiscriminator  0.28  0.16  0.988
# Only the last layers (dc5/W, bn4/gamma: 1e-2) stay inside the per-layer 2e-2: every earlier gradient has passed through
# BatchNorm backward's gy - mean(gy) - xhat*mean(gy*xhat), which cancels most of gy and leaves bf16's 2^-9 storage rounding
# of gy (and the (Leaky)ReLU masks flipped by the rounded forward activations) as a 5-25 % L2 perturbation — the
# discriminator losses back-propagate from ONE sample (updater.py:25-26), so nothing averages it out.  The bounds are the
# measured worst case with ~1.4x margin; the <= 2e-2 per-layer bound on identical inputs lives in test_kernels_gpu.py.
BF16_GRAD_BOUNDS = {"ImageGenerator": (0.70, 0.33), "ImageDiscriminator": (0.50, 0.20), "VideoDiscriminator": (0.40, 0.22)}
BF16_GRAD_COS = 0.96


def dump_report(tag, report):
    path = os.environ.get("MCG_PARITY_REPORT")
    if path:
        data = {}
        if os.path.exists(path):
            with open(path) as f:
                data = json.load(f)
        data[tag] = report
        with open(path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)


def build_pair(config, n_filters, dtype_mode, seed=3):
    from mocogan_chainer_b200 import chainer
    from mocogan_chainer_b200.model.net import ImageDiscriminator, ImageGenerator, VideoDiscriminator
    from mocogan_chainer_b200.model.updater import Updater
    chainer.config.compute_dtype = dtype_mode
    model, oG, oI, oV = ref.build_models(config, dtype=np.float64, seed=seed, n_filters=n_filters)
    prng = np.random.default_rng(11)
    for net in (oG, oI, oV):  # non-trivial biases / gamma / beta so every term is exercised
        for k, v in net.params.items():
            if k.endswith("/b") or k.endswith("beta"):
                v += 0.05 * prng.standard_normal(v.shape)
            if k.endswith("gamma"):
                v += 0.1 * prng.standard_normal(v.shape)
    C = oG.out_channels
    G = ImageGenerator(50, 10, oG.dim_zl, C, n_filters, 16)
    Di = ImageDiscriminator(oI.in_channels, oI.out_channels, n_filters, True, 0.2)   # cgan: clip + label planes
    Dv = VideoDiscriminator(oV.in_channels, oV.out_channels, n_filters, True, 0.2)
    for mine, theirs in ((G, oG), (Di, oI), (Dv, oV)):
        mine.arena()
        for path, p in mine.namedparams():
            p.data = theirs.params[path.lstrip("/")].astype(np.float32)
        # the oracle continues from the fp32-rounded weights the device holds
        for k in theirs.params:
            theirs.params[k] = theirs.params[k].astype(np.float32).astype(np.float64)

    def make_opt(m):
        o = chainer.optimizers.Adam(alpha=2e-4, beta1=5e-5)
        o.setup(m)
        o.add_hook(chainer.optimizer.WeightDecay(1e-5), 'hook_dec')
        return o

    class _It(object):
        epoch, is_new_epoch, epoch_detail = 0, False, 0.0

    up = Updater(model=model, models=(G, Di, Dv), video_length=16, img_size=64, channel=C, dim_zl=oG.dim_zl,
                 tensorboard_writer=None, iterator=_It(),
                 optimizer={'image_gen': make_opt(G), 'image_dis': make_opt(Di), 'video_dis': make_opt(Dv)}, device=0,
                 keep_forward=True)
    oup = ref.Updater(model, oG, oI, oV)
    return model, (G, Di, Dv), (oG, oI, oV), up, oup


def run_step_case(config, n_filters, N, dtype_mode, tol_out, tol_grad, steps=1, ref32=False, override=True, tol_fwd=None,
                  tag=None):
    """override=False: pass C is compared WITHOUT handing the device's updated discriminator weights to the oracle, so
    the oracle's pass C runs on the weights its own passes A/B produced (a pass-B -> pass-C ordering bug on the device
    cannot hide behind the hand-over); tol_grad then has to absorb Adam's amplification of round-off (see below)."""
    from mocogan_chainer_b200 import kernels as K
    from mocogan_chainer_b200 import random as mrandom
    model, (G, Di, Dv), (oG, oI, oV), up, oup = build_pair(config, n_filters, dtype_mode)
    C = oG.out_channels
    report = {"config": config, "n_filters": n_filters, "N": N, "dtype": dtype_mode, "grads": []}
    try:
        _run_steps(config, n_filters, N, dtype_mode, tol_out, tol_grad, steps, ref32, override, tol_fwd, report, K, mrandom,
                   model, (G, Di, Dv), (oG, oI, oV), up, oup)
    finally:
        dump_report(tag or "%s_nf%d_N%d_%s" % (config, n_filters, N, dtype_mode), report)
    return up


def _run_steps(config, n_filters, N, dtype_mode, tol_out, tol_grad, steps, ref32, override, tol_fwd, report, K, mrandom, model,
               nets, onets, up, oup):
    (G, Di, Dv), (oG, oI, oV) = nets, onets
    C = oG.out_channels
    for step in range(steps):
        x_real = np.random.default_rng(1234 + step).uniform(-1, 1, size=(N, C, 16, 64, 64)).astype(np.float32)
        t_real = np.random.default_rng(5 + step).integers(0, 6, size=N) if oG.dim_zl else None
        r = ref.draw_step_randoms(np.random.default_rng(100 + step), np.random.default_rng(200 + step), oG, oI, oV, N,
                                  x_real.shape, t=(7 + 3 * step) % 16, dtype=np.float32)
        pre = {id(n): {k: v.copy() for k, v in n.params.items()} for n in (oG, oI, oV)}

        mrandom.set_source(mrandom.InjectedRandom(r))
        xr = torch.from_numpy(x_real).cuda()
        tr = None if t_real is None else torch.from_numpy(t_real).int().cuda()
        up.step_on_device(xr, tr)
        torch.cuda.synchronize()
        assert K.tc_error_flag() == 0
        # pass C is judged on identical updated discriminator weights (see oracle.update_core docstring)
        d_override = {key: {path.lstrip("/"): p.data.float().cpu().numpy().astype(np.float64) for path, p in m.namedparams()}
                      for key, m in (("image_dis", Di), ("video_dis", Dv))}
        if not override:
            d_override = None
        trace = {}
        trace32 = None
        if dtype_mode == "fp32" and ref32:
            # the same step through the oracle in float32 from the same pre-step state: the reference's own precision
            m32, g32, i32, v32 = ref.build_models(config, dtype=np.float64, seed=3, n_filters=n_filters)
            for net, src in ((g32, oG), (i32, oI), (v32, oV)):
                net.dtype = np.float32
                net.params = {k: v.astype(np.float32) for k, v in pre[id(src)].items()}
                net.persistent = {k: v.astype(np.float32) for k, v in net.persistent.items()}
            trace32 = {}
            ref.Updater(m32, g32, i32, v32).update_core(x_real, t_real, r, trace=trace32)
            d_override = None
        olosses = oup.update_core(x_real.astype(np.float64), t_real, r, trace=trace, d_override=d_override)

        check_forward(up, trace, oG, tol_fwd if tol_fwd is not None else tol_out, report)
        kink = ref.kink_margin(trace)
        gerrs = {}
        report["grads"].append(gerrs)
        for name, oname in (("ImageDiscriminator", "image_dis/loss"), ("VideoDiscriminator", "video_dis/loss"),
                            ("ImageGenerator", "image_gen/loss")):
            got = float(up.losses[name])
            assert abs(got - olosses[oname]) <= tol_out * max(1.0, abs(olosses[oname])), (step, name, got, olosses[oname])
        for mine, theirs, key in ((Di, oI, "grads_di"), (Dv, oV, "grads_dv"), (G, oG, "grads_g")):
            bn_fed = ("dc1/b", "dc2/b", "dc3/b", "dc4/b") if mine is G else ("dc2/b", "dc3/b", "dc4/b")
            for path, p in mine.namedparams():
                k = path.lstrip("/")
                g_ref = trace[key][k]
                g = p.grad.float().cpu().numpy()
                if k in bn_fed:   # exactly-zero gradient by construction on the device; round-off in the oracle
                    assert np.abs(g).max() == 0.0 and np.abs(g_ref).max() < 1e-9
                    continue
                gerrs["%s/%s" % (mine.name, k)] = {
                    "relerr": relerr(g, g_ref),
                    "l2": float(np.linalg.norm(g - g_ref) / max(np.linalg.norm(g_ref), 1e-30)),
                    "cos": float((g * g_ref).sum() / (np.linalg.norm(g) * np.linalg.norm(g_ref) + 1e-30))}
                if dtype_mode == "fp32":
                    # float32 itself limits how well ANY fp32 implementation can follow the float64 truth through
                    # pass C (Adam's m/(sqrt(v)+eps) amplifies 1e-7 gradient differences into the updated D weights):
                    # the bound is tol_grad, or 10x the error of the oracle run in float32 ("as the reference runs"),
                    # capped at 2e-2.  Measured at n_filters=64: oracle-fp32 vs fp64 is 4e-4..7e-3 on G's gradients.
                    bound = tol_grad
                    if trace32 is not None:
                        bound = min(2e-2, max(tol_grad, 10 * relerr(trace32[key][k], g_ref)))
                    e = relerr(g, g_ref)
                    if e >= bound and kink < 2e-5:
                        # With ~1e7 (Leaky)ReLU inputs per step some pre-activation always sits within float32 round-off
                        # of the kink (kink_margin ~ 1e-7); if float32 lands on the other side of it than the float64
                        # truth, ONE mask bit flips and perturbs the gradients below it by ~1e-3.  Such a case must
                        # still agree in the L2 sense and stay within 2e-2 element-wise.
                        l2 = np.linalg.norm(g - g_ref) / max(np.linalg.norm(g_ref), 1e-30)
                        assert l2 < 2e-3 and e < 2e-2, (step, mine.name, k, e, l2, kink)
                    else:
                        assert e < bound, (step, mine.name, k, e, bound)
                else:
                    # bf16 storage perturbs activations by ~1%, which flips a few LeakyReLU/ReLU masks; whole-network
                    # gradients are therefore judged by direction and a loose magnitude bound, while the <= 2e-2
                    # per-layer bound is enforced on identical inputs in test_kernels_gpu.py.
                    if g.size <= 8:  # e.g. Di.dc5/b: a difference of two ~0.5/N terms; direction is meaningless
                        assert np.abs(g - g_ref).max() < 2e-2 or relerr(g, g_ref) < tol_grad, (step, mine.name, k)
                        continue
                    cos = float((g * g_ref).sum() / (np.linalg.norm(g) * np.linalg.norm(g_ref) + 1e-30))
                    l2 = float(np.linalg.norm(g - g_ref) / max(np.linalg.norm(g_ref), 1e-30))
                    bmax, bl2 = BF16_GRAD_BOUNDS[mine.name]
                    assert cos > BF16_GRAD_COS and relerr(g, g_ref) < min(tol_grad, bmax) and l2 < bl2, \
                        (step, mine.name, k, cos, relerr(g, g_ref), l2)
            # Adam + WeightDecay: replay the oracle's rule on the device's own gradients from the pre-step weights
            opt = up.get_optimizer({"ImageGenerator": "image_gen", "ImageDiscriminator": "image_dis",
                                    "VideoDiscriminator": "video_dis"}[mine.name])
            if step == 0:
                chk = {k: v.copy() for k, v in pre[id(theirs)].items()}
                st = ops.AdamState(chk)
                st.update(chk, {path.lstrip("/"): p.grad.float().cpu().numpy().astype(np.float64)
                                for path, p in mine.namedparams()})
                for path, p in mine.namedparams():
                    k = path.lstrip("/")
                    assert np.abs(p.data.float().cpu().numpy() - chk[k]).max() < 2e-6, (mine.name, k)
            assert opt.t == step + 1
        # BatchNorm running statistics (two updates per discriminator BN per step, one per generator BN)
        for mine, theirs in ((Di, oI), (Dv, oV), (G, oG)):
            for path, link, n in mine.namedpersistents():
                if n == "N":
                    continue
                v = getattr(link, n).cpu().numpy()
                assert relerr(v, theirs.persistent[path.lstrip("/")]) < max(tol_out, 1e-4), (mine.name, path)
        # keep the two sides in lock-step for multi-step runs: the oracle adopts the device's weights
        if steps > 1:
            for mine, theirs in ((Di, oI), (Dv, oV), (G, oG)):
                for path, p in mine.namedparams():
                    theirs.params[path.lstrip("/")][...] = p.data.float().cpu().numpy()


def test_step_fp32_strict_mnist_shape():
    """BASELINE config 1 shape family (C=1, no labels), narrow filters, strict fp32 path."""
    run_step_case("mnist_normal", 16, 3, "fp32", 1e-5, 1e-4)


def test_step_fp32_strict_infogan():
    run_step_case("mug_infogan", 8, 2, "fp32", 1e-5, 1e-4)


def test_step_fp32_strict_cgan():
    """--model cgan (train.py:74-79, updater.py:65-76,93-95,104-106): label planes on the real AND the attached fake
    clip, 9-channel discriminators, the generator receives the clip channels' slice of the video gradient."""
    run_step_case("mug_cgan", 8, 2, "fp32", 1e-5, 1e-4)


def test_step_bf16_tcgen05_cgan():
    run_step_case("mug_cgan", 64, 2, "bf16", 2e-2, 0.75)


def test_step_bf16_tcgen05_mug_normal():
    """Full-width (n_filters=64) so the tcgen05 kernels carry the convolutions, reduced batch."""
    run_step_case("mug_normal", 64, 4, "bf16", 2e-2, 0.75)


def test_step_bf16_two_steps_infogan():
    run_step_case("mug_infogan", 64, 2, "bf16", 2e-2, 0.75, steps=2)


def test_step_fp32_full_width():
    """n_filters = 64 (the width BASELINE quotes) on the strict path, reduced batch."""
    run_step_case("mug_normal", 64, 2, "fp32", 1e-5, 1e-4, ref32=True)


def test_step_fp32_pass_c_without_weight_handover():
    """Pass C judged with NO hand-over of the device's updated discriminator weights to the oracle: each side runs pass C
    on the weights its own passes A and B produced.  Adam with beta1 = 5e-5 moves every discriminator weight by
    ~alpha * g / (|g| + eps): fp32 round-off on near-zero gradients becomes ~1e-5 weight differences, which pass C's
    data gradients see — hence 5e-3 here instead of 1e-4; a device that ran pass C on STALE discriminator weights (or
    before their update finished) misses by ~alpha / |w| ~ 1e-2 .. 1e-1 on the generator's gradients and fails."""
    run_step_case("mug_normal", 8, 2, "fp32", 1e-5, 5e-3, override=False, tag="mug_normal_nf8_N2_fp32_nohandover")


def test_step_bf16_baseline_config2_batch35():
    """BASELINE config 2 exactly (normal model, n_filters 64, batch 35): the tile shapes pick_tile chooses at N = 35 /
    B = 560 (the benchmarked instantiations) inside a whole step, against the float64 oracle."""
    run_step_case("mug_normal", 64, 35, "bf16", 2e-2, 0.75)


def test_step_bf16_baseline_config4_infogan_batch35():
    """BASELINE config 4 (infogan: 7-way discriminator outputs, categorical terms of updater.py:28-37,53-56), batch 35."""
    run_step_case("mug_infogan", 64, 35, "bf16", 2e-2, 0.75)


def test_losses_track_oracle_over_100_steps():
    """north_star: "losses tracking over 100 steps".  Three runs go FREE for 100 consecutive update_core steps from the
    same initial weights on the same injected clips / latents / noise (no weight hand-over in between): the strict fp32
    device path, the oracle in float64 (truth) and the oracle in float32 (the reference's own precision).
    What can be asked: with beta1 = 5e-5 Adam is close to sign-SGD (dp ~ alpha*g/(|g|+eps)), so a float32 round-off
    difference on a near-zero gradient moves that weight by ~2e-4 per step and a GAN does not contract it; the float32
    ORACLE itself drifts from float64 to ~8e-2 in the losses by step 100 (tools/track_losses.py prints the curves).
    The device path must (i) agree to 1e-2 through the first 10 steps, (ii) never leave a 0.25 band (losses are
    O(1)), and (iii) drift on average no more than 5x the float32 oracle's own drift (+1e-2)."""
    from mocogan_chainer_b200 import kernels as K
    from mocogan_chainer_b200 import random as mrandom
    model, (G, Di, Dv), (oG, oI, oV), up, oup = build_pair("mug_normal", 8, "fp32")
    m32, g32, i32, v32 = ref.build_models("mug_normal", dtype=np.float64, seed=3, n_filters=8)
    for net, src in ((g32, oG), (i32, oI), (v32, oV)):
        net.dtype = np.float32
        net.params = {k: v.astype(np.float32) for k, v in src.params.items()}
        net.persistent = {k: v.astype(np.float32) for k, v in net.persistent.items()}
    o32 = ref.Updater(m32, g32, i32, v32)
    N, C = 2, oG.out_channels
    names = (("ImageDiscriminator", "image_dis/loss"), ("VideoDiscriminator", "video_dis/loss"),
             ("ImageGenerator", "image_gen/loss"))
    dev, o32dev = [], []
    for step in range(100):
        x_real = np.random.default_rng(1234 + step).uniform(-1, 1, size=(N, C, 16, 64, 64)).astype(np.float32)
        t_real = np.random.default_rng(5 + step).integers(0, 6, size=N)
        r = ref.draw_step_randoms(np.random.default_rng(100 + step), np.random.default_rng(200 + step), oG, oI, oV, N,
                                  x_real.shape, t=(7 + 3 * step) % 16, dtype=np.float32)
        mrandom.set_source(mrandom.InjectedRandom(r))
        up.step_on_device(torch.from_numpy(x_real).cuda(), torch.from_numpy(t_real).int().cuda())
        l64 = oup.update_core(x_real.astype(np.float64), t_real, r)
        l32 = o32.update_core(x_real, t_real, r)
        d = max(abs(float(up.losses[a]) - l64[b]) for a, b in names)
        dev.append(d)
        o32dev.append(max(abs(float(l32[b]) - l64[b]) for a, b in names))
        assert np.isfinite(d) and d < (1e-2 if step < 10 else 0.25), (step, d)
    assert K.tc_error_flag() == 0
    assert np.mean(dev) <= 5 * np.mean(o32dev) + 1e-2, (np.mean(dev), np.mean(o32dev))
