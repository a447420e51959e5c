"""The oracle against its committed golden vectors (CPU), and the CUDA path against the same vectors (GPU)."""
import os

import numpy as np
import pytest

from tests.golden.make_golden import CASES, grad_sample, make_inputs, run_case

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    gold = np.load(os.path.join(HERE, "step_%s.npz" % name))
    out = run_case(**CASES[name])
    assert set(gold.files) == set(out.keys())
    for k in gold.files:
        a, b = np.asarray(out[k], np.float64), np.asarray(gold[k], np.float64)
        assert a.shape == b.shape, k
        assert np.abs(a - b).max() <= 1e-6 * max(np.abs(b).max(), 1e-12) + 1e-12, k


# the cgan vectors pin the oracle on the CPU side; the device's cgan step is compared with the oracle directly in
# tests/test_step_gpu.py::test_step_{fp32_strict,bf16_tcgen05}_cgan
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(set(CASES) - {"mug_cgan"}))
def test_cuda_step_matches_golden(name):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from mocogan_chainer_b200 import random as mrandom
    from tests.test_step_gpu import build_pair
    kw = CASES[name]
    gold = np.load(os.path.join(HERE, "step_%s.npz" % name))
    model, (G, Di, Dv), _, up, _ = build_pair(kw["config"], kw["nf"], "fp32")
    _, oG, oI, oV, x_real, t_real, r = make_inputs(kw["config"], kw["nf"], kw["N"])
    for mine, theirs in ((G, oG), (Di, oI), (Dv, oV)):
        for path, p in mine.namedparams():
            p.data = theirs.params[path.lstrip("/")].astype(np.float32)
    mrandom.set_source(mrandom.InjectedRandom(r))
    up.step_on_device(torch.from_numpy(x_real).cuda(), None if t_real is None else torch.from_numpy(t_real).int().cuda())
    torch.cuda.synchronize()
    rel = lambda a, b: np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30)
    # the generator's clip (sub-sampled as stored) and the four discriminator outputs: fp32 path <= 1e-5
    C = oG.out_channels
    xf = up.last_forward["x_fake"].data.float().permute(2, 0, 1, 3, 4)[:, :, :C].cpu().numpy()
    assert rel(xf[::5, :, :, ::16, ::16], gold["x_fake_sample"]) < 1e-5
    assert abs(xf.mean() - gold["x_fake_mean_std"][0]) < 1e-5 and abs(xf.std() - gold["x_fake_mean_std"][1]) < 1e-5
    for key in ("y_real_i", "y_real_v", "y_fake_i", "y_fake_v"):
        assert rel(up.last_forward[key].data.float().cpu().numpy(), gold[key]) < 1e-5, key
    for mine_name, key in (("ImageDiscriminator", "loss_image_dis_loss"), ("VideoDiscriminator", "loss_video_dis_loss"),
                           ("ImageGenerator", "loss_image_gen_loss")):
        assert abs(float(up.losses[mine_name]) - float(gold[key])) < 1e-5 * max(1.0, abs(float(gold[key])))
    for tag, net in (("g", G), ("di", Di), ("dv", Dv)):
        for path, p in net.namedparams():
            k = path.lstrip("/")
            key = "gradnorm_%s_%s" % (tag, k.replace("/", "_"))
            want = float(gold[key])
            if want < 1e-9:      # BN-fed biases: exactly zero on the device
                continue
            g = p.grad.float().cpu().numpy()
            got = float(np.linalg.norm(g))
            assert abs(got - want) < 2e-3 * want, (tag, k, got, want)
            # element-wise (the whole tensor, or the stored stride through it): fp32 gradients <= 1e-4; pass C (the
            # generator) is compared without the weight hand-over of test_step_gpu.py, so Adam's amplification of
            # round-off in the updated discriminator weights is inside its bound
            gs = gold["grad_%s_%s" % (tag, k.replace("/", "_"))]
            assert rel(grad_sample(g), gs) < (5e-3 if tag == "g" else 1e-4), (tag, k, rel(grad_sample(g), gs))
        for path, link, n in net.namedpersistents():
            if n == "N":
                continue
            key = "running_%s_%s" % (tag, path.lstrip("/").replace("/", "_"))
            v = getattr(link, n).cpu().numpy()
            assert np.abs(v - gold[key]).max() < 1e-4 * max(np.abs(gold[key]).max(), 1e-6), key
