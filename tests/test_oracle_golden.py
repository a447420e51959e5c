"""The oracle against its committed golden vectors (CPU), and the CUDA path against the same vectors (GPU)."""
import os

import numpy as np
import pytest

from tests.golden.make_golden import CASES, make_inputs, run_case

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
            got = float(np.linalg.norm(p.grad.float().cpu().numpy()))
            assert abs(got - want) < 2e-3 * want, (tag, k, got, want)
        for path, link, n in net.namedpersistents():
            if n == "N":
                continue
            key = "running_%s_%s" % (tag, path.lstrip("/").replace("/", "_"))
            v = getattr(link, n).cpu().numpy()
            assert np.abs(v - gold[key]).max() < 1e-4 * max(np.abs(gold[key]).max(), 1e-6), key
