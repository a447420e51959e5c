"""The oracle's hand-written backward passes and step semantics vs independent torch float64 autograd."""
import numpy as np
import pytest

from oracle import mocogan_ref as ref
from tests.torch_ref import torch_step


BN_FED_BIAS = {"image_gen": ["dc1/b", "dc2/b", "dc3/b", "dc4/b"], "image_dis": ["dc2/b", "dc3/b", "dc4/b"],
               "video_dis": ["dc2/b", "dc3/b", "dc4/b"]}


def _relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("config", ["mnist_normal", "mug_normal", "mug_infogan", "mug_cgan"])
def test_step_matches_torch_autograd(config):
    nf, N = 4, 3
    model, G, Di, Dv = ref.build_models(config, dtype=np.float64, seed=3, n_filters=nf)
    # make the test bite: non-trivial biases / gamma / beta
    prng = np.random.default_rng(11)
    for net in (G, Di, Dv):
        for k, v in net.params.items():
            if k.endswith("/b") or k.endswith("beta"):
                v += 0.1 * prng.standard_normal(v.shape)
            if k.endswith("gamma"):
                v += 0.2 * prng.standard_normal(v.shape)
    C = G.out_channels
    x_real = np.random.default_rng(1234).uniform(-1, 1, size=(N, C, 16, 64, 64))
    t_real = np.random.default_rng(5).integers(0, 6, size=N)
    r = ref.draw_step_randoms(np.random.default_rng(1), np.random.default_rng(2), G, Di, Dv, N, x_real.shape,
                              t=7, dtype=np.float64)
    t_losses, t_grads, t_post, t_xfake = torch_step(model, G, Di, Dv, x_real, t_real, r)
    up = ref.Updater(model, G, Di, Dv)
    trace = {}
    losses = up.update_core(x_real, t_real, r, trace=trace)
    assert _relerr(trace["x_fake"], t_xfake) < 1e-10
    for k in losses:
        assert abs(losses[k] - t_losses[k]) < 1e-10 * max(1, abs(t_losses[k])), k
    for name, key in (("image_dis", "grads_di"), ("video_dis", "grads_dv"), ("image_gen", "grads_g")):
        for p, g in trace[key].items():
            tg = t_grads[name][p]
            assert tg is not None, (name, p)
            if p in BN_FED_BIAS[name]:  # true gradient is exactly 0; both sides hold only round-off
                assert np.abs(g).max() < 1e-12 and np.abs(tg).max() < 1e-12
                continue
            scale = max(np.abs(tg).max(), 1e-12)
            assert np.abs(g - tg).max() / scale < 1e-7, (name, p, np.abs(g - tg).max(), scale)
    # post-step weights: biases feeding a BN have true gradient 0 -> Adam amplifies round-off; skip them
    skip = BN_FED_BIAS
    for name, net in (("image_gen", G), ("image_dis", Di), ("video_dis", Dv)):
        for p, v in net.params.items():
            if p in skip[name]:
                continue
            assert np.abs(v - t_post[name][p]).max() < 1e-7, (name, p)


def test_as_executed_gives_same_update():
    """The dead back-propagation the reference performs must not change any observable."""
    outs = []
    for as_exec in (False, True):
        model, G, Di, Dv = ref.build_models("mug_normal", dtype=np.float64, seed=3, n_filters=4)
        x_real = np.random.default_rng(1234).uniform(-1, 1, size=(2, 3, 16, 64, 64))
        r = ref.draw_step_randoms(np.random.default_rng(1), np.random.default_rng(2), G, Di, Dv, 2, x_real.shape,
                                  t=3, dtype=np.float64)
        up = ref.Updater(model, G, Di, Dv)
        losses = up.update_core(x_real, np.zeros(2, np.int64), r, as_executed=as_exec)
        outs.append((losses, {k: v.copy() for k, v in G.params.items()}, {k: v.copy() for k, v in Dv.params.items()}))
    assert outs[0][0] == outs[1][0]
    for k in outs[0][1]:
        assert np.array_equal(outs[0][1][k], outs[1][1][k]), k
    for k in outs[0][2]:
        assert np.array_equal(outs[0][2][k], outs[1][2][k]), k
