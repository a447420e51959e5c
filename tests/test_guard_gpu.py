"""Out-of-bounds writes: whole training steps with a red zone around every device buffer the package allocates
(tests/guard_alloc.py).  Every kernel of the step runs at the benchmark's own sizes (BASELINE configs 2 and 4: batch
35, n_filters 64, bf16 — the class-fused dgrad, the window-view 3-channel layers, stream-K wgrad, the last-block column
reductions) and at small odd sizes in both dtypes (the dispatch paths the big layers do not take), eagerly, and no red
zone may change.  This is the check `compute-sanitizer --tool memcheck` would do; that tool is not available on the
B200 pool."""
import numpy as np
import pytest
import torch

from tests.guard_alloc import GUARD, guarded_allocations

pytestmark = pytest.mark.gpu


def test_guard_catches_a_write_one_element_past_the_end():
    from mocogan_chainer_b200 import kernels as K
    with guarded_allocations() as g:
        t = torch.empty(1000, dtype=torch.float32, device="cuda")
        assert t.data_ptr() % 128 == 0
        K.fill_zero(t)
        assert g.check() == 1
        t = torch.empty(1000, dtype=torch.float32, device="cuda")
        flat, nbytes, _ = g.buffers[-1]
        K.fill_zero(flat[GUARD:GUARD + nbytes + 16].view(torch.float32))      # handed 16 bytes too many
        with pytest.raises(AssertionError, match="16 bytes above"):
            g.check()


def _losses_ok(up):
    losses = {k: float(v) for k, v in up.losses.items()}
    assert len(losses) == 3 and all(np.isfinite(v) and abs(v) < 1e3 for v in losses.values()), losses


@pytest.mark.parametrize("model,poison", [("normal", False), ("infogan", False), ("normal", True), ("infogan", True)])
def test_no_write_outside_any_buffer_baseline_batch35(model, poison):
    """poison=True: every `torch.empty` payload starts as NaN bytes, so the step must not consume anything it (or a
    zero-fill) did not write first — losses and all weights stay finite."""
    import bench
    from mocogan_chainer_b200 import kernels as K
    K._ws_cache.clear()                      # scratch of earlier tests was allocated without red zones
    with guarded_allocations(poison=poison) as g:
        up, it = bench.build_updater(35, 1234, use_graph=False, model=model)
        up.step_host_inputs(it.x[0].cuda(), it.t[0].cuda())       # device-resident batch (bench `value`)
        up.update_core()                                          # iterator -> pinned host batch -> H2D -> step (bench `e2e`)
        torch.cuda.synchronize()
        assert K.tc_error_flag() == 0
        n = g.check()
    assert n > 200, n                        # activations, gradients, arenas, Adam state, scratch: all were guarded
    _losses_ok(up)
    for o in up.get_all_optimizers().values():
        assert bool(torch.isfinite(o.target.arena().data).all()), o.target.name
    K._ws_cache.clear()


@pytest.mark.parametrize("poison", [False, True])
@pytest.mark.parametrize("model,dtype_mode,nf,N", [("cgan", "fp32", 16, 3), ("cgan", "bf16", 64, 3), ("infogan", "bf16", 64, 5),
                                                   ("normal", "fp32", 8, 2)])
def test_no_write_outside_any_buffer_small_odd_sizes(model, dtype_mode, nf, N, poison):
    from mocogan_chainer_b200 import chainer, train
    from mocogan_chainer_b200 import kernels as K
    from mocogan_chainer_b200 import random as mrandom
    from mocogan_chainer_b200.model.updater import Updater
    chainer.config.compute_dtype = dtype_mode
    K._ws_cache.clear()
    np.random.seed(0)

    class _It(object):
        epoch, is_new_epoch, epoch_detail = 0, False, 0.0

    with guarded_allocations(poison=poison) as g:
        G, Di, Dv = train.build_models(model, 50, 10, 6, 3, nf, 16, True, 0.2)
        opts = {k: train.make_optimizer(m, 2e-4, 5e-5) for k, m in (("image_gen", G), ("image_dis", Di), ("video_dis", Dv))}
        mrandom.set_source(mrandom.DeviceRandom(seed=7, device="cuda", video_length=16))
        up = Updater(model=model, models=(G, Di, Dv), video_length=16, img_size=64, channel=3, dim_zl=6,
                     tensorboard_writer=None, iterator=_It(), optimizer=opts, device=0)
        x = torch.rand((N, 3, 16, 64, 64), device="cuda") * 2 - 1
        t = torch.randint(0, 6, (N,), device="cuda", dtype=torch.int32)
        for _ in range(2):
            up.step_on_device(x, t)
        torch.cuda.synchronize()
        assert K.tc_error_flag() == 0
        assert g.check() > 100
    _losses_ok(up)
    for m in (G, Di, Dv):
        assert bool(torch.isfinite(m.arena().data).all()), m.name
    K._ws_cache.clear()
