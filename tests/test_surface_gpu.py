"""train.py / generate_samples.py drop-in surface on the GPU: a short synthetic training run (eager and CUDA-graph),
npz snapshots with the reference's key schema, and generate_samples loading them."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _needs_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")


def test_train_then_generate(tmp_path, monkeypatch):
    from mocogan_chainer_b200 import generate_samples, train
    monkeypatch.chdir(tmp_path)
    up = train.main(["-g", "0", "--synthetic", "12", "--batchsize", "4", "--max_epoch", "2", "--snapshot_interval", "1",
                     "--n_filters_gen", "64", "--save_name", "run", "--model", "infogan", "--graph"])
    assert up.iteration == 6 and up.epoch == 2 and up._graph is not None
    for v in up.losses.values():
        assert np.isfinite(float(v))
    out = tmp_path / "result" / "run"
    for f in ("image_gen_epoch_1.npz", "video_dis_epoch_2.npz", "image_dis_epoch_fianl.npz"):
        assert (out / f).exists()
    with np.load(out / "image_gen_epoch_fianl.npz") as f:
        assert f["dc1/W"].shape == (60, 512, 4, 4) and f["g0/W_r/W"].shape == (10, 16)
        assert f["bn1/avg_mean"].shape == (512,) and "bn4/N" in f.files
        assert np.abs(f["bn1/avg_mean"]).max() > 0      # running statistics were updated by training
    videos = generate_samples.main([str(out / "image_gen_epoch_fianl.npz"), str(tmp_path / "gen"), "-n", "4"])
    assert videos.shape == (16, 4, 3, 64, 64) and videos.dtype == np.uint8
    assert os.path.exists(tmp_path / "gen" / "videos.npy")
    from oracle import mocogan_ref as ref
    grid = np.load(tmp_path / "gen" / "grid.npy")
    assert grid.shape == (16, 3, 128, 128) and np.array_equal(grid, ref.to_grid(videos, 2))   # util.py:30-51


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_video_to_uint8_and_grid_match_reference_postprocessing(dtype):
    """generate_samples.py:39 `((v / 2. + 0.5) * 255).astype(np.uint8)` + util.py:30-51 `to_grid`, bit-exact (5 videos on
    a 3x3 grid: four cells stay black; values hit both ends of the tanh range)."""
    from mocogan_chainer_b200 import kernels as K
    from oracle import mocogan_ref as ref
    T, N, C, H, W = 3, 5, 3, 8, 8
    v = torch.tanh(torch.randn((T, N, C, H, W), generator=torch.Generator().manual_seed(0)) * 2).to(dtype)
    v.view(-1)[:4] = torch.tensor([-1.0, 1.0, 0.0, 0.999], dtype=dtype)
    phys = v.cuda().permute(0, 1, 3, 4, 2).contiguous().reshape(T * N, 1, H, W, C)
    u8, grid = K.video_to_uint8(phys, T, N, True, 3)
    want = ref.to_uint8(v.float().numpy())
    assert np.array_equal(u8.cpu().numpy(), want)
    assert np.array_equal(grid.cpu().numpy(), ref.to_grid(want, 3))


def test_graph_replay_equals_eager_with_same_seed():
    """The captured step must be the same computation as the eager step: identical weights after 4 steps."""
    from mocogan_chainer_b200 import chainer
    from mocogan_chainer_b200 import random as mrandom
    from tests.test_step_gpu import build_pair
    res = []
    for graph in (False, True):
        np.random.seed(0)
        chainer.config.compute_dtype = "fp32"
        _, (G, Di, Dv), _, up, _ = build_pair("mug_normal", 8, "fp32")
        up.use_graph, up.graph_warmup = graph, 1
        mrandom.set_source(mrandom.DeviceRandom(seed=7, video_length=16))
        x = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, (2, 3, 16, 64, 64)).astype(np.float32)).cuda()
        t = torch.tensor([1, 4], dtype=torch.int32, device="cuda")
        for _ in range(4):
            up.step_host_inputs(x, t)
        torch.cuda.synchronize()
        assert (up._graph is not None) == graph
        res.append(torch.cat([m.arena().data.clone() for m in (G, Di, Dv)]))
    # wgrad uses fp32 atomics (order-dependent rounding), so compare to fp32 round-off amplified by 4 Adam steps
    assert (res[0] - res[1]).abs().max().item() < 5e-4
    assert (res[0] - res[1]).abs().mean().item() < 1e-5


def test_trainer_snapshot_keys_and_resume(tmp_path, monkeypatch):
    """§8f rank 2: `snapshot_epoch_N.npz` (extensions.snapshot, train.py:137-138) carries Chainer's trainer key schema
    under `updater/` and `--resume` (train.py:162-163) restores models, Adam moments / step counts, the iterator and the
    iteration count, so a resumed run continues where the first one stopped."""
    from mocogan_chainer_b200 import train
    monkeypatch.chdir(tmp_path)
    common = ["-g", "0", "--synthetic", "8", "--batchsize", "4", "--snapshot_interval", "1", "--n_filters_gen", "8",
              "--dtype", "fp32", "--seed", "3"]
    up1 = train.main(common + ["--max_epoch", "2", "--save_name", "a"])
    snap = tmp_path / "result" / "a" / "snapshot_epoch_1.npz"
    with np.load(snap) as f:
        keys = set(f.files)
        for k in ("updater/iteration", "updater/iterator:main/current_position", "updater/iterator:main/epoch",
                  "updater/iterator:main/is_new_epoch", "updater/iterator:main/order",
                  "updater/model:image_gen/dc1/W", "updater/model:image_gen/g0/W_r/W", "updater/model:video_dis/bn2/avg_mean",
                  "updater/optimizer:image_gen/t", "updater/optimizer:image_gen/epoch",
                  "updater/optimizer:video_dis/dc4/W/m", "updater/optimizer:video_dis/dc4/W/v",
                  "updater/optimizer:video_dis/dc4/W/t", "updater/optimizer:image_dis/bn3/gamma/m"):
            assert k in keys, k
        assert int(f["updater/iteration"]) == 2 and int(f["updater/optimizer:image_dis/t"]) == 2
        assert f["updater/optimizer:video_dis/dc4/W/m"].shape == f["updater/model:video_dis/dc4/W"].shape == (64, 32, 4, 4, 4)
        assert np.abs(f["updater/optimizer:video_dis/dc4/W/v"]).max() > 0
        saved = {k: f[k].copy() for k in f.files}
    # resume from epoch 1 and train to epoch 2: same iteration / step counts as the uninterrupted run
    up2 = train.main(common + ["--max_epoch", "2", "--save_name", "b", "--resume", str(snap)])
    assert up2.iteration == up1.iteration == 4 and up2.epoch == 2
    for name in ("image_gen", "image_dis", "video_dis"):
        assert up2.get_optimizer(name).t == 4 and int(up2.get_optimizer(name).t_dev.item()) == 4
    # and loading alone reproduces the saved state bit for bit
    from mocogan_chainer_b200 import chainer
    chainer.serializers.load_npz(snap, chainer.training.TrainerState(up2))
    again = chainer.serializers._collect(chainer.training.TrainerState(up2))
    for k, v in saved.items():
        assert np.array_equal(np.asarray(again[k]), v), k


def test_uint8_batch_equals_float_batch():
    """A uint8 channels-last batch (the clip cache's) through the step == the same clips normalised on the host
    (datasets.py:91) through the step: the (v-128)/128 fused into mcg_pack_video is exact."""
    from mocogan_chainer_b200 import chainer
    from mocogan_chainer_b200 import random as mrandom
    from tests.test_step_gpu import build_pair
    rng = np.random.default_rng(2)
    u8 = rng.integers(0, 256, size=(2, 16, 64, 64, 3), dtype=np.uint8)               # (N, T, H, W, C)
    xf = ((u8.astype(np.float32) - 128.) / 128.).transpose(0, 4, 1, 2, 3).copy()    # (N, C, T, H, W) float32
    t = torch.tensor([1, 4], dtype=torch.int32, device="cuda")
    losses = []
    for x in (torch.from_numpy(u8).cuda().permute(0, 4, 1, 2, 3), torch.from_numpy(xf).cuda()):
        np.random.seed(0)
        chainer.config.compute_dtype = "fp32"
        _, _, _, up, _ = build_pair("mug_normal", 8, "fp32")
        mrandom.set_source(mrandom.DeviceRandom(seed=7, video_length=16))
        up.step_on_device(x, t)
        torch.cuda.synchronize()
        losses.append({k: float(v) for k, v in up.losses.items()})
    for k in losses[0]:
        assert abs(losses[0][k] - losses[1][k]) < 1e-6, (k, losses)


@pytest.mark.parametrize("dtype_mode,nf,tol", [("fp32", 16, 1e-5), ("bf16", 64, 2e-2)])
def test_generator_eval_mode_uses_running_statistics(dtype_mode, nf, tol):
    """§8f rank 1: under chainer.using_config('train', False) — the state util.py:92 `log_tensorboard` runs the
    generator in — BatchNorm normalises with the RUNNING statistics (F.fixed_batch_normalization) and
    updates nothing.  Non-trivial statistics are loaded into both sides; the clip is compared with the oracle's
    `batchnorm_fixed` forward."""
    from mocogan_chainer_b200 import chainer
    from mocogan_chainer_b200 import random as mrandom
    from oracle import mocogan_ref as ref
    from tests.test_step_gpu import build_pair, relerr
    _, (G, Di, Dv), (oG, oI, oV), up, _ = build_pair("mug_normal", nf, dtype_mode)
    rng = np.random.default_rng(21)
    G.arena()
    before = {}
    for path, link, n in G.namedpersistents():
        if n == "N":
            continue
        key = path.lstrip("/")
        v = (0.3 * rng.standard_normal(oG.persistent[key].shape)) if n == "avg_mean" else rng.uniform(0.5, 2.0, oG.persistent[key].shape)
        v = v.astype(np.float32)
        oG.persistent[key] = v.astype(np.float64)
        getattr(link, n).copy_(torch.from_numpy(v))
        before[key] = v
    N = 3
    r = ref.draw_step_randoms(np.random.default_rng(100), np.random.default_rng(200), oG, oI, oV, N, (N, 3, 16, 64, 64), t=3,
                              dtype=np.float32)
    want, _ = oG.forward(N, r["latents"], update_running=False, train=False)
    mrandom.set_source(mrandom.InjectedRandom(r))
    with chainer.using_config('train', False), chainer.no_backprop_mode():
        x, labels = G(N)
    torch.cuda.synchronize()
    assert tuple(x.shape) == (16, N, 3, 64, 64)
    assert relerr(x.data.float().cpu().numpy(), want) < tol
    for path, link, n in G.namedpersistents():      # eval mode leaves the running statistics alone
        if n != "N":
            assert np.array_equal(getattr(link, n).cpu().numpy(), before[path.lstrip("/")])
    # and it is a different function from the training-mode forward (batch statistics)
    mrandom.set_source(mrandom.InjectedRandom(r))
    with chainer.no_backprop_mode():
        x_train, _ = G(N)
    assert relerr(x_train.data.float().cpu().numpy(), want) > 10 * tol


def test_log_tensorboard_hook_writes_eval_mode_grid_frames(tmp_path):
    """util.py:89-115: four grid frames ('{:02d}th frame' at np.linspace(0, T, 4, endpoint=False)) and the first clips as
    frame strips, generated with train = False; the grid equals util.py:30-51 `to_grid` of the uint8 clips."""
    from mocogan_chainer_b200 import chainer, util
    from mocogan_chainer_b200 import random as mrandom
    from mocogan_chainer_b200.model.net import ImageGenerator
    from oracle import mocogan_ref as ref
    chainer.config.compute_dtype = "bf16"
    np.random.seed(0)
    G = ImageGenerator(50, 10, 6, 3, 64, 16)
    G.arena()
    G.bn1.avg_var.fill_(1.0), G.bn2.avg_var.fill_(1.0), G.bn3.avg_var.fill_(1.0), G.bn4.avg_var.fill_(1.0)

    class Rec(object):
        def __init__(self):
            self.images = {}

        def add_image(self, tag, img, step):
            self.images[tag] = (np.asarray(img), step)

    class Up(object):
        epoch = 7

    rec = Rec()
    mrandom.set_source(mrandom.DeviceRandom(seed=5, video_length=16))
    train_before = chainer.config.train
    util.log_tensorboard(G, 16, 16, rec)(Up())
    assert chainer.config.train == train_before
    assert sorted(k for k in rec.images if k.endswith("frame")) == ["00th frame", "04th frame", "08th frame", "12th frame"]
    assert sorted(k for k in rec.images if k.startswith("video_")) == ["video_%02d" % i for i in range(10)]
    img, step = rec.images["04th frame"]
    assert img.shape == (3, 256, 256) and img.dtype == np.uint8 and step == 7
    assert rec.images["video_03"][0].shape == (3, 64, 16 * 64)
    # the same latents again (same Philox state and call ids): the written grid frame is to_grid of the uint8 clips
    mrandom.set_source(mrandom.DeviceRandom(seed=5, video_length=16))
    videos, grid = util.sample_videos(G, 16, train=False)
    assert np.array_equal(grid.cpu().numpy(), ref.to_grid(videos.cpu().numpy(), 4))
    assert np.array_equal(grid[4].cpu().numpy(), img)
    assert np.array_equal(ref.to_grid(videos.cpu().numpy(), 4)[:, :, :64, 3 * 64:4 * 64], videos[:, 3].cpu().numpy())
    w = util.ImageLogWriter(tmp_path / "runs")
    w.add_image("00th frame", img, 1)
    w.add_scalar("loss:ImageGenerator", 0.5, 1)
    assert (tmp_path / "runs" / "00th_frame_000001.png").exists() and (tmp_path / "runs" / "scalars.jsonl").exists()


@pytest.mark.parametrize("dtype_mode,nf,tol", [("fp32", 8, 1e-5), ("bf16", 64, 3e-2)])
def test_generator128_extension_matches_composed_oracle_ops(dtype_mode, nf, tol):
    """model/net128.py (EXTENSION, parity unpinned by the reference — it cannot emit 128x128): the six-stage generator
    against the oracle's GRU / deconvolution / BatchNorm restatements composed the same way.  bf16: six chained bf16
    layers compared on tanh's absolute scale, like x_fake in test_step_gpu.py (FORWARD_SLACK)."""
    from mocogan_chainer_b200 import chainer
    from mocogan_chainer_b200 import random as mrandom
    from mocogan_chainer_b200.model.net128 import ImageGenerator128
    from oracle import chainer_ops as ops
    from oracle import mocogan_ref as ref
    chainer.config.compute_dtype = dtype_mode
    np.random.seed(4)
    T, N = 4, 3
    G = ImageGenerator128(50, 10, 6, 3, nf, T)
    G.arena()
    assert sorted(G._children) == ["bn1", "bn2", "bn3", "bn4", "bn5", "dc1", "dc2", "dc3", "dc4", "dc5", "dc6", "g0"]
    P = {path.lstrip("/"): p.data.float().cpu().numpy().astype(np.float64) for path, p in G.namedparams()}
    assert P["dc1/W"].shape == (60, nf * 16, 4, 4) and P["dc6/W"].shape == (nf, 3, 4, 4)
    lat = ref.ImageGenerator.draw_latents(np.random.default_rng(9), N, 50, 10, 6, T, np.float32)
    # the oracle's motion path (net.py:61-107 restated), then six deconvolution stages
    g0 = {k[3:]: v for k, v in P.items() if k.startswith("g0/")}
    zl = np.eye(6)[lat["labels"]]
    h, hs = lat["h0"].astype(np.float64), []
    for t in range(T):
        h, _ = ops.gru_step_fwd(g0, h, np.concatenate((zl, lat["eps"][t]), axis=1))
        hs.append(h)
    z = np.concatenate((np.tile(lat["zc"][None], (T, 1, 1)), np.stack(hs)), axis=2).reshape(T * N, 60, 1, 1)
    x = z
    for i in range(1, 7):
        s, p = ((1, 1), (0, 0)) if i == 1 else ((2, 2), (1, 1))
        y = ops.deconv_nd_fwd(x, P["dc%d/W" % i], P["dc%d/b" % i], s, p)
        x = np.maximum(ops.batchnorm_fwd(y, P["bn%d/gamma" % i], P["bn%d/beta" % i])[0], 0) if i < 6 else np.tanh(y)
    want = x.reshape(T, N, 3, 128, 128)
    r = {"t": 0, "noise_i_real": [], "noise_v_real": [], "noise_i_fake": [], "noise_v_fake": [], "latents": lat}
    mrandom.set_source(mrandom.InjectedRandom(r))
    with chainer.no_backprop_mode():
        got, labels = G(N)
    torch.cuda.synchronize()
    assert tuple(got.shape) == (T, N, 3, 128, 128)
    err = np.abs(got.data.float().cpu().numpy() - want).max() / np.abs(want).max()
    assert err < tol, err
    assert abs(G.forward_gflop_per_frame() - 2 * 16 * (60 * 16 * nf + 16 * nf * 8 * nf * 16 + 8 * nf * 4 * nf * 64 + 4 * nf * 2 * nf * 256
                                                         + 2 * nf * nf * 1024 + nf * 3 * 4096) / 1e9) < 1e-9


def test_backward_sums_dense_gradients_across_streams():
    """A variable with two dense consumers whose nodes ran on different streams (chainer.config.branch_streams): the sum
    of the two gradients waits for both producers and is itself ordered for whoever reads it next."""
    from mocogan_chainer_b200 import chainer
    from mocogan_chainer_b200.chainer import FunctionNode, Variable

    class SlowScale(FunctionNode):
        def __init__(self, k):
            super(SlowScale, self).__init__()
            self.k = k

        def forward(self, inputs):
            return inputs[0] * self.k,

        def backward(self, idx, gys):
            g = gys[0]
            for _ in range(200):              # keep this stream busy so an unordered sum would read a stale buffer
                g = g * 1.0
            return g * self.k,

    class Sum2(FunctionNode):
        def forward(self, inputs):
            return (inputs[0] + inputs[1]).sum(),

        def backward(self, idx, gys):
            a, b = self.inputs
            return tuple(torch.ones_like(v.data) for v in (a, b))[:len(idx)]

    old = chainer.config.branch_streams
    chainer.config.branch_streams = True
    try:
        x = Variable(torch.randn(1 << 20, device="cuda"), requires_grad=True)
        h = SlowScale(1.0).apply((x,))[0]                      # x -> h on the caller's stream
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            a = SlowScale(2.0).apply((h,))[0]                  # consumer 1 of h, on `side`
        b = SlowScale(3.0).apply((h,))[0]                      # consumer 2 of h, on the caller's stream
        torch.cuda.current_stream().wait_stream(side)
        loss = Sum2().apply((a, b))[0]
        loss.backward()
        torch.cuda.synchronize()
        assert torch.equal(x.grad, torch.full_like(x.data, 5.0))
    finally:
        chainer.config.branch_streams = old
