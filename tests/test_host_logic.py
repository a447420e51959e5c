"""Host-side logic that needs no GPU: the autograd sweep, dead-branch pruning, link/parameter naming (the npz key
schema of SURVEY.md App. D), iterators, flags."""
import numpy as np
import pytest
import torch

from mocogan_chainer_b200 import chainer
from oracle import mocogan_ref as ref
from mocogan_chainer_b200.chainer import FunctionNode, Variable


class Mul(FunctionNode):
    calls = []

    def __init__(self, tag):
        super(Mul, self).__init__()
        self.tag = tag

    def forward(self, inputs):
        return inputs[0] * inputs[1],

    def backward(self, idx, gys):
        Mul.calls.append((self.tag, idx))
        a, b = self.inputs
        out = {0: gys[0] * b.data, 1: gys[0] * a.data}
        return tuple(out[i] for i in idx)


def test_backward_accumulates_and_orders_by_rank():
    Mul.calls = []
    x = Variable(torch.tensor([2.0]))
    w = Variable(torch.tensor([3.0]))
    h = Mul("h").apply((x, w))[0]          # 6
    y = Mul("y").apply((h, h))[0]          # 36 = (xw)^2
    y.backward()
    assert float(x.grad) == 2 * 6 * 3 and float(w.grad) == 2 * 6 * 2
    assert [c[0] for c in Mul.calls] == ["y", "h"]


def test_dead_branches_are_never_launched():
    """SURVEY §3.2 pts 2,4: a stopped variable prunes every backward kernel that only feeds it."""
    Mul.calls = []
    x = Variable(torch.tensor([2.0]))
    frozen = Variable(torch.tensor([5.0]))
    w = Variable(torch.tensor([3.0]))
    h = Mul("h").apply((x, frozen))[0]
    y = Mul("y").apply((h, w))[0]
    x.stop = frozen.stop = True
    y.backward()
    assert Mul.calls == [("y", (1,))]       # no dgrad towards h, node "h" never runs
    assert x.grad is None and float(w.grad) == 10.0
    x.stop = False
    Mul.calls = []
    w.cleargrad()
    y.backward()
    assert Mul.calls == [("y", (0, 1)), ("h", (0,))]


def test_requires_grad_false_inputs_get_no_gradient():
    Mul.calls = []
    x = Variable(torch.tensor([2.0]), requires_grad=False)
    w = Variable(torch.tensor([3.0]))
    y = Mul("y").apply((x, w))[0]
    y.backward()
    assert Mul.calls == [("y", (1,))]


def test_param_names_match_chainer_npz_schema():
    from mocogan_chainer_b200.model.net import ImageDiscriminator, ImageGenerator, VideoDiscriminator
    np.random.seed(0)
    G = ImageGenerator(50, 10, 6, 3, 8, 16)
    names = [n.lstrip("/") for n, _ in G.namedparams()]
    for k in ("g0/W_r/W", "g0/U_z/b", "g0/W/W", "g0/U/b", "dc1/W", "dc5/b", "bn1/gamma", "bn4/beta"):
        assert k in names
    shapes = {n.lstrip("/"): p.shape for n, p in G.namedparams()}
    assert shapes["g0/W_r/W"] == (10, 16) and shapes["g0/U_r/W"] == (10, 10)
    assert shapes["dc1/W"] == (60, 64, 4, 4) and shapes["dc5/W"] == (8, 3, 4, 4)      # deconv: (in, out, kh, kw)
    Dv = VideoDiscriminator(3, 7, 8, True, 0.2)
    shapes = {n.lstrip("/"): p.shape for n, p in Dv.namedparams()}
    assert shapes["dc1/W"] == (8, 3, 4, 4, 4) and shapes["dc5/W"] == (7, 64, 4, 4, 4)
    pers = [p.lstrip("/") for p, _, _ in Dv.namedpersistents()]
    assert "bn2/avg_mean" in pers and "bn4/avg_var" in pers and "bn3/N" in pers
    assert ImageDiscriminator().count_params() == 2766529 and ImageGenerator(50, 10, 6).count_params() == 3250827
    assert VideoDiscriminator().count_params() == 11057857
    assert G.name == "ImageGenerator" and Dv.name == "VideoDiscriminator" and G.dc1.name == "dc1"


def test_internal_weight_layout_is_channels_last_of_chainer_layout():
    from mocogan_chainer_b200.chainer import Parameter
    a = np.arange(2 * 3 * 4 * 5, dtype=np.float32).reshape(2, 3, 4, 5)
    p = Parameter(a, channels_last_weight=True)
    assert p.internal_shape == (2, 4, 5, 3)
    assert np.array_equal(p._internal_init(), np.moveaxis(a, 1, -1))
    t = torch.from_numpy(p._internal_init().copy())
    assert np.array_equal(p._to_logical(t).numpy(), a)


def test_serial_iterator_epochs_and_concat():
    from mocogan_chainer_b200.chainer.dataset import concat_examples
    from mocogan_chainer_b200.chainer.iterators import SerialIterator
    data = [(np.full((1, 2), i, np.float32), i % 3) for i in range(10)]
    it = SerialIterator(data, 4, shuffle=False)
    b1 = it.next()
    assert [int(t) for _, t in b1] == [0, 1, 2] + [0] and not it.is_new_epoch
    it.next()
    b3 = it.next()
    assert it.is_new_epoch and it.epoch == 1 and len(b3) == 4
    x, t = concat_examples(b1)
    assert x.shape == (4, 1, 2) and list(t) == [0, 1, 2, 0]
    x, t = concat_examples([(np.zeros(3, np.float32), None)] * 2)   # Moving-MNIST labels are None (datasets.py:166)
    assert t is None and x.shape == (2, 3)


def test_glorot_and_lecun_statistics():
    from mocogan_chainer_b200.chainer import initializers as I
    np.random.seed(1)
    w = I.GlorotNormal()((256, 128, 4, 4))
    assert abs(w.std() - np.sqrt(2.0 / ((128 + 256) * 16))) < 2e-4
    w = I.LeCunNormal()((10, 16))
    assert w.shape == (10, 16) and w.dtype == np.float32


def test_adam_lr_schedule_matches_chainer_formula():
    import math
    opt = chainer.optimizers.Adam(alpha=2e-4, beta1=5e-5)
    opt.t = 3
    assert abs(opt.lr - 2e-4 * math.sqrt(1 - 0.999 ** 3) / (1 - (5e-5) ** 3)) < 1e-18


def test_train_flags_match_reference_defaults():
    from mocogan_chainer_b200.train import build_parser
    a = build_parser().parse_args([])
    assert (a.gpu, a.dataset_type, a.dataset, a.batchsize, a.max_epoch, a.model) == (-1, 'mug', 'data/dataset/train', 100,
                                                                                    1000, 'normal')
    assert (a.display_interval, a.snapshot_interval, a.log_tensorboard_interval, a.num_gen_samples) == (1, 10, 10, 36)
    assert (a.dim_zc, a.dim_zm, a.n_filters_gen, a.n_filters_idis, a.n_filters_vdis, a.resume) == (50, 10, 64, 64, 64, '')
    a = build_parser().parse_args(['-g', '0', '-r', 'x.npz', '--model', 'infogan', '--dataset_type', 'mnist'])
    assert a.gpu == 0 and a.resume == 'x.npz' and a.model == 'infogan'
    from mocogan_chainer_b200.generate_samples import build_parser as gp
    g = gp().parse_args(['w.npz', 'out'])
    assert (g.model_weight, g.save_path, g.num, g.gpu) == ('w.npz', 'out', 36, -1)


def test_model_wiring_follows_train_py():
    import pytest
    from mocogan_chainer_b200.train import build_models
    np.random.seed(0)
    g, di, dv = build_models("infogan", 50, 10, 6, 3, 8, 16, True, 0.2)
    assert di.out_channels == 7 and dv.out_channels == 7 and g.dim_zl == 6 and di.use_noise and dv.noise_sigma == 0.2
    g, di, dv = build_models("cgan", 50, 10, 6, 3, 8, 16, True, 0.2)
    assert di.in_channels == 9 and dv.in_channels == 9
    with pytest.raises(ValueError):
        build_models("cgan", 50, 10, 0, 3, 8, 16, True, 0.2)


def test_oracle_to_grid_restates_util_py():
    """oracle.to_uint8 / to_grid against a literal transcription of the reference's arithmetic on a small case
    (util.py:30-51: blank videos appended, cell (i, j) = video i*size + j; generate_samples.py:39)."""
    import numpy as np
    from oracle import mocogan_ref as ref
    rng = np.random.default_rng(0)
    v = np.tanh(rng.standard_normal((2, 3, 1, 2, 2)))
    u = ref.to_uint8(v)
    assert u.dtype == np.uint8 and u.min() >= 0 and np.array_equal(u, ((v / 2. + 0.5) * 255).astype(np.uint8))
    g = ref.to_grid(u, 2)
    assert g.shape == (2, 1, 4, 4)
    assert np.array_equal(g[:, :, 0:2, 2:4], u[:, 1]) and np.array_equal(g[:, :, 2:4, 0:2], u[:, 2])
    assert (g[:, :, 2:4, 2:4] == 0).all()          # the fourth cell has no video: black


def test_uint8_clip_cache_follows_datasets_py():
    """§8f rank 3: the clip cache draws sub-sequences exactly like datasets.py:72-88 (same np.random stream) and its
    uint8 batch, normalised (v-128)/128, is the reference's float32 batch (datasets.py:91-104)."""
    import numpy as np
    from mocogan_chainer_b200.chainer.dataset import concat_examples
    from mocogan_chainer_b200.datasets import Uint8ClipCache
    from oracle import mocogan_ref as ref
    rng = np.random.default_rng(0)
    lens = [16, 20, 33, 40, 17]                      # 33 and 40 > 16 * 2: the extract_speed branch
    videos = [rng.integers(0, 256, size=(n, 8, 8, 3), dtype=np.uint8) for n in lens]
    labels = [0, 1, 2, 3, 4]
    np.random.seed(7)
    cache = Uint8ClipCache(videos, labels, batch_size=3, video_length=16, extract_speed=2, shuffle=True, pin=False)
    order = cache._order.copy()
    state = np.random.get_state()
    x, t = concat_examples(cache.next())
    assert tuple(x.shape) == (3, 3, 16, 8, 8) and x.dtype.is_floating_point is False
    np.random.set_state(state)                       # replay the draws the cache made
    for b, vid in enumerate(order[:3]):
        idx = ref.subsequence_indices(lens[vid], 16, 2, lambda gap: np.random.randint(0, gap, 1)[0])
        want = ref.normalize_clip(videos[vid][idx])                       # (C, T, H, W) float32
        got = (x[b].numpy().astype(np.float32) - 128.) / 128.
        assert np.array_equal(got, want) and int(t[b]) == labels[vid]
    assert cache.epoch == 0 and not cache.is_new_epoch
    cache.next()
    assert cache.epoch == 1 and cache.is_new_epoch and cache.current_position == 1
    # linspace branch really subsamples: frames two apart
    idx = ref.subsequence_indices(40, 16, 2, lambda gap: 3)
    assert idx[0] == 3 and idx[-1] == 33 and np.all(np.diff(idx) == 2)


def test_concat_label_video_node_matches_oracle_and_slices_the_video_gradient():
    """updater.py:65-76 / :104-106 (cgan): the node's clip + label planes equal the oracle's, on a generator-style
    transposed view; backward keeps only the clip channels of both halves of the lazy video gradient."""
    from mocogan_chainer_b200.chainer import VideoGrad
    from mocogan_chainer_b200.chainer import functions as F
    from oracle import mocogan_ref as ref
    rng = np.random.default_rng(0)
    T, N, C, H, W, L = 4, 3, 3, 2, 2, 6
    x_tn = torch.from_numpy(rng.standard_normal((T, N, C, H, W)).astype(np.float32))
    lab = torch.tensor([5, 0, 2], dtype=torch.int32)
    x = Variable(x_tn.permute(1, 2, 0, 3, 4))                    # updater.py:102, a strided view
    y = F.concat_label_video(x, Variable(lab, requires_grad=False), L)
    want = ref.concat_label_video(x.data.numpy(), lab.numpy(), L)
    assert tuple(y.shape) == (N, C + L, T, H, W) and y.requires_grad
    assert np.array_equal(y.data.numpy(), want)
    assert F.physical_view(y.data).is_contiguous()               # channels-last storage for the input pass
    gv = torch.from_numpy(rng.standard_normal((N, T, H, W, C + L)).astype(np.float32))
    gi = torch.from_numpy(rng.standard_normal((N, 1, H, W, C + L)).astype(np.float32))
    fp = torch.tensor([1], dtype=torch.int32)
    out, = y.creator_node.backward((0,), (VideoGrad(gv=gv, gi=gi, frame_ptr=fp),))
    assert out.frame_ptr is fp and out.gv.is_contiguous() and out.gi.is_contiguous()
    assert torch.equal(out.gv, gv[..., :C]) and torch.equal(out.gi, gi[..., :C])
    # a real (uint8, unattached) clip is normalised as datasets.py:91 does and needs no gradient
    u8 = torch.from_numpy(rng.integers(0, 256, size=(N, C, T, H, W), dtype=np.uint8))
    yr = F.concat_label_video(Variable(u8, requires_grad=False), lab, L)
    assert not yr.requires_grad
    assert np.array_equal(yr.data.numpy(), ref.concat_label_video((u8.numpy().astype(np.float32) - 128.) / 128., lab.numpy(), L))


def _write_video(path, frames):
    from PIL import Image
    path.mkdir(parents=True, exist_ok=True)
    for j, img in enumerate(frames):
        Image.fromarray(img).save(path / "{:02d}.jpg".format(j), quality=95)


def test_mug_dataset_follows_datasets_py(tmp_path):
    """datasets.py:29-107 — directory scan, category labels, short videos discarded, sub-sequence rule with the same
    np.random draws, (v - 128) / 128, (C, T, H, W); and the uint8 clip cache built from it hands out the same clips."""
    from mocogan_chainer_b200.datasets import MugDataset, read_video_u8
    rng = np.random.default_rng(0)
    lens = {("happiness", "a"): 40, ("happiness", "b"): 20, ("fear", "c"): 33, ("fear", "short"): 10}
    for (cat, vid), n in lens.items():
        _write_video(tmp_path / cat / vid, rng.integers(0, 256, size=(n, 64, 64, 3), dtype=np.uint8))
    ds = MugDataset(tmp_path, video_length=16)
    assert len(ds) == 3 and ds.num_labels == 2 and ds.extract_speed == 2
    labels = {p.name: lab for p, lab in ds.videos}
    assert labels == {"a": 2, "b": 2, "c": 3}
    for i, (path, lab) in enumerate(ds.videos):
        stored = read_video_u8(sorted(str(p) for p in path.glob("*.jpg")))
        np.random.seed(7 + i)
        video, categ = ds.get_example(i)
        assert video.shape == (3, 16, 64, 64) and video.dtype == np.float32 and categ == lab
        draws = []
        np.random.seed(7 + i)
        idx = ref.subsequence_indices(len(stored), 16, 2, lambda gap: draws.append(gap) or np.random.randint(0, gap, 1)[0])
        assert np.array_equal(video, ref.normalize_clip(stored[idx]))
        if len(stored) == 40:          # > 16 * 2 frames: every 2nd frame (np.linspace over 31 frames)
            assert idx[-1] - idx[0] == 30
    cache = ds.clip_cache(2, shuffle=False, pin=False)
    np.random.seed(3)
    b = cache.next()
    np.random.seed(3)
    want = [ds.get_example(i) for i in (0, 1)]
    got = (b.x.numpy().astype(np.float32) - 128.) / 128.
    assert b.x.dtype == torch.uint8 and tuple(b.x.shape) == (2, 3, 16, 64, 64)
    assert np.array_equal(got, np.stack([w[0] for w in want])) and b.t.tolist() == [w[1] for w in want]


def test_moving_mnist_dataset_follows_datasets_py(tmp_path):
    """datasets.py:110-167 — (T, N, H, W) .npy written out once as 3-channel JPEG frames, contiguous 16-frame windows,
    label None."""
    from mocogan_chainer_b200.datasets import MovingMnistDataset
    arr = np.random.default_rng(1).integers(0, 256, size=(20, 3, 64, 64), dtype=np.uint8)
    np.save(tmp_path / "mnist.npy", arr)
    ds = MovingMnistDataset(tmp_path / "mnist.npy", 16, save_path=tmp_path / "pre")
    assert len(ds) == 3 and len(list((tmp_path / "pre" / "00000").glob("*.jpg"))) == 20
    np.random.seed(0)
    video, label = ds.get_example(1)
    assert video.shape == (3, 16, 64, 64) and video.dtype == np.float32 and label is None
    assert -1.0 <= video.min() and video.max() <= 127. / 128.
    assert np.array_equal(video[0], video[1])       # grey digits tiled to three channels
    ds2 = MovingMnistDataset(tmp_path / "mnist.npy", 16, save_path=tmp_path / "pre")   # second run: no re-preprocessing
    assert len(ds2) == 3


def test_clip_cache_waits_for_the_copy_that_read_a_staging_buffer():
    """A staging buffer is refilled only after the event reported for the copy that last read it has completed."""
    from mocogan_chainer_b200.datasets import Uint8ClipCache
    vids = [np.full((20, 4, 4, 3), i, dtype=np.uint8) for i in range(4)]
    cache = Uint8ClipCache(vids, [0, 1, 2, 3], 2, video_length=16, extract_speed=2, shuffle=False, pin=False)

    class Ev(object):
        def __init__(self):
            self.waited = False

        def synchronize(self):
            self.waited = True

    evs = []
    for _ in range(3):
        b = cache.next()
        ev = Ev()
        b.copied(ev)
        evs.append((b.slot, ev))
    assert not any(e.waited for _, e in evs)          # three buffers: nothing re-used yet
    b = cache.next()                                  # fourth batch re-uses the first buffer
    assert b.slot == evs[0][0] and evs[0][1].waited and not evs[1][1].waited


def test_report_keys_and_train_device_rules(monkeypatch):
    """chainer.report keys follow the optimizer names ('image_dis/loss', train.py:146-151); --gpu -1 is rejected."""
    from mocogan_chainer_b200 import chainer, train
    from mocogan_chainer_b200.chainer.training import StandardUpdater

    class L_(chainer.Link):
        pass

    class O_(object):
        def __init__(self, t):
            self.target = t

    link = L_()
    link.name = "ImageDiscriminator"
    StandardUpdater(iterator=None, optimizer={"image_dis": O_(link)})
    chainer.report({"loss": 1.5}, link)
    assert chainer.get_report()["image_dis/loss"] == 1.5
    with pytest.raises(ValueError):
        train.main(["--synthetic", "4", "--batchsize", "2"])       # default --gpu -1


def test_util_to_sequence_and_image_log_writer(tmp_path):
    """util.py:13-28 `to_sequence` (frames side by side / stacked) and the SummaryWriter stand-in's two methods."""
    import json
    from mocogan_chainer_b200 import util
    video = np.arange(3 * 2 * 4 * 5, dtype=np.uint8).reshape(3, 2, 4, 5)       # (num, channel, height, width)
    seq = util.to_sequence(video)
    assert seq.shape == (2, 4, 15) and np.array_equal(seq[:, :, 5:10], video[1])
    assert util.to_sequence(video, horizontally=False).shape == (2, 12, 5)
    w = util.ImageLogWriter(tmp_path / "runs")
    w.add_image("00th frame", np.zeros((3, 8, 8), np.uint8), 3)
    w.add_image("grey", np.full((1, 8, 8), 0.5, np.float32), 3)               # [0, 1] floats are scaled like tensorboard does
    w.add_scalar("loss:ImageGenerator", 1.25, 3)
    assert (tmp_path / "runs" / "00th_frame_000003.png").exists() and (tmp_path / "runs" / "grey_000003.png").exists()
    rec = json.loads((tmp_path / "runs" / "scalars.jsonl").read_text().strip())
    assert rec == {"tag": "loss:ImageGenerator", "value": 1.25, "step": 3}


def test_bench_traffic_is_reported_only_for_the_profiled_build(tmp_path, monkeypatch):
    """bench.load_traffic: the committed ncu figures carry the csrc sha of the build they were captured on; a different
    tree gets `stale` instead of numbers."""
    import json
    import bench
    sha = bench.csrc_sha()
    assert len(sha) == 16 and sha == bench.csrc_sha()
    prof = tmp_path / "profiles"
    prof.mkdir()
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    monkeypatch.setattr(bench, "csrc_sha", lambda: sha)
    (prof / "r02_traffic.json").write_text(json.dumps({"tc_conv_dgrad:Dv.dc2": 1.0, "csrc_sha": sha}))
    d, name = bench.load_traffic()
    assert name == "r02_traffic.json" and d["tc_conv_dgrad:Dv.dc2"] == 1.0 and not d.get("stale")
    (prof / "r02_traffic.json").write_text(json.dumps({"tc_conv_dgrad:Dv.dc2": 1.0, "csrc_sha": "0" * 16}))
    d, _ = bench.load_traffic()
    assert d.get("stale") and "tc_conv_dgrad:Dv.dc2" not in d


def test_guard_allocator_hands_out_red_zoned_tensors_and_sees_an_overrun():
    """tests/guard_alloc.py (the bounds check of tests/test_guard_gpu.py) on host tensors: shapes, strides, dtypes and
    zero-fill are those of the plain call; one byte written past a tensor is reported; torch is restored on exit."""
    from tests.guard_alloc import GUARD, guarded_allocations
    plain = torch.empty
    with guarded_allocations(cuda_only=False) as g:
        a = torch.empty((3, 5), dtype=torch.float32)
        b = torch.zeros(7, dtype=torch.bfloat16)
        like = plain(2, 3, 4).permute(2, 0, 1)
        c = torch.empty_like(like)
        assert a.data_ptr() % 64 == 0 and c.stride() == like.stride() and c.shape == like.shape
        assert float(b.float().abs().sum()) == 0.0 and float(torch.zeros_like(a).abs().sum()) == 0.0
        a.fill_(1.0)
        c.fill_(2.0)
        assert g.check() == 4
        torch.empty(10)
        g.buffers[-1][0][GUARD + 40] = 0
        with pytest.raises(AssertionError, match="1 bytes above"):
            g.check()
    assert torch.empty is plain
    with guarded_allocations(cuda_only=False, poison=True) as g:       # unwritten payloads read as NaN / -1
        assert bool(torch.isnan(torch.empty(9)).all()) and bool(torch.isnan(torch.empty(9, dtype=torch.bfloat16).float()).all())
        assert bool((torch.empty_like(plain(4, dtype=torch.int32)) == -1).all()) and float(torch.zeros(5).sum()) == 0.0
        assert g.check() == 4
