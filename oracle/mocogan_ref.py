"""oracle/mocogan_ref.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement of raahii/mocogan-chainer's hot path: the three networks of `model/net.py` and the
training step `Updater.update_core` of `model/updater.py:78-113`, on top of oracle/chainer_ops.py.

PARITY UNPINNED (no reference tests / fixtures exist; Chainer 3.1.0 cannot be imported here) — see the
header of chainer_ops.py for how this file earns trust instead.

Randomness is *injected*, never replayed: every tensor the reference draws from the global NumPy RNG
(frame index t, the 16 discriminator noise tensors, labels, h0, eps_t, z_c — draw order in SURVEY.md §3.2)
is passed in explicitly so the CUDA path and this oracle consume identical numbers.
"""
import numpy as np

from . import chainer_ops as ops


def _p(d, prefix):
    n = len(prefix) + 1
    return {k[n:]: v for k, v in d.items() if k.startswith(prefix + "/")}


# ================================================================================================
# ImageGenerator  (net.py:17-117)
# ================================================================================================
class ImageGenerator:
    def __init__(self, dim_zc=50, dim_zm=10, dim_zl=0, out_channels=3, n_filters=64, video_len=16,
                 rng=None, dtype=np.float32):
        self.dim_zc, self.dim_zm, self.dim_zl = dim_zc, dim_zm, dim_zl
        self.out_channels, self.n_filters, self.video_len = out_channels, n_filters, video_len
        self.n_hidden = dim_zc + dim_zm
        self.use_label = dim_zl != 0
        self.name = "ImageGenerator"
        self.dtype = dtype
        rng = rng if rng is not None else np.random.default_rng(0)
        nf, H, I = n_filters, dim_zm, dim_zm + dim_zl
        p = {}
        for lin in ops.GRU_LINEARS:  # net.py:39-41  StatelessGRU(in=dim_zm+dim_zl, out=dim_zm)
            insz = I if lin.startswith("W") else H
            p["g0/%s/W" % lin] = ops.lecun_normal(rng, (H, insz), dtype)
            p["g0/%s/b" % lin] = np.zeros(H, dtype)
        chans = [self.n_hidden, nf * 8, nf * 4, nf * 2, nf, out_channels]  # net.py:44-48
        for i in range(5):
            p["dc%d/W" % (i + 1)] = ops.glorot_normal(rng, (chans[i], chans[i + 1], 4, 4), dtype)
            p["dc%d/b" % (i + 1)] = np.zeros(chans[i + 1], dtype)
        self.persistent = {}
        for i in range(1, 5):  # net.py:50-53
            p["bn%d/gamma" % i] = np.ones(chans[i], dtype)
            p["bn%d/beta" % i] = np.zeros(chans[i], dtype)
            self.persistent["bn%d/avg_mean" % i] = np.zeros(chans[i], dtype)
            self.persistent["bn%d/avg_var" % i] = np.zeros(chans[i], dtype)
        self.params = p
        self.strides = [(1, 1), (2, 2), (2, 2), (2, 2), (2, 2)]
        self.pads = [(0, 0), (1, 1), (1, 1), (1, 1), (1, 1)]

    @staticmethod
    def draw_latents(rng, batchsize, dim_zc, dim_zm, dim_zl, video_len, dtype=np.float32):
        """Same quantities, same order as net.py:92,66,71,102 (labels, h0, eps_1..T, zc)."""
        lat = {}
        lat["labels"] = rng.integers(0, dim_zl, size=batchsize).astype(np.int64) if dim_zl else None
        lat["h0"] = rng.normal(0, 0.33, size=(batchsize, dim_zm)).astype(dtype)
        lat["eps"] = rng.normal(0, 0.33, size=(video_len, batchsize, dim_zm)).astype(dtype)
        lat["zc"] = rng.normal(0, 0.33, size=(batchsize, dim_zc)).astype(dtype)
        return lat

    def forward(self, batchsize, lat, update_running=True, train=True):
        """train=False: chainer.config.train == False, i.e. L.BatchNormalization runs F.fixed_batch_normalization with
        the running statistics (the util.py:92 `log_tensorboard` path); no statistics are updated."""
        p, T, N, dt = self.params, self.video_len, batchsize, self.dtype
        cache = {"N": N}
        # make_zm  net.py:61-81
        zl = np.eye(self.dim_zl, dtype=dt)[lat["labels"]] if self.use_label else None
        g0 = _p(p, "g0")
        h = lat["h0"].astype(dt)
        gru_caches, hs = [], []
        for t in range(T):
            et = lat["eps"][t].astype(dt)
            xt = np.concatenate((zl, et), axis=1) if self.use_label else et
            h, c = ops.gru_step_fwd(g0, h, xt)
            gru_caches.append(c)
            hs.append(h)
        zm = np.stack(hs, axis=0)  # (T,N,dim_zm)
        zc = np.tile(lat["zc"].astype(dt)[None], (T, 1, 1))  # net.py:102-103
        z = np.concatenate((zc, zm), axis=2).reshape(T * N, self.n_hidden, 1, 1)  # net.py:106-107
        cache["gru"] = gru_caches
        x = z
        acts = []
        for i in range(1, 6):
            W, b = p["dc%d/W" % i], p["dc%d/b" % i]
            y = ops.deconv_nd_fwd(x, W, b, self.strides[i - 1], self.pads[i - 1])
            if i < 5 and not train:
                bn = ops.batchnorm_fixed(y, p["bn%d/gamma" % i], p["bn%d/beta" % i], self.persistent["bn%d/avg_mean" % i],
                                         self.persistent["bn%d/avg_var" % i])
                out = np.maximum(bn, 0)
                acts.append((x, y, None, out, bn))
            elif i < 5:
                am = self.persistent["bn%d/avg_mean" % i] if update_running else None
                av = self.persistent["bn%d/avg_var" % i] if update_running else None
                bn, stats = ops.batchnorm_fwd(y, p["bn%d/gamma" % i], p["bn%d/beta" % i], am, av)
                out = np.maximum(bn, 0)
                acts.append((x, y, stats, out, bn))
            else:
                out = np.tanh(y)
                acts.append((x, y, None, out, None))
            x = out
        cache["acts"] = acts
        cache["z"] = z
        return x.reshape(T, N, self.out_channels, 64, 64), cache  # net.py:115

    def backward(self, cache, gx):
        """gx: (T,N,C,64,64).  Returns dict of parameter gradients."""
        p, T, N = self.params, self.video_len, cache["N"]
        grads = {}
        g = gx.reshape(T * N, self.out_channels, 64, 64).astype(self.dtype)
        for i in range(5, 0, -1):
            x, y, stats, out, _ = cache["acts"][i - 1]
            if i == 5:
                g = g * (1 - out * out)
            else:
                g = g * (out > 0)
                g, gg, gb_ = ops.batchnorm_bwd(y, p["bn%d/gamma" % i], stats, g)
                grads["bn%d/gamma" % i], grads["bn%d/beta" % i] = gg, gb_
            g, gW, gb = ops.deconv_nd_bwd(x, p["dc%d/W" % i], g, self.strides[i - 1], self.pads[i - 1])
            grads["dc%d/W" % i], grads["dc%d/b" % i] = gW, gb
        gz = g.reshape(T, N, self.n_hidden)
        gzm = gz[:, :, self.dim_zc:]
        g0 = _p(p, "g0")
        gg0 = {k: np.zeros_like(v) for k, v in g0.items()}
        gh = np.zeros((N, self.dim_zm), self.dtype)
        for t in range(T - 1, -1, -1):
            gh = gh + gzm[t]
            gh, _ = ops.gru_step_bwd(g0, cache["gru"][t], gh, gg0)
        for k, v in gg0.items():
            grads["g0/" + k] = v
        return grads


# ================================================================================================
# ImageDiscriminator / VideoDiscriminator  (net.py:119-199)
# ================================================================================================
class _Discriminator:
    nd = 2

    def __init__(self, in_channels=3, out_channels=1, n_filters=64, use_noise=False, noise_sigma=0.2,
                 rng=None, dtype=np.float32):
        self.in_channels, self.out_channels, self.n_filters = in_channels, out_channels, n_filters
        self.use_noise, self.noise_sigma, self.dtype = use_noise, noise_sigma, dtype
        rng = rng if rng is not None else np.random.default_rng(0)
        nf, k = n_filters, (4,) * self.nd
        chans = [in_channels, nf, nf * 2, nf * 4, nf * 8, out_channels]
        p = {}
        for i in range(5):
            p["dc%d/W" % (i + 1)] = ops.glorot_normal(rng, (chans[i + 1], chans[i]) + k, dtype)
            p["dc%d/b" % (i + 1)] = np.zeros(chans[i + 1], dtype)
        self.persistent = {}
        for i in (2, 3, 4):
            p["bn%d/gamma" % i] = np.ones(chans[i], dtype)
            p["bn%d/beta" % i] = np.zeros(chans[i], dtype)
            self.persistent["bn%d/avg_mean" % i] = np.zeros(chans[i], dtype)
            self.persistent["bn%d/avg_var" % i] = np.zeros(chans[i], dtype)
        self.params = p

    def noise_shapes(self, x_shape):
        """Shapes of the four add_noise tensors for an input of x_shape (net.py:148,150,152,154)."""
        shapes = [tuple(x_shape)]
        n, sp = x_shape[0], tuple(x_shape[2:])
        nf = self.n_filters
        for i in range(3):
            sp = tuple(ops.conv_out_size(d, 4, s, pp) for d, s, pp in zip(sp, self.strides[i], self.pads[i]))
            shapes.append((n, nf * 2 ** i) + sp)
        return shapes

    def forward(self, x, noises=None, update_running=True):
        """noises: list of 4 arrays = sigma*randn already scaled? No: raw N(0,1) draws; sigma applied here."""
        p, dt = self.params, self.dtype
        acts = []
        h = x.astype(dt)
        for i in range(1, 6):
            if i <= 4 and self.use_noise and noises is not None:  # add_noise, net.py:10-15
                h = h + (self.noise_sigma * noises[i - 1]).astype(dt)
            y = ops.conv_nd_fwd(h, p["dc%d/W" % i], p["dc%d/b" % i], self.strides[i - 1], self.pads[i - 1])
            stats = None
            pre = y
            if i in (2, 3, 4):
                am = self.persistent["bn%d/avg_mean" % i] if update_running else None
                av = self.persistent["bn%d/avg_var" % i] if update_running else None
                pre, stats = ops.batchnorm_fwd(y, p["bn%d/gamma" % i], p["bn%d/beta" % i], am, av)
            out = ops.leaky_relu(pre, 0.2) if i <= 4 else pre
            acts.append((h, y, stats, pre))
            h = out
        return h, {"acts": acts}

    def backward(self, cache, gy, need_gx=False, need_gw=True):
        """Uses the *current* self.params (so Pass C sees updated weights with stale activations)."""
        p = self.params
        grads = {}
        g = gy.astype(self.dtype)
        for i in range(5, 0, -1):
            x_in, y, stats, pre = cache["acts"][i - 1]
            if i <= 4:
                g = ops.leaky_relu_grad(pre, g, 0.2)
            if i in (2, 3, 4):
                g, gg, gb_ = ops.batchnorm_bwd(y, p["bn%d/gamma" % i], stats, g)
                if need_gw:
                    grads["bn%d/gamma" % i], grads["bn%d/beta" % i] = gg, gb_
            want_gx = need_gx or i > 1
            g, gW, gb = ops.conv_nd_bwd(x_in, p["dc%d/W" % i], g, self.strides[i - 1], self.pads[i - 1],
                                        need_gx=want_gx, need_gw=need_gw)
            if need_gw:
                grads["dc%d/W" % i], grads["dc%d/b" % i] = gW, gb
        return grads, g


class ImageDiscriminator(_Discriminator):
    nd = 2
    name = "ImageDiscriminator"
    strides = [(2, 2)] * 4 + [(1, 1)]
    pads = [(1, 1)] * 4 + [(0, 0)]


class VideoDiscriminator(_Discriminator):
    nd = 3
    name = "VideoDiscriminator"
    strides = [(1, 2, 2)] * 4 + [(1, 3, 3)]
    pads = [(0, 1, 1)] * 4 + [(0, 0, 0)]


# ================================================================================================
# Losses  (updater.py:21-63)
# ================================================================================================
def loss_dis(model, dis_name, y_real, y_fake, t_real, t_fake):
    """Returns (loss, gy_real, gy_fake).  Keeps the reference's `[:1]` row slice (sample 0 only)."""
    n = len(y_fake)
    dt = y_real.dtype.type
    inv_n = dt(1.0 / n)
    loss = ops.softplus(-y_real[:1]).sum() * inv_n + ops.softplus(y_fake)[:1].sum() * inv_n
    gr = np.zeros_like(y_real)
    gf = np.zeros_like(y_fake)
    gr[:1] = -ops.sigmoid(-y_real[:1]) * inv_n
    gf[:1] = ops.sigmoid(y_fake[:1]) * inv_n
    if model == "infogan" and dis_name == "VideoDiscriminator":
        N, C = y_real.shape[0], y_real.shape[1]
        yr, yf = y_real.reshape(N, C), y_fake.reshape(N, C)
        l1, g1 = ops.softmax_cross_entropy(yr[:, 1:], t_real)
        l2, g2 = ops.softmax_cross_entropy(yf[:, 1:], t_fake)
        loss = loss + l1 + l2
        gr.reshape(N, C)[:, 1:] += g1
        gf.reshape(N, C)[:, 1:] += g2
    return loss, gr, gf


def loss_gen(model, y_fake_i, y_fake_v, t_fake):
    n = len(y_fake_i)
    dt = y_fake_i.dtype.type
    inv_n = dt(1.0 / n)
    loss = ops.softplus(-y_fake_i[:, 0]).sum() * inv_n + ops.softplus(-y_fake_v[:, 0]).sum() * inv_n
    gi = np.zeros_like(y_fake_i)
    gv = np.zeros_like(y_fake_v)
    gi[:, 0] = -ops.sigmoid(-y_fake_i[:, 0]) * inv_n
    gv[:, 0] = -ops.sigmoid(-y_fake_v[:, 0]) * inv_n
    if model == "infogan":
        l1, g1 = ops.softmax_cross_entropy(y_fake_i[:, 1:, 0, 0], t_fake)
        l2, g2 = ops.softmax_cross_entropy(y_fake_v[:, 1:, 0, 0, 0], t_fake)
        loss = loss + l1 + l2
        gi[:, 1:, 0, 0] += g1
        gv[:, 1:, 0, 0, 0] += g2
    return loss, gi, gv


# ================================================================================================
# Updater.update_core  (updater.py:78-113)
# ================================================================================================
def concat_label_video(video, label, dim_zl):
    """updater.py:65-76 (cgan): dim_zl planes of -1 with +1 at each clip's label plane, appended on the channel axis.
    (N,C,T,H,W) -> (N,C+dim_zl,T,H,W).  The reference indexes with a Variable (`label_video[np.arange(N), label]`),
    which Chainer v3 rejects; this is what the line means with `label` as an integer array."""
    N, C, T, H, W = video.shape
    label_video = -1.0 * np.ones((N, dim_zl, T, H, W), dtype=video.dtype)
    label_video[np.arange(N), np.asarray(label).astype(np.int64)] = 1.
    return np.concatenate((video, label_video), axis=1)


def draw_step_randoms(rng_lat, rng_noise, gen, image_dis, video_dis, batchsize, x_shape, t=None, dtype=np.float32):
    """All random tensors one step consumes, in the reference's draw order (SURVEY.md §3.2 pt 5).  `x_shape` is the
    clip's; the discriminators' input noise takes THEIR channel count (cgan: clip + label planes, updater.py:93-95)."""
    N, C, T, H, W = x_shape
    C = image_dis.in_channels
    x_shape = (N, C, T, H, W)
    r = {}
    r["t"] = int(rng_noise.integers(0, T)) if t is None else int(t)
    r["noise_i_real"] = [rng_noise.standard_normal(s).astype(dtype) for s in image_dis.noise_shapes((N, C, H, W))]
    r["noise_v_real"] = [rng_noise.standard_normal(s).astype(dtype) for s in video_dis.noise_shapes(x_shape)]
    r["latents"] = ImageGenerator.draw_latents(rng_lat, batchsize, gen.dim_zc, gen.dim_zm, gen.dim_zl,
                                               gen.video_len, dtype)
    r["noise_i_fake"] = [rng_noise.standard_normal(s).astype(dtype) for s in image_dis.noise_shapes((N, C, H, W))]
    r["noise_v_fake"] = [rng_noise.standard_normal(s).astype(dtype) for s in video_dis.noise_shapes(x_shape)]
    return r


class Updater:
    def __init__(self, model, gen, image_dis, video_dis, alpha=2e-4, beta1=5e-5, weight_decay=1e-5):
        self.model = model
        self.image_gen, self.image_dis, self.video_dis = gen, image_dis, video_dis
        self.opt = {  # train.py:93-101
            "image_gen": ops.AdamState(gen.params, alpha, beta1, weight_decay=weight_decay),
            "image_dis": ops.AdamState(image_dis.params, alpha, beta1, weight_decay=weight_decay),
            "video_dis": ops.AdamState(video_dis.params, alpha, beta1, weight_decay=weight_decay),
        }

    def update_core(self, x_real, t_real, r, as_executed=False, trace=None, d_override=None):
        """One training step.  `as_executed=True` additionally performs the back-propagation work the
        reference executes and then discards (SURVEY.md §3.2 pts 2,4) — used only for CPU-baseline timing.
        `trace`, if a dict, receives intermediate tensors for per-layer parity tests.
        `d_override` ({'image_dis': params, 'video_dis': params}) replaces the discriminators' post-update weights
        just before pass C: Adam's m/(sqrt(v)+eps) turns 1e-6-level gradient differences into 1e-5-level weight
        differences, so a test that wants to judge pass C's kernels alone feeds both sides the same updated weights."""
        G, Di, Dv = self.image_gen, self.image_dis, self.video_dis
        N = x_real.shape[0]
        t = r["t"]
        t_real = None if t_real is None else np.asarray(t_real).astype(np.int64)
        cgan = self.model == "cgan"
        Cx = x_real.shape[1]
        if cgan:   # updater.py:93-95
            x_real = concat_label_video(x_real, t_real, G.dim_zl)
        # forward — updater.py:97-108
        y_real_i, c_ri = Di.forward(x_real[:, :, t], r["noise_i_real"])
        y_real_v, c_rv = Dv.forward(x_real, r["noise_v_real"])
        x_fake_tn, c_g = G.forward(N, r["latents"])
        t_fake = r["latents"]["labels"]
        x_fake = x_fake_tn.transpose(1, 2, 0, 3, 4)  # (N,C,T,H,W), not detached
        if cgan:   # updater.py:104-106: F.concat -> its backward hands the clip channels' slice to the generator
            x_fake = concat_label_video(x_fake, t_fake, G.dim_zl)
        y_fake_i, c_fi = Di.forward(x_fake[:, :, t], r["noise_i_fake"])
        y_fake_v, c_fv = Dv.forward(x_fake, r["noise_v_fake"])
        if trace is not None:
            trace.update(y_real_i=y_real_i, y_real_v=y_real_v, x_fake=x_fake_tn, y_fake_i=y_fake_i,
                         y_fake_v=y_fake_v, cache_g=c_g, cache_ri=c_ri, cache_rv=c_rv, cache_fi=c_fi,
                         cache_fv=c_fv)

        def dead_generator_backward(gx_fake_nct):
            G.backward(c_g, gx_fake_nct[:, :Cx].transpose(2, 0, 1, 3, 4))

        # PASS A — updater.py:111
        loss_di, gr, gf = loss_dis(self.model, Di.name, y_real_i, y_fake_i, t_real, t_fake)
        g1, gxr = Di.backward(c_ri, gr, need_gx=as_executed)
        g2, gxf = Di.backward(c_fi, gf, need_gx=as_executed)
        grads_di = {k: g1[k] + g2[k] for k in g1}
        if as_executed:
            gfull = np.zeros_like(x_fake)
            gfull[:, :, t] = gxf
            dead_generator_backward(gfull)
        self.opt["image_dis"].update(Di.params, grads_di)
        # PASS B — updater.py:112
        loss_dv, gr, gf = loss_dis(self.model, Dv.name, y_real_v, y_fake_v, t_real, t_fake)
        g1, gxr = Dv.backward(c_rv, gr, need_gx=as_executed)
        g2, gxf = Dv.backward(c_fv, gf, need_gx=as_executed)
        grads_dv = {k: g1[k] + g2[k] for k in g1}
        if as_executed:
            dead_generator_backward(gxf)
        self.opt["video_dis"].update(Dv.params, grads_dv)
        # PASS C — updater.py:113: fresh D weights, stale activations
        if d_override is not None:
            for net, key in ((Di, "image_dis"), (Dv, "video_dis")):
                for k, v in d_override[key].items():
                    net.params[k][...] = v
        loss_g, gi, gv = loss_gen(self.model, y_fake_i, y_fake_v, t_fake)
        _, gx_i = Di.backward(c_fi, gi, need_gx=True, need_gw=as_executed)
        _, gx_v = Dv.backward(c_fv, gv, need_gx=True, need_gw=as_executed)
        gx_fake = gx_v.copy()
        gx_fake[:, :, t] += gx_i
        gx_fake = gx_fake[:, :Cx]   # cgan: the label planes are constants
        grads_g = G.backward(c_g, gx_fake.transpose(2, 0, 1, 3, 4))
        self.opt["image_gen"].update(G.params, grads_g)
        if trace is not None:
            trace.update(grads_di=grads_di, grads_dv=grads_dv, grads_g=grads_g, gx_fake=gx_fake)
        return {"image_dis/loss": float(loss_di), "video_dis/loss": float(loss_dv), "image_gen/loss": float(loss_g)}


def build_models(config, dtype=np.float32, seed=0, n_filters=64):
    """config: 'mnist_normal' (BASELINE config 1), 'mug_normal' (2/3), 'mug_infogan' (4), 'mug_cgan' (train.py:74-79)."""
    rng = np.random.default_rng(seed)
    if config == "mnist_normal":
        C, zl, out, model = 1, 0, 1, "normal"
    elif config == "mug_normal":
        C, zl, out, model = 3, 6, 1, "normal"
    elif config == "mug_infogan":
        C, zl, out, model = 3, 6, 7, "infogan"
    elif config == "mug_cgan":
        C, zl, out, model = 3, 6, 1, "cgan"
    else:
        raise ValueError(config)
    Cd = C + zl if model == "cgan" else C
    G = ImageGenerator(50, 10, zl, C, n_filters, 16, rng=rng, dtype=dtype)
    Di = ImageDiscriminator(Cd, out, n_filters, True, 0.2, rng=rng, dtype=dtype)
    Dv = VideoDiscriminator(Cd, out, n_filters, True, 0.2, rng=rng, dtype=dtype)
    return model, G, Di, Dv


def kink_margin(trace):
    """Smallest |pre-activation| / rms over every ReLU / LeakyReLU input of the step (generator and the four
    discriminator calls).  The gradient is discontinuous at 0, so when this margin is within float32 round-off
    (~1e-6) a float32 implementation may legitimately land on the other side of the kink than the float64 truth;
    parity tests relax their gradient bound for that case instead of failing on a measure-zero event."""
    m = np.inf
    for key in ("cache_ri", "cache_rv", "cache_fi", "cache_fv"):
        for (_, _, _, pre) in trace[key]["acts"][:4]:
            m = min(m, float(np.abs(pre).min() / np.sqrt(np.mean(pre * pre))))
    for a in trace["cache_g"]["acts"][:4]:
        bn = a[4]
        m = min(m, float(np.abs(bn).min() / np.sqrt(np.mean(bn * bn))))
    return m


# ------------------------------------------------------------------------------------------------ sample post-processing
def to_uint8(videos):
    """generate_samples.py:39 / util.py:100-101: ((videos / 2. + 0.5) * 255).astype(np.uint8) — float -> uint8 by
    truncation toward zero, videos (t, bs, c, h, w) in tanh range."""
    return ((np.asarray(videos) / 2. + 0.5) * 255).astype(np.uint8)


def to_grid(videos, size):
    """util.py:30-51 `to_grid`: (t, bs, c, h, w) uint8 -> (t, c, size*h, size*w); videos beyond bs are black."""
    t, bs, c, h, w = videos.shape
    grid = np.zeros((t, c, size * h, size * w), dtype=videos.dtype)
    for i in range(size):
        for j in range(size):
            if i * size + j < bs:
                grid[:, :, i * h:i * h + h, j * w:j * w + w] = videos[:, i * size + j]
    return grid


# ------------------------------------------------------------------------------------------------ input pipeline
def subsequence_indices(video_len, video_length, extract_speed, randint):
    """datasets.py:72-88 (MugDataset.get_example; MovingMnistDataset uses the `else` branch only, :141-147): which frames
    of a stored video make up one training clip.  `randint(gap)` stands for `np.random.randint(0, gap, 1)[0]`.
    Long videos (> video_length * extract_speed) are sampled every ~extract_speed-th frame (np.linspace with int32
    truncation), shorter ones give a contiguous window."""
    if video_len < video_length:
        raise ValueError('invalid video length: {} < {}'.format(video_len, video_length))
    if extract_speed and video_len > video_length * extract_speed:
        needed = extract_speed * (video_length - 1)
        gap = video_len - needed
        start = 0 if gap == 0 else randint(gap)
        return np.linspace(start, start + needed, video_length, endpoint=True, dtype=np.int32)
    gap = video_len - video_length
    start = 0 if gap == 0 else randint(gap)
    return np.arange(start, start + video_length)


def normalize_clip(frames_u8):
    """datasets.py:91-104: frames (T, H, W, C) uint8 -> float32 clip (C, T, H, W) = (v - 128) / 128."""
    video = (np.asarray(frames_u8, dtype=np.float32) - 128.) / 128.
    return video.transpose(3, 0, 1, 2)
