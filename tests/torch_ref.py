"""tests/torch_ref.py — an independent float64 torch-CPU autograd statement of the MoCoGAN step.

Used only to cross-check the NumPy oracle (oracle/mocogan_ref.py): every gradient here comes from torch
autograd over primitive ops (conv2d / conv3d / conv_transpose2d / mean / var / softplus / cross_entropy),
never from hand-written backward code.  BatchNorm and the GRU are written out from primitives because
torch.nn.GRUCell / torch.optim.Adam use different formulas from Chainer v3.1.0.

The stale-activation semantics of updater.py:111-113 are reproduced the same way Chainer produces them:
one shared graph, three backward() calls, parameters updated in place through `.data` between them.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

EPS = 2e-5


def _bn(y, gamma, beta):
    axes = (0,) + tuple(range(2, y.dim()))
    ex = (None, slice(None)) + (None,) * (y.dim() - 2)
    mean = y.mean(dim=axes)
    var = y.var(dim=axes, unbiased=False)
    # x_hat first, then the affine map: with in-place-updated gamma this is what makes autograd reproduce
    # Chainer's BN backward (gamma read at backward time, x_hat rebuilt from the saved batch statistics).
    x_hat = (y - mean[ex]) / torch.sqrt(var[ex] + EPS)
    return gamma[ex] * x_hat + beta[ex]


def dis_forward(P, x, noises, sigma, nd, strides, pads):
    conv = F.conv2d if nd == 2 else F.conv3d
    h = x
    for i in range(1, 6):
        if i <= 4 and noises is not None:
            h = h + sigma * noises[i - 1]
        y = conv(h, P["dc%d/W" % i], P["dc%d/b" % i], stride=strides[i - 1], padding=pads[i - 1])
        if i in (2, 3, 4):
            y = _bn(y, P["bn%d/gamma" % i], P["bn%d/beta" % i])
        h = F.leaky_relu(y, 0.2) if i <= 4 else y
    return h


def gen_forward(P, gen, N, lat):
    T = gen.video_len
    f = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))
    zl = torch.eye(gen.dim_zl, dtype=torch.float64)[torch.from_numpy(lat["labels"])] if gen.dim_zl else None
    h = f(lat["h0"])
    lin = lambda name, v: v @ P["g0/%s/W" % name].T + P["g0/%s/b" % name]
    hs = []
    for t in range(T):
        et = f(lat["eps"][t])
        x = torch.cat((zl, et), dim=1) if zl is not None else et
        r = torch.sigmoid(lin("W_r", x) + lin("U_r", h))
        z = torch.sigmoid(lin("W_z", x) + lin("U_z", h))
        hb = torch.tanh(lin("W", x) + lin("U", r * h))
        h = z * hb + (1 - z) * h
        hs.append(h)
    zm = torch.stack(hs, 0)
    zc = f(lat["zc"])[None].repeat(T, 1, 1)
    x = torch.cat((zc, zm), dim=2).reshape(T * N, gen.n_hidden, 1, 1)
    for i in range(1, 6):
        x = F.conv_transpose2d(x, P["dc%d/W" % i], P["dc%d/b" % i], stride=gen.strides[i - 1],
                               padding=gen.pads[i - 1])
        x = torch.relu(_bn(x, P["bn%d/gamma" % i], P["bn%d/beta" % i])) if i < 5 else torch.tanh(x)
    return x.reshape(T, N, gen.out_channels, 64, 64)


class TorchAdam:
    def __init__(self, P, alpha=2e-4, beta1=5e-5, beta2=0.999, eps=1e-8, wd=1e-5):
        self.P, self.a, self.b1, self.b2, self.eps, self.wd, self.t = P, alpha, beta1, beta2, eps, wd, 0
        self.m = {k: torch.zeros_like(v) for k, v in P.items()}
        self.v = {k: torch.zeros_like(v) for k, v in P.items()}

    def update(self):
        self.t += 1
        lr = self.a * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for k, p in self.P.items():
            g = p.grad + self.wd * p.data
            self.m[k] += (1 - self.b1) * (g - self.m[k])
            self.v[k] += (1 - self.b2) * (g * g - self.v[k])
            p.data -= lr * self.m[k] / (torch.sqrt(self.v[k]) + self.eps)


def params_to_torch(params):
    return {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in params.items()}


def torch_step(model, G, Di, Dv, x_real, t_real, r):
    """G, Di, Dv: oracle model objects (only their hyper-parameters and *initial* params are read).
    Returns losses, the three gradient dicts and the three post-step parameter dicts (numpy float64)."""
    PG, PI, PV = params_to_torch(G.params), params_to_torch(Di.params), params_to_torch(Dv.params)
    oG, oI, oV = TorchAdam(PG), TorchAdam(PI), TorchAdam(PV)
    f = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))
    fl = lambda lst: [f(a) for a in lst]
    x_real = f(x_real)
    N, t = x_real.shape[0], r["t"]
    t_fake = None if r["latents"]["labels"] is None else torch.from_numpy(r["latents"]["labels"])
    t_real_t = None if t_real is None else torch.from_numpy(np.asarray(t_real).astype(np.int64))

    def with_labels(video, label):   # updater.py:65-76 written independently: one-hot planes mapped {0,1} -> {-1,+1}
        onehot = F.one_hot(label.long(), G.dim_zl).to(video.dtype) * 2.0 - 1.0
        planes = onehot[:, :, None, None, None].expand(-1, -1, *video.shape[2:])
        return torch.cat((video, planes), dim=1)

    if model == "cgan":
        x_real = with_labels(x_real, t_real_t)
    y_real_i = dis_forward(PI, x_real[:, :, t], fl(r["noise_i_real"]), Di.noise_sigma, 2, Di.strides, Di.pads)
    y_real_v = dis_forward(PV, x_real, fl(r["noise_v_real"]), Dv.noise_sigma, 3, Dv.strides, Dv.pads)
    x_fake_tn = gen_forward(PG, G, N, r["latents"])
    x_fake = x_fake_tn.permute(1, 2, 0, 3, 4)
    if model == "cgan":
        x_fake = with_labels(x_fake, t_fake)
    y_fake_i = dis_forward(PI, x_fake[:, :, t], fl(r["noise_i_fake"]), Di.noise_sigma, 2, Di.strides, Di.pads)
    y_fake_v = dis_forward(PV, x_fake, fl(r["noise_v_fake"]), Dv.noise_sigma, 3, Dv.strides, Dv.pads)

    def loss_dis(name, y_real, y_fake):
        loss = F.softplus(-y_real[:1]).sum() / N + F.softplus(y_fake)[:1].sum() / N
        if model == "infogan" and name == "VideoDiscriminator":
            C = y_real.shape[1]
            loss = loss + F.cross_entropy(y_real.reshape(N, C)[:, 1:], t_real_t)
            loss = loss + F.cross_entropy(y_fake.reshape(N, C)[:, 1:], t_fake)
        return loss

    def grads_of(P):
        return {k: (v.grad.detach().numpy().copy() if v.grad is not None else None) for k, v in P.items()}

    def zero(P):
        for v in P.values():
            v.grad = None

    losses, grads = {}, {}
    # pass A
    zero(PI)
    la = loss_dis("ImageDiscriminator", y_real_i, y_fake_i)
    la.backward(retain_graph=True)
    grads["image_dis"] = grads_of(PI)
    oI.update()
    # pass B
    zero(PV)
    lb = loss_dis("VideoDiscriminator", y_real_v, y_fake_v)
    lb.backward(retain_graph=True)
    grads["video_dis"] = grads_of(PV)
    oV.update()
    # pass C
    zero(PG)
    lc = F.softplus(-y_fake_i[:, 0]).sum() / N + F.softplus(-y_fake_v[:, 0]).sum() / N
    if model == "infogan":
        lc = lc + F.cross_entropy(y_fake_i[:, 1:, 0, 0], t_fake) + F.cross_entropy(y_fake_v[:, 1:, 0, 0, 0], t_fake)
    lc.backward()
    grads["image_gen"] = grads_of(PG)
    oG.update()
    losses = {"image_dis/loss": float(la.detach()), "video_dis/loss": float(lb.detach()), "image_gen/loss": float(lc.detach())}
    post = {"image_gen": {k: v.detach().numpy() for k, v in PG.items()},
            "image_dis": {k: v.detach().numpy() for k, v in PI.items()},
            "video_dis": {k: v.detach().numpy() for k, v in PV.items()}}
    return losses, grads, post, x_fake_tn.detach().numpy()
