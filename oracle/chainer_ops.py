"""oracle/chainer_ops.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy restatement of the Chainer v3.1.0 *CPU* operators that raahii/mocogan-chainer reaches from
`model/net.py` and `model/updater.py` (reference call sites cited per function).  Chainer itself
(`requirements.txt:1`, `chainer==3.1.0`) is a third-party, un-vendored dependency that is neither under
/root/reference nor installable here, so the algorithm below restates its published CPU path
(im2col/col2im + tensordot, float32 ufuncs).

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path.
Trust in this file comes from tests/test_oracle_*.py instead: every backward here is checked against an
independent torch-CPU float64 autograd graph and against finite differences.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  Nothing under mocogan_chainer_b200/ does.

All functions are dtype-generic: float64 is the "truth" mode, float32 the "as the reference runs" mode.
"""
import itertools

import numpy as np


# ----------------------------------------------------------------------------------------------
# Initialisers  (net.py:35,131,172 GlorotNormal for every conv/deconv; Linear default LeCunNormal)
# ----------------------------------------------------------------------------------------------
def _fans(shape):
    rec = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    return shape[1] * rec, shape[0] * rec  # fan_in, fan_out


def glorot_normal(rng, shape, dtype=np.float32):
    fan_in, fan_out = _fans(shape)
    return rng.normal(0.0, np.sqrt(2.0 / (fan_in + fan_out)), size=shape).astype(dtype)


def lecun_normal(rng, shape, dtype=np.float32):
    fan_in, _ = _fans(shape)
    return rng.normal(0.0, np.sqrt(1.0 / fan_in), size=shape).astype(dtype)


# ----------------------------------------------------------------------------------------------
# Elementwise functions  (net.py:110-114,149-155,190-196; updater.py:25-26,50-51)
# ----------------------------------------------------------------------------------------------
def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def softplus(x):
    # chainer.functions.softplus(beta=1): max(x,0) + log1p(exp(-|x|))
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def leaky_relu(x, slope=0.2):
    return np.where(x >= 0, x, slope * x).astype(x.dtype)


def leaky_relu_grad(x, gy, slope=0.2):
    return np.where(x >= 0, gy, slope * gy).astype(gy.dtype)


# ----------------------------------------------------------------------------------------------
# im2col / col2im  (chainer.utils.conv_nd.im2col_nd_cpu / col2im_nd_cpu, cover_all=False)
# ----------------------------------------------------------------------------------------------
def conv_out_size(size, k, s, p):
    return (size + 2 * p - k) // s + 1


def deconv_out_size(size, k, s, p):
    return s * (size - 1) + k - 2 * p


def im2col_nd(x, ksize, stride, pad):
    n, c = x.shape[:2]
    dims = x.shape[2:]
    outs = tuple(conv_out_size(d, k, s, p) for d, k, s, p in zip(dims, ksize, stride, pad))
    xp = np.pad(x, ((0, 0), (0, 0)) + tuple((p, p + s - 1) for p, s in zip(pad, stride)), mode="constant")
    col = np.empty((n, c) + tuple(ksize) + outs, dtype=x.dtype)
    for taps in itertools.product(*[range(k) for k in ksize]):
        src = tuple(slice(kk, kk + s * o, s) for kk, s, o in zip(taps, stride, outs))
        col[(slice(None), slice(None)) + taps] = xp[(slice(None), slice(None)) + src]
    return col


def col2im_nd(col, stride, pad, dims):
    n, c = col.shape[:2]
    nd = len(dims)
    ksize = col.shape[2:2 + nd]
    outs = col.shape[2 + nd:]
    img = np.zeros((n, c) + tuple(d + 2 * p + s - 1 for d, p, s in zip(dims, pad, stride)), dtype=col.dtype)
    for taps in itertools.product(*[range(k) for k in ksize]):
        dst = tuple(slice(kk, kk + s * o, s) for kk, s, o in zip(taps, stride, outs))
        img[(slice(None), slice(None)) + dst] += col[(slice(None), slice(None)) + taps]
    crop = tuple(slice(p, p + d) for p, d in zip(pad, dims))
    return img[(slice(None), slice(None)) + crop]


# ----------------------------------------------------------------------------------------------
# L.Convolution2D / L.ConvolutionND  (net.py:133-137,174-178).  W: (out, in, *k); cross-correlation.
# ----------------------------------------------------------------------------------------------
def conv_nd_fwd(x, W, b, stride, pad):
    nd = W.ndim - 2
    ksize = W.shape[2:]
    col = im2col_nd(x, ksize, stride, pad)  # (n, c, *k, *out)
    axes = tuple(range(1, 2 + nd))
    y = np.tensordot(col, W, (axes, axes)).astype(x.dtype, copy=False)  # (n, *out, O)
    y = np.moveaxis(y, -1, 1)
    if b is not None:
        y = y + b.reshape((1, -1) + (1,) * nd)
    return np.ascontiguousarray(y)


def conv_nd_bwd(x, W, gy, stride, pad, need_gx=True, need_gw=True):
    """Returns (gx, gW, gb) — the three outputs of Chainer's conv backward (None where not requested)."""
    nd = W.ndim - 2
    ksize = W.shape[2:]
    gW = gb = gx = None
    if need_gw:
        col = im2col_nd(x, ksize, stride, pad)
        out_axes_gy = (0,) + tuple(range(2, 2 + nd))
        out_axes_col = (0,) + tuple(range(2 + nd, 2 + 2 * nd))
        gW = np.tensordot(gy, col, (out_axes_gy, out_axes_col)).astype(W.dtype, copy=False)
        gb = gy.sum(axis=(0,) + tuple(range(2, 2 + nd)))
    if need_gx:
        gcol = np.tensordot(W, gy, (0, 1)).astype(gy.dtype, copy=False)  # (c, *k, n, *out)
        gcol = np.moveaxis(gcol, 1 + nd, 0)
        gx = col2im_nd(gcol, stride, pad, x.shape[2:])
    return gx, gW, gb


# ----------------------------------------------------------------------------------------------
# L.DeconvolutionND(2, ...)  (net.py:44-48).  W: (in, out, kh, kw).
# ----------------------------------------------------------------------------------------------
def deconv_nd_fwd(x, W, b, stride, pad):
    nd = W.ndim - 2
    ksize = W.shape[2:]
    dims = tuple(deconv_out_size(d, k, s, p) for d, k, s, p in zip(x.shape[2:], ksize, stride, pad))
    gcol = np.tensordot(W, x, (0, 1)).astype(x.dtype, copy=False)  # (out, *k, n, *in_spatial)
    gcol = np.moveaxis(gcol, 1 + nd, 0)
    y = col2im_nd(gcol, stride, pad, dims)
    if b is not None:
        y = y + b.reshape((1, -1) + (1,) * nd)
    return np.ascontiguousarray(y)


def deconv_nd_bwd(x, W, gy, stride, pad, need_gx=True, need_gw=True):
    nd = W.ndim - 2
    ksize = W.shape[2:]
    gx = gW = gb = None
    if need_gx:
        gx = conv_nd_fwd(gy, W, None, stride, pad)  # W read as (O=in, C=out, *k)
    if need_gw:
        col = im2col_nd(gy, ksize, stride, pad)  # (n, out, *k, *in_spatial)
        ax_x = (0,) + tuple(range(2, 2 + nd))
        ax_col = (0,) + tuple(range(2 + nd, 2 + 2 * nd))
        gW = np.tensordot(x, col, (ax_x, ax_col)).astype(W.dtype, copy=False)
        gb = gy.sum(axis=(0,) + tuple(range(2, 2 + nd)))
    return gx, gW, gb


# ----------------------------------------------------------------------------------------------
# L.BatchNormalization, train mode  (net.py:50-53,139-141,180-182).  decay 0.9, eps 2e-5.
# ----------------------------------------------------------------------------------------------
BN_EPS = 2e-5
BN_DECAY = 0.9


def batchnorm_fwd(x, gamma, beta, avg_mean=None, avg_var=None, eps=BN_EPS, decay=BN_DECAY):
    axis = (0,) + tuple(range(2, x.ndim))
    ex = (None, slice(None)) + (None,) * (x.ndim - 2)
    mean = x.mean(axis=axis)
    var = x.var(axis=axis)
    var = var + x.dtype.type(eps)
    std = np.sqrt(var)
    x_hat = (x - mean[ex]) / std[ex]
    y = gamma[ex] * x_hat + beta[ex]
    if avg_mean is not None:
        m = x.size // gamma.size
        adjust = m / max(m - 1.0, 1.0)
        avg_mean *= decay
        avg_mean += (1 - decay) * mean
        avg_var *= decay
        avg_var += (1 - decay) * adjust * var  # CPU-path quirk: var already includes eps here
    return y.astype(x.dtype, copy=False), (mean, std)


def batchnorm_bwd(x, gamma, stats, gy):
    mean, std = stats
    axis = (0,) + tuple(range(2, x.ndim))
    ex = (None, slice(None)) + (None,) * (x.ndim - 2)
    m = x.size // gamma.size
    x_hat = (x - mean[ex]) / std[ex]
    gbeta = gy.sum(axis=axis)
    ggamma = (gy * x_hat).sum(axis=axis)
    inv_m = x.dtype.type(1.0 / m)
    gx = (gamma / std)[ex] * (gy - (x_hat * ggamma[ex] + gbeta[ex]) * inv_m)
    return gx.astype(x.dtype, copy=False), ggamma, gbeta


def batchnorm_fixed(x, gamma, beta, avg_mean, avg_var, eps=BN_EPS):
    ex = (None, slice(None)) + (None,) * (x.ndim - 2)
    return gamma[ex] * (x - avg_mean[ex]) / np.sqrt(avg_var[ex] + eps) + beta[ex]


# ----------------------------------------------------------------------------------------------
# L.StatelessGRU  (net.py:39-41,76): six Linear links W_r,U_r,W_z,U_z,W,U, all with bias.
# ----------------------------------------------------------------------------------------------
GRU_LINEARS = ("W_r", "U_r", "W_z", "U_z", "W", "U")


def gru_step_fwd(p, h, x):
    """p: dict 'W_r/W','W_r/b',...  h: (N,H)  x: (N,I).  Returns h_new and the per-step cache."""
    r = sigmoid(x @ p["W_r/W"].T + p["W_r/b"] + h @ p["U_r/W"].T + p["U_r/b"])
    z = sigmoid(x @ p["W_z/W"].T + p["W_z/b"] + h @ p["U_z/W"].T + p["U_z/b"])
    rh = r * h
    h_bar = np.tanh(x @ p["W/W"].T + p["W/b"] + rh @ p["U/W"].T + p["U/b"])
    h_new = z * h_bar + (1 - z) * h  # F.linear_interpolate(z, h_bar, h)
    return h_new, (h, x, r, z, h_bar)


def gru_step_bwd(p, cache, gh_new, grads):
    """Accumulates parameter grads into `grads`; returns (gh, gx)."""
    h, x, r, z, h_bar = cache
    gz = gh_new * (h_bar - h)
    gh_bar = gh_new * z
    gh = gh_new * (1 - z)
    ga = gh_bar * (1 - h_bar * h_bar)  # through tanh
    grads["W/W"] += ga.T @ x
    grads["W/b"] += ga.sum(0)
    grads["U/W"] += ga.T @ (r * h)
    grads["U/b"] += ga.sum(0)
    gx = ga @ p["W/W"]
    grh = ga @ p["U/W"]
    gr = grh * h
    gh = gh + grh * r
    gzp = gz * z * (1 - z)
    grp = gr * r * (1 - r)
    grads["W_z/W"] += gzp.T @ x
    grads["W_z/b"] += gzp.sum(0)
    grads["U_z/W"] += gzp.T @ h
    grads["U_z/b"] += gzp.sum(0)
    grads["W_r/W"] += grp.T @ x
    grads["W_r/b"] += grp.sum(0)
    grads["U_r/W"] += grp.T @ h
    grads["U_r/b"] += grp.sum(0)
    gx = gx + gzp @ p["W_z/W"] + grp @ p["W_r/W"]
    gh = gh + gzp @ p["U_z/W"] + grp @ p["U_r/W"]
    return gh, gx


# ----------------------------------------------------------------------------------------------
# F.softmax_cross_entropy (normalize=True)  (updater.py:36-37,55-56)
# ----------------------------------------------------------------------------------------------
def softmax_cross_entropy(x, t):
    xm = x - x.max(axis=1, keepdims=True)
    logp = xm - np.log(np.exp(xm).sum(axis=1, keepdims=True))
    n = x.shape[0]
    loss = -logp[np.arange(n), t].sum() / n
    gx = np.exp(logp)
    gx[np.arange(n), t] -= 1
    gx = gx / n
    return x.dtype.type(loss), gx.astype(x.dtype, copy=False)


# ----------------------------------------------------------------------------------------------
# chainer.optimizers.Adam + chainer.optimizer.WeightDecay  (train.py:93-101)
# ----------------------------------------------------------------------------------------------
class AdamState:
    """Chainer v3.1.0 Adam: alpha, beta1, beta2=0.999, eps=1e-8; lr_t = alpha*sqrt(1-b2^t)/(1-b1^t)."""

    def __init__(self, params, alpha=2e-4, beta1=5e-5, beta2=0.999, eps=1e-8, weight_decay=1e-5):
        self.alpha, self.beta1, self.beta2, self.eps, self.wd = alpha, beta1, beta2, eps, weight_decay
        self.t = 0
        self.m = {k: np.zeros_like(v) for k, v in params.items()}
        self.v = {k: np.zeros_like(v) for k, v in params.items()}

    def lr(self):
        import math
        fix1 = 1.0 - math.pow(self.beta1, self.t)
        fix2 = 1.0 - math.pow(self.beta2, self.t)
        return self.alpha * math.sqrt(fix2) / fix1

    def update(self, params, grads):
        """WeightDecay hook (g += rate*p), t += 1, then the in-place Adam rule on every parameter."""
        self.t += 1
        lr = self.lr()
        for k, p in params.items():
            g = grads[k]
            dt = p.dtype.type
            g = g + dt(self.wd) * p
            m, v = self.m[k], self.v[k]
            m += dt(1 - self.beta1) * (g - m)
            v += dt(1 - self.beta2) * (g * g - v)
            p -= dt(lr) * m / (np.sqrt(v) + dt(self.eps))
