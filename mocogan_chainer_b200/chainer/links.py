"""chainer.links used by model/net.py: Convolution2D, ConvolutionND, DeconvolutionND, BatchNormalization, Linear,
StatelessGRU — same constructor signatures, same child/parameter names (they are the npz keys)."""
import numpy as np

from . import Chain, Link, Parameter, config
from . import functions as F
from . import initializers


def _tup(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


class ConvolutionND(Link):
    """L.ConvolutionND(ndim, in, out, ksize, stride, pad, initialW) — W (out, in, *k), b zeros (net.py:174-178)."""
    deconv = False

    def __init__(self, ndim, in_channels, out_channels, ksize, stride=1, pad=0, nobias=False, initialW=None,
                 initial_bias=None):
        super(ConvolutionND, self).__init__()
        self.ndim = ndim
        self.ksize, self.stride, self.pad = _tup(ksize, ndim), _tup(stride, ndim), _tup(pad, ndim)
        self.out_channels = out_channels
        self.feeds_bn = False   # set by the model: bias followed by BatchNorm has an exactly-zero gradient
        self.out_dtype = None
        wshape = ((in_channels, out_channels) if self.deconv else (out_channels, in_channels)) + self.ksize
        with self.init_scope():
            self.W = Parameter(initializers.generate(initialW, wshape), channels_last_weight=True)
            self.b = None if nobias else Parameter(initializers.generate(initial_bias if initial_bias is not None else 0.0,
                                                                          (out_channels,)))

    def __call__(self, x):
        node = F.ConvolutionND(self.stride, self.pad, deconv=self.deconv, out_dtype=self.out_dtype,
                               bias_grad=not self.feeds_bn)
        return node.apply((x, self.W, self.b))[0]


class Convolution2D(ConvolutionND):
    """L.Convolution2D(in, out, ksize, stride, pad, initialW) (net.py:133-137)."""

    def __init__(self, in_channels, out_channels, ksize=None, stride=1, pad=0, nobias=False, initialW=None,
                 initial_bias=None):
        super(Convolution2D, self).__init__(2, in_channels, out_channels, ksize, stride, pad, nobias, initialW, initial_bias)


class DeconvolutionND(ConvolutionND):
    """L.DeconvolutionND(ndim, in, out, ksize, stride, pad, initialW) — W (in, out, *k) (net.py:44-48)."""
    deconv = True


class BatchNormalization(Link):
    """L.BatchNormalization(size): decay 0.9, eps 2e-5, gamma 1, beta 0, persistents avg_mean, avg_var, N."""

    def __init__(self, size, decay=0.9, eps=2e-5):
        super(BatchNormalization, self).__init__()
        self.decay, self.eps = decay, eps
        with self.init_scope():
            self.gamma = Parameter(np.ones(size, np.float32))
            self.beta = Parameter(np.zeros(size, np.float32))
        self.add_persistent("avg_mean", np.zeros(size, np.float32))
        self.add_persistent("avg_var", np.zeros(size, np.float32))
        self.add_persistent("N", 0)

    def __call__(self, x, finetune=False):
        return F.bn_act_noise(x, bn=self)


class Linear(Link):
    """L.Linear(in, out): W (out, in) LeCunNormal, b zeros.  Only ever evaluated inside the fused GRU kernel."""

    def __init__(self, in_size, out_size, initialW=None, initial_bias=None):
        super(Linear, self).__init__()
        with self.init_scope():
            self.W = Parameter(initializers.generate(initialW, (out_size, in_size)))
            self.b = Parameter(initializers.generate(initial_bias if initial_bias is not None else 0.0, (out_size,)))


class StatelessGRU(Chain):
    """L.StatelessGRU(in_size, out_size) (net.py:39-41): children W_r, U_r, W_z, U_z, W, U, all with bias.
    `sequence()` runs the whole T-step recurrence of make_zm (net.py:61-81) in one persistent kernel."""

    def __init__(self, in_size, out_size):
        super(StatelessGRU, self).__init__()
        self.in_size, self.out_size = in_size, out_size
        with self.init_scope():
            self.W_r = Linear(in_size, out_size)
            self.U_r = Linear(out_size, out_size)
            self.W_z = Linear(in_size, out_size)
            self.U_z = Linear(out_size, out_size)
            self.W = Linear(in_size, out_size)
            self.U = Linear(out_size, out_size)

    def param_list(self):
        return [p for n in ("W_r", "U_r", "W_z", "U_z", "W", "U") for p in (getattr(self, n).W, getattr(self, n).b)]

    def sequence(self, h0, eps, zc, labels=None):
        n_labels = self.in_size - self.out_size
        node = F.GRUSequence(labels, n_labels, h0, eps, zc)
        return node.apply(tuple(self.param_list()))[0]

    def __call__(self, h, x):
        raise NotImplementedError("single-step StatelessGRU is not on the hot path; use sequence() (net.py:61-81 fused)")
