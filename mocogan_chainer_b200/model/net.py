"""model/net.py of raahii/mocogan-chainer, re-hosted on libmcg.so.

Same classes, constructor signatures, attribute and child names as the reference (net.py:17-199); the bodies of
`__call__` keep the reference's layer order but call fused nodes, so each line cites the reference lines it covers.
"""
import numpy as np

from .. import chainer
from .. import random as mrandom
from ..chainer import Variable
from ..chainer import functions as F
from ..chainer import links as L


def add_noise(x, use_noise, sigma):
    """net.py:10-15 — standalone (unfused) form: x + sigma * randn(*x.shape) while training."""
    if chainer.config.train and use_noise:
        return F.bn_act_noise(x, noise=mrandom.get_source().noise(sigma))
    return x


class ImageGenerator(chainer.Chain):
    def __init__(self, dim_zc=50, dim_zm=10, dim_zl=0, out_channels=3,
                 n_filters=64, video_len=16):
        super(ImageGenerator, self).__init__()

        self.dim_zc = dim_zc
        self.dim_zm = dim_zm
        self.dim_zl = dim_zl
        self.out_channels = out_channels
        self.n_filters = n_filters
        self.video_len = video_len

        n_hidden = dim_zc + dim_zm
        self.n_hidden = n_hidden
        self.use_label = dim_zl != 0
        self.name = self.__class__.__name__

        with self.init_scope():
            w = chainer.initializers.GlorotNormal()

            # Rm
            if self.use_label:
                self.g0 = L.StatelessGRU(self.dim_zm + self.dim_zl, self.dim_zm)
            else:
                self.g0 = L.StatelessGRU(self.dim_zm, self.dim_zm)

            # G
            self.dc1 = L.DeconvolutionND(2, n_hidden, n_filters * 8, 4, stride=1, pad=0, initialW=w)
            self.dc2 = L.DeconvolutionND(2, n_filters * 8, n_filters * 4, 4, stride=2, pad=1, initialW=w)
            self.dc3 = L.DeconvolutionND(2, n_filters * 4, n_filters * 2, 4, stride=2, pad=1, initialW=w)
            self.dc4 = L.DeconvolutionND(2, n_filters * 2, n_filters, 4, stride=2, pad=1, initialW=w)
            self.dc5 = L.DeconvolutionND(2, n_filters, out_channels, 4, stride=2, pad=1, initialW=w)

            self.bn1 = L.BatchNormalization(n_filters * 8)
            self.bn2 = L.BatchNormalization(n_filters * 4)
            self.bn3 = L.BatchNormalization(n_filters * 2)
            self.bn4 = L.BatchNormalization(n_filters)
        for dc in (self.dc1, self.dc2, self.dc3, self.dc4):
            dc.feeds_bn = True

    def make_hidden(self, batchsize, size):
        """net.py:55-56 — N(0, 0.33^2); drawn on the device (or injected) instead of NumPy-on-host + H2D."""
        return mrandom.get_source().normal((batchsize, size), 0.33)

    def to_one_hot(self, zl, xp=None):
        import torch
        return torch.eye(self.dim_zl, device=zl.device)[zl.long()]

    def make_zm(self, batchsize, labels, zc):
        """net.py:61-81 + :102-107 — h0, T x (eps_t, [zl|eps_t] -> g0), stack h_1..h_T, tile zc, concat: one kernel."""
        src = mrandom.get_source()
        h0 = self.make_hidden(batchsize, self.dim_zm)
        eps = src.normal((self.video_len, batchsize, self.dim_zm), 0.33)
        return self.g0.sequence(h0, eps, zc, labels)

    def __call__(self, batchsize, xp=np):
        """
        output shape: (video_length, batchsize, channel, x, y)
        """
        src = mrandom.get_source()
        self.arena()
        # make zl  (net.py:91-96)
        labels = src.randint(self.dim_zl, batchsize) if self.use_label else None

        # make zm, zc, [zc, zm]  (net.py:99-107).  Draw order h0, eps, zc differs from the reference's only in that
        # zc is consumed by the same fused kernel; InjectedRandom supplies them by name so parity is unaffected.
        h0 = self.make_hidden(batchsize, self.dim_zm)
        eps = src.normal((self.video_len, batchsize, self.dim_zm), 0.33)
        zc = self.make_hidden(batchsize, self.dim_zc)
        z = self.g0.sequence(h0, eps, zc, labels)
        z = F.reshape(z, (self.video_len * batchsize, self.n_hidden, 1, 1))

        # G(z)  (net.py:110-114)
        self.dc1.out_dtype = chainer.act_dtype()
        x = F.bn_act_noise(self.dc1(z), bn=self.bn1, act="relu")
        x = F.bn_act_noise(self.dc2(x), bn=self.bn2, act="relu")
        x = F.bn_act_noise(self.dc3(x), bn=self.bn3, act="relu")
        x = F.bn_act_noise(self.dc4(x), bn=self.bn4, act="relu")
        x = F.bn_act_noise(self.dc5(x), act="tanh")
        x = F.reshape(x, (self.video_len, batchsize, self.out_channels, 64, 64))  # net.py:115

        return x, labels


class _Discriminator(chainer.Chain):
    """Shared body of ImageDiscriminator / VideoDiscriminator (net.py:143-158 and :184-199 are the same sequence)."""

    def _forward(self, x, frame=None):
        self.arena()
        spec = lambda: F.add_noise_spec(self.use_noise, self.noise_sigma)
        # add_noise(x) [+ x[:,:,t] + layout/cast]                         net.py:148 / :189
        y = F.pack_video(x, frame=frame, noise=spec())
        # leaky_relu(dc1(y)); add_noise                                    net.py:149-150 / :190-191
        y = F.bn_act_noise(self.dc1(y), act="leaky_relu", slope=0.2, noise=spec())
        # leaky_relu(bn2(dc2(y))); add_noise                               net.py:151-152 / :192-193
        y = F.bn_act_noise(self.dc2(y), bn=self.bn2, act="leaky_relu", slope=0.2, noise=spec())
        y = F.bn_act_noise(self.dc3(y), bn=self.bn3, act="leaky_relu", slope=0.2, noise=spec())
        # leaky_relu(bn4(dc4(y)))  (no noise before dc5)                   net.py:155 / :196
        y = F.bn_act_noise(self.dc4(y), bn=self.bn4, act="leaky_relu", slope=0.2)
        y = self.dc5(y)                                                  # net.py:156 / :197
        return y

    def _mark(self):
        import torch
        for dc in (self.dc2, self.dc3, self.dc4):
            dc.feeds_bn = True
        self.dc5.out_dtype = torch.float32


class ImageDiscriminator(_Discriminator):
    def __init__(self, in_channels=3, out_channels=1, n_filters=64, use_noise=False, noise_sigma=0.2):
        super(ImageDiscriminator, self).__init__()

        self.in_channels = in_channels
        self.out_channels = out_channels
        self.n_filters = n_filters
        self.use_noise = use_noise
        self.noise_sigma = noise_sigma
        self.name = self.__class__.__name__

        with self.init_scope():
            w = chainer.initializers.GlorotNormal()

            self.dc1 = L.Convolution2D(in_channels, n_filters, 4, stride=2, pad=1, initialW=w)
            self.dc2 = L.Convolution2D(n_filters, n_filters * 2, 4, stride=2, pad=1, initialW=w)
            self.dc3 = L.Convolution2D(n_filters * 2, n_filters * 4, 4, stride=2, pad=1, initialW=w)
            self.dc4 = L.Convolution2D(n_filters * 4, n_filters * 8, 4, stride=2, pad=1, initialW=w)
            self.dc5 = L.Convolution2D(n_filters * 8, out_channels, 4, stride=1, pad=0, initialW=w)

            self.bn2 = L.BatchNormalization(n_filters * 2)
            self.bn3 = L.BatchNormalization(n_filters * 4)
            self.bn4 = L.BatchNormalization(n_filters * 8)
        self._mark()

    def __call__(self, x, frame=None):
        """
        input shape:  (batchsize, 3, 64, 64) — or a whole clip (batchsize, 3, T, 64, 64) with `frame` = t, which
        fuses the reference's `x[:, :, t]` (updater.py:97,107) into the input pass.
        output shape: (batchsize, out, 1, 1)
        """
        return self._forward(x, frame)


class VideoDiscriminator(_Discriminator):
    def __init__(self, in_channels=3, out_channels=1, n_filters=64, use_noise=False, noise_sigma=0.2):
        super(VideoDiscriminator, self).__init__()

        self.in_channels = in_channels
        self.out_channels = out_channels
        self.n_filters = n_filters
        self.use_noise = use_noise
        self.noise_sigma = noise_sigma
        self.name = self.__class__.__name__

        with self.init_scope():
            w = chainer.initializers.GlorotNormal()

            self.dc1 = L.ConvolutionND(3, in_channels, n_filters, 4, stride=(1, 2, 2), pad=(0, 1, 1), initialW=w)
            self.dc2 = L.ConvolutionND(3, n_filters, n_filters * 2, 4, stride=(1, 2, 2), pad=(0, 1, 1), initialW=w)
            self.dc3 = L.ConvolutionND(3, n_filters * 2, n_filters * 4, 4, stride=(1, 2, 2), pad=(0, 1, 1), initialW=w)
            self.dc4 = L.ConvolutionND(3, n_filters * 4, n_filters * 8, 4, stride=(1, 2, 2), pad=(0, 1, 1), initialW=w)
            self.dc5 = L.ConvolutionND(3, n_filters * 8, out_channels, 4, stride=(1, 3, 3), pad=(0, 0, 0), initialW=w)

            self.bn2 = L.BatchNormalization(n_filters * 2)
            self.bn3 = L.BatchNormalization(n_filters * 4)
            self.bn4 = L.BatchNormalization(n_filters * 8)
        self._mark()

    def __call__(self, x):
        """
        input shape:  (batchsize, 3, 16, 64, 64)
        output shape: (batchsize, out, 1, 1, 1)
        """
        return self._forward(x)
