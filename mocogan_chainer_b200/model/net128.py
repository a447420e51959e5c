"""EXTENSION — not part of raahii/mocogan-chainer.  A 128x128 variant of the frame generator, so that BASELINE config 5
("32-frame 128x128 clips, batch 256") can be answered as it is worded.

The reference's ImageGenerator cannot produce 128x128 frames: its five deconvolutions end at 64x64 and net.py:115
hard-codes `64, 64` in the final reshape.  The architecture below is the obvious continuation
of net.py:44-53 — ONE more stride-2 deconvolution stage in front of the output layer, the first layer widened to
n_filters*16 so that every stage still halves the channel count:

    z (B, dim_zc + dim_zm, 1, 1)
      dc1  deconv 4x4 s1 p0   -> (B, 16nf,   4,   4)  bn1  relu
      dc2  deconv 4x4 s2 p1   -> (B,  8nf,   8,   8)  bn2  relu
      dc3  deconv 4x4 s2 p1   -> (B,  4nf,  16,  16)  bn3  relu
      dc4  deconv 4x4 s2 p1   -> (B,  2nf,  32,  32)  bn4  relu
      dc5  deconv 4x4 s2 p1   -> (B,   nf,  64,  64)  bn5  relu
      dc6  deconv 4x4 s2 p1   -> (B,    C, 128, 128)  tanh

PARITY UNPINNED: there is no reference implementation, weights or output to compare with; the layers are the same
FunctionNodes (and kernels) that the parity-tested 64x64 generator uses, and tests/test_surface_gpu.py checks this
class against the oracle's deconvolution / BatchNorm restatement composed the same way.  The motion path (GRU, latent
draws) is inherited unchanged from ImageGenerator (net.py:55-107).
"""
import numpy as np

from .. import chainer
from .. import random as mrandom
from ..chainer import functions as F
from ..chainer import links as L
from .net import ImageGenerator


class ImageGenerator128(ImageGenerator):
    size = 128

    def __init__(self, dim_zc=50, dim_zm=10, dim_zl=0, out_channels=3, n_filters=64, video_len=16):
        super(ImageGenerator128, self).__init__(dim_zc, dim_zm, dim_zl, out_channels, n_filters, video_len)
        nf = n_filters
        with self.init_scope():
            w = chainer.initializers.GlorotNormal()
            self.dc1 = L.DeconvolutionND(2, self.n_hidden, nf * 16, 4, stride=1, pad=0, initialW=w)
            self.dc2 = L.DeconvolutionND(2, nf * 16, nf * 8, 4, stride=2, pad=1, initialW=w)
            self.dc3 = L.DeconvolutionND(2, nf * 8, nf * 4, 4, stride=2, pad=1, initialW=w)
            self.dc4 = L.DeconvolutionND(2, nf * 4, nf * 2, 4, stride=2, pad=1, initialW=w)
            self.dc5 = L.DeconvolutionND(2, nf * 2, nf, 4, stride=2, pad=1, initialW=w)
            self.dc6 = L.DeconvolutionND(2, nf, out_channels, 4, stride=2, pad=1, initialW=w)
            self.bn1 = L.BatchNormalization(nf * 16)
            self.bn2 = L.BatchNormalization(nf * 8)
            self.bn3 = L.BatchNormalization(nf * 4)
            self.bn4 = L.BatchNormalization(nf * 2)
            self.bn5 = L.BatchNormalization(nf)
        # the base class registered dc1..dc5 / bn1..bn4 already; re-assigning replaced the links, only the new names are added
        for name in ("_children",):
            seen, uniq = set(), []
            for n in getattr(self, name):
                if n not in seen:
                    seen.add(n)
                    uniq.append(n)
            setattr(self, name, uniq)
        for dc in (self.dc1, self.dc2, self.dc3, self.dc4, self.dc5):
            dc.feeds_bn = True

    def forward_gflop_per_frame(self):
        """2 * Cin * Cout * 16 taps per INPUT pixel of every deconvolution (Appendix C's counting)."""
        nf, h = self.n_filters, self.n_hidden
        chans = [h, nf * 16, nf * 8, nf * 4, nf * 2, nf, self.out_channels]
        in_px = [1, 16, 64, 256, 1024, 4096]
        return sum(2.0 * chans[i] * chans[i + 1] * 16 * in_px[i] for i in range(6)) / 1e9

    def __call__(self, batchsize, xp=np):
        """output shape: (video_length, batchsize, channel, 128, 128)"""
        src = mrandom.get_source()
        self.arena()
        labels = src.randint(self.dim_zl, batchsize) if self.use_label else None
        h0 = self.make_hidden(batchsize, self.dim_zm)
        eps = src.normal((self.video_len, batchsize, self.dim_zm), 0.33)
        zc = self.make_hidden(batchsize, self.dim_zc)
        z = self.g0.sequence(h0, eps, zc, labels)
        z = F.reshape(z, (self.video_len * batchsize, self.n_hidden, 1, 1))
        self.dc1.out_dtype = chainer.act_dtype()
        x = F.bn_act_noise(self.dc1(z), bn=self.bn1, act="relu")
        x = F.bn_act_noise(self.dc2(x), bn=self.bn2, act="relu")
        x = F.bn_act_noise(self.dc3(x), bn=self.bn3, act="relu")
        x = F.bn_act_noise(self.dc4(x), bn=self.bn4, act="relu")
        x = F.bn_act_noise(self.dc5(x), bn=self.bn5, act="relu")
        x = F.bn_act_noise(self.dc6(x), act="tanh")
        x = F.reshape(x, (self.video_len, batchsize, self.out_channels, self.size, self.size))
        return x, labels
