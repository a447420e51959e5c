"""chainer.optimizers.Adam, Chainer v3.1.0 rule, one fused kernel per model."""
import torch

from .. import kernels as K
from .optimizer import Optimizer, WeightDecay


class Adam(Optimizer):
    def __init__(self, alpha=0.001, beta1=0.9, beta2=0.999, eps=1e-8):
        super(Adam, self).__init__()
        self.alpha, self.beta1, self.beta2, self.eps = alpha, beta1, beta2, eps
        self.m = self.v = self.t_dev = None
        self.grad_transform = None   # set by the data-parallel layer: all-reduce of the flat gradient
        self.grad_scale = 1.0
        self.frozen_links = ()       # links whose parameter gradients are dead work in this pass
        self.stop_variables = ()     # variables backward must not pass while this optimizer updates
        # additive: stream for the gradient all-reduce + Adam kernel (None = the caller's).  Whoever reads the updated
        # weights next must be ordered after it; the Updater puts the video discriminator's update on the stream of the
        # only branch that needs it before the end of the step.
        self.update_stream = None
        self.grad_buckets = None     # data-parallel layer: early all-reduce of the large weight gradients

    def setup(self, link):
        super(Adam, self).setup(link)
        arena = link.arena()
        self.m = torch.zeros_like(arena.data)
        self.v = torch.zeros_like(arena.data)
        self.t_dev = torch.zeros(1, dtype=torch.int32, device=arena.data.device)
        return self

    @property
    def lr(self):
        import math
        fix1 = 1.0 - math.pow(self.beta1, self.t)
        fix2 = 1.0 - math.pow(self.beta2, self.t)
        return self.alpha * math.sqrt(fix2) / fix1

    def update(self, lossfun=None, *args, **kwds):
        """GradientMethod.update: loss = lossfun(*args); target.cleargrads(); loss.backward(); hooks; t += 1; rule.
        Dead back-propagation (SURVEY.md §3.2 pts 2,4) is skipped by marking variables/links for this call only."""
        if lossfun is not None:
            loss = lossfun(*args, **kwds)
            self.target.cleargrads()
            if self.grad_buckets is not None:
                self.grad_buckets.begin()
            marked = []
            for v in self.stop_variables:
                if not v.stop:
                    v.stop = True
                    marked.append(v)
            for link in self.frozen_links:
                for p in link.params():
                    if not p.stop:
                        p.stop = True
                        marked.append(p)
            try:
                loss.backward()
            finally:
                for v in marked:
                    v.stop = False
            del loss
        arena = self.target.arena()
        wd = 0.0
        for h in self._hooks.values():
            if isinstance(h, WeightDecay):
                wd += h.rate
        self.t += 1

        def apply():
            if self.grad_transform is not None:
                self.grad_transform(arena.grad)
            K.int_add(self.t_dev, 1)
            K.adam_step(arena.data, arena.grad, self.m, self.v, arena.bf16, self.alpha, self.beta1, self.beta2, self.eps,
                        wd, self.grad_scale, self.t_dev)

        if self.update_stream is not None and torch.cuda.is_available():
            self.update_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.update_stream):
                apply()
        else:
            apply()

    def serialize(self, serializer):
        """GradientMethod.serialize of Chainer v3: `t`, `epoch`, then for every parameter its update rule's `t` and Adam
        state `m`, `v` under `<param path>/` — arrays in Chainer's layouts (the flat moment buffers are sliced and
        permuted exactly like Parameter.data)."""
        if getattr(serializer, "is_saver", False) and self.t_dev is not None:
            self.t = int(self.t_dev.item())   # a replayed CUDA graph advances only the device-side counter
        serializer("t", (self, "t"))
        serializer("epoch", (self, "epoch"))
        arena = self.target.arena()
        off = 0
        pad = lambda n: (n + 7) // 8 * 8
        for path, p in self.target.namedparams():
            sub = serializer[path]
            sub("t", (self, "t"))
            n = p.size
            for key, buf in (("m", self.m), ("v", self.v)):
                sub(key, p._to_logical(buf[off:off + n].view(p.internal_shape)))
            off += pad(n)
        assert off == arena.n
        if self.t_dev is not None:   # the device-side step counter follows a load
            self.t_dev.fill_(int(self.t))
