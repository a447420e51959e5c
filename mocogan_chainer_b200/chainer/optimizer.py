"""chainer.optimizer: the WeightDecay hook (train.py:96) and the GradientMethod.update() protocol."""


class WeightDecay(object):
    """g += rate * p, applied before the update rule (fused into the Adam kernel here)."""
    name = "WeightDecay"

    def __init__(self, rate):
        self.rate = rate


class Optimizer(object):
    def __init__(self):
        self.target = None
        self.t = 0
        self.epoch = 0
        self._hooks = {}

    def setup(self, link):
        self.target = link
        self.t = 0
        self._hooks = {}
        return self

    def add_hook(self, hook, name=None):
        self._hooks[name or getattr(hook, "name", hook.__class__.__name__)] = hook

    def remove_hook(self, name):
        del self._hooks[name]

    def new_epoch(self):
        self.epoch += 1
