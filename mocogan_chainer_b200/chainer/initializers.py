"""chainer.initializers used by the reference (net.py:35,131,172): GlorotNormal; Linear's default LeCunNormal.
Arrays are drawn on the host with NumPy's global RNG, as Chainer does."""
import numpy as np


def _fans(shape):
    rec = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    return shape[1] * rec, shape[0] * rec


class Initializer(object):
    def __call__(self, shape):
        raise NotImplementedError


class GlorotNormal(Initializer):
    def __init__(self, scale=1.0):
        self.scale = scale

    def __call__(self, shape):
        fan_in, fan_out = _fans(shape)
        return np.random.normal(0.0, self.scale * np.sqrt(2.0 / (fan_in + fan_out)), size=shape).astype(np.float32)


class LeCunNormal(Initializer):
    def __init__(self, scale=1.0):
        self.scale = scale

    def __call__(self, shape):
        fan_in, _ = _fans(shape)
        return np.random.normal(0.0, self.scale * np.sqrt(1.0 / fan_in), size=shape).astype(np.float32)


class Constant(Initializer):
    def __init__(self, value):
        self.value = value

    def __call__(self, shape):
        return np.full(shape, self.value, dtype=np.float32)


def generate(initializer, shape):
    if initializer is None:
        initializer = LeCunNormal()
    if isinstance(initializer, np.ndarray):
        assert tuple(initializer.shape) == tuple(shape)
        return initializer.astype(np.float32)
    if np.isscalar(initializer):
        return np.full(shape, initializer, dtype=np.float32)
    return initializer(tuple(shape))
