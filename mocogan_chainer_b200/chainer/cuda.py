"""chainer.cuda: the handful of entry points the reference calls (net.py:11, updater.py:91, train.py:88)."""
import numpy as np
import torch


class _XP(object):
    """Stand-in for the `xp` module object the reference threads through (numpy or cupy).  Device arrays are torch
    tensors; random draws go through mocogan_chainer_b200.random so they can be device-side or injected."""
    float32 = np.float32

    @staticmethod
    def asarray(a, dtype=None):
        if torch.is_tensor(a):
            return a
        t = torch.from_numpy(np.asarray(a))
        return t.cuda()


cupy = _XP()


def get_array_module(*args):
    return cupy


class _Dev(object):
    def __init__(self, i):
        self.id = i

    def use(self):
        torch.cuda.set_device(self.id)


def get_device_from_id(i):
    return _Dev(i)


def to_gpu(a, device=None):
    return cupy.asarray(a)


def to_cpu(a):
    return a.detach().cpu().numpy() if torch.is_tensor(a) else a
