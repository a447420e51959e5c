"""Sources of the random numbers one training step draws (SURVEY.md §3.2 pt 5 lists the reference's draw order).

DeviceRandom   — product mode: counter-based Philox inside the kernels, keyed by a device-resident (seed, step)
                 so that a captured CUDA graph replays with fresh numbers and nothing round-trips to the host
                 (the reference draws float64 NumPy/MT19937 numbers on the host: net.py:13,56,92; updater.py:96).
InjectedRandom — parity mode: the tensors are supplied by the caller (generated once by the oracle) and consumed
                 in the reference's draw order, so both sides see identical numbers.  MT19937 stream parity on the
                 GPU is a non-goal.
"""
import torch

from . import kernels as K


class DeviceRandom(object):
    def __init__(self, seed=0, device="cuda", video_length=16):
        self.device = device
        self.state = K.step_state_new(seed, device)
        self.video_length = video_length
        self._call = 0

    def begin_step(self):
        K.step_advance(self.state, self.video_length)
        self._call = 0

    def _next_id(self):
        self._call += 1
        return self._call

    def frame(self):
        """updater.py:96 `t = xp.random.randint(0, video_length)` — a device int32[1] view, never read on the host."""
        return self.state[3:4]

    def noise(self, sigma):
        return ("philox", sigma, self.state, self._next_id())

    def normal(self, shape, sigma):
        out = torch.empty(tuple(shape), device=self.device)
        K.randn(out, float(sigma), self.state, self._next_id())
        return out

    def randint(self, high, n):
        out = torch.empty(int(n), dtype=torch.int32, device=self.device)
        K.randint(out, int(high), self.state, self._next_id())
        return out


class InjectedRandom(object):
    """r: the dict produced by oracle.mocogan_ref.draw_step_randoms (numpy arrays), consumed in draw order."""

    def __init__(self, r, device="cuda"):
        f = lambda a: torch.from_numpy(a).float().to(device).contiguous()
        self.device = device
        self.t = int(r["t"])
        self._frame = torch.tensor([self.t], dtype=torch.int32, device=device)
        self._noise = [f(a) for key in ("noise_i_real", "noise_v_real", "noise_i_fake", "noise_v_fake") for a in r[key]]
        lat = r["latents"]
        self._labels = None if lat["labels"] is None else torch.from_numpy(lat["labels"]).int().to(device)
        self._normals = [f(lat["h0"]), f(lat["eps"]), f(lat["zc"])]

    def begin_step(self):
        pass

    def frame(self):
        return self._frame

    def noise(self, sigma):
        return ("tensor", sigma, self._noise.pop(0))

    def normal(self, shape, sigma):
        t = self._normals.pop(0)
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))
        return t

    def randint(self, high, n):
        return self._labels


_source = None


def set_source(src):
    global _source
    _source = src
    return src


def get_source():
    global _source
    if _source is None:
        _source = DeviceRandom(0)
    return _source
