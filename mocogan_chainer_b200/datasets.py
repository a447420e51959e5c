"""Datasets.  Only the OUTPUT CONTRACT of the reference's datasets.py matters to the hot path:
float32 clips (C, T, H, W) in [-1, 0.992] and an int label or None (datasets.py:105-107,162-166).  The JPEG-decoding
MUG / Moving-MNIST readers are out of scope; SyntheticClipDataset produces clips of that contract."""
import numpy as np

from .chainer.dataset import DatasetMixin


class SyntheticClipDataset(DatasetMixin):
    def __init__(self, n=140, channels=3, video_len=16, size=64, num_labels=6, seed=1234):
        rng = np.random.default_rng(seed)
        self.x = rng.uniform(-1, 1, size=(n, channels, video_len, size, size)).astype(np.float32)
        self.t = rng.integers(0, num_labels, size=n).astype(np.int32) if num_labels else None

    def __len__(self):
        return len(self.x)

    def get_example(self, i):
        return self.x[i], (None if self.t is None else int(self.t[i]))
