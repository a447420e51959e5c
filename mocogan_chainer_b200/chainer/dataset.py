"""chainer.dataset: DatasetMixin and concat_examples (updater.py:4,89)."""
import numpy as np
import torch


class DatasetMixin(object):
    def __getitem__(self, i):
        return self.get_example(i)

    def __len__(self):
        raise NotImplementedError

    def get_example(self, i):
        raise NotImplementedError


def concat_examples(batch, device=None, padding=None):
    """list of (x, label) -> (stacked x, stacked labels); labels of None stay None (datasets.py:166).  A batch object
    that is already stacked (attributes x, t — the uint8 clip cache's) is handed through."""
    if hasattr(batch, "x") and hasattr(batch, "t"):
        return batch.x, batch.t
    first = batch[0]
    if isinstance(first, tuple):
        cols = []
        for j in range(len(first)):
            col = [ex[j] for ex in batch]
            cols.append(None if col[0] is None else _stack(col, device))
        return tuple(cols)
    return _stack(batch, device)


def _stack(items, device):
    if torch.is_tensor(items[0]):
        out = torch.stack(items)
    else:
        out = np.stack([np.asarray(a) for a in items])
    return to_device(device, out) if device is not None and device >= 0 else out


def to_device(device, x):
    if x is None:
        return None
    if torch.is_tensor(x):
        return x.cuda(non_blocking=True)
    t = torch.from_numpy(np.ascontiguousarray(x))
    return t.cuda(non_blocking=True)
