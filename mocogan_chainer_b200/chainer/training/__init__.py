"""chainer.training.StandardUpdater surface used by model/updater.py:9-19,80-90."""
from .. import dataset as _dataset


class StandardUpdater(object):
    def __init__(self, iterator, optimizer, converter=_dataset.concat_examples, device=None, loss_func=None):
        self._iterators = iterator if isinstance(iterator, dict) else {"main": iterator}
        self._optimizers = optimizer if isinstance(optimizer, dict) else {"main": optimizer}
        self.converter = converter if converter is not _dataset.concat_examples else _to_device
        self.device = device
        self.iteration = 0

    @property
    def epoch(self):
        return self._iterators["main"].epoch

    @property
    def epoch_detail(self):
        return self._iterators["main"].epoch_detail

    @property
    def is_new_epoch(self):
        return self._iterators["main"].is_new_epoch

    def get_optimizer(self, name):
        return self._optimizers[name]

    def get_all_optimizers(self):
        return dict(self._optimizers)

    def get_iterator(self, name):
        return self._iterators[name]

    def update(self):
        self.update_core()
        self.iteration += 1

    def update_core(self):
        raise NotImplementedError

    def serialize(self, serializer):
        serializer("iteration", (self, "iteration"))


def _to_device(x, device=None):
    """`self.converter(x_real, self.device)` at updater.py:90 is applied to an already-stacked array."""
    return _dataset.to_device(device, x)
