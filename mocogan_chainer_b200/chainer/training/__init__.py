"""chainer.training.StandardUpdater surface used by model/updater.py:9-19,80-90."""
from .. import dataset as _dataset


class StandardUpdater(object):
    def __init__(self, iterator, optimizer, converter=_dataset.concat_examples, device=None, loss_func=None):
        self._iterators = iterator if isinstance(iterator, dict) else {"main": iterator}
        self._optimizers = optimizer if isinstance(optimizer, dict) else {"main": optimizer}
        self.converter = converter if converter is not _dataset.concat_examples else _to_device
        self.device = device
        self.iteration = 0
        # Chainer's trainer registers every optimizer's target with the reporter under the optimizer's name, which is
        # why updater.py:40,59 `chainer.report({'loss': ...}, dis)` yields 'image_dis/loss' etc. (train.py:146-148)
        for name, opt in self._optimizers.items():
            target = getattr(opt, "target", None)
            if target is not None:
                object.__setattr__(target, "report_name", name)

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
        """StandardUpdater.serialize of Chainer v3: iterators, then each optimizer and its target model, then the
        iteration count — the `updater/...` part of a trainer snapshot (train.py:137-138)."""
        for name, iterator in self._iterators.items():
            if hasattr(iterator, "serialize"):
                iterator.serialize(serializer["iterator:" + name])
        for name, optimizer in self._optimizers.items():
            optimizer.serialize(serializer["optimizer:" + name])
            optimizer.target.serialize(serializer["model:" + name])
        serializer("iteration", (self, "iteration"))


class TrainerState(object):
    """What `extensions.snapshot()` writes for `training.Trainer` and `--resume` reads back (train.py:132-138,162-163),
    restricted to the state that determines the continuation of training: everything under `updater/`.  (The
    reference's trainer also stores LogReport / trigger bookkeeping under `extensions/`; that plumbing is out of
    scope, so those keys are neither written nor required.)"""

    def __init__(self, updater):
        self.updater = updater

    def serialize(self, serializer):
        self.updater.serialize(serializer["updater"])


def _to_device(x, device=None):
    """`self.converter(x_real, self.device)` at updater.py:90 is applied to an already-stacked array."""
    return _dataset.to_device(device, x)
