"""chainer.iterators.SerialIterator (train.py:66): shuffled, repeating mini-batches with epoch bookkeeping."""
import numpy as np


class SerialIterator(object):
    def __init__(self, dataset, batch_size, repeat=True, shuffle=True):
        self.dataset, self.batch_size, self._repeat, self._shuffle = dataset, batch_size, repeat, shuffle
        self.reset()

    def reset(self):
        n = len(self.dataset)
        self._order = np.random.permutation(n) if self._shuffle else np.arange(n)
        self.current_position = 0
        self.epoch = 0
        self.is_new_epoch = False

    def __iter__(self):
        return self

    def __next__(self):
        n = len(self.dataset)
        if not self._repeat and self.epoch > 0:
            raise StopIteration
        i, i_end = self.current_position, self.current_position + self.batch_size
        batch = [self.dataset[int(j)] for j in self._order[i:i_end]]
        if i_end >= n:
            if self._repeat:
                rest = i_end - n
                if self._shuffle:
                    self._order = np.random.permutation(n)
                if rest > 0:
                    batch.extend(self.dataset[int(j)] for j in self._order[:rest])
                self.current_position = rest
            else:
                self.current_position = 0
            self.epoch += 1
            self.is_new_epoch = True
        else:
            self.is_new_epoch = False
            self.current_position = i_end
        return batch

    next = __next__

    @property
    def epoch_detail(self):
        return self.epoch + self.current_position / float(len(self.dataset))

    def serialize(self, serializer):
        """SerialIterator.serialize of Chainer v3: current_position, epoch, is_new_epoch and the shuffled order."""
        serializer("current_position", (self, "current_position"))
        serializer("epoch", (self, "epoch"))
        serializer("is_new_epoch", (self, "is_new_epoch"))
        if self._order is not None:
            serializer("order", (self, "_order"))
