"""`mx.io` — DataIter / DataBatch / DataDesc / PrefetchingIter (data/data.py:18-108, train.py:60-66, validation.py:26-28)."""
from collections import namedtuple


class DataDesc(namedtuple("DataDesc", ["name", "shape"])):
    def __new__(cls, name, shape, dtype=None, layout="NCHW"):
        ret = super().__new__(cls, name, tuple(shape))
        ret.dtype, ret.layout = dtype, layout
        return ret


class DataBatch:
    def __init__(self, data, label=None, pad=None, index=None, provide_data=None, provide_label=None):
        self.data, self.label, self.pad, self.index = data, label, pad, index
        self.provide_data, self.provide_label = provide_data, provide_label


class DataIter:
    def __init__(self, batch_size=0):
        self.batch_size = batch_size

    def __iter__(self):
        return self

    def reset(self):
        pass

    def next(self):
        raise StopIteration

    def __next__(self):
        return self.next()


class PrefetchingIter(DataIter):
    """The reference prefetches with a thread; the synthetic iterators here have nothing to overlap, so this passes through."""

    def __init__(self, iters, rename_data=None, rename_label=None):
        self.iter = iters[0] if isinstance(iters, (list, tuple)) else iters
        super().__init__(getattr(self.iter, "batch_size", 0))

    @property
    def provide_data(self):
        return self.iter.provide_data

    @property
    def provide_label(self):
        return self.iter.provide_label

    def reset(self):
        self.iter.reset()

    def next(self):
        return self.iter.next()
