"""`mxnet.gluon.data` — a minimal DataLoader over an indexable dataset (the synthetic datasets of compat/data)."""
import numpy as np

from .. import ndarray as nd
from ..context import cpu


class Dataset:
    def __len__(self):
        raise NotImplementedError

    def __getitem__(self, idx):
        raise NotImplementedError


class DataLoader:
    def __init__(self, dataset, batch_size=1, shuffle=False, last_batch="discard", num_workers=0, **kwargs):
        self.dataset, self.batch_size, self.shuffle = dataset, batch_size, shuffle

    def __len__(self):
        return len(self.dataset) // self.batch_size

    def __iter__(self):
        order = np.arange(len(self.dataset))
        if self.shuffle:
            np.random.shuffle(order)
        for b in range(len(self)):
            items = [self.dataset[int(i)] for i in order[b * self.batch_size:(b + 1) * self.batch_size]]
            yield tuple(nd.array(np.stack([it[k] for it in items]), ctx=cpu()) for k in range(len(items[0])))
