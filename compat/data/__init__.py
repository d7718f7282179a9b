"""`from data import ClipBatchIter, get_ucf101trainval, get_simple_meitu_dataloader` with SYNTHETIC clips.

Video decode / datasets are out of scope (OpenCV / pynvvl / FFmpeg plumbing, datasets not available);
the hot path is exercised with clips of the right shape and distribution: U[0,1) per value then per-batch per-channel
normalisation as videos_reader.py:93-97 (ClipBatchIter) / ImageNet statistics as data/ucf101.py:124-128 (gluon loaders),
labels as the datasets produce them (class index; multi-hot with 1-4 tags, data/simple_meitu.py:134-136).
FVT_COMPAT_CLIPS sets the number of synthetic clips per split (default 24)."""
import os
import random

import numpy as np

import mxnet as mx

_N = int(os.environ.get("FVT_COMPAT_CLIPS", "24"))
_MEAN = np.array([0.485, 0.456, 0.406], np.float32).reshape(3, 1, 1, 1)
_STD = np.array([0.229, 0.224, 0.225], np.float32).reshape(3, 1, 1, 1)


def _clip(seed, n_frame, crop):
    return np.random.default_rng(seed).random((3, n_frame, crop, crop), dtype=np.float32)


class ClipBatchIter(mx.io.DataIter):
    """data/data.py:18-108: UCF101 batches for the Module API (train.py:60-66, validation.py:26-27)."""

    def __init__(self, datadir, batch_size=8, n_frame=32, crop_size=112, scale_w=171, scale_h=128, train=True, temporal_center=False):
        super(ClipBatchIter, self).__init__(batch_size)
        self.datadir, self.batch_size, self.n_frame, self.crop_size = datadir, batch_size, n_frame, crop_size
        self.train = train
        rng = np.random.default_rng(1 if train else 2)
        self.clip_lst = [("synthetic/%s_%04d.avi" % ("train" if train else "test", i), int(rng.integers(0, 101))) for i in range(_N)]
        self.reset()

    @property
    def provide_data(self):
        return [mx.io.DataDesc(name="data", shape=(self.batch_size, 3, self.n_frame, self.crop_size, self.crop_size), dtype=np.float32, layout="NCDHW")]

    @property
    def provide_label(self):
        return [mx.io.DataDesc(name="softmax_label", shape=(self.batch_size,), dtype=np.float32, layout="N")]

    def reset(self):
        self.clip_p = 0
        if self.train:
            random.shuffle(self.clip_lst)

    def next(self):
        if self.clip_p >= len(self.clip_lst):
            raise StopIteration
        batch = self.clip_lst[self.clip_p: self.clip_p + self.batch_size]
        if len(batch) < self.batch_size:
            batch += random.sample(self.clip_lst, self.batch_size - len(batch))
        names, labels = zip(*batch)
        data = np.stack([_clip(abs(hash(nm)) % (1 << 31), self.n_frame, self.crop_size) for nm in names])
        m = data.mean(axis=(0, 2, 3, 4), keepdims=True)                      # videos_reader.py:93-97
        s = data.std(axis=(0, 2, 3, 4), keepdims=True)
        data = (data - m) / (s + 1e-3)
        self.clip_p += self.batch_size
        return mx.io.DataBatch([mx.nd.array(data)], [mx.nd.array(labels)])


class _Synthetic(mx.gluon.data.Dataset):
    def __init__(self, n, n_frame, crop, num_class, multilabel, seed):
        self.n, self.n_frame, self.crop, self.num_class, self.multilabel, self.seed = n, n_frame, crop, num_class, multilabel, seed

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        rng = np.random.default_rng(self.seed * 100003 + idx)
        x = (_clip(self.seed * 7919 + idx, self.n_frame, self.crop) - _MEAN) / _STD       # data/ucf101.py:124-128
        if self.multilabel:
            y = np.zeros(self.num_class, np.float32)
            y[rng.choice(self.num_class, size=int(rng.integers(1, 5)), replace=False)] = 1.0
        else:
            y = np.float32(rng.integers(0, self.num_class))
        return x.astype(np.float32), y


def get_ucf101trainval(datadir, batch_size=8, n_frame=32, crop_size=112, scale_h=128, scale_w=171, num_workers=6):
    """data/ucf101.py:130-148 -> (train_loader, val_loader) of (clips (B,3,T,H,W), labels (B,))."""
    mk = lambda seed, shuffle: mx.gluon.data.DataLoader(_Synthetic(_N, n_frame, crop_size, 101, False, seed), batch_size=batch_size, shuffle=shuffle)   # noqa: E731
    return mk(1, True), mk(2, False)


def get_simple_meitu_dataloader(datadir, batch_size=4, n_frame=32, crop_size=112, scale_h=128, scale_w=171, num_workers=6):
    """data/simple_meitu.py:145-164 -> (train_loader, val_loader) of (clips, multi-hot tags (B, 63))."""
    mk = lambda seed, shuffle: mx.gluon.data.DataLoader(_Synthetic(_N, n_frame, crop_size, 63, True, seed), batch_size=batch_size, shuffle=shuffle)    # noqa: E731
    return mk(3, True), mk(4, False)
