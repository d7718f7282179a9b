"""gluon.Trainer-shaped optimiser front end (reference train_simple_r3d.py:95-97,106,124) with the MXNet kvstore
gradient reduction replaced by NCCL all-reduce over NVLink (one process per GPU, torch.distributed for the plumbing).

    trainer = Trainer(net, 'sgd', {'learning_rate': lr, 'momentum': 0.9, 'wd': wd}, kvstore='device')
    logits = net(x); loss = criterion(logits, y); loss.backward(); trainer.step(global_batch)

step(batch_size) = [all-reduce(sum) of the flat gradient buffer, launched bucket by bucket DURING backward as layers
finish, reverse layer order] -> one fused SGD-momentum launch with rescale = 1/batch_size (MXNet sgd_mom_update:
g' = rescale*g + wd*w; mom = momentum*mom - lr*g'; w += mom).  Every rank applies the identical update, which is what
MXNet's update-on-kvstore + pull amounts to.  BatchNorm running statistics are not reduced (per-device, as in the
reference).
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .ops import _ptr, _stream

_SGD_DTYPE = np.dtype([("w", "<u8"), ("g", "<u8"), ("m", "<u8"), ("numel", "<u8"), ("wd", "<f4"), ("lr_mult", "<f4")])
_CHUNK = 1 << 16


class Trainer:
    def __init__(self, net, optimizer="sgd", optimizer_params=None, kvstore="device", bucket_bytes=25 << 20,
                 wd_policy="gluon", tail_bytes=6 << 20):
        if optimizer != "sgd":
            raise NotImplementedError("the reference trains with 'sgd' only (train_simple_r3d.py:95, train.py:69)")
        op = dict(optimizer_params or {})
        self._lr = float(op.get("learning_rate", 0.01))
        self.momentum = float(op.get("momentum", 0.0))
        self.wd = float(op.get("wd", 0.0))
        self.net = net
        self.kvstore = kvstore
        self.bucket_elems = max(1, bucket_bytes // 4)
        # The gradients of the FIRST layers (stem, conv2_x: < 1 M elements) become final last, and whatever is reduced after
        # them is exposed at the end of backward.  Everything pending is therefore flushed once when backward reaches the
        # first `tail_bytes` of the buffer (that all-reduce overlaps the conv2_x backward, the longest part of the step),
        # which leaves only a few MB for the exposed, final all-reduce (round 1: up to a full 25 MB bucket, 0.64 ms at N = 8).
        self.tail_elems = max(0, tail_bytes // 4)
        self._tail_flushed = False
        self.wd_policy = wd_policy           # 'gluon': wd on every tensor; 'module': only *_weight and *_gamma (MXNet Module API)
        self._table = None
        self._handles = []
        self._pending = None
        self._table_key = None
        net._attach_trainer(self)

    @property
    def _distributed(self):
        """Evaluated lazily: the process group may be initialised after the Trainer is built."""
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    # ---- gluon API
    @property
    def learning_rate(self):
        return self._lr

    def set_learning_rate(self, lr):
        self._lr = float(lr)

    # ---- gradient reduction, overlapped with backward
    def on_grads_ready(self, lo, hi):
        """Called by TrainPlan.backward when flat.g[lo:hi) is final.  Ready ranges arrive from the END of the buffer
        towards its start (reverse layer order), so pending ranges stay contiguous."""
        if not self._distributed:
            return
        if self._pending is None:
            self._pending = [lo, hi]
        else:
            self._pending[0] = min(self._pending[0], lo)
            self._pending[1] = max(self._pending[1], hi)
        if self._pending[1] - self._pending[0] >= self.bucket_elems:
            self._flush()
        elif not self._tail_flushed and self._pending[0] <= self.tail_elems:
            self._tail_flushed = True
            self._flush()

    def _flush(self):
        if self._pending is None:
            return
        lo, hi = self._pending
        self._pending = None
        g = self.net._flat.g[lo:hi]
        self._handles.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True))

    # ---- update
    def _build_table(self):
        flat = self.net._flat
        rows, chunk_t, chunk_o = [], [], []
        for i, (name, (off, numel, shape, store)) in enumerate(flat.slots.items()):
            wd = self.wd
            if self.wd_policy == "module" and not (name.endswith("_weight") or name.endswith("_gamma")):
                wd = 0.0
            base = flat.w.data_ptr() + 4 * off, flat.g.data_ptr() + 4 * off, flat.m.data_ptr() + 4 * off
            rows.append((base[0], base[1], base[2], store, wd, 1.0))
            nchunks = (store + _CHUNK - 1) // _CHUNK
            chunk_t += [i] * nchunks
            chunk_o += list(range(nchunks))
        table = np.array(rows, dtype=_SGD_DTYPE)
        dev = flat.w.device
        self._table = torch.from_numpy(table.view(np.uint8).copy()).to(dev)
        self._chunk_t = torch.tensor(chunk_t, dtype=torch.int32, device=dev)
        self._chunk_o = torch.tensor(chunk_o, dtype=torch.int32, device=dev)
        self._nchunks = len(chunk_t)

    def allreduce_grads(self):
        self._flush()
        for h in self._handles:
            h.wait()
        self._handles = []
        self._tail_flushed = False

    def step(self, batch_size, ignore_stale_grad=False):
        net = self.net
        if net._flat is None:
            raise RuntimeError("Trainer.step() before any training-mode forward/backward")
        self.allreduce_grads()
        key = (net._flat.w.data_ptr(), net._flat.g.data_ptr(), net._flat.m.data_ptr())
        if self._table is None or key != self._table_key:       # the table caches raw addresses of the flat buffers
            self._build_table()
            self._table_key = key
        lib = _lib.load()
        _lib.check(lib.fvt_sgd_momentum_multi(_lib.handle(), _ptr(self._table), _ptr(self._chunk_t), _ptr(self._chunk_o), self._nchunks,
                                              _CHUNK, ctypes.c_float(self._lr), ctypes.c_float(self.momentum),
                                              ctypes.c_float(1.0 / float(batch_size)), _stream()))
        net._weights_changed()


# ------------------------------------------------------------------------------------------------ LR schedules
class FactorScheduler:
    """`mx.lr_scheduler.FactorScheduler(step, factor)` as used at reference train.py:75-77: the rate is multiplied by
    `factor` every `step` updates (MXNet semantics: applied once `num_update` EXCEEDS count + step; floor
    `stop_factor_lr`).  Third-party MXNet behaviour, restated."""

    def __init__(self, step, factor=1.0, stop_factor_lr=1e-8, base_lr=0.01):
        if step < 1:
            raise ValueError("Schedule step must be greater or equal than 1 round")
        if factor > 1.0:
            raise ValueError("Factor must be no more than 1 to make lr reduce")
        self.step, self.factor, self.stop_factor_lr, self.base_lr = step, factor, stop_factor_lr, base_lr
        self.count = 0

    def __call__(self, num_update):
        while num_update > self.count + self.step:
            self.count += self.step
            self.base_lr *= self.factor
            if self.base_lr < self.stop_factor_lr:
                self.base_lr = self.stop_factor_lr
        return self.base_lr


class MultiFactorScheduler:
    """`mx.lr_scheduler.MultiFactorScheduler(step=[...], factor)` — reference train_simple_r3d.py:99-100, evaluated once
    per EPOCH there (`trainer.set_learning_rate(lr_sch(epoch))`, :106): the rate is multiplied by `factor` each time
    `num_update` exceeds the next entry of `step`."""

    def __init__(self, step, factor=1.0, base_lr=0.01):
        if not isinstance(step, (list, tuple)) or len(step) < 1:
            raise ValueError("step must be a non-empty list")
        for i, s in enumerate(step):
            if i != 0 and step[i] <= step[i - 1]:
                raise ValueError("Schedule step must be an increasing integer list")
            if s < 1:
                raise ValueError("Schedule step must be greater or equal than 1 round")
        if factor > 1.0:
            raise ValueError("Factor must be no more than 1 to make lr reduce")
        self.step, self.factor, self.base_lr = list(step), factor, base_lr
        self.cur_step_ind, self.count = 0, 0

    def __call__(self, num_update):
        while self.cur_step_ind <= len(self.step) - 1:
            if num_update > self.step[self.cur_step_ind]:
                self.count = self.step[self.cur_step_ind]
                self.cur_step_ind += 1
                self.base_lr *= self.factor
            else:
                return self.base_lr
        return self.base_lr


def split_and_load(data, ctx_list, batch_axis=0, even_split=True):
    """`gluon.utils.split_and_load` (reference train_simple_r3d.py:110-111): slice a batch along `batch_axis` into one
    piece per device.  With one process per GPU `ctx_list` normally has one entry."""
    n = len(ctx_list)
    size = data.shape[batch_axis]
    if even_split and size % n != 0:
        raise ValueError("data with shape %s cannot be evenly split into %d slices along axis %d" % (tuple(data.shape), n, batch_axis))
    step = size // n
    out = []
    for i, ctx in enumerate(ctx_list):
        lo = i * step
        hi = size if i == n - 1 else (i + 1) * step
        piece = data.narrow(batch_axis, lo, hi - lo)
        out.append(piece.to(ctx) if ctx is not None else piece)
    return out
