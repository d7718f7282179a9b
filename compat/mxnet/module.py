"""`mx.module.Module` over an `R3DSymbol` (train.py:54,79-94; validation.py:23-50): bind / set_params / forward /
get_outputs / prepare / fit, one executor on ONE GPU per process (the reference binds one executor per context in one
process; here each GPU is a process and `fit` sums gradients with NCCL through fastvideotagging_b200.trainer.Trainer)."""
import logging
import time
from collections import namedtuple

import torch

from fastvideotagging_b200.trainer import Trainer

from . import metric as _metric
from . import ndarray as nd
from .context import one_device

BatchEndParam = namedtuple("BatchEndParams", ["epoch", "nbatch", "eval_metric", "locals"])


class Module:
    def __init__(self, symbol, data_names=("data",), label_names=("softmax_label",), logger=logging, context=None,
                 work_load_list=None, fixed_param_names=None, state_names=None):
        self.symbol = symbol
        self._device = one_device(context)
        self._exec = None
        self._for_training = False
        self.binded = self.params_initialized = self.optimizer_initialized = False

    # ---- binding / parameters
    def bind(self, data_shapes, label_shapes=None, for_training=True, inputs_need_grad=False, force_rebind=False,
             shared_module=None, grad_req="write"):
        if self.binded and not force_rebind:
            return
        torch.cuda.set_device(self._device)
        self._exec = self.symbol.bind(ctx=self._device)
        self._for_training = for_training
        self.binded = True

    @property
    def _net(self):
        return self._exec.net

    def init_params(self, initializer=None, arg_params=None, aux_params=None, allow_missing=False, force_init=False, allow_extra=False):
        if self.params_initialized and not force_init:
            return
        ft, mag = (initializer.factor_type, initializer.magnitude) if initializer is not None and hasattr(initializer, "factor_type") else ("avg", 3.0)
        self._net.initialize(ctx=self._device, factor_type=ft, magnitude=mag)
        self.set_params(arg_params or {}, aux_params or {}, allow_missing=True)
        self.params_initialized = True

    def set_params(self, arg_params, aux_params, allow_missing=False, force_init=True, allow_extra=False):
        merged = {}
        for d in (arg_params or {}, aux_params or {}):
            for k, v in d.items():
                merged[k] = v.asnumpy() if hasattr(v, "asnumpy") else v
        own = set(self._net._param_names + self._net._aux_names)
        merged = {k: v for k, v in merged.items() if k in own}
        if merged:
            self._net.load_param_dict(merged, with_dense=True, strict=not allow_missing)
        self.params_initialized = True

    def get_params(self):
        net = self._net
        arg = {k: nd.NDArray(getattr(net, k).detach().clone()) for k in net._param_names}
        aux = {k: nd.NDArray(getattr(net, k).detach().clone()) for k in net._aux_names}
        return arg, aux

    def init_optimizer(self, kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.01),), force_init=False):
        op = dict(optimizer_params)
        self._lr_scheduler = op.pop("lr_scheduler", None)
        # Module API: weight decay only on *_weight and *_gamma
        self._trainer = Trainer(self._net, optimizer, op, kvstore=getattr(kvstore, "type", kvstore), wd_policy="module")
        if self._lr_scheduler is not None:
            self._lr_scheduler.base_lr = op.get("learning_rate", 0.01)
        self._num_update = 0
        self.optimizer_initialized = True

    # ---- computation
    def _to_dev(self, arr):
        t = arr._t if isinstance(arr, nd.NDArray) else torch.as_tensor(arr)
        return t.to(self._device, non_blocking=True)

    def forward(self, data_batch, is_train=None):
        is_train = self._for_training if is_train is None else is_train
        x = self._to_dev(data_batch.data[0]).float()
        y = self._to_dev(data_batch.label[0]).float() if (is_train and data_batch.label) else None
        self._exec.forward(is_train=is_train, data=x, softmax_label=y)
        self._batch_size = x.shape[0]

    def backward(self, out_grads=None):
        self._exec.backward()

    def update(self):
        self._num_update += 1
        if self._lr_scheduler is not None:
            self._trainer.set_learning_rate(self._lr_scheduler(self._num_update))
        import torch.distributed as dist
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._trainer.step(self._batch_size * world)            # rescale_grad = 1 / total batch (Module.fit's default)

    def get_outputs(self, merge_multi_context=True):
        return [nd.NDArray(o) for o in self._exec.outputs]

    def prepare(self, data_batch, sparse_row_id_fn=None):
        pass

    def update_metric(self, eval_metric, labels, pre_sliced=False):
        eval_metric.update(labels, self.get_outputs())

    def score(self, eval_data, eval_metric, num_batch=None, reset=True, epoch=0):
        eval_metric = _metric.create(eval_metric)
        eval_metric.reset()
        if reset:
            eval_data.reset()
        for nbatch, batch in enumerate(eval_data):
            if num_batch is not None and nbatch == num_batch:
                break
            self.forward(batch, is_train=False)
            self.update_metric(eval_metric, batch.label)
        return [eval_metric.get()]

    def save_checkpoint(self, prefix, epoch, save_optimizer_states=False):
        from . import model
        arg, aux = self.get_params()
        model.save_checkpoint(prefix, epoch, self.symbol, arg, aux)

    def fit(self, train_data, eval_data=None, eval_metric="acc", epoch_end_callback=None, batch_end_callback=None,
            kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.01),), eval_end_callback=None,
            eval_batch_end_callback=None, initializer=None, arg_params=None, aux_params=None, allow_missing=False,
            force_rebind=False, force_init=False, begin_epoch=0, num_epoch=None, validation_metric=None, monitor=None):
        """mx.module.BaseModule.fit as train.py:79-94 calls it."""
        assert num_epoch is not None, "please specify number of epochs"
        self.bind(data_shapes=train_data.provide_data, label_shapes=train_data.provide_label, for_training=True, force_rebind=force_rebind)
        self.init_params(initializer=initializer, arg_params=arg_params, aux_params=aux_params, allow_missing=allow_missing, force_init=force_init)
        self.init_optimizer(kvstore=kvstore, optimizer=optimizer, optimizer_params=optimizer_params)
        eval_metric = _metric.create(eval_metric)
        validation_metric = validation_metric or eval_metric
        for epoch in range(begin_epoch, num_epoch):
            tic = time.time()
            eval_metric.reset()
            train_data.reset()
            for nbatch, batch in enumerate(train_data):
                self.forward(batch, is_train=True)
                self.backward()
                self.update()
                self.update_metric(eval_metric, batch.label)
                if batch_end_callback is not None:
                    p = BatchEndParam(epoch=epoch, nbatch=nbatch, eval_metric=eval_metric, locals=locals())
                    for cb in (batch_end_callback if isinstance(batch_end_callback, (list, tuple)) else [batch_end_callback]):
                        cb(p)
            name, val = eval_metric.get()
            logging.info("Epoch[%d] Train-%s=%f", epoch, name, val)
            logging.info("Epoch[%d] Time cost=%.3f", epoch, time.time() - tic)
            arg, aux = self.get_params()
            if epoch_end_callback is not None:
                for cb in (epoch_end_callback if isinstance(epoch_end_callback, (list, tuple)) else [epoch_end_callback]):
                    cb(epoch, self.symbol, arg, aux)
            if eval_data is not None:
                for name, val in self.score(eval_data, validation_metric, epoch=epoch):
                    logging.info("Epoch[%d] Validation-%s=%f", epoch, name, val)
