"""`mxnet.autograd` — record()/pause() scopes (train_simple_r3d.py:116).  Inside record() the network runs in training
mode (batch-statistics BatchNorm) with gradients enabled; outside it runs in inference mode."""
import contextlib

import torch

_state = {"recording": False, "training": False}


def is_recording():
    return _state["recording"]


def is_training():
    return _state["training"]


@contextlib.contextmanager
def record(train_mode=True):
    prev = dict(_state)
    _state["recording"], _state["training"] = True, bool(train_mode)
    try:
        with torch.enable_grad():
            yield
    finally:
        _state.update(prev)


@contextlib.contextmanager
def pause(train_mode=False):
    prev = dict(_state)
    _state["recording"], _state["training"] = False, bool(train_mode)
    try:
        with torch.no_grad():
            yield
    finally:
        _state.update(prev)


def train_mode():
    return record(True)


def predict_mode():
    return pause(False)


def backward(heads, head_grads=None):
    heads = heads if isinstance(heads, (list, tuple)) else [heads]
    for h in heads:
        h.backward()
