"""`gluon.utils.split_and_load` (train_simple_r3d.py:110-111): one slice per context — one context per process here."""
from .. import ndarray as nd
from ..context import one_device


def split_and_load(data, ctx_list, batch_axis=0, even_split=True):
    dev = one_device(ctx_list)
    t = data._t if isinstance(data, nd.NDArray) else data
    return [nd.NDArray(t.to(dev, non_blocking=True))]


def split_data(data, num_slice, batch_axis=0, even_split=True):
    if num_slice != 1:
        raise NotImplementedError("one process per GPU: num_slice must be 1")
    return [data]
