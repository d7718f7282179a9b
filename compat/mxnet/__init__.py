"""The part of the `mxnet` namespace that the reference's train.py, train_simple_r3d.py and validation.py touch,
mapped onto fastvideotagging_b200: torch tensors for device memory, libfvt_b200.so (hand-written sm_100a CUDA) for every
operator.  MXNet itself does not exist for this platform (`mxnet-cu90`, no sm_100 build) — this module is the drop-in
boundary's "host side", not a re-implementation of MXNet: anything outside that surface raises."""
import numpy as _np
import torch as _torch

from . import ndarray as nd          # noqa: F401
from . import ndarray                # noqa: F401
from . import autograd, callback, context, gluon, init, initializer, io, kvstore, lr_scheduler, metric, model, module, optimizer  # noqa: F401
from . import module as mod          # noqa: F401
from . import kvstore as kv          # noqa: F401
from .context import Context, cpu, gpu, current_context      # noqa: F401

__version__ = "1.3.0-fvt-b200-compat"


class _Random:
    @staticmethod
    def seed(seed_state, ctx="all"):
        """mx.random.seed (train_simple_r3d.py:27)."""
        _torch.manual_seed(int(seed_state))
        _np.random.seed(int(seed_state) & 0xffffffff)


random = _Random()


class _Viz:
    @staticmethod
    def plot_network(*args, **kwargs):
        raise NotImplementedError("mx.viz.plot_network needs graphviz and an MXNet symbol graph (train.py --plot): not provided")


viz = _Viz()
