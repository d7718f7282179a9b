"""`mxnet.gluon` — Trainer, utils.split_and_load, loss, nn (train_simple_r3d.py:12-17,95-124)."""
from . import loss, nn, utils, data      # noqa: F401
from .trainer import Trainer             # noqa: F401
from .block import Block, HybridBlock    # noqa: F401
