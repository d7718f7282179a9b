"""`mx.lr_scheduler` — FactorScheduler (train.py:75-77), MultiFactorScheduler (train_simple_r3d.py:99-100)."""
from fastvideotagging_b200.trainer import FactorScheduler, MultiFactorScheduler  # noqa: F401
