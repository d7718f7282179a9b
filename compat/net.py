"""`from net import create_r3d` (reference net.py:110-170, train.py:10,35)."""
from fastvideotagging_b200.net import BLOCK_CONFIG, ModelBuilder, create_r3d      # noqa: F401
