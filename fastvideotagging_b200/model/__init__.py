"""Mirror of reference model/__init__.py:1-2 (the importable part of it)."""
from .R2Plus1 import R2Plus2D, BLOCK_CONFIG  # noqa: F401
from .blocks import get_spatial_temporal_conv, R3DBlock, get_R2plus1d  # noqa: F401
from .mlc_loss import LsepLoss, LSEP_funcLoss, WarpLoss, WARP_funcLoss, LsepLossHy  # noqa: F401
from .mlc_loss import SigmoidBinaryCrossEntropyLoss, SoftmaxCrossEntropyLoss, SoftmaxOutput  # noqa: F401
from .heads import R2Plus2D_MT, Decision_thresh, ECOLite3DHead  # noqa: F401  (multi_taskR3d.py:93, decision_model.py:4, ECO paper)
