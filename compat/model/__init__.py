"""`from model import R2Plus2D, LsepLoss, LSEP_funcLoss, WarpLoss, WARP_funcLoss, Decision_thresh` (reference
model/__init__.py:1-5, train_simple_r3d.py:6,24-25) as gluon Blocks over fastvideotagging_b200.model.  (The reference's
own `model` package cannot be imported on any interpreter: model/unified_model.py:25 is a SyntaxError.)"""
from fastvideotagging_b200 import model as _m
from fastvideotagging_b200.model import BLOCK_CONFIG      # noqa: F401
from mxnet.gluon.block import Block as _Block
from mxnet.gluon.loss import _LossBlock


class R2Plus2D(_Block):
    """model/R2Plus1.py:93-254 — same constructor arguments."""

    def __init__(self, num_class, model_depth, final_spatial_kernel=7, final_temporal_kernel=2, with_bias=False):
        self._impl = _m.R2Plus2D(num_class, model_depth, final_spatial_kernel, final_temporal_kernel, with_bias)

    def extract_features(self, x):
        from mxnet import ndarray as nd
        return nd.NDArray(self._impl.extract_features(x._t if isinstance(x, nd.NDArray) else x))

    def load_from_sym_params(self, f, ctx=None, with_dense=False):
        return self._impl.load_from_sym_params(f, None, with_dense)


class R2Plus2D_MT(_Block):
    """model/multi_taskR3d.py:93-267."""

    def __init__(self, num_scenes, num_actions, model_depth, final_spatial_kernel=7, final_temporal_kernel=2, with_bias=False, **kw):
        self._impl = _m.R2Plus2D_MT(num_scenes, num_actions, model_depth, final_spatial_kernel, final_temporal_kernel, with_bias, **kw)


class Decision_thresh(_Block):
    """model/decision_model.py:4-14."""

    def __init__(self, thresh_size=63):
        self._impl = _m.Decision_thresh(thresh_size)


class LsepLoss(_LossBlock):
    def __init__(self):
        super().__init__(_m.LsepLoss())


class LsepLossHy(_LossBlock):
    def __init__(self, batch_size=4, num_class=63):
        super().__init__(_m.LsepLossHy(batch_size, num_class))


class LSEP_funcLoss(_LossBlock):
    def __init__(self):
        super().__init__(_m.LSEP_funcLoss())


class WarpLoss(_LossBlock):
    def __init__(self, label_size=62):
        super().__init__(_m.WarpLoss(label_size))


class WARP_funcLoss(_LossBlock):
    def __init__(self, label_size=62):
        super().__init__(_m.WARP_funcLoss(label_size))
