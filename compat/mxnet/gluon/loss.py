"""`mxnet.gluon.loss` — the two gluon losses the reference instantiates (train_simple_r3d.py:43,70-76,237), on the loss kernels."""
from fastvideotagging_b200.model import mlc_loss as _ml

from .block import Block


class _LossBlock(Block):
    def __init__(self, impl):
        self._impl = impl

    def __call__(self, pred, label, *rest):
        from .. import ndarray as nd
        out = self._impl(pred._t if isinstance(pred, nd.NDArray) else pred, label._t if isinstance(label, nd.NDArray) else label)
        return nd.NDArray(out)


class SigmoidBinaryCrossEntropyLoss(_LossBlock):
    def __init__(self, from_sigmoid=False, weight=None, batch_axis=0, **kwargs):
        super().__init__(_ml.SigmoidBinaryCrossEntropyLoss(from_sigmoid=from_sigmoid))


SigmoidBCELoss = SigmoidBinaryCrossEntropyLoss


class SoftmaxCrossEntropyLoss(_LossBlock):
    def __init__(self, axis=-1, sparse_label=True, from_logits=False, weight=None, batch_axis=0, **kwargs):
        if not sparse_label or from_logits:
            raise NotImplementedError("the reference uses the default sparse-label, from-logits=False form")
        super().__init__(_ml.SoftmaxCrossEntropyLoss())


SoftmaxCELoss = SoftmaxCrossEntropyLoss
