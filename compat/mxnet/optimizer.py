"""`mxnet.optimizer` is imported by train_simple_r3d.py:14 and never used; the only optimiser the reference trains with is
'sgd' with momentum (train_simple_r3d.py:95-97, train.py:69-73), implemented by fvt_sgd_momentum_multi."""


class SGD:
    def __init__(self, learning_rate=0.01, momentum=0.0, wd=0.0, lr_scheduler=None, rescale_grad=1.0, **kwargs):
        self.learning_rate, self.momentum, self.wd, self.lr_scheduler, self.rescale_grad = learning_rate, momentum, wd, lr_scheduler, rescale_grad
