"""`gluon.Trainer(params, 'sgd', {...}, kvstore=...)` (train_simple_r3d.py:95-97): set_learning_rate / learning_rate / step."""
from fastvideotagging_b200.trainer import Trainer as _Trainer


class Trainer(_Trainer):
    def __init__(self, params, optimizer, optimizer_params=None, kvstore="device", compression_params=None, update_on_kvstore=None):
        net = getattr(params, "net", None)
        if net is None:
            raise TypeError("gluon.Trainer needs the ParameterDict returned by net.collect_params()")
        super().__init__(net, optimizer, optimizer_params, kvstore=kvstore)
