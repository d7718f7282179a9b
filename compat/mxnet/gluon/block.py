"""gluon.Block protocol on top of fastvideotagging_b200's modules: __call__ on NDArrays, autograd.record() selects
training mode (batch-statistics BatchNorm + gradients) exactly like MXNet, initialize / collect_params / save / load."""
import torch

from .. import autograd
from .. import ndarray as nd
from ..context import one_device


class ParameterDict(dict):
    """What collect_params() returns: {name: tensor}, remembering the network so that gluon.Trainer can find it."""
    net = None


class Block:
    """Wraps a torch module `self._impl` whose forward takes/returns torch tensors."""
    _impl = None

    def __call__(self, *args):
        ins = [a._t if isinstance(a, nd.NDArray) else a for a in args]
        if autograd.is_recording():
            self._impl.train(autograd.is_training())
            with torch.enable_grad():
                out = self._impl(*ins)
        else:
            self._impl.eval()
            with torch.no_grad():
                out = self._impl(*ins)
        if isinstance(out, (tuple, list)):
            return tuple(nd.NDArray(o) for o in out)
        return nd.NDArray(out)

    def forward(self, *args):
        return self(*args)

    def hybridize(self, active=True, **kwargs):
        pass                                # the execution plans are already static (CUDA graphs)

    def initialize(self, init=None, ctx=None, verbose=False, force_reinit=False):
        dev = one_device(ctx) if ctx is not None else torch.device("cuda", torch.cuda.current_device())
        torch.cuda.set_device(dev)
        ft, mag = (init.factor_type, init.magnitude) if init is not None and hasattr(init, "factor_type") else ("avg", 3.0)
        if hasattr(self._impl, "initialize"):
            self._impl.initialize(ctx=dev, factor_type=ft, magnitude=mag)
        else:
            self._impl.to(dev)
        return self

    def collect_params(self, select=None):
        d = ParameterDict(self._impl.collect_params() if hasattr(self._impl, "collect_params") else dict(self._impl.named_parameters()))
        d.net = self._impl
        return d

    def save_parameters(self, filename):
        self._impl.save_parameters(filename)

    save_params = save_parameters

    def load_parameters(self, filename, ctx=None, allow_missing=False, ignore_extra=False):
        self._impl.load_parameters(filename, ctx=one_device(ctx) if ctx is not None else None, allow_missing=allow_missing)

    load_params = load_parameters

    def __getattr__(self, name):            # anything else (load_from_sym_params, extract_features, name lists ...) -> the module
        impl = self.__dict__.get("_impl")
        if impl is None:
            raise AttributeError(name)
        return getattr(impl, name)


HybridBlock = Block
