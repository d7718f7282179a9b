"""A minimal stand-in for the `mxnet` namespace, backed by torch-CPU, used ONLY by tests/golden/make_golden.py to
execute the reference's own model/R2Plus1.py and model/mlc_loss.py source (unmodified, read from /root/reference)
and record golden outputs.  MXNet itself (`mxnet-cu90`, requirements.txt:6) is not installable in this image.

Semantics implemented are the MXNet 1.x ones the reference relies on: float32 default dtype, comparison ops
returning 0/1 float arrays, Conv3D/Dense deferred shape inference, BatchNorm(axis=1) with biased batch variance in
training and moving statistics otherwise, AvgPool3D 'valid', gluon child/parameter ordering and naming.
Nothing in the product path imports this package.
"""
import contextlib
import sys
import types
from collections import OrderedDict

import numpy as np
import torch

_recording = [False]


def _t(x):
    if isinstance(x, NDArray):
        return x._t
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x, dtype=np.float32))


class Context:
    def __init__(self, kind="cpu", idx=0):
        self.kind, self.idx = kind, idx

    def __repr__(self):
        return "%s(%d)" % (self.kind, self.idx)


def cpu(i=0):
    return Context("cpu", i)


def gpu(i=0):
    return Context("cpu", i)


class NDArray:
    def __init__(self, t):
        self._t = t
        self._grad_holder = None

    # -- basic protocol
    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def context(self):
        return cpu()

    @property
    def grad(self):
        return NDArray(self._t.grad) if self._t.grad is not None else None

    def attach_grad(self):
        self._t = self._t.detach().clone().requires_grad_(True)

    def detach(self):
        return NDArray(self._t.detach())

    def backward(self, out_grad=None):
        self._t.backward(_t(out_grad) if out_grad is not None else torch.ones_like(self._t))

    def asnumpy(self):
        return self._t.detach().cpu().numpy()

    def asscalar(self):
        return self._t.detach().reshape(-1)[0].item() if self._t.numel() == 1 else (_ for _ in ()).throw(ValueError("not a scalar"))

    def astype(self, dt):
        return NDArray(self._t.to(torch.float32))

    def as_in_context(self, ctx):
        return self

    def wait_to_read(self):
        return None

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return NDArray(self._t.reshape(*shape))

    def mean(self):
        return NDArray(self._t.mean())

    def sum(self, axis=None, keepdims=False):
        return _sum(self, axis=axis, keepdims=keepdims)

    def argmax(self, axis):
        return NDArray(self._t.argmax(dim=axis).float())

    def __len__(self):
        return self._t.shape[0]

    def __iter__(self):
        for i in range(self._t.shape[0]):
            yield NDArray(self._t[i])

    def __bool__(self):
        if self._t.numel() != 1:
            raise ValueError("The truth value of an NDArray with multiple elements is ambiguous.")
        return bool(self._t.reshape(-1)[0].item() != 0)

    def __float__(self):
        return float(self._t.reshape(-1)[0].item())

    def __getitem__(self, idx):
        return NDArray(self._t[idx])

    def __setitem__(self, idx, v):
        with torch.no_grad():
            self._t[idx] = _t(v) if isinstance(v, (NDArray, torch.Tensor)) else v

    # -- arithmetic (broadcasting like numpy; MXNet requires explicit broadcast_* for mismatched shapes,
    #    which the reference uses where needed)
    def _bin(self, o, f):
        return NDArray(f(self._t, _t(o) if isinstance(o, (NDArray, torch.Tensor, np.ndarray, list)) else o))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    __radd__ = __add__
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    __rmul__ = __mul__
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: b / a)
    def __neg__(self): return NDArray(-self._t)

    def __iadd__(self, o):
        with torch.no_grad():
            self._t += _t(o) if isinstance(o, (NDArray, torch.Tensor)) else o
        return self

    def _cmp(self, o, f):
        return NDArray(f(self._t, _t(o) if isinstance(o, (NDArray, torch.Tensor)) else o).to(torch.float32))

    def __eq__(self, o): return self._cmp(o, lambda a, b: a == b)
    def __ne__(self, o): return self._cmp(o, lambda a, b: a != b)
    def __lt__(self, o): return self._cmp(o, lambda a, b: a < b)
    def __le__(self, o): return self._cmp(o, lambda a, b: a <= b)
    def __gt__(self, o): return self._cmp(o, lambda a, b: a > b)
    def __ge__(self, o): return self._cmp(o, lambda a, b: a >= b)
    __hash__ = None

    def __repr__(self):
        return "\n%s\n<NDArray %s @cpu(0)>" % (self.asnumpy(), "x".join(map(str, self.shape)))


def _sum(x, axis=None, keepdims=False):
    t = _t(x)
    return NDArray(t.sum().reshape(1)) if axis is None else NDArray(t.sum(dim=axis, keepdim=keepdims))


# ------------------------------------------------------------------------------------------------ mxnet.nd
nd = types.ModuleType("mxnet.nd")
nd.NDArray = NDArray
nd.array = lambda src, ctx=None, dtype=None: NDArray(torch.as_tensor(np.asarray(src.asnumpy() if isinstance(src, NDArray) else src, dtype=np.float32)))
nd.zeros = lambda shape, ctx=None, dtype=None: NDArray(torch.zeros(shape if isinstance(shape, (tuple, list)) else (shape,)))
nd.ones = lambda shape, ctx=None, dtype=None: NDArray(torch.ones(shape if isinstance(shape, (tuple, list)) else (shape,)))
nd.zeros_like = lambda x: NDArray(torch.zeros_like(_t(x)))
nd.ones_like = lambda x: NDArray(torch.ones_like(_t(x)))
nd.greater = lambda a, b: a > b
nd.lesser_equal = lambda a, b: a <= b
nd.equal = lambda a, b: a == b
nd.broadcast_minus = lambda a, b: a - b
nd.broadcast_mul = lambda a, b: a * b
nd.exp = lambda x: NDArray(torch.exp(_t(x)))
nd.log = lambda x: NDArray(torch.log(_t(x)))
nd.relu = lambda x: NDArray(torch.relu(_t(x)))
nd.sum = _sum
nd.mean = lambda x: NDArray(_t(x).mean().reshape(1))
nd.one_hot = lambda idx, depth: NDArray(torch.nn.functional.one_hot(_t(idx).long(), depth).float())
nd.random = types.SimpleNamespace(
    normal=lambda loc=0, scale=1, shape=None, ctx=None: NDArray(torch.randn(shape) * scale + loc),
    uniform=lambda low=0, high=1, shape=None, ctx=None: NDArray(torch.rand(shape) * (high - low) + low))


def _nd_load(fname):
    z = np.load(fname)
    return {k: nd.array(z[k]) for k in z.files}


nd.load = _nd_load
ndarray = nd

# ------------------------------------------------------------------------------------------------ mxnet.autograd
autograd = types.ModuleType("mxnet.autograd")


@contextlib.contextmanager
def _record(train_mode=True):
    _recording.append(True)
    try:
        yield
    finally:
        _recording.pop()


autograd.record = _record
autograd.is_recording = lambda: _recording[-1]


class Function:
    """mx.autograd.Function: forward runs un-recorded on NDArrays; backward supplies input gradients."""

    def __init__(self):
        self.saved_tensors = ()

    def save_for_backward(self, *args):
        self.saved_tensors = args

    def __call__(self, *inputs):
        outer = self

        class _Bridge(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *tin):
                with torch.no_grad():
                    out = outer.forward(*[NDArray(t) for t in tin])
                return _t(out).clone()

            @staticmethod
            def backward(ctx, g):
                grads = outer.backward(NDArray(g))
                if not isinstance(grads, (tuple, list)):
                    grads = (grads,)
                res = []
                for t, gr in zip([_t(i) for i in inputs], grads):
                    gt = _t(gr)
                    res.append(gt if gt.shape == t.shape else None)    # the reference returns a dummy for `target`
                return tuple(res)

        return NDArray(_Bridge.apply(*[_t(i) for i in inputs]))


autograd.Function = Function

# ------------------------------------------------------------------------------------------------ mxnet.gluon
gluon = types.ModuleType("mxnet.gluon")
nn = types.ModuleType("mxnet.gluon.nn")
gloss = types.ModuleType("mxnet.gluon.loss")
_name_counters = {}


class Parameter:
    def __init__(self, name, shape=None):
        self.name, self.shape, self._data = name, shape, None

    def _load_init(self, data, ctx=None):
        self._data = _t(data).detach().clone().to(torch.float32)

    def set_data(self, data):
        self._load_init(data)

    def data(self, ctx=None):
        return NDArray(self._data)


class _NameScope:
    def __init__(self, block):
        self.block = block

    def __enter__(self):
        Block._scope.append(self.block)
        return self

    def __exit__(self, *a):
        Block._scope.pop()


class Block:
    _scope = []

    def __init__(self, prefix=None, params=None):
        hint = type(self).__name__.lower()
        parent = Block._scope[-1] if Block._scope else None
        counters = parent._counters if parent is not None else _name_counters
        if prefix is None:
            i = counters.get(hint, 0)
            counters[hint] = i + 1
            prefix = "%s%d_" % (hint, i)
        self.__dict__["_children"] = OrderedDict()
        self.__dict__["_counters"] = {}
        self.__dict__["_own_params"] = OrderedDict()
        self.prefix = (parent.prefix if parent is not None else "") + prefix
        self._local_prefix = prefix

    @property
    def name(self):
        return self._local_prefix[:-1] if self._local_prefix.endswith("_") else self._local_prefix

    def name_scope(self):
        return _NameScope(self)

    def __setattr__(self, k, v):
        if isinstance(v, Block):
            self._children[k] = v
        object.__setattr__(self, k, v)

    def register_child(self, b):
        self._children[str(len(self._children))] = b

    def collect_params(self):
        out = OrderedDict()
        for k, p in self._own_params.items():
            out[self._local_prefix + k] = p
        for ck, c in self._children.items():
            for k, p in c.collect_params().items():
                out[ck + "/" + k] = p
        return out

    def initialize(self, init=None, ctx=None, **kw):
        return None

    def __call__(self, *a):
        return self.forward(*a)


class HybridBlock(Block):
    def forward(self, *a):
        return self.hybrid_forward(nd, *a)


class Sequential(Block):
    def add(self, *blocks):
        with self.name_scope():
            pass
        for b in blocks:
            self.register_child(b)

    def forward(self, x):
        for b in self._children.values():
            x = b(x)
        return x

    def __iter__(self):
        return iter(self._children.values())


def _scoped(cls):
    """Layers created inside `with blk.name_scope()` or as constructor args pick up the enclosing block's counters."""
    return cls


class Conv3D(Block):
    def __init__(self, channels, kernel_size, strides=(1, 1, 1), padding=(0, 0, 0), use_bias=True, **kw):
        super().__init__()
        self.channels, self.k = channels, tuple(kernel_size)
        self.s, self.p, self.use_bias = tuple(strides), tuple(padding), use_bias
        self._own_params["weight"] = Parameter("weight")
        if use_bias:
            self._own_params["bias"] = Parameter("bias")

    def forward(self, x):
        w = self._own_params["weight"]._data
        b = self._own_params["bias"]._data if self.use_bias else None
        return NDArray(torch.nn.functional.conv3d(_t(x), w, b, stride=self.s, padding=self.p))


class BatchNorm(Block):
    def __init__(self, axis=1, momentum=0.9, epsilon=1e-5, **kw):
        super().__init__()
        self.momentum, self.eps = momentum, epsilon
        for k in ("gamma", "beta", "running_mean", "running_var"):
            self._own_params[k] = Parameter(k)

    def forward(self, x):
        t = _t(x)
        P = self._own_params
        g, b = P["gamma"]._data, P["beta"]._data
        sh = (1, -1, 1, 1, 1)
        if autograd.is_recording():
            mean = t.mean(dim=(0, 2, 3, 4))
            var = t.var(dim=(0, 2, 3, 4), unbiased=False)
            with torch.no_grad():
                P["running_mean"]._data = self.momentum * P["running_mean"]._data + (1 - self.momentum) * mean
                P["running_var"]._data = self.momentum * P["running_var"]._data + (1 - self.momentum) * var
        else:
            mean, var = P["running_mean"]._data, P["running_var"]._data
        scale = g / torch.sqrt(var + self.eps)
        return NDArray(t * scale.reshape(sh) + (b - mean * scale).reshape(sh))


class Activation(Block):
    def __init__(self, activation, **kw):
        super().__init__()
        assert activation == "relu"

    def forward(self, x):
        return NDArray(torch.relu(_t(x)))


class AvgPool3D(Block):
    def __init__(self, pool_size, strides=None, padding=0, **kw):
        Block.__init__(self, prefix=None)
        self._local_prefix = "pool%d_" % 0
        self.k, self.s, self.p = tuple(pool_size), tuple(strides), padding

    def forward(self, x):
        return NDArray(torch.nn.functional.avg_pool3d(_t(x), self.k, stride=self.s))


class Dense(Block):
    def __init__(self, units, activation=None, use_bias=True, **kw):
        super().__init__()
        self.units = units
        self._own_params["weight"] = Parameter("weight")
        self._own_params["bias"] = Parameter("bias")

    def forward(self, x):
        t = _t(x)
        t = t.reshape(t.shape[0], -1)
        return NDArray(t @ self._own_params["weight"]._data.t() + self._own_params["bias"]._data)


for _c in (Block, HybridBlock, Sequential, Conv3D, BatchNorm, Activation, AvgPool3D, Dense):
    setattr(nn, _c.__name__, _c)
gluon.nn = nn
gluon.loss = gloss
gluon.Block = Block
gluon.HybridBlock = HybridBlock

sym = types.ModuleType("mxnet.sym")
optimizer = types.ModuleType("mxnet.optimizer")
init = types.ModuleType("mxnet.init")
init.Xavier = lambda *a, **k: None

for _name, _mod in (("mxnet.nd", nd), ("mxnet.ndarray", nd), ("mxnet.autograd", autograd), ("mxnet.gluon", gluon),
                    ("mxnet.gluon.nn", nn), ("mxnet.gluon.loss", gloss), ("mxnet.sym", sym),
                    ("mxnet.optimizer", optimizer), ("mxnet.init", init)):
    sys.modules[_name] = _mod
