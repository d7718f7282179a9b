"""`mxnet.nd` — the NDArray surface the reference scripts use (.shape, .asscalar(), .asnumpy(), .argmax(axis),
.argsort(), .astype, .mean(), .backward(), slicing, nd.array, nd.mean ...), as a thin wrapper over a torch tensor."""
import numpy as np
import torch

from .context import Context, cpu, current_context

_DTYPES = {"float32": torch.float32, "float64": torch.float64, "float16": torch.float16, "int32": torch.int32,
           "int64": torch.int64, "uint8": torch.uint8, np.float32: torch.float32, np.float64: torch.float64,
           np.int32: torch.int32, np.int64: torch.int64, np.uint8: torch.uint8}


def _unwrap(v):
    return v._t if isinstance(v, NDArray) else v


class NDArray:
    __slots__ = ("_t",)
    __array_priority__ = 100

    def __init__(self, t):
        self._t = t

    # ---- introspection
    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def dtype(self):
        return {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64,
                torch.uint8: np.uint8, torch.bool: np.bool_, torch.float16: np.float16}.get(self._t.dtype, np.float32)

    @property
    def context(self):
        return Context("gpu", self._t.device.index or 0) if self._t.is_cuda else cpu()

    ctx = context

    @property
    def size(self):
        return self._t.numel()

    @property
    def ndim(self):
        return self._t.dim()

    @property
    def grad(self):
        return NDArray(self._t.grad) if self._t.grad is not None else None

    def __len__(self):
        return self._t.shape[0]

    def __repr__(self):
        return "\n%s\n<NDArray %s @%s>" % (self._t.detach().cpu().numpy(), "x".join(map(str, self.shape)), self.context)

    # ---- conversion
    def asnumpy(self):
        return self._t.detach().cpu().numpy()

    def asscalar(self):
        if self._t.numel() != 1:
            raise ValueError("The current array is not a scalar")
        return self._t.detach().reshape(()).cpu().item()

    def astype(self, dtype, copy=True):
        return NDArray(self._t.to(_DTYPES.get(dtype, dtype)))

    def as_in_context(self, ctx):
        return NDArray(self._t.to(ctx.device))

    def copyto(self, other):
        if isinstance(other, Context):
            return NDArray(self._t.to(other.device).clone())
        other._t.copy_(self._t)
        return other

    def copy(self):
        return NDArray(self._t.clone())

    def detach(self):
        return NDArray(self._t.detach())

    def attach_grad(self, grad_req="write"):
        self._t.requires_grad_(True)

    def wait_to_read(self):
        if self._t.is_cuda:
            torch.cuda.current_stream(self._t.device).synchronize()

    # ---- autograd
    def backward(self, out_grad=None, retain_graph=False, train_mode=True):
        head = _unwrap(out_grad) if out_grad is not None else torch.ones_like(self._t)      # MXNet: head gradient of ones
        self._t.backward(head, retain_graph=retain_graph)

    # ---- reductions / indexing
    def mean(self, axis=None, keepdims=False):
        t = self._t.float() if not self._t.is_floating_point() else self._t
        return NDArray(t.mean() if axis is None else t.mean(dim=axis, keepdim=keepdims))

    def sum(self, axis=None, keepdims=False):
        return NDArray(self._t.sum() if axis is None else self._t.sum(dim=axis, keepdim=keepdims))

    def max(self, axis=None):
        return NDArray(self._t.max() if axis is None else self._t.max(dim=axis).values)

    def min(self, axis=None):
        return NDArray(self._t.min() if axis is None else self._t.min(dim=axis).values)

    def argmax(self, axis=None):
        return NDArray(self._t.argmax(dim=axis).float())                 # MXNet returns float32 indices

    def argsort(self, axis=-1, is_ascend=True):
        # stable ascending sort (the reference reverses it: ties come out larger-index first, train_simple_r3d.py:183)
        return NDArray(torch.sort(self._t, dim=axis, descending=not is_ascend, stable=True).indices.float())

    def reshape(self, *shape):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape
        return NDArray(self._t.reshape(*[int(s) for s in shape]))

    def transpose(self, *axes):
        axes = axes[0] if len(axes) == 1 and isinstance(axes[0], (tuple, list)) else axes
        return NDArray(self._t.permute(*axes) if axes else self._t.t())

    def __getitem__(self, key):
        keys = key if isinstance(key, tuple) else (key,)
        t = self._t
        flips, norm = [], []
        for d, k in enumerate(keys):
            if isinstance(k, slice) and k.step is not None and k.step < 0:      # torch has no negative steps: slice forwards, then flip
                n = t.shape[d]
                idx = list(range(n))[k]
                if idx:
                    lo, hi, st = idx[-1], idx[0] + 1, -k.step
                    norm.append(slice(lo, hi, st))
                else:
                    norm.append(slice(0, 0))
                flips.append(d)
            else:
                norm.append(_unwrap(k).long() if isinstance(k, NDArray) else k)
        out = t[tuple(norm)]
        if flips:
            removed = [d for d, k in enumerate(norm) if isinstance(k, int)]
            dims = [d - len([r for r in removed if r < d]) for d in flips]
            out = torch.flip(out, dims)
        return NDArray(out)

    def __setitem__(self, key, value):
        self._t[key] = _unwrap(value)

    def __iter__(self):
        for i in range(self._t.shape[0]):
            yield NDArray(self._t[i])

    # ---- arithmetic
    def _bin(self, other, fn, reverse=False):
        o = _unwrap(other)
        if isinstance(o, torch.Tensor) and o.device != self._t.device:
            o = o.to(self._t.device)
        return NDArray(fn(o, self._t) if reverse else fn(self._t, o))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, lambda a, b: a - b, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.true_divide)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: a / b, True)
    def __neg__(self): return NDArray(-self._t)
    def __iadd__(self, o):
        self._t = self._t + _unwrap(o)
        return self
    def __isub__(self, o):
        self._t = self._t - _unwrap(o)
        return self
    def __imul__(self, o):
        self._t = self._t * _unwrap(o)
        return self

    def _cmp(self, other, fn):
        o = _unwrap(other)
        if isinstance(o, torch.Tensor):
            o = o.to(self._t.device, self._t.dtype) if o.dtype != self._t.dtype or o.device != self._t.device else o
        return NDArray(fn(self._t, o).to(self._t.dtype if self._t.is_floating_point() else torch.float32))

    def __eq__(self, o): return self._cmp(o, torch.eq)
    def __ne__(self, o): return self._cmp(o, torch.ne)
    def __gt__(self, o): return self._cmp(o, torch.gt)
    def __ge__(self, o): return self._cmp(o, torch.ge)
    def __lt__(self, o): return self._cmp(o, torch.lt)
    def __le__(self, o): return self._cmp(o, torch.le)
    __hash__ = None

    def __float__(self):
        return float(self.asscalar())

    def __bool__(self):
        return bool(self.asscalar())


def array(source, ctx=None, dtype=None):
    if isinstance(source, NDArray):
        t = source._t.clone()
    else:
        a = np.asarray(source)
        if dtype is None and a.dtype != np.uint8:
            a = a.astype(np.float32)                      # mx.nd.array defaults to float32
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(_DTYPES.get(dtype, dtype))
    ctx = ctx or current_context()
    return NDArray(t.to(ctx.device))


def zeros(shape, ctx=None, dtype="float32"):
    return NDArray(torch.zeros(shape, dtype=_DTYPES.get(dtype, dtype), device=(ctx or current_context()).device))


def ones(shape, ctx=None, dtype="float32"):
    return NDArray(torch.ones(shape, dtype=_DTYPES.get(dtype, dtype), device=(ctx or current_context()).device))


def mean(x, axis=None, keepdims=False):
    return x.mean(axis, keepdims)


def sum(x, axis=None, keepdims=False):      # noqa: A001  (mxnet.nd.sum)
    return x.sum(axis, keepdims)


def relu(x):
    return NDArray(torch.relu(x._t))


def sigmoid(x):
    return NDArray(torch.sigmoid(x._t))


def softmax(x, axis=-1):
    return NDArray(torch.softmax(x._t, dim=axis))


def concat(*arrays, dim=1):
    return NDArray(torch.cat([a._t for a in arrays], dim=dim))


def argmax(x, axis=None):
    return x.argmax(axis)


def save(fname, data):
    from fastvideotagging_b200 import params_io
    params_io.nd_save(fname, {k: v.asnumpy() for k, v in data.items()})


def load(fname):
    from fastvideotagging_b200 import params_io
    return {k: array(v, ctx=cpu()) for k, v in params_io.nd_load(fname).items()}


class _Random:
    @staticmethod
    def uniform(low=0, high=1, shape=(1,), ctx=None, dtype="float32"):
        return NDArray(torch.empty(shape, device=(ctx or current_context()).device).uniform_(low, high))

    @staticmethod
    def normal(loc=0, scale=1, shape=(1,), ctx=None, dtype="float32"):
        return NDArray(torch.empty(shape, device=(ctx or current_context()).device).normal_(loc, scale))


random = _Random()
