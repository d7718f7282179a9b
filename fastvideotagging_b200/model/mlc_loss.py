"""Host-side mirror of reference model/mlc_loss.py: same class names, constructor arguments and call convention
(`loss = criterion(pred, target)`; `loss.backward()`), computed by the one-launch forward+backward kernels of
libfvt_b200.so (csrc/loss_kernels.cu).  pred/target are (batch, num_class) fp32 CUDA tensors.

Differences that are deliberate and documented:
  * WARP negative sampling uses the counter-based Philox stream (seed, global sample index, class, trial) instead of
    np.random.choice, so results do not depend on host RNG state or on how the batch is sharded over GPUs.
  * A row without negatives (the reference loops forever, mlc_loss.py:140-142) yields NaN.
"""
import ctypes

import torch

from .. import _lib
from ..ops import _ptr, _stream, require_cuda


def _prep(pred, target):
    require_cuda(pred, "pred")
    require_cuda(target, "target")
    if pred.dim() != 2 or pred.shape != target.shape:
        raise ValueError("pred and target must both be (batch, num_class); got %s and %s" % (tuple(pred.shape), tuple(target.shape)))
    return pred.detach().float().contiguous(), target.detach().float().contiguous()


def _workspace(batch, device):
    return torch.empty(batch + 4, dtype=torch.float32, device=device)


class _LsepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mode):
        lib = _lib.load()
        p, t = _prep(pred, target)
        b, c = p.shape
        loss = torch.empty(1, dtype=torch.float32, device=p.device)
        grad = torch.empty_like(p)
        _lib.check(lib.fvt_lsep_fwd_bwd(_lib.handle(), _ptr(p), _ptr(t), b, c, mode, _ptr(loss), _ptr(grad), _ptr(_workspace(b, p.device)), _stream()))
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g.reshape(1, 1), None, None


class LsepLoss(torch.nn.Module):
    """Log-sum-exp pairwise ranking loss, reference mlc_loss.py:57-86.  Returns a (1,) tensor:
    log(1 + sum over the whole batch of exp(p_neg - p_pos))."""

    def forward(self, pred, target):
        return _LsepFn.apply(pred, target, 0)


class LsepLossHy(LsepLoss):
    """Hybridizable twin, reference mlc_loss.py:89-108 (same value; batch_size/num_class are shape hints there)."""

    def __init__(self, batch_size=4, num_class=63):
        super().__init__()
        self.batch, self.dim = batch_size, num_class


class LSEP_funcLoss(torch.nn.Module):
    """autograd.Function variant exactly as written in the reference (mlc_loss.py:8-54), including the shadowed
    batch index in forward and the -1/loss factor in backward.  Prefer LsepLoss; this exists for drop-in parity."""
    name = "LSEP_funcLoss"

    def forward(self, pred, target, max_num_trials=None):
        return _LsepFn.apply(pred, target, 1)


class _WarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mode, label_size, max_trials, seed, sample_offset, rank_in):
        lib = _lib.load()
        p, t = _prep(pred, target)
        b, c = p.shape
        loss = torch.empty(1, dtype=torch.float32, device=p.device)
        grad = torch.empty_like(p)
        rank = torch.empty_like(p)
        trials = torch.empty((b, c), dtype=torch.int32, device=p.device)
        _lib.check(lib.fvt_warp_fwd_bwd(_lib.handle(), _ptr(p), _ptr(t), b, c, label_size, max_trials, mode,
                                        ctypes.c_uint64(seed), ctypes.c_uint64(sample_offset), _ptr(rank_in),
                                        _ptr(rank), _ptr(trials), _ptr(loss), _ptr(grad),
                                        _ptr(_workspace(b, p.device)), _stream()))
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(rank, trials)
        return loss, rank, trials

    @staticmethod
    def backward(ctx, g, _gr, _gt):
        (grad,) = ctx.saved_tensors
        return (grad * g.reshape(1, 1),) + (None,) * 7


class WarpLoss(torch.nn.Module):
    """WARP loss, reference mlc_loss.py:110-174.  `label_size` sizes the harmonic rank-weight table and
    max_num_trails = label_size - 1 (:117-120).  `seed` keys the Philox stream; `sample_offset` is the global index of
    row 0 of this shard so sampled ranks do not depend on the GPU count; it advances by the batch size per call unless
    `auto_advance=False`."""

    def __init__(self, label_size=62, seed=123, sample_offset=0, auto_advance=True):
        super().__init__()
        self.label_size = label_size
        self.max_num_trails = label_size - 1
        self.seed, self.sample_offset, self.auto_advance = seed, sample_offset, auto_advance
        self.rank_weights = [1.0 / 1]
        for i in range(1, label_size):
            self.rank_weights.append(self.rank_weights[i - 1] + 1.0 / (i + 1))
        self.last_rank = None
        self.last_trials = None

    _mode = 0

    def _max_trials(self, target):
        return self.max_num_trails

    def forward(self, pred, target, rank_weights=None):
        loss, rank, trials = _WarpFn.apply(pred, target, self._mode, self.label_size, self._max_trials(target),
                                           self.seed, self.sample_offset, rank_weights)
        self.last_rank, self.last_trials = rank, trials
        if self.auto_advance:
            self.sample_offset += pred.shape[0]
        return loss


class WARP_funcLoss(WarpLoss):
    """autograd.Function variant, reference mlc_loss.py:177-233: max_num_trials = num_class - 1 (:192) and the
    linear (non-hinge) value sum_b (sum_j L_bj) * sum_c (1 - pos*p + neg*p)."""
    name = "WARP_funcLoss"
    _mode = 1

    def _max_trials(self, target):
        return target.shape[1] - 1


class _BceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, from_sigmoid):
        lib = _lib.load()
        p, t = _prep(pred, target)
        b, c = p.shape
        loss = torch.empty(b, dtype=torch.float32, device=p.device)
        grad = torch.empty_like(p)
        _lib.check(lib.fvt_bce_fwd_bwd(_lib.handle(), _ptr(p), _ptr(t), b, c, int(from_sigmoid), _ptr(loss), _ptr(grad), _stream()))
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g.reshape(-1, 1), None, None


class SigmoidBinaryCrossEntropyLoss(torch.nn.Module):
    """gluon.loss.SigmoidBinaryCrossEntropyLoss as the reference uses it (train_simple_r3d.py:76,237): returns the
    per-sample mean over classes, shape (batch,)."""

    def __init__(self, from_sigmoid=False):
        super().__init__()
        self._from_sigmoid = from_sigmoid

    def forward(self, pred, label):
        return _BceFn.apply(pred, label, self._from_sigmoid)


class _SoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, label, mode):
        lib = _lib.load()
        require_cuda(logits, "logits")
        x = logits.detach().float().contiguous()
        lab = label.detach().float().contiguous().reshape(-1)
        b, c = x.shape
        out = torch.empty((b,) if mode == 0 else (b, c), dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x)
        _lib.check(lib.fvt_softmax_fwd_bwd(_lib.handle(), _ptr(x), _ptr(lab), b, c, mode, _ptr(out), _ptr(grad), _stream()))
        ctx.save_for_backward(grad)
        ctx.mode = mode
        return out

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        if ctx.mode == 0:
            return grad * g.reshape(-1, 1), None, None
        return grad, None, None       # SoftmaxOutput ignores the head gradient (MXNet: out_grad is not used)


class SoftmaxCrossEntropyLoss(torch.nn.Module):
    """gluon.loss.SoftmaxCrossEntropyLoss with sparse labels (train_simple_r3d.py:43): (batch,) losses."""

    def forward(self, pred, label):
        return _SoftmaxFn.apply(pred, label, 0)


def SoftmaxOutput(data, label):
    """mx.sym.SoftmaxOutput(multi_output=True, use_ignore=True, normalization='null') (net.py:167-169): returns the
    class probabilities; its backward is (p - onehot) regardless of the incoming gradient, label -1 is ignored."""
    return _SoftmaxFn.apply(data, label, 1)
