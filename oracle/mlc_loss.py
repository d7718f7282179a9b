"""CPU ORACLE (test infrastructure, not product code) for reference model/mlc_loss.py and the gluon losses the
training scripts select (train_simple_r3d.py:43,70-78).  numpy only; every function cites the lines it restates.

Pinned by tests/golden/mlc_loss_golden.json, which holds the outputs of the reference's own model/mlc_loss.py source
executed (unmodified, from /root/reference) on a torch-backed stand-in for the `mxnet` namespace
(tests/golden/make_golden.py) — the closest available thing to running the reference, since MXNet itself cannot be
installed here.
"""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """Philox4x32-10 (Random123).  counter: 4 uint32, key: 2 uint32 -> 4 uint32.  Known answers in
    tests/test_oracle_losses.py (Random123 kat_vectors)."""
    c = [int(x) & MASK for x in counter]
    k = [int(x) & MASK for x in key]
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = (p0 >> 32) & MASK, p0 & MASK
        hi1, lo1 = (p1 >> 32) & MASK, p1 & MASK
        c = [hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return c


def rank_weights(label_size):
    """WarpLoss.__init__ (mlc_loss.py:117-119): rank_weights[k] = H_{k+1}, python-float accumulation."""
    rw = [1.0 / 1]
    for i in range(1, label_size):
        rw.append(rw[i - 1] + 1.0 / (i + 1))
    return rw


def lsep_loss(pred, target, dtype=np.float32):
    """LsepLoss.forward (mlc_loss.py:63-86) and the gradient MXNet autodiff produces for it.
    Returns (loss scalar, grad[B,C])."""
    p = np.asarray(pred, dtype)
    t = np.asarray(target)
    dist = p[:, :, None] - p[:, None, :]                       # dist[b,i,j] = p_i - p_j  (:68)
    pos = (t > 0)[:, :, None].astype(dtype)                    # :69
    neg = (t == 0)[:, None, :].astype(dtype)                   # :70
    e = pos * neg * np.exp(-dist)                              # :85
    S = e.sum(dtype=dtype)
    loss = np.log(dtype(1) + S)
    grad = (e.sum(axis=1) - e.sum(axis=2)) / (dtype(1) + S)    # d/dp_j: +e_ij ; d/dp_i: -e_ij
    return loss, grad.astype(dtype)


def lsep_func_loss(pred, target, dtype=np.float32):
    """LSEP_funcLoss.forward/backward AS WRITTEN (mlc_loss.py:15-54): the inner enumerate shadows the batch index
    (:27-29) and backward uses fac = -1/loss with the one-hot form (:36,51-53).  head gradient = 1."""
    p = np.asarray(pred, dtype)
    t = np.asarray(target)
    B, C = p.shape
    loss = dtype(0)
    for b in range(B):
        pos = [j for j in range(C) if t[b, j] > 0]
        neg = [j for j in range(C) if t[b, j] <= 0]
        for q, pj in enumerate(pos):                           # `for i,pj in enumerate(pos)` — row index is q
            if q >= B:
                continue                                       # IndexError in the reference; skipped here and in the kernel
            for nj in neg:
                loss += np.exp(p[q, nj] - p[q, pj])
    loss = np.log(dtype(1) + loss)
    fac = dtype(-1) / loss
    grad = np.zeros_like(p)
    for b in range(B):
        npos = int((t[b] > 0).sum())
        nneg = int((t[b] <= 0).sum())
        for k in range(C):
            if t[b, k] > 0:
                grad[b, k] += nneg * np.exp(-p[b, k])
            if t[b, k] <= 0:
                grad[b, k] -= npos * np.exp(p[b, k])
    return loss, (grad * fac).astype(dtype)


def warp_sample(pred, target, max_trials, table, seed=123, sample_offset=0):
    """The sampling loop of WarpLoss.forward / WARP_funcLoss.forward (mlc_loss.py:129-147, :198-215) with
    np.random.choice replaced by the counter-based stream of the contract:
        u = philox4x32_10((sample_offset + b, j, trial, 0), (seed_lo, seed_hi))[0];  neg = negatives[u % n_neg].
    Returns (L[B,C] float32, trials[B,C] int)."""
    p = np.asarray(pred, np.float32)
    t = np.asarray(target)
    B, C = p.shape
    L = np.zeros((B, C), np.float32)
    trials = np.zeros((B, C), np.int32)
    key = (seed & MASK, (seed >> 32) & MASK)
    for b in range(B):
        negs = [j for j in range(C) if t[b, j] == 0]
        for j in range(C):
            if t[b, j] == 1:
                if not negs:
                    L[b, j] = np.nan                           # the reference loops forever here (:140-142)
                    continue
                margin, n = -1.0, 0
                while margin < 0 and n < max_trials:
                    n += 1
                    u = philox4x32_10(((sample_offset + b) & MASK, j, n, 0), key)[0]
                    margin = p[b, negs[u % len(negs)]] - p[b, j]
                r_j = int(np.floor(max_trials / n))
                L[b, j] = table[r_j]
                trials[b, j] = n
    return L, trials


def warp_loss(pred, target, L, dtype=np.float32):
    """WarpLoss.forward after sampling (mlc_loss.py:151-174) + autodiff gradient (L is a constant)."""
    p = np.asarray(pred, dtype)
    t = np.asarray(target)
    dist = p[:, :, None] - p[:, None, :]
    pos = (t > 0)[:, :, None].astype(dtype)
    neg = (t == 0)[:, None, :].astype(dtype)
    filt = pos * neg
    el = np.maximum(1 + filt * (-dist), 0)
    Lr = np.asarray(L, dtype)[:, :, None]
    loss = (Lr * el).sum(dtype=dtype)
    active = (el > 0).astype(dtype) * filt * Lr                # d el / d(p_j - p_i)
    grad = active.sum(axis=1) - active.sum(axis=2)
    return loss, grad.astype(dtype)


def warp_func_loss(pred, target, L, dtype=np.float32):
    """WARP_funcLoss forward value and hand-written backward (mlc_loss.py:217-232), head gradient 1."""
    p = np.asarray(pred, dtype)
    t = np.asarray(target)
    pos = (t > 0).astype(dtype)
    neg = (t == 0).astype(dtype)
    Ls = np.asarray(L, dtype).sum(axis=1, keepdims=True)
    loss = (Ls * (1 - pos * p + neg * p).sum(axis=1, keepdims=True)).sum(dtype=dtype)
    grad = Ls * (neg - pos)
    return loss, grad.astype(dtype)


def sigmoid_bce(pred, target, from_sigmoid=False, dtype=np.float32):
    """gluon.loss.SigmoidBinaryCrossEntropyLoss (MXNet 1.x): per-sample mean over classes; gradient of sum_b loss_b."""
    x = np.asarray(pred, dtype)
    z = np.asarray(target, dtype)
    C = x.shape[1]
    if not from_sigmoid:
        l = np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))
        g = 1 / (1 + np.exp(-x)) - z
    else:
        eps = dtype(1e-12)
        l = -(np.log(x + eps) * z + np.log(1 - x + eps) * (1 - z))
        g = -(z / (x + eps) - (1 - z) / (1 - x + eps))
    return l.mean(axis=1).astype(dtype), (g / C).astype(dtype)


def softmax_ce(logits, label, dtype=np.float32):
    """gluon.loss.SoftmaxCrossEntropyLoss, sparse labels (train_simple_r3d.py:43): loss[B], grad = softmax - onehot."""
    x = np.asarray(logits, dtype)
    m = x.max(axis=1, keepdims=True)
    lse = m + np.log(np.exp(x - m).sum(axis=1, keepdims=True))
    logp = x - lse
    idx = np.asarray(label).astype(int)
    loss = -logp[np.arange(x.shape[0]), idx]
    g = np.exp(logp)
    g[np.arange(x.shape[0]), idx] -= 1
    return loss.astype(dtype), g.astype(dtype)


def softmax_output(logits, label, ignore_label=-1, dtype=np.float32):
    """mx.sym.SoftmaxOutput(multi_output=True, use_ignore=True, normalization='null') (net.py:167-169):
    forward = softmax; backward = p - onehot, zero for ignored rows."""
    x = np.asarray(logits, dtype)
    m = x.max(axis=1, keepdims=True)
    e = np.exp(x - m)
    prob = e / e.sum(axis=1, keepdims=True)
    g = prob.copy()
    lab = np.asarray(label)
    for b in range(x.shape[0]):
        if lab[b] == ignore_label:
            g[b] = 0
        else:
            g[b, int(lab[b])] -= 1
    return prob.astype(dtype), g.astype(dtype)
