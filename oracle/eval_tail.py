"""CPU restatement (TEST INFRASTRUCTURE ONLY) of the steps next to the hot path: evaluation tail and clip normalisation.
Only tests/ may import this module.  Each function cites the reference lines it follows."""
import numpy as np


def softmax(x):
    x = np.asarray(x, dtype=np.float64)
    e = np.exp(x - x.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


def multi_clip_accuracy(logits_per_pass, labels):
    """validation.py:39-66: outputs += softmax(batch) for every pass, argmax, mean(pred == label)."""
    acc = np.zeros_like(np.asarray(logits_per_pass[0], dtype=np.float64))
    for lg in logits_per_pass:
        acc += softmax(lg)
    pred = np.argmax(acc, axis=1)
    return acc, pred, float((pred == np.asarray(labels)).sum()) / len(labels)


def topk_iou_counts(y_hat, y, k=4):
    """train_simple_r3d.py:170-193, literally: argsort ascending, reversed; label set = value > 0.1; set arithmetic."""
    inter = np.zeros(k, dtype=np.int64)
    union = np.zeros(k, dtype=np.int64)
    pred_order = np.argsort(np.asarray(y_hat), axis=1, kind="stable")[:, ::-1]
    for pred_vec, y_vec in zip(pred_order, np.asarray(y)):
        label_set = set(index for index, value in enumerate(y_vec) if value > 0.1)
        pred_topk = [set(pred_vec[0:kk].tolist()) for kk in range(1, k + 1)]
        inter += np.array([len(p_k.intersection(label_set)) for p_k in pred_topk])
        union += np.array([len(p_k.union(label_set)) for p_k in pred_topk])
    return inter, union


def normalize_batch(clips_u8_nthwc, flip=None):
    """videos_reader.py:69-76,93-97: DHWC -> CDHW, optional horizontal flip, per-channel batch mean / std,
    (x - m) / (std + 1e-3)."""
    x = np.asarray(clips_u8_nthwc).astype(np.float32).transpose(0, 4, 1, 2, 3).copy()      # N, C, T, H, W
    if flip is not None:
        for i, f in enumerate(flip):
            if f:
                x[i] = np.flip(x[i], 3)
    m = np.mean(x.astype(np.float64), axis=(0, 2, 3, 4))
    std = np.std(x.astype(np.float64), axis=(0, 2, 3, 4))
    out = np.empty_like(x)
    for i in range(3):
        out[:, i] = (x[:, i] - m[i]) / (std[i] + 1e-3)
    return out, m, std


def normalize_imagenet(clips_u8_nthwc):
    """data/ucf101.py:124-128: ToTensor (x / 255, HWC -> CHW) then Normalize(mean, std)."""
    x = np.asarray(clips_u8_nthwc).astype(np.float32).transpose(0, 4, 1, 2, 3) / 255.0
    mean = np.array([0.485, 0.456, 0.406], dtype=np.float32).reshape(1, 3, 1, 1, 1)
    std = np.array([0.229, 0.224, 0.225], dtype=np.float32).reshape(1, 3, 1, 1, 1)
    return (x - mean) / std
