"""Oracle of the evaluation tail / clip normalisation against hand-checkable cases (CPU)."""
import numpy as np

from oracle import eval_tail as et


def test_topk_iou_hand_case():
    # scores descending: classes 2, 0, 3, 1; labels {0, 3}
    y_hat = np.array([[0.5, -1.0, 0.9, 0.1]])
    y = np.array([[1.0, 0.0, 0.0, 1.0]])
    inter, union = et.topk_iou_counts(y_hat, y, 4)
    assert inter.tolist() == [0, 1, 2, 2] and union.tolist() == [3, 3, 3, 4]


def test_topk_iou_tie_order_is_reversed_stable_argsort():
    # equal scores: argsort ascending-stable gives [0, 1, 2, 3], reversed -> class 3 first
    y_hat = np.zeros((1, 4))
    y = np.array([[0.0, 0.0, 0.0, 1.0]])
    inter, union = et.topk_iou_counts(y_hat, y, 2)
    assert inter.tolist() == [1, 1] and union.tolist() == [1, 2]


def test_multi_clip_accuracy_sums_softmax_not_logits():
    a = np.array([[10.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    b = np.array([[0.0, 2.0, 2.0], [0.0, 1.0, 0.0]])
    acc, pred, accuracy = et.multi_clip_accuracy([a, b], [0, 1])
    assert pred.tolist() == [0, 1] and accuracy == 1.0
    assert np.allclose(acc.sum(axis=1), 2.0)


def test_normalize_batch_matches_formula():
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, size=(2, 3, 4, 5, 3), dtype=np.uint8)
    out, m, std = et.normalize_batch(x, flip=[0, 1])
    assert out.shape == (2, 3, 3, 4, 5)
    assert np.allclose(out[0, 1, 2, 3, 4], (float(x[0, 2, 3, 4, 1]) - m[1]) / (std[1] + 1e-3), atol=1e-5)
    assert np.allclose(out[1, 2, 0, 1, 0], (float(x[1, 0, 1, 4, 2]) - m[2]) / (std[2] + 1e-3), atol=1e-5)      # flipped clip
