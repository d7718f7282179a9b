"""GPU parity of the rows next to the hot path (N2 clip normalisation, N3 evaluation tail) against oracle/eval_tail.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,c", [(7, 101), (16, 63), (3, 4), (33, 400)])
def test_multi_clip_accumulator(cuda_device, rows, c):
    from fastvideotagging_b200.evaluate import MultiClipAccumulator
    from oracle import eval_tail as et
    rng = np.random.default_rng(rows * 1000 + c)
    passes = [rng.normal(size=(rows, c)).astype(np.float32) * 3 for _ in range(5)]
    labels = rng.integers(0, c, size=rows)
    acc = MultiClipAccumulator(rows, c, cuda_device)
    for lg in passes:
        half = rows // 2
        acc.add(0, torch.from_numpy(lg[:half]).to(cuda_device)) if half else None
        acc.add(half, torch.from_numpy(lg[half:]).to(cuda_device))
    pred, accuracy = acc.predictions(torch.from_numpy(labels))
    ref_acc, ref_pred, ref_accuracy = et.multi_clip_accuracy(passes, labels)
    assert np.abs(acc.acc.cpu().numpy() - ref_acc).max() <= 1e-5
    assert (pred.cpu().numpy() == ref_pred).all() and accuracy == ref_accuracy


@pytest.mark.parametrize("rows,c", [(16, 63), (5, 101), (4, 3), (9, 40)])
def test_topk_iou_counts_bit_exact(cuda_device, rows, c):
    from fastvideotagging_b200.evaluate import TopkIoU
    from oracle import eval_tail as et
    rng = np.random.default_rng(rows + c)
    y_hat = np.round(rng.normal(size=(rows, c)), 1).astype(np.float32)          # rounded: plenty of exact ties
    y = (rng.random((rows, c)) < 0.1).astype(np.float32)
    k = min(4, c)
    m = TopkIoU(k, cuda_device)
    m.update(torch.from_numpy(y_hat).to(cuda_device), torch.from_numpy(y).to(cuda_device))
    m.update(torch.from_numpy(y_hat[::-1].copy()).to(cuda_device), torch.from_numpy(y[::-1].copy()).to(cuda_device))
    inter, union = et.topk_iou_counts(y_hat, y, k)
    assert m.inter.cpu().numpy().tolist() == (2 * inter).tolist()
    assert m.union.cpu().numpy().tolist() == (2 * union).tolist()
    assert np.allclose(m.value(), (2 * inter + 1e-4) / (2 * union + 1e-4))


def test_normalize_clips_batch_and_imagenet(cuda_device):
    from fastvideotagging_b200.evaluate import normalize_clips
    from oracle import eval_tail as et
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, size=(3, 8, 28, 36, 3), dtype=np.uint8)
    flip = np.array([1, 0, 1], dtype=np.uint8)
    out, mean, std = normalize_clips(torch.from_numpy(x).to(cuda_device), torch.from_numpy(flip), mode="batch")
    ref, m, s = et.normalize_batch(x, flip)
    assert np.allclose(mean.numpy(), m, rtol=1e-6) and np.allclose(std.numpy(), s, rtol=1e-6)
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max()
    out2 = normalize_clips(torch.from_numpy(x).to(cuda_device), mode="imagenet")
    assert np.abs(out2.cpu().numpy() - et.normalize_imagenet(x)).max() <= 1e-5


def test_normalized_uint8_clips_feed_the_network(cuda_device):
    """uint8 frames -> normalize_clips -> R2Plus2D gives the logits of the same clips normalised on the host."""
    from fastvideotagging_b200.evaluate import normalize_clips
    from fastvideotagging_b200.model import R2Plus2D
    from oracle import eval_tail as et
    from oracle import r2plus1d as orc
    rng = np.random.default_rng(4)
    x = rng.integers(0, 256, size=(2, 8, 112, 112, 3), dtype=np.uint8)
    net = R2Plus2D(101, 10, final_spatial_kernel=7, final_temporal_kernel=1).to(cuda_device)
    net.load_param_dict(orc.randomize_bn(orc.init_params(10, 101, seed=0), seed=1))
    net.eval()
    with torch.no_grad():
        a = net(normalize_clips(torch.from_numpy(x).to(cuda_device), mode="batch")[0]).cpu().numpy()
        b = net(torch.from_numpy(et.normalize_batch(x)[0]).to(cuda_device)).cpu().numpy()
    assert np.abs(a - b).max() <= 2e-2 * np.abs(b).max() and (a.argmax(1) == b.argmax(1)).all()
