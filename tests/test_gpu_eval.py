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


@pytest.mark.parametrize("hpair", [True, False])
def test_clip_unfold_u8_is_the_two_pass_route_in_one_kernel(cuda_device, hpair):
    """N2 as SURVEY 8f states it: uint8 frames -> crop / flip / normalise -> W-unfolded NDHWC bf16 stem input in ONE kernel
    (fvt_clip_unfold_u8) == crop on the host, fvt_clip_normalize_u8 (fp32 NCDHW), then fvt_stem_unfold[_hpair], bit for
    bit; and == the oracle's normalisation (oracle/eval_tail.py, videos_reader.py:93-97) to bf16 rounding."""
    from fastvideotagging_b200 import ops
    from fastvideotagging_b200.evaluate import normalize_clips, batch_statistics
    from oracle import eval_tail as et
    rng = np.random.default_rng(5)
    n, t, hs, ws, h, w = 3, 4, 64, 86, 56, 56                       # decoded 64x86 frames, random 56x56 crops
    frames = rng.integers(0, 256, size=(n, t, hs, ws, 3), dtype=np.uint8)
    crop = np.stack([rng.integers(0, hs - h + 1, size=n), rng.integers(0, ws - w + 1, size=n)], axis=1).astype(np.int32)
    flip = np.array([1, 0, 1], dtype=np.uint8)
    cropped = np.stack([frames[i, :, crop[i, 0]:crop[i, 0] + h, crop[i, 1]:crop[i, 1] + w] for i in range(n)])
    cd = torch.from_numpy(cropped).to(cuda_device)
    mean, std = batch_statistics(cd)
    inv = [1.0 / (s + 1e-3) for s in std]
    two_pass_f32, m2, s2 = normalize_clips(cd, torch.from_numpy(flip), mode="batch")
    assert np.allclose(m2.numpy(), mean, rtol=1e-6) and np.allclose(s2.numpy(), std, rtol=1e-6)
    unfold = ops.stem_unfold_hpair if hpair else ops.stem_unfold
    ref = unfold(two_pass_f32)
    got = torch.full_like(ref, float("nan"))
    ops.clip_unfold_u8(torch.from_numpy(frames).to(cuda_device), got, 1.0, mean, inv, hpair, flip=torch.from_numpy(flip),
                       crop_yx=torch.from_numpy(crop), crop_hw=(h, w))
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
    # against the oracle's normalisation: channel kw*3 + c of output column ow is pixel ow*2 - 3 + kw
    o_ref, _, _ = et.normalize_batch(cropped, flip)                  # (N, 3, T, H, W) fp32
    g = got.float().cpu().numpy()
    if hpair:
        g = g.reshape(n, t, h // 2, w // 2, 2, 32).transpose(0, 1, 2, 4, 3, 5).reshape(n, t, h, w // 2, 32)
    for kw in (0, 3, 6):
        for ow in (0, 5, w // 2 - 1):
            iw = ow * 2 - 3 + kw
            want = o_ref[:, :, :, :, iw].transpose(0, 2, 3, 1) if 0 <= iw < w else np.zeros((n, t, h, 3), np.float32)
            assert np.abs(g[:, :, :, ow, kw * 3:kw * 3 + 3] - want).max() <= 2 ** -8 * max(1.0, np.abs(want).max())


def test_uint8_clips_run_the_network_directly(cuda_device):
    """R2Plus2D accepts decoded uint8 frames (N, T, H, W, 3) once the normalisation constants are set: same logits as the
    fp32 route, eval and training forward (the host ships 1 byte per value instead of 4)."""
    from fastvideotagging_b200.evaluate import normalize_clips, batch_statistics
    from fastvideotagging_b200.model import R2Plus2D
    from oracle import r2plus1d as orc
    rng = np.random.default_rng(6)
    x = rng.integers(0, 256, size=(2, 8, 112, 112, 3), dtype=np.uint8)
    xd = torch.from_numpy(x).to(cuda_device)
    net = R2Plus2D(101, 10, final_spatial_kernel=7, final_temporal_kernel=1).to(cuda_device)
    net.load_param_dict(orc.randomize_bn(orc.init_params(10, 101, seed=0), seed=1))
    mean, std = batch_statistics(xd)
    net.set_input_normalization(mean, std, scale=1.0, std_eps=1e-3)
    x32 = normalize_clips(xd, mode="batch")[0]
    net.eval()
    with torch.no_grad():
        a, b = net(xd), net(x32)
    assert torch.equal(a, b)
    net.train()
    a = net(xd).detach().clone()
    b = net(x32).detach().clone()
    assert torch.equal(a, b)
