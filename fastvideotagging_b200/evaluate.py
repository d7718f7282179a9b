"""Evaluation tail and clip pre-processing on the device — the steps directly after and
directly before the R(2+1)D hot path, so that an evaluation loop has no per-batch `asnumpy()` synchronisation and a
training loop uploads uint8 frames instead of normalised fp32 clips.

    MultiClipAccumulator   validation.py:39-66   softmax of every sampled clip summed per video, argmax, accuracy
    TopkIoU                train_simple_r3d.py:169-197   top-k (k = 1..4) intersection-over-union of multi-label predictions
    normalize_clips        videos_reader.py:69-76,93-97 (per-batch statistics) / data/ucf101.py:124-128 (ImageNet statistics)
"""
import ctypes

import torch

from . import _lib
from .ops import _ptr, _stream, require_cuda

check = _lib.check


class MultiClipAccumulator:
    """outputs[i_batch] += softmax(batch) over `clips_per_video` passes, then argmax (validation.py:39-66)."""

    def __init__(self, n_rows, num_class, device):
        self.acc = torch.zeros((n_rows, num_class), dtype=torch.float32, device=device)
        self.n_rows, self.num_class = n_rows, num_class

    def add(self, row0, logits):
        """Accumulate softmax(logits) into rows [row0, row0 + len(logits))."""
        require_cuda(logits, "logits")
        logits = logits.float().contiguous()
        rows = logits.shape[0]
        assert logits.shape[1] == self.num_class and 0 <= row0 and row0 + rows <= self.n_rows
        dst = self.acc[row0:row0 + rows]
        check(_lib.load().fvt_softmax_accumulate(_lib.handle(), _ptr(logits), _ptr(dst), rows, self.num_class, _stream()))

    def predictions(self, labels=None):
        """-> (pred (n_rows,) int32, accuracy or None).  One device->host read for the accuracy, none without labels."""
        pred = torch.empty(self.n_rows, dtype=torch.int32, device=self.acc.device)
        correct = torch.zeros(1, dtype=torch.int64, device=self.acc.device)
        lab = labels.to(self.acc.device, torch.int32).contiguous() if labels is not None else None
        check(_lib.load().fvt_argmax_correct(_lib.handle(), _ptr(self.acc), _ptr(lab), self.n_rows, self.num_class, _ptr(pred), _ptr(correct), _stream()))
        return pred, (correct.item() / float(self.n_rows) if labels is not None else None)


class TopkIoU:
    """topk_inter / topk_union of train_simple_r3d.py:170-193 (both start at 1e-4, the ratio is the logged IoU)."""

    def __init__(self, k=4, device=None):
        self.k = k
        self.inter = torch.zeros(k, dtype=torch.int64, device=device)
        self.union = torch.zeros(k, dtype=torch.int64, device=device)

    def update(self, y_hat, y):
        require_cuda(y_hat, "y_hat")
        y_hat = y_hat.float().contiguous()
        y = y.to(y_hat.device, torch.float32).contiguous()
        check(_lib.load().fvt_topk_iou(_lib.handle(), _ptr(y_hat), _ptr(y), y_hat.shape[0], y_hat.shape[1], self.k, _ptr(self.inter),
                                       _ptr(self.union), _stream()))

    def value(self):
        return ((self.inter.double() + 1e-4) / (self.union.double() + 1e-4)).cpu().numpy()


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def batch_statistics(clips_u8):
    """Per-channel mean and population std over a whole uint8 clip batch (N, T, H, W, 3), exact integer sums on the device
    (videos_reader.py:93-94).  Returns (mean[3], std[3]) as python floats, for R2Plus2D.set_input_normalization(mean, std,
    scale=1.0, std_eps=1e-3)."""
    require_cuda(clips_u8, "clips")
    assert clips_u8.dtype == torch.uint8 and clips_u8.dim() == 5 and clips_u8.shape[-1] == 3 and clips_u8.is_contiguous()
    lib = _lib.load()
    sums = torch.empty(6, dtype=torch.int64, device=clips_u8.device)
    pixels = clips_u8.numel() // 3
    check(lib.fvt_clip_stats_u8(_lib.handle(), _ptr(clips_u8), pixels, _ptr(sums), _stream()))
    s = sums.cpu().double()
    mean = s[:3] / pixels
    std = torch.sqrt(torch.clamp(s[3:] / pixels - mean * mean, min=0.0))
    return mean.tolist(), std.tolist()


def normalize_clips(clips_u8, flip=None, mode="batch"):
    """clips_u8: (N, T, H, W, 3) uint8 CUDA tensor of decoded, cropped frames -> (N, 3, T, H, W) fp32, the layout and
    values the reference feeds the network.

    mode 'batch'    (videos_reader.py:93-97): per-channel mean / population std over the whole batch,
                    (x - mean) / (std + 1e-3);  returns (clips, mean, std)
    mode 'imagenet' (data/ucf101.py:124-128): ToTensor (x / 255) then (x - mean) / std with the ImageNet constants.
    flip: optional (N,) bool/uint8 tensor — horizontal flip per clip (videos_reader.py:74-75)."""
    require_cuda(clips_u8, "clips")
    assert clips_u8.dtype == torch.uint8 and clips_u8.dim() == 5 and clips_u8.shape[-1] == 3 and clips_u8.is_contiguous()
    lib = _lib.load()
    n, t, h, w, _ = clips_u8.shape
    out = torch.empty((n, 3, t, h, w), dtype=torch.float32, device=clips_u8.device)
    flip_t = flip.to(clips_u8.device, torch.uint8).contiguous() if flip is not None else None
    f3 = ctypes.c_float * 3
    if mode == "batch":
        sums = torch.empty(6, dtype=torch.int64, device=clips_u8.device)
        pixels = n * t * h * w
        check(lib.fvt_clip_stats_u8(_lib.handle(), _ptr(clips_u8), pixels, _ptr(sums), _stream()))
        s = sums.cpu().double()
        mean = s[:3] / pixels
        std = torch.sqrt(torch.clamp(s[3:] / pixels - mean * mean, min=0.0))
        inv = 1.0 / (std + 1e-3)
        check(lib.fvt_clip_normalize_u8(_lib.handle(), _ptr(clips_u8), _ptr(flip_t), _ptr(out), n, t, h, w, ctypes.c_float(1.0),
                                        f3(*mean.tolist()), f3(*inv.tolist()), _stream()))
        return out, mean.float(), std.float()
    if mode == "imagenet":
        check(lib.fvt_clip_normalize_u8(_lib.handle(), _ptr(clips_u8), _ptr(flip_t), _ptr(out), n, t, h, w, ctypes.c_float(1.0 / 255.0),
                                        f3(*IMAGENET_MEAN), f3(*[1.0 / v for v in IMAGENET_STD]), _stream()))
        return out
    raise ValueError("mode must be 'batch' or 'imagenet'")
