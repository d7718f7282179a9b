"""Data-gradient convolutions of the training plan (BASELINE configs[2]: batch 4, 32 x 112 x 112), plain vs fused with the
consumer BatchNorm's backward sums (FVT_CONV_BN_BWD), each timed alone with CUDA events; the BatchNorm backward that
follows is timed too (two passes after the plain form, one after the fused one).
usage: gpu_dgrad_fused.py [reps]   (reps = 1: single launches, for ncu)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastvideotagging_b200 import ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
N = 4
SHAPES = [  # name, t, h, w, cin, cout of the FORWARD conv, kernel, pad   (the data gradient maps cout -> cin)
    ("conv2_x 3x1x1 144->64", 32, 56, 56, 144, 64, (3, 1, 1), (1, 0, 0)),
    ("conv2_x 1x3x3 64->144", 32, 56, 56, 64, 144, (1, 3, 3), (0, 1, 1)),
    ("conv3_x 3x1x1 288->128", 16, 28, 28, 288, 128, (3, 1, 1), (1, 0, 0)),
    ("conv3_x 1x3x3 128->288", 16, 28, 28, 128, 288, (1, 3, 3), (0, 1, 1)),
    ("conv4_x 3x1x1 576->256", 8, 14, 14, 576, 256, (3, 1, 1), (1, 0, 0)),
    ("conv4_x 1x3x3 256->576", 8, 14, 14, 256, 576, (1, 3, 3), (0, 1, 1)),
    ("conv5_x 3x1x1 1152->512", 4, 7, 7, 1152, 512, (3, 1, 1), (1, 0, 0)),
    ("conv5_x 1x3x3 512->1152", 4, 7, 7, 512, 1152, (1, 3, 3), (0, 1, 1)),
]
only = os.environ.get("FVT_ONLY", "")
for kv in os.environ.get("FVT_DBG_OPTS", "").split(","):
    if kv:
        k_, v_ = kv.split("="); assert ops.set_option(k_, int(v_)) == 0
def timeit(fn):
    fn(); torch.cuda.synchronize()
    if reps <= 1:
        return 0.0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
print("%-26s %8s | %9s %9s | %9s %9s | %9s %9s" % ("forward layer", "rows", "dgrad", "fused", "bn 2pass", "bn 1pass", "sum plain", "sum fused"))
for name, t, h, w, cin, cout, k, p in SHAPES:
    if only and only not in name:
        continue
    cin_s, cout_s = ops.pad16(cin), ops.pad16(cout)
    fwd = ops.conv_desc(N, t, h, w, cin_s, cout_s, k, (1, 1, 1), p)
    dd = ops.dgrad_desc(fwd)
    wm = torch.randn(cout, k[0], k[1], k[2], cin, device=dev) / (cout * k[0] * k[1] * k[2]) ** 0.5
    wpd = ops.pack_conv_weight_dgrad(dd, wm, ohwi=True)
    dy = torch.randn(N, t, h, w, cout_s, device=dev).to(torch.bfloat16)
    raw = torch.randn(N, t, h, w, cin_s, device=dev).to(torch.bfloat16)
    rows = N * t * h * w
    gamma = torch.rand(cin, device=dev) + 0.5
    mean = torch.zeros(cin_s, device=dev); invstd = torch.ones(cin_s, device=dev)
    scale = torch.ones(cin_s, device=dev); shift = torch.zeros(cin_s, device=dev)
    out = torch.empty_like(raw); draw = torch.empty_like(raw)
    sums = torch.empty(2 * cin_s, device=dev)
    acc = ops.stats_buffer(cin_s, dev)
    d2 = ops.ConvDesc(*dd.key())
    d2.flags = ops.FVT_CONV_STATS | ops.FVT_CONV_BN_BWD | ops.FVT_CONV_RESIDUAL
    t_plain = timeit(lambda: ops.conv3d_fwd(dd, dy, wpd, out=out))
    t_fused = timeit(lambda: ops.conv3d_fwd(d2, dy, wpd, scale=scale, shift=shift, residual=raw, out=out, stats=acc))
    t_bn2 = timeit(lambda: ops.bn_backward(raw, out, None, mean, invstd, gamma, sums, draw, relu_scale=scale, relu_shift=shift))
    t_bn1 = timeit(lambda: ops.bn_backward(raw, out, None, mean, invstd, gamma, sums, draw, sums_acc=acc, dz_in=2))
    print("%-26s %8d | %9.1f %9.1f | %9.1f %9.1f | %9.1f %9.1f" % (name, rows, t_plain, t_fused, t_bn2, t_bn1, t_plain + t_bn2, t_fused + t_bn1), flush=True)
