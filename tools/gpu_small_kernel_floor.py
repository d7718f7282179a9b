"""What one launch of the small-layer kernels costs INSIDE a replayed CUDA graph (conv4_x / conv5_x shapes of BASELINE
configs[2], batch 4): 40 dependent launches of one kernel captured in a graph, replayed; per-launch time = replay / 40.
usage: gpu_small_kernel_floor.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastvideotagging_b200 import ops
dev = torch.device("cuda:0")
N, REP = 4, 40

def graph_time(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 5 / REP * 1e3

for name, t, hw, c in (("conv3_x 288ch", 16, 28, 288), ("conv4_x 256ch", 8, 14, 256), ("conv4_x 576ch", 8, 14, 576), ("conv5_x 512ch", 4, 7, 512), ("conv5_x 1152ch", 4, 7, 1152)):
    rows = N * t * hw * hw
    raw = torch.randn(N, t, hw, hw, c, device=dev).to(torch.bfloat16)
    dz = torch.randn_like(raw); out = torch.empty_like(raw); draw = torch.empty_like(raw)
    gamma = torch.rand(c, device=dev) + 0.5; beta = torch.zeros(c, device=dev)
    mean = torch.zeros(c, device=dev); invstd = torch.ones(c, device=dev); scale = torch.ones(c, device=dev); shift = torch.zeros(c, device=dev)
    rm = torch.zeros(c, device=dev); rv = torch.ones(c, device=dev)
    sums = torch.empty(2 * c, device=dev)
    acc = ops.stats_buffer(c, dev); stats = ops.stats_buffer(c, dev)
    t1 = graph_time(lambda: ops.bn_backward(raw, dz, None, mean, invstd, gamma, sums, draw, sums_acc=acc, dz_in=2))
    t2 = graph_time(lambda: ops.bn_backward(raw, dz, None, mean, invstd, gamma, sums, draw, relu_scale=scale, relu_shift=shift))
    t3 = graph_time(lambda: ops.bn_finalize_apply(stats, gamma, beta, rm, rv, c, rows, 1e-5, 0.9, scale, shift, mean, invstd, raw, out, True))
    t4 = graph_time(lambda: ops.bn_apply(raw, scale, shift, out, True))
    by = raw.numel() * 2
    print("%-16s rows %6d (%5.1f MB): bn_bwd apply-only %5.1f us | bn_bwd two-pass %5.1f us | bn finalize+apply %5.1f us | bn apply %5.1f us | copy floor %4.1f us"
          % (name, rows, by / 1e6, t1, t2, t3, t4, 2 * by / 6.5e12 * 1e6), flush=True)
# small convs
for name, t, hw, cin, cout, k, p in (("conv4_x 1x3x3 256->576", 8, 14, 256, 576, (1, 3, 3), (0, 1, 1)), ("conv4_x 3x1x1 576->256", 8, 14, 576, 256, (3, 1, 1), (1, 0, 0)),
                                      ("conv5_x 1x3x3 512->1152", 4, 7, 512, 1152, (1, 3, 3), (0, 1, 1)), ("conv5_x 3x1x1 1152->512", 4, 7, 1152, 512, (3, 1, 1), (1, 0, 0))):
    fwd = ops.conv_desc(N, t, hw, hw, cin, cout, k, (1, 1, 1), p, ops.FVT_CONV_STATS)
    x = torch.randn(N, t, hw, hw, cin, device=dev).to(torch.bfloat16)
    wm = torch.randn(cout, k[0], k[1], k[2], cin, device=dev) * 0.02
    wp = ops.pack_conv_weight(fwd, wm.permute(0, 4, 1, 2, 3).contiguous())
    y = torch.empty(N, t, hw, hw, cout, device=dev, dtype=torch.bfloat16)
    stats = ops.stats_buffer(cout, dev)
    t1 = graph_time(lambda: ops.conv3d_fwd(fwd, x, wp, out=y, stats=stats))
    fl = 2.0 * N * t * hw * hw * cout * cin * k[0] * k[1] * k[2]
    print("%-26s forward + statistics %5.1f us   (tensor roofline %4.1f us)" % (name, t1, fl / 1382.8e12 * 1e6), flush=True)
