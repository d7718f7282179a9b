"""Small-M convolutions (conv4_x / conv5_x at batch 4) under different N tile widths (desc.block_n), with and without split-K,
timed inside a replayed CUDA graph (40 dependent launches).  usage: gpu_small_conv_tiles.py"""
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

for name, t, hw, cin, cout, k, p in (("conv4_x 1x3x3 256->576", 8, 14, 256, 576, (1, 3, 3), (0, 1, 1)), ("conv4_x 3x1x1 576->256", 8, 14, 576, 256, (3, 1, 1), (1, 0, 0)),
                                      ("conv5_x 1x3x3 512->1152", 4, 7, 512, 1152, (1, 3, 3), (0, 1, 1)), ("conv5_x 3x1x1 1152->512", 4, 7, 1152, 512, (3, 1, 1), (1, 0, 0)),
                                      ("conv4_x dgrad 576->256 (1x3x3)", 8, 14, 576, 256, (1, 3, 3), (0, 1, 1)), ("conv5_x dgrad 1152->512 (1x3x3)", 4, 7, 1152, 512, (1, 3, 3), (0, 1, 1))):
    x = torch.randn(N, t, hw, hw, cin, device=dev).to(torch.bfloat16)
    wm = torch.randn(cout, k[0], k[1], k[2], cin, device=dev) * 0.02
    y = torch.empty(N, t, hw, hw, cout, device=dev, dtype=torch.bfloat16)
    stats = ops.stats_buffer(cout, dev)
    res = []
    for bn in (0, 64, 96, 128, 144, 192, 256):
        if bn and (cout % bn):
            continue
        for split in (1, 0):
            ops.set_option("disable_split_k", 0 if split else 1)
            try:
                fwd = ops.conv_desc(N, t, hw, hw, cin, cout, k, (1, 1, 1), p, ops.FVT_CONV_STATS, bn)
                wp = ops.pack_conv_weight(fwd, wm.permute(0, 4, 1, 2, 3).contiguous())
                us = graph_time(lambda: ops.conv3d_fwd(fwd, x, wp, out=y, stats=stats))
                res.append("bn %3d%s %5.1f" % (bn, " +splitK" if split else "        ", us))
            except Exception as e:
                res.append("bn %3d%s  n/a" % (bn, " +splitK" if split else "        "))
    ops.set_option("disable_split_k", 0)
    fl = 2.0 * N * t * hw * hw * cout * cin * k[0] * k[1] * k[2]
    print("%-32s roofline %4.1f us | %s" % (name, fl / 1382.8e12 * 1e6, " | ".join(res)), flush=True)

# data gradient + the BatchNorm backward it feeds: plain (split-K allowed) + two passes  vs  fused epilogue + one pass
print()
for name, t, hw, cin, cout, k, p in (("conv3_x 1x3x3 128->288", 16, 28, 128, 288, (1, 3, 3), (0, 1, 1)), ("conv3_x 3x1x1 288->128", 16, 28, 288, 128, (3, 1, 1), (1, 0, 0)),
                                      ("conv4_x 1x3x3 256->576", 8, 14, 256, 576, (1, 3, 3), (0, 1, 1)), ("conv4_x 3x1x1 576->256", 8, 14, 576, 256, (3, 1, 1), (1, 0, 0)),
                                      ("conv5_x 1x3x3 512->1152", 4, 7, 512, 1152, (1, 3, 3), (0, 1, 1)), ("conv5_x 3x1x1 1152->512", 4, 7, 1152, 512, (3, 1, 1), (1, 0, 0))):
    fwd = ops.conv_desc(N, t, hw, hw, cin, cout, k, (1, 1, 1), p)
    dd = ops.dgrad_desc(fwd)
    wm = torch.randn(cout, k[0], k[1], k[2], cin, device=dev) * 0.02
    wpd = ops.pack_conv_weight_dgrad(dd, wm, ohwi=True)
    dy = torch.randn(N, t, hw, hw, cout, device=dev).to(torch.bfloat16)
    raw = torch.randn(N, t, hw, hw, cin, device=dev).to(torch.bfloat16)
    gamma = torch.rand(cin, device=dev) + 0.5
    mean = torch.zeros(cin, device=dev); invstd = torch.ones(cin, device=dev); scale = torch.ones(cin, device=dev); shift = torch.zeros(cin, device=dev)
    out = torch.empty_like(raw); draw = torch.empty_like(raw)
    sums = torch.empty(2 * cin, device=dev); acc = ops.stats_buffer(cin, dev); acc2 = ops.stats_buffer(cin, dev)
    d2 = ops.ConvDesc(*dd.key()); d2.flags = ops.FVT_CONV_STATS | ops.FVT_CONV_BN_BWD | ops.FVT_CONV_RESIDUAL
    def plain():
        ops.conv3d_fwd(dd, dy, wpd, out=out)
        ops.bn_backward(raw, out, None, mean, invstd, gamma, sums, draw, relu_scale=scale, relu_shift=shift, sums_acc=acc2)
    def fused():
        acc.zero_()
        ops.conv3d_fwd(d2, dy, wpd, scale=scale, shift=shift, residual=raw, out=out, stats=acc)
        ops.bn_backward(raw, out, None, mean, invstd, gamma, sums, draw, sums_acc=acc, dz_in=2)
    print("%-26s data gradient + BatchNorm backward: plain + two passes %5.1f us | fused + one pass %5.1f us" % (name, graph_time(plain), graph_time(fused)), flush=True)
