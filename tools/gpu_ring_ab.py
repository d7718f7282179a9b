"""Frame-ring (K1t) vs im2col (K1) timing of 3x1x1 convs for several channel counts (experiments)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastvideotagging_b200 import ops, _lib
lib = _lib.load()
dev = torch.device("cuda:0")
if len(sys.argv) > 1: ops.set_option("ring_prefetch", int(sys.argv[1]))
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for (n, t, h, w, cin, cout) in [(48, 32, 56, 56, 144, 64), (48, 32, 56, 56, 128, 64), (48, 32, 56, 56, 64, 64), (48, 32, 56, 56, 64, 144),
                                (48, 32, 56, 56, 192, 64), (4, 32, 56, 56, 144, 64), (4, 32, 56, 56, 64, 144)]:
    x = (torch.randn(n, t, h, w, cin, device=dev) * 0.5).to(torch.bfloat16)
    wt = torch.randn(cout, cin, 3, 1, 1, device=dev) / (cin * 3) ** 0.5
    d = ops.conv_desc(n, t, h, w, cin, cout, (3, 1, 1), (1, 1, 1), (1, 0, 0), ops.FVT_CONV_RELU)
    wp = ops.pack_conv_weight(d, wt)
    y = torch.empty(n, t, h, w, cout, device=dev, dtype=torch.bfloat16)
    res = []
    for ring in (1, 0):
        ops.set_option("disable_frame_ring", 0 if ring else 1)
        res.append(timeit(lambda: ops.conv3d_fwd(d, x, wp, out=y)))
    ops.set_option("disable_frame_ring", 0)
    gb = (x.numel() + y.numel()) * 2 / 1e9
    print("n=%d cin=%d cout=%d: ring %7.1f us (%.2f TB/s) | im2col %7.1f us (%.2f TB/s)" % (n, cin, cout, res[0], gb / res[0] * 1e3, res[1], gb / res[1] * 1e3), flush=True)
