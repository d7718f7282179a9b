"""1x3x3 slab conv timing for several channel counts at the conv2_x geometry (experiments)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastvideotagging_b200 import ops, _lib
lib = _lib.load()
dev = torch.device("cuda:0")
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for (n, cin, cout) in [(4, 144, 64), (4, 128, 64), (4, 192, 64), (4, 64, 64), (4, 64, 144), (4, 64, 128), (4, 128, 128), (4, 64, 256)]:
    t, h, w = 32, 56, 56
    x = (torch.randn(n, t, h, w, cin, device=dev) * 0.5).to(torch.bfloat16)
    wt = torch.randn(cout, cin, 1, 3, 3, device=dev) / (cin * 9) ** 0.5
    d = ops.conv_desc(n, t, h, w, cin, cout, (1, 3, 3), (1, 1, 1), (0, 1, 1), 0)
    wp = ops.pack_conv_weight(d, wt)
    y = torch.empty(n, t, h, w, cout, device=dev, dtype=torch.bfloat16)
    us = timeit(lambda: ops.conv3d_fwd(d, x, wp, out=y))
    tiles = n * t * 28
    mmas = 9 * (cin // 16)
    floor = mmas * max(64, cout // 2)
    print("cin=%3d cout=%3d: %7.1f us  %6.0f TF/s | per tile %6.0f clk @1.85GHz, MMA floor %5d clk, %d MMAs" % (
        cin, cout, us, 2.0 * n * t * h * w * cout * cin * 9 / us / 1e6, us * 1850 / (tiles / 148.0), floor, mmas), flush=True)
