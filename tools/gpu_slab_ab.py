"""A/B of the slab kernel's load options (L2 prefetch distance, TMA boxes per slab) at BASELINE layer shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastvideotagging_b200 import ops, _lib
lib = _lib.load()
dev = torch.device("cuda:0")
CASES = [
    ("conv2 spatial 64->144 b48", 48, 32, 56, 56, 64, 144),
    ("conv3 spatial 128->288 b48", 48, 16, 28, 28, 128, 288),
    ("conv2 dgrad 144->64 b4", 4, 32, 56, 56, 144, 64),
    ("conv2 spatial 64->144 b4", 4, 32, 56, 56, 64, 144),
]
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for name, n, t, h, w, cin, cout in CASES:
    x = (torch.randn(n, t, h, w, cin, device=dev) * 0.5).to(torch.bfloat16)
    wt = torch.randn(cout, cin, 1, 3, 3, device=dev) / (cin * 9) ** 0.5
    d = ops.conv_desc(n, t, h, w, cin, cout, (1, 3, 3), (1, 1, 1), (0, 1, 1), ops.FVT_CONV_RELU)
    wp = ops.pack_conv_weight(d, wt)
    y0 = None
    line = "%-28s" % name
    for pf in (0, 1, 2, 4):
        for br in (0, 1, 2):
            ops.set_option("slab_prefetch", pf); ops.set_option("slab_box_rows", br)
            y = torch.empty(n, t, h, w, cout, device=dev, dtype=torch.bfloat16)
            us = timeit(lambda: ops.conv3d_fwd(d, x, wp, out=y))
            if y0 is None: y0 = y.clone()
            line += " | pf%d br%d %6.1f%s" % (pf, br, us, "" if torch.equal(y, y0) else " MISMATCH")
    print(line, flush=True)
ops.set_option("slab_prefetch", 2); ops.set_option("slab_box_rows", 0)
