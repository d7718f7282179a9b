"""A/B timing of single conv layers at BASELINE shapes with the epilogue partially disabled (experiments only):
tells whether a kernel is paced by its MMAs/loads or by its epilogue."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastvideotagging_b200 import ops, _lib
lib = _lib.load()
dev = torch.device("cuda:0")
for kv in os.environ.get("FVT_DBG_OPTS", "").split(","):
    if kv:
        k_, v_ = kv.split("="); assert ops.set_option(k_, int(v_)) == 0
CASES = [
    ("conv2 spatial 64->144 b48", 48, 32, 56, 56, 64, 144, (1, 3, 3), (0, 1, 1), False),
    ("conv2 temporal 144->64 b48", 48, 32, 56, 56, 144, 64, (3, 1, 1), (1, 0, 0), False),
    ("conv2 temporal+res b48", 48, 32, 56, 56, 144, 64, (3, 1, 1), (1, 0, 0), True),
    ("conv3 spatial 128->288 b48", 48, 16, 28, 28, 128, 288, (1, 3, 3), (0, 1, 1), False),
    ("conv3 temporal 288->128 b48", 48, 16, 28, 28, 288, 128, (3, 1, 1), (1, 0, 0), False),
    ("conv3 temporal 256->128 b48 (aligned)", 48, 16, 28, 28, 256, 128, (3, 1, 1), (1, 0, 0), False),
    ("conv3 temporal 320->128 b48 (aligned)", 48, 16, 28, 28, 320, 128, (3, 1, 1), (1, 0, 0), False),
    ("conv3 temporal 288->128 b48 +res", 48, 16, 28, 28, 288, 128, (3, 1, 1), (1, 0, 0), True),
    ("conv4 temporal 576->256 b48", 48, 8, 14, 14, 576, 256, (3, 1, 1), (1, 0, 0), False),
    ("conv2 dgrad-spatial 144->64 b4", 4, 32, 56, 56, 144, 64, (1, 3, 3), (0, 1, 1), False),
    ("dgrad-part 128->64 1x3x3 b4", 4, 32, 56, 56, 128, 64, (1, 3, 3), (0, 1, 1), False),
    ("dgrad-part 16->64 1x3x3 +res b4", 4, 32, 56, 56, 16, 64, (1, 3, 3), (0, 1, 1), True),
    ("conv2 spatial 64->144 b4", 4, 32, 56, 56, 64, 144, (1, 3, 3), (0, 1, 1), False),
    ("conv3 spatial 128->288 b4", 4, 16, 28, 28, 128, 288, (1, 3, 3), (0, 1, 1), False),
    ("conv3 dgrad-spatial 288->128 b4", 4, 16, 28, 28, 288, 128, (1, 3, 3), (0, 1, 1), False),
]
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
if len(sys.argv) > 1:
    ops.set_option("ring_prefetch", int(sys.argv[1]))
    CASES = [c for c in CASES if "temporal" in c[0]]
for name, n, t, h, w, cin, cout, k, p, res in CASES:
    x = (torch.randn(n, t, h, w, cin, device=dev) * 0.5).to(torch.bfloat16)
    wt = torch.randn(cout, cin, *k, device=dev) / (cin * k[0] * k[1] * k[2]) ** 0.5
    flags = ops.FVT_CONV_RELU | (ops.FVT_CONV_RESIDUAL if res else 0)
    d = ops.conv_desc(n, t, h, w, cin, cout, k, (1, 1, 1), p, flags)
    wp = ops.pack_conv_weight(d, wt)
    sc = torch.ones(cout, device=dev); sh = torch.zeros(cout, device=dev)
    y = torch.empty(n, t, h, w, cout, device=dev, dtype=torch.bfloat16)
    r = torch.randn_like(y) if res else None
    out = []
    for dbg in (0, 256, 512):
        ops.set_option("debug_flags", dbg)
        out.append(timeit(lambda: ops.conv3d_fwd(d, x, wp, sc, sh, r, out=y)))
    ops.set_option("debug_flags", 0)
    st = torch.zeros(2 * cout, device=dev)
    ds = ops.conv_desc(n, t, h, w, cin, cout, k, (1, 1, 1), p, ops.FVT_CONV_STATS)
    out.append(timeit(lambda: ops.conv3d_fwd(ds, x, wp, out=y, stats=st)))
    fl = 2.0 * n * t * h * w * cout * cin * k[0] * k[1] * k[2]
    print("%-32s full %7.1f us (%5.0f TF/s) | no-store %7.1f | no-epilogue %7.1f | train(stats) %7.1f" % (
        name, out[0], fl / out[0] / 1e6, out[1], out[2], out[3]), flush=True)
