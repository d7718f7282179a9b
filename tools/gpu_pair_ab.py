"""CTA-pair slab kernel (one N tile per cluster, block-granular input ring) against the single-CTA streamed-filter slab kernel
on the layers that pick it by default."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastvideotagging_b200 import ops, _lib
lib = _lib.load()
dev = torch.device("cuda:0")
CASES = [
    ("conv3 spatial 128->288 b48", 48, 16, 28, 28, 128, 288, 0),
    ("conv3 spatial 128->288 b4 stats", 4, 16, 28, 28, 128, 288, 1),
    ("conv2 dgrad 144->64 b4", 4, 32, 56, 56, 144, 64, 0),
    ("conv2 dgrad 144->64 b16 T16", 16, 16, 56, 56, 144, 64, 0),
    ("conv3 temporal 288->128 b48", 48, 16, 28, 28, 288, 128, 0, (3, 1, 1), (1, 0, 0)),
    ("conv3 temporal 288->128 b4 stats", 4, 16, 28, 28, 288, 128, 1, (3, 1, 1), (1, 0, 0)),
    ("conv3 t-dgrad 128->288 b4", 4, 16, 28, 28, 128, 288, 0, (3, 1, 1), (1, 0, 0)),
    ("conv3 temporal 288->128 b16 T8", 16, 8, 28, 28, 288, 128, 0, (3, 1, 1), (1, 0, 0)),
]


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for case in CASES:
    name, n, t, h, w, cin, cout, stats = case[:8]
    kern, pad = (case[8], case[9]) if len(case) > 8 else ((1, 3, 3), (0, 1, 1))
    x = (torch.randn(n, t, h, w, cin, device=dev) * 0.5).to(torch.bfloat16)
    taps = kern[0] * kern[1] * kern[2]
    wt = torch.randn(cout, cin, *kern, device=dev) / (cin * taps) ** 0.5
    sc, sh = 0.5 + torch.rand(cout, device=dev), torch.randn(cout, device=dev)
    d = ops.conv_desc(n, t, h, w, cin, cout, kern, (1, 1, 1), pad, ops.FVT_CONV_STATS if stats else ops.FVT_CONV_RELU)
    wp = ops.pack_conv_weight(d, wt)
    y = torch.empty(n, t, h, w, cout, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(2 * cout, device=dev)
    gflop = 2.0 * n * t * h * w * cout * cin * taps / 1e9
    line = "%-34s" % name
    for mode in (0, 1):
        ops.set_option("slab_pair_auto", mode)
        fn = (lambda: ops.conv3d_fwd(d, x, wp, out=y, stats=st)) if stats else (lambda: ops.conv3d_fwd(d, x, wp, sc, sh, out=y))
        us = timeit(fn)
        line += " | %s %7.1f us (%5.0f TF/s)" % ("pair  " if mode else "single", us, gflop / us * 1e3)
        for dl, bits in (("no-epi", 512),):
            ops.set_option("debug_flags", bits)
            line += " %s %7.1f" % (dl, timeit(fn))
            ops.set_option("debug_flags", 0)
    ops.set_option("slab_pair_auto", 1)
    print(line, flush=True)

# ---- K1p: generic im2col convolution on CTA pairs (wide streamed-weight layers)
print("== K1p (fvt_set_option igemm_pair) on the conv4_x / conv5_x layers")
for name, n, t, hh, cin, cout, kern, strd, pad in (("conv4 spatial 256->576 b48", 48, 8, 14, 256, 576, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
                                                  ("conv4 temporal 576->256 b48", 48, 8, 14, 576, 256, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
                                                  ("conv5 spatial 512->1152 b48", 48, 4, 7, 512, 1152, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
                                                  ("conv5 temporal 1152->512 b48", 48, 4, 7, 1152, 512, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
                                                  ("conv4 first spatial 128->464 s2 b48", 48, 16, 28, 128, 464, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
                                                  ("conv3 first temporal 240->128 s2 b48", 48, 32, 28, 240, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
                                                  ("conv4 spatial 256->576 b16 T16", 16, 4, 14, 256, 576, (1, 3, 3), (1, 1, 1), (0, 1, 1))):
    x = (torch.randn(n, t, hh, hh, cin, device=dev) * 0.5).to(torch.bfloat16)
    taps = kern[0] * kern[1] * kern[2]
    wt = torch.randn(cout, cin, *kern, device=dev) / (cin * taps) ** 0.5
    sc, sh = 0.5 + torch.rand(cout, device=dev), torch.randn(cout, device=dev)
    d = ops.conv_desc(n, t, hh, hh, cin, cout, kern, strd, pad, ops.FVT_CONV_RELU)
    to, ho, wo = ops.conv_out_shape(d)
    wp = ops.pack_conv_weight(d, wt)
    y = torch.empty(n, to, ho, wo, cout, device=dev, dtype=torch.bfloat16)
    gflop = 2.0 * n * to * ho * wo * cout * cin * taps / 1e9
    line = "%-38s" % name
    for mode in (0, 2):
        ops.set_option("igemm_pair", mode)
        us = timeit(lambda: ops.conv3d_fwd(d, x, wp, sc, sh, out=y))
        line += " | %s %7.1f us (%5.0f TF/s)" % ("pair  " if mode else "single", us, gflop / us * 1e3)
    ops.set_option("igemm_pair", 1)
    print(line, flush=True)

# ---- layers whose filter is stationary on ONE SM (conv2_x 1x3x3 64 -> 144, the row-paired stem): single CTA vs forced pair
print("== forced pair (fvt_set_option slab_pair) on single-SM-stationary layers")
for name, n, t, cin, cout, kern, pad in (("conv2 spatial 64->144 b48", 48, 32, 64, 144, (1, 3, 3), (0, 1, 1)),
                                         ("conv2 spatial 64->144 b4", 4, 32, 64, 144, (1, 3, 3), (0, 1, 1)),
                                         ("conv2 spatial 64->144 b16 T16", 16, 16, 64, 144, (1, 3, 3), (0, 1, 1)),
                                         ("conv2 dgrad(t) 64->144... stem 64->48 b4", 4, 32, 64, 48, (1, 5, 1), (0, 2, 0))):
    h = w = 56
    x = (torch.randn(n, t, h, w, cin, device=dev) * 0.5).to(torch.bfloat16)
    taps = kern[1] * kern[2]
    wt = torch.randn(cout, cin, *kern, device=dev) / (cin * taps) ** 0.5
    sc, sh = 0.5 + torch.rand(cout, device=dev), torch.randn(cout, device=dev)
    d_inf = ops.conv_desc(n, t, h, w, cin, cout, kern, (1, 1, 1), pad, ops.FVT_CONV_RELU)
    d_trn = ops.conv_desc(n, t, h, w, cin, cout, kern, (1, 1, 1), pad, ops.FVT_CONV_STATS)
    wp = ops.pack_conv_weight(d_inf, wt)
    y = torch.empty(n, t, h, w, cout, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(2 * cout, device=dev)
    line = "%-40s" % name
    for mode in (0, 2, 1):
        ops.set_option("slab_pair", mode)
        a = timeit(lambda: ops.conv3d_fwd(d_inf, x, wp, sc, sh, out=y))
        b = timeit(lambda: ops.conv3d_fwd(d_trn, x, wp, out=y, stats=st))
        line += " | pair=%d inference %7.1f train(stats) %7.1f" % (mode, a, b)
    ops.set_option("slab_pair", 0)
    print(line, flush=True)
