"""Fused (2+1)D unit (K2f, one launch) against the two-launch form at the conv2_x shapes of BASELINE configs[1]/[2]:
time per unit, effective TFLOP/s, and the debug variants (no stores / no epilogue data) that locate the bound."""
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


for name, n, t, residual in (("conv2_x unit b48", 48, 32, False), ("conv2_x unit+res b48", 48, 32, True),
                             ("conv2_x unit+res b4", 4, 32, True), ("conv2_x unit+res b16 T16", 16, 16, True)):
    h = w = 56
    mid = 144
    x = (torch.randn(n, t, h, w, 64, device=dev) * 0.5).to(torch.bfloat16)
    w_s = torch.randn(mid, 64, 1, 3, 3, device=dev) / 24.0
    w_t = torch.randn(64, mid, 3, 1, 1, device=dev) / (3 * mid) ** 0.5
    sc_m, sh_m = 0.5 + torch.rand(mid, device=dev), 0.3 * torch.randn(mid, device=dev)
    sc_o, sh_o = 0.5 + torch.rand(64, device=dev), 0.3 * torch.randn(64, device=dev)
    res = torch.randn(n, t, h, w, 64, device=dev).to(torch.bfloat16) if residual else None
    d_s = ops.conv_desc(n, t, h, w, 64, mid, (1, 3, 3), (1, 1, 1), (0, 1, 1), ops.FVT_CONV_RELU)
    d_t = ops.conv_desc(n, t, h, w, mid, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), ops.FVT_CONV_RELU | (ops.FVT_CONV_RESIDUAL if residual else 0))
    wp_s, wp_t = ops.pack_conv_weight(d_s, w_s), ops.pack_conv_weight(d_t, w_t)
    y_mid = torch.empty(n, t, h, w, mid, device=dev, dtype=torch.bfloat16)
    y_two = torch.empty(n, t, h, w, 64, device=dev, dtype=torch.bfloat16)
    y_f = torch.empty_like(y_two)

    def two():
        ops.conv3d_fwd(d_s, x, wp_s, sc_m, sh_m, out=y_mid)
        ops.conv3d_fwd(d_t, y_mid, wp_t, sc_o, sh_o, res, out=y_two)

    def fused():
        ops.unit2p1_fwd(d_s, d_t, x, wp_s, sc_m, sh_m, wp_t, sc_o, sh_o, res, out=y_f)

    gflop = 2.0 * n * t * h * w * (mid * 576 + 64 * 3 * mid) / 1e9
    t2 = timeit(two)
    line = "%-26s two launches %7.1f us (%5.0f TF/s)" % (name, t2, gflop / t2 * 1e3)
    for label, opt in (("fused IS", 1), ("fused OS", 0)):
        ops.set_option("unit_input_stationary", opt)
        y_f.fill_(float("nan"))
        tf = timeit(fused)
        diff = (y_f.float() - y_two.float()).abs().max().item()
        line += " | %s %7.1f us (%5.0f TF/s) diff %.3g" % (label, tf, gflop / tf * 1e3, diff)
        for dl, bits in (("no-store", 256), ("no-epi", 512)):
            ops.set_option("debug_flags", bits)
            line += " %s %7.1f" % (dl, timeit(fused))
            ops.set_option("debug_flags", 0)
    ops.set_option("unit_input_stationary", 1)
    print(line, flush=True)
