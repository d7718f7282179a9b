"""Weight-gradient kernels on the distinct layer shapes of the R(2+1)D-34 training plan (BASELINE configs[2]: batch 4,
32 x 112 x 112), each timed alone with CUDA events against its tensor-roofline and HBM-roofline time.
usage: gpu_wgrad_shapes.py [reps]      (reps = 1: one launch per shape, for ncu)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastvideotagging_b200 import ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
PEAK, HBM = 1382.8e12, 6.5e12
N = 4
SHAPES = [  # name, t, h, w, cin, cout, kernel, stride, pad
    ("stem 1x7x7 (row-paired 1x5x1) 64->45", 32, 56, 56, 64, 45, (1, 5, 1), (1, 1, 1), (0, 2, 0)),
    ("stem 3x1x1 45->64", 32, 56, 56, 45, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("conv2_x 1x3x3 64->144", 32, 56, 56, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("conv2_x 3x1x1 144->64", 32, 56, 56, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("conv3_x 1x3x3/s2 64->230", 32, 56, 56, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    ("conv3_x 3x1x1/s2 230->128", 32, 28, 28, 230, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    ("conv3_x 1x3x3 128->288", 16, 28, 28, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("conv3_x 3x1x1 288->128", 16, 28, 28, 288, 128, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("conv4_x 1x3x3/s2 128->460", 16, 28, 28, 128, 460, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    ("conv4_x 3x1x1/s2 460->256", 16, 14, 14, 460, 256, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    ("conv4_x 1x3x3 256->576", 8, 14, 14, 256, 576, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("conv4_x 3x1x1 576->256", 8, 14, 14, 576, 256, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("conv5_x 1x3x3/s2 256->921", 8, 14, 14, 256, 921, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    ("conv5_x 3x1x1/s2 921->512", 8, 7, 7, 921, 512, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    ("conv5_x 1x3x3 512->1152", 4, 7, 7, 512, 1152, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("conv5_x 3x1x1 1152->512", 4, 7, 7, 1152, 512, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("shortcut 1x1x1/s2 64->128", 32, 56, 56, 64, 128, (1, 1, 1), (2, 2, 2), (0, 0, 0)),
    ("shortcut 1x1x1/s2 256->512", 8, 14, 14, 256, 512, (1, 1, 1), (2, 2, 2), (0, 0, 0)),
]
only = os.environ.get("FVT_ONLY", "")
for kv in os.environ.get("FVT_DBG_OPTS", "").split(","):
    if kv:
        k_, v_ = kv.split("="); assert ops.set_option(k_, int(v_)) == 0
tot = tot_ideal = 0.0
for name, t, h, w, cin, cout, k, s, p in SHAPES:
    if only and only not in name:
        continue
    cin_s, cout_s = ops.pad16(cin), ops.pad16(cout)
    fwd = ops.conv_desc(N, t, h, w, cin_s, cout_s, k, s, p)
    to, ho, wo = ops.conv_out_shape(fwd)
    x = torch.randn(N, t, h, w, cin_s, device=dev).to(torch.bfloat16)
    dy = torch.randn(N, to, ho, wo, cout_s, device=dev).to(torch.bfloat16)
    dw = torch.empty(cout, k[0], k[1], k[2], cin, device=dev)
    def run():
        ops.conv3d_wgrad(fwd, x, dy, dw, cout, cin, ohwi=True)
    run(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if reps > 1:
        a.record()
        for _ in range(reps): run()
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / reps * 1e3
    else:
        us = 0.0
    rows = N * to * ho * wo
    fl = 2.0 * rows * cout * cin * k[0] * k[1] * k[2]
    by = (x.numel() + dy.numel()) * 2 + dw.numel() * 4
    ideal = max(fl / PEAK, by / HBM) * 1e6
    tot += us; tot_ideal += ideal
    print("%-36s rows %7d  %8.1f us   tensor %6.1f us  hbm %6.1f us  -> %4.2f of the roofline" % (
        name, rows, us, fl / PEAK * 1e6, by / HBM * 1e6, ideal / us if us else 0.0), flush=True)
print("sum %.1f us   rooflines %.1f us" % (tot, tot_ideal))
