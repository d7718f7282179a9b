"""Grouped weight gradients (ops.WgradGroup) of the stride-1 layers of each residual stage of the R(2+1)D-34 training plan
(BASELINE configs[2]: batch 4, 32 x 112 x 112) against the same layers launched one by one.
usage: gpu_wgrad_group.py [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastvideotagging_b200 import ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
N = 4
STAGES = {  # stage: (units, t, hw, c, mid)
    "conv2_x": (6, 32, 56, 64, 144),
    "conv3_x": (7, 16, 28, 128, 288),
    "conv4_x": (11, 8, 14, 256, 576),
    "conv5_x": (5, 4, 7, 512, 1152),
}
only = os.environ.get("FVT_ONLY", "")
for kv in os.environ.get("FVT_DBG_OPTS", "").split(","):
    if kv:
        k_, v_ = kv.split("="); assert ops.set_option(k_, int(v_)) == 0
for stage, (units, t, hw, c, mid) in STAGES.items():
    if only and only not in stage:
        continue
    layers = []
    for u in range(units):
        for cin, cout, k, p in ((c, mid, (1, 3, 3), (0, 1, 1)), (mid, c, (3, 1, 1), (1, 0, 0))):
            fwd = ops.conv_desc(N, t, hw, hw, ops.pad16(cin), ops.pad16(cout), k, (1, 1, 1), p)
            x = torch.randn(N, t, hw, hw, ops.pad16(cin), device=dev).to(torch.bfloat16)
            dy = torch.randn(N, t, hw, hw, ops.pad16(cout), device=dev).to(torch.bfloat16)
            dw = torch.empty(cout, k[0], k[1], k[2], cin, device=dev)
            layers.append((fwd, x, dy, dw, cout, cin))
    group = ops.WgradGroup(layers, dev)
    def single():
        for fwd, x, dy, dw, co, ci in layers:
            ops.conv3d_wgrad(fwd, x, dy, dw, co, ci, ohwi=True)
    def timeit(fn):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3
    t1 = timeit(single) if reps > 1 else 0.0
    ref = [L[3].clone() for L in layers]
    t2 = timeit(group.run) if reps > 1 else (group.run(), 0.0)[1]
    err = max(((L[3] - r).abs().max() / r.abs().max()).item() for L, r in zip(layers, ref)) if reps > 1 else 0.0
    fl = sum(2.0 * N * t * hw * hw * L[4] * L[5] * L[0].kt * L[0].kh * L[0].kw for L in layers)
    print("%s: %d layers  one by one %8.1f us   grouped %8.1f us (grid %d, reduce blocks %d, workspace %.1f MB)  tensor roofline %6.1f us  max rel diff %.2e"
          % (stage, len(layers), t1, t2, group.grid, group.red_blocks, group.ws_bytes / 1e6, fl / 1382.8e12 * 1e6, err), flush=True)
