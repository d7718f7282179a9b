"""BASELINE configs[4]: forward of the ECO-Lite 3D head (3x3x3 residual stages on 96x16x28x28 stacked trunk features),
batch 32, timed with CUDA events.  No reference code exists for ECO (model/ECO.py is empty), so only throughput and
the roofline fraction are reported.  usage: gpu_eco_head.py [batch] [steps]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastvideotagging_b200.model import ECOLite3DHead
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
torch.manual_seed(0)
head = ECOLite3DHead(101).to(dev).eval()
x = torch.rand(batch, 16, 28, 28, 96, device=dev).to(torch.bfloat16)          # NDHWC bf16, as the 2D trunk would emit
with torch.no_grad():
    for _ in range(3):
        y = head(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        y = head(x)
    b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
gf = ECOLite3DHead.conv_gflop_per_clip(16, 28)
print(json.dumps({"workload": "ECO-Lite 3D head forward, 96x16x28x28 features, batch %d (BASELINE configs[4])" % batch,
                  "clips_per_s": batch / ms * 1e3, "ms_per_step": ms, "gflop_per_clip": gf,
                  "tflops": batch * gf / ms, "frac_of_sustained_bf16_peak": batch * gf / ms / 1382.8,
                  "finite": bool(torch.isfinite(y).all())}))
