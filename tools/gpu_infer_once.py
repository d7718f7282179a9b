"""One eager inference forward of R(2+1)D-34 at the BASELINE batch (for ncu captures of individual launches).
usage: FVT_INFER_GRAPHS=0 python tools/gpu_infer_once.py [batch]"""
import sys, os
os.environ.setdefault("FVT_INFER_GRAPHS", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_clips, oracle_params, NUM_CLASS, MODEL_DEPTH, HW, T
from fastvideotagging_b200.model import R2Plus2D
b = int(sys.argv[1]) if len(sys.argv) > 1 else 48
dev = torch.device("cuda:0")
net = R2Plus2D(NUM_CLASS, MODEL_DEPTH, final_spatial_kernel=HW // 16, final_temporal_kernel=T // 8).to(dev)
net.load_param_dict(oracle_params()); net.eval()
x = torch.from_numpy(synthetic_clips(b, seed=1)).to(dev)
with torch.no_grad():
    for _ in range(2):
        y = net(x)
torch.cuda.synchronize()
print("logits", tuple(y.shape), float(y.float().abs().max()))
