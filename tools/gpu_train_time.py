"""Times the R(2+1)D-34 training step (BASELINE configs[2]) with CUDA events; used for A/B runs through env switches
(FVT_SIDE_STREAM, FVT_CUDA_GRAPHS).  usage: gpu_train_time.py [batch] [steps] [T]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_clips, oracle_params, NUM_CLASS, MODEL_DEPTH, HW
import bench
from fastvideotagging_b200.model import R2Plus2D, SigmoidBinaryCrossEntropyLoss
from fastvideotagging_b200.trainer import Trainer
tb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
T = int(sys.argv[3]) if len(sys.argv) > 3 else 32
bench.T = T
dev = torch.device("cuda:0")
net = R2Plus2D(NUM_CLASS, MODEL_DEPTH, final_spatial_kernel=HW // 16, final_temporal_kernel=T // 8).to(dev)
net.load_param_dict(oracle_params()); net.train()
trainer = Trainer(net, "sgd", {"learning_rate": 1e-4, "momentum": 0.9, "wd": 1e-4})
xt = torch.from_numpy(synthetic_clips(tb, seed=7)).to(dev)
torch.manual_seed(0)
lab = (torch.rand(tb, NUM_CLASS, device=dev) < 0.03).float(); lab[:, 0] = 1
crit = SigmoidBinaryCrossEntropyLoss()
losses = []
def step():
    loss = crit(net(xt), lab).mean(); loss.backward(); trainer.step(tb); return loss
for i in range(4):
    losses.append(step().item())
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(steps):
    l = step()
b.record(); torch.cuda.synchronize()
print("side=%s graphs=%s batch=%d T=%d: %.3f ms/step  %.1f clips/s  losses %s -> %.6f" % (
    os.environ.get("FVT_SIDE_STREAM", "1"), os.environ.get("FVT_CUDA_GRAPHS", "1"), tb, T,
    a.elapsed_time(b) / steps, tb * steps / a.elapsed_time(b) * 1e3, ["%.6f" % v for v in losses], l.item()))
