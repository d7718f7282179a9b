"""Runs a few R(2+1)D-34 training steps (BASELINE configs[2] shape) — the command profiled for profiles/*train*."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_clips, oracle_params, NUM_CLASS, MODEL_DEPTH, T, HW
from fastvideotagging_b200.model import R2Plus2D, SigmoidBinaryCrossEntropyLoss
from fastvideotagging_b200.trainer import Trainer
tb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
net = R2Plus2D(NUM_CLASS, MODEL_DEPTH, final_spatial_kernel=HW // 16, final_temporal_kernel=T // 8).to(dev)
net.load_param_dict(oracle_params()); net.train()
trainer = Trainer(net, "sgd", {"learning_rate": 1e-4, "momentum": 0.9, "wd": 1e-4})
xt = torch.from_numpy(synthetic_clips(tb, seed=7)).to(dev)
lab = (torch.rand(tb, NUM_CLASS, device=dev) < 0.03).float(); lab[:, 0] = 1
crit = SigmoidBinaryCrossEntropyLoss()
for i in range(steps):
    loss = crit(net(xt), lab).mean(); loss.backward(); trainer.step(tb)
torch.cuda.synchronize()
print("loss", loss.item())
