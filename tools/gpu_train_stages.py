"""Per-stage timeline of the REPLAYED training step (BASELINE configs[2] / configs[3] shapes): external timing events
recorded on the main stream inside the captured forward / backward graphs (FVT_STAGE_EVENTS=1, engine.TrainPlan._mark),
read after a replay.  Side-stream work (weight packing, weight gradients) shows up only where the main stream waits for it.
usage: gpu_train_stages.py [batch] [T] [steps]"""
import sys, os
os.environ["FVT_STAGE_EVENTS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bench import synthetic_clips, oracle_params, NUM_CLASS, MODEL_DEPTH, HW
from fastvideotagging_b200.model import R2Plus2D, SigmoidBinaryCrossEntropyLoss
from fastvideotagging_b200.trainer import Trainer
tb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
bench.T = T
dev = torch.device("cuda:0")
net = R2Plus2D(NUM_CLASS, MODEL_DEPTH, final_spatial_kernel=HW // 16, final_temporal_kernel=T // 8).to(dev)
net.load_param_dict(oracle_params()); net.train()
trainer = Trainer(net, "sgd", {"learning_rate": 1e-4, "momentum": 0.9, "wd": 1e-4})
xt = torch.from_numpy(synthetic_clips(tb, seed=7)).to(dev)
lab = (torch.rand(tb, NUM_CLASS, device=dev) < 0.03).float(); lab[:, 0] = 1
crit = SigmoidBinaryCrossEntropyLoss()
def step():
    loss = crit(net(xt), lab).mean(); loss.backward(); trainer.step(tb); return loss
for _ in range(4):
    step()
plan = list(net._train_plans.values())[0]
acc = {}
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.0
for _ in range(steps):
    a.record(); step(); b.record()
    for label, ms in plan.stage_times():
        acc[label] = acc.get(label, 0.0) + ms
    tot += a.elapsed_time(b)
def group(label):
    ph, what = label.split(":")
    if what.startswith("block"):
        i = int(what[5:])
        what = "conv2_x" if i < 3 else "conv3_x" if i < 7 else "conv4_x" if i < 13 else "conv5_x"
    return ph + ":" + what
g = {}
for label, ms in acc.items():
    g[group(label)] = g.get(group(label), 0.0) + ms / steps
print("batch %d T %d: %.3f ms/step (wall between events, includes Python)" % (tb, T, tot / steps))
for k, v in g.items():
    print("  %-14s %8.3f ms" % (k, v))
print("  fwd total %.3f  bwd total %.3f" % (sum(v for k, v in g.items() if k.startswith("fwd") and k != "fwd:start"),
                                            sum(v for k, v in g.items() if k.startswith("bwd") and k != "bwd:start")))
