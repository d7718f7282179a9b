"""Debug: run the small training step repeatedly and report which gradient elements differ between runs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import r2plus1d as orc
from fastvideotagging_b200.model import R2Plus2D, SigmoidBinaryCrossEntropyLoss
dev = torch.device("cuda:0")
from fastvideotagging_b200 import _lib
for kv in os.environ.get("FVT_DBG_OPTS", "").split(","):
    if kv:
        k, v = kv.split("="); assert ops.set_option(k, int(v)) == 0
depth, n, t, hw, eps, num_class = 10, 4, 8, 64, 10.0, 101
pool = (t // 8, hw // 16, hw // 16)
params = orc.randomize_bn(orc.init_params(depth, num_class, seed=0), seed=1)
x = np.random.default_rng(123).random((n, 3, t, hw, hw), dtype=np.float32)
labels = np.zeros((n, num_class), np.float32); labels[:, 0] = 1; labels[0, 7] = 1
runs = []
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    net = R2Plus2D(num_class, depth, final_spatial_kernel=pool[1], final_temporal_kernel=pool[0], bn_eps=eps).to(dev)
    net.load_param_dict(params); net.train()
    logits = net(torch.from_numpy(x).to(dev))
    loss = SigmoidBinaryCrossEntropyLoss()(logits, torch.from_numpy(labels).to(dev)).sum()
    loss.backward(); torch.cuda.synchronize()
    runs.append({k: getattr(net, k).grad.detach().cpu().numpy().copy() for k in net._param_names})
    # also per-run intermediate checks
    plan = list(net._train_plans.values())[0]
    del net
ref = orc.Net(params, depth, pool, eps=eps)
ref.require_grad()
rl, _ = ref.forward(x, train=True)
z = torch.from_numpy(labels)
bce = (torch.relu(rl) - rl * z + torch.log1p(torch.exp(-rl.abs()))).mean(dim=1).sum()
bce.backward()
f = {k: v.grad.numpy() for k, v in ref.p.items() if v.grad is not None}
for i, r in enumerate(runs):
    errs = {k: float(np.abs(r[k] - f[k]).max() / (np.abs(f[k]).max() + 1e-12)) for k in r}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:3]
    print("run", i, "worst vs f32:", [(k, round(e, 4)) for k, e in worst])
    k = worst[0][0]
    if worst[0][1] > 0.15:
        d = np.abs(r[k] - f[k]); idx = np.argwhere(d > 0.1 * np.abs(f[k]).max())
        print("    ", k, f[k].shape, "n_bad", len(idx), idx[:5].tolist(), idx[-2:].tolist(), "max|f|", np.abs(f[k]).max())
