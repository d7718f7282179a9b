import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import r2plus1d as orc
from fastvideotagging_b200.model import R2Plus2D, SigmoidBinaryCrossEntropyLoss
dev = torch.device("cuda:0")
def rel(a, b): return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
def cos(a, b): return float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
EPS = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-5
for depth, n, t, hw in ((10, 4, 8, 64), (18, 2, 8, 64)):
    pool = (t // 8, hw // 16, hw // 16)
    params = orc.randomize_bn(orc.init_params(depth, 101, seed=0), seed=1)
    x = np.random.default_rng(123).random((n, 3, t, hw, hw), dtype=np.float32)
    labels = np.zeros((n, 101), np.float32); labels[:, 0] = 1
    net = R2Plus2D(101, depth, final_spatial_kernel=pool[1], final_temporal_kernel=pool[0], bn_eps=EPS).to(dev)
    net.load_param_dict(params); net.train()
    logits = net(torch.from_numpy(x).to(dev))
    loss = SigmoidBinaryCrossEntropyLoss()(logits, torch.from_numpy(labels).to(dev)).sum()
    loss.backward(); torch.cuda.synchronize()
    res = {}
    for tag, kw in (("bf16", dict(bf16_storage=True)), ("f32", dict())):
        ref = orc.Net(params, depth, pool, eps=EPS, **kw); ref.require_grad()
        taps = {}
        rl, _ = ref.forward(x, train=True, taps=taps)
        z = torch.from_numpy(labels)
        bce = (torch.relu(rl) - rl * z + torch.log1p(torch.exp(-rl.abs()))).mean(dim=1).sum(); bce.backward()
        res[tag] = (rl.detach().numpy(), {k: v.grad.numpy() for k, v in ref.p.items() if v.grad is not None}, taps)
    lg = logits.detach().cpu().numpy()
    print("R%d n=%d t=%d hw=%d: logits kernel-vs-bf16oracle %.4f | bf16oracle-vs-f32 %.4f" % (depth, n, t, hw, rel(lg, res["bf16"][0]), rel(res["bf16"][0], res["f32"][0])))
    # per-layer activation check against the bf16 oracle taps
    plan = list(net._train_plans.values())[0]
    for name in ("conv1_middle", "conv1", "comp_0_conv_1_middle", "comp_0_conv_1", "comp_0_conv_2"):
        L = plan.layers[name]
        raw = L.raw.float().cpu()[..., :L.cout_real].permute(0, 4, 1, 2, 3).numpy()
        print("   raw %-22s rel %.5f" % (name, rel(raw, res["bf16"][2][name].detach().numpy())))
    out0 = plan.layers["comp_0_conv_2"].act.float().cpu()[..., :64].permute(0, 4, 1, 2, 3).numpy()
    print("   block0 out rel %.5f" % rel(out0, res["bf16"][2]["comp_0_out"].detach().numpy()))
    names = [k for k in net._param_names]
    worst = []
    for k in names:
        g = getattr(net, k).grad.detach().cpu().numpy()
        worst.append((rel(g, res["bf16"][1][k]), cos(g, res["bf16"][1][k]), rel(res["bf16"][1][k], res["f32"][1][k]), k))
    kf = [(rel(getattr(net, k).grad.detach().cpu().numpy(), res["f32"][1][k]), rel(res["bf16"][1][k], res["f32"][1][k]), k) for k in names]
    print("   kernel-vs-f32 median %.4f max %.4f | bf16oracle-vs-f32 median %.4f max %.4f" % (np.median([a for a, b, k in kf]), max(a for a, b, k in kf), np.median([b for a, b, k in kf]), max(b for a, b, k in kf)))
    print("   worst ratio kernel/oracle:", sorted([(a / (b + 1e-3), a, b, k) for a, b, k in kf], reverse=True)[:4])
    worst.sort(reverse=True)
    for w in worst[:6] + worst[-3:]:
        print("   grad %-36s kernel-vs-bf16oracle rel %.4f cos %.5f | bf16oracle-vs-f32 rel %.4f" % (w[3], w[0], w[1], w[2]))
    print("   median rel", float(np.median([w[0] for w in worst])))
