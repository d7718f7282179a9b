"""Run-to-run difference of the SAME backward pass (fixed forward, fixed dlogits), per parameter tensor: shows where the
atomics-order noise enters and how it grows towards the input."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import r2plus1d as orc
from fastvideotagging_b200.model import R2Plus2D, SigmoidBinaryCrossEntropyLoss
from fastvideotagging_b200 import _lib
for kv in os.environ.get("FVT_DBG_OPTS", "").split(","):
    if kv:
        k, v = kv.split("="); assert ops.set_option(k, int(v)) == 0
dev = torch.device("cuda:0")
depth, n, t, hw = 34, 4, 32, 112
params = orc.randomize_bn(orc.init_params(depth, 101, seed=0), seed=1)
x = np.random.default_rng(123).random((n, 3, t, hw, hw), dtype=np.float32)
net = R2Plus2D(101, depth, final_spatial_kernel=hw // 16, final_temporal_kernel=t // 8).to(dev)
net.load_param_dict(params); net.train()
xd = torch.from_numpy(x).to(dev)
lab = torch.zeros(n, 101, device=dev); lab[:, 3] = 1
loss = SigmoidBinaryCrossEntropyLoss()(net(xd), lab).mean(); loss.backward()
plan = list(net._train_plans.values())[0]
gen = torch.Generator(device="cpu").manual_seed(5)
d1 = (torch.randn(n, 101, generator=gen) * 1e-2).to(dev)
def grads():
    plan.flat.g.zero_(); plan._backward_body(d1.contiguous()); torch.cuda.synchronize(); return plan.flat.g.clone()
a, b = grads(), grads()
rows = []
for name, (off, numel, shape, store) in plan.flat.slots.items():
    sl = slice(off, off + numel)
    e = ((a[sl] - b[sl]).norm() / (a[sl].norm() + 1e-30)).item()
    rows.append((name, e, a[sl].norm().item()))
print("whole buffer: %.3e" % ((a - b).norm() / a.norm()).item())
for r in rows[::-1]:
    if r[0].endswith("_weight"): print("%-40s noise %.3e  norm %.3e" % r)

# ---- step-by-step determinism of the last block's backward (same inputs, two executions, bitwise comparison)
print("---- last block, step by step (max abs diff between two executions / max abs value)")
B = plan.bufs
comp, xin_name, xin_shape, a_, b_, c_, d_, sc_ = plan.blocks[-1]
from fastvideotagging_b200 import ops
fl = plan.flat
def run_once():
    out = {}
    g_cur = torch.empty(plan.final_shape, dtype=torch.bfloat16, device=dev)
    fl.g.zero_()
    ops.pool_fc_bwd(d1.contiguous(), plan.pooled, fl.view(fl.w, "final_fc_weight"), fl.view(fl.g, "final_fc_weight"),
                    fl.view(fl.g, "final_fc_bias"), g_cur)
    out["g_pool"] = g_cur.clone()
    gmask = torch.empty(d_.out_shape, dtype=torch.bfloat16, device=dev)
    draw_d = torch.empty(d_.out_shape, dtype=torch.bfloat16, device=dev)
    plan._bn_bwd(d_, g_cur, d_.act, draw_d, dz_out=gmask)
    out["draw_d"] = draw_d.clone(); out["gmask"] = gmask.clone()
    gc = torch.empty(c_.out_shape, dtype=torch.bfloat16, device=dev)
    plan._dgrad(d_, draw_d, gc)
    out["gc"] = gc.clone()
    draw_c = torch.empty(c_.out_shape, dtype=torch.bfloat16, device=dev)
    plan._bn_bwd(c_, gc, True, draw_c)
    out["draw_c"] = draw_c.clone()
    gb = torch.empty(b_.out_shape, dtype=torch.bfloat16, device=dev)
    plan._dgrad(c_, draw_c, gb)
    out["gb"] = gb.clone()
    draw_b = torch.empty(b_.out_shape, dtype=torch.bfloat16, device=dev)
    plan._bn_bwd(b_, gb, True, draw_b)
    out["draw_b"] = draw_b.clone()
    out["sums_c"] = fl.g[fl.slots[c_.spec.bn + "_gamma"][0]:fl.slots[c_.spec.bn + "_gamma"][0] + 2 * c_.cout_s].clone()
    torch.cuda.synchronize()
    return out
r1, r2 = run_once(), run_once()
for k in r1:
    x1, x2 = r1[k].float(), r2[k].float()
    print("%-8s %.3e  (rel L2 %.3e, nonfinite %d)" % (k, ((x1 - x2).abs().max() / (x1.abs().max() + 1e-30)).item(),
                                                   ((x1 - x2).norm() / (x1.norm() + 1e-30)).item(), int((~torch.isfinite(x1)).sum())))
