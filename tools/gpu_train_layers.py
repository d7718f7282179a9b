"""Per-layer device times of the training step's kernels (BASELINE configs[2] shapes), each timed alone with CUDA events:
conv fwd(+stats) | bn finalize+apply | bn backward | wgrad | dgrad (+zero-insert), against the tensor-roofline time of
the conv's FLOPs.  usage: gpu_train_layers.py [batch] [T] [csv]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bench import synthetic_clips, oracle_params, NUM_CLASS, MODEL_DEPTH, HW
from fastvideotagging_b200 import ops
from fastvideotagging_b200.model import R2Plus2D, SigmoidBinaryCrossEntropyLoss
tb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
bench.T = T
os.environ["FVT_CUDA_GRAPHS"] = "0"
os.environ["FVT_SIDE_STREAM"] = "0"
dev = torch.device("cuda:0")
from fastvideotagging_b200 import _lib
for kv in os.environ.get("FVT_DBG_OPTS", "").split(","):
    if kv:
        k_, v_ = kv.split("="); assert ops.set_option(k_, int(v_)) == 0
net = R2Plus2D(NUM_CLASS, MODEL_DEPTH, final_spatial_kernel=HW // 16, final_temporal_kernel=T // 8).to(dev)
net.load_param_dict(oracle_params()); net.train()
xt = torch.from_numpy(synthetic_clips(tb, seed=7)).to(dev)
lab = (torch.rand(tb, NUM_CLASS, device=dev) < 0.03).float(); lab[:, 0] = 1
loss = SigmoidBinaryCrossEntropyLoss()(net(xt), lab).mean(); loss.backward(); torch.cuda.synchronize()
plan = list(net._train_plans.values())[0]
PEAK = 1382.8e12

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

rows = []
tot = [0.0] * 7
print("%-28s %8s %6s %6s | %7s %7s %7s %7s %7s %7s | %7s" % ("layer", "M", "N", "K", "fwd", "bnfwd", "bnbwd", "wgrad", "dgrad", "zins", "ideal"))
for name, L in plan.layers.items():
    src = plan.bufs[L.src]
    g = torch.zeros_like(L.raw)
    g.normal_()
    draw = torch.empty_like(L.raw)
    t_fwd = timeit(lambda: ops.conv3d_fwd(L.fwd, src, L.wp, out=L.raw, stats=L.stats))
    gname, bname, mname, vname = plan._bn_names(L)
    def bnf():
        ops.bn_finalize(L.stats, plan.flat.view(plan.flat.w, gname), plan.flat.view(plan.flat.w, bname), plan.aux[mname],
                        plan.aux[vname], L.cout_s, L.rows, plan.eps, plan.momentum, L.scale, L.shift, L.mean, L.invstd)
        ops.bn_apply(L.raw, L.scale, L.shift, L.act, True)
    t_bnf = timeit(bnf)
    t_bnb = timeit(lambda: plan._bn_bwd(L, g, True, draw))
    x_in = src
    t_wg = timeit(lambda: plan._wgrad_now(L, x_in, draw))
    t_dg = t_zi = 0.0
    if L.need_dgrad:
        out = torch.empty(L.in_shape, dtype=torch.bfloat16, device=dev)
        if L.strided:
            up = plan._view("up", (L.fwd.n, L.fwd.t, L.fwd.h, L.fwd.w, L.cout_s))
            t_zi = timeit(lambda: ops.zero_insert(draw, L.fwd, out=up))
            t_dg = timeit(lambda: ops.conv3d_fwd(L.dgr, up, L.wpd, out=out))
        else:
            t_dg = timeit(lambda: ops.conv3d_fwd(L.dgr, draw, L.wpd, out=out))
    k = L.spec.kernel
    K = L.cin_real * k[0] * k[1] * k[2]
    fl = 2.0 * L.rows * L.cout_real * K
    ideal = fl / PEAK * 1e6
    vals = [t_fwd, t_bnf, t_bnb, t_wg, t_dg, t_zi, ideal]
    for i, v in enumerate(vals): tot[i] += v
    rows.append((name, L.rows, L.cout_real, K) + tuple(vals))
    print("%-28s %8d %6d %6d | %7.1f %7.1f %7.1f %7.1f %7.1f %7.1f | %7.1f" % rows[-1], flush=True)
print("%-28s %8s %6s %6s | %7.1f %7.1f %7.1f %7.1f %7.1f %7.1f | %7.1f" % (("TOTAL", "", "", "") + tuple(tot)))
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as fh:
        fh.write("layer,M,N,K,fwd_us,bnfwd_us,bnbwd_us,wgrad_us,dgrad_us,zero_insert_us,ideal_us\n")
        for r in rows:
            fh.write("%s,%d,%d,%d,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f\n" % r)
