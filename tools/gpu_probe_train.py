"""GPU probe for the training kernels: dgrad (via K1), wgrad (K3), zero-insert, BatchNorm fwd/bwd against torch-CPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
from fastvideotagging_b200 import ops, _lib
from oracle import r2plus1d as orc

torch.manual_seed(0)
dev = torch.device("cuda:0")
STATE = {"ok": True}

def conv_case(name, n, t, h, w, cin, cout, k, s, p):
    cin_s, cout_s = ops.pad16(cin), ops.pad16(cout)
    x = (torch.randn(n, t, h, w, cin_s) * 0.5); x[..., cin:] = 0; x = x.to(torch.bfloat16)
    wt = torch.randn(cout, cin, *k) * (1.0 / (cin * k[0] * k[1] * k[2]) ** 0.5)
    fwd = ops.conv_desc(n, t, h, w, cin_s, cout_s, k, s, p)
    to, ho, wo = ops.conv_out_shape(fwd)
    dy = (torch.randn(n, to, ho, wo, cout_s) * 0.5); dy[..., cout:] = 0; dy = dy.to(torch.bfloat16)
    # reference via autograd (fp32 on bf16-rounded operands)
    xr = x[..., :cin].float().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    wr = wt.to(torch.bfloat16).float().requires_grad_(True)
    y = F.conv3d(xr, wr, stride=s, padding=p)
    y.backward(dy[..., :cout].float().permute(0, 4, 1, 2, 3).contiguous())
    dx_ref = xr.grad.permute(0, 2, 3, 4, 1)
    dw_ref = wr.grad
    xd, dyd, wd = x.to(dev), dy.to(dev), wt.to(dev)
    # wgrad
    dw = torch.zeros(cout, cin, *k, device=dev)
    ops.conv3d_wgrad(fwd, xd, dyd, dw, cout, cin)
    torch.cuda.synchronize()
    e = (dw.cpu() - dw_ref).abs().max().item(); sc = dw_ref.abs().max().item()
    okw = e <= 1e-2 * sc + 1e-3
    # dgrad
    dd = ops.dgrad_desc(fwd)
    wp = ops.pack_conv_weight_dgrad(dd, wd)
    src = dyd if s == (1, 1, 1) else ops.zero_insert(dyd, fwd)
    dx = ops.conv3d_fwd(dd, src, wp)
    torch.cuda.synchronize()
    dxc = dx.float().cpu()[..., :cin]
    e2 = (dxc - dx_ref).abs().max().item(); sc2 = dx_ref.abs().max().item()
    okd = e2 <= 1.5e-2 * sc2 + 1e-3 and float(dx.float().cpu()[..., cin:].abs().max()) == 0.0 if cin_s > cin else e2 <= 1.5e-2 * sc2 + 1e-3
    print("%-26s wgrad err %.4f/%.3f %s | dgrad err %.4f/%.3f %s" % (name, e, sc, "OK" if okw else "FAIL", e2, sc2, "OK" if okd else "FAIL"), flush=True)
    STATE["ok"] &= bool(okw and okd)
    return bool(okw and okd)

CONV_CASES = [
    ("1x3x3 64->144", 2, 4, 14, 14, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("3x1x1 144->64", 2, 4, 14, 14, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("1x3x3 s2 64->230", 2, 4, 28, 28, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    ("3x1x1 s2 230->128", 2, 8, 14, 14, 230, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    ("1x1x1 s2 64->128", 2, 4, 28, 28, 64, 128, (1, 1, 1), (2, 2, 2), (0, 0, 0)),
    ("1x3x3 128->288", 2, 4, 14, 14, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("3x1x1 288->128", 2, 4, 14, 14, 288, 128, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("1x3x3 512->1152 7x7", 2, 2, 7, 7, 512, 1152, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("3x1x1 1152->512 7x7", 2, 2, 7, 7, 1152, 512, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("stem-eq 1x7x1 21->45", 1, 2, 56, 28, 21, 45, (1, 7, 1), (1, 2, 1), (0, 3, 0)),
    ("3x1x1 45->64 stem", 1, 4, 28, 28, 45, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("1x3x3 64->144 big", 4, 8, 56, 56, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("1x3x3 128->288 28x28", 2, 4, 28, 28, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("1x3x3 256->576 14x14", 2, 4, 14, 14, 256, 576, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("1x3x3 64->64 odd 13x9", 1, 3, 13, 9, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("1x3x3 192->80 odd 11x17", 2, 2, 11, 17, 192, 80, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
]
# ---- BatchNorm forward (stats from K1 epilogue -> finalize -> apply) and backward
def bn_case(rows_shape, c, relu, with_mask):
    cs = ops.pad16(c)
    n, t, h, w = rows_shape
    rows = n * t * h * w
    raw = (torch.randn(n, t, h, w, cs) * 1.5 + 0.3); raw[..., c:] = 0; raw = raw.to(torch.bfloat16)
    gamma = torch.rand(c) + 0.5; beta = torch.randn(c) * 0.2
    rm = torch.randn(c) * 0.1; rv = torch.rand(c) + 0.5
    x = raw[..., :c].float().permute(0, 4, 1, 2, 3).numpy().astype(np.float64)
    y, nrm, nrv, mean, inv = orc.np_batchnorm_train(x, gamma.numpy().astype(np.float64), beta.numpy().astype(np.float64),
                                                     rm.numpy().astype(np.float64), rv.numpy().astype(np.float64), 1e-5)
    if relu: y = np.maximum(y, 0)
    rawd = raw.to(dev)
    rf = rawd.float().reshape(-1, cs)
    stats = ops.stats_encode(torch.cat([rf.sum(0), (rf * rf).sum(0)]))
    scale = torch.empty(cs, device=dev); shift = torch.empty(cs, device=dev); mean_d = torch.empty(cs, device=dev); inv_d = torch.empty(cs, device=dev)
    rmd, rvd = rm.to(dev), rv.to(dev)
    gd, bd = gamma.to(dev), beta.to(dev)
    ops.bn_finalize(stats, gd, bd, rmd, rvd, cs, rows, 1e-5, 0.9, scale, shift, mean_d, inv_d)
    act = torch.empty_like(rawd)
    ops.bn_apply(rawd, scale, shift, act, relu)
    torch.cuda.synchronize()
    ya = act.float().cpu()[..., :c].permute(0, 4, 1, 2, 3).numpy()
    e1 = np.abs(ya - y).max()
    e2 = max(np.abs(rmd.cpu().numpy() - nrm).max(), np.abs(rvd.cpu().numpy() - nrv).max())
    # backward
    dact = (torch.randn(n, t, h, w, cs)); dact[..., c:] = 0; dact = dact.to(torch.bfloat16)
    g = dact[..., :c].float().permute(0, 4, 1, 2, 3).numpy().astype(np.float64)
    if with_mask: g = g * (y > 0)
    dx, dg, db = orc.np_batchnorm_backward(x, g, gamma.numpy().astype(np.float64), mean, inv)
    sums = torch.empty(2 * cs, device=dev); draw = torch.empty_like(rawd)
    ops.bn_backward(rawd, dact.to(dev), act if with_mask else None, mean_d, inv_d, gd, sums, draw)
    torch.cuda.synchronize()
    dxa = draw.float().cpu()[..., :c].permute(0, 4, 1, 2, 3).numpy()
    e3 = np.abs(dxa - dx).max() / (np.abs(dx).max() + 1e-9)
    e4 = max(np.abs(sums[:c].cpu().numpy() - dg).max() / (np.abs(dg).max() + 1e-9), np.abs(sums[cs:cs + c].cpu().numpy() - db).max() / (np.abs(db).max() + 1e-9))
    ok = e1 < 3e-2 and e2 < 1e-4 and e3 < 1.5e-2 and e4 < 2e-3
    if with_mask and relu:
        # the ReLU directly follows this BatchNorm: mask recomputed from raw must give the same result as mask = act
        sums_b = torch.empty(2 * cs, device=dev); draw_b = torch.empty_like(rawd)
        ops.bn_backward(rawd, dact.to(dev), None, mean_d, inv_d, gd, sums_b, draw_b, relu_scale=scale, relu_shift=shift)
        torch.cuda.synchronize()
        # (different masks near zero: compare with a tolerance)
        dsum = (sums_b - sums).abs().max().item() / (sums.abs().max().item() + 1e-9)
        ddraw = (draw_b.float() - draw.float()).abs().max().item() / (draw.float().abs().max().item() + 1e-9)
        same = dsum < 1e-5 and ddraw < 2 ** -7
        if not same:
            print("   self-mask variant differs from mask=act: sums rel %.3g draw rel %.3g" % (dsum, ddraw))
        ok = ok and same
    print("bn rows=%d c=%d relu=%d mask=%d: apply err %.4f running err %.2e dx rel %.4f dgamma/dbeta rel %.2e %s" % (rows, c, relu, with_mask, e1, e2, e3, e4, "OK" if ok else "FAIL"), flush=True)
    STATE["ok"] &= bool(ok)
    return bool(ok)

BN_CASES = [((2, 4, 14, 14), 144, True, True), ((2, 4, 14, 14), 64, False, False), ((1, 2, 7, 7), 1152, True, True),
            ((2, 8, 28, 28), 230, True, True), ((1, 4, 28, 28), 45, True, True)]

def main():
    for c in CONV_CASES:
        try:
            conv_case(*c)
        except Exception as ex:
            print("%-26s EXC %r" % (c[0], ex), flush=True); STATE["ok"] = False

    for shp, c, relu, m in BN_CASES:
        try:
            bn_case(shp, c, relu, m)
        except Exception as ex:
            print("bn EXC %r" % (ex,), flush=True); STATE["ok"] = False
    print("ALL OK" if STATE["ok"] else "SOME FAILED")
    sys.exit(0 if STATE["ok"] else 1)


if __name__ == "__main__":
    main()
