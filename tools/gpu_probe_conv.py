"""First-contact GPU probe for K1: runs a spread of conv shapes against a torch-CPU fp32 reference.
Prints one line per case; exits non-zero on any mismatch.  Run under gpurun."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from fastvideotagging_b200 import ops, _lib

torch.manual_seed(0)
dev = torch.device("cuda:0")

def ref_conv(x_ndhwc, w, stride, pad):
    x = x_ndhwc.float().permute(0, 4, 1, 2, 3).contiguous()
    y = F.conv3d(x, w.to(torch.bfloat16).float(), stride=stride, padding=pad)
    return y.permute(0, 2, 3, 4, 1).contiguous()

def run_case(name, n, t, h, w, cin, cout, k, s, p, relu=False, res=False, affine=False, stats=False, block_n=0):
    cin_s, cout_s = ops.pad16(cin), ops.pad16(cout)
    x = torch.randn(n, t, h, w, cin_s) * 0.5
    x[..., cin:] = 0
    x = x.to(torch.bfloat16)
    wt = torch.randn(cout, cin, *k) * (1.0 / (cin * k[0] * k[1] * k[2]) ** 0.5)
    flags = (ops.FVT_CONV_RELU if relu else 0) | (ops.FVT_CONV_RESIDUAL if res else 0) | (ops.FVT_CONV_STATS if stats else 0)
    d = ops.conv_desc(n, t, h, w, cin_s, cout_s, k, s, p, flags, block_n)
    to, ho, wo = ops.conv_out_shape(d)
    yref = ref_conv(x[..., :cin], wt, s, p)   # fp32 (n,to,ho,wo,cout)
    raw = yref.clone()
    scale = shift = None
    if affine:
        scale = torch.rand(cout_s) + 0.5
        shift = torch.randn(cout_s) * 0.1
        scale[cout:] = 1; shift[cout:] = 0
        yref = yref * scale[:cout] + shift[:cout]
    r = None
    if res:
        r = (torch.randn(n, to, ho, wo, cout_s) * 0.5).to(torch.bfloat16)
        r[..., cout:] = 0
        yref = yref + r[..., :cout].float()
    if relu:
        yref = yref.relu()
    xd = x.to(dev); wd = wt.to(dev)
    wp = ops.pack_conv_weight(d, wd)
    st = ops.stats_buffer(cout_s, dev) if stats else None
    t0 = time.time()
    y = ops.conv3d_fwd(d, xd, wp, scale.to(dev) if affine else None, shift.to(dev) if affine else None,
                       r.to(dev) if res else None, stats=st)
    torch.cuda.synchronize()
    dt = time.time() - t0
    yc = y.float().cpu()
    err = (yc[..., :cout] - yref).abs()
    # north-star tolerance for the bf16 path: 1e-2 relative per element; the absolute floor (values near zero) is 1e-3 of the
    # tensor's max — the kernel's only roundings are the bf16 store (2^-9 relative) and fp32 summation order
    tol = 1e-3 * yref.abs().max() + 1e-2 * yref.abs()
    bad = (err > tol)
    padbad = yc[..., cout:].abs().max().item() if cout_s > cout else 0.0
    ok = (not bad.any().item()) and padbad == 0.0
    msg = "%-28s M=%7d N=%4d K=%5d bn=%3d max_err=%.4f max_ref=%.3f pad=%.1f %s (%.1f ms)" % (
        name, n * to * ho * wo, cout_s, cin_s * k[0] * k[1] * k[2], _lib.load().fvt_conv3d_block_n(d), err.max().item(),
        yref.abs().max().item(), padbad, "OK" if ok else "FAIL", dt * 1e3)
    if stats:
        rawb = raw.to(torch.bfloat16).float()
        s1 = rawb.sum(dim=(0, 1, 2, 3)); s2 = (rawb * rawb).sum(dim=(0, 1, 2, 3))
        g = ops.stats_decode(st).cpu()
        e1 = (g[:cout] - s1).abs().max().item() / (s1.abs().max().item() + 1e-6)
        e2 = (g[cout_s:cout_s + cout] - s2).abs().max().item() / (s2.abs().max().item() + 1e-6)
        sok = e1 < 2e-2 and e2 < 2e-2
        msg += " stats rel err %.4f %.4f %s" % (e1, e2, "OK" if sok else "FAIL")
        ok = ok and sok
    print(msg, flush=True)
    if not ok and bad.any():
        idx = bad.nonzero()[:5]
        for i in idx:
            i = tuple(i.tolist())
            print("   mismatch at", i, "got", yc[i].item(), "ref", yref[i].item())
        print("   bad fraction %.4f; bad rows(first) %s" % (bad.float().mean().item(), sorted(set(bad.nonzero()[:, 3].tolist()))[:10]))
    return ok

CASES = [
    # name, n,t,h,w, cin,cout, k, s, p
    ("1x1x1 64->64 tiny", 1, 2, 8, 8, 64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    ("1x1x1 64->64 M=128", 1, 2, 8, 8, 64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    ("1x3x3 64->144", 2, 4, 14, 14, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("3x1x1 144->64", 2, 4, 14, 14, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("1x3x3 s2 64->230", 2, 4, 28, 28, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    ("3x1x1 s2 230->128", 2, 8, 14, 14, 230, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    ("1x1x1 s2 shortcut 64->128", 2, 4, 28, 28, 64, 128, (1, 1, 1), (2, 2, 2), (0, 0, 0)),
    ("1x3x3 512->1152 7x7", 2, 4, 7, 7, 512, 1152, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("3x1x1 1152->512 7x7", 2, 4, 7, 7, 1152, 512, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("stem-eq 1x7x1 32->45", 1, 2, 112, 56, 32, 45, (1, 7, 1), (1, 2, 1), (0, 3, 0)),
    ("3x1x1 45->64 stem", 1, 4, 56, 56, 45, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("3x3x3 96->128", 1, 4, 14, 14, 96, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("1x3x3 256->576 big", 4, 8, 14, 14, 256, 576, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("1x3x3 64->144 56x56 slab", 2, 8, 56, 56, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("1x3x3 128->288 28x28 slab", 2, 8, 28, 28, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("1x3x3 64->64 odd 13x9", 1, 3, 13, 9, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("1x3x3 144->64 56x56 slab", 2, 4, 56, 56, 144, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("1x3x3 176->96 28x28 slab", 1, 4, 28, 28, 176, 96, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("3x1x1 144->64 56x56 B-stat", 2, 8, 56, 56, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
]
def main():
    print("device", torch.cuda.get_device_name(0), "check", _lib.load().fvt_device_check(0), flush=True)
    allok = True
    for c in CASES:
        try:
            allok &= run_case(*c)
        except Exception as e:
            print("%-28s EXC %r" % (c[0], e), flush=True)
            allok = False
            break
    try:
        allok &= run_case("epilogue affine+res+relu", 2, 4, 14, 14, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), True, True, True)
        allok &= run_case("epilogue stats", 2, 4, 14, 14, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), False, False, False, True)
        allok &= run_case("block_n=64 on 144", 2, 4, 14, 14, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), block_n=64)
        allok &= run_case("multi-tile persistent", 8, 8, 28, 28, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1), True, True, True, True)
    except Exception as e:
        print("EXC", repr(e)); allok = False
    print("ALL OK" if allok else "SOME FAILED")
    sys.exit(0 if allok else 1)


if __name__ == "__main__":
    main()
