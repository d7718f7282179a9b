"""Guard-zone tests (compute-sanitizer is closed on this pool): every output tensor is carved out of a sentinel-filled
allocation; after the kernel the zones on both sides must be untouched.  Covers the kernels whose addressing changed
this round: BatchNorm passes, fp32 conv, row-paired stem unfold, channels-last weight gradients, the K1i TMA store."""
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 4096          # elements on each side


def _guarded(shape, dtype, device, fill=None):
    numel = 1
    for s in shape:
        numel *= s
    big = torch.empty(numel + 2 * GUARD, dtype=dtype, device=device)
    if dtype in (torch.float32,):
        big.fill_(-12345.0)
    else:
        big.fill_(-123.0)
    view = big[GUARD:GUARD + numel].view(shape)
    if fill is not None:
        view.fill_(fill)
    return big, view


def _check(big, what):
    sentinel = -12345.0 if big.dtype == torch.float32 else -123.0
    lo, hi = big[:GUARD].float(), big[-GUARD:].float()
    assert torch.all(lo == sentinel).item(), "%s: write before the tensor" % what
    assert torch.all(hi == sentinel).item(), "%s: write past the tensor" % what


def test_batchnorm_passes_stay_in_bounds(cuda_device):
    from fastvideotagging_b200 import ops
    gen = torch.Generator().manual_seed(0)
    for rows, c_real in [(777, 45), (1000, 144), (50, 921), (4099, 64)]:
        cs = ops.pad16(c_real)
        raw = torch.zeros(rows, cs)
        raw[:, :c_real] = torch.randn(rows, c_real, generator=gen)
        raw = raw.to(torch.bfloat16).to(cuda_device)
        rf = raw.float()
        stats = ops.stats_encode(torch.cat([rf.sum(0), (rf * rf).sum(0)]).contiguous())
        gamma = torch.ones(c_real, device=cuda_device)
        beta = torch.zeros(c_real, device=cuda_device)
        rm, rv = torch.zeros(c_real, device=cuda_device), torch.ones(c_real, device=cuda_device)
        outs = [_guarded((cs,), torch.float32, cuda_device) for _ in range(4)]
        big_y, y = _guarded((rows, cs), torch.bfloat16, cuda_device)
        ops.bn_finalize_apply(stats, gamma, beta, rm, rv, cs, rows, 1e-5, 0.9, *[o[1] for o in outs], raw, y, True)
        dact = torch.randn(rows, cs, generator=gen).to(torch.bfloat16).to(cuda_device)
        big_d, draw = _guarded((rows, cs), torch.bfloat16, cuda_device)
        big_z, dz = _guarded((rows, cs), torch.bfloat16, cuda_device)
        big_s, sums = _guarded((2 * cs,), torch.float32, cuda_device)
        ops.bn_backward(raw, dact, y, outs[2][1], outs[3][1], gamma, sums, draw, dz_out=dz)
        ops.bn_backward(raw, dact, None, outs[2][1], outs[3][1], gamma, sums, draw, relu_scale=outs[0][1], relu_shift=outs[1][1])
        torch.cuda.synchronize()
        for b, name in [(big_y, "bn_finalize_apply out"), (big_d, "bn_backward draw"), (big_z, "bn_backward dz"), (big_s, "bn sums")] + \
                       [(o[0], "bn_finalize_apply stat %d" % i) for i, o in enumerate(outs)]:
            _check(b, "%s rows=%d c=%d" % (name, rows, c_real))
        assert torch.isfinite(draw.float()).all()


def test_fp32_conv_and_stem_unfold_stay_in_bounds(cuda_device):
    from fastvideotagging_b200 import ops
    gen = torch.Generator().manual_seed(1)
    n, t, h, w, cin, cout = 2, 3, 9, 7, 5, 37
    d = ops.conv_desc(n, t, h, w, cin, cout, (3, 3, 3), (1, 2, 1), (1, 1, 1), ops.FVT_CONV_RELU)
    x = torch.randn(n, t, h, w, cin, generator=gen).to(cuda_device)
    wt = torch.randn(3, 3, 3, cin, cout, generator=gen).to(cuda_device)
    big, y = _guarded((n, 3, 5, 7, cout), torch.float32, cuda_device)
    ops.conv3d_fwd_f32(d, x, wt, out=y)
    torch.cuda.synchronize()
    _check(big, "conv3d_fwd_f32")
    clips = torch.rand(2, 3, 3, 10, 18, generator=gen).to(cuda_device)
    wo = (18 + 6 - 7) // 2 + 1
    big_u, u = _guarded((2, 3, 5, wo, 64), torch.bfloat16, cuda_device)
    ops.stem_unfold_hpair(clips, out=u)
    big_v, v = _guarded((2, 3, 10, wo, 32), torch.bfloat16, cuda_device)
    ops.stem_unfold(clips, out=v)
    torch.cuda.synchronize()
    _check(big_u, "stem_unfold_hpair")
    _check(big_v, "stem_unfold")
    # the row-paired unfold is the plain unfold with rows 2*h2, 2*h2+1 side by side
    assert torch.equal(u.reshape(2, 3, 5, wo, 2, 32).permute(0, 1, 2, 4, 3, 5).reshape(2, 3, 10, wo, 32), v)


def test_weight_gradients_and_tma_store_stay_in_bounds(cuda_device, lib):
    from fastvideotagging_b200 import ops
    gen = torch.Generator().manual_seed(2)
    for (n, t, h, w, cin, cout, k, s, p) in [(2, 4, 14, 14, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
                                             (2, 4, 14, 14, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
                                             (1, 4, 28, 28, 45, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
                                             (2, 4, 16, 16, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1))]:
        cin_s, cout_s = ops.pad16(cin), ops.pad16(cout)
        fwd = ops.conv_desc(n, t, h, w, cin_s, cout_s, k, s, p, 0)
        to, ho, wo = ops.conv_out_shape(fwd)
        x = torch.randn(n, t, h, w, cin_s, generator=gen).to(torch.bfloat16).to(cuda_device)
        dy = torch.randn(n, to, ho, wo, cout_s, generator=gen).to(torch.bfloat16).to(cuda_device)
        for ohwi in (False, True):
            shape = (cout, *k, cin) if ohwi else (cout, cin, *k)
            big, dw = _guarded(shape, torch.float32, cuda_device, fill=0.0)
            ops.conv3d_wgrad(fwd, x, dy, dw, cout, cin, ohwi=ohwi)
            torch.cuda.synchronize()
            _check(big, "conv3d_wgrad ohwi=%s %s" % (ohwi, (cin, cout, k, s)))
    # K1i with the TMA-store epilogue on: partial last position block (196 positions), guarded output
    n, t, h, w = 2, 5, 14, 14
    x = torch.randn(n, t, h, w, 144, generator=gen).to(torch.bfloat16).to(cuda_device)
    wt = (torch.randn(64, 144, 3, 1, 1, generator=gen) / 20).to(cuda_device)
    d = ops.conv_desc(n, t, h, w, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), ops.FVT_CONV_RELU)
    wp = ops.pack_conv_weight(d, wt)
    big, y = _guarded((n, t, h, w, 64), torch.bfloat16, cuda_device)
    assert ops.set_option("disable_tis_tma_store", 0) == 0
    try:
        ops.conv3d_fwd(d, x, wp, out=y)
        torch.cuda.synchronize()
    finally:
        ops.set_option("disable_tis_tma_store", 1)
    _check(big, "K1i TMA store")
    assert (y.float() != -123.0).any()
