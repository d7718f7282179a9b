"""Teacher-forced per-kernel parity on the GPU: every kernel is fed identical (bf16-rounded) inputs as a torch-CPU /
numpy-oracle evaluation of the same operator and must agree within the bf16 tolerance (rel 1e-2 of the tensor's max).
This is the strict gate for the training path: end-to-end training gradients of a BatchNorm network are chaotic under
1-ulp bf16 perturbations ('conditioning'), so kernel correctness is established here, per kernel."""
import pytest

pytestmark = pytest.mark.gpu


def _probe_conv():
    from tools import gpu_probe_conv
    return gpu_probe_conv


def _probe_train():
    from tools import gpu_probe_train
    return gpu_probe_train


@pytest.mark.parametrize("idx", range(19))
def test_conv_forward_shapes(cuda_device, idx):
    m = _probe_conv()
    assert m.run_case(*m.CASES[idx])


def test_conv_forward_epilogues(cuda_device):
    m = _probe_conv()
    base = (2, 4, 14, 14, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    assert m.run_case("affine+res+relu", *base, True, True, True)
    assert m.run_case("stats", *base, False, False, False, True)
    assert m.run_case("block_n=64", *base, block_n=64)
    # FVT_CONV_STATS describes the RAW output (training forward): it comes without affine / residual / ReLU
    assert m.run_case("multi-tile persistent", 8, 8, 28, 28, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1), True, True, True)
    assert m.run_case("multi-tile persistent, stats", 8, 8, 28, 28, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1), False, False, False, True)
    assert m.run_case("ragged M tail", 1, 3, 7, 9, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), True, False, True)
    assert m.run_case("ragged M tail, stats", 1, 3, 7, 9, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), False, False, False, True)


def test_slab_and_im2col_kernels_agree(cuda_device, lib):
    """K1s (slab) and K1 (im2col) are two schedules of the same convolution: identical inputs -> results equal up to
    fp32 summation order (<= 1 bf16 ulp)."""
    import torch
    from fastvideotagging_b200 import ops
    torch.manual_seed(3)
    x = (torch.randn(2, 4, 28, 28, 64) * 0.5).to(torch.bfloat16).to(cuda_device)
    w = torch.randn(144, 64, 1, 3, 3, device=cuda_device) / 24.0
    d = ops.conv_desc(2, 4, 28, 28, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), ops.FVT_CONV_RELU)
    wp = ops.pack_conv_weight(d, w)
    y_slab = ops.conv3d_fwd(d, x, wp)
    assert ops.set_option("disable_slab", 1) == 0
    try:
        y_gen = ops.conv3d_fwd(d, x, wp)
    finally:
        ops.set_option("disable_slab", 0)
    torch.cuda.synchronize()
    diff = (y_slab.float() - y_gen.float()).abs().max().item()
    assert diff <= 2 ** -7 * y_gen.float().abs().max().item()


def test_stationary_and_streamed_weights_agree(cuda_device, lib):
    """K1 with the filter resident in shared memory vs. streamed per tile: same MMAs in the same order -> bit-identical."""
    import torch
    from fastvideotagging_b200 import ops
    torch.manual_seed(4)
    x = (torch.randn(2, 8, 56, 56, 144) * 0.5).to(torch.bfloat16).to(cuda_device)
    w = torch.randn(64, 144, 3, 1, 1, device=cuda_device) / 20.0
    d = ops.conv_desc(2, 8, 56, 56, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), ops.FVT_CONV_RELU)
    wp = ops.pack_conv_weight(d, w)
    y_stat = ops.conv3d_fwd(d, x, wp)
    assert ops.set_option("disable_b_stationary", 1) == 0
    try:
        y_str = ops.conv3d_fwd(d, x, wp)
    finally:
        ops.set_option("disable_b_stationary", 0)
    torch.cuda.synchronize()
    assert torch.equal(y_stat, y_str)


@pytest.mark.parametrize("shape", [(2, 8, 56, 56, 64, 64), (1, 16, 28, 28, 64, 48), (3, 5, 14, 14, 48, 64), (1, 32, 20, 13, 32, 64),
                                   (2, 8, 56, 56, 144, 64), (3, 4, 28, 28, 144, 64), (1, 32, 20, 13, 128, 64), (5, 3, 14, 14, 96, 128),
                                   (2, 1, 28, 28, 144, 64), (2, 2, 28, 28, 80, 48)])
def test_frame_ring_and_im2col_kernels_agree(cuda_device, lib, shape):
    """K1t (frame ring, one channel block) / K1i (input-stationary, several channel blocks) and K1 (im2col) are
    schedules of the same 3x1x1 convolution (with residual + ReLU + statistics): equal up to fp32 summation order."""
    import torch
    from fastvideotagging_b200 import ops
    n, t, h, w_, cin, cout = shape
    torch.manual_seed(6)
    x = (torch.randn(n, t, h, w_, cin) * 0.5).to(torch.bfloat16).to(cuda_device)
    res = (torch.randn(n, t, h, w_, cout) * 0.5).to(torch.bfloat16).to(cuda_device)
    w = torch.randn(cout, cin, 3, 1, 1, device=cuda_device) / (3 * cin) ** 0.5
    sc = (torch.rand(cout, device=cuda_device) + 0.5)
    sh = torch.randn(cout, device=cuda_device) * 0.1
    d = ops.conv_desc(n, t, h, w_, cin, cout, (3, 1, 1), (1, 1, 1), (1, 0, 0), ops.FVT_CONV_RELU | ops.FVT_CONV_RESIDUAL)
    d_st = ops.conv_desc(n, t, h, w_, cin, cout, (3, 1, 1), (1, 1, 1), (1, 0, 0), ops.FVT_CONV_STATS)     # raw output + statistics
    wp = ops.pack_conv_weight(d, w)
    st_ring = ops.stats_buffer(cout, cuda_device)
    y_ring = ops.conv3d_fwd(d, x, wp, sc, sh, res)
    r_ring = ops.conv3d_fwd(d_st, x, wp, stats=st_ring)
    assert ops.set_option("disable_frame_ring", 1) == 0 and ops.set_option("disable_temporal_is", 1) == 0
    try:
        st_gen = ops.stats_buffer(cout, cuda_device)
        y_gen = ops.conv3d_fwd(d, x, wp, sc, sh, res)
        r_gen = ops.conv3d_fwd(d_st, x, wp, stats=st_gen)
    finally:
        ops.set_option("disable_frame_ring", 0)
        ops.set_option("disable_temporal_is", 0)
    torch.cuda.synchronize()
    scale = y_gen.float().abs().max().item()
    assert (y_ring.float() - y_gen.float()).abs().max().item() <= 2 ** -7 * scale
    assert (r_ring.float() - r_gen.float()).abs().max().item() <= 2 ** -7 * r_gen.float().abs().max().item()
    st_ring, st_gen = ops.stats_decode(st_ring), ops.stats_decode(st_gen)
    assert (st_ring - st_gen).abs().max().item() <= 1e-3 * st_gen.abs().max().item()
    # the statistics are those of the stored (bf16-rounded) raw output
    rf = r_ring.float().reshape(-1, cout)
    ref = torch.cat([rf.sum(0), (rf * rf).sum(0)])
    assert (st_ring - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()


@pytest.mark.parametrize("shape", [(2, 4, 7, 7, 512, 1152, (1, 3, 3), (0, 1, 1)), (4, 4, 7, 7, 1152, 512, (3, 1, 1), (1, 0, 0)),
                                   (1, 8, 14, 14, 576, 256, (3, 1, 1), (1, 0, 0)), (1, 3, 5, 9, 256, 80, (1, 3, 3), (0, 1, 1))])
def test_split_k_matches_single_pass(cuda_device, lib, shape):
    """Small-M convolutions split their reduction over several CTAs (fp32 partial tiles in per-split slices of the
    caller's workspace + one finalize pass that adds the slices in split order).  Same result as the single-pass kernel up
    to fp32 summation order, for both epilogue flavours; the workspace may hold garbage on entry; two runs agree bit for bit."""
    import torch
    from fastvideotagging_b200 import ops
    n, t, h, w_, cin, cout, k, pad = shape
    torch.manual_seed(8)
    x = (torch.randn(n, t, h, w_, cin) * 0.5).to(torch.bfloat16).to(cuda_device)
    res = (torch.randn(n, t, h, w_, cout) * 0.5).to(torch.bfloat16).to(cuda_device)
    w = torch.randn(cout, cin, *k, device=cuda_device) / (cin * k[0] * k[1] * k[2]) ** 0.5
    sc = torch.rand(cout, device=cuda_device) + 0.5
    sh = torch.randn(cout, device=cuda_device) * 0.1
    outs = {}
    for split in (1, 0):
        assert ops.set_option("disable_split_k", 0 if split else 1) == 0
        try:
            d1 = ops.conv_desc(n, t, h, w_, cin, cout, k, (1, 1, 1), pad, ops.FVT_CONV_RELU | ops.FVT_CONV_RESIDUAL)
            wp = ops.pack_conv_weight(d1, w)
            y1 = ops.conv3d_fwd(d1, x, wp, sc, sh, res)
            d2 = ops.conv_desc(n, t, h, w_, cin, cout, k, (1, 1, 1), pad, ops.FVT_CONV_STATS)
            ops.workspace(cuda_device).fill_(float("nan"))          # contents are irrelevant on entry
            st = ops.stats_buffer(cout, cuda_device)
            y2 = ops.conv3d_fwd(d2, x, wp, stats=st)
            st_b = ops.stats_buffer(cout, cuda_device)
            y2_b = ops.conv3d_fwd(d2, x, wp, stats=st_b)
            torch.cuda.synchronize()
            assert torch.equal(y2, y2_b) and torch.equal(st, st_b)      # deterministic, statistics included
            outs[split] = (y1.float(), y2.float(), ops.stats_decode(st))
        finally:
            ops.set_option("disable_split_k", 0)
    for a, b in zip(outs[1][:2], outs[0][:2]):
        assert (a - b).abs().max().item() <= 2 ** -7 * b.abs().max().item()
    assert (outs[1][2] - outs[0][2]).abs().max().item() <= 2e-3 * outs[0][2].abs().max().item()


def test_wgrad_slab_and_im2col_kernels_agree(cuda_device, lib):
    """K3s (slab, stacked taps) and K3 (im2col) compute the same weight gradient up to fp32 summation order."""
    import torch
    from fastvideotagging_b200 import ops
    torch.manual_seed(5)
    for (n, t, h, w_, cin, cout) in ((2, 4, 56, 56, 64, 144), (2, 4, 28, 28, 128, 288), (1, 2, 7, 7, 512, 1152)):
        x = (torch.randn(n, t, h, w_, cin) * 0.5).to(torch.bfloat16).to(cuda_device)
        dy = (torch.randn(n, t, h, w_, cout) * 0.5).to(torch.bfloat16).to(cuda_device)
        d = ops.conv_desc(n, t, h, w_, cin, cout, (1, 3, 3), (1, 1, 1), (0, 1, 1))
        dw_slab = torch.zeros(cout, cin, 1, 3, 3, device=cuda_device)
        ops.conv3d_wgrad(d, x, dy, dw_slab, cout, cin)
        assert ops.set_option("disable_wgrad_slab", 1) == 0
        try:
            dw_gen = torch.zeros_like(dw_slab)
            ops.conv3d_wgrad(d, x, dy, dw_gen, cout, cin)
        finally:
            ops.set_option("disable_wgrad_slab", 0)
        torch.cuda.synchronize()
        scale = dw_gen.abs().max().item()
        assert (dw_slab - dw_gen).abs().max().item() <= 2e-4 * scale + 1e-4, (h, cin, cout)


@pytest.mark.parametrize("shape", [(2, 8, 56, 56, 144, 64), (2, 6, 28, 28, 288, 128), (1, 4, 14, 14, 576, 256), (3, 5, 20, 13, 48, 64)])
def test_wgrad_temporal_slab_and_im2col_kernels_agree(cuda_device, lib, shape):
    """K3s in temporal mode (one tap per CTA, channel blocks stacked in M) vs K3 (im2col) on 3x1x1 convs."""
    import torch
    from fastvideotagging_b200 import ops
    n, t, h, w_, cin, cout = shape
    torch.manual_seed(7)
    x = (torch.randn(n, t, h, w_, cin) * 0.5).to(torch.bfloat16).to(cuda_device)
    dy = (torch.randn(n, t, h, w_, cout) * 0.5).to(torch.bfloat16).to(cuda_device)
    d = ops.conv_desc(n, t, h, w_, cin, cout, (3, 1, 1), (1, 1, 1), (1, 0, 0))
    cin_r = cin - 3 if cin == 48 else cin                     # stem-like: 45 real channels in 48 stored
    dw_slab = torch.zeros(cout, cin_r, 3, 1, 1, device=cuda_device)
    ops.conv3d_wgrad(d, x, dy, dw_slab, cout, cin_r)
    assert ops.set_option("disable_wgrad_slab", 1) == 0
    try:
        dw_gen = torch.zeros_like(dw_slab)
        ops.conv3d_wgrad(d, x, dy, dw_gen, cout, cin_r)
    finally:
        ops.set_option("disable_wgrad_slab", 0)
    torch.cuda.synchronize()
    scale = dw_gen.abs().max().item()
    assert (dw_slab - dw_gen).abs().max().item() <= 2e-4 * scale + 1e-4


@pytest.mark.parametrize("idx", range(16))
def test_conv_wgrad_dgrad(cuda_device, idx):
    m = _probe_train()
    assert m.conv_case(*m.CONV_CASES[idx])


@pytest.mark.parametrize("idx", range(5))
def test_batchnorm_forward_backward(cuda_device, idx):
    m = _probe_train()
    assert m.bn_case(*m.BN_CASES[idx])


@pytest.mark.parametrize("rows,c_real,res_mode", [(4096, 144, 0), (1000, 64, 1), (777, 230, 2), (50176, 288, 0)])
def test_bn_finalize_apply_matches_the_two_pass_form(cuda_device, lib, rows, c_real, res_mode):
    """fvt_bn_finalize_apply (one launch) == fvt_bn_finalize + fvt_bn_apply, bit for bit: outputs, scale/shift/mean/invstd
    and the running statistics (MXNet momentum convention)."""
    import torch
    from fastvideotagging_b200 import ops
    gen = torch.Generator().manual_seed(rows + c_real)
    cs = ops.pad16(c_real)
    raw = torch.zeros(rows, cs)
    raw[:, :c_real] = torch.randn(rows, c_real, generator=gen) * 2 + 0.5
    raw = raw.to(torch.bfloat16).to(cuda_device)
    rf = raw.float()
    stats = ops.stats_encode(torch.cat([rf.sum(0), (rf * rf).sum(0)]).contiguous())
    gamma = (0.5 + torch.rand(c_real, generator=gen)).to(cuda_device)
    beta = torch.randn(c_real, generator=gen).to(cuda_device)
    res = torch.randn(rows, cs, generator=gen).to(torch.bfloat16).to(cuda_device) if res_mode else None
    rs = torch.rand(cs, generator=gen).to(cuda_device) if res_mode == 2 else None
    rh = torch.randn(cs, generator=gen).to(cuda_device) if res_mode == 2 else None

    def fresh():
        return (torch.full((c_real,), 0.25, device=cuda_device), torch.full((c_real,), 1.5, device=cuda_device),
                [torch.empty(cs, device=cuda_device) for _ in range(4)], torch.empty_like(raw))

    rm1, rv1, o1, y1 = fresh()
    ops.bn_finalize(stats, gamma, beta, rm1, rv1, cs, rows, 1e-5, 0.9, *o1)
    ops.bn_apply(raw, o1[0], o1[1], y1, True, res=res, res_scale=rs, res_shift=rh)
    rm2, rv2, o2, y2 = fresh()
    ops.bn_finalize_apply(stats, gamma, beta, rm2, rv2, cs, rows, 1e-5, 0.9, *o2, raw, y2, True, res=res, res_scale=rs, res_shift=rh)
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)
    for a, b in zip(o1, o2):
        assert torch.equal(a, b)
    assert torch.equal(rm1, rm2) and torch.equal(rv1, rv2)
    # and against the definition (biased variance, reference semantics A4)
    mean = rf[:, :c_real].double().mean(0)
    var = rf[:, :c_real].double().var(0, unbiased=False)
    assert torch.allclose(o2[2][:c_real].double().cpu(), mean.cpu(), rtol=1e-4, atol=1e-4)
    assert torch.allclose(rv2.double().cpu(), (0.9 * 1.5 + 0.1 * var).cpu(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("shape", [(2, 4, 14, 14, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
                                   (2, 4, 14, 14, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
                                   (2, 4, 28, 28, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
                                   (2, 8, 14, 14, 230, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
                                   (1, 4, 56, 56, 45, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
                                   (2, 4, 7, 7, 512, 1152, (1, 3, 3), (1, 1, 1), (0, 1, 1))])
def test_channels_last_weight_layout_matches_reference_layout(cuda_device, lib, shape):
    """FVT_CONV_W_OHWI: packing from, and weight gradients into, (O, kT, kH, kW, I) storage give the same numbers as the
    reference's (O, I, kT, kH, kW) layout (packing bit for bit; gradients up to the fp32 atomics order)."""
    import torch
    from fastvideotagging_b200 import ops
    n, t, h, w, cin, cout, k, s, p = shape
    gen = torch.Generator().manual_seed(cin * 7 + cout)
    cin_s, cout_s = ops.pad16(cin), ops.pad16(cout)
    wt = (torch.randn(cout, cin, *k, generator=gen) / (cin * k[0] * k[1] * k[2]) ** 0.5).to(cuda_device)
    wt_ohwi = wt.permute(0, 2, 3, 4, 1).contiguous()
    fwd = ops.conv_desc(n, t, h, w, cin_s, cout_s, k, s, p, 0)
    assert torch.equal(ops.pack_conv_weight(fwd, wt), ops.pack_conv_weight(fwd, wt_ohwi, ohwi=True))
    dgr = ops.dgrad_desc(fwd)
    assert torch.equal(ops.pack_conv_weight_dgrad(dgr, wt), ops.pack_conv_weight_dgrad(dgr, wt_ohwi, ohwi=True))
    x = torch.zeros(n, t, h, w, cin_s)
    x[..., :cin] = torch.randn(n, t, h, w, cin, generator=gen)
    x = x.to(torch.bfloat16).to(cuda_device)
    to, ho, wo = ops.conv_out_shape(fwd)
    dy = torch.zeros(n, to, ho, wo, cout_s)
    dy[..., :cout] = torch.randn(n, to, ho, wo, cout, generator=gen)
    dy = dy.to(torch.bfloat16).to(cuda_device)
    dw_a = torch.zeros(cout, cin, *k, device=cuda_device)
    dw_b = torch.zeros(cout, *k, cin, device=cuda_device)
    ops.conv3d_wgrad(fwd, x, dy, dw_a, cout, cin)
    ops.conv3d_wgrad(fwd, x, dy, dw_b, cout, cin, ohwi=True)
    torch.cuda.synchronize()
    ref = dw_a.permute(0, 2, 3, 4, 1)
    assert (dw_b - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-6
    assert dw_b.abs().max().item() > 0


def test_temporal_is_tma_store_epilogue_matches_register_stores(cuda_device, lib):
    """K1i's optional TMA-store epilogue (off by default: measured slower) writes bit-identical outputs, including the
    clipped partial position block (14*14 = 196 positions) and the residual + ReLU + affine epilogue."""
    import torch
    from fastvideotagging_b200 import ops
    gen = torch.Generator().manual_seed(11)
    for (n, t, h, w) in [(2, 6, 14, 14), (1, 4, 56, 56)]:
        x = (torch.randn(n, t, h, w, 144, generator=gen) * 0.5).to(torch.bfloat16).to(cuda_device)
        wt = (torch.randn(64, 144, 3, 1, 1, generator=gen) / 432 ** 0.5).to(cuda_device)
        res = torch.randn(n, t, h, w, 64, generator=gen).to(torch.bfloat16).to(cuda_device)
        sc = (0.5 + torch.rand(64, generator=gen)).to(cuda_device)
        sh = torch.randn(64, generator=gen).to(cuda_device)
        d = ops.conv_desc(n, t, h, w, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), ops.FVT_CONV_RELU | ops.FVT_CONV_RESIDUAL)
        wp = ops.pack_conv_weight(d, wt)
        a = ops.conv3d_fwd(d, x, wp, sc, sh, res).clone()
        assert ops.set_option("disable_tis_tma_store", 0) == 0
        try:
            b = ops.conv3d_fwd(d, x, wp, sc, sh, res).clone()
        finally:
            ops.set_option("disable_tis_tma_store", 1)
        torch.cuda.synchronize()
        assert torch.equal(a, b)
        assert a.float().abs().max().item() > 0


@pytest.mark.parametrize("shape", [(2, 4, 56, 56, 64, 144, (1, 3, 3), (0, 1, 1)),      # conv2_x 1x3x3
                                   (1, 3, 56, 56, 64, 144, (1, 3, 3), (0, 1, 1)),      # odd number of tiles: one dummy tile
                                   (1, 2, 14, 14, 64, 144, (1, 3, 3), (0, 1, 1)),      # tiles with a clipped bottom row
                                   (2, 3, 56, 56, 64, 48, (1, 5, 1), (0, 2, 0))])      # the row-paired stem conv
def test_cta_pair_slab_kernel_matches_single_cta(cuda_device, lib, shape):
    """K1s2 (tcgen05.mma.cta_group::2 over a CTA pair, staged TMA store) == K1s on one CTA, bit for bit: plain, with the
    affine + residual + ReLU epilogue, and with BatchNorm statistics."""
    import torch
    from fastvideotagging_b200 import ops
    n, t, h, w, cin, cout, k, p = shape
    gen = torch.Generator().manual_seed(n * 100 + h)
    x = (torch.randn(n, t, h, w, cin, generator=gen) * 0.5).to(torch.bfloat16).to(cuda_device)
    wt = (torch.randn(cout, cin, *k, generator=gen) / (cin * k[1] * k[2]) ** 0.5).to(cuda_device)
    res = torch.randn(n, t, h, w, cout, generator=gen).to(torch.bfloat16).to(cuda_device)
    sc = (0.5 + torch.rand(cout, generator=gen)).to(cuda_device)
    sh = torch.randn(cout, generator=gen).to(cuda_device)
    d_plain = ops.conv_desc(n, t, h, w, cin, cout, k, (1, 1, 1), p, 0)
    d_full = ops.conv_desc(n, t, h, w, cin, cout, k, (1, 1, 1), p, ops.FVT_CONV_RELU | ops.FVT_CONV_RESIDUAL)
    d_stat = ops.conv_desc(n, t, h, w, cin, cout, k, (1, 1, 1), p, ops.FVT_CONV_STATS)
    wp = ops.pack_conv_weight(d_plain, wt)

    def run_all():
        a = ops.conv3d_fwd(d_plain, x, wp).clone()
        b = ops.conv3d_fwd(d_full, x, wp, sc, sh, res).clone()
        st = ops.stats_buffer(cout, cuda_device)
        c = ops.conv3d_fwd(d_stat, x, wp, stats=st).clone()
        torch.cuda.synchronize()
        return a, b, c, ops.stats_decode(st)

    out = {}
    try:
        for mode in (0, 1, 2):
            assert ops.set_option("slab_pair", mode) == 0
            out[mode] = run_all()
    finally:
        ops.set_option("slab_pair", 0)
    for mode in (1, 2):
        for i in range(3):
            assert torch.equal(out[0][i], out[mode][i]), (mode, i)
        ref = out[0][3]
        assert (out[mode][3] - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-3     # fp32 atomics order
    assert out[0][0].float().abs().max().item() > 0


def _unit_case(device, n, t, h, w, mid, seed, residual, kernel=(1, 3, 3)):
    import torch
    from fastvideotagging_b200 import ops
    gen = torch.Generator().manual_seed(seed)
    x = (torch.randn(n, t, h, w, 64, generator=gen) * 0.5).to(torch.bfloat16).to(device)
    w_s = (torch.randn(mid, 64, *kernel, generator=gen) / (64 * kernel[1] * kernel[2]) ** 0.5).to(device)
    w_t = (torch.randn(64, mid, 3, 1, 1, generator=gen) / (3 * mid) ** 0.5).to(device)
    sc_m = (0.5 + torch.rand(mid, generator=gen)).to(device)
    sh_m = (0.3 * torch.randn(mid, generator=gen)).to(device)
    sc_o = (0.5 + torch.rand(64, generator=gen)).to(device)
    sh_o = (0.3 * torch.randn(64, generator=gen)).to(device)
    res = torch.randn(n, t, h, w, 64, generator=gen).to(torch.bfloat16).to(device) if residual else None
    d_s = ops.conv_desc(n, t, h, w, 64, mid, kernel, (1, 1, 1), (0, kernel[1] // 2, kernel[2] // 2), ops.FVT_CONV_RELU)
    d_t = ops.conv_desc(n, t, h, w, mid, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0),
                        ops.FVT_CONV_RELU | (ops.FVT_CONV_RESIDUAL if residual else 0))
    wp_s, wp_t = ops.pack_conv_weight(d_s, w_s), ops.pack_conv_weight(d_t, w_t)
    return x, w_s, w_t, sc_m, sh_m, sc_o, sh_o, res, d_s, d_t, wp_s, wp_t


@pytest.mark.parametrize("shape", [(2, 4, 56, 56, 144, True),      # conv2_x, second unit of a block (residual)
                                   (1, 6, 56, 56, 144, False),     # conv2_x, first unit; more frames than P / D ring slots
                                   (1, 1, 56, 56, 144, False),     # one frame: only the centre temporal tap
                                   (3, 2, 28, 28, 144, True),      # 7 row tiles per frame: the last pair has a dummy tile
                                   (2, 3, 14, 14, 144, True),      # second tile clipped at the bottom edge
                                   (1, 5, 56, 56, 96, True),       # another mid width (two 64-channel blocks, 6 K steps)
                                   (20, 3, 56, 56, 144, True),     # more units than CTA pairs: several clips per cluster
                                   (2, 5, 56, 56, 48, False, (1, 5, 1)),     # the row-paired stem: (1,5,1) conv 64 -> 48, then 48 -> 64
                                   (1, 4, 28, 28, 48, True, (1, 1, 3))])     # another filter shape
@pytest.mark.parametrize("input_stationary", [1, 0])
def test_fused_unit_matches_two_launches(cuda_device, lib, shape, input_stationary):
    """K2f (one launch, mid in tensor memory, cta_group::2, A operand from TMEM) == spatial conv launch + temporal conv
    launch on identical inputs: the bf16 rounding of mid is the same, so results differ by fp32 summation order only
    (<= 1 bf16 ulp per element); and both agree with a torch fp32 evaluation at the bf16 tolerance (1e-2 of max).
    Both schedules of the temporal conv: input-stationary (one N = 192 MMA chain per mid frame, rotating tap window,
    slots zeroed by the output warps) and output-stationary (three N = 64 chains per output frame)."""
    import torch
    import torch.nn.functional as F
    from fastvideotagging_b200 import ops
    n, t, h, w, mid, residual = shape[:6]
    kernel = shape[6] if len(shape) > 6 else (1, 3, 3)
    if kernel != (1, 3, 3) and not input_stationary:
        pytest.skip("the output-stationary form handles 3x3 filters only")
    x, w_s, w_t, sc_m, sh_m, sc_o, sh_o, res, d_s, d_t, wp_s, wp_t = _unit_case(cuda_device, n, t, h, w, mid, n * 10 + t, residual, kernel)
    assert ops.unit2p1_supported(d_s, d_t)
    y_mid = ops.conv3d_fwd(d_s, x, wp_s, sc_m, sh_m)
    y_two = ops.conv3d_fwd(d_t, y_mid, wp_t, sc_o, sh_o, res)
    y_fused = torch.full_like(y_two, float("nan"))
    assert ops.set_option("unit_input_stationary", input_stationary) == 0
    try:
        ops.unit2p1_fwd(d_s, d_t, x, wp_s, sc_m, sh_m, wp_t, sc_o, sh_o, res, out=y_fused)
        torch.cuda.synchronize()
    finally:
        ops.set_option("unit_input_stationary", 1)
    a, b = y_fused.float(), y_two.float()
    assert torch.isfinite(a).all()
    tol = 2 ** -7 * b.abs() + 2 ** -7 * 1e-2 * b.abs().max()
    assert ((a - b).abs() <= tol).all(), ((a - b).abs().max().item(), b.abs().max().item())
    assert (a != b).float().mean().item() < 0.02
    # torch fp32 evaluation from the same bf16 inputs / bf16-rounded weights, mid rounded to bf16 as both paths store it
    xf = x.float().permute(0, 4, 1, 2, 3)
    m = F.conv3d(xf, w_s.to(torch.bfloat16).float(), padding=(0, kernel[1] // 2, kernel[2] // 2))
    m = torch.relu(m * sc_m.view(1, -1, 1, 1, 1) + sh_m.view(1, -1, 1, 1, 1)).to(torch.bfloat16).float()
    o = F.conv3d(m, w_t.to(torch.bfloat16).float(), padding=(1, 0, 0)) * sc_o.view(1, -1, 1, 1, 1) + sh_o.view(1, -1, 1, 1, 1)
    if res is not None:
        o = o + res.float().permute(0, 4, 1, 2, 3)
    o = torch.relu(o).permute(0, 2, 3, 4, 1)
    assert (a - o).abs().max().item() <= 1e-2 * o.abs().max().item()
    assert a.abs().max().item() > 0


def test_fused_unit_rejects_other_geometries(cuda_device, lib):
    """Strided / wider units are not eligible: the caller keeps the two-launch path (no silent approximation)."""
    from fastvideotagging_b200 import ops, _lib
    d_s = ops.conv_desc(1, 4, 28, 28, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1), ops.FVT_CONV_RELU)
    d_t = ops.conv_desc(1, 4, 28, 28, 288, 128, (3, 1, 1), (1, 1, 1), (1, 0, 0), ops.FVT_CONV_RELU)
    assert not ops.unit2p1_supported(d_s, d_t)
    d_s2 = ops.conv_desc(1, 4, 56, 56, 64, 144, (1, 3, 3), (1, 2, 2), (0, 1, 1), ops.FVT_CONV_RELU)
    d_t2 = ops.conv_desc(1, 4, 28, 28, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), ops.FVT_CONV_RELU)
    assert not ops.unit2p1_supported(d_s2, d_t2)
    import ctypes
    with pytest.raises(_lib.FvtError):          # the C entry point itself refuses (bad descriptor pair), nothing is launched
        _lib.check(lib.fvt_unit2p1_fwd(_lib.handle(), ctypes.byref(d_s), ctypes.byref(d_t), *([None] * 10)))


@pytest.mark.parametrize("shape", [(2, 4, 28, 28, 128, 288, (1, 3, 3), (0, 1, 1)),     # conv3_x 1x3x3: two N tiles, two channel blocks
                                   (1, 3, 28, 28, 128, 288, (1, 3, 3), (0, 1, 1)),     # odd number of row tiles (21): a dummy tile
                                   (2, 3, 56, 56, 144, 64, (1, 3, 3), (0, 1, 1)),      # conv2_x data-gradient shape: 2.25 channel blocks
                                   (1, 2, 14, 14, 128, 288, (1, 3, 3), (0, 1, 1)),     # clipped bottom rows
                                   (40, 4, 28, 28, 128, 288, (1, 3, 3), (0, 1, 1))])   # more tile pairs than clusters
def test_cta_pair_auto_matches_streamed_single_cta(cuda_device, lib, shape):
    """Layers whose filter fits two SMs but not one run on the CTA-pair kernel by default (one N tile per cluster, input ring
    per channel block).  Same convolution as the single-CTA streamed-filter slab kernel; the channel-block-major MMA order
    changes the fp32 summation order only (<= 1 bf16 ulp): plain, affine + residual + ReLU, and BatchNorm statistics."""
    import torch
    from fastvideotagging_b200 import ops
    n, t, h, w, cin, cout, k, p = shape
    gen = torch.Generator().manual_seed(n * 100 + h + cin)
    x = (torch.randn(n, t, h, w, cin, generator=gen) * 0.5).to(torch.bfloat16).to(cuda_device)
    wt = (torch.randn(cout, cin, *k, generator=gen) / (cin * k[1] * k[2]) ** 0.5).to(cuda_device)
    res = torch.randn(n, t, h, w, cout, generator=gen).to(torch.bfloat16).to(cuda_device)
    sc = (0.5 + torch.rand(cout, generator=gen)).to(cuda_device)
    sh = torch.randn(cout, generator=gen).to(cuda_device)
    d_plain = ops.conv_desc(n, t, h, w, cin, cout, k, (1, 1, 1), p, 0)
    d_full = ops.conv_desc(n, t, h, w, cin, cout, k, (1, 1, 1), p, ops.FVT_CONV_RELU | ops.FVT_CONV_RESIDUAL)
    d_stat = ops.conv_desc(n, t, h, w, cin, cout, k, (1, 1, 1), p, ops.FVT_CONV_STATS)
    wp = ops.pack_conv_weight(d_plain, wt)

    def run_all():
        a = ops.conv3d_fwd(d_plain, x, wp).clone()
        b = ops.conv3d_fwd(d_full, x, wp, sc, sh, res).clone()
        st = ops.stats_buffer(cout, cuda_device)
        c = ops.conv3d_fwd(d_stat, x, wp, stats=st).clone()
        torch.cuda.synchronize()
        return a, b, c, ops.stats_decode(st)

    out = {}
    try:
        for mode in (0, 1):
            assert ops.set_option("slab_pair_auto", mode) == 0
            out[mode] = run_all()
    finally:
        ops.set_option("slab_pair_auto", 1)
    for i in range(3):
        a, b = out[1][i].float(), out[0][i].float()
        assert torch.isfinite(a).all()
        tol = 2 ** -7 * b.abs() + 2 ** -7 * 1e-2 * b.abs().max()
        assert ((a - b).abs() <= tol).all(), (i, (a - b).abs().max().item(), b.abs().max().item())
    ref = out[0][3]
    assert (out[1][3] - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3
    assert out[0][0].float().abs().max().item() > 0


@pytest.mark.parametrize("shape", [(2, 16, 28, 28, 288, 128),      # conv3_x 3x1x1: 4.5 channel blocks, whole clip per tile
                                   (2, 16, 28, 28, 128, 288),      # its data gradient: two N tiles
                                   (3, 8, 14, 14, 288, 128),       # 8 frames x 16 positions per tile, ragged last column chunk
                                   (2, 5, 7, 9, 288, 128),         # 4 frames x 32 positions, clipped frame tile and column chunk
                                   (1, 32, 28, 28, 288, 128),      # two frame tiles per clip: temporal halo crosses tiles
                                   (30, 16, 28, 28, 288, 128)])    # more tile pairs than clusters
def test_cta_pair_temporal_matches_im2col(cuda_device, lib, shape):
    """Stride-1 3x1x1 convs whose filter fits two SMs but not one run on the CTA-pair slab kernel (frames as image rows, H*W
    positions as image columns).  Same convolution as the generic im2col kernel K1: results equal up to fp32 summation
    order (<= 1 bf16 ulp): plain, affine + residual + ReLU, and BatchNorm statistics."""
    import torch
    from fastvideotagging_b200 import ops
    n, t, h, w, cin, cout = shape
    gen = torch.Generator().manual_seed(n * 100 + t + cin)
    x = (torch.randn(n, t, h, w, cin, generator=gen) * 0.5).to(torch.bfloat16).to(cuda_device)
    wt = (torch.randn(cout, cin, 3, 1, 1, generator=gen) / (cin * 3) ** 0.5).to(cuda_device)
    res = torch.randn(n, t, h, w, cout, generator=gen).to(torch.bfloat16).to(cuda_device)
    sc = (0.5 + torch.rand(cout, generator=gen)).to(cuda_device)
    sh = torch.randn(cout, generator=gen).to(cuda_device)
    k, p = (3, 1, 1), (1, 0, 0)
    d_plain = ops.conv_desc(n, t, h, w, cin, cout, k, (1, 1, 1), p, 0)
    d_full = ops.conv_desc(n, t, h, w, cin, cout, k, (1, 1, 1), p, ops.FVT_CONV_RELU | ops.FVT_CONV_RESIDUAL)
    d_stat = ops.conv_desc(n, t, h, w, cin, cout, k, (1, 1, 1), p, ops.FVT_CONV_STATS)
    wp = ops.pack_conv_weight(d_plain, wt)

    def run_all():
        a = ops.conv3d_fwd(d_plain, x, wp).clone()
        b = ops.conv3d_fwd(d_full, x, wp, sc, sh, res).clone()
        st = ops.stats_buffer(cout, cuda_device)
        c = ops.conv3d_fwd(d_stat, x, wp, stats=st).clone()
        torch.cuda.synchronize()
        return a, b, c, ops.stats_decode(st)

    out = {}
    try:
        for mode in (0, 2):                      # 2: the pair kernel even below the problem size where it pays off
            assert ops.set_option("slab_pair_auto", mode) == 0
            out[mode] = run_all()
    finally:
        ops.set_option("slab_pair_auto", 1)
    for i in range(3):
        a, b = out[2][i].float(), out[0][i].float()
        assert torch.isfinite(a).all()
        tol = 2 ** -7 * b.abs() + 2 ** -7 * 1e-2 * b.abs().max()
        assert ((a - b).abs() <= tol).all(), (i, (a - b).abs().max().item(), b.abs().max().item())
    ref = out[0][3]
    assert (out[2][3] - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3
    assert out[0][0].float().abs().max().item() > 0


@pytest.mark.parametrize("shape", [(4, 8, 14, 14, 256, 576, (1, 3, 3), (1, 1, 1), (0, 1, 1)),      # conv4_x 1x3x3: three N tiles
                                   (4, 8, 14, 14, 576, 256, (3, 1, 1), (1, 1, 1), (1, 0, 0)),      # conv4_x 3x1x1
                                   (3, 4, 7, 7, 512, 1152, (1, 3, 3), (1, 1, 1), (0, 1, 1)),       # conv5_x: 5 tiles (odd: dummy tile), ragged rows
                                   (2, 8, 28, 28, 128, 464, (1, 3, 3), (1, 2, 2), (0, 1, 1)),      # strided first conv of conv4_x
                                   (2, 8, 14, 14, 464, 256, (3, 1, 1), (2, 1, 1), (1, 0, 0)),      # strided temporal conv
                                   (24, 8, 14, 14, 256, 576, (1, 3, 3), (1, 1, 1), (0, 1, 1))])    # more items than clusters
def test_cta_pair_im2col_matches_single_cta(cuda_device, lib, shape):
    """K1p (generic im2col convolution on CTA pairs, cta_group::2, half of the weight tile per CTA) == K1 on one CTA:
    same MMAs in the same order -> bit-identical outputs (plain, affine + residual + ReLU), statistics up to atomics order."""
    import torch
    from fastvideotagging_b200 import ops
    n, t, h, w, cin, cout, k, s, p = shape
    gen = torch.Generator().manual_seed(n * 100 + h + cin)
    x = (torch.randn(n, t, h, w, cin, generator=gen) * 0.5).to(torch.bfloat16).to(cuda_device)
    wt = (torch.randn(cout, cin, *k, generator=gen) / (cin * k[0] * k[1] * k[2]) ** 0.5).to(cuda_device)
    d_plain = ops.conv_desc(n, t, h, w, cin, cout, k, s, p, 0)
    to, ho, wo = ops.conv_out_shape(d_plain)
    res = torch.randn(n, to, ho, wo, cout, generator=gen).to(torch.bfloat16).to(cuda_device)
    sc = (0.5 + torch.rand(cout, generator=gen)).to(cuda_device)
    sh = torch.randn(cout, generator=gen).to(cuda_device)
    d_full = ops.conv_desc(n, t, h, w, cin, cout, k, s, p, ops.FVT_CONV_RELU | ops.FVT_CONV_RESIDUAL)
    d_stat = ops.conv_desc(n, t, h, w, cin, cout, k, s, p, ops.FVT_CONV_STATS)
    wp = ops.pack_conv_weight(d_plain, wt)

    def run_all():
        a = ops.conv3d_fwd(d_plain, x, wp).clone()
        b = ops.conv3d_fwd(d_full, x, wp, sc, sh, res).clone()
        st = ops.stats_buffer(cout, cuda_device)
        c = ops.conv3d_fwd(d_stat, x, wp, stats=st).clone()
        torch.cuda.synchronize()
        return a, b, c, ops.stats_decode(st)

    out = {}
    try:
        assert ops.set_option("disable_split_k", 1) == 0          # same single-pass reduction on both sides
        for mode in (0, 2):                                              # 2: the pair kernel whatever the problem size
            assert ops.set_option("igemm_pair", mode) == 0
            out[mode] = run_all()
    finally:
        ops.set_option("igemm_pair", 1)
        ops.set_option("disable_split_k", 0)
    for i in range(3):
        assert torch.equal(out[0][i], out[2][i]), i
    ref = out[0][3]
    assert (out[2][3] - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-3
    assert out[0][0].float().abs().max().item() > 0


@pytest.mark.parametrize("shape", [(4, 8, 56, 56, 64, 144, (1, 3, 3), (0, 1, 1)),      # conv2_x 1x3x3
                                   (2, 8, 28, 28, 128, 288, (1, 3, 3), (0, 1, 1)),     # conv3_x 1x3x3
                                   (2, 8, 56, 56, 144, 64, (3, 1, 1), (1, 0, 0)),      # conv2_x 3x1x1 (temporal mode)
                                   (2, 4, 28, 28, 288, 128, (3, 1, 1), (1, 0, 0))])    # conv3_x 3x1x1
def test_wgrad_split_reduction_is_deterministic_and_overwrites(cuda_device, lib, shape):
    """Weight gradients reduce their pixel splits through per-split slices of the caller's workspace, added in split
    order: two runs are bit-identical, dw is OVERWRITTEN (grad_req='write', whatever it held), garbage in the workspace
    is harmless, and the result equals the single-split launch (no workspace) up to fp32 summation order."""
    import torch
    from fastvideotagging_b200 import ops
    n, t, h, w, cin, cout, k, p = shape
    gen = torch.Generator().manual_seed(cin + cout)
    x = (torch.randn(n, t, h, w, cin, generator=gen) * 0.5).to(torch.bfloat16).to(cuda_device)
    dy = (torch.randn(n, t, h, w, cout, generator=gen) * 0.1).to(torch.bfloat16).to(cuda_device)
    d = ops.conv_desc(n, t, h, w, cin, cout, k, (1, 1, 1), p, 0)
    taps = k[0] * k[1] * k[2]
    want = ops.conv_workspace_bytes(ops._with_ohwi(d), "wgrad", cout, cin)
    assert want > 0 and want % (cout * cin * taps * 4) == 0          # whole dW-shaped slices, >= 2 of them

    def run(ws, base):
        dw = torch.full((cout, taps, cin), base, dtype=torch.float32, device=cuda_device)
        ops.conv3d_wgrad(d, x, dy, dw, cout, cin, ohwi=True, ws=ws)
        torch.cuda.synchronize()
        return dw

    ops.wgrad_workspace(cuda_device).fill_(float("nan"))
    ref = run(None, 0.0)                     # no workspace: one split per dW tile
    a = run("auto", 0.0)
    b = run("auto", 1.5)                     # overwritten, not accumulated
    assert torch.equal(a, b)
    scale = ref.abs().max().item()
    assert scale > 0 and torch.isfinite(a).all()
    assert (a - ref).abs().max().item() <= 1e-3 * scale          # fp32 sums over 1e5 pixels in two different orders


def test_multi_tensor_pack_matches_single_launch_packs(cuda_device, lib):
    """fvt_pack_conv_weights_multi (all operand copies of a step in one launch) == fvt_pack_conv_weight /
    fvt_pack_conv_weight_dgrad per tensor, bit for bit, for the layer shapes of the network (padded channel counts,
    several N tiles, 1 / 3 / 9 taps)."""
    import torch
    from fastvideotagging_b200 import ops
    gen = torch.Generator().manual_seed(3)
    shapes = [(64, 144, (1, 3, 3), (0, 1, 1)), (144, 64, (3, 1, 1), (1, 0, 0)), (64, 230, (1, 3, 3), (0, 1, 1)),
              (230, 128, (3, 1, 1), (1, 0, 0)), (45, 64, (3, 1, 1), (1, 0, 0)), (256, 921, (1, 3, 3), (0, 1, 1)),
              (921, 512, (3, 1, 1), (1, 0, 0)), (64, 128, (1, 1, 1), (0, 0, 0)), (512, 1152, (1, 3, 3), (0, 1, 1))]
    tf, td = ops.PackTable(cuda_device), ops.PackTable(cuda_device)
    cases = []
    for cin, cout, k, p in shapes:
        w = torch.randn(cout, k[0], k[1], k[2], cin, generator=gen).to(cuda_device)            # (O, kT, kH, kW, I) master
        d = ops.conv_desc(2, 4, 14, 14, ops.pad16(cin), ops.pad16(cout), k, (1, 1, 1), p)
        dd = ops.dgrad_desc(d)
        cases.append((d, dd, w, tf.add_fwd(d, w), td.add_dgrad(dd, w)))
    tf.run()
    td.run()
    for d, dd, w, got_f, got_d in cases:
        ref_f = ops.pack_conv_weight(d, w, ohwi=True)
        ref_d = ops.pack_conv_weight_dgrad(dd, w, ohwi=True)
        torch.cuda.synchronize()
        assert torch.equal(got_f.view(torch.int16), ref_f.view(torch.int16)), d.key()
        assert torch.equal(got_d.view(torch.int16), ref_d.view(torch.int16)), dd.key()


@pytest.mark.parametrize("case", [
    (2, 8, 28, 28, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1)),      # first conv of conv3_x: 1x3x3 / s(1,2,2)
    (2, 8, 14, 14, 230, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0)),     # its temporal conv: 3x1x1 / s(2,1,1)
    (2, 8, 28, 28, 64, 128, (1, 1, 1), (2, 2, 2), (0, 0, 0)),      # projection shortcut: 1x1x1 / s(2,2,2)
    (1, 2, 14, 14, 256, 921, (1, 3, 3), (1, 2, 2), (0, 1, 1)),     # conv5_x first conv at T/4 = 2 -> dY has one frame pair
    (1, 2, 7, 7, 921, 512, (3, 1, 1), (2, 1, 1), (1, 0, 0)),       # dY temporal extent 1
    (1, 4, 9, 11, 32, 48, (3, 3, 3), (2, 2, 2), (1, 1, 1)),        # odd extents, full 3x3x3 / s2 (ECO-style)
    (2, 2, 7, 7, 512, 256, (1, 3, 3), (1, 2, 2), (0, 0, 0)),       # multi-task scene conv: stride 2 WITHOUT padding (low padding in the sub-convolution)
])
def test_strided_dgrad_parity_classes_match_zero_insert_and_autograd(cuda_device, lib, case):
    """Data gradient of a strided convolution as per-parity-class sub-convolutions of dY (fvt_conv3d_fwd_ex + pack kind 2)
    == torch autograd on the same bf16 operands (fp32 accumulate) at 1e-2 of max, and == the zero-insert route up to
    fp32 summation order; with a residual the classes add it at their lattice positions."""
    import torch
    import torch.nn.functional as F
    from fastvideotagging_b200 import ops
    n, t, h, w, cin, cout, k, s, p = case
    gen = torch.Generator().manual_seed(cin * 7 + cout)
    cin_s, cout_s = ops.pad16(cin), ops.pad16(cout)
    fwd = ops.conv_desc(n, t, h, w, cin_s, cout_s, k, s, p)
    to, ho, wo = ops.conv_out_shape(fwd)
    wm = (torch.randn(cout, k[0], k[1], k[2], cin, generator=gen) / (cin * k[0] * k[1] * k[2]) ** 0.5).to(cuda_device)   # (O,kT,kH,kW,I)
    dy = torch.zeros(n, to, ho, wo, cout_s)
    dy[..., :cout] = torch.randn(n, to, ho, wo, cout, generator=gen)
    dy = dy.to(torch.bfloat16).to(cuda_device)
    res = torch.zeros(n, t, h, w, cin_s)
    res[..., :cin] = torch.randn(n, t, h, w, cin, generator=gen)
    res = res.to(torch.bfloat16).to(cuda_device)
    table = ops.PackTable(cuda_device)
    plan = ops.DgradPlan(fwd, wm, table)
    table.run()
    got = plan.run(dy, torch.full((n, t, h, w, cin_s), float("nan"), dtype=torch.bfloat16, device=cuda_device))
    got_res = plan.run(dy, torch.full((n, t, h, w, cin_s), float("nan"), dtype=torch.bfloat16, device=cuda_device), residual=res)
    # reference 1: autograd of the forward convolution on the bf16-rounded operands
    w_ref = wm.to(torch.bfloat16).float().permute(0, 4, 1, 2, 3).contiguous()            # (O, I, kT, kH, kW)
    x0 = torch.zeros(n, cin, t, h, w, device=cuda_device, requires_grad=True)
    y = F.conv3d(x0, w_ref, stride=s, padding=p)
    y.backward(dy[..., :cout].float().permute(0, 4, 1, 2, 3))
    ref = x0.grad.permute(0, 2, 3, 4, 1)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    assert scale > 0
    assert (got[..., :cin].float() - ref).abs().max().item() <= 1e-2 * scale
    assert got[..., cin:].abs().max().item() == 0 if cin_s > cin else True
    assert (got_res[..., :cin].float() - (ref + res[..., :cin].float())).abs().max().item() <= 1e-2 * (scale + res.float().abs().max().item())
    # reference 2: the zero-insert route of round 1
    dd = ops.dgrad_desc(fwd)
    wpd = ops.pack_conv_weight_dgrad(dd, wm, ohwi=True)
    if all(2 * pp == kk - 1 for pp, kk in zip(p, k)):      # the zero-insert route only exists for 'same'-padded filters
        old = ops.conv3d_fwd(dd, ops.zero_insert(dy, fwd), wpd)
        torch.cuda.synchronize()
        assert (got.float() - old.float()).abs().max().item() <= 2 ** -7 * scale


@pytest.mark.parametrize("case", [
    ("conv2_x temporal dgrad 64->144 (K1i)", 2, 8, 56, 56, 144, 64, (3, 1, 1), (1, 0, 0)),
    ("conv2_x spatial dgrad 144->64 (K1s2)", 2, 8, 56, 56, 64, 144, (1, 3, 3), (0, 1, 1)),
    ("conv3_x spatial dgrad 288->128", 2, 4, 28, 28, 128, 288, (1, 3, 3), (0, 1, 1)),
    ("conv3_x temporal dgrad 128->288 (pair)", 4, 16, 28, 28, 288, 128, (3, 1, 1), (1, 0, 0)),
    ("conv4_x spatial dgrad 576->256 (K1)", 2, 4, 14, 14, 256, 576, (1, 3, 3), (0, 1, 1)),
    ("conv5_x temporal dgrad 512->1152 (K1, narrow N)", 2, 2, 7, 7, 1152, 512, (3, 1, 1), (1, 0, 0)),
    ("stem temporal dgrad 64->45 (K1t)", 2, 8, 56, 56, 45, 64, (3, 1, 1), (1, 0, 0)),
])
def test_fused_dgrad_bn_backward_matches_the_two_pass_form(cuda_device, lib, case):
    """FVT_CONV_BN_BWD: the data-gradient convolution masks with the consumer BatchNorm's ReLU and accumulates its backward
    sums in the epilogue; fvt_bn_backward(dz_in=2) then needs one pass.  Against the unfused sequence (dgrad -> two-pass
    BatchNorm backward) on the same inputs: dz bit-identical, dgamma/dbeta to fp32 summation order, d(raw) to one bf16
    rounding; twice the same bits.  One case per kernel the training plan dispatches data gradients to."""
    import torch
    from fastvideotagging_b200 import ops
    name, n, t, h, w, cin, cout, k, p = case        # forward conv cin -> cout; its data gradient maps cout -> cin
    gen = torch.Generator().manual_seed(cin + 3 * cout)
    cin_s, cout_s = ops.pad16(cin), ops.pad16(cout)
    fwd = ops.conv_desc(n, t, h, w, cin_s, cout_s, k, (1, 1, 1), p)
    dd = ops.dgrad_desc(fwd)
    wm = (torch.randn(cout, k[0], k[1], k[2], cin, generator=gen) / (cout * k[0] * k[1] * k[2]) ** 0.5).to(cuda_device)
    wpd = ops.pack_conv_weight_dgrad(dd, wm, ohwi=True)
    dy = torch.zeros(n, t, h, w, cout_s)
    dy[..., :cout] = torch.randn(n, t, h, w, cout, generator=gen)
    dy = dy.to(torch.bfloat16).to(cuda_device)
    # the consumer BatchNorm (of the layer that produced this conv's input): raw, batch statistics, forward constants
    raw = torch.zeros(n, t, h, w, cin_s)
    raw[..., :cin] = torch.randn(n, t, h, w, cin, generator=gen) * 1.3 + 0.2
    raw = raw.to(torch.bfloat16).to(cuda_device)
    rows = n * t * h * w
    rf = raw.float().reshape(rows, cin_s)
    gamma = (0.5 + torch.rand(cin, generator=gen)).to(cuda_device)
    beta = (0.3 * torch.randn(cin, generator=gen)).to(cuda_device)
    mean = torch.zeros(cin_s, device=cuda_device)
    invstd = torch.zeros(cin_s, device=cuda_device)
    mean[:cin] = rf[:, :cin].mean(0)
    invstd[:cin] = 1.0 / torch.sqrt(rf[:, :cin].var(0, unbiased=False) + 1e-5)
    scale = torch.zeros(cin_s, device=cuda_device)
    shift = torch.zeros(cin_s, device=cuda_device)
    scale[:cin] = gamma * invstd[:cin]
    shift[:cin] = beta - mean[:cin] * scale[:cin]
    # ---- unfused: dgrad, then the two-pass BatchNorm backward with the mask recomputed from raw
    assert ops.set_option("disable_split_k", 1) == 0      # the fused form never splits K: same summation order on both sides
    try:
        dact = ops.conv3d_fwd(dd, dy, wpd)
    finally:
        ops.set_option("disable_split_k", 0)
    sums_a = torch.empty(2 * cin_s, device=cuda_device)
    draw_a = torch.empty_like(raw)
    dz_a = torch.empty_like(raw)
    ops.bn_backward(raw, dact, None, mean, invstd, gamma, sums_a, draw_a, relu_scale=scale, relu_shift=shift)
    pre = rf * scale + shift
    dz_ref = torch.where(pre > 0, dact.float().reshape(rows, cin_s), torch.zeros_like(pre)).to(torch.bfloat16)
    # ---- fused
    def fused():
        d2 = ops.ConvDesc(*dd.key())
        d2.flags = ops.FVT_CONV_STATS | ops.FVT_CONV_BN_BWD | ops.FVT_CONV_RESIDUAL
        acc = ops.stats_buffer(cin_s, cuda_device)
        dz = ops.conv3d_fwd(d2, dy, wpd, scale=scale, shift=shift, residual=raw, stats=acc)
        sums = torch.empty(2 * cin_s, device=cuda_device)
        draw = torch.empty_like(raw)
        ops.bn_backward(raw, dz, None, mean, invstd, gamma, sums, draw, sums_acc=acc, dz_in=2)
        torch.cuda.synchronize()
        return dz, sums, draw
    dz_b, sums_b, draw_b = fused()
    dz_c, sums_c, draw_c = fused()
    assert torch.equal(dz_b, dz_c) and torch.equal(sums_b, sums_c) and torch.equal(draw_b, draw_c), name
    assert torch.equal(dz_b.reshape(rows, cin_s).view(torch.int16), dz_ref.view(torch.int16)), name
    sg, sb = sums_a[:cin].abs().max().item(), sums_a[cin_s:cin_s + cin].abs().max().item()
    assert (sums_b[:cin] - sums_a[:cin]).abs().max().item() <= 1e-3 * sg + 1e-4, name
    assert (sums_b[cin_s:cin_s + cin] - sums_a[cin_s:cin_s + cin]).abs().max().item() <= 1e-4 * sb + 1e-5, name
    scale_d = draw_a.float().abs().max().item()
    assert (draw_b.float() - draw_a.float()).abs().max().item() <= 2 ** -6 * scale_d, name
    assert float(draw_b[..., cin:].float().abs().max()) == 0.0 if cin_s > cin else True


def test_grouped_wgrad_matches_single_launches(cuda_device, lib):
    """fvt_conv3d_wgrad_group_plan/_run: the weight gradients of several layers in one grid (conv4_x / conv5_x shapes of
    the training plan, spatial and temporal, plus layers the group does not take) == the one-layer launches up to fp32
    summation order (the group uses fewer pixel splits), == torch autograd on the same bf16 operands at 1e-2 of max; twice
    the same bits."""
    import torch
    import torch.nn.functional as F
    from fastvideotagging_b200 import ops
    gen = torch.Generator().manual_seed(11)
    shapes = [  # n, t, h, w, cin, cout, k, stride, pad
        (2, 4, 14, 14, 256, 576, (1, 3, 3), (1, 1, 1), (0, 1, 1)),      # conv4_x spatial
        (2, 4, 14, 14, 576, 256, (3, 1, 1), (1, 1, 1), (1, 0, 0)),      # conv4_x temporal (flattened T*H*W tiles)
        (2, 2, 7, 7, 512, 1152, (1, 3, 3), (1, 1, 1), (0, 1, 1)),       # conv5_x spatial
        (2, 2, 7, 7, 1152, 512, (3, 1, 1), (1, 1, 1), (1, 0, 0)),       # conv5_x temporal: 49-position frames
        (2, 8, 28, 28, 288, 128, (3, 1, 1), (1, 1, 1), (1, 0, 0)),      # conv3_x temporal: long pixel range -> split + slice reduction
        (2, 4, 14, 14, 256, 921, (1, 3, 3), (1, 2, 2), (0, 1, 1)),      # strided: not in the group (run behind it)
        (2, 4, 14, 14, 230, 128, (3, 1, 1), (1, 1, 1), (1, 0, 0)),      # channel counts with padding (230 -> 240 stored)
        (2, 6, 28, 28, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),       # conv2_x temporal: taps on the N side (N = 192)
        (2, 6, 28, 28, 45, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),        # stem temporal: one (padded) channel block, taps on N
    ]
    layers, refs = [], []
    for n, t, h, w, cin, cout, k, s, p in shapes:
        cin_s, cout_s = ops.pad16(cin), ops.pad16(cout)
        fwd = ops.conv_desc(n, t, h, w, cin_s, cout_s, k, s, p)
        to, ho, wo = ops.conv_out_shape(fwd)
        x = torch.zeros(n, t, h, w, cin_s)
        x[..., :cin] = torch.randn(n, t, h, w, cin, generator=gen)
        dy = torch.zeros(n, to, ho, wo, cout_s)
        dy[..., :cout] = torch.randn(n, to, ho, wo, cout, generator=gen)
        x, dy = x.to(torch.bfloat16).to(cuda_device), dy.to(torch.bfloat16).to(cuda_device)
        dw = torch.full((cout, k[0], k[1], k[2], cin), float("nan"), device=cuda_device)
        layers.append((fwd, x, dy, dw, cout, cin))
        single = ops.conv3d_wgrad(fwd, x, dy, torch.empty_like(dw), cout, cin, ohwi=True)
        w0 = torch.zeros(cout, cin, *k, device=cuda_device, requires_grad=True)
        y = F.conv3d(x[..., :cin].float().permute(0, 4, 1, 2, 3), w0, stride=s, padding=p)
        y.backward(dy[..., :cout].float().permute(0, 4, 1, 2, 3))
        refs.append((single, w0.grad.permute(0, 2, 3, 4, 1)))
    group = ops.WgradGroup(layers, cuda_device)
    assert group.in_group == [True, True, True, True, True, False, True, True, True]
    assert group.red_blocks > 0                      # at least the conv3_x layer is split
    group.run()
    torch.cuda.synchronize()
    first = [L[3].clone() for L in layers]
    for (fwd, x, dy, dw, cout, cin), (single, ref) in zip(layers, refs):
        scale = ref.abs().max().item()
        assert torch.isfinite(dw).all()
        assert (dw - ref).abs().max().item() <= 1e-2 * scale
        assert (dw - single).abs().max().item() <= 1e-4 * scale
    for L in layers:
        L[3].fill_(float("nan"))
    group.run()
    torch.cuda.synchronize()
    for L, f in zip(layers, first):
        assert torch.equal(L[3], f)
