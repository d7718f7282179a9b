"""Host-side mirror of the reference's builder API: names, argument lists and structural known-answers.
CPU only (no kernel is launched): the GPU behaviour of these classes is in test_gpu_builders.py."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_model_package_exports_reference_names():
    import fastvideotagging_b200.model as m
    for name in ("R2Plus2D", "R3DBlock", "get_spatial_temporal_conv", "get_R2plus1d", "BLOCK_CONFIG", "LsepLoss", "LsepLossHy",
                 "LSEP_funcLoss", "WarpLoss", "WARP_funcLoss"):
        assert hasattr(m, name), name
    assert m.BLOCK_CONFIG[34] == (3, 4, 6, 3) and m.BLOCK_CONFIG[18] == (2, 2, 2, 2)     # model/R2Plus1.py:84-90


def test_create_r3d_argument_lists_match_reference_log(capsys):
    """r2plus1d_output/log.txt:38 — 'symbol has 349 = 211 arg + 138 aux' for depth 34 / 101 classes."""
    from fastvideotagging_b200.net import create_r3d
    sym = create_r3d(101, no_bias=1, model_depth=34, final_spatial_kernel=7, final_temporal_kernel=4)
    assert capsys.readouterr().out.strip() == "16"             # net.py:162 prints the block count
    args, aux = sym.list_arguments(), sym.list_auxiliary_states()
    assert len(args) == 211 and len(aux) == 138
    assert args[0] == "data" and args[-1] == "softmax_label" and args[-3:-1] == ["final_fc_weight", "final_fc_bias"]
    assert args[1:7] == ["conv1_middle_weight", "conv1_middle_spatbn_relu_gamma", "conv1_middle_spatbn_relu_beta",
                         "conv1_weight", "conv1_spatbn_relu_gamma", "conv1_spatbn_relu_beta"]
    assert "shortcut_projection_3_weight" in args and "shortcut_projection_0_weight" not in args
    assert len(set(args)) == len(args)
    gold = json.load(open(os.path.join(HERE, "golden", "structure_golden.json")))
    if "r34_param_names" in gold:
        assert [a for a in args if a not in ("data", "softmax_label")] == gold["r34_param_names"]
    a_shapes, o_shapes, x_shapes = sym.infer_shape((4, 3, 32, 112, 112))
    assert o_shapes == [(4, 101)] and a_shapes[0] == (4, 3, 32, 112, 112)
    assert dict(zip(args, a_shapes))["comp_0_conv_1_middle_weight"] == (144, 64, 1, 3, 3)
    assert dict(zip(args, a_shapes))["comp_13_conv_1_middle_weight"] == (921, 256, 1, 3, 3)


def test_model_builder_rejects_wrong_channel_chain():
    from fastvideotagging_b200.net import ModelBuilder, _Node
    b = ModelBuilder(no_bias=1)
    body = b.add_r3d_block(_Node(64, 1, 2, []), 64, 64)
    with pytest.raises(ValueError):
        b.add_r3d_block(body, 128, 128)


def test_block_builders_have_reference_structure():
    from fastvideotagging_b200.model import R3DBlock, get_spatial_temporal_conv
    unit = get_spatial_temporal_conv(64, 128, [2, 2, 2])
    assert unit.middle_filters == 230 and unit.conv_middle.strides == (1, 2, 2) and unit.conv.strides == (2, 1, 1)
    assert tuple(unit.conv_middle.weight.shape) == (230, 64, 1, 3, 3) and tuple(unit.conv.weight.shape) == (128, 230, 3, 1, 1)
    blk = R3DBlock(64, 128, comp_index=3, downsampling=True)
    assert hasattr(blk, "branch_conv") and blk.branch_conv.strides == (2, 2, 2) and blk.use_striding == [2, 2, 2]
    same = R3DBlock(64, 64, comp_index=0)
    assert not hasattr(same, "branch_conv")


def test_blocks_refuse_cpu_tensors():
    import torch
    from fastvideotagging_b200.model import R3DBlock
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        R3DBlock(64, 64, comp_index=0)(torch.zeros(1, 64, 2, 8, 8))


def test_row_paired_stem_geometry_is_the_same_convolution():
    """engine.StemGeometry (CPU, torch only): the stride-1 (1,5,1) conv over the row-paired W-unfolded clip with the
    re-expressed filter equals the reference's Conv3D(45,(1,7,7),s(1,2,2),p(0,3,3)) (model/R2Plus1.py:100-104), and
    weight_grad is the adjoint of weight (so gradients map back exactly)."""
    import torch
    import torch.nn.functional as F
    from fastvideotagging_b200.engine import StemGeometry
    gen = torch.Generator().manual_seed(0)
    n, t, h, w = 2, 3, 16, 20
    x = torch.randn(n, 3, t, h, w, generator=gen, dtype=torch.float64)
    wt = torch.randn(45, 3, 1, 7, 7, generator=gen, dtype=torch.float64)
    ref = F.conv3d(x, wt, stride=(1, 2, 2), padding=(0, 3, 3))
    for hpair in (True, False):
        geo = StemGeometry(h if hpair else h + 1)
        assert geo.hpair == hpair
        wo = (w + 6 - 7) // 2 + 1
        xp = F.pad(x, (3, 3))
        # W-unfold by definition: u[n,t,h,ow,kw*3+ci] = x[n,ci,t,h,2*ow-3+kw]
        u = torch.stack([xp[..., 2 * ow:2 * ow + 7] for ow in range(wo)], dim=-2)       # (n,3,t,h,wo,7)
        u = u.permute(0, 2, 3, 4, 5, 1).reshape(n, t, h, wo, 21)
        if hpair:
            u32 = torch.zeros(n, t, h, wo, 32, dtype=x.dtype)
            u32[..., :21] = u
            u2 = u32.reshape(n, t, h // 2, 2, wo, 32).permute(0, 1, 2, 4, 3, 5).reshape(n, t, h // 2, wo, 64)
            inp, cin = u2, 64
        else:
            inp, cin = u, 21
        weq = geo.weight(wt)
        assert tuple(weq.shape) == (45, cin, *geo.kernel)
        got = F.conv3d(inp.permute(0, 4, 1, 2, 3), weq, stride=geo.stride, padding=geo.pad)
        assert got.shape == ref.shape
        assert torch.allclose(got, ref, atol=1e-10)
        g = torch.randn(weq.shape, generator=gen, dtype=torch.float64)
        lhs = (weq * g).sum()
        rhs = (wt * geo.weight_grad(g)).sum()
        assert abs(lhs.item() - rhs.item()) < 1e-9 * max(1.0, abs(lhs.item()))


def test_lr_schedulers_follow_mxnet_semantics():
    """FactorScheduler / MultiFactorScheduler as the reference drives them (train.py:75-77; train_simple_r3d.py:99-106,
    per epoch with steps "2,5,10")."""
    from fastvideotagging_b200.trainer import FactorScheduler, MultiFactorScheduler, split_and_load
    import numpy as np
    import torch
    s = MultiFactorScheduler([2, 5, 10], 0.1)
    s.base_lr = 1e-2
    got = [s(e) for e in range(13)]
    exp = [1e-2] * 3 + [1e-3] * 3 + [1e-4] * 5 + [1e-5] * 2          # the factor applies once the epoch EXCEEDS a step
    assert np.allclose(got, exp, rtol=1e-12)
    f = FactorScheduler(step=4, factor=0.5, base_lr=1.0)
    assert [f(i) for i in (0, 4, 5, 8, 9, 13)] == [1.0, 1.0, 0.5, 0.5, 0.25, 0.125]
    with pytest.raises(ValueError):
        MultiFactorScheduler([5, 2], 0.1)
    parts = split_and_load(torch.arange(12).reshape(6, 2), [None, None, None])
    assert [p.shape[0] for p in parts] == [2, 2, 2] and parts[2][0, 0].item() == 8
    with pytest.raises(ValueError):
        split_and_load(torch.zeros(5, 2), [None, None])


def test_strided_dgrad_parity_classes_are_the_data_gradient():
    """ops.dgrad_parity_classes (the index arithmetic behind the direct strided data gradient, fvt_conv3d_fwd_ex): composing the
    per-class stride-1 sub-convolutions of dY reproduces torch autograd's conv3d input gradient exactly (fp64), for every
    strided convolution geometry of the network, the multi-task scene conv (stride 2 without padding) and a 3x3x3/s2 stage
    with odd extents."""
    import torch
    import torch.nn.functional as F
    from fastvideotagging_b200 import ops
    cases = [((8, 28, 28), (1, 3, 3), (1, 2, 2), (0, 1, 1)), ((8, 14, 14), (3, 1, 1), (2, 1, 1), (1, 0, 0)),
             ((8, 28, 28), (1, 1, 1), (2, 2, 2), (0, 0, 0)), ((2, 7, 7), (1, 3, 3), (1, 2, 2), (0, 0, 0)),
             ((4, 9, 11), (3, 3, 3), (2, 2, 2), (1, 1, 1)), ((2, 1, 1), (3, 1, 1), (2, 1, 1), (1, 0, 0))]
    gen = torch.Generator().manual_seed(0)
    for x_ext, k, s, p in cases:
        cin, cout = 3, 4
        w = torch.randn(cout, cin, *k, dtype=torch.float64, generator=gen)
        o_ext, cls = ops.dgrad_parity_classes(x_ext, k, s, p)
        dy = torch.randn(1, cout, *o_ext, dtype=torch.float64, generator=gen)
        x0 = torch.zeros(1, cin, *x_ext, dtype=torch.float64, requires_grad=True)
        F.conv3d(x0, w, stride=s, padding=p).backward(dy)
        dx = torch.zeros(1, cin, *x_ext, dtype=torch.float64)
        for par, sub_k, tap_a, pad_lo, pad_hi in cls:
            g = torch.zeros(cin, cout, *sub_k, dtype=torch.float64)
            for ut in range(sub_k[0]):
                for uh in range(sub_k[1]):
                    for uw in range(sub_k[2]):
                        g[:, :, ut, uh, uw] = w[:, :, tap_a[0] - s[0] * ut, tap_a[1] - s[1] * uh, tap_a[2] - s[2] * uw].t()
            dyp = F.pad(dy, (pad_lo[2], pad_hi[2], pad_lo[1], pad_hi[1], pad_lo[0], pad_hi[0]))
            dx[:, :, par[0]::s[0], par[1]::s[1], par[2]::s[2]] = F.conv3d(dyp, g)
        assert float((dx - x0.grad).abs().max()) < 1e-12, (x_ext, k, s, p)
