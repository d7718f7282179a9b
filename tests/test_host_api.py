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
