"""Heads next to the trunk on the same conv kernels: multi-task scene/action heads, Decision_thresh,
ECO-Lite 3D head — against oracle/heads.py (torch-CPU fp32, bf16 storage emulated)."""
import numpy as np
import pytest
import torch

from oracle import heads as oh
from oracle import r2plus1d as orc

pytestmark = pytest.mark.gpu


def _rand_bn(mod, gen):
    with torch.no_grad():
        mod.gamma.copy_(0.5 + torch.rand(mod.channels, generator=gen))
        mod.beta.copy_(0.2 * torch.randn(mod.channels, generator=gen))
        mod.running_mean.copy_(0.1 * torch.randn(mod.channels, generator=gen))
        mod.running_var.copy_(0.5 + torch.rand(mod.channels, generator=gen))


def _bn_tuple(mod):
    return tuple(t.detach().float().cpu() for t in (mod.gamma, mod.beta, mod.running_mean, mod.running_var))


def test_multitask_heads_match_oracle(cuda_device):
    from fastvideotagging_b200.model import R2Plus2D_MT
    depth, n, t, hw = 18, 2, 8, 112
    gen = torch.Generator().manual_seed(0)
    net = R2Plus2D_MT(num_scenes=21, num_actions=63, model_depth=depth, final_spatial_kernel=7, final_temporal_kernel=1)
    params = orc.randomize_bn(orc.init_params(depth, 63, seed=0), seed=1)
    net.trunk.load_param_dict(params)
    for bn in (net.scene_bn, net.action_bn):
        _rand_bn(bn, gen)
    with torch.no_grad():
        net.scene_conv.bias.copy_(0.3 * torch.randn(256, generator=gen))
        net.action_conv.bias.copy_(0.3 * torch.randn(512, generator=gen))
        net.scene_output.bias.copy_(0.1 * torch.randn(21, generator=gen))
        net.action_output.bias.copy_(0.1 * torch.randn(63, generator=gen))
    net.to(cuda_device).eval()
    x = np.random.default_rng(123).random((n, 3, t, hw, hw), dtype=np.float32)
    with torch.no_grad():
        scene, action = net(torch.from_numpy(x).to(cuda_device))
    assert tuple(scene.shape) == (n, 21) and tuple(action.shape) == (n, 63)
    taps = {}
    orc.Net(params, depth, (1, 7, 7), bf16_storage=True).forward(x, taps=taps)
    feat = taps["comp_7_out"]
    g = _bn_tuple(net.scene_bn)
    a = _bn_tuple(net.action_bn)
    p = {"scene_conv_weight": net.scene_conv.weight.detach().cpu(), "scene_conv_bias": net.scene_conv.bias.detach().cpu(),
         "scene_bn_gamma": g[0], "scene_bn_beta": g[1], "scene_bn_mean": g[2], "scene_bn_var": g[3],
         "scene_dense_weight": net.scene_output.weight.detach().cpu(), "scene_dense_bias": net.scene_output.bias.detach().cpu(),
         "action_conv_weight": net.action_conv.weight.detach().cpu(), "action_conv_bias": net.action_conv.bias.detach().cpu(),
         "action_bn_gamma": a[0], "action_bn_beta": a[1], "action_bn_mean": a[2], "action_bn_var": a[3],
         "action_dense_weight": net.action_output.weight.detach().cpu(), "action_dense_bias": net.action_output.bias.detach().cpu()}
    ref_s, ref_a = oh.multitask_heads(feat, p, (1, 7, 7), bf16_storage=True)
    for got, ref in ((scene, ref_s), (action, ref_a)):
        got, ref = got.float().cpu().numpy(), ref.numpy()
        assert np.abs(got - ref).max() <= 1e-2 * np.abs(ref).max() + 1e-4, (np.abs(got - ref).max(), np.abs(ref).max())
        assert (got.argmax(1) == ref.argmax(1)).all()


def test_decision_thresh(cuda_device):
    from fastvideotagging_b200.model import Decision_thresh
    m = Decision_thresh(63).to(cuda_device)
    with torch.no_grad():
        m.thresh.copy_(torch.linspace(-1, 1, 63).reshape(1, 63))
    x = torch.randn(5, 63, device=cuda_device)
    assert torch.equal(m(x).detach().cpu(), oh.decision_thresh(x.cpu(), m.thresh.detach().cpu()))


@pytest.mark.parametrize("n,t", [(2, 4), (1, 16)])
def test_eco_lite_3d_head_matches_torch_restatement(cuda_device, n, t):
    from fastvideotagging_b200.model import ECOLite3DHead
    gen = torch.Generator().manual_seed(1)
    head = ECOLite3DHead(num_class=101)
    blocks = []
    for blk in head.blocks:
        for bn in [blk.bn1, blk.bn2] + ([blk.down_bn] if blk.project else []):
            _rand_bn(bn, gen)
        d = {"w1": blk.conv1.weight.detach().clone(), "bn1": _bn_tuple(blk.bn1), "w2": blk.conv2.weight.detach().clone(),
             "bn2": _bn_tuple(blk.bn2), "stride": blk.conv1.strides[0]}
        if blk.project:
            d["wd"], d["bnd"] = blk.down.weight.detach().clone(), _bn_tuple(blk.down_bn)
        blocks.append(d)
    head.to(cuda_device).eval()
    x = torch.rand(n, 96, t, 28, 28, generator=gen)
    with torch.no_grad():
        got = head(x.to(cuda_device)).float().cpu().numpy()
    ref = oh.eco_lite_3d_head(x, blocks, head.dense.weight.detach().cpu(), head.dense.bias.detach().cpu(), bf16_storage=True).numpy()
    assert got.shape == (n, 101)
    assert np.abs(got - ref).max() <= 1e-2 * np.abs(ref).max() + 1e-4, (np.abs(got - ref).max(), np.abs(ref).max())
    assert (got.argmax(1) == ref.argmax(1)).all()
    assert abs(ECOLite3DHead.conv_gflop_per_clip(16, 28) - 83.24) < 0.01
