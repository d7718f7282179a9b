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


def _ref_group(x, w, b, gamma, beta, res, stride, pad, relu, eps=1e-5, relu_mask=None):
    """torch-CPU fp32 restatement of Conv3D (+bias) -> BatchNorm(batch statistics, biased variance) (+ residual) (-> ReLU)
    on bf16-rounded operands, with the raw conv output rounded to bf16 as the kernels store it (straight-through).
    relu_mask (teacher forcing): the ReLU passes exactly the positions the kernel's output passed — a pre-activation within
    rounding of zero may fall on either side, which changes the output by nothing but single gradient entries by O(|x||dy|)."""
    import torch.nn.functional as F
    raw = F.conv3d(x, w, None, stride=stride, padding=pad)
    raw = raw + (raw.to(torch.bfloat16).float() - raw).detach()
    if b is not None:
        raw = raw + b.reshape(1, -1, 1, 1, 1)
    mean = raw.mean(dim=(0, 2, 3, 4), keepdim=True)
    var = raw.var(dim=(0, 2, 3, 4), unbiased=False, keepdim=True)
    y = (raw - mean) / torch.sqrt(var + eps) * gamma.reshape(1, -1, 1, 1, 1) + beta.reshape(1, -1, 1, 1, 1)
    if res is not None:
        y = y + res
    if relu and relu_mask is not None:
        return y * relu_mask
    return torch.relu(y) if relu else y


@pytest.mark.parametrize("case", [
    ("scene conv 1x3x3/s(1,2,2) + bias", 2, 2, 7, 7, 512, 256, (1, 3, 3), (1, 2, 2), (0, 0, 0), True, True, False),
    ("action conv 1x3x3 + bias", 2, 2, 7, 7, 512, 512, (1, 3, 3), (1, 1, 1), (0, 1, 1), True, True, False),
    ("ECO 3x3x3/s2 96->128", 2, 8, 28, 28, 96, 128, (3, 3, 3), (2, 2, 2), (1, 1, 1), False, True, False),
    ("ECO 3x3x3 128->128 + residual", 2, 4, 14, 14, 128, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), False, True, True),
    ("projection 1x1x1/s2, no ReLU", 2, 4, 14, 14, 64, 128, (1, 1, 1), (2, 2, 2), (0, 0, 0), False, False, False),
])
def test_training_mode_conv_bn_group_matches_torch_autograd(cuda_device, case):
    """blocks._ConvBnFn (the training-mode node of the N4 heads and the stand-alone blocks): forward output, data gradient,
    weight gradient and BatchNorm parameter gradients against torch autograd on the same bf16 operands, 1e-2 of max;
    running statistics follow the MXNet convention (a conv bias only shifts the running mean and gets a zero gradient)."""
    from fastvideotagging_b200.model.blocks import Conv3D, BatchNorm, to_ndhwc, to_ncdhw
    name, n, t, h, w, cin, cout, k, s, p, use_bias, relu, with_res = case
    gen = torch.Generator().manual_seed(cin + cout)
    conv = Conv3D(cin, cout, k, s, p, use_bias=use_bias)
    bn = BatchNorm(cout)
    _rand_bn(bn, gen)
    if use_bias:
        with torch.no_grad():
            conv.bias.copy_(0.3 * torch.randn(cout, generator=gen))
    conv.to(cuda_device), bn.to(cuda_device)
    x = torch.randn(n, cin, t, h, w, generator=gen).to(torch.bfloat16).float()
    xd = to_ndhwc(x.to(cuda_device)).requires_grad_(True)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    to_, ho, wo = (t + 2 * p[0] - k[0]) // s[0] + 1, (h + 2 * p[1] - k[1]) // s[1] + 1, (w + 2 * p[2] - k[2]) // s[2] + 1
    res = torch.randn(n, cout, to_, ho, wo, generator=gen).to(torch.bfloat16).float() if with_res else None
    resd = to_ndhwc(res.to(cuda_device)).requires_grad_(True) if with_res else None
    out = conv.run(xd, bn, relu=relu, residual=resd, training=True)
    dout = (torch.randn(n, cout, to_, ho, wo, generator=gen) * 0.1).to(torch.bfloat16).float()
    out.backward(to_ndhwc(dout.to(cuda_device)))
    torch.cuda.synchronize()
    # reference
    xr = x.clone().requires_grad_(True)
    wr = conv.weight.detach().cpu().to(torch.bfloat16).float().requires_grad_(True)
    br = conv.bias.detach().cpu().clone().requires_grad_(True) if use_bias else None
    gr, btr = bn.gamma.detach().cpu().clone().requires_grad_(True), bn.beta.detach().cpu().clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True) if with_res else None
    kmask = (to_ncdhw(out.detach(), cout).cpu() > 0).float() if relu else None
    ref = _ref_group(xr, wr, br, gr, btr, rr, s, p, relu, relu_mask=kmask)
    ref.backward(dout)
    with torch.no_grad():          # the forced mask differs from the reference's own ReLU only where the pre-activation is ~0
        ref_free = _ref_group(xr, wr, br, gr, btr, rr, s, p, relu)
        assert (ref_free - ref).abs().max().item() <= 1e-2 * ref_free.abs().max().item()

    def close(what, got, want, tol=1e-2):
        scale = want.abs().max().item()
        err = (got - want).abs().max().item()
        assert err <= tol * scale + 1e-6, "%s %s: err %.4g vs max %.4g" % (name, what, err, scale)

    close("output", to_ncdhw(out.detach(), cout).cpu(), ref.detach())
    close("dx", to_ncdhw(xd.grad, cin).cpu(), xr.grad)
    close("dW", conv.weight.grad.cpu(), wr.grad)
    close("dgamma", bn.gamma.grad.cpu(), gr.grad, 2e-2)
    close("dbeta", bn.beta.grad.cpu(), btr.grad, 2e-2)
    if with_res:
        close("dresidual", to_ncdhw(resd.grad, cout).cpu(), rr.grad)
    if use_bias:
        assert float(conv.bias.grad.abs().max()) == 0.0            # exactly zero in front of a batch-statistics BatchNorm
    with torch.no_grad():
        import torch.nn.functional as F
        raw = F.conv3d(x, wr.detach(), br.detach() if use_bias else None, stride=s, padding=p)
        m, v = raw.mean(dim=(0, 2, 3, 4)), raw.var(dim=(0, 2, 3, 4), unbiased=False)
    assert torch.allclose(bn.running_mean.cpu(), 0.9 * rm0.cpu() + 0.1 * m, rtol=1e-2, atol=2e-3)
    assert torch.allclose(bn.running_var.cpu(), 0.9 * rv0.cpu() + 0.1 * v, rtol=1e-2, atol=1e-3)


def test_multitask_network_trains_end_to_end(cuda_device):
    """R2Plus2D_MT in training mode (multi_taskR3d.py:246-267 inside autograd.record()): scene + action losses backward
    through both heads AND the trunk (flat gradient buffer), batch statistics update the running statistics, two SGD steps
    lower the loss; the head gradients match torch autograd given the trunk's own conv5 map (teacher-forced)."""
    from fastvideotagging_b200.model import R2Plus2D_MT
    from fastvideotagging_b200.model.heads import _FlattenDenseFn, _PoolFcFn
    from fastvideotagging_b200.trainer import Trainer
    depth, n, t, hw = 10, 4, 8, 112
    torch.manual_seed(0)
    net = R2Plus2D_MT(num_scenes=21, num_actions=63, model_depth=depth, final_spatial_kernel=7, final_temporal_kernel=1)
    net.trunk.load_param_dict(orc.randomize_bn(orc.init_params(depth, 63, seed=0), seed=1))
    net.to(cuda_device).train()
    net.dropout = 0.0                                              # deterministic comparison below
    x = torch.from_numpy(np.random.default_rng(1).random((n, 3, t, hw, hw), dtype=np.float32)).to(cuda_device)
    ys = torch.randint(0, 21, (n,), device=cuda_device)
    ya = (torch.rand(n, 63, device=cuda_device) < 0.05).float()
    trainer = Trainer(net.trunk, "sgd", {"learning_rate": 1e-2, "momentum": 0.9, "wd": 0.0})
    head_params = [p for nm, p in net.named_parameters() if not nm.startswith("trunk.")]
    opt = torch.optim.SGD(head_params, lr=1e-2, momentum=0.9)
    rv0 = net.action_bn.running_var.clone()
    losses = []
    for _ in range(3):
        opt.zero_grad()
        scene, action = net(x)
        loss = torch.nn.functional.cross_entropy(scene, ys) + torch.nn.functional.binary_cross_entropy_with_logits(action, ya)
        loss.backward()
        g = net.trunk._flat.g
        assert torch.isfinite(g).all() and float(g.abs().max()) > 0
        for p in head_params:
            assert p.grad is not None and torch.isfinite(p.grad).all()
        trainer.step(1)
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses
    assert not torch.equal(net.action_bn.running_var, rv0)
    # flatten + Dense and pool + Dense nodes against torch autograd on the same inputs
    gen = torch.Generator().manual_seed(3)
    s_in = torch.randn(n, 1, 3, 3, 256, generator=gen).to(torch.bfloat16).to(cuda_device).requires_grad_(True)
    dy = torch.randn(n, 21, generator=gen).to(torch.bfloat16).float().to(cuda_device)
    w, b = net.scene_output.weight, net.scene_output.bias
    w.grad = b.grad = None
    out = _FlattenDenseFn.apply(s_in, w, b)
    out.backward(dy)
    flat = s_in.detach().float().permute(0, 4, 1, 2, 3).reshape(n, -1).requires_grad_(True)          # NCDHW flatten order
    wq = w.detach().to(torch.bfloat16).float().requires_grad_(True)
    ref = flat @ wq.t() + b.detach()
    ref.backward(dy)
    assert (out.detach() - ref.detach()).abs().max().item() <= 1e-2 * ref.detach().abs().max().item()
    assert (w.grad - wq.grad).abs().max().item() <= 1e-2 * wq.grad.abs().max().item()
    gx = s_in.grad.float().permute(0, 4, 1, 2, 3).reshape(n, -1)
    assert (gx - flat.grad).abs().max().item() <= 1e-2 * flat.grad.abs().max().item()
    assert torch.allclose(b.grad, dy.sum(0), rtol=1e-5, atol=1e-5)
    a_in = torch.randn(n, 1, 7, 7, 512, generator=gen).to(torch.bfloat16).to(cuda_device).requires_grad_(True)
    dy2 = torch.randn(n, 63, generator=gen).to(cuda_device)
    w2, b2 = net.action_output.weight, net.action_output.bias
    w2.grad = b2.grad = None
    out2 = _PoolFcFn.apply(a_in, w2, b2, 512)
    out2.backward(dy2)
    pooled = a_in.detach().float().mean(dim=(1, 2, 3)).requires_grad_(True)
    w2r = w2.detach().clone().requires_grad_(True)
    ref2 = pooled @ w2r.t() + b2.detach()
    ref2.backward(dy2)
    assert (out2.detach() - ref2.detach()).abs().max().item() <= 1e-3 * ref2.detach().abs().max().item() + 1e-5
    assert (w2.grad - w2r.grad).abs().max().item() <= 1e-3 * w2r.grad.abs().max().item() + 1e-6
    assert (a_in.grad.float()[:, 0, 0, 0] - pooled.grad / 49.0).abs().max().item() <= 1e-2 * (pooled.grad / 49.0).abs().max().item()


def test_eco_lite_head_trains(cuda_device):
    """ECOLite3DHead in training mode: batch-statistics forward, backward through the 3x3x3 weight / data gradient kernels
    (incl. the strided stages via parity sub-convolutions), gradients finite and a few SGD steps lower the loss."""
    from fastvideotagging_b200.model import ECOLite3DHead
    torch.manual_seed(0)
    head = ECOLite3DHead(num_class=11).to(cuda_device).train()
    x = torch.rand(4, 4, 28, 28, 96, device=cuda_device).to(torch.bfloat16)
    y = torch.randint(0, 11, (4,), device=cuda_device)
    opt = torch.optim.SGD(head.parameters(), lr=0.05, momentum=0.9)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(head(x), y)
        loss.backward()
        for p in head.parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < 0.7 * losses[0], losses
