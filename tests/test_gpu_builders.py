"""GPU parity of the stand-alone builders (get_spatial_temporal_conv, R3DBlock, get_R2plus1d) and of the symbol-API
executor (create_r3d(...).bind()) against torch-CPU / oracle evaluations of the same operators."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bn_eval(x, bn):
    return F.batch_norm(x, bn.running_mean.cpu(), bn.running_var.cpu(), bn.gamma.detach().cpu(), bn.beta.detach().cpu(), False, 0.0, bn.eps)


def _randomize(mod, seed):
    g = torch.Generator().manual_seed(seed)
    from fastvideotagging_b200.model.blocks import BatchNorm
    for m in mod.modules():
        if isinstance(m, BatchNorm):
            m.gamma.data = torch.rand(m.channels, generator=g) + 0.5
            m.beta.data = torch.randn(m.channels, generator=g) * 0.1
            m.running_mean.copy_(torch.randn(m.channels, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.channels, generator=g) + 0.5)


def _unit_ref(unit, x):
    q = lambda t: t.to(torch.bfloat16).float()
    y = F.conv3d(q(x), q(unit.conv_middle.weight.detach().cpu()), stride=unit.conv_middle.strides, padding=unit.conv_middle.padding)
    y = q(_bn_eval(y, unit.bn_middle).relu())
    return F.conv3d(y, q(unit.conv.weight.detach().cpu()), stride=unit.conv.strides, padding=unit.conv.padding)


@pytest.mark.parametrize("cin,cout,down", [(64, 64, False), (64, 128, True), (128, 128, False)])
def test_r3d_block_matches_torch_reference(cuda_device, cin, cout, down):
    from fastvideotagging_b200.model import R3DBlock
    torch.manual_seed(1)
    blk = R3DBlock(cin, cout, comp_index=0, downsampling=down)
    _randomize(blk, 2)
    blk.eval()
    x = torch.randn(2, cin, 4, 14, 14) * 0.5
    q = lambda t: t.to(torch.bfloat16).float()
    y = q(_bn_eval(_unit_ref(blk.spatial_temporal_conv1, x), blk.bn1).relu())
    y = _bn_eval(_unit_ref(blk.spatial_temporal_conv2, y), blk.bn2)
    sc = q(x)
    if hasattr(blk, "branch_conv"):
        sc = q(_bn_eval(F.conv3d(q(x), q(blk.branch_conv.weight.detach().cpu()), stride=blk.branch_conv.strides), blk.branch_bn))
    ref = (y + sc).relu()
    with torch.no_grad():
        got = blk.to(cuda_device)(x.to(cuda_device)).cpu()
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-3


def test_spatial_temporal_conv_training_mode_uses_batch_statistics(cuda_device):
    """Inside autograd.record() the reference's unit normalises with batch statistics and updates the running ones
    (MXNet convention: running = 0.9*running + 0.1*batch, biased variance)."""
    from fastvideotagging_b200.model import get_spatial_temporal_conv
    torch.manual_seed(3)
    unit = get_spatial_temporal_conv(64, 64, [1, 1, 1]).to(cuda_device)
    unit.train()
    x = torch.randn(2, 64, 4, 14, 14) * 0.5
    got = unit(x.to(cuda_device)).cpu()
    q = lambda t: t.to(torch.bfloat16).float()
    raw = q(F.conv3d(q(x), q(unit.conv_middle.weight.detach().cpu()), padding=(0, 1, 1)))
    mean = raw.mean(dim=(0, 2, 3, 4)); var = raw.var(dim=(0, 2, 3, 4), unbiased=False)
    y = q(((raw - mean[None, :, None, None, None]) / torch.sqrt(var[None, :, None, None, None] + 1e-5)).relu())
    ref = F.conv3d(y, q(unit.conv.weight.detach().cpu()), padding=(1, 0, 0))
    assert (got - ref).abs().max().item() <= 1.5e-2 * ref.abs().max().item() + 1e-3
    assert torch.allclose(unit.bn_middle.running_mean.cpu(), 0.1 * mean, atol=1e-4)
    assert torch.allclose(unit.bn_middle.running_var.cpu(), 0.9 + 0.1 * var, atol=1e-4)


def test_symbol_executor_matches_oracle_with_symbol_eps(cuda_device):
    """create_r3d(...).bind(): BatchNorm eps = 1e-3 (net.py:44) and a SoftmaxOutput head; prediction = oracle's."""
    from fastvideotagging_b200.net import create_r3d
    from oracle import r2plus1d as orc
    params = orc.randomize_bn(orc.init_params(18, 101, seed=0), seed=1)
    sym = create_r3d(101, no_bias=1, model_depth=18, final_spatial_kernel=7, final_temporal_kernel=1)
    exe = sym.bind(cuda_device, arg_params=params)
    x = np.random.default_rng(5).random((2, 3, 8, 112, 112), dtype=np.float32)
    prob = exe.forward(is_train=False, data=torch.from_numpy(x).to(cuda_device))[0].cpu().numpy()
    logits = orc.Net(params, 18, (1, 7, 7), eps=1e-3, bf16_storage=True).forward(x)[0].numpy()
    e = np.exp(logits - logits.max(axis=1, keepdims=True)); ref = e / e.sum(axis=1, keepdims=True)
    assert np.abs(prob - ref).max() <= 5e-3
    assert (prob.argmax(1) == ref.argmax(1)).all()
    assert np.allclose(prob.sum(1), 1.0, atol=1e-5)
    # training step through the executor: SoftmaxOutput backward = p - onehot, label -1 ignored
    lab = torch.tensor([3.0, -1.0], device=cuda_device)
    exe.forward(is_train=True, data=torch.from_numpy(x).to(cuda_device), softmax_label=lab)
    exe.backward()
    g = exe.net.final_fc_bias.grad.cpu().numpy()
    p0 = exe.outputs[0][0].cpu().numpy()
    onehot = np.zeros(101, dtype=np.float32); onehot[3] = 1
    assert np.abs(g - (p0 - onehot)).max() <= 1e-4          # row 1 (label -1) contributes nothing


def test_get_r2plus1d_sequential_variant_runs_and_is_sigmoid(cuda_device):
    from fastvideotagging_b200.model import get_R2plus1d
    net = get_R2plus1d(num_class=63, model_depth=10, final_spatial_kernel=7, final_temporal_kernel=1).to(cuda_device).eval()
    x = torch.rand(1, 3, 8, 112, 112, device=cuda_device)
    with torch.no_grad():
        y = net(x)
    assert tuple(y.shape) == (1, 63) and float(y.min()) > 0.0 and float(y.max()) < 1.0
