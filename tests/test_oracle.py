"""CPU tests: the oracle against the reference's recorded known-answers and against golden vectors produced by
executing the reference's own model source (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import r2plus1d as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_mid_filter_sequence_matches_reference_log():
    # r2plus1d_output/log.txt:6-37: 6x144, 230, 7x288, 460, 11x576, 921, 5x1152
    mids = [c["cout"] for c in orc.all_convs(34) if c["kernel"] == (1, 3, 3)]
    assert mids == [144] * 6 + [230] + [288] * 7 + [460] + [576] * 11 + [921] + [1152] * 5
    assert [orc.mid_filters(a, b) for a, b in ((64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 512), (512, 512))] == \
        [144, 230, 288, 460, 576, 921, 1152]


def test_symbol_argument_counts_match_reference_log():
    # r2plus1d_output/log.txt:38 "symbol has 349 = 211 arg + 138 aux" for depth 34 / 101 classes
    args, aux = orc.param_names(34, 101)
    assert (len(args), len(aux)) == (211, 138)
    assert len(orc.all_convs(34)) == 69 and len(orc.all_convs(18)) == 37
    p = orc.init_params(34, 101)
    trainable = [k for k in p if "moving" not in k]
    assert len(trainable) == 209
    assert sum(int(np.prod(p[k].shape)) for k in trainable) == 63543788
    p18 = orc.init_params(18, 101)
    assert sum(int(np.prod(v.shape)) for k, v in p18.items() if "moving" not in k) == 33217452


def test_flop_count_matches_baseline_md():
    assert abs(orc.conv_flops(34, 32, 112, 112)[0] / 1e9 - 304.71) < 0.01
    assert abs(orc.conv_flops(34, 16, 112, 112)[0] / 1e9 - 152.36) < 0.01
    assert abs(orc.conv_flops(18, 8, 112, 112)[0] / 1e9 - 41.50) < 0.01


def test_structure_matches_reference_source():
    """Parameter order/shapes collected from the reference's R2Plus2D (run on the mxnet stand-in) zip exactly onto the
    oracle's symbol-name plan — the same zip load_from_sym_params does (model/R2Plus1.py:267-279)."""
    with open(os.path.join(GOLD, "structure_golden.json")) as fh:
        struct = json.load(fh)
    for depth in (18, 34):
        p = orc.init_params(depth, 101)
        got = [(n, list(p[n].shape)) for n, _ in struct[str(depth)]]
        assert got == [(n, s) for n, s in struct[str(depth)]]
        names = [n for n, _ in struct[str(depth)]]
        assert set(names) == set(p.keys())


@pytest.mark.parametrize("tag,depth,n,t,hw", [("small_r18_1x8x64", 18, 1, 8, 64), ("c1_r18_2x8x112", 18, 2, 8, 112),
                                              ("r34_1x16x64", 34, 1, 16, 64)])
def test_torch_engine_matches_reference_source_logits(tag, depth, n, t, hw):
    g = np.load(os.path.join(GOLD, "r2plus1d_golden.npz"))
    params = orc.randomize_bn(orc.init_params(depth, 101, seed=0), seed=1)
    assert abs(float(sum(np.abs(v).sum() for v in params.values())) - float(g[tag + "_param_checksum"][0])) < 1e-2 * float(g[tag + "_param_checksum"][0]) * 1e-3
    x = np.random.default_rng(123).random((n, 3, t, hw, hw), dtype=np.float32)
    pool = (t // 8, hw // 16, hw // 16)
    net = orc.Net(params, depth, pool)
    logits, pooled = net.forward(x)
    np.testing.assert_allclose(logits.numpy(), g[tag + "_eval_logits"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(pooled.reshape(n, -1).numpy(), g[tag + "_features"], rtol=1e-4, atol=1e-5)
    net = orc.Net(params, depth, pool)
    logits_t, _ = net.forward(x, train=True)
    np.testing.assert_allclose(logits_t.numpy(), g[tag + "_train_logits"], rtol=2e-3, atol=2e-4)
    # MXNet running-stat convention: momentum multiplies the OLD value, biased variance
    np.testing.assert_allclose(net.running["conv1_middle_spatbn_relu_moving_mean"].numpy(), g[tag + "_stem_bn_running_mean"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(net.running["conv1_middle_spatbn_relu_moving_var"].numpy(), g[tag + "_stem_bn_running_var"], rtol=1e-4, atol=1e-6)


def test_numpy_definition_matches_torch_engine_and_golden():
    """The plain-numpy operator definitions (the oracle proper) agree with the torch engine on a tiny net."""
    params = orc.randomize_bn(orc.init_params(10, 7, seed=3), seed=4)
    x = np.random.default_rng(5).random((2, 3, 8, 32, 32), dtype=np.float32)
    a, _ = orc.np_forward(params, x, 10, pool=(1, 2, 2))
    b, _ = orc.Net(params, 10, (1, 2, 2)).forward(x)
    np.testing.assert_allclose(a, b.numpy(), rtol=1e-4, atol=1e-5)
    a, _ = orc.np_forward(params, x, 10, pool=(1, 2, 2), train=True)
    b, _ = orc.Net(params, 10, (1, 2, 2)).forward(x, train=True)
    np.testing.assert_allclose(a, b.numpy(), rtol=2e-3, atol=2e-4)


def test_numpy_conv_gradients_match_autograd():
    import torch
    rng = np.random.default_rng(0)
    for k, s, p in (((1, 3, 3), (1, 2, 2), (0, 1, 1)), ((3, 1, 1), (2, 1, 1), (1, 0, 0)), ((1, 1, 1), (2, 2, 2), (0, 0, 0)),
                    ((1, 7, 7), (1, 2, 2), (0, 3, 3))):
        x = rng.normal(size=(2, 3, 4, 9, 10)).astype(np.float64)
        w = rng.normal(size=(5, 3) + k).astype(np.float64)
        y = orc.np_conv3d(x, w, s, p)
        xt = torch.tensor(x, requires_grad=True)
        wt = torch.tensor(w, requires_grad=True)
        yt = torch.nn.functional.conv3d(xt, wt, stride=s, padding=p)
        np.testing.assert_allclose(y, yt.detach().numpy(), rtol=1e-10, atol=1e-10)
        dy = rng.normal(size=y.shape)
        yt.backward(torch.tensor(dy))
        np.testing.assert_allclose(orc.np_conv3d_wgrad(x, dy, k, s, p), wt.grad.numpy(), rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(orc.np_conv3d_dgrad(dy, w, x.shape, s, p), xt.grad.numpy(), rtol=1e-9, atol=1e-9)


def test_batchnorm_conventions():
    rng = np.random.default_rng(1)
    x = rng.normal(1.0, 2.0, size=(3, 4, 2, 5, 5)).astype(np.float64)
    g, b = rng.uniform(0.5, 1.5, 4), rng.normal(size=4)
    rm, rv = np.zeros(4), np.ones(4)
    y, nrm, nrv, mean, inv = orc.np_batchnorm_train(x, g, b, rm, rv, eps=1e-5)
    m = x.size // 4
    np.testing.assert_allclose(nrm, 0.1 * x.mean(axis=(0, 2, 3, 4)))
    np.testing.assert_allclose(nrv, 0.9 + 0.1 * x.var(axis=(0, 2, 3, 4)))          # biased variance, not m/(m-1)
    # backward against finite differences of sum(y * r)
    r = rng.normal(size=x.shape)
    dx, dg, db = orc.np_batchnorm_backward(x, r, g, mean, inv)
    eps = 1e-6
    idx = (1, 2, 1, 3, 4)
    xp = x.copy(); xp[idx] += eps
    xm = x.copy(); xm[idx] -= eps
    f = lambda z: (orc.np_batchnorm_train(z, g, b, rm, rv, 1e-5)[0] * r).sum()
    assert abs((f(xp) - f(xm)) / (2 * eps) - dx[idx]) < 1e-5
    np.testing.assert_allclose(db, r.sum(axis=(0, 2, 3, 4)))


def test_xavier_and_sgd_conventions():
    w = orc.xavier_uniform(np.random.default_rng(0), (64, 32, 1, 3, 3))
    s = np.sqrt(3.0 / ((32 * 9 + 64 * 9) / 2.0))
    assert np.abs(w).max() <= s and np.abs(w).max() > 0.98 * s
    w2 = orc.xavier_uniform(np.random.default_rng(0), (64, 32, 1, 3, 3), "in", 2.34)
    assert np.abs(w2).max() <= np.sqrt(2.34 / (32 * 9))
    wn, mn = orc.sgd_momentum_step(np.array([1.0]), np.array([4.0]), np.array([0.5]), lr=0.1, momentum=0.9, wd=0.01, rescale=0.25)
    assert np.allclose(mn, 0.9 * 0.5 - 0.1 * (0.25 * 4.0 + 0.01 * 1.0)) and np.allclose(wn, 1.0 + mn)
