"""CPU tests: oracle/mlc_loss.py against golden outputs of the reference's own model/mlc_loss.py source."""
import json
import os

import numpy as np
import pytest

from oracle import mlc_loss as orl

GOLD = os.path.join(os.path.dirname(__file__), "golden")
with open(os.path.join(GOLD, "mlc_loss_golden.json")) as fh:
    CASES = json.load(fh)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert orl.philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orl.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orl.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_survey_known_answers():
    c = CASES["readme_example_2x4"]
    loss, grad = orl.lsep_loss(c["pred"], c["target"], np.float64)
    assert abs(loss - 1.8583357044738) < 1e-6
    np.testing.assert_allclose(grad, [[-0.1819578, -0.2999977, 0.27685574, 0.20509977],
                                      [0.172011, -0.1991018, 0.19010156, -0.16301076]], atol=1e-6)
    loss, grad = orl.lsep_func_loss(c["pred"], c["target"], np.float64)
    assert abs(loss - 2.0615215945708) < 1e-6
    np.testing.assert_allclose(grad, [[-0.39443648, -0.65031581, 1.5995188, 1.18495267],
                                      [1.07218951, -0.53243356, 1.18495267, -0.43591973]], atol=1e-6)
    rw = orl.rank_weights(62)
    assert abs(rw[61] - 4.712392887832752) < 1e-12 and rw[1] == 1.5
    assert [61 // n for n in range(1, 7)] == [61, 30, 20, 15, 12, 10]


@pytest.mark.parametrize("name", sorted(CASES))
def test_lsep_matches_reference_source(name):
    c = CASES[name]
    loss, grad = orl.lsep_loss(c["pred"], c["target"])
    assert abs(loss - c["lsep"]["loss"]) <= 1e-5 * abs(c["lsep"]["loss"])
    np.testing.assert_allclose(grad, c["lsep"]["grad"], rtol=1e-4, atol=1e-6)
    if "lsep_func" in c:
        loss, grad = orl.lsep_func_loss(c["pred"], c["target"])
        assert abs(loss - c["lsep_func"]["loss"]) <= 1e-5 * abs(c["lsep_func"]["loss"])
        np.testing.assert_allclose(grad, c["lsep_func"]["grad"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", sorted(CASES))
def test_warp_matches_reference_source(name):
    c = CASES[name]
    pred, target = np.array(c["pred"], np.float32), np.array(c["target"], np.float32)
    for key, fn in (("warp", orl.warp_loss), ("warp_func", orl.warp_func_loss)):
        w = c[key]
        L, trials = orl.warp_sample(pred, target, w["max_trials"], orl.rank_weights(w["label_size"]), w["seed"], w["sample_offset"])
        assert (trials == np.array(w["trials"])).all()                      # sampled ranks are bit-exact
        loss, grad = fn(pred, target, L)
        assert abs(loss - w["loss"]) <= 1e-5 * abs(w["loss"])
        np.testing.assert_allclose(grad, w["grad"], rtol=1e-5, atol=1e-5)


def test_warp_quirks():
    # a positive that never violates still gets rank_weights[1] = 1.5 (mlc_loss.py:144-147)
    pred = np.array([[10.0, -5.0, -6.0, -7.0]], np.float32)
    target = np.array([[1, 0, 0, 0]], np.float32)
    L, trials = orl.warp_sample(pred, target, 3, orl.rank_weights(4))
    assert trials[0, 0] == 3 and L[0, 0] == np.float32(1.5)
    # masked-out pairs contribute relu(1) = 1 each (constant offset)
    loss, grad = orl.warp_loss(pred, target, L)
    assert abs(loss - 1.5 * 1.0) < 1e-6 and np.all(grad == 0)            # 3 inactive hinge terms (=0) + 1 masked pair (=1)
    # a row without negatives never terminates in the reference; the contract marks it NaN
    L, _ = orl.warp_sample(np.zeros((1, 3), np.float32), np.ones((1, 3), np.float32), 2, orl.rank_weights(3))
    assert np.isnan(L).all()


def test_bce_and_softmax_closed_forms():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(5, 7)).astype(np.float64)
    z = (rng.random((5, 7)) < 0.3).astype(np.float64)
    loss, grad = orl.sigmoid_bce(x, z, False, np.float64)
    p = 1 / (1 + np.exp(-x))
    np.testing.assert_allclose(loss, -(z * np.log(p) + (1 - z) * np.log(1 - p)).mean(axis=1), rtol=1e-10)
    loss2, _ = orl.sigmoid_bce(p, z, True, np.float64)
    np.testing.assert_allclose(loss2, loss, rtol=1e-8)
    lab = np.array([0, 3, 6, 2, 2])
    l, g = orl.softmax_ce(x, lab, np.float64)
    prob, g2 = orl.softmax_output(x, lab.astype(np.float64), dtype=np.float64)
    np.testing.assert_allclose(l, -np.log(prob[np.arange(5), lab]), rtol=1e-10)
    np.testing.assert_allclose(g, g2, rtol=1e-10)
    _, g3 = orl.softmax_output(x, np.array([-1, 3, 6, 2, 2.0]), dtype=np.float64)
    assert np.all(g3[0] == 0)
