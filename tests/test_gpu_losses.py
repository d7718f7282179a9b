"""GPU parity of the loss kernels (through the C ABI / the reference-named classes) against the CPU oracle and the
golden outputs of the reference's own model/mlc_loss.py source.  Tolerance: 1e-5 relative on loss values
(north-star), sampled WARP ranks bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import mlc_loss as orl

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
with open(os.path.join(GOLD, "mlc_loss_golden.json")) as fh:
    CASES = json.load(fh)


def _dev(a, device):
    return torch.tensor(np.asarray(a, np.float32), device=device)


@pytest.mark.parametrize("name", sorted(CASES))
def test_lsep_kernels_match_reference_source(cuda_device, name):
    from fastvideotagging_b200.model import LsepLoss, LSEP_funcLoss
    c = CASES[name]
    for key, cls in (("lsep", LsepLoss), ("lsep_func", LSEP_funcLoss)):
        if key not in c:
            continue
        p = _dev(c["pred"], cuda_device).requires_grad_(True)
        loss = cls()(p, _dev(c["target"], cuda_device))
        assert tuple(loss.shape) == (1,)
        loss.backward()
        assert abs(loss.item() - c[key]["loss"]) <= 1e-5 * abs(c[key]["loss"])
        np.testing.assert_allclose(p.grad.cpu().numpy(), c[key]["grad"], rtol=2e-4, atol=1e-5)


@pytest.mark.parametrize("name", sorted(CASES))
def test_warp_kernels_match_reference_source(cuda_device, name):
    from fastvideotagging_b200.model import WarpLoss, WARP_funcLoss
    c = CASES[name]
    for key, cls in (("warp", WarpLoss), ("warp_func", WARP_funcLoss)):
        w = c[key]
        crit = cls(label_size=w["label_size"], seed=w["seed"], sample_offset=w["sample_offset"])
        p = _dev(c["pred"], cuda_device).requires_grad_(True)
        loss = crit(p, _dev(c["target"], cuda_device))
        loss.backward()
        assert (crit.last_trials.cpu().numpy() == np.array(w["trials"])).all()          # bit-exact sampled ranks
        assert abs(loss.item() - w["loss"]) <= 1e-5 * abs(w["loss"])
        np.testing.assert_allclose(p.grad.cpu().numpy(), w["grad"], rtol=1e-5, atol=1e-5)
        assert crit.sample_offset == w["sample_offset"] + len(c["pred"])


def test_warp_sampling_is_shard_invariant(cuda_device):
    """Ranks depend on the GLOBAL sample index: two shards of 8 reproduce one batch of 16."""
    from fastvideotagging_b200.model import WarpLoss
    c = CASES["meitu_16x63"]
    p, t = _dev(c["pred"], cuda_device), _dev(c["target"], cuda_device)
    full = WarpLoss(label_size=63, seed=9)
    full(p, t)
    a = WarpLoss(label_size=63, seed=9, sample_offset=0)
    b = WarpLoss(label_size=63, seed=9, sample_offset=8)
    a(p[:8], t[:8])
    b(p[8:], t[8:])
    assert torch.equal(full.last_trials, torch.cat([a.last_trials, b.last_trials]))
    L, trials = orl.warp_sample(np.array(c["pred"], np.float32), np.array(c["target"]), 62, orl.rank_weights(63), 9, 0)
    assert (full.last_trials.cpu().numpy() == trials).all()
    assert np.array_equal(full.last_rank.cpu().numpy(), L)


def test_warp_statistics_match_uniform_sampling(cuda_device):
    """Agreement with numpy's uniform negative sampling is statistical: mean #trials for a positive that is violated
    by a fraction f of the negatives is ~1/f."""
    from fastvideotagging_b200.model import WarpLoss
    B, C = 512, 63
    pred = np.zeros((B, C), np.float32)
    target = np.zeros((B, C), np.float32)
    target[:, 0] = 1
    pred[:, 0] = 0.5
    pred[:, 1:32] = 1.0                       # 31 of 62 negatives violate -> f = 0.5
    crit = WarpLoss(label_size=63, seed=1)
    crit(_dev(pred, cuda_device), _dev(target, cuda_device))
    tr = crit.last_trials[:, 0].float().mean().item()
    assert 1.8 < tr < 2.2


def test_edge_cases(cuda_device):
    from fastvideotagging_b200.model import LsepLoss, WarpLoss
    # empty positive set -> S = 0, loss = log(1) = 0, zero gradient
    p = torch.randn(3, 10, device=cuda_device, requires_grad=True)
    loss = LsepLoss()(p, torch.zeros(3, 10, device=cuda_device))
    loss.backward()
    assert loss.item() == 0.0 and float(p.grad.abs().max()) == 0.0
    # rows without negatives poison the WARP value (the reference would never return)
    loss = WarpLoss(label_size=10)(torch.randn(2, 10, device=cuda_device), torch.ones(2, 10, device=cuda_device))
    assert np.isnan(loss.item())
    # default label_size=62 with 63 classes in WARP_funcLoss indexes past the table (IndexError in the reference)
    from fastvideotagging_b200.model import WARP_funcLoss
    from fastvideotagging_b200._lib import FvtError
    with pytest.raises(FvtError, match="rank_weights"):
        WARP_funcLoss(label_size=62)(torch.randn(2, 63, device=cuda_device), torch.zeros(2, 63, device=cuda_device))


def test_bce_and_softmax_heads(cuda_device):
    from fastvideotagging_b200.model import SigmoidBinaryCrossEntropyLoss, SoftmaxCrossEntropyLoss, SoftmaxOutput
    rng = np.random.default_rng(0)
    x = rng.normal(0, 2, (16, 101)).astype(np.float32)
    z = (rng.random((16, 101)) < 0.03).astype(np.float32)
    for from_sigmoid in (False, True):
        xin = 1 / (1 + np.exp(-x)) if from_sigmoid else x
        p = _dev(xin, cuda_device).requires_grad_(True)
        loss = SigmoidBinaryCrossEntropyLoss(from_sigmoid=from_sigmoid)(p, _dev(z, cuda_device))
        assert tuple(loss.shape) == (16,)
        loss.sum().backward()
        rl, rg = orl.sigmoid_bce(xin, z, from_sigmoid)
        np.testing.assert_allclose(loss.detach().cpu().numpy(), rl, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(p.grad.cpu().numpy(), rg, rtol=1e-4, atol=1e-7)
    lab = rng.integers(0, 101, 16).astype(np.float32)
    p = _dev(x, cuda_device).requires_grad_(True)
    loss = SoftmaxCrossEntropyLoss()(p, _dev(lab, cuda_device))
    loss.sum().backward()
    rl, rg = orl.softmax_ce(x, lab)
    np.testing.assert_allclose(loss.detach().cpu().numpy(), rl, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(p.grad.cpu().numpy(), rg, rtol=1e-4, atol=1e-6)
    lab[3] = -1
    p = _dev(x, cuda_device).requires_grad_(True)
    prob = SoftmaxOutput(p, _dev(lab, cuda_device))
    prob.backward(torch.ones_like(prob))
    rp, rg = orl.softmax_output(x, lab)
    np.testing.assert_allclose(prob.detach().cpu().numpy(), rp, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(p.grad.cpu().numpy(), rg, rtol=1e-4, atol=1e-6)
    assert float(p.grad[3].abs().max()) == 0.0
