"""Generates the golden fixtures in this directory by EXECUTING THE REFERENCE'S OWN SOURCE
(/root/reference/model/R2Plus1.py and /root/reference/model/mlc_loss.py, unmodified, loaded by file path) on the
torch-CPU-backed `mxnet` stand-in in tests/golden/mxnet_shim (MXNet itself cannot be installed in this image).

Run here (the container that has /root/reference); the outputs are committed:
    python tests/golden/make_golden.py
        -> tests/golden/mlc_loss_golden.json, tests/golden/r2plus1d_golden.npz, tests/golden/structure_golden.json
Nothing under tests/ reads /root/reference at test time.
"""
import contextlib
import importlib.util
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "mxnet_shim"))
sys.path.insert(0, ROOT)
REF = "/root/reference"

import mxnet  # noqa: E402  (the shim)
from mxnet import nd, autograd  # noqa: E402
from oracle import r2plus1d as orc  # noqa: E402
from oracle import mlc_loss as orl  # noqa: E402


def load_ref(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod


class PhiloxChooser:
    """Drop-in for np.random.choice inside WarpLoss/WARP_funcLoss that follows the repo's sampling contract
    (counter = (sample_offset + row, class j, trial, 0), key = seed) while the reference's own loop runs."""

    def __init__(self, pred, target, max_trials, seed, offset):
        self.pred, self.target, self.max_trials = pred, target, max_trials
        self.key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        self.offset = offset
        self.todo = [(b, j) for b in range(target.shape[0]) for j in range(target.shape[1]) if target[b, j] == 1]
        self.cur, self.trial = 0, 0
        self.trials = np.zeros(target.shape, np.int32)

    def __call__(self, candidates, replace=False):
        b, j = self.todo[self.cur]
        self.trial += 1
        u = orl.philox4x32_10(((self.offset + b) & 0xFFFFFFFF, j, self.trial, 0), self.key)[0]
        neg = candidates[u % len(candidates)]
        margin = self.pred[b, neg] - self.pred[b, j]
        if margin >= 0 or self.trial >= self.max_trials:
            self.trials[b, j] = self.trial
            self.cur += 1
            self.trial = 0
        return neg


def loss_cases():
    cases = [("readme_example_2x4",
              np.array([[0.9, 0.4, 0.5, 0.2], [0.1, 0.6, 0.2, 0.8]], np.float32),      # mlc_loss.py:248-249
              np.array([[1, 1, 0, 0], [0, 1, 0, 1]], np.float32))]
    rng = np.random.default_rng(7)
    for name, B, C in (("meitu_16x63", 16, 63), ("ucf_4x101", 4, 101), ("ragged_3x10", 3, 10)):
        pred = rng.normal(0, 1.0, (B, C)).astype(np.float32)
        target = np.zeros((B, C), np.float32)
        for b in range(B):
            k = int(rng.integers(1, 5))                       # 1-4 tags per clip (simple_meitu.py:134-136)
            target[b, rng.choice(C, size=k, replace=False)] = 1
        cases.append((name, pred, target))
    return cases


def make_loss_golden(ml):
    out = {}
    for name, pred, target in loss_cases():
        B, C = pred.shape
        rec = {"pred": pred.tolist(), "target": target.tolist()}
        # LsepLoss (block) + autograd gradient
        p = nd.array(pred); p.attach_grad()
        with autograd.record():
            loss = ml.LsepLoss()(p, nd.array(target))
        loss.backward()
        rec["lsep"] = {"loss": float(loss.asnumpy().reshape(-1)[0]), "grad": p.grad.asnumpy().tolist()}
        # LSEP_funcLoss as written (only defined when #positives per row <= batch: row index = enumerate counter)
        if int(target.sum(axis=1).max()) <= B:
            p = nd.array(pred); p.attach_grad()
            with autograd.record():
                loss = ml.LSEP_funcLoss()(p, nd.array(target))
            loss.backward(mxnet.nd.ones_like(loss))
            rec["lsep_func"] = {"loss": float(loss.asnumpy().reshape(-1)[0]), "grad": p.grad.asnumpy().tolist()}
        # WarpLoss with the contract RNG plugged into the reference's np.random.choice call
        label_size = C
        seed, offset = 123, 5
        chooser = PhiloxChooser(pred, target, label_size - 1, seed, offset)
        saved = ml.np.random.choice
        ml.np.random.choice = chooser
        try:
            p = nd.array(pred); p.attach_grad()
            with autograd.record():
                loss = ml.WarpLoss(label_size=label_size)(p, nd.array(target))
            loss.backward()
        finally:
            ml.np.random.choice = saved
        rec["warp"] = {"label_size": label_size, "max_trials": label_size - 1, "seed": seed, "sample_offset": offset,
                       "loss": float(loss.asnumpy().reshape(-1)[0]), "grad": p.grad.asnumpy().tolist(),
                       "trials": chooser.trials.tolist()}
        # WARP_funcLoss (max_trials = C - 1, table of label_size entries)
        chooser = PhiloxChooser(pred, target, C - 1, seed, offset)
        ml.np.random.choice = chooser
        try:
            p = nd.array(pred); p.attach_grad()
            with autograd.record():
                loss = ml.WARP_funcLoss(label_size=C)(p, nd.array(target))
            loss.backward()
        finally:
            ml.np.random.choice = saved
        rec["warp_func"] = {"label_size": C, "max_trials": C - 1, "seed": seed, "sample_offset": offset,
                            "loss": float(loss.asnumpy().reshape(-1)[0]), "grad": p.grad.asnumpy().tolist(),
                            "trials": chooser.trials.tolist()}
        out[name] = rec
    with open(os.path.join(HERE, "mlc_loss_golden.json"), "w") as fh:
        json.dump(out, fh)
    return out


def build_ref_net(R, depth, num_class, pool, params):
    """Instantiate the reference's R2Plus2D and load `params` through its own load_from_sym_params."""
    mxnet._name_counters.clear()
    with contextlib.redirect_stdout(io.StringIO()):
        net = R.R2Plus2D(num_class=num_class, model_depth=depth, final_spatial_kernel=pool[1], final_temporal_kernel=pool[0])
    tmp = os.path.join(HERE, "_tmp_params.npz")
    np.savez(tmp, **{"arg:" + k: v for k, v in params.items()})
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            net.load_from_sym_params(tmp, with_dense=True)
    finally:
        os.remove(tmp)
    return net


def make_net_golden(R):
    out = {}
    struct = {}
    for depth in (18, 34):
        params = orc.init_params(depth, 101, seed=0)
        net = build_ref_net(R, depth, 101, (1, 7, 7), params)
        order = []
        for child in net._children.values():
            if child.name == "pool0":
                continue
            keys = list(child.collect_params().keys())
            names = getattr(net, child.name + "_name")
            assert len(keys) == len(names), (child.name, len(keys), len(names))
            for k, nme in zip(keys, names):
                order.append([nme, list(child.collect_params()[k]._data.shape)])
        struct[str(depth)] = order
    with open(os.path.join(HERE, "structure_golden.json"), "w") as fh:
        json.dump(struct, fh)

    for tag, depth, n, t, hw in (("c1_r18_2x8x112", 18, 2, 8, 112), ("small_r18_1x8x64", 18, 1, 8, 64), ("r34_1x16x64", 34, 1, 16, 64)):
        pool = (t // 8, hw // 16, hw // 16)
        params = orc.randomize_bn(orc.init_params(depth, 101, seed=0), seed=1)
        x = np.random.default_rng(123).random((n, 3, t, hw, hw), dtype=np.float32)
        net = build_ref_net(R, depth, 101, pool, params)
        y_eval = net(nd.array(x)).asnumpy()
        feat = net.extract_features(nd.array(x)).asnumpy()
        net = build_ref_net(R, depth, 101, pool, params)
        with autograd.record():
            y_train = net(nd.array(x)).asnumpy()
        rm = net.base._children["1"]._own_params["running_mean"]._data.numpy()
        rv = net.base._children["1"]._own_params["running_var"]._data.numpy()
        out[tag + "_eval_logits"] = y_eval
        out[tag + "_features"] = feat.reshape(n, -1)
        out[tag + "_train_logits"] = y_train
        out[tag + "_stem_bn_running_mean"] = rm
        out[tag + "_stem_bn_running_var"] = rv
        out[tag + "_param_checksum"] = np.array([float(sum(np.abs(v).sum() for v in params.values()))])
    np.savez_compressed(os.path.join(HERE, "r2plus1d_golden.npz"), **out)
    return out


if __name__ == "__main__":
    ml = load_ref("ref_mlc_loss", "model/mlc_loss.py")
    R = load_ref("ref_R2Plus1", "model/R2Plus1.py")
    lg = make_loss_golden(ml)
    print("loss golden:", {k: (round(v["lsep"]["loss"], 6), round(v["warp"]["loss"], 4)) for k, v in lg.items()})
    ng = make_net_golden(R)
    print("net golden keys:", sorted(ng)[:6], "...")
