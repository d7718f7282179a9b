"""The reference's OWN entry scripts, unmodified, on the B200 hot path (north_star: "train.py, train_simple_r3d.py and
validation.py work unchanged as a drop-in"; SURVEY 8b "Who calls it").

`compat/` provides what the scripts import (an `mxnet`-named module over fastvideotagging_b200, top-level `model`, `net`,
`utils`, `util`, `data` with synthetic clips).  The script files themselves are NOT part of this repository: a copy made by
`python compat/fetch_reference_scripts.py` (git-ignored `compat/_ref/`, taken verbatim from /root/reference) is executed when
present — it travels to the GPU box with the working-tree snapshot — otherwise the tests skip.  One process per GPU:
`--gpus 0`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "compat", "_ref")


def _run(script, args, cwd, clips):
    path = os.path.join(REF, script)
    if not os.path.exists(path):
        pytest.skip("no copy of the reference's %s (run compat/fetch_reference_scripts.py where /root/reference exists)" % script)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT]), FVT_COMPAT_CLIPS=str(clips))
    out = subprocess.run([sys.executable, "-W", "ignore", path] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    log = out.stdout + out.stderr
    assert out.returncode == 0, log[-4000:]
    return log


# (the script's ucf101 branch cannot run in the reference either: train_simple_r3d.py:70-78 replaces its SoftmaxCrossEntropyLoss
#  by the multi-label loss of --loss_type, which rejects the (B,) class-index labels)
@pytest.mark.parametrize("dataset,loss", [("meitu", "lsep_nn"), ("meitu", "warp_nn"), ("meitu", "lsep_fn"), ("meitu", "bce")])
def test_train_simple_r3d_main_loop_runs_unmodified(cuda_device, tmp_path, dataset, loss):
    """train_simple_r3d.py:26-208 with --debug: R(2+1)D-34 on 16x112x112 clips, gluon.Trainer('sgd'), two training
    iterations under autograd.record(), save_parameters, then the evaluation loop (top-k IoU / accuracy)."""
    out_dir = str(tmp_path / "out")
    # the script builds its save path as './{output}/...' (train_simple_r3d.py:135): the output directory is given relative to cwd
    log = _run("train_simple_r3d.py", ["--gpus", "0", "--dataset", dataset, "--loss_type", loss, "--num_epoch", "1", "--debug",
                                        "--log_interval", "2", "--batch_per_device", "2", "--output", "out", "--lr", "1e-4"],
               str(tmp_path), clips=12)
    assert "finished" in log and "training loss=" in log, log[-3000:]
    loss_lines = [ln for ln in log.splitlines() if "] training loss=" in ln and "Iter" not in ln]
    assert loss_lines, log[-3000:]
    value = float(loss_lines[-1].split("training loss=")[1].split()[0])
    if not (dataset == "meitu" and loss == "bce"):        # from_sigmoid=True on raw logits is NaN-prone in the reference too
        assert value == value and abs(value) < 1e9, loss_lines
    saved = [f for f in os.listdir(out_dir) if f.endswith(".params")]
    assert saved, os.listdir(out_dir)
    # the checkpoint is an MXNet NDArray file keyed by the symbol-API names
    from fastvideotagging_b200 import params_io
    arrays = params_io.nd_load(os.path.join(out_dir, saved[0]))
    assert "conv1_middle_weight" in arrays and arrays["comp_0_conv_1_middle_weight"].shape == (144, 64, 1, 3, 3)


def test_train_py_then_validation_py_run_unmodified(cuda_device, tmp_path):
    """train.py:14-94 (create_r3d symbol, mx.module.Module.fit, do_checkpoint) for one epoch of three batches, then
    validation.py:13-66 (load_checkpoint, Module.bind/set_params/forward/get_outputs, multi-clip averaging) on what it saved."""
    out_dir = str(tmp_path / "models")
    common = ["--gpus", "0", "--batch_per_device", "4", "--n_frame", "16", "--num_class", "101", "--datadir", "synthetic", "--output", "models"]
    log = _run("train.py", common + ["--pretrained", "", "--num_epoch", "1", "--model_depth", "34", "--lr", "1e-4"], str(tmp_path), clips=10)
    assert "Train-accuracy" in log, log[-3000:]
    assert os.path.exists(os.path.join(out_dir, "test-0001.params")) and os.path.exists(os.path.join(out_dir, "test-symbol.json"))
    log = _run("validation.py", common + ["--model_prefix", "test", "--eval_epoch", "1", "--clips_per_video", "2"], str(tmp_path), clips=10)
    lines = [ln for ln in log.splitlines() if ln.startswith("epoch ") and " acc " in ln]
    assert len(lines) == 2, log[-3000:]
    acc = float(lines[-1].split(" acc ")[1])
    assert 0.0 <= acc <= 1.0


_OWN_LOOP = r'''
# A training + evaluation loop written for this test against the SAME mxnet / gluon calls the reference's
# train_simple_r3d.py makes (gluon.Trainer, gluon.utils.split_and_load, autograd.record, nd.mean(..).asscalar(),
# L.backward(), trainer.step, save_parameters / load_parameters, lr scheduling) — so that the compat tree is exercised on
# the GPU box even where no copy of the reference's scripts is available.
import os, sys
import numpy as np
import mxnet as mx
from mxnet import gluon, nd, autograd
from model import R2Plus2D, LsepLoss, WarpLoss
from data import get_simple_meitu_dataloader

ctx = [mx.gpu(0)]
net = R2Plus2D(num_class=63, model_depth=18, final_temporal_kernel=2, final_spatial_kernel=7)
net.initialize(mx.init.Xavier(), ctx=ctx)
trainer = gluon.Trainer(net.collect_params(), 'sgd', {'learning_rate': 1e-4, 'momentum': 0.9, 'wd': 1e-4}, kvstore='device')
train_loader, val_loader = get_simple_meitu_dataloader(datadir='synthetic', batch_size=2, n_frame=16, crop_size=112,
                                                       scale_h=128, scale_w=171, num_workers=0)
losses = []
for crit in (LsepLoss(), WarpLoss(label_size=63)):
    for i, (data, label) in enumerate(train_loader):
        xs = gluon.utils.split_and_load(data, ctx_list=ctx, batch_axis=0)
        ys = gluon.utils.split_and_load(label, ctx_list=ctx, batch_axis=0)
        Ls = []
        with autograd.record():
            for x, y in zip(xs, ys):
                L = crit(net(x), y)
                Ls.append(L)
                losses.append(nd.mean(L).asscalar())
            for L in Ls:
                L.backward()
        trainer.step(data.shape[0])
        if i == 1:
            break
trainer.set_learning_rate(trainer.learning_rate * 0.1)
net.save_parameters(sys.argv[1])
net2 = R2Plus2D(num_class=63, model_depth=18, final_temporal_kernel=2, final_spatial_kernel=7)
net2.load_parameters(sys.argv[1], ctx=ctx)
for data, label in val_loader:
    x = gluon.utils.split_and_load(data, ctx_list=ctx, batch_axis=0)[0]
    a, b = net(x).asnumpy(), net2(x).asnumpy()
    assert np.array_equal(a, b), "a reloaded network must give the same eval logits"
    break
assert all(np.isfinite(v) for v in losses), losses
print("own-loop ok", len(losses), "iterations, lr now", trainer.learning_rate)
'''


def test_compat_tree_runs_a_gluon_training_loop(cuda_device, tmp_path):
    """The compat `mxnet` / `model` / `data` modules under a loop written here with the reference's call sequence
    (train_simple_r3d.py:95-137, 169-197): runs wherever the repo runs, independent of a copy of the reference's scripts."""
    script = tmp_path / "own_loop.py"
    script.write_text(_OWN_LOOP)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT]), FVT_COMPAT_CLIPS="8")
    out = subprocess.run([sys.executable, "-W", "ignore", str(script), str(tmp_path / "net.params")], cwd=str(tmp_path), env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout + out.stderr)[-4000:]
    assert "own-loop ok 4 iterations" in out.stdout, out.stdout[-2000:]


_OWN_MODULE_LOOP = r'''
# The symbol / Module API the reference's train.py and validation.py use (create_r3d, mx.module.Module.fit with
# do_checkpoint + Speedometer + FactorScheduler, load_checkpoint, bind / set_params / forward / get_outputs), driven by a loop
# written for this test.
import os, sys
import numpy as np
import mxnet as mx
from net import create_r3d
from data import ClipBatchIter

out = sys.argv[1]
os.makedirs(out, exist_ok=True)
net = create_r3d(num_class=101, no_bias=True, model_depth=18, final_spatial_kernel=7, final_temporal_kernel=2, bn_mom=0.9,
                 cudnn_tune='off', workspace=512)
m = mx.module.Module(net, context=[mx.gpu(0)])
train = mx.io.PrefetchingIter(ClipBatchIter(datadir='synthetic', batch_size=4, n_frame=16, crop_size=112, train=True))
val = mx.io.PrefetchingIter(ClipBatchIter(datadir='synthetic', batch_size=4, n_frame=16, crop_size=112, train=False, temporal_center=True))
m.fit(train_data=train, eval_data=val, eval_metric='accuracy',
      epoch_end_callback=mx.callback.do_checkpoint(out + '/own', 1), batch_end_callback=mx.callback.Speedometer(4, 1),
      kvstore=mx.kvstore.create('device'), optimizer='sgd',
      optimizer_params={'learning_rate': 1e-4, 'momentum': 0.9, 'wd': 1e-4,
                        'lr_scheduler': mx.lr_scheduler.FactorScheduler(step=2, factor=0.5)},
      initializer=mx.init.Xavier(factor_type='in', magnitude=2.34), arg_params={}, aux_params={}, allow_missing=True,
      begin_epoch=0, num_epoch=1)
sym, arg_params, aux_params = mx.model.load_checkpoint(out + '/own', 1)
m2 = mx.module.Module(sym, context=[mx.gpu(0)])
it = ClipBatchIter(datadir='synthetic', batch_size=4, n_frame=16, crop_size=112, train=False)
m2.bind(data_shapes=it.provide_data, label_shapes=it.provide_label, for_training=False)
m2.set_params(arg_params, aux_params, allow_missing=True)
n = 0
for batch in it:
    m2.forward(batch, is_train=False)
    probs = m2.get_outputs()[0].asnumpy()
    assert probs.shape == (4, 101) and np.isfinite(probs).all()
    assert np.allclose(probs.sum(axis=1), 1.0, atol=1e-3), "SoftmaxOutput returns probabilities"
    n += 1
print("own-module-loop ok", n, "validation batches")
'''


def test_compat_tree_runs_a_module_api_loop(cuda_device, tmp_path):
    """The compat `mxnet.module.Module` / `net.create_r3d` / `data.ClipBatchIter` surface under a loop written here with the
    reference's call sequence (train.py:35-94, validation.py:18-66): symbol network, fit with checkpoint callback and
    learning-rate schedule, reload through load_checkpoint, inference through bind / set_params / forward / get_outputs."""
    script = tmp_path / "own_module_loop.py"
    script.write_text(_OWN_MODULE_LOOP)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT]), FVT_COMPAT_CLIPS="8")
    out = subprocess.run([sys.executable, "-W", "ignore", str(script), str(tmp_path / "models")], cwd=str(tmp_path), env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout + out.stderr)[-4000:]
    assert "own-module-loop ok 2 validation batches" in out.stdout, out.stdout[-2000:]
    assert os.path.exists(str(tmp_path / "models" / "own-0001.params")) and os.path.exists(str(tmp_path / "models" / "own-symbol.json"))
