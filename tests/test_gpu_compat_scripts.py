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
