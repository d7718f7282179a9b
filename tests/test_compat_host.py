"""Host-side behaviour of the compat tree (compat/mxnet, compat/data) that needs no GPU: the NDArray wrapper, the
iterators' protocol and shapes, metrics, learning-rate schedules, checkpoint files, kvstore and context plumbing — the
parts of the MXNet surface the reference's scripts touch outside the hot path (train.py:23-94,
train_simple_r3d.py:37-137, validation.py:18-66, data/data.py:18-108)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r'''
import os, sys, tempfile
import numpy as np
import mxnet as mx
from mxnet import nd, gluon

# ---- NDArray surface (asscalar / asnumpy / argmax / slicing / arithmetic), on the CPU context
a = nd.array(np.arange(12, dtype=np.float32).reshape(3, 4))
assert a.shape == (3, 4) and a.dtype == np.float32 and a.context.device_type == "cpu"
assert abs(nd.mean(a).asscalar() - 5.5) < 1e-6
assert (a.argmax(axis=1).asnumpy() == 3).all()
assert np.allclose((a[1:3] * 2 + 1).asnumpy(), np.arange(12, dtype=np.float32).reshape(3, 4)[1:3] * 2 + 1)
assert np.allclose(nd.softmax(a).asnumpy().sum(axis=1), 1.0, atol=1e-6)
assert nd.concat(a, a, dim=1).shape == (3, 8) and nd.zeros((2, 3)).asnumpy().sum() == 0 and nd.ones((2,)).asnumpy().sum() == 2

# ---- contexts and batch splitting as train_simple_r3d.py:110-111 uses them
assert mx.gpu(1).device_id == 1 and mx.cpu().device_type == "cpu"
try:                                          # the hot path has no CPU fallback: a CPU context list is refused, loudly
    gluon.utils.split_and_load(nd.array(np.zeros((4, 2), np.float32)), ctx_list=[mx.cpu()], batch_axis=0)
    raise AssertionError("split_and_load onto a CPU context must raise")
except RuntimeError as e:
    assert "no CPU fallback" in str(e)

# ---- data iterators (synthetic clips): Module-API iterator and gluon loaders
from data import ClipBatchIter, get_ucf101trainval, get_simple_meitu_dataloader
it = ClipBatchIter(datadir="synthetic", batch_size=3, n_frame=4, crop_size=16, train=False)
desc = it.provide_data[0]
assert desc.name == "data" and desc.shape == (3, 3, 4, 16, 16) and it.provide_label[0].shape == (3,)
batches = list(it)
assert len(batches) == (len(it.clip_lst) + 2) // 3
assert batches[0].data[0].shape == (3, 3, 4, 16, 16) and batches[0].label[0].shape == (3,)
x0 = batches[0].data[0].asnumpy()
assert abs(float(x0.mean())) < 0.2 and 0.5 < float(x0.std()) < 1.5            # per-batch normalisation, videos_reader.py:93-97
it.reset()
assert np.array_equal(next(iter(it)).data[0].asnumpy(), x0)                     # test split: same clips after reset
pre = mx.io.PrefetchingIter(it)
assert pre.provide_data[0].shape == desc.shape and pre.batch_size == 3
tr, va = get_simple_meitu_dataloader(datadir="synthetic", batch_size=2, n_frame=4, crop_size=16, scale_h=20, scale_w=24, num_workers=0)
data, label = next(iter(va))
assert data.shape == (2, 3, 4, 16, 16) and label.shape == (2, 63)
lab = label.asnumpy()
assert set(np.unique(lab)) <= {0.0, 1.0} and (lab.sum(axis=1) >= 1).all() and (lab.sum(axis=1) <= 4).all()   # 1-4 tags, data/simple_meitu.py:134-136

# ---- metric, schedules, kvstore
m = mx.metric.create("accuracy")
m.update([nd.array(np.array([1, 0, 2], np.float32))], [nd.array(np.eye(3, dtype=np.float32)[[1, 0, 0]])])
assert m.get() == ("accuracy", 2.0 / 3.0)
s = mx.lr_scheduler.FactorScheduler(step=2, factor=0.5)
s.base_lr = 1.0
assert [s(i) for i in (1, 2, 3, 5)] == [1.0, 1.0, 0.5, 0.25]
kv = mx.kvstore.create("device")
assert kv.type == "device" and kv.num_workers == 1 and kv.rank == 0

# ---- NDArray files: nd.save / nd.load round trip through the MXNet container format
path = os.path.join(tempfile.mkdtemp(), "x.params")
nd.save(path, {"arg:w": a, "aux:m": nd.array(np.ones(5, np.float32))})
back = nd.load(path)
assert set(back) == {"arg:w", "aux:m"} and np.array_equal(back["arg:w"].asnumpy(), a.asnumpy())
print("compat host ok")
'''


def test_compat_host_surface_without_a_gpu(tmp_path):
    script = tmp_path / "compat_host.py"
    script.write_text(_SCRIPT)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT]), FVT_COMPAT_CLIPS="7", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-W", "ignore", str(script)], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout + out.stderr)[-4000:]
    assert "compat host ok" in out.stdout
