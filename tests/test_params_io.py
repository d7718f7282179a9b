"""Weight-file formats either side of the hot path: MXNet NDArray-dict files and the Caffe2 pickle
mapping.  CPU only."""
import logging
import pickle
import struct

import numpy as np
import pytest

from fastvideotagging_b200 import params_io
from fastvideotagging_b200 import engine


def test_nd_file_bytes_follow_the_mxnet_list_layout(tmp_path):
    """Hand-assembled expected bytes (MXNet 1.x NDArray::Save list form, V2 records) for a two-entry dict."""
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    b = np.array([7, 8], dtype=np.int32)
    f = tmp_path / "two.params"
    params_io.nd_save(str(f), {"arg:w": a, "aux:m": b})
    exp = struct.pack("<QQQ", 0x112, 0, 2)
    exp += struct.pack("<Ii", 0xF993FAC9, 0) + struct.pack("<I", 2) + struct.pack("<qq", 2, 3) + struct.pack("<ii", 1, 0)
    exp += struct.pack("<i", 0) + a.tobytes()
    exp += struct.pack("<Ii", 0xF993FAC9, 0) + struct.pack("<I", 1) + struct.pack("<q", 2) + struct.pack("<ii", 1, 0)
    exp += struct.pack("<i", 4) + b.tobytes()
    exp += struct.pack("<Q", 2) + struct.pack("<Q", 5) + b"arg:w" + struct.pack("<Q", 5) + b"aux:m"
    assert f.read_bytes() == exp
    back = params_io.nd_load(str(f))
    assert list(back) == ["arg:w", "aux:m"]
    assert back["arg:w"].dtype == np.float32 and np.array_equal(back["arg:w"], a)
    assert back["aux:m"].dtype == np.int32 and np.array_equal(back["aux:m"], b)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.float16, np.uint8, np.int32, np.int8, np.int64])
def test_nd_roundtrip_all_type_flags(tmp_path, dtype):
    rng = np.random.default_rng(0)
    arrs = {"x": (rng.random((3, 1, 4)) * 50).astype(dtype), "empty": np.zeros((0, 5), dtype), "one": np.ones((1,), dtype)}
    f = str(tmp_path / "t.params")
    params_io.nd_save(f, arrs)
    back = params_io.nd_load(f)
    for k in arrs:
        assert back[k].dtype == np.dtype(dtype) and back[k].shape == arrs[k].shape and np.array_equal(back[k], arrs[k])


def test_nd_list_without_names_and_older_record_versions(tmp_path):
    f = str(tmp_path / "l.params")
    params_io.nd_save(f, [np.ones((2, 2), np.float32), np.zeros(3, np.float32)])
    out = params_io.nd_load(f)
    assert isinstance(out, list) and out[0].shape == (2, 2) and out[1].shape == (3,)
    a = np.arange(4, dtype=np.float32)
    # V1 record (no storage-type word) and the pre-magic legacy record (uint32 dims)
    v1 = struct.pack("<I", 0xF993FAC8) + struct.pack("<I", 1) + struct.pack("<q", 4) + struct.pack("<iii", 1, 0, 0) + a.tobytes()
    legacy = struct.pack("<I", 2) + struct.pack("<II", 2, 2) + struct.pack("<iii", 2, 3, 0) + a.tobytes()
    blob = struct.pack("<QQQ", 0x112, 0, 2) + v1 + legacy + struct.pack("<Q", 2)
    blob += struct.pack("<Q", 1) + b"a" + struct.pack("<Q", 1) + b"b"
    g = tmp_path / "old.params"
    g.write_bytes(blob)
    out = params_io.nd_load(str(g))
    assert np.array_equal(out["a"], a) and out["b"].shape == (2, 2) and np.array_equal(out["b"].ravel(), a)


def test_nd_load_rejects_garbage_truncation_and_sparse(tmp_path):
    f = tmp_path / "bad.params"
    f.write_bytes(b"PK\x03\x04 not an mxnet file....")
    with pytest.raises(params_io.ParamsFormatError):
        params_io.nd_load(str(f))
    good = tmp_path / "good.params"
    params_io.nd_save(str(good), {"w": np.ones((4, 4), np.float32)})
    raw = good.read_bytes()
    f.write_bytes(raw[:-30])
    with pytest.raises(params_io.ParamsFormatError):
        params_io.nd_load(str(f))
    sparse = struct.pack("<QQQ", 0x112, 0, 1) + struct.pack("<Ii", 0xF993FAC9, 1)
    f.write_bytes(sparse + b"\0" * 64)
    with pytest.raises(params_io.ParamsFormatError):
        params_io.nd_load(str(f))


def test_checkpoint_prefix_convention(tmp_path):
    pshapes, ashapes = engine.parameter_shapes(18, 101)
    rng = np.random.default_rng(1)
    arg = {k: rng.standard_normal(s).astype(np.float32) for k, s in list(pshapes.items())[:6]}
    aux = {k: rng.random(s).astype(np.float32) for k, s in list(ashapes.items())[:4]}
    fname = params_io.save_checkpoint(str(tmp_path / "r3d"), 7, arg, aux)
    assert fname.endswith("r3d-0007.params")
    a2, x2 = params_io.load_checkpoint(str(tmp_path / "r3d"), 7)
    assert set(a2) == set(arg) and set(x2) == set(aux)
    for k in arg:
        assert np.array_equal(a2[k], arg[k])
    # unprefixed dicts are split by suffix
    a3, x3 = params_io.split_checkpoint({**arg, **aux})
    assert set(a3) == set(arg) and set(x3) == set(aux)


def _caffe2_blobs(depth, rng):
    """A Caffe2-style blob dict for R(2+1)D-`depth` pretrained on Kinetics: canonical names with Caffe2 suffixes plus the
    400-way head `last_out_L400_{w,b}` (what r2.5d_d34_l32.pkl holds, per the reference log)."""
    pshapes, ashapes = engine.parameter_shapes(depth, 101)
    blobs = {}
    for name, shape in pshapes.items():
        if name.startswith("final_fc"):
            continue
        for suf_mx, suf_c2 in (("_weight", "_w"), ("_beta", "_b"), ("_gamma", "_s")):
            if name.endswith(suf_mx):
                blobs[name[:-len(suf_mx)] + suf_c2] = rng.standard_normal(shape).astype(np.float32)
    for name, shape in ashapes.items():
        if name.endswith("_moving_mean"):
            blobs[name[:-len("_moving_mean")] + "_rm"] = rng.standard_normal(shape).astype(np.float32)
        else:
            blobs[name[:-len("_moving_var")] + "_riv"] = (0.5 + rng.random(shape)).astype(np.float32)
    blobs["last_out_L400_w"] = rng.standard_normal((400, 512)).astype(np.float32)
    blobs["last_out_L400_b"] = rng.standard_normal((400,)).astype(np.float32)
    return blobs


def test_caffe2_mapping_reproduces_the_reference_log(tmp_path, caplog):
    """Known answer: reference r2plus1d_output/log.txt:38-48 (R34, 101 classes, Kinetics pickle) — 349 = 211 arg + 138 aux
    in the symbol, 347 = 209 arg + 138 aux loaded, `data` / `final_fc_weight` / `final_fc_bias` / `softmax_label` not
    loaded, `last_out_L400_beta` / `last_out_L400_weight` not used, every aux loaded."""
    from fastvideotagging_b200.net import create_r3d
    rng = np.random.default_rng(3)
    blobs = _caffe2_blobs(34, rng)
    f = tmp_path / "r2.5d_d34_l32.pkl"
    with open(f, "wb") as fh:
        pickle.dump({"blobs": blobs}, fh, protocol=2)
    sym = create_r3d(num_class=101, no_bias=1, model_depth=34, final_spatial_kernel=7, final_temporal_kernel=4)
    with caplog.at_level(logging.INFO, logger="utils"):
        arg_p, aux_p = params_io.load_from_caffe2_pkl(str(f), sym)
    msgs = [r.getMessage() for r in caplog.records]
    assert "symbol has 349 = 211 arg + 138 aux" in msgs
    assert "model loaded has 347 = 209 arg + 138 aux" in msgs
    i0, i1, i2 = msgs.index("testing arg loaded"), msgs.index("testing arg used in net"), msgs.index("testing aux")
    assert msgs[i0 + 1:i1] == ["arg data not loaded", "arg final_fc_weight not loaded", "arg final_fc_bias not loaded",
                               "arg softmax_label not loaded"]
    assert sorted(msgs[i1 + 1:i2]) == ["arg last_out_L400_beta not used in net", "arg last_out_L400_weight not used in net"]
    assert msgs[i2 + 1:] == []
    assert len(arg_p) == 209 and len(aux_p) == 138
    # _riv is an inverse variance (utils.py:33)
    k = "comp_0_spatbn_1_middle"
    assert np.allclose(aux_p[k + "_moving_var"], 1.0 / blobs[k + "_riv"])
    assert np.array_equal(arg_p[k + "_gamma"], blobs[k + "_s"]) and np.array_equal(arg_p[k + "_beta"], blobs[k + "_b"])
    assert np.array_equal(arg_p["conv1_middle_weight"], blobs["conv1_middle_w"])


def test_block_save_load_and_caffe2_pickle_on_the_host_module(tmp_path):
    """R2Plus2D.save_parameters / load_parameters / load_from_sym_params / load_from_caffe2_pickle (host logic, CPU)."""
    from fastvideotagging_b200.model import R2Plus2D
    net = R2Plus2D(101, 18, final_spatial_kernel=7, final_temporal_kernel=1)
    net.initialize(seed=4)
    f = str(tmp_path / "net.params")
    net.save_parameters(f)
    assert params_io.is_nd_file(f)
    other = R2Plus2D(101, 18, final_spatial_kernel=7, final_temporal_kernel=1)
    other.load_parameters(f)
    for k, v in net.collect_params().items():
        assert np.array_equal(v.detach().numpy(), other.collect_params()[k].detach().numpy()), k
    # a Module-API checkpoint (arg:/aux: prefixes) through load_from_sym_params, dense layer skipped by default
    arg, aux = params_io.split_checkpoint({k: v.detach().numpy() for k, v in net.collect_params().items()})
    ck = params_io.save_checkpoint(str(tmp_path / "sym"), 1, arg, aux)
    third = R2Plus2D(101, 18, final_spatial_kernel=7, final_temporal_kernel=1)
    third.initialize(seed=9)
    fc_before = third.final_fc_weight.detach().clone()
    third.load_from_sym_params(ck)
    assert np.array_equal(third.conv1_weight.detach().numpy(), net.conv1_weight.detach().numpy())
    assert np.array_equal(third.final_fc_weight.detach().numpy(), fc_before.numpy())
    # Caffe2 pickle
    blobs = _caffe2_blobs(18, np.random.default_rng(5))
    pk = tmp_path / "r2.5d_d18.pkl"
    with open(pk, "wb") as fh:
        pickle.dump({"blobs": blobs}, fh, protocol=2)
    rep = third.load_from_caffe2_pickle(str(pk))
    assert rep["not_used"] == ["last_out_L400_beta", "last_out_L400_weight"]
    assert np.array_equal(third.comp_3_conv_1_middle_weight.detach().numpy(), blobs["comp_3_conv_1_middle_w"])
    assert np.allclose(third.conv1_spatbn_relu_moving_var.numpy(), 1.0 / blobs["conv1_spatbn_relu_riv"])
    assert np.array_equal(third.final_fc_weight.detach().numpy(), fc_before.numpy())
