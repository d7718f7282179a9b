"""Weight files either side of the hot path: MXNet `.params` NDArray-dict files and the Caffe2
R(2+1)D pickle the reference fine-tunes from.

* `nd_save` / `nd_load` — the binary container `mx.nd.save` / `mx.nd.load` read and write, i.e. what
  `mx.callback.do_checkpoint` (reference train.py:83 -> `prefix-0001.params`, keys `arg:<name>` / `aux:<name>`),
  `mx.model.load_checkpoint` (validation.py:22) and gluon `save_parameters` / `load_parameters`
  (train_simple_r3d.py:92,137,225) exchange.  MXNet is a third-party dependency that is not vendored under the
  reference and cannot be installed here, so the layout is restated from MXNet 1.x `src/ndarray/ndarray.cc`
  (`NDArray::Save/Load`, list form) and is **unpinned against a file written by a real MXNet**:

      uint64  0x112                      kMXAPINDArrayListMagic
      uint64  0                          reserved
      uint64  n_arrays
      n_arrays x { uint32 0xF993FAC9     NDARRAY_V2_MAGIC (V1 = ...C8: no stype field; V3 = ...CA: numpy shape semantics)
                   int32  stype          0 = dense (the only kind handled; sparse raises)
                   uint32 ndim, int64 dims[ndim]
                   int32  dev_type, int32 dev_id      (1, 0) = cpu(0)
                   int32  type_flag      0 f32, 1 f64, 2 f16, 3 u8, 4 i32, 5 i8, 6 i64
                   raw little-endian data, C order }
      uint64  n_names
      n_names x { uint64 len, bytes }

  The pre-magic legacy layout (`uint32 ndim, uint32 dims[ndim]` first) is accepted on load.

* `caffe2_blobs_to_params` / `load_from_caffe2_pkl` — the blob-name mapping of reference utils.py:13-55
  (`_w -> _weight`, `_b -> _beta`, `_s -> _gamma`, `_rm -> _moving_mean`, `_riv -> _moving_var = 1 / riv`) with the
  same "not loaded" / "not used" report the reference logs (r2plus1d_output/log.txt:38-48 is the known answer: 209 arg
  + 138 aux loaded, `final_fc_weight` / `final_fc_bias` not loaded, `last_out_L400_{beta,weight}` not used).

Host-side format code only: nothing here touches the GPU.
"""
import logging
import pickle
import struct

import numpy as np

logger = logging.getLogger("utils")

LIST_MAGIC = 0x112
NDARRAY_V1_MAGIC = 0xF993FAC8
NDARRAY_V2_MAGIC = 0xF993FAC9
NDARRAY_V3_MAGIC = 0xF993FACA

_TYPE_FLAGS = {0: np.float32, 1: np.float64, 2: np.float16, 3: np.uint8, 4: np.int32, 5: np.int8, 6: np.int64}
_FLAG_OF = {np.dtype(v): k for k, v in _TYPE_FLAGS.items()}


class ParamsFormatError(ValueError):
    pass


def _to_numpy(v):
    if hasattr(v, "detach"):                     # torch tensor (possibly on the GPU, possibly bf16)
        v = v.detach().cpu()
        if str(v.dtype) == "torch.bfloat16":
            v = v.float()
        v = v.numpy()
    a = np.asarray(v)
    if a.dtype not in _FLAG_OF:
        if a.dtype.kind == "f":
            a = a.astype(np.float32)
        elif a.dtype.kind in "iub":
            a = a.astype(np.int64)
        else:
            raise ParamsFormatError("dtype %s has no MXNet type flag" % a.dtype)
    return np.ascontiguousarray(a)


def _pack_array(a):
    a = _to_numpy(a)
    out = [struct.pack("<Ii", NDARRAY_V2_MAGIC, 0), struct.pack("<I", a.ndim), struct.pack("<%dq" % a.ndim, *a.shape)]
    if a.ndim == 0:
        # MXNet 1.x (V2) has no 0-d arrays: ndim 0 means "none"; scalars are stored as shape (1,)
        a = a.reshape(1)
        out = [struct.pack("<Ii", NDARRAY_V2_MAGIC, 0), struct.pack("<I", 1), struct.pack("<q", 1)]
    out.append(struct.pack("<ii", 1, 0))
    out.append(struct.pack("<i", _FLAG_OF[a.dtype]))
    out.append(a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes())
    return b"".join(out)


def nd_save(fname, data):
    """`mx.nd.save(fname, data)`: `data` is a dict {name: array} or a list of arrays (saved without names)."""
    if isinstance(data, dict):
        names, arrays = list(data.keys()), list(data.values())
    else:
        names, arrays = [], list(data)
    chunks = [struct.pack("<QQQ", LIST_MAGIC, 0, len(arrays))]
    chunks += [_pack_array(a) for a in arrays]
    chunks.append(struct.pack("<Q", len(names)))
    for n in names:
        b = n.encode("utf-8")
        chunks.append(struct.pack("<Q", len(b)))
        chunks.append(b)
    with open(fname, "wb") as fh:
        fh.write(b"".join(chunks))


class _Reader:
    def __init__(self, buf):
        self.buf, self.pos = buf, 0

    def take(self, fmt):
        size = struct.calcsize(fmt)
        if self.pos + size > len(self.buf):
            raise ParamsFormatError("truncated NDArray file (need %d bytes at offset %d of %d)" % (size, self.pos, len(self.buf)))
        v = struct.unpack_from(fmt, self.buf, self.pos)
        self.pos += size
        return v

    def raw(self, n):
        if self.pos + n > len(self.buf):
            raise ParamsFormatError("truncated NDArray file (need %d data bytes at offset %d of %d)" % (n, self.pos, len(self.buf)))
        v = self.buf[self.pos:self.pos + n]
        self.pos += n
        return v


def _read_array(r):
    (magic,) = r.take("<I")
    if magic in (NDARRAY_V2_MAGIC, NDARRAY_V3_MAGIC):
        (stype,) = r.take("<i")
        if stype != 0:
            raise ParamsFormatError("sparse NDArray (stype %d) is not supported" % stype)
        (ndim,) = r.take("<I")
        shape = r.take("<%dq" % ndim)
        if ndim == 0 and magic == NDARRAY_V2_MAGIC:
            return None
    elif magic == NDARRAY_V1_MAGIC:
        (ndim,) = r.take("<I")
        shape = r.take("<%dq" % ndim)
        if ndim == 0:
            return None
    else:                                          # legacy: the word just read is ndim, dims are uint32
        ndim = magic
        if ndim > 32:
            raise ParamsFormatError("not an NDArray record (magic 0x%08X)" % magic)
        shape = r.take("<%dI" % ndim)
        if ndim == 0:
            return None
    r.take("<ii")                                  # context it was saved from; arrays are loaded to host memory
    (flag,) = r.take("<i")
    if flag not in _TYPE_FLAGS:
        raise ParamsFormatError("unknown type flag %d" % flag)
    dt = np.dtype(_TYPE_FLAGS[flag]).newbyteorder("<")
    count = 1
    for s in shape:
        if s < 0:
            raise ParamsFormatError("negative dimension in shape %s" % (shape,))
        count *= s
    data = r.raw(count * dt.itemsize)
    return np.frombuffer(data, dtype=dt, count=count).reshape(shape).astype(dt.newbyteorder("="), copy=True)


def nd_load(fname):
    """`mx.nd.load(fname)`: dict {name: numpy array} when names were saved, else a list."""
    with open(fname, "rb") as fh:
        buf = fh.read()
    r = _Reader(buf)
    header, _reserved = r.take("<QQ")
    if header != LIST_MAGIC:
        raise ParamsFormatError("%s is not an MXNet NDArray list file (header 0x%X)" % (fname, header))
    (n,) = r.take("<Q")
    arrays = [_read_array(r) for _ in range(n)]
    (n_names,) = r.take("<Q")
    names = []
    for _ in range(n_names):
        (ln,) = r.take("<Q")
        names.append(r.raw(ln).decode("utf-8"))
    if n_names == 0:
        return arrays
    if n_names != n:
        raise ParamsFormatError("%d names for %d arrays" % (n_names, n))
    return dict(zip(names, arrays))


def is_nd_file(fname):
    with open(fname, "rb") as fh:
        head = fh.read(8)
    return len(head) == 8 and struct.unpack("<Q", head)[0] == LIST_MAGIC


def load_any(fname):
    """{name: array} from an MXNet NDArray file, or from a torch.save()d dict (this repo's round-1 checkpoints)."""
    if is_nd_file(fname):
        out = nd_load(fname)
        if not isinstance(out, dict):
            raise ParamsFormatError("%s holds unnamed arrays; a parameter file needs names" % fname)
        return out
    import torch
    return {k: (v.numpy() if hasattr(v, "numpy") else np.asarray(v)) for k, v in torch.load(fname, map_location="cpu", weights_only=True).items()}


def split_checkpoint(params):
    """`mx.model.load_checkpoint` convention: keys `arg:<name>` / `aux:<name>` -> (arg_params, aux_params).  Keys without
    a prefix are sorted by suffix (moving_mean / moving_var are auxiliary states)."""
    arg, aux = {}, {}
    for k, v in params.items():
        if k.startswith("arg:"):
            arg[k[4:]] = v
        elif k.startswith("aux:"):
            aux[k[4:]] = v
        elif k.endswith("_moving_mean") or k.endswith("_moving_var"):
            aux[k] = v
        else:
            arg[k] = v
    return arg, aux


def save_checkpoint(prefix, epoch, arg_params, aux_params):
    """`prefix-%04d.params` as `mx.callback.do_checkpoint` writes it (reference train.py:83)."""
    d = {"arg:" + k: v for k, v in arg_params.items()}
    d.update({"aux:" + k: v for k, v in aux_params.items()})
    fname = "%s-%04d.params" % (prefix, epoch)
    nd_save(fname, d)
    return fname


def load_checkpoint(prefix, epoch):
    return split_checkpoint(nd_load("%s-%04d.params" % (prefix, epoch)))


# ------------------------------------------------------------------------------------------------ Caffe2 pickle
def caffe2_blobs_to_params(blobs):
    """Reference utils.py:21-33, rule by rule (a blob may match only one suffix; `_riv` is an inverse variance)."""
    args_loaded, auxs_loaded = {}, {}
    for k, v in blobs.items():
        v = np.asarray(v)
        if k.endswith("_w"):
            args_loaded[k[:-2] + "_weight"] = v
        if k.endswith("_b"):
            args_loaded[k[:-2] + "_beta"] = v
        if k.endswith("_s"):
            args_loaded[k[:-2] + "_gamma"] = v
        if k.endswith("_rm"):
            auxs_loaded[k[:-3] + "_moving_mean"] = v
        if k.endswith("_riv"):
            auxs_loaded[k[:-4] + "_moving_var"] = (1.0 / v).astype(v.dtype, copy=False)
    return args_loaded, auxs_loaded


def load_from_caffe2_pkl(filepath, net):
    """Reference utils.py:13-55.  `net` needs `list_arguments()` / `list_auxiliary_states()` (the symbol returned by
    `create_r3d`); returns (args_loaded, auxs_loaded) as numpy arrays and logs the same coverage report."""
    with open(filepath, "rb") as fopen:
        blobs = pickle.load(fopen, encoding="latin1")["blobs"]
    print("len of blobs %d" % len(blobs))
    args_loaded, auxs_loaded = caffe2_blobs_to_params(blobs)
    report = coverage_report(net.list_arguments(), net.list_auxiliary_states(), args_loaded, auxs_loaded)
    for line in report["lines"]:
        logger.info(line)
    return args_loaded, auxs_loaded


def coverage_report(args_symbol, auxs_symbol, args_loaded, auxs_loaded):
    lines = ["symbol has %d = %d arg + %d aux" % (len(args_symbol) + len(auxs_symbol), len(args_symbol), len(auxs_symbol)),
             "model loaded has %d = %d arg + %d aux" % (len(args_loaded) + len(auxs_loaded), len(args_loaded), len(auxs_loaded)),
             "testing arg loaded"]
    not_loaded = [a for a in args_symbol if a not in args_loaded]
    lines += ["arg %s not loaded" % a for a in not_loaded]
    lines.append("testing arg used in net")
    not_used = [a for a in args_loaded if a not in args_symbol]
    lines += ["arg %s not used in net" % a for a in not_used]
    lines.append("testing aux")
    aux_missing = [a for a in auxs_symbol if a not in auxs_loaded]
    lines += ["aux %s not loaded" % a for a in aux_missing]
    return {"lines": lines, "not_loaded": not_loaded, "not_used": not_used, "aux_not_loaded": aux_missing}
