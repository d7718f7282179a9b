"""ctypes binding of libfvt_b200.so — the only door between the Python host code and the sm_100a kernels.

There is no CPU fallback: `load()` raises if the shared library is missing, and every compute entry point
returns FVT_ERR_UNSUPPORTED_ARCH on a non-sm_100 device, which `check()` turns into an exception.
"""
import ctypes
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libfvt_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "fvt_b200.h")

FVT_CONV_RELU = 1
FVT_CONV_RESIDUAL = 2
FVT_CONV_STATS = 4
FVT_CONV_W_OHWI = 8
FVT_CONV_BN_BWD = 16


class FvtError(RuntimeError):
    pass


class ConvExt(ctypes.Structure):
    """Mirror of `fvt_conv_ext` (include/fvt_b200.h): per-axis high padding + output lattice of fvt_conv3d_fwd_ex."""
    _fields_ = [("pad_hi", ctypes.c_int32 * 3), ("out_extent", ctypes.c_int32 * 3), ("out_stride", ctypes.c_int32 * 3),
                ("out_offset", ctypes.c_int32 * 3)]


class ConvDesc(ctypes.Structure):
    """Mirror of `fvt_conv_desc` (include/fvt_b200.h)."""
    _fields_ = [(k, ctypes.c_int32) for k in (
        "n", "t", "h", "w", "cin", "cout", "kt", "kh", "kw", "st", "sh", "sw", "pt", "ph", "pw", "flags", "block_n")]

    def key(self):
        return tuple(getattr(self, k) for k, _ in self._fields_)


_lib = None


def header_symbols():
    """Names of every function declared in include/fvt_b200.h (used by the export test)."""
    with open(HEADER_PATH) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fvt_[a-z0-9_]+)\s*\(", text)))


def load():
    """Load the shared library (building it is `fastvideotagging_b200.build.build()`'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FvtError(
            "libfvt_b200.so is not built (%s). Run `python -m fastvideotagging_b200.build` — there is no "
            "CPU fallback for the R(2+1)D hot path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, fp = ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p
    dp = ctypes.POINTER(ConvDesc)
    ip = ctypes.POINTER(ctypes.c_int32)
    hp = ctypes.c_void_p          # fvt_handle_t
    sigs = {
        "fvt_version": (ctypes.c_int, []),
        "fvt_last_error": (ctypes.c_char_p, []),
        "fvt_device_check": (ctypes.c_int, [ctypes.c_int]),
        "fvt_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
        "fvt_destroy": (ctypes.c_int, [hp]),
        "fvt_set_option": (ctypes.c_int, [hp, ctypes.c_char_p, ctypes.c_int]),
        "fvt_get_option": (ctypes.c_int, [hp, ctypes.c_char_p, ip]),
        "fvt_stats_bytes": (ctypes.c_size_t, [i32]),
        "fvt_stats_encode": (ctypes.c_int, [hp, fp, vp, i32, vp]),
        "fvt_stats_decode": (ctypes.c_int, [hp, vp, fp, i32, vp]),
        "fvt_conv3d_out_shape": (ctypes.c_int, [dp, ip, ip, ip]),
        "fvt_conv3d_block_n": (ctypes.c_int, [dp]),
        "fvt_conv3d_packed_weight_elems": (ctypes.c_size_t, [dp]),
        "fvt_pack_conv_weight": (ctypes.c_int, [hp, dp, fp, i32, i32, vp, vp]),
        "fvt_conv3d_fwd": (ctypes.c_int, [hp, dp, vp, vp, fp, fp, vp, vp, vp, vp, ctypes.c_size_t, vp]),
        "fvt_conv3d_workspace_bytes": (ctypes.c_size_t, [hp, dp, i32, i32, i32]),
        "fvt_conv3d_fwd_ex": (ctypes.c_int, [hp, dp, ctypes.POINTER(ConvExt), vp, vp, fp, fp, vp, vp, vp]),
        "fvt_unit2p1_supported": (ctypes.c_int, [hp, dp, dp]),
        "fvt_unit2p1_fwd": (ctypes.c_int, [hp, dp, dp, vp, vp, fp, fp, vp, fp, fp, vp, vp, vp]),
        "fvt_stem_unfold": (ctypes.c_int, [hp, fp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
        "fvt_stem_unfold_hpair": (ctypes.c_int, [hp, fp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
        "fvt_conv3d_fwd_f32": (ctypes.c_int, [hp, dp, fp, fp, fp, fp, fp, fp, vp]),
        "fvt_pool_fc_fwd_f32": (ctypes.c_int, [hp, fp, i32, i32, i32, fp, fp, i32, fp, fp, vp]),
        "fvt_pool_fc_fwd": (ctypes.c_int, [hp, vp, i32, i32, i32, i32, fp, fp, i32, fp, fp, vp]),
        "fvt_pack_conv_weight_dgrad": (ctypes.c_int, [hp, dp, fp, i32, i32, vp, vp]),
        "fvt_pack_entry_blocks": (ctypes.c_uint32, [i32, i32, i32, i32]),
        "fvt_pack_conv_weights_multi": (ctypes.c_int, [hp, vp, i32, ctypes.c_uint32, vp]),
        "fvt_conv3d_wgrad": (ctypes.c_int, [hp, dp, vp, vp, fp, i32, i32, vp, ctypes.c_size_t, vp]),
        "fvt_conv3d_wgrad_group_plan": (ctypes.c_int, [hp, i32, dp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp), ip, ip,
                                                       vp, ctypes.c_size_t, vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t),
                                                       ctypes.POINTER(ctypes.c_size_t), ip]),
        "fvt_conv3d_wgrad_group_run": (ctypes.c_int, [hp, vp, vp, vp]),
        "fvt_zero_insert": (ctypes.c_int, [hp, vp, vp] + [i32] * 11 + [vp]),
        "fvt_bn_fold_multi": (ctypes.c_int, [hp, vp, i32, vp]),
        "fvt_bn_finalize": (ctypes.c_int, [hp, vp, fp, fp, fp, fp, i32, i32, ctypes.c_int64, ctypes.c_float, ctypes.c_float,
                                           fp, fp, fp, fp, vp]),
        "fvt_bn_apply": (ctypes.c_int, [hp, vp, fp, fp, vp, fp, fp, vp, ctypes.c_int64, i32, i32, vp]),
        "fvt_bn_finalize_apply": (ctypes.c_int, [hp, vp, fp, fp, fp, fp, i32, i32, ctypes.c_int64, ctypes.c_float, ctypes.c_float,
                                                 fp, fp, fp, fp, vp, vp, fp, fp, vp, i32, vp]),
        "fvt_bn_backward": (ctypes.c_int, [hp, vp, vp, vp, fp, fp, fp, fp, fp, fp, vp, vp, vp, ctypes.c_int64, i32, i32, i32, vp]),
        "fvt_pool_fc_bwd": (ctypes.c_int, [hp, fp, fp, fp, i32, i32, i32, i32, fp, fp, vp, i32, vp]),
        "fvt_sgd_momentum_multi": (ctypes.c_int, [hp, vp, vp, vp, i32, ctypes.c_uint32, ctypes.c_float, ctypes.c_float,
                                                  ctypes.c_float, vp]),
        "fvt_clip_stats_u8": (ctypes.c_int, [hp, vp, ctypes.c_int64, vp, vp]),
        "fvt_clip_normalize_u8": (ctypes.c_int, [hp, vp, vp, fp, i32, i32, i32, i32, ctypes.c_float,
                                                 ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), vp]),
        "fvt_clip_unfold_u8": (ctypes.c_int, [hp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, ctypes.c_float,
                                              ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), i32, i32, i32, i32, i32, vp]),
        "fvt_softmax_accumulate": (ctypes.c_int, [hp, fp, fp, i32, i32, vp]),
        "fvt_argmax_correct": (ctypes.c_int, [hp, fp, vp, i32, i32, vp, vp, vp]),
        "fvt_topk_iou": (ctypes.c_int, [hp, fp, fp, i32, i32, i32, vp, vp, vp]),
        "fvt_loss_workspace_bytes": (ctypes.c_size_t, [i32]),
        "fvt_lsep_fwd_bwd": (ctypes.c_int, [hp, fp, fp, i32, i32, i32, fp, fp, vp, vp]),
        "fvt_warp_fwd_bwd": (ctypes.c_int, [hp, fp, fp, i32, i32, i32, i32, i32, ctypes.c_uint64, ctypes.c_uint64, fp, fp,
                                            vp, fp, fp, vp, vp]),
        "fvt_bce_fwd_bwd": (ctypes.c_int, [hp, fp, fp, i32, i32, i32, fp, fp, vp]),
        "fvt_softmax_fwd_bwd": (ctypes.c_int, [hp, fp, fp, i32, i32, i32, fp, fp, vp]),
        "fvt_philox4x32_10": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32),
                                             ctypes.POINTER(ctypes.c_uint32)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_handles = {}
_handles_lock = __import__("threading").Lock()


def handle(device_index=None):
    """The fvt_handle_t of (calling thread, CUDA device): created on first use, kept for the life of the process.
    device_index=None: torch's current device.  One handle per (thread, device) is the library's threading contract
    (include/fvt_b200.h); the handle carries the tuning switches (`set_option`)."""
    import threading
    import torch
    if device_index is None:
        device_index = torch.cuda.current_device()
    key = (threading.get_ident(), int(device_index))
    h = _handles.get(key)
    if h is None:
        lib = load()
        out = ctypes.c_void_p()
        check(lib.fvt_create(ctypes.byref(out), int(device_index)))
        h = ctypes.c_void_p(out.value)
        # FVT_PDL=1: programmatic dependent launch for every hot-path kernel of this handle (csrc/pdl.cuh)
        if os.environ.get("FVT_PDL", "0") == "1":
            check(lib.fvt_set_option(h, b"pdl", 1))
        with _handles_lock:
            _handles[key] = h
    return h


def set_option(name, value, device_index=None):
    """fvt_set_option on the calling thread's handle for the (current) device."""
    return check(load().fvt_set_option(handle(device_index), name.encode() if isinstance(name, str) else name, int(value)))


def get_option(name, device_index=None):
    out = ctypes.c_int32()
    check(load().fvt_get_option(handle(device_index), name.encode() if isinstance(name, str) else name, ctypes.byref(out)))
    return out.value


def check(status):
    if status < 0:
        msg = load().fvt_last_error()
        raise FvtError("libfvt_b200 error %d: %s" % (status, msg.decode() if msg else "?"))
    return status
