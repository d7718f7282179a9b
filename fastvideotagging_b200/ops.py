"""Tensor-level wrappers over the C ABI: torch is used for device memory and streams only."""
import ctypes

import torch

from . import _lib
from ._lib import ConvDesc, ConvExt, FVT_CONV_RELU, FVT_CONV_RESIDUAL, FVT_CONV_STATS, FVT_CONV_W_OHWI, FVT_CONV_BN_BWD, check


def pad16(c):
    return (c + 15) // 16 * 16


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _h(t=None):
    """The calling thread's library handle for the device `t` lives on (the current device when t is None).  The
    library refuses a handle whose device is not the current one, so a tensor on another device is an error here, not
    a launch on the wrong GPU."""
    return _lib.handle(t.device.index if (t is not None and t.is_cuda) else None)


set_option = _lib.set_option
get_option = _lib.get_option


# ---------------------------------------------------------------------------------------------------------------------
# exact per-channel accumulators (fvt_stats_bytes): 4 int64 limbs per value
# ---------------------------------------------------------------------------------------------------------------------
STAT_LIMBS = 4


def stats_buffer(c_store, device):
    """Zeroed [2][c_store] accumulators for fvt_conv3d_fwd(FVT_CONV_STATS) / fvt_bn_finalize: int64 (2*c_store, 4)."""
    return torch.zeros((2 * c_store, STAT_LIMBS), dtype=torch.int64, device=device)


def stats_encode(values):
    """float tensor (n,) -> accumulators (n, 4) holding exactly those values (tests / interop)."""
    lib = _lib.load()
    v = values.detach().to(torch.float32).contiguous()
    out = torch.empty((v.numel(), STAT_LIMBS), dtype=torch.int64, device=v.device)
    check(lib.fvt_stats_encode(_h(v), _ptr(v), _ptr(out), v.numel(), _stream()))
    return out


def stats_decode(acc):
    """accumulators (n, 4) -> float32 tensor (n,)."""
    lib = _lib.load()
    assert acc.dtype == torch.int64 and acc.is_contiguous()
    n = acc.numel() // STAT_LIMBS
    out = torch.empty(n, dtype=torch.float32, device=acc.device)
    check(lib.fvt_stats_decode(_h(acc), _ptr(acc), _ptr(out), n, _stream()))
    return out


def require_cuda(t, name):
    if not t.is_cuda:
        raise _lib.FvtError("%s must live on a CUDA device: the R(2+1)D hot path has no CPU fallback" % name)


def conv_desc(n, t, h, w, cin, cout, kernel, stride=(1, 1, 1), pad=(0, 0, 0), flags=0, block_n=0):
    return ConvDesc(n, t, h, w, cin, cout, kernel[0], kernel[1], kernel[2], stride[0], stride[1], stride[2],
                    pad[0], pad[1], pad[2], flags, block_n)


def conv_out_shape(desc):
    lib = _lib.load()
    a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    check(lib.fvt_conv3d_out_shape(ctypes.byref(desc), ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
    return a.value, b.value, c.value


def _with_ohwi(desc):
    d = ConvDesc(*desc.key())
    d.flags |= FVT_CONV_W_OHWI
    return d


def pack_conv_weight(desc, w_oidhw, out=None, ohwi=False):
    """fp32 (O, I, kT, kH, kW) device tensor -> packed bf16 weights for `desc` (zero padded).  ohwi=True: the tensor
    is the (O, kT, kH, kW, I) storage used by engine.FlatParams (FVT_CONV_W_OHWI)."""
    lib = _lib.load()
    require_cuda(w_oidhw, "weight")
    w = w_oidhw.detach().to(torch.float32).contiguous()
    elems = lib.fvt_conv3d_packed_weight_elems(ctypes.byref(desc))
    if elems == 0:
        check(-1)
    if out is None:
        out = torch.empty(elems, dtype=torch.bfloat16, device=w.device)
    assert out.numel() == elems and out.dtype == torch.bfloat16
    d = _with_ohwi(desc) if ohwi else desc
    cout, cin = (w.shape[0], w.shape[4]) if ohwi else (w.shape[0], w.shape[1])
    check(lib.fvt_pack_conv_weight(_h(w), ctypes.byref(d), _ptr(w), cout, cin, _ptr(out), _stream()))
    return out


_WORKSPACE = {}
WORKSPACE_BYTES = 128 << 20


def workspace(device, stream=None):
    """Caller-owned fp32 scratch handed to fvt_conv3d_fwd (split-K slices of small-M convolutions), one per (device,
    stream): launches that share a workspace must be stream-ordered.  Contents are irrelevant between calls."""
    if stream is None:
        stream = torch.cuda.current_stream(device).cuda_stream
    key = (device.type, device.index, int(stream))
    ws = _WORKSPACE.get(key)
    if ws is None:
        ws = torch.empty(WORKSPACE_BYTES // 4, dtype=torch.float32, device=device)
        _WORKSPACE[key] = ws
    return ws


def conv3d_fwd(desc, x, w_packed, scale=None, shift=None, residual=None, out=None, stats=None):
    """x: (N, T, H, W, Cin) bf16 contiguous -> (N, To, Ho, Wo, Cout) bf16."""
    lib = _lib.load()
    require_cuda(x, "x")
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    assert tuple(x.shape) == (desc.n, desc.t, desc.h, desc.w, desc.cin), (tuple(x.shape), desc.key())
    to, ho, wo = conv_out_shape(desc)
    if out is None:
        out = torch.empty((desc.n, to, ho, wo, desc.cout), dtype=torch.bfloat16, device=x.device)
    ws = workspace(x.device)
    if stats is not None:
        assert stats.dtype == torch.int64 and stats.numel() == 2 * desc.cout * STAT_LIMBS, "stats: ops.stats_buffer(cout)"
    check(lib.fvt_conv3d_fwd(_h(x), ctypes.byref(desc), _ptr(x), _ptr(w_packed), _ptr(scale), _ptr(shift), _ptr(residual),
                             _ptr(out), _ptr(stats), _ptr(ws), ws.numel() * 4, _stream()))
    return out


def conv3d_fwd_ex(desc, ext, x, w_packed, out, scale=None, shift=None, residual=None):
    """fvt_conv3d_fwd_ex: `desc` with per-axis high padding and the result written onto a lattice of `out` (see the header;
    the building block of strided data gradients).  out: (N, T, H, W, cout) bf16, partially written."""
    lib = _lib.load()
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and out.is_contiguous()
    assert tuple(x.shape) == (desc.n, desc.t, desc.h, desc.w, desc.cin), (tuple(x.shape), desc.key())
    assert tuple(out.shape) == (desc.n, ext.out_extent[0], ext.out_extent[1], ext.out_extent[2], desc.cout), tuple(out.shape)
    check(lib.fvt_conv3d_fwd_ex(_h(x), ctypes.byref(desc), ctypes.byref(ext), _ptr(x), _ptr(w_packed), _ptr(scale), _ptr(shift),
                                _ptr(residual), _ptr(out), _stream()))
    return out


def dgrad_parity_classes(x_ext, k, s, p):
    """Pure index arithmetic behind DgradPlan.  For a convolution with input extent x_ext, filter k, stride s, padding p
    (3-tuples over t, h, w): the data gradient splits into one STRIDE-1 convolution of dY per parity class `par` of dX,
        dX[s*j + par] = sum_e dY[j + e] * w[par + p - s*e],   e = e_lo .. e_hi  (the e with 0 <= par + p - s*e < k).
    Returns (dY extent, [(par, sub_k, tap_a, pad_lo, pad_hi)]) with sub-filter tap u = e - e_lo copying source tap
    tap_a - s*u, low padding pad_lo = -e_lo, high padding such that the output has the class's extent.  Classes no filter
    tap reaches are left out (dX is zero there)."""
    o_ext = tuple((x_ext[a] + 2 * p[a] - k[a]) // s[a] + 1 for a in range(3))
    out = []
    for par_t in range(s[0]):
        for par_h in range(s[1]):
            for par_w in range(s[2]):
                par = (par_t, par_h, par_w)
                sub_k, tap_a, pad_lo, pad_hi, empty = [], [], [], [], False
                for a in range(3):
                    count = (x_ext[a] - par[a] + s[a] - 1) // s[a]             # dX positions of this class on the axis
                    es = sorted((par[a] + p[a] - kk) // s[a] for kk in range(k[a]) if (par[a] + p[a] - kk) % s[a] == 0)
                    if count <= 0 or not es:
                        empty = True
                        break
                    if es != list(range(es[0], es[-1] + 1)):
                        raise NotImplementedError("non-contiguous parity taps")
                    if es[0] > 0:
                        raise NotImplementedError("padding beyond the filter's reach is not used by the reference's convolutions")
                    sub_k.append(es[-1] - es[0] + 1)
                    pad_lo.append(-es[0])
                    tap_a.append(par[a] + p[a] - s[a] * es[0])
                    pad_hi.append(count - o_ext[a] + sub_k[-1] - 1 - pad_lo[-1])
                    if pad_hi[-1] < 0:
                        raise NotImplementedError("parity class smaller than dY")
                if not empty:
                    out.append((par, tuple(sub_k), tuple(tap_a), tuple(pad_lo), tuple(pad_hi)))
    return o_ext, out


class DgradPlan:
    """Data gradient of a STRIDED convolution as one stride-1 sub-convolution of dY per parity class of dX (header:
    fvt_conv3d_fwd_ex; index arithmetic: dgrad_parity_classes).  fwd: the forward descriptor (stored channel counts);
    w_ohwi: its fp32 master (O, kT, kH, kW, I); the sub-filters are registered in `pack_table` (re-packed with the other
    operand copies).  run(dy, out, residual=None): out (N, T, H, W, cin) bf16 <- dgrad (+ residual)."""

    def __init__(self, fwd, w_ohwi, pack_table):
        self.fwd = fwd
        x_ext, s = (fwd.t, fwd.h, fwd.w), (fwd.st, fwd.sh, fwd.sw)
        o_ext, cls = dgrad_parity_classes(x_ext, (fwd.kt, fwd.kh, fwd.kw), s, (fwd.pt, fwd.ph, fwd.pw))
        assert o_ext == conv_out_shape(fwd)
        self.needs_clear = len(cls) < s[0] * s[1] * s[2]
        self.classes = []
        for par, sub_k, tap_a, pad_lo, pad_hi in cls:
            sub = ConvDesc(fwd.n, o_ext[0], o_ext[1], o_ext[2], fwd.cout, fwd.cin, sub_k[0], sub_k[1], sub_k[2], 1, 1, 1,
                           pad_lo[0], pad_lo[1], pad_lo[2], 0, 0)
            ext = ConvExt((ctypes.c_int32 * 3)(*pad_hi), (ctypes.c_int32 * 3)(*x_ext), (ctypes.c_int32 * 3)(*s),
                          (ctypes.c_int32 * 3)(*par))
            wp = pack_table.add_dgrad_sub(sub, w_ohwi, tap_a, s)
            self.classes.append((sub, ext, wp))

    def run(self, dy, out, residual=None):
        if self.needs_clear:
            if residual is not None:
                out.copy_(residual)              # classes no filter tap reaches carry the residual alone
            else:
                out.zero_()
        for sub, ext, wp in self.classes:
            d = sub
            if residual is not None:
                d = ConvDesc(*sub.key())
                d.flags = FVT_CONV_RESIDUAL
            conv3d_fwd_ex(d, ext, dy, wp, out, residual=residual)
        return out


def unit2p1_supported(d_spatial, d_temporal):
    """True when the (1x3x3, 3x1x1) descriptor pair can run as ONE fused launch (fvt_unit2p1_fwd) on this device."""
    lib = _lib.load()
    return check(lib.fvt_unit2p1_supported(_h(), ctypes.byref(d_spatial), ctypes.byref(d_temporal))) == 1


def unit2p1_fwd(d_spatial, d_temporal, x, w_spatial, scale_mid, shift_mid, w_temporal, scale_out, shift_out,
                residual=None, out=None):
    """The factorised unit + the BatchNorm/ReLU(/residual) around it in one launch (eval mode):
    x (N, T, H, W, 64) bf16 -> relu(bn(conv3x1x1(relu(bn(conv1x3x3(x))))) [+ residual]) (N, T, H, W, 64) bf16."""
    lib = _lib.load()
    require_cuda(x, "x")
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    assert tuple(x.shape) == (d_spatial.n, d_spatial.t, d_spatial.h, d_spatial.w, d_spatial.cin), (tuple(x.shape), d_spatial.key())
    if out is None:
        out = torch.empty((d_temporal.n, d_temporal.t, d_temporal.h, d_temporal.w, d_temporal.cout), dtype=torch.bfloat16,
                          device=x.device)
    check(lib.fvt_unit2p1_fwd(_h(x), ctypes.byref(d_spatial), ctypes.byref(d_temporal), _ptr(x), _ptr(w_spatial), _ptr(scale_mid),
                              _ptr(shift_mid), _ptr(w_temporal), _ptr(scale_out), _ptr(shift_out), _ptr(residual),
                              _ptr(out), _stream()))
    return out


def stem_unfold(x_ncdhw, kw_taps=7, sw=2, pw=3, cu=32, out=None):
    """(N, 3, T, H, W) fp32 -> (N, T, H, Wo, cu) bf16 with u[..., kw*3+ci] = x[n, ci, t, h, ow*sw-pw+kw]."""
    lib = _lib.load()
    require_cuda(x_ncdhw, "clips")
    assert x_ncdhw.dtype == torch.float32 and x_ncdhw.is_contiguous() and x_ncdhw.shape[1] == 3
    n, _, t, h, w = x_ncdhw.shape
    wo = (w + 2 * pw - kw_taps) // sw + 1
    if out is None:
        out = torch.empty((n, t, h, wo, cu), dtype=torch.bfloat16, device=x_ncdhw.device)
    check(lib.fvt_stem_unfold(_h(x_ncdhw), _ptr(x_ncdhw), _ptr(out), n, t, h, w, kw_taps, sw, pw, cu, _stream()))
    return out


def stem_unfold_hpair(x_ncdhw, kw_taps=7, sw=2, pw=3, cu=32, out=None):
    """(N, 3, T, H, W) fp32, H even -> (N, T, H/2, Wo, 2*cu) bf16 with u2[..., (h&1)*cu + kw*3+ci] = x[n, ci, t, h, ow*sw-pw+kw]."""
    lib = _lib.load()
    require_cuda(x_ncdhw, "clips")
    assert x_ncdhw.dtype == torch.float32 and x_ncdhw.is_contiguous() and x_ncdhw.shape[1] == 3
    n, _, t, h, w = x_ncdhw.shape
    wo = (w + 2 * pw - kw_taps) // sw + 1
    if out is None:
        out = torch.empty((n, t, h // 2, wo, 2 * cu), dtype=torch.bfloat16, device=x_ncdhw.device)
    check(lib.fvt_stem_unfold_hpair(_h(x_ncdhw), _ptr(x_ncdhw), _ptr(out), n, t, h, w, kw_taps, sw, pw, cu, _stream()))
    return out


def clip_unfold_u8(clips_u8, out, scale, mean, inv_std, hpair, flip=None, crop_yx=None, crop_hw=None, kw_taps=7, sw=2, pw=3, cu=32):
    """Decoded uint8 frames (N, T, Hs, Ws, 3) -> crop / flip / normalise -> the W-unfolded bf16 stem input, one pass
    (fvt_clip_unfold_u8).  out: the tensor stem_unfold / stem_unfold_hpair would write.  scale, mean[3], inv_std[3]:
    value = (v*scale - mean[c]) * inv_std[c].  crop_yx: (N, 2) int32 device tensor of (y0, x0); crop_hw: (h, w)."""
    lib = _lib.load()
    require_cuda(clips_u8, "clips")
    assert clips_u8.dtype == torch.uint8 and clips_u8.dim() == 5 and clips_u8.shape[-1] == 3 and clips_u8.is_contiguous()
    n, t, hs, ws, _ = clips_u8.shape
    h, w = crop_hw if crop_hw is not None else (hs, ws)
    f3 = ctypes.c_float * 3
    flip_t = flip.to(clips_u8.device, torch.uint8).contiguous() if flip is not None else None
    crop_t = crop_yx.to(clips_u8.device, torch.int32).contiguous() if crop_yx is not None else None
    check(lib.fvt_clip_unfold_u8(_h(clips_u8), _ptr(clips_u8), _ptr(flip_t), _ptr(crop_t), _ptr(out), n, t, hs, ws, h, w,
                                 ctypes.c_float(scale), f3(*[float(v) for v in mean]), f3(*[float(v) for v in inv_std]),
                                 kw_taps, sw, pw, cu, int(bool(hpair)), _stream()))
    return out


def pool_fc_fwd(x, c_real, weight, bias, want_pooled=False):
    """x: (N, T, H, W, C) bf16 -> logits (N, num_class) fp32 [and pooled (N, c_real) fp32]."""
    lib = _lib.load()
    require_cuda(x, "x")
    n = x.shape[0]
    c = x.shape[-1]
    positions = x.numel() // (n * c)
    num_class = weight.shape[0] if weight is not None else 0
    logits = torch.empty((n, num_class), dtype=torch.float32, device=x.device) if weight is not None else None
    pooled = torch.empty((n, c_real), dtype=torch.float32, device=x.device) if want_pooled else None
    check(lib.fvt_pool_fc_fwd(_h(x), _ptr(x), n, positions, c, c_real, _ptr(weight), _ptr(bias), num_class, _ptr(pooled),
                              _ptr(logits), _stream()))
    return (logits, pooled) if want_pooled else logits


# ---------------------------------------------------------------------------------------------------------------------
# training ops
# ---------------------------------------------------------------------------------------------------------------------
_WGRAD_WS = {}
WGRAD_WS_BYTES = 192 << 20


def wgrad_workspace(device, stream=None):
    """Caller-owned scratch for fvt_conv3d_wgrad (dW-shaped slices of the pixel splits, reduced in split order), one per
    (device, stream)."""
    if stream is None:
        stream = torch.cuda.current_stream(device).cuda_stream
    key = (device.type, device.index, int(stream))
    ws = _WGRAD_WS.get(key)
    if ws is None:
        ws = torch.empty(WGRAD_WS_BYTES // 4, dtype=torch.float32, device=device)
        _WGRAD_WS[key] = ws
    return ws


def conv_workspace_bytes(desc, op="fwd", cout_real=0, cin_real=0):
    """fvt_conv3d_workspace_bytes: what a fwd / wgrad call with this descriptor would like."""
    lib = _lib.load()
    return lib.fvt_conv3d_workspace_bytes(_h(), ctypes.byref(desc), 1 if op == "wgrad" else 0, cout_real, cin_real)


def dgrad_desc(fwd, block_n=0, flags=0):
    """Descriptor of the stride-1 convolution that computes the data gradient of `fwd` from dY (for strided `fwd`,
    dY must first be zero-inserted onto the input lattice, see zero_insert): channels swapped, padding k-1-p."""
    return ConvDesc(fwd.n, fwd.t, fwd.h, fwd.w, fwd.cout, fwd.cin, fwd.kt, fwd.kh, fwd.kw, 1, 1, 1,
                    fwd.kt - 1 - fwd.pt, fwd.kh - 1 - fwd.ph, fwd.kw - 1 - fwd.pw, flags, block_n)


def pack_conv_weight_dgrad(ddesc, w_oidhw, out=None, ohwi=False):
    lib = _lib.load()
    w = w_oidhw.detach().to(torch.float32).contiguous()
    elems = lib.fvt_conv3d_packed_weight_elems(ctypes.byref(ddesc))
    if out is None:
        out = torch.empty(elems, dtype=torch.bfloat16, device=w.device)
    assert out.numel() == elems and out.dtype == torch.bfloat16
    d = _with_ohwi(ddesc) if ohwi else ddesc
    cout, cin = (w.shape[0], w.shape[4]) if ohwi else (w.shape[0], w.shape[1])
    check(lib.fvt_pack_conv_weight_dgrad(_h(w), ctypes.byref(d), _ptr(w), cout, cin, _ptr(out), _stream()))
    return out


class BnFoldTable:
    """Device table for fvt_bn_fold_multi: eval-mode BatchNorm of every layer folded into (scale, shift) in one launch."""

    _DTYPE = [("gamma", "<u8"), ("beta", "<u8"), ("mean", "<u8"), ("var", "<u8"), ("scale", "<u8"), ("shift", "<u8"),
              ("c_real", "<i4"), ("c_store", "<i4"), ("eps", "<f4"), ("reserved", "<i4")]

    def __init__(self, device):
        self.device, self.rows, self._keep, self._dev = device, [], [], None

    def add(self, gamma, beta, mean, var, eps, c_store):
        """Registers one BatchNorm; returns the (scale, shift) tensors of c_store floats the launch fills."""
        ts = [t.detach() for t in (gamma, beta, mean, var)]
        for t in ts:
            assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda, "BatchNorm tensors must be contiguous fp32 CUDA tensors"
        scale = torch.empty(c_store, dtype=torch.float32, device=self.device)
        shift = torch.empty(c_store, dtype=torch.float32, device=self.device)
        self.rows.append(tuple(t.data_ptr() for t in ts) + (scale.data_ptr(), shift.data_ptr(), ts[0].numel(), c_store, float(eps), 0))
        self._keep.append((ts, scale, shift))
        self._dev = None
        return scale, shift

    def run(self):
        import numpy as np
        if not self.rows:
            return
        if self._dev is None:
            arr = np.array(self.rows, dtype=np.dtype(self._DTYPE))
            assert arr.dtype.itemsize == 64
            self._dev = torch.from_numpy(arr.view(np.uint8).copy()).to(self.device)
        check(_lib.load().fvt_bn_fold_multi(_h(self._dev), _ptr(self._dev), len(self.rows), _stream()))


class PackTable:
    """Device table for fvt_pack_conv_weights_multi: every (tensor, layout) operand copy of a training step in one launch.
    add_fwd / add_dgrad register fp32 masters stored (O, kT, kH, kW, I) and allocate the packed bf16 buffers."""

    _DTYPE = [("w", "<u8"), ("out", "<u8"), ("kind", "<i4"), ("taps", "<i4"), ("k_store", "<i4"), ("rows", "<i4"),
              ("cout_real", "<i4"), ("cin_real", "<i4"), ("block0", "<u4"), ("nblocks", "<u4"),
              ("sub", "<i4", (3,)), ("src_k", "<i4", (3,)), ("tap_a", "<i4", (3,)), ("tap_s", "<i4", (3,))]

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.total_blocks = 0
        self._dev = None
        self._keep = []

    def _add(self, kind, desc, w_ohwi, cout_real, cin_real, sub=(0, 0, 0), src_k=(0, 0, 0), tap_a=(0, 0, 0), tap_s=(0, 0, 0)):
        lib = _lib.load()
        assert w_ohwi.dtype == torch.float32 and w_ohwi.is_contiguous() and w_ohwi.is_cuda
        elems = lib.fvt_conv3d_packed_weight_elems(ctypes.byref(desc))
        if elems == 0:
            check(-1)
        taps = desc.kt * desc.kh * desc.kw
        rows = elems // (taps * desc.cin)
        out = torch.empty(elems, dtype=torch.bfloat16, device=self.device)
        nb = lib.fvt_pack_entry_blocks(kind, taps, desc.cin, rows)
        self.rows.append((w_ohwi.data_ptr(), out.data_ptr(), kind, taps, desc.cin, rows, cout_real, cin_real, self.total_blocks, nb,
                          tuple(sub), tuple(src_k), tuple(tap_a), tuple(tap_s)))
        self.total_blocks += nb
        self._keep.append((w_ohwi, out))
        self._dev = None
        return out

    def add_fwd(self, desc, w_ohwi):
        """Forward-layout copy for `desc` of a master (O, kT, kH, kW, I)."""
        return self._add(0, desc, w_ohwi, w_ohwi.shape[0], w_ohwi.shape[4])

    def add_dgrad(self, ddesc, w_ohwi):
        """Data-gradient-layout copy for the dgrad descriptor `ddesc` of the FORWARD master (O, kT, kH, kW, I)."""
        return self._add(1, ddesc, w_ohwi, w_ohwi.shape[0], w_ohwi.shape[4])

    def add_dgrad_sub(self, sub_desc, w_ohwi, tap_a, tap_s):
        """One parity sub-filter of a strided convolution's data gradient (fvt_pack_entry kind 2): `sub_desc` is the
        stride-1 convolution over dY (its kt/kh/kw = sub-filter extent), source tap per axis = tap_a - tap_s*u."""
        return self._add(2, sub_desc, w_ohwi, w_ohwi.shape[0], w_ohwi.shape[4], sub=(sub_desc.kt, sub_desc.kh, sub_desc.kw),
                         src_k=tuple(w_ohwi.shape[1:4]), tap_a=tap_a, tap_s=tap_s)

    def run(self):
        import numpy as np
        if not self.rows:
            return
        if self._dev is None:
            arr = np.array(self.rows, dtype=np.dtype(self._DTYPE))
            assert arr.dtype.itemsize == 96
            self._dev = torch.from_numpy(arr.view(np.uint8).copy()).to(self.device)
        lib = _lib.load()
        check(lib.fvt_pack_conv_weights_multi(_h(self._dev), _ptr(self._dev), len(self.rows), self.total_blocks, _stream()))


def zero_insert(dy, fwd, out=None):
    """dy: (N, To, Ho, Wo, C) -> (N, T, H, W, C) with dy on the stride lattice of `fwd`'s input."""
    lib = _lib.load()
    n, to, ho, wo, c = dy.shape
    if out is None:
        out = torch.empty((n, fwd.t, fwd.h, fwd.w, c), dtype=torch.bfloat16, device=dy.device)
    check(lib.fvt_zero_insert(_h(dy), _ptr(dy), _ptr(out), n, fwd.t, fwd.h, fwd.w, to, ho, wo, fwd.st, fwd.sh, fwd.sw, c, _stream()))
    return out


def conv3d_wgrad(fwd, x, dy, dw, cout_real, cin_real, ohwi=False, ws="auto"):
    """dw (fp32, (cout_real, cin_real, kT, kH, kW); ohwi=True: (cout_real, kT, kH, kW, cin_real)) = wgrad(x, dy)
    (overwritten: MXNet grad_req='write'; deterministic — pixel splits meet through workspace slices, not atomics)."""
    lib = _lib.load()
    assert dw.dtype == torch.float32 and dw.is_contiguous()
    d = _with_ohwi(fwd) if ohwi else fwd
    if isinstance(ws, str):
        ws = wgrad_workspace(x.device)                       # ws=None: no workspace (one pixel split per dW tile)
    check(lib.fvt_conv3d_wgrad(_h(x), ctypes.byref(d), _ptr(x), _ptr(dy), _ptr(dw), cout_real, cin_real, _ptr(ws),
                               ws.numel() * 4 if ws is not None else 0, _stream()))
    return dw


def wgrad_group_eligible(descs, cout_real, cin_real, device):
    """Which of these layers a grouped weight-gradient launch would take (fvt_conv3d_wgrad_group_plan without buffers)."""
    lib = _lib.load()
    n = len(descs)
    arr = (ConvDesc * n)(*[_with_ohwi(d) for d in descs])
    co = (ctypes.c_int32 * n)(*cout_real)
    ci = (ctypes.c_int32 * n)(*cin_real)
    member = (ctypes.c_int32 * n)()
    tb, wb = ctypes.c_size_t(0), ctypes.c_size_t(0)
    check(lib.fvt_conv3d_wgrad_group_plan(_lib.handle(device.index), n, arr, None, None, None, co, ci, None, 0, None, 0,
                                          ctypes.byref(tb), ctypes.byref(wb), member))
    return [bool(m) for m in member]


class WgradGroup:
    """Weight gradients of several layers as ONE launch (fvt_conv3d_wgrad_group_plan / _run).  layers: list of
    (fwd_desc, x, dy, dw, cout_real, cin_real) with dw in the (O, kT, kH, kW, I) layout; all tensors are kept by reference
    (their pointers are baked into the launch table).  Layers the grouped kernel does not take (strided, 1x1x1) are listed
    in `.rest` — run() launches them one by one behind the group."""

    def __init__(self, layers, device):
        lib = _lib.load()
        n = len(layers)
        self.layers = layers
        self.device = device
        descs = (ConvDesc * n)(*[_with_ohwi(L[0]) for L in layers])
        self._descs = descs
        vp = ctypes.c_void_p
        xs = (vp * n)(*[L[1].data_ptr() for L in layers])
        dys = (vp * n)(*[L[2].data_ptr() for L in layers])
        dws = (vp * n)(*[L[3].data_ptr() for L in layers])
        for L in layers:
            assert L[3].dtype == torch.float32 and L[3].is_contiguous()
        co = (ctypes.c_int32 * n)(*[L[4] for L in layers])
        ci = (ctypes.c_int32 * n)(*[L[5] for L in layers])
        member = (ctypes.c_int32 * n)()
        tb, wb = ctypes.c_size_t(0), ctypes.c_size_t(0)
        h = _lib.handle(device.index)
        check(lib.fvt_conv3d_wgrad_group_plan(h, n, descs, xs, dys, dws, co, ci, None, 0, None, 0, ctypes.byref(tb), ctypes.byref(wb), member))
        self.in_group = [bool(m) for m in member]
        self.rest = [L for L, m in zip(layers, self.in_group) if not m]
        self.table_bytes, self.ws_bytes = tb.value, wb.value
        self.ws = torch.empty(max(wb.value, 16) // 4, dtype=torch.float32, device=device) if wb.value else None
        self.host = self.dev = None
        if tb.value:
            self.host = (ctypes.c_uint8 * tb.value)()
            check(lib.fvt_conv3d_wgrad_group_plan(h, n, descs, xs, dys, dws, co, ci, _ptr(self.ws), wb.value, self.host, tb.value,
                                                  ctypes.byref(tb), ctypes.byref(wb), member))
            raw = torch.frombuffer(self.host, dtype=torch.uint8).clone()
            pad = torch.empty(tb.value + 128, dtype=torch.uint8, device=device)
            off = (-pad.data_ptr()) % 128
            self.dev = pad[off:off + tb.value]
            self.dev.copy_(raw)
            self._pad = pad
            hdr = torch.frombuffer(self.host, dtype=torch.int32, count=6)
            self.grid, self.red_blocks = int(hdr[2]), int(hdr[3])

    def run(self):
        lib = _lib.load()
        if self.dev is not None:
            check(lib.fvt_conv3d_wgrad_group_run(_lib.handle(self.device.index), self.host, _ptr(self.dev), _stream()))
        for fwd, x, dy, dw, co, ci in self.rest:
            conv3d_wgrad(fwd, x, dy, dw, co, ci, ohwi=True)


def bn_finalize(stats, gamma, beta, running_mean, running_var, c_store, rows, eps, momentum, scale, shift, mean, invstd):
    lib = _lib.load()
    check(lib.fvt_bn_finalize(_h(stats), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), c_store,
                              gamma.numel(), rows, eps, momentum, _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd), _stream()))


def bn_apply(raw, scale, shift, out, relu, res=None, res_scale=None, res_shift=None):
    lib = _lib.load()
    c = raw.shape[-1]
    rows = raw.numel() // c
    check(lib.fvt_bn_apply(_h(raw), _ptr(raw), _ptr(scale), _ptr(shift), _ptr(res), _ptr(res_scale), _ptr(res_shift), _ptr(out),
                           rows, c, int(relu), _stream()))
    return out


def bn_finalize_apply(stats, gamma, beta, running_mean, running_var, c_store, rows, eps, momentum, scale, shift, mean, invstd,
                      raw, out, relu, res=None, res_scale=None, res_shift=None):
    """bn_finalize + bn_apply in one launch (training forward)."""
    lib = _lib.load()
    check(lib.fvt_bn_finalize_apply(_h(raw), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), c_store,
                                    gamma.numel(), rows, eps, momentum, _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd),
                                    _ptr(raw), _ptr(res), _ptr(res_scale), _ptr(res_shift), _ptr(out), int(relu), _stream()))
    return out


_BN_ACC = {}


def _bn_acc(c, device):
    """Scratch accumulators behind fvt_bn_backward's `sums` (zeroed inside the call), one per (device, stream)."""
    key = (device.index, int(torch.cuda.current_stream(device).cuda_stream))
    acc = _BN_ACC.get(key)
    if acc is None or acc.shape[0] < 2 * c:
        acc = torch.zeros((2 * max(c, 2048), STAT_LIMBS), dtype=torch.int64, device=device)
        _BN_ACC[key] = acc
    return acc


def bn_backward(raw, dact, mask, mean, invstd, gamma, sums, draw, dz_out=None, relu_scale=None, relu_shift=None, sums_acc=None,
                dz_in=False):
    """mask: tensor whose sign gates the gradient (ReLU after a residual add), or None; relu_scale/relu_shift: the
    forward scale/shift of this BatchNorm when the ReLU follows it directly (mask recomputed from raw).
    dz_in=2: `dact` already is dz and `sums_acc` already holds [sum dz*raw, sum dz] — written by the data-gradient
    convolution's epilogue (FVT_CONV_BN_BWD) — so only the apply pass runs (dz_in=1: the same with sum dz*(raw-mean))."""
    lib = _lib.load()
    c = raw.shape[-1]
    rows = raw.numel() // c
    if sums_acc is None:
        assert not dz_in
        sums_acc = _bn_acc(c, raw.device)
    check(lib.fvt_bn_backward(_h(raw), _ptr(raw), _ptr(dact), _ptr(mask), _ptr(mean), _ptr(invstd), _ptr(gamma),
                              _ptr(relu_scale), _ptr(relu_shift), _ptr(sums), _ptr(sums_acc), _ptr(draw), _ptr(dz_out), rows, c,
                              gamma.numel(), int(dz_in), _stream()))        # dz_in: 0 two passes, 1 / 2 apply only (see the header)
    return draw


def pool_fc_bwd(dlogits, pooled, weight, dw, db, dx):
    lib = _lib.load()
    n, k = dlogits.shape
    c = pooled.shape[1]
    c_store = dx.shape[-1]
    positions = dx.numel() // (n * c_store)
    check(lib.fvt_pool_fc_bwd(_h(dlogits), _ptr(dlogits), _ptr(pooled), _ptr(weight), n, k, c, positions, _ptr(dw), _ptr(db), _ptr(dx),
                              c_store, _stream()))


# ------------------------------------------------------------------------------------------------ fp32 path
def conv3d_fwd_f32(desc, x, w_thwio, scale=None, shift=None, residual=None, out=None):
    """fp32 NDHWC conv (+ folded BN / residual / ReLU per desc.flags) on the CUDA cores; weights (kT, kH, kW, I, O)."""
    lib = _lib.load()
    require_cuda(x, "x")
    assert x.dtype == torch.float32 and x.is_contiguous() and w_thwio.dtype == torch.float32 and w_thwio.is_contiguous()
    to, ho, wo = ((desc.t + 2 * desc.pt - desc.kt) // desc.st + 1, (desc.h + 2 * desc.ph - desc.kh) // desc.sh + 1,
                  (desc.w + 2 * desc.pw - desc.kw) // desc.sw + 1)
    if out is None:
        out = torch.empty((desc.n, to, ho, wo, desc.cout), dtype=torch.float32, device=x.device)
    check(lib.fvt_conv3d_fwd_f32(_h(x), ctypes.byref(desc), _ptr(x), _ptr(w_thwio), _ptr(scale), _ptr(shift), _ptr(residual),
                                 _ptr(out), _stream()))
    return out


def pool_fc_fwd_f32(x, weight, bias, want_pooled=False):
    """x: (N, T, H, W, C) fp32 -> logits (N, num_class) fp32 [and pooled (N, C)]."""
    lib = _lib.load()
    require_cuda(x, "x")
    n, c = x.shape[0], x.shape[-1]
    positions = x.numel() // (n * c)
    logits = torch.empty((n, weight.shape[0]), dtype=torch.float32, device=x.device)
    pooled = torch.empty((n, c), dtype=torch.float32, device=x.device) if want_pooled else None
    check(lib.fvt_pool_fc_fwd_f32(_h(x), _ptr(x), n, positions, c, _ptr(weight), _ptr(bias), weight.shape[0], _ptr(pooled),
                                  _ptr(logits), _stream()))
    return (logits, pooled) if want_pooled else logits
