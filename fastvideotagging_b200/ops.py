"""Tensor-level wrappers over the C ABI: torch is used for device memory and streams only."""
import ctypes

import torch

from . import _lib
from ._lib import ConvDesc, FVT_CONV_RELU, FVT_CONV_RESIDUAL, FVT_CONV_STATS, FVT_CONV_W_OHWI, check


def pad16(c):
    return (c + 15) // 16 * 16


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


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
    check(lib.fvt_pack_conv_weight(ctypes.byref(d), _ptr(w), cout, cin, _ptr(out), _stream()))
    return out


_WORKSPACE = {}
WORKSPACE_BYTES = 64 << 20


def workspace(device):
    """Per-device zeroed fp32 scratch handed to fvt_conv3d_fwd (split-K of small-M convolutions).  The library keeps it
    zeroed, so it is allocated and cleared exactly once."""
    key = (device.type, device.index)
    ws = _WORKSPACE.get(key)
    if ws is None:
        ws = torch.zeros(WORKSPACE_BYTES // 4, dtype=torch.float32, device=device)
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
    check(lib.fvt_conv3d_fwd(ctypes.byref(desc), _ptr(x), _ptr(w_packed), _ptr(scale), _ptr(shift), _ptr(residual),
                             _ptr(out), _ptr(stats), _ptr(ws), ws.numel() * 4, _stream()))
    return out


def unit2p1_supported(d_spatial, d_temporal):
    """True when the (1x3x3, 3x1x1) descriptor pair can run as ONE fused launch (fvt_unit2p1_fwd) on this device."""
    lib = _lib.load()
    return check(lib.fvt_unit2p1_supported(ctypes.byref(d_spatial), ctypes.byref(d_temporal))) == 1


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
    check(lib.fvt_unit2p1_fwd(ctypes.byref(d_spatial), ctypes.byref(d_temporal), _ptr(x), _ptr(w_spatial), _ptr(scale_mid),
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
    check(lib.fvt_stem_unfold(_ptr(x_ncdhw), _ptr(out), n, t, h, w, kw_taps, sw, pw, cu, _stream()))
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
    check(lib.fvt_stem_unfold_hpair(_ptr(x_ncdhw), _ptr(out), n, t, h, w, kw_taps, sw, pw, cu, _stream()))
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
    check(lib.fvt_pool_fc_fwd(_ptr(x), n, positions, c, c_real, _ptr(weight), _ptr(bias), num_class, _ptr(pooled),
                              _ptr(logits), _stream()))
    return (logits, pooled) if want_pooled else logits


# ---------------------------------------------------------------------------------------------------------------------
# training ops
# ---------------------------------------------------------------------------------------------------------------------
_WGRAD_WS = {}
WGRAD_WS_BYTES = 160 << 20


def wgrad_workspace(device, enable=True):
    """Registers (once per device) the scratch the slab weight-gradient kernels reduce their pixel splits through
    (fvt_set_wgrad_workspace) — plain stores + one reduce pass instead of fp32 atomics.  enable=False withdraws it."""
    lib = _lib.load()
    key = (device.type, device.index)
    with torch.cuda.device(device):
        if not enable:
            check(lib.fvt_set_wgrad_workspace(None, 0))
            return None
        ws = _WGRAD_WS.get(key)
        if ws is None:
            ws = torch.empty(WGRAD_WS_BYTES // 4, dtype=torch.float32, device=device)
            _WGRAD_WS[key] = ws
        check(lib.fvt_set_wgrad_workspace(_ptr(ws), ws.numel() * 4))
    return ws


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
    check(lib.fvt_pack_conv_weight_dgrad(ctypes.byref(d), _ptr(w), cout, cin, _ptr(out), _stream()))
    return out


def zero_insert(dy, fwd, out=None):
    """dy: (N, To, Ho, Wo, C) -> (N, T, H, W, C) with dy on the stride lattice of `fwd`'s input."""
    lib = _lib.load()
    n, to, ho, wo, c = dy.shape
    if out is None:
        out = torch.empty((n, fwd.t, fwd.h, fwd.w, c), dtype=torch.bfloat16, device=dy.device)
    check(lib.fvt_zero_insert(_ptr(dy), _ptr(out), n, fwd.t, fwd.h, fwd.w, to, ho, wo, fwd.st, fwd.sh, fwd.sw, c, _stream()))
    return out


def conv3d_wgrad(fwd, x, dy, dw, cout_real, cin_real, ohwi=False):
    """dw (fp32, (cout_real, cin_real, kT, kH, kW); ohwi=True: (cout_real, kT, kH, kW, cin_real)) += wgrad(x, dy)."""
    lib = _lib.load()
    assert dw.dtype == torch.float32 and dw.is_contiguous()
    d = _with_ohwi(fwd) if ohwi else fwd
    check(lib.fvt_conv3d_wgrad(ctypes.byref(d), _ptr(x), _ptr(dy), _ptr(dw), cout_real, cin_real, _stream()))
    return dw


def bn_finalize(stats, gamma, beta, running_mean, running_var, c_store, rows, eps, momentum, scale, shift, mean, invstd):
    lib = _lib.load()
    check(lib.fvt_bn_finalize(_ptr(stats), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), c_store,
                              gamma.numel(), rows, eps, momentum, _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd), _stream()))


def bn_apply(raw, scale, shift, out, relu, res=None, res_scale=None, res_shift=None):
    lib = _lib.load()
    c = raw.shape[-1]
    rows = raw.numel() // c
    check(lib.fvt_bn_apply(_ptr(raw), _ptr(scale), _ptr(shift), _ptr(res), _ptr(res_scale), _ptr(res_shift), _ptr(out),
                           rows, c, int(relu), _stream()))
    return out


def bn_finalize_apply(stats, gamma, beta, running_mean, running_var, c_store, rows, eps, momentum, scale, shift, mean, invstd,
                      raw, out, relu, res=None, res_scale=None, res_shift=None):
    """bn_finalize + bn_apply in one launch (training forward)."""
    lib = _lib.load()
    check(lib.fvt_bn_finalize_apply(_ptr(stats), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), c_store,
                                    gamma.numel(), rows, eps, momentum, _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd),
                                    _ptr(raw), _ptr(res), _ptr(res_scale), _ptr(res_shift), _ptr(out), int(relu), _stream()))
    return out


def bn_backward(raw, dact, mask, mean, invstd, gamma, sums, draw, dz_out=None, relu_scale=None, relu_shift=None):
    """mask: tensor whose sign gates the gradient (ReLU after a residual add), or None; relu_scale/relu_shift: the
    forward scale/shift of this BatchNorm when the ReLU follows it directly (mask recomputed from raw)."""
    lib = _lib.load()
    c = raw.shape[-1]
    rows = raw.numel() // c
    check(lib.fvt_bn_backward(_ptr(raw), _ptr(dact), _ptr(mask), _ptr(mean), _ptr(invstd), _ptr(gamma),
                              _ptr(relu_scale), _ptr(relu_shift), _ptr(sums), _ptr(draw), _ptr(dz_out), rows, c,
                              gamma.numel(), _stream()))
    return draw


def pool_fc_bwd(dlogits, pooled, weight, dw, db, dx):
    lib = _lib.load()
    n, k = dlogits.shape
    c = pooled.shape[1]
    c_store = dx.shape[-1]
    positions = dx.numel() // (n * c_store)
    check(lib.fvt_pool_fc_bwd(_ptr(dlogits), _ptr(pooled), _ptr(weight), n, k, c, positions, _ptr(dw), _ptr(db), _ptr(dx),
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
    check(lib.fvt_conv3d_fwd_f32(ctypes.byref(desc), _ptr(x), _ptr(w_thwio), _ptr(scale), _ptr(shift), _ptr(residual),
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
    check(lib.fvt_pool_fc_fwd_f32(_ptr(x), n, positions, c, _ptr(weight), _ptr(bias), weight.shape[0], _ptr(pooled),
                                  _ptr(logits), _stream()))
    return (logits, pooled) if want_pooled else logits
