"""Execution plans for the R(2+1)D network on top of the C ABI.

`InferencePlan` runs R2Plus2D.forward (reference model/R2Plus1.py:232-245) in eval mode with every BatchNorm folded
into the producing convolution's epilogue:  one K1 launch per Conv3D (+BN +ReLU [+residual]), one unfold launch for
the stem input, one pool+dense launch — 71 launches for depth 34 against ~280 operator launches in the reference.

The layer list is generated here from the same formulas the reference uses (mid-filter count, strides, block plan);
parameter names are the symbol-API names (net.py:42-51,80,96-98,123-132,166) that the reference's own
`load_from_sym_params` (model/R2Plus1.py:256-279) maps Gluon parameters onto.
"""
import torch

from . import ops
from .ops import FVT_CONV_RELU, FVT_CONV_RESIDUAL, pad16

BLOCK_CONFIG = {
    10: (1, 1, 1, 1),
    16: (2, 2, 2, 1),
    18: (2, 2, 2, 2),
    26: (2, 3, 4, 3),
    34: (3, 4, 6, 3),
}

STEM_UNFOLD_CH = 32   # 7 taps x 3 channels = 21 -> stored 32


def middle_filters(in_filters, out_filter):
    """model/R2Plus1.py:22-24 (true division, then int())."""
    i = 3 * in_filters * out_filter * 3 * 3
    i /= in_filters * 3 * 3 + 3 * out_filter
    return int(i)


class ConvSpec:
    """One Conv3D + the BatchNorm that follows it."""
    __slots__ = ("name", "bn", "cin", "cout", "kernel", "stride", "pad", "relu", "role")

    def __init__(self, name, bn, cin, cout, kernel, stride, pad, relu, role):
        self.name, self.bn, self.cin, self.cout = name, bn, cin, cout
        self.kernel, self.stride, self.pad, self.relu, self.role = kernel, stride, pad, relu, role


def stem_specs():
    return [
        ConvSpec("conv1_middle", "conv1_middle_spatbn_relu", 3, 45, (1, 7, 7), (1, 2, 2), (0, 3, 3), True, "stem_spatial"),
        ConvSpec("conv1", "conv1_spatbn_relu", 45, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), True, "stem_temporal"),
    ]


def block_specs(comp, cin, cout, downsampling):
    """R3DBlock (model/R2Plus1.py:42-82): returns (main path specs, shortcut spec or None)."""
    s = 2 if downsampling else 1
    mid1, mid2 = middle_filters(cin, cout), middle_filters(cout, cout)
    main = [
        ConvSpec("comp_%d_conv_1_middle" % comp, "comp_%d_spatbn_1_middle" % comp, cin, mid1, (1, 3, 3), (1, s, s), (0, 1, 1), True, "spatial"),
        ConvSpec("comp_%d_conv_1" % comp, "comp_%d_spatbn_1" % comp, mid1, cout, (3, 1, 1), (s, 1, 1), (1, 0, 0), True, "temporal"),
        ConvSpec("comp_%d_conv_2_middle" % comp, "comp_%d_spatbn_2_middle" % comp, cout, mid2, (1, 3, 3), (1, 1, 1), (0, 1, 1), True, "spatial"),
        ConvSpec("comp_%d_conv_2" % comp, "comp_%d_spatbn_2" % comp, mid2, cout, (3, 1, 1), (1, 1, 1), (1, 0, 0), False, "temporal_out"),
    ]
    short = None
    if cin != cout or downsampling:
        short = ConvSpec("shortcut_projection_%d" % comp, "shortcut_projection_%d_spatbn" % comp, cin, cout,
                         (1, 1, 1), (s, s, s), (0, 0, 0), False, "shortcut")
    return main, short


def network_blocks(model_depth):
    """[(comp_index, cin, cout, downsampling)] in forward order (model/R2Plus1.py:118-165)."""
    n2, n3, n4, n5 = BLOCK_CONFIG[model_depth]
    out, comp = [], 0
    for cin, cout, nb, down in ((64, 64, n2, False), (64, 128, n3, True), (128, 256, n4, True), (256, 512, n5, True)):
        for b in range(nb):
            out.append((comp, cin if b == 0 else cout, cout, bool(down and b == 0)))
            comp += 1
    return out


def parameter_shapes(model_depth, num_class):
    """Ordered {name: shape} of every trainable tensor and {name: shape} of every BN running statistic."""
    params, aux = {}, {}

    def add(spec):
        params[spec.name + "_weight"] = (spec.cout, spec.cin) + tuple(spec.kernel)
        params[spec.bn + "_gamma"] = (spec.cout,)
        params[spec.bn + "_beta"] = (spec.cout,)
        aux[spec.bn + "_moving_mean"] = (spec.cout,)
        aux[spec.bn + "_moving_var"] = (spec.cout,)

    for s in stem_specs():
        add(s)
    for comp, cin, cout, down in network_blocks(model_depth):
        main, short = block_specs(comp, cin, cout, down)
        for s in main:
            add(s)
        if short is not None:
            add(short)
    params["final_fc_weight"] = (num_class, 512)
    params["final_fc_bias"] = (num_class,)
    return params, aux


def fold_bn(gamma, beta, mean, var, eps, c_store):
    """Eval-mode BatchNorm as y = x*scale + shift; pad channels get (0, 0) so they stay exactly zero."""
    scale = gamma.float() / torch.sqrt(var.float() + eps)
    shift = beta.float() - mean.float() * scale
    s = torch.zeros(c_store, dtype=torch.float32, device=gamma.device)
    b = torch.zeros(c_store, dtype=torch.float32, device=gamma.device)
    s[: scale.numel()] = scale
    b[: shift.numel()] = shift
    return s, b


def stem_equivalent_weight(w):
    """(45, 3, 1, 7, 7) stem filter -> (45, 21, 1, 7, 1) filter over the W-unfolded input:
    w_eq[o, kw*3+ci, 0, kh, 0] = w[o, ci, 0, kh, kw]."""
    o, ci, kt, kh, kw = w.shape
    return w.permute(0, 4, 1, 2, 3).reshape(o, kw * ci, kt, kh, 1).contiguous()


class _Layer:
    __slots__ = ("spec", "desc", "w_packed", "scale", "shift", "out_shape", "src", "dst", "res")


class InferencePlan:
    """Shape-specialised eval-mode forward.  Buffers are allocated once and reused across calls."""

    def __init__(self, params, aux, model_depth, num_class, pool, eps, n, t, h, w, device):
        self.device = device
        self.n, self.t, self.h, self.w = n, t, h, w
        self.num_class = num_class
        self.pool = pool
        self.layers = []
        self.launches = 0
        bufs = {}

        def new_buf(key, shape):
            # buffers are keyed by role and grown to the largest request so blocks ping-pong in place
            numel = 1
            for s in shape:
                numel *= s
            cur = bufs.get(key)
            if cur is None or cur.numel() < numel:
                bufs[key] = torch.empty(numel, dtype=torch.bfloat16, device=device)
            return key, tuple(shape)

        def add_layer(spec, in_shape, src, dst_key, res, w_override=None, kernel=None, stride=None, pad=None, cin_store=None):
            L = _Layer()
            L.spec = spec
            k = kernel or spec.kernel
            s = stride or spec.stride
            p = pad or spec.pad
            cin_s = cin_store or pad16(spec.cin)
            cout_s = pad16(spec.cout)
            flags = (FVT_CONV_RELU if (spec.relu or res is not None) else 0) | (FVT_CONV_RESIDUAL if res is not None else 0)
            nn_, tt, hh, ww = in_shape[:4]
            L.desc = ops.conv_desc(nn_, tt, hh, ww, cin_s, cout_s, k, s, p, flags)
            to, ho, wo = ops.conv_out_shape(L.desc)
            L.out_shape = (nn_, to, ho, wo, cout_s)
            wt = w_override if w_override is not None else params[spec.name + "_weight"]
            L.w_packed = ops.pack_conv_weight(L.desc, wt)
            L.scale, L.shift = fold_bn(params[spec.bn + "_gamma"], params[spec.bn + "_beta"],
                                       aux[spec.bn + "_moving_mean"], aux[spec.bn + "_moving_var"], eps, cout_s)
            L.src = src
            L.dst = new_buf(dst_key, L.out_shape)
            L.res = res
            self.layers.append(L)
            return L.dst, L.out_shape

        # ---- stem: unfold + (1,7,1)/s(1,2,1) conv on K1, then the 3x1x1 temporal conv
        s_sp, s_tm = stem_specs()
        wo_unf = (w + 2 * 3 - 7) // 2 + 1
        self.unfold = new_buf("unfold", (n, t, h, wo_unf, STEM_UNFOLD_CH))
        cur, shp = add_layer(s_sp, (n, t, h, wo_unf), self.unfold, "mid", None,
                             w_override=stem_equivalent_weight(params["conv1_middle_weight"].detach()),
                             kernel=(1, 7, 1), stride=(1, 2, 1), pad=(0, 3, 0), cin_store=STEM_UNFOLD_CH)
        cur, shp = add_layer(s_tm, shp, cur, "x0", None)
        # ---- residual blocks; block input alternates between x0/x1, intermediates reuse mid / y / sc
        flip = 0
        for comp, cin, cout, down in network_blocks(model_depth):
            main, short = block_specs(comp, cin, cout, down)
            x_in, x_shape = cur, shp
            a, sa = add_layer(main[0], x_shape, x_in, "mid", None)
            b, sb = add_layer(main[1], sa, a, "y", None)
            c, sc_ = add_layer(main[2], sb, b, "mid", None)
            if short is not None:
                res, _ = add_layer(short, x_shape, x_in, "sc", None)
            else:
                res = x_in
            flip ^= 1
            cur, shp = add_layer(main[3], sc_, c, "x1" if flip else "x0", res)
        self.final = cur
        self.bufs = bufs
        self.fc_w = params["final_fc_weight"].detach().float().contiguous()
        self.fc_b = params["final_fc_bias"].detach().float().contiguous()
        tp, hp, wp = shp[1] - pool[0] + 1, shp[2] - pool[1] + 1, shp[3] - pool[2] + 1
        if (tp, hp, wp) != (1, 1, 1):
            raise ValueError("AvgPool3D%s over a %s map leaves %s: only a global pool (1x1x1 output) is supported; "
                             "pass final_temporal_kernel = T/8 and final_spatial_kernel = H/16 as the reference callers do"
                             % (pool, shp[1:4], (tp, hp, wp)))
        self.launches = 1 + len(self.layers) + 1

    def _view(self, ref):
        key, shape = ref
        numel = 1
        for s in shape:
            numel *= s
        return self.bufs[key][:numel].view(shape)

    def forward(self, x, want_features=False):
        """x: (N, 3, T, H, W) fp32 CUDA -> logits (N, num_class) fp32 [, pooled features (N, 512)]."""
        assert tuple(x.shape) == (self.n, 3, self.t, self.h, self.w), (tuple(x.shape), (self.n, 3, self.t, self.h, self.w))
        ops.stem_unfold(x.contiguous(), out=self._view(self.unfold))
        for L in self.layers:
            ops.conv3d_fwd(L.desc, self._view(L.src), L.w_packed, L.scale, L.shift,
                           self._view(L.res) if L.res is not None else None, out=self._view(L.dst))
        return ops.pool_fc_fwd(self._view(self.final), 512, self.fc_w, self.fc_b, want_pooled=want_features)
