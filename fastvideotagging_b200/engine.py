"""Execution plans for the R(2+1)D network on top of the C ABI.

`InferencePlan` runs R2Plus2D.forward (reference model/R2Plus1.py:232-245) in eval mode with every BatchNorm folded
into the producing convolution's epilogue:  one K1 launch per Conv3D (+BN +ReLU [+residual]), one unfold launch for
the stem input, one pool+dense launch — 71 launches for depth 34 against ~280 operator launches in the reference.

The layer list is generated here from the same formulas the reference uses (mid-filter count, strides, block plan);
parameter names are the symbol-API names (net.py:42-51,80,96-98,123-132,166) that the reference's own
`load_from_sym_params` (model/R2Plus1.py:256-279) maps Gluon parameters onto.
"""
import os

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


def stem_equivalent_weight(w):
    """(45, 3, 1, 7, 7) stem filter -> (45, 21, 1, 7, 1) filter over the W-unfolded input:
    w_eq[o, kw*3+ci, 0, kh, 0] = w[o, ci, 0, kh, kw]."""
    o, ci, kt, kh, kw = w.shape
    return w.permute(0, 4, 1, 2, 3).reshape(o, kw * ci, kt, kh, 1).contiguous()


class StemGeometry:
    """How the 1x7x7 / s(1,2,2) stem conv (reference model/R2Plus1.py:100-104) is laid onto the conv kernels.

    hpair (even H): the clip is W-unfolded with rows 2*h2, 2*h2+1 side by side in 64 channels (ops.stem_unfold_hpair) and
    the conv becomes a stride-1 (1,5,1) conv, pad (0,2,0), over h2 — eligible for the slab kernels, which read every
    input row once (the (1,7,1)/s(1,2,1) form re-reads the unfolded input 7 times from L2: 0.65 -> 0.2 ms at batch 48).
      w2[o, par*32 + kw*3+ci, 0, kh2, 0] = w[o, ci, 0, kh = 2*kh2 + par - 1, kw]    (zero where kh is outside 0..6)
    otherwise: the (1,7,1)/s(1,2,1) conv over the 32-channel unfold."""

    def __init__(self, h):
        self.hpair = (h % 2 == 0) and os.environ.get("FVT_STEM_HPAIR", "1") != "0"
        if self.hpair:
            self.cin_store, self.kernel, self.stride, self.pad, self.cin_real = 64, (1, 5, 1), (1, 1, 1), (0, 2, 0), 64
        else:
            self.cin_store, self.kernel, self.stride, self.pad, self.cin_real = STEM_UNFOLD_CH, (1, 7, 1), (1, 2, 1), (0, 3, 0), 21

    def unfold_shape(self, n, t, h, w):
        wo = (w + 2 * 3 - 7) // 2 + 1
        return (n, t, h // 2, wo, 64) if self.hpair else (n, t, h, wo, STEM_UNFOLD_CH)

    def unfold(self, x, out, norm=None):
        """x: the reference's (N, 3, T, H, W) fp32 clip batch, or decoded uint8 frames (N, T, H, W, 3) with
        norm = (scale, mean[3], inv_std[3]) — normalised on the fly (ops.clip_unfold_u8: SURVEY 8f N2)."""
        if x.dtype == torch.uint8:
            if norm is None:
                raise ValueError("uint8 clips need the normalisation constants: R2Plus2D.set_input_normalization(...)")
            return ops.clip_unfold_u8(x, out, norm[0], norm[1], norm[2], self.hpair)
        return ops.stem_unfold_hpair(x, out=out) if self.hpair else ops.stem_unfold(x, out=out)

    def weight(self, w):
        """(45, 3, 1, 7, 7) -> the equivalent filter over the unfolded input, reference (O, I, kT, kH, kW) layout."""
        if not self.hpair:
            return stem_equivalent_weight(w)
        o = w.shape[0]
        wp = torch.nn.functional.pad(w[:, :, 0], (0, 0, 1, 2))                 # (o, 3, kh' = kh+1 in 0..9, 7)
        wp = wp.reshape(o, 3, 5, 2, 7).permute(0, 3, 4, 1, 2).reshape(o, 2, 21, 5)      # [o, par, kw*3+ci, kh2]
        out = torch.zeros((o, 2, 32, 5), dtype=w.dtype, device=w.device)
        out[:, :, :21] = wp
        return out.reshape(o, 64, 1, 5, 1).contiguous()

    def weight_grad(self, dweq):
        """Inverse map for the gradient: equivalent-filter gradient -> (45, 3, 1, 7, 7)."""
        o = dweq.shape[0]
        if not self.hpair:
            return dweq.reshape(o, 7, 3, 1, 7).permute(0, 2, 3, 4, 1)         # dW[o, ci, 0, kh, kw] = dW_eq[o, kw*3+ci, 0, kh, 0]
        g = dweq.reshape(o, 2, 32, 5)[:, :, :21].reshape(o, 2, 7, 3, 5)       # [o, par, kw, ci, kh2]
        g = g.permute(0, 3, 4, 1, 2).reshape(o, 3, 10, 7)[:, :, 1:8]          # [o, ci, kh, kw]
        return g.unsqueeze(2)


class _Layer:
    __slots__ = ("spec", "desc", "w_packed", "scale", "shift", "out_shape", "src", "dst", "res", "is_stem")


class InferencePlan:
    """Shape-specialised eval-mode forward.  Buffers are allocated once and reused across calls."""

    def __init__(self, params, aux, model_depth, num_class, pool, eps, n, t, h, w, device):
        self.device = device
        self.n, self.t, self.h, self.w = n, t, h, w
        self.num_class = num_class
        self.pool = pool
        self.layers = []
        self.launches = 0
        self.eps = eps
        self.stale = False               # set when the weights moved: refresh() re-packs / re-folds IN PLACE (graphs stay valid)
        self._fold = ops.BnFoldTable(device)
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
            L.is_stem = w_override is not None
            L.w_packed = ops.pack_conv_weight(L.desc, wt)
            # eval-mode BatchNorm folded into the conv epilogue: all layers in one launch (fvt_bn_fold_multi), re-run in place
            # by refresh() when the weights / running statistics move
            L.scale, L.shift = self._fold.add(params[spec.bn + "_gamma"], params[spec.bn + "_beta"],
                                              aux[spec.bn + "_moving_mean"], aux[spec.bn + "_moving_var"], eps, cout_s)
            L.src = src
            L.dst = new_buf(dst_key, L.out_shape)
            L.res = res
            self.layers.append(L)
            return L.dst, L.out_shape

        # ---- stem: unfold + (1,7,1)/s(1,2,1) conv on K1, then the 3x1x1 temporal conv
        self._stage_of = {}              # block index -> stage key (the stage's output width)
        s_sp, s_tm = stem_specs()
        self.stem = StemGeometry(h)
        ushape = self.stem.unfold_shape(n, t, h, w)
        self.unfold = new_buf("unfold", ushape)
        cur, shp = add_layer(s_sp, ushape[:4], self.unfold, "mid", None,
                             w_override=self.stem.weight(params["conv1_middle_weight"].detach()),
                             kernel=self.stem.kernel, stride=self.stem.stride, pad=self.stem.pad, cin_store=self.stem.cin_store)
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
        self._fold.run()
        self._ptrs = self._param_ptrs(params, aux)
        self.fc_w = params["final_fc_weight"].detach().float().contiguous()
        self.fc_b = params["final_fc_bias"].detach().float().contiguous()
        tp, hp, wp = shp[1] - pool[0] + 1, shp[2] - pool[1] + 1, shp[3] - pool[2] + 1
        if (tp, hp, wp) != (1, 1, 1):
            raise ValueError("AvgPool3D%s over a %s map leaves %s: only a global pool (1x1x1 output) is supported; "
                             "pass final_temporal_kernel = T/8 and final_spatial_kernel = H/16 as the reference callers do"
                             % (pool, shp[1:4], (tp, hp, wp)))
        # ---- K2f: a stride-1 64 -> mid -> 64 unit (conv2_x, and the row-paired stem: (1,5,1) conv 64 -> 48 + 3x1x1 48 -> 64)
        #      runs as ONE launch with `mid` kept in tensor memory
        self.fused = {}
        if os.environ.get("FVT_FUSED_UNIT", "1") != "0":
            for i in range(len(self.layers) - 1):
                a, b = self.layers[i], self.layers[i + 1]
                if (a.spec.role in ("spatial", "stem_spatial") and b.spec.role in ("temporal", "temporal_out", "stem_temporal") and b.src == a.dst
                        and a.res is None and (i - 1) not in self.fused and ops.unit2p1_supported(a.desc, b.desc)):
                    self.fused[i] = b
        self.launches = 1 + len(self.layers) - len(self.fused) + 1
        self.use_graphs = os.environ.get("FVT_INFER_GRAPHS", "1") != "0"
        self._graphs, self._seen = {}, {}
        self.input_norm = None           # (scale, mean[3], inv_std[3]) for uint8 clips (R2Plus2D.set_input_normalization)

    @staticmethod
    def _param_ptrs(params, aux):
        return tuple(t.data_ptr() for t in list(params.values()) + list(aux.values()))

    def refresh(self, params, aux):
        """The weights or running statistics changed (optimiser step, load): re-pack the bf16 operand copies and re-fold the
        BatchNorms INTO THE EXISTING BUFFERS — 69 pack launches + one fold launch, no allocation, and the captured forward
        graphs (which hold those buffers' addresses) stay valid.  Returns False when the parameter storage itself moved
        (first training forward re-homes the parameters into the flat buffer; .to(device)): the caller rebuilds the plan."""
        if self._param_ptrs(params, aux) != self._ptrs:
            return False
        for L in self.layers:
            w = params[L.spec.name + "_weight"]
            ops.pack_conv_weight(L.desc, self.stem.weight(w.detach()) if L.is_stem else w, out=L.w_packed)
        self._fold.run()
        self.fc_w.copy_(params["final_fc_weight"].detach())
        self.fc_b.copy_(params["final_fc_bias"].detach())
        self.stale = False
        return True

    def _view(self, ref):
        key, shape = ref
        numel = 1
        for s in shape:
            numel *= s
        return self.bufs[key][:numel].view(shape)

    def forward(self, x, want_features=False, want_map=False):
        """x: (N, 3, T, H, W) fp32 CUDA -> logits (N, num_class) fp32 [, pooled features (N, 512)].

        The ~64 launches of a forward are captured into a CUDA graph per input buffer (keyed by the clip tensor's
        address, captured the third time the same buffer is seen, at most 4 graphs) and replayed: the launches then run
        back to back (12.32 -> 12.1 ms/step at batch 48).  FVT_INFER_GRAPHS=0 keeps every launch eager."""
        want = (self.n, self.t, self.h, self.w, 3) if x.dtype == torch.uint8 else (self.n, 3, self.t, self.h, self.w)
        assert tuple(x.shape) == want, (tuple(x.shape), want)
        x = x.contiguous()
        if not self.use_graphs or torch.cuda.is_current_stream_capturing():
            return self._forward_body(x, want_features, want_map)
        key = (x.data_ptr(), bool(want_features), bool(want_map))
        entry = self._graphs.get(key)
        if entry is None:
            seen = self._seen.get(key, 0) + 1
            if len(self._seen) > 64:
                self._seen.clear()
            self._seen[key] = seen
            if seen < 3 or len(self._graphs) >= 4:
                return self._forward_body(x, want_features, want_map)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with _capture(graph):
                out = self._forward_body(x, want_features, want_map)
            entry = (graph, out, x)                  # the clip tensor is kept alive: the graph reads its address
            self._graphs[key] = entry
        entry[0].replay()
        out = entry[1]
        return tuple(o.clone() for o in out) if isinstance(out, tuple) else out.clone()

    def _forward_body(self, x, want_features=False, want_map=False):
        self.stem.unfold(x, self._view(self.unfold), self.input_norm)
        skip = False
        for i, L in enumerate(self.layers):
            if skip:
                skip = False
                continue
            B = self.fused.get(i)
            if B is not None:
                ops.unit2p1_fwd(L.desc, B.desc, self._view(L.src), L.w_packed, L.scale, L.shift, B.w_packed, B.scale, B.shift,
                                self._view(B.res) if B.res is not None else None, out=self._view(B.dst))
                skip = True
                continue
            ops.conv3d_fwd(L.desc, self._view(L.src), L.w_packed, L.scale, L.shift,
                           self._view(L.res) if L.res is not None else None, out=self._view(L.dst))
        if want_map:                   # conv5_x output (N, T/8, H/16, W/16, 512) bf16 for heads other than pool + Dense
            return self._view(self.final).clone()
        return ops.pool_fc_fwd(self._view(self.final), 512, self.fc_w, self.fc_b, want_pooled=want_features)


class InferencePlanF32:
    """Eval-mode forward in plain fp32 (north-star "fp32 path", BASELINE configs[0]): the same layer plan on
    ops.conv3d_fwd_f32 / pool_fc_fwd_f32, activations NDHWC fp32 with the real channel counts.  Verification-grade speed
    (CUDA cores); it pins the layer semantics against an fp32 reference at 1e-4, which bf16 storage cannot."""

    def __init__(self, params, aux, model_depth, num_class, pool, eps, n, t, h, w, device):
        self.n, self.t, self.h, self.w = n, t, h, w
        self.pool = pool
        self.layers = []

        def add_layer(spec, in_shape, res):
            flags = (FVT_CONV_RELU if (spec.relu or res) else 0) | (FVT_CONV_RESIDUAL if res else 0)
            nn_, tt, hh, ww = in_shape[:4]
            d = ops.conv_desc(nn_, tt, hh, ww, spec.cin, spec.cout, spec.kernel, spec.stride, spec.pad, flags)
            wt = params[spec.name + "_weight"].detach().float().permute(2, 3, 4, 1, 0).contiguous()      # (kT,kH,kW,I,O)
            g, b = params[spec.bn + "_gamma"].detach().float(), params[spec.bn + "_beta"].detach().float()
            m, v = aux[spec.bn + "_moving_mean"].float(), aux[spec.bn + "_moving_var"].float()
            scale = (g / torch.sqrt(v + eps)).contiguous()
            shift = (b - m * scale).contiguous()
            to, ho, wo = ((tt + 2 * spec.pad[0] - spec.kernel[0]) // spec.stride[0] + 1,
                          (hh + 2 * spec.pad[1] - spec.kernel[1]) // spec.stride[1] + 1,
                          (ww + 2 * spec.pad[2] - spec.kernel[2]) // spec.stride[2] + 1)
            self.layers.append((d, wt, scale, shift))
            return (nn_, to, ho, wo, spec.cout)

        s_sp, s_tm = stem_specs()
        shp = add_layer(s_sp, (n, t, h, w), False)
        shp = add_layer(s_tm, shp, False)
        self.blocks = []
        for comp, cin, cout, down in network_blocks(model_depth):
            main, short = block_specs(comp, cin, cout, down)
            first = len(self.layers)
            x_shape = shp
            sa = add_layer(main[0], x_shape, False)
            sb = add_layer(main[1], sa, False)
            sc_ = add_layer(main[2], sb, False)
            has_short = short is not None
            if has_short:
                add_layer(short, x_shape, False)
            shp = add_layer(main[3], sc_, True)
            self.blocks.append((first, has_short))
        tp, hp, wp = shp[1] - pool[0] + 1, shp[2] - pool[1] + 1, shp[3] - pool[2] + 1
        if (tp, hp, wp) != (1, 1, 1):
            raise ValueError("AvgPool3D%s over a %s map leaves %s: only a global pool is supported" % (pool, shp[1:4], (tp, hp, wp)))
        self.fc_w = params["final_fc_weight"].detach().float().contiguous()
        self.fc_b = params["final_fc_bias"].detach().float().contiguous()
        self.launches = len(self.layers) + 1

    def _run(self, idx, x, res=None):
        d, wt, scale, shift = self.layers[idx]
        return ops.conv3d_fwd_f32(d, x, wt, scale, shift, res)

    def forward(self, x, want_features=False, want_map=False):
        assert tuple(x.shape) == (self.n, 3, self.t, self.h, self.w)
        cur = x.float().permute(0, 2, 3, 4, 1).contiguous()              # NCDHW -> NDHWC (layout only)
        cur = self._run(1, self._run(0, cur))
        for first, has_short in self.blocks:
            a = self._run(first, cur)
            b = self._run(first + 1, a)
            c = self._run(first + 2, b)
            if has_short:
                res = self._run(first + 3, cur)
                cur = self._run(first + 4, c, res)
            else:
                cur = self._run(first + 3, c, cur)
        if want_map:
            return cur
        return ops.pool_fc_fwd_f32(cur, self.fc_w, self.fc_b, want_pooled=want_features)


# =====================================================================================================================
# training
# =====================================================================================================================
class FlatParams:
    """All trainable tensors in ONE fp32 buffer (plus same-shaped gradient and momentum buffers).

    Layout, in forward order: conv weight (O, I, kT, kH, kW) | gamma padded to c_store | beta padded to c_store | ...
    | final_fc_weight | final_fc_bias.  BatchNorm gamma/beta are stored padded so the BN-backward kernel can write
    [dgamma | dbeta] straight into the gradient buffer; the user-visible parameters are views of the first c entries.
    A contiguous gradient buffer is what makes the NCCL all-reduce bucketing copy-free (slices of one tensor) and the
    SGD step a single launch.
    """

    def __init__(self, model_depth, num_class, device):
        self.slots = {}                 # name -> (offset, numel, shape)
        off = 0
        pshapes, _ = parameter_shapes(model_depth, num_class)
        for name, shape in pshapes.items():
            numel = 1
            for s in shape:
                numel *= s
            store = pad16(shape[0]) if (name.endswith("_gamma") or name.endswith("_beta")) else numel
            off = (off + 3) // 4 * 4                                  # 16-byte aligned slots
            self.slots[name] = (off, numel, tuple(shape), store)
            off += store
        self.total = (off + 3) // 4 * 4
        self.w = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.g = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.m = torch.zeros(self.total, dtype=torch.float32, device=device)

    def view(self, buf, name):
        """The tensor in the reference's shape.  Conv weights are STORED (O, kT, kH, kW, I) — input channels innermost, so
        the weight-gradient kernels add 32 consecutive floats per warp (FVT_CONV_W_OHWI) — and returned as a permuted
        (O, I, kT, kH, kW) view; everything else is stored as shaped."""
        off, numel, shape, _ = self.slots[name]
        if len(shape) == 5:
            o, i, kt, kh, kw = shape
            return buf[off:off + numel].view(o, kt, kh, kw, i).permute(0, 4, 1, 2, 3)
        return buf[off:off + numel].view(shape)

    def raw(self, buf, name):
        """Contiguous storage of a conv weight: (O, kT, kH, kW, I)."""
        off, numel, shape, _ = self.slots[name]
        o, i, kt, kh, kw = shape
        return buf[off:off + numel].view(o, kt, kh, kw, i)

    def padded(self, buf, name):
        off, _, _, store = self.slots[name]
        return buf[off:off + store]


import contextlib
import gc


@contextlib.contextmanager
def _capture(graph):
    """torch.cuda.graph(graph) with Python's cyclic garbage collector held off for the duration of the capture.  The
    capture runs in CUDA's global capture mode, where e.g. destroying another plan's graphs, streams or events is "not
    permitted while a stream is capturing" and invalidates the capture; torch collects garbage once on entry, but a
    capture of several hundred launches allocates enough Python objects to trigger further collections — which then
    finalise whatever cyclic garbage an earlier network left behind (seen as a flaky cudaErrorStreamCaptureInvalidated
    when two networks are trained one after the other in one process)."""
    was = gc.isenabled()
    gc.collect()
    gc.disable()
    try:
        with torch.cuda.graph(graph):
            yield
    finally:
        if was:
            gc.enable()


class _TLayer:
    __slots__ = ("spec", "fwd", "dgr", "w_name", "cin_real", "cout_real", "cin_s", "cout_s", "rows", "in_shape",
                 "out_shape", "src", "raw", "act", "stats", "scale", "shift", "mean", "invstd", "wp", "wpd", "w_eq",
                 "strided", "need_dgrad", "dplan", "bnacc", "dgr_bn", "draw_own", "group")


class TrainPlan:
    """Shape-specialised training step: forward with batch-statistics BatchNorm (reference semantics inside
    autograd.record(), model/R2Plus1.py:73-82 + A4) and the full backward pass, on the C-ABI kernels.

    forward:   per conv  K1(+stats) -> finalize + apply(+residual)(+ReLU) in one launch; bf16 operand copies of the weights
               re-packed on the side stream (first stage first)
    backward:  per conv  BatchNorm backward (one pass when the data gradient that produced its input fused the sums, else
               reduce + apply) -> data gradient (K1 / K1s2 / K1p; strided layers as parity sub-convolutions);
               weight gradients: deferred per residual stage into ONE grouped launch (K3g) on the side stream — strided and
               1x1x1 layers launch their own (K3) — and handed to the gradient reducer from that stream
    Everything is deterministic (exact BatchNorm sums, slice reductions in split order): two runs give the same bits.
    """

    def __init__(self, flat, aux, model_depth, num_class, pool, eps, n, t, h, w, device, momentum=0.9):
        self.flat, self.aux = flat, aux
        self.device = device
        self.eps, self.momentum = eps, momentum
        self.n, self.t, self.h, self.w = n, t, h, w
        # The whole step is deterministic: weight gradients reduce their pixel splits through workspace slices in a fixed
        # order (ops.wgrad_workspace, one per stream), BatchNorm sums use exact accumulators, split-K uses slices.  Every
        # gradient slot is OVERWRITTEN by backward() (MXNet grad_req='write'), so the gradient buffer is never zeroed.
        self.fwd_generation = 0          # bumped by every training forward: backward() refuses a stale forward
        self.num_class = num_class
        self.bufs = {}
        self.layers = {}
        self.blocks = []
        self.grad_hook = None            # callable(lo, hi): gradient slice [lo, hi) of flat.g is final
        self.weights_version = -1

        def buf(name, shape, dtype=torch.bfloat16):
            tns = torch.empty(shape, dtype=dtype, device=device)
            self.bufs[name] = tns
            return tns

        def make(spec, in_shape, src_name, kernel=None, stride=None, pad=None, cin_store=None, need_dgrad=True):
            L = _TLayer()
            L.spec = spec
            k, s, p = kernel or spec.kernel, stride or spec.stride, pad or spec.pad
            L.cin_real, L.cout_real = spec.cin, spec.cout
            L.cin_s, L.cout_s = cin_store or pad16(spec.cin), pad16(spec.cout)
            nn_, tt, hh, ww = in_shape[:4]
            L.fwd = ops.conv_desc(nn_, tt, hh, ww, L.cin_s, L.cout_s, k, s, p, ops.FVT_CONV_STATS)
            to, ho, wo = ops.conv_out_shape(L.fwd)
            L.in_shape = (nn_, tt, hh, ww, L.cin_s)
            L.out_shape = (nn_, to, ho, wo, L.cout_s)
            L.rows = nn_ * to * ho * wo
            L.strided = tuple(s) != (1, 1, 1)
            L.need_dgrad = need_dgrad
            L.dgr = ops.dgrad_desc(L.fwd) if need_dgrad else None
            L.w_name = spec.name + "_weight"
            L.src = src_name
            L.raw = buf(spec.name + ":raw", L.out_shape)
            L.act = buf(spec.name + ":act", L.out_shape)
            L.stats = None                     # view of self.stats_all, assigned once every layer is known
            for nm in ("scale", "shift", "mean", "invstd"):
                setattr(L, nm, buf(spec.name + ":" + nm, (L.cout_s,), torch.float32))
            L.wp = L.wpd = L.w_eq = None
            L.dplan = None
            L.bnacc = L.dgr_bn = None
            L.draw_own = L.group = None
            self.layers[spec.name] = L
            return L

        self._stage_of = {}              # block index -> stage key (the stage's output width)
        s_sp, s_tm = stem_specs()
        self.stem = StemGeometry(h)
        ushape = self.stem.unfold_shape(n, t, h, w)
        self.unfold = buf("unfold", ushape)
        self.stem0 = make(s_sp, ushape[:4], "unfold", kernel=self.stem.kernel, stride=self.stem.stride, pad=self.stem.pad,
                          cin_store=self.stem.cin_store, need_dgrad=False)
        self.stem0.cin_real = self.stem.cin_real
        self.stem1 = make(s_tm, self.stem0.out_shape, s_sp.name + ":act")
        cur_name, cur_shape = s_tm.name + ":act", self.stem1.out_shape
        max_elems = max(self.stem0.raw.numel(), self.stem1.raw.numel())
        for comp, cin, cout, down in network_blocks(model_depth):
            main, short = block_specs(comp, cin, cout, down)
            a = make(main[0], cur_shape, cur_name)
            b = make(main[1], a.out_shape, main[0].name + ":act")
            c = make(main[2], b.out_shape, main[1].name + ":act")
            d = make(main[3], c.out_shape, main[2].name + ":act")
            sc = make(short, cur_shape, cur_name) if short is not None else None
            self.blocks.append((comp, cur_name, cur_shape, a, b, c, d, sc))
            self._stage_of[comp] = cout
            cur_name, cur_shape = main[3].name + ":act", d.out_shape     # block output lives in d.act
            for L in (a, b, c, d):
                max_elems = max(max_elems, L.raw.numel(), L.in_shape[0] * L.in_shape[1] * L.in_shape[2] * L.in_shape[3] * L.in_shape[4])
        self.final_name, self.final_shape = cur_name, cur_shape
        # per-channel (sum, sum^2) exact accumulators of every conv in ONE buffer: a single memset per step
        self.stats_all = ops.stats_buffer(sum(L.cout_s for L in self.layers.values()), device)
        # exact accumulators of the BatchNorm backward sums that the fused data-gradient epilogues fill (FVT_CONV_BN_BWD)
        self.fuse_bnbwd = os.environ.get("FVT_FUSE_BNBWD", "1") != "0"
        # Measured (batch 4): the fused epilogue pays where launches dominate (conv3_x .. conv5_x); on the large-M layers
        # (stem, conv2_x: 401408 positions) its per-row reads and 31 shuffles per 16 columns make a short-K data gradient
        # epilogue-bound (conv2_x 64 -> 144 temporal: 41 -> 122 us, more than the reduce pass it saves)
        self.fuse_bnbwd_rows = int(os.environ.get("FVT_FUSE_BNBWD_ROWS", "200000"))
        self.bnacc_all = ops.stats_buffer(sum(L.cout_s for L in self.layers.values()), device)
        off = 0
        for L in self.layers.values():
            L.stats = self.stats_all[off:off + 2 * L.cout_s]
            L.bnacc = self.bnacc_all[off:off + 2 * L.cout_s]
            off += 2 * L.cout_s
        tp, hp, wp = cur_shape[1] - pool[0] + 1, cur_shape[2] - pool[1] + 1, cur_shape[3] - pool[2] + 1
        if (tp, hp, wp) != (1, 1, 1):
            raise ValueError("AvgPool3D%s over a %s map leaves %s: only a global pool is supported" % (pool, cur_shape[1:4], (tp, hp, wp)))
        # backward scratch: two activation-gradient ping-pong buffers, one raw-gradient buffer, the masked block
        # gradient, the shortcut gradient and the zero-insert staging area
        up_elems = 1
        for L in self.layers.values():
            if L.strided and L.need_dgrad:
                up_elems = max(up_elems, L.fwd.n * L.fwd.t * L.fwd.h * L.fwd.w * L.cout_s)
        # raw-gradient scratch buffers rotate so that a weight gradient still reading one on the side stream never sees it
        # overwritten; more of them let the data-gradient chain run further ahead of the weight gradients
        self._n_draw = max(3, int(os.environ.get("FVT_DRAW_BUFS", "6")))
        for nm in ["gA", "gB", "gmask", "gshort", "draw_s"] + ["draw%d" % i for i in range(self._n_draw)]:
            buf(nm, (max_elems,))
        buf("up", (up_elems,))
        # Deferred, grouped weight gradients (ops.WgradGroup): in the small-feature-map stages a single layer cannot fill the
        # machine (conv4_x / conv5_x at batch 4: 35-55 us per layer against 2-12 us of tensor work, and every launch
        # occupies all SMs, so the data-gradient chain waits behind it).  A weight gradient feeds nothing but the optimiser:
        # the stride-1 layers of such a stage keep their raw gradient in a buffer of their own and ONE launch at the end of
        # the stage's backward computes all of them.  FVT_WGRAD_GROUP=0: every layer launches its own (round-1 behaviour);
        # FVT_WGRAD_GROUP_ROWS: largest output-position count per layer that is still deferred.
        self._groups = {}                # stage key -> [layers]  (ops.WgradGroup built on first use: needs flat.g)
        self._group_obj = {}
        if os.environ.get("FVT_WGRAD_GROUP", "1") != "0":
            max_rows = int(os.environ.get("FVT_WGRAD_GROUP_ROWS", "1000000"))
            for comp, xin_name, xin_shape, a, b, c, d, sc in self.blocks:
                cand = [L for L in (a, b, c, d) if L.rows <= max_rows]
                if not cand:
                    continue
                ok = ops.wgrad_group_eligible([L.fwd for L in cand], [L.cout_real for L in cand], [L.cin_real for L in cand], device)
                for L, e in zip(cand, ok):
                    if e:
                        L.group = self._stage_of[comp]
                        L.draw_own = buf(L.spec.name + ":draw", L.out_shape)
                        self._groups.setdefault(L.group, []).append(L)
        self.pooled = None
        self._dweq = None                # gradient of the stem's equivalent filter (re-mapped into conv1_middle_weight's slot)
        self.launches_fwd = self.launches_bwd = 0
        # CUDA graphs: the ~700 launches of a step are captured once (forward graph, backward graph) and replayed, so
        # the step is not paced by Python/ctypes launch overhead.  FVT_CUDA_GRAPHS=0 runs every launch eagerly.
        self.use_graphs = os.environ.get("FVT_CUDA_GRAPHS", "1") != "0"
        self.x_static = torch.empty((n, 3, t, h, w), dtype=torch.float32, device=device)
        self.dlogits_static = torch.empty((n, num_class), dtype=torch.float32, device=device)
        self._fwd_graph = self._bwd_graph = None
        self._logits_static = None
        self._warm_fwd = self._warm_bwd = 0
        self.finish_hook = None          # callable(): wait for gradient reductions launched by grad_hook
        self.input_norm = None           # (scale, mean[3], inv_std[3]) for uint8 clips
        # Second stream (a parallel branch of the captured graphs): weight re-packing runs beside the first forward
        # layers, and every weight gradient runs beside the data-gradient chain it does not feed.  At batch 4 the
        # conv4_x / conv5_x launches fill 14-49 of the 148 SMs, so the two branches genuinely overlap.
        self.side = torch.cuda.Stream(device=device) if os.environ.get("FVT_SIDE_STREAM", "1") != "0" else None
        self._packed_ev = self._packed_ev_early = None
        self._pack_f = self._pack_d = self._pack_f_early = None
        self._early = set()
        self.dgrad_direct = os.environ.get("FVT_DGRAD_DIRECT", "1") != "0"
        self._busy = {}                  # scratch buffer name -> event recorded after its last reader on the side stream
        self._draw_i = 0
        # timing experiments only (tools/gpu_train_ablate.py): FVT_SKIP=wgrad,dgrad,bnbwd leaves those launches out, which
        # shows each family's marginal cost inside the replayed graph (results are garbage then)
        self._skip = set(filter(None, os.environ.get("FVT_SKIP", "").split(",")))
        # timing experiments only (tools/gpu_train_stages.py): FVT_STAGE_EVENTS=1 records an external timing event on the
        # main stream after the stem and after every residual block, inside the captured graphs too, so the replayed step
        # can be read as a per-stage timeline
        self._marks = {} if os.environ.get("FVT_STAGE_EVENTS", "0") == "1" else None

    def _mark(self, label):
        if self._marks is None:
            return
        ev = self._marks.get(label)
        if ev is None:
            ev = self._marks[label] = torch.cuda.Event(enable_timing=True, external=True)
        ev.record(torch.cuda.current_stream(self.device))

    def stage_times(self):
        """[(label, ms since the previous mark)] of the last executed step (FVT_STAGE_EVENTS=1; synchronises)."""
        torch.cuda.synchronize(self.device)
        labels = list(self._marks)
        return [(b, self._marks[a].elapsed_time(self._marks[b])) for a, b in zip(labels, labels[1:])]

    def set_hooks(self, grad_hook, finish_hook):
        """Attach (or change) the gradient-reduction hooks.  They are baked into the captured backward graph, so a change
        drops the graphs (re-captured on the next steps)."""
        if grad_hook is self.grad_hook and finish_hook is self.finish_hook:
            return
        self.grad_hook, self.finish_hook = grad_hook, finish_hook
        self._fwd_graph = self._bwd_graph = None
        self._warm_fwd = self._warm_bwd = 0

    # ------------------------------------------------------------------ weights
    def _w(self, L):
        return self.flat.view(self.flat.w, L.w_name)

    def _w_raw(self, L):
        return self.flat.raw(self.flat.w, L.w_name)

    def _build_pack_tables(self):
        """Operand copies of all conv weights in TWO launches (forward layouts, then data-gradient layouts) instead of 137
        (ops.PackTable); the stem's equivalent filter (a 45 x 64 x 5 tensor re-expressed over the W-unfolded input) keeps
        its own small path."""
        self._pack_f, self._pack_d = ops.PackTable(self.device), ops.PackTable(self.device)
        # the stem's temporal conv and the first stage (conv2_x: 1 MB of weights) get a launch of their own, so that the
        # forward pass does not wait for the 250 MB of conv3_x..conv5_x filters before its second convolution (measured:
        # the main stream idled ~0.19 ms at the start of every step behind the one-launch pack)
        self._pack_f_early = ops.PackTable(self.device)
        first_stage = self._stage_of[self.blocks[0][0]] if self.blocks else None
        self._early = {self.stem1.spec.name}
        for comp, _, _, a, b, c, d, sc in self.blocks:
            if self._stage_of[comp] == first_stage:
                self._early.update(L.spec.name for L in (a, b, c, d) + ((sc,) if sc is not None else ()))
        for L in self.layers.values():
            if L is self.stem0:
                continue
            L.wp = (self._pack_f_early if L.spec.name in self._early else self._pack_f).add_fwd(L.fwd, self._w_raw(L))
        for L in reversed(list(self.layers.values())):
            if not L.need_dgrad:
                continue
            if L.strided and self.dgrad_direct:
                # data gradient of a strided convolution: one stride-1 sub-convolution of dY per parity class of dX
                # (ops.DgradPlan) instead of zero-insert + a full convolution over the input extent
                L.dplan = ops.DgradPlan(L.fwd, self._w_raw(L), self._pack_d)
            else:
                L.wpd = self._pack_d.add_dgrad(L.dgr, self._w_raw(L))

    def refresh_weights(self, version):
        """Re-pack bf16 operand copies of the fp32 master weights (forward and data-gradient layouts).  With the side
        stream the packing launches form a branch parallel to the forward pass: the forward-layout copies (the stem's
        first, then everything else in one launch; the first conv after the stem waits for it, `_wait_packed`), then the
        data-gradient copies, joined by `_join_side()` at the end of the forward pass."""
        self._packed_ev = self._packed_ev_early = None
        if version == self.weights_version:
            return
        if self._pack_f is None:
            self._build_pack_tables()
        main = torch.cuda.current_stream(self.device)
        L0 = self.stem0
        if self.side is None:
            L0.wp = ops.pack_conv_weight(L0.fwd, self.stem.weight(self._w(L0)), out=L0.wp)
            self._pack_f_early.run()
            self._pack_f.run()
            self._pack_d.run()
        else:
            self.side.wait_stream(main)
            L0.wp = ops.pack_conv_weight(L0.fwd, self.stem.weight(self._w(L0)), out=L0.wp)     # tiny: on the main stream
            with torch.cuda.stream(self.side):
                self._pack_f_early.run()
                self._packed_ev_early = torch.cuda.Event()
                self._packed_ev_early.record(self.side)
                self._pack_f.run()
                self._packed_ev = torch.cuda.Event()
                self._packed_ev.record(self.side)
                self._pack_d.run()
        self.weights_version = version

    def _wait_packed(self, L):
        if L is self.stem0:
            return
        if L.spec.name in self._early:
            if self._packed_ev_early is not None:
                torch.cuda.current_stream(self.device).wait_event(self._packed_ev_early)
                self._packed_ev_early = None
            return
        if self._packed_ev is not None:
            torch.cuda.current_stream(self.device).wait_event(self._packed_ev)
            self._packed_ev = None

    def _join_side(self):
        if self.side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.side)

    # ------------------------------------------------------------------ forward
    def _bn_names(self, L):
        return L.spec.bn + "_gamma", L.spec.bn + "_beta", L.spec.bn + "_moving_mean", L.spec.bn + "_moving_var"

    def _conv_bn(self, L, src, apply=None):
        """conv (+ per-channel statistics) -> BatchNorm.  apply=None: finalize only (projection shortcut, whose affine is
        applied inside the block's last fused pass); apply=dict(relu, res, res_scale, res_shift): finalize + apply in
        one launch."""
        gname, bname, mname, vname = self._bn_names(L)
        self._wait_packed(L)
        ops.conv3d_fwd(L.fwd, src, L.wp, out=L.raw, stats=L.stats)
        args = (L.stats, self.flat.view(self.flat.w, gname), self.flat.view(self.flat.w, bname),
                self.aux[mname], self.aux[vname], L.cout_s, L.rows, self.eps, self.momentum,
                L.scale, L.shift, L.mean, L.invstd)
        if apply is None:
            ops.bn_finalize(*args)
        else:
            ops.bn_finalize_apply(*args, L.raw, L.act, True, **apply)

    def forward(self, x, weights_version):
        """Training-mode forward: re-pack bf16 operand copies if the weights changed, run the network.  Returns fp32
        logits (N, num_class)."""
        want = (self.n, self.t, self.h, self.w, 3) if x.dtype == torch.uint8 else (self.n, 3, self.t, self.h, self.w)
        assert tuple(x.shape) == want, (tuple(x.shape), want)
        if x.dtype != self.x_static.dtype:           # the captured graphs read one static input buffer: re-capture for the other form
            self.x_static = torch.empty(want, dtype=x.dtype, device=self.device)
            self._fwd_graph = None
            self._warm_fwd = 0
        self.fwd_generation += 1
        if not self.use_graphs:
            self.refresh_weights(weights_version)
            out = self._forward_body(x.contiguous())
            self._join_side()
            return out
        self.x_static.copy_(x)
        if self._fwd_graph is None:
            if self._warm_fwd < 1:                      # first call: eager (one-time initialisation inside the library)
                self._warm_fwd += 1
                self.refresh_weights(weights_version)
                out = self._forward_body(self.x_static)
                self._join_side()
                return out
            graph = torch.cuda.CUDAGraph()
            with _capture(graph):
                self.weights_version = -1               # the graph always re-packs: weights change every step
                self.refresh_weights(0)
                self._logits_static = self._forward_body(self.x_static)
                self._join_side()
            self._fwd_graph = graph
        self._fwd_graph.replay()
        return self._logits_static.clone()

    def forward_map(self, x, weights_version):
        """Training-mode forward of the TRUNK only (eager launches): returns the conv5_x output (N, T/8, H/16, W/16, 512)
        bf16 — the input of heads other than pool + Dense (multi-task scene/action heads).  The buffer is
        plan-owned: backward_map() must run before the next forward of this shape."""
        self.fwd_generation += 1
        self.refresh_weights(weights_version)
        self._forward_body(x.contiguous(), head=False)
        self._join_side()
        return self.bufs[self.final_name]

    def backward_map(self, dmap):
        """Backward of forward_map(): dmap is the gradient w.r.t. the conv5_x output, bf16, same shape."""
        self._backward_body(None, dmap=dmap.contiguous())

    def _forward_body(self, x, head=True):
        self._mark("fwd:start")
        self.stats_all.zero_()
        self.stem.unfold(x, self.unfold, self.input_norm)
        B = self.bufs
        for L in (self.stem0, self.stem1):
            self._conv_bn(L, B[L.src], apply={})
        self._mark("fwd:stem")
        for comp, xin_name, xin_shape, a, b, c, d, sc in self.blocks:
            xin = B[xin_name]
            for L in (a, b, c):
                self._conv_bn(L, B[L.src], apply={})
            if sc is not None:
                self._conv_bn(sc, xin)
                self._conv_bn(d, B[d.src], apply=dict(res=sc.raw, res_scale=sc.scale, res_shift=sc.shift))
            else:
                self._conv_bn(d, B[d.src], apply=dict(res=xin))
            self._mark("fwd:block%s" % comp)
        if not head:
            return None
        logits, self.pooled = ops.pool_fc_fwd(B[self.final_name], 512, self.flat.view(self.flat.w, "final_fc_weight"),
                                              self.flat.view(self.flat.w, "final_fc_bias"), want_pooled=True)
        self._mark("fwd:head")
        return logits

    # ------------------------------------------------------------------ backward
    def _view(self, name, shape):
        numel = 1
        for s in shape:
            numel *= s
        return self.bufs[name][:numel].view(shape)

    def _ready(self, *names):
        if self.grad_hook is None:
            return
        lo = min(self.flat.slots[n][0] for n in names)
        hi = max(self.flat.slots[n][0] + self.flat.slots[n][3] for n in names)
        self.grad_hook(lo, hi)

    def _ready_from_side(self, name_lists):
        """Hand finished gradient slices to the reducer FROM the side stream: a block's (or a stage's grouped) weight
        gradients finish there, and the all-reduce is then ordered behind them without the main stream waiting — it goes
        on with the next block's data-gradient chain.  (Round 1 joined the side stream into the main stream here; with the
        grouped launches that serialised 1.6 ms of weight-gradient kernels with the chain whenever a reducer was attached:
        2 GPUs 9.76 -> 9.43 ms/step, the single-GPU time.)"""
        if self.grad_hook is None:
            return
        if self.side is None:
            for nm in name_lists:
                self._ready(*nm)
            return
        self.side.wait_stream(torch.cuda.current_stream(self.device))     # BatchNorm parameter gradients come from the main stream
        with torch.cuda.stream(self.side):
            for nm in name_lists:
                self._ready(*nm)

    def _bn_bwd(self, L, dact, mask, draw, dz_out=None):
        """mask: True = ReLU directly after this BatchNorm (mask recomputed from raw); a tensor = ReLU after a residual join
        (mask by that tensor's sign); None = no ReLU; "fused" = `dact` already is the masked gradient and L.bnacc holds
        [sum dz*raw, sum dz], both written by the data-gradient convolution's epilogue (_dgrad(fuse_bn=L))."""
        if "bnbwd" in self._skip:
            return
        gname, bname, _, _ = self._bn_names(L)
        sums = self.flat.padded(self.flat.g, gname)           # [dgamma(c_store) | dbeta(c_store)] adjacent slots
        off_g, off_b = self.flat.slots[gname][0], self.flat.slots[bname][0]
        assert off_b == off_g + L.cout_s, "gamma/beta slots must be adjacent"
        sums2 = self.flat.g[off_g:off_g + 2 * L.cout_s]
        if isinstance(mask, str):
            ops.bn_backward(L.raw, dact, None, L.mean, L.invstd, self.flat.view(self.flat.w, gname), sums2, draw, None,
                            sums_acc=L.bnacc, dz_in=2)
        elif mask is True:        # ReLU directly after this BatchNorm: recompute the mask from raw (one read less)
            ops.bn_backward(L.raw, dact, None, L.mean, L.invstd, self.flat.view(self.flat.w, gname), sums2, draw, dz_out,
                            relu_scale=L.scale, relu_shift=L.shift)
        else:
            ops.bn_backward(L.raw, dact, mask, L.mean, L.invstd, self.flat.view(self.flat.w, gname), sums2, draw, dz_out)

    def _draw(self, shape, key=None, own=None):
        """Next raw-gradient scratch buffer (three rotate, so a weight gradient still reading one on the side stream
        never sees it overwritten); waits for that buffer's last side-stream reader.  own: a layer whose weight gradient is
        deferred to its stage's grouped launch keeps its raw gradient in a buffer of its own."""
        if own is not None and own.draw_own is not None:
            v = own.draw_own
            v._fvt_key = None
            return v
        if key is None:
            key = "draw%d" % self._draw_i
            self._draw_i = (self._draw_i + 1) % self._n_draw
        ev = self._busy.pop(key, None)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
        v = self._view(key, shape)
        v._fvt_key = key
        return v

    def _wgrad(self, L, x_in, draw):
        """Weight gradient of L; on the side stream when there is one (it feeds nothing but the gradient buffer)."""
        if L.group is not None:           # deferred: computed by the stage's grouped launch (_run_group)
            assert draw is L.draw_own
            return
        if self.side is None:
            return self._wgrad_now(L, x_in, draw)
        main = torch.cuda.current_stream(self.device)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            self._wgrad_now(L, x_in, draw)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self._busy[draw._fvt_key] = ev

    def _run_group(self, key):
        """The deferred weight gradients of a stage in one launch (ops.WgradGroup), on the side stream when there is one."""
        layers = self._groups.get(key)
        if not layers or "wgrad" in self._skip:
            return
        g = self._group_obj.get(key)
        if g is None or g.grad_ptr != self.flat.g.data_ptr():
            B = self.bufs
            g = ops.WgradGroup([(L.fwd, B[L.src], L.draw_own, self.flat.raw(self.flat.g, L.w_name), L.cout_real, L.cin_real)
                                for L in layers], self.device)
            assert all(g.in_group)
            g.grad_ptr = self.flat.g.data_ptr()
            self._group_obj[key] = g
        if self.side is None:
            return g.run()
        self.side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.side):
            g.run()

    def _wgrad_now(self, L, x_in, draw):
        if "wgrad" in self._skip:
            return
        if L is self.stem0:
            k = self.stem.kernel
            if self._dweq is None:
                self._dweq = torch.zeros((45, self.stem.cin_real, k[0], k[1], k[2]), dtype=torch.float32, device=self.device)
            ops.conv3d_wgrad(L.fwd, x_in, draw, self._dweq, 45, self.stem.cin_real)
            self.flat.view(self.flat.g, L.w_name).copy_(self.stem.weight_grad(self._dweq))
        else:
            ops.conv3d_wgrad(L.fwd, x_in, draw, self.flat.raw(self.flat.g, L.w_name), L.cout_real, L.cin_real, ohwi=True)

    def _can_fuse_bn(self, L):
        """The data gradient of L feeds exactly one BatchNorm backward whose ReLU mask comes from its own raw output: fusable
        unless L is strided (its data gradient runs as parity sub-convolutions onto a lattice)."""
        if not (self.fuse_bnbwd and L.dplan is None and not L.strided and L.rows <= self.fuse_bnbwd_rows and
                "dgrad" not in self._skip and "bnbwd" not in self._skip):
            return False
        # The fused epilogue needs the whole reduction in one CTA: a data gradient with few output tiles and a long
        # reduction (conv5_x 1152 -> 512: 28 tiles, 162 k-blocks) runs faster split over K with the two-pass BatchNorm
        # backward behind it (measured inside a replayed graph: 27.7 us against 46.9 us fused)
        k = L.spec.kernel
        tiles = ((L.rows + 127) // 128) * ((L.cin_s + 127) // 128)
        k_blocks = k[0] * k[1] * k[2] * ((L.cout_s + 63) // 64)
        return not (2 * tiles <= 148 and k_blocks >= 32)

    def _dgrad(self, L, draw, out, residual=None, fuse_bn=None):
        """Data gradient of L.  fuse_bn = the producer layer P of L's input (act_P = relu(bn(raw_P))): the epilogue masks with
        P's ReLU and accumulates P's BatchNorm backward sums (FVT_CONV_BN_BWD), `out` then holds dz and _bn_bwd(P, out,
        "fused", ...) runs its apply pass only."""
        if "dgrad" in self._skip:
            return out
        if fuse_bn is not None:
            P = fuse_bn
            if L.dgr_bn is None:
                L.dgr_bn = ops.ConvDesc(*L.dgr.key())
                L.dgr_bn.flags = ops.FVT_CONV_STATS | ops.FVT_CONV_BN_BWD | ops.FVT_CONV_RESIDUAL
            ops.conv3d_fwd(L.dgr_bn, draw, L.wpd, scale=P.scale, shift=P.shift, residual=P.raw, out=out, stats=P.bnacc)
            return out
        if L.dplan is not None:
            return L.dplan.run(draw, out, residual)
        src = draw
        if L.strided:
            src = ops.zero_insert(draw, L.fwd, out=self._view("up", (L.fwd.n, L.fwd.t, L.fwd.h, L.fwd.w, L.cout_s)))
        d = L.dgr
        if residual is not None:
            d = ops.ConvDesc(*d.key())
            d.flags = ops.FVT_CONV_RESIDUAL
        ops.conv3d_fwd(d, src, L.wpd, residual=residual, out=out)
        return out

    def backward(self, dlogits):
        """dlogits: (N, num_class) fp32.  Overwrites every slot of flat.g (grad_req='write')."""
        if not self.use_graphs:
            return self._backward_body(dlogits.contiguous())
        self.dlogits_static.copy_(dlogits)
        if self._bwd_graph is None:
            if self._warm_bwd < 1:
                self._warm_bwd += 1
                return self._backward_body(self.dlogits_static)
            graph = torch.cuda.CUDAGraph()
            with _capture(graph):
                self._backward_body(self.dlogits_static)
                if self.finish_hook is not None:        # gradient all-reduces launched during capture join here
                    self.finish_hook()
            self._bwd_graph = graph
        self._bwd_graph.replay()

    def _backward_body(self, dlogits, dmap=None):
        B = self.bufs
        fl = self.flat
        g_cur = self._view("gA", self.final_shape)
        self._mark("bwd:start")
        if self.fuse_bnbwd:
            self.bnacc_all.zero_()                 # one memset for every fused BatchNorm backward of the step
        if dmap is not None:                       # the trunk's own pool + Dense head is not part of this graph
            g_cur.copy_(dmap)
            fl.view(fl.g, "final_fc_weight").zero_()
            fl.view(fl.g, "final_fc_bias").zero_()
        else:
            ops.pool_fc_bwd(dlogits, self.pooled, fl.view(fl.w, "final_fc_weight"), fl.view(fl.g, "final_fc_weight"),
                            fl.view(fl.g, "final_fc_bias"), g_cur)
        self._ready("final_fc_weight", "final_fc_bias")
        self._mark("bwd:head")
        cur_key = "gA"
        pending = []                                  # parameter names of a deferred stage's blocks, reduced after its group
        rblocks = list(reversed(self.blocks))
        for bi, (comp, xin_name, xin_shape, a, b, c, d, sc) in enumerate(rblocks):
            xin = B[xin_name]
            other = "gB" if cur_key == "gA" else "gA"
            gmask = self._view("gmask", d.out_shape)
            draw_d = self._draw(d.out_shape, own=d)
            self._bn_bwd(d, g_cur, d.act, draw_d, dz_out=gmask)           # out = relu(bn2 + shortcut): mask by out > 0
            if sc is not None:
                draw_s = self._draw(sc.out_shape, key="draw_s")
                self._bn_bwd(sc, gmask, None, draw_s)
                self._wgrad(sc, xin, draw_s)
                gshort = self._dgrad(sc, draw_s, self._view("gshort", xin_shape))
            else:
                gshort = gmask
            self._wgrad(d, c.act, draw_d)
            # (data gradient -> the BatchNorm backward it feeds) pairs: the ReLU mask and the per-channel sums are taken in
            # the convolution's epilogue where possible, so the BatchNorm backward is one pass instead of two
            f = self._can_fuse_bn(d)
            gc = self._dgrad(d, draw_d, self._view(other, c.out_shape), fuse_bn=c if f else None)
            draw_c = self._draw(c.out_shape, own=c)
            self._bn_bwd(c, gc, "fused" if f else True, draw_c)
            self._wgrad(c, b.act, draw_c)
            f = self._can_fuse_bn(c)
            gb = self._dgrad(c, draw_c, self._view(cur_key, b.out_shape), fuse_bn=b if f else None)
            draw_b = self._draw(b.out_shape, own=b)
            self._bn_bwd(b, gb, "fused" if f else True, draw_b)
            self._wgrad(b, a.act, draw_b)
            f = self._can_fuse_bn(b)
            ga = self._dgrad(b, draw_b, self._view(other, a.out_shape), fuse_bn=a if f else None)
            draw_a = self._draw(a.out_shape, own=a)
            self._bn_bwd(a, ga, "fused" if f else True, draw_a)
            self._wgrad(a, xin, draw_a)
            g_cur = self._dgrad(a, draw_a, self._view(cur_key, xin_shape), residual=gshort)
            names = []
            for L in (a, b, c, d) + ((sc,) if sc is not None else ()):
                names += [L.w_name, L.spec.bn + "_gamma", L.spec.bn + "_beta"]
            stage = self._stage_of[comp]
            stage_ends = bi + 1 == len(rblocks) or self._stage_of[rblocks[bi + 1][0]] != stage
            if stage in self._groups:
                pending.append(names)
                if stage_ends:
                    self._run_group(stage)
                    if bi + 1 < len(rblocks):             # the last stage's slices go out together with the stem's (below):
                        self._ready_from_side(pending)    # one small all-reduce at the end of the step instead of two
                        pending = []
            else:
                self._ready_from_side([names])
            self._mark("bwd:block%s" % comp)
        # stem
        other = "gB" if cur_key == "gA" else "gA"
        draw1 = self._draw(self.stem1.out_shape)
        self._bn_bwd(self.stem1, g_cur, True, draw1)
        self._wgrad(self.stem1, self.stem0.act, draw1)
        f = self._can_fuse_bn(self.stem1)
        g0 = self._dgrad(self.stem1, draw1, self._view(other, self.stem0.out_shape), fuse_bn=self.stem0 if f else None)
        draw0 = self._draw(self.stem0.out_shape)
        self._bn_bwd(self.stem0, g0, "fused" if f else True, draw0)
        self._wgrad(self.stem0, self.unfold, draw0)
        names = []
        for L in (self.stem0, self.stem1):
            names += [L.w_name, L.spec.bn + "_gamma", L.spec.bn + "_beta"]
        self._mark("bwd:stem")
        self._ready_from_side(pending + [names])          # behind the stem's weight gradients on the side stream
        self._join_side()
        self._busy.clear()
        self._mark("bwd:joined")
