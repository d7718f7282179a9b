"""Stand-alone block builders with the reference's names (model/R2Plus1.py:19-40, 42-82, 295-347).

`R2Plus2D` (R2Plus1.py in this package) runs the whole network through shape-specialised plans; the builders here are
the composable pieces a reference user can also instantiate on their own:

    get_spatial_temporal_conv(in_filters, out_filter, stride, use_bias=False)   -> the (2+1)D unit
    R3DBlock(input_filter, num_filter, comp_index, downsampling, ...)            -> one residual block
    get_R2plus1d(num_class, no_bias, model_depth, final_spatial_kernel, final_temporal_kernel) -> Sequential net

Every convolution / BatchNorm / ReLU / add goes through the same C-ABI kernels (ops.conv3d_fwd with the BatchNorm
folded into the epilogue in eval mode; conv + statistics -> bn_finalize -> bn_apply in training mode).  Inputs and
outputs are the reference's NCDHW fp32 tensors on a CUDA device; inside, activations are NDHWC bf16.  In training mode
(module.train() with gradients enabled) every Conv3D -> BatchNorm (-> +residual) (-> ReLU) group is ONE autograd node
(`_ConvBnFn`) whose backward runs the BatchNorm-backward, weight-gradient and data-gradient kernels — eager launches, for
the heads and stand-alone blocks; the full network trains through R2Plus2D's captured plans.
"""
import torch

from .. import ops
from ..engine import BLOCK_CONFIG, middle_filters
from ..ops import FVT_CONV_RELU, FVT_CONV_RESIDUAL, FVT_CONV_STATS, pad16

BN_EPS = 1e-5          # gluon nn.BatchNorm() default
BN_MOMENTUM = 0.9


def _xavier(shape):
    fan_in = shape[1] * shape[2] * shape[3] * shape[4]
    fan_out = shape[0] * shape[2] * shape[3] * shape[4]
    s = (3.0 / ((fan_in + fan_out) / 2.0)) ** 0.5
    return torch.empty(shape).uniform_(-s, s)


def to_ndhwc(x, c_store=None):
    """(N, C, T, H, W) fp32 -> (N, T, H, W, C_store) bf16 with zero pad channels."""
    n, c, t, h, w = x.shape
    cs = c_store or pad16(c)
    out = torch.zeros((n, t, h, w, cs), dtype=torch.bfloat16, device=x.device)
    out[..., :c] = x.permute(0, 2, 3, 4, 1)
    return out


def to_ncdhw(y, c):
    return y[..., :c].permute(0, 4, 1, 2, 3).float().contiguous()


class _ToNdhwc(torch.autograd.Function):
    """Layout change at a block's boundary that gradients pass through: (N,C,T,H,W) fp32 <-> (N,T,H,W,C_store) bf16."""

    @staticmethod
    def forward(ctx, x):
        ctx.c = x.shape[1]
        return to_ndhwc(x)

    @staticmethod
    def backward(ctx, g):
        return to_ncdhw(g, ctx.c)


class _FromNdhwc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, c):
        ctx.cs = y.shape[-1]
        return to_ncdhw(y, c)

    @staticmethod
    def backward(ctx, g):
        return to_ndhwc(g, ctx.cs), None


def _to_ndhwc(x):
    return _ToNdhwc.apply(x) if (x.requires_grad and torch.is_grad_enabled()) else to_ndhwc(x)


def _from_ndhwc(y, c):
    return _FromNdhwc.apply(y, c) if (y.requires_grad and torch.is_grad_enabled()) else to_ncdhw(y, c)


class _ConvBnFn(torch.autograd.Function):
    """Training-mode Conv3D -> BatchNorm(batch statistics) [-> + residual] [-> ReLU] on NDHWC bf16 activations:
    forward  K1(+statistics) -> bn_finalize (MXNet running-stat update) -> bn_apply;
    backward bn_backward (ReLU mask recomputed from raw, or taken from the block output when a residual joined)
             -> conv3d_wgrad -> data gradient (fvt_conv3d_fwd with the transposed filter; parity sub-convolutions when strided).
    A conv bias in front of a batch-statistics BatchNorm cancels in the output (the batch mean absorbs it): it only shifts the
    running mean and its gradient is exactly zero."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, residual, conv, bn, relu):
        n, t, h, w, cs = x.shape
        cout, cout_s = conv.channels, pad16(conv.channels)
        d = ops.conv_desc(n, t, h, w, cs, cout_s, conv.kernel, conv.strides, conv.padding, FVT_CONV_STATS)
        wp = ops.pack_conv_weight(d, weight.detach())
        stats = ops.stats_buffer(cout_s, x.device)
        raw = ops.conv3d_fwd(d, x, wp, stats=stats)
        rows = raw.numel() // cout_s
        scale, shift, mean, invstd = (torch.empty(cout_s, dtype=torch.float32, device=x.device) for _ in range(4))
        ops.bn_finalize(stats, gamma.detach(), beta.detach(), bn.running_mean, bn.running_var, cout_s, rows, bn.eps, bn.momentum,
                        scale, shift, mean, invstd)
        if bias is not None:
            bn.running_mean.add_((1.0 - bn.momentum) * bias.detach())
        out = torch.empty_like(raw)
        ops.bn_apply(raw, scale, shift, out, relu, res=residual)
        ctx.conv, ctx.d, ctx.relu, ctx.has_res, ctx.has_bias = conv, d, relu, residual is not None, bias is not None
        ctx.save_for_backward(x, raw, out if (relu and residual is not None) else None, scale, shift, mean, invstd, gamma.detach(), weight.detach())
        return out

    @staticmethod
    def backward(ctx, dout):
        x, raw, out, scale, shift, mean, invstd, gamma, weight = ctx.saved_tensors
        conv, d = ctx.conv, ctx.d
        cout, cout_s, cin = conv.channels, raw.shape[-1], conv.in_channels
        dout = dout.contiguous()
        sums = torch.empty(2 * cout_s, dtype=torch.float32, device=raw.device)
        draw = torch.empty_like(raw)
        dres = None
        if ctx.has_res:
            dres = torch.empty_like(raw)
            ops.bn_backward(raw, dout, out if ctx.relu else None, mean, invstd, gamma, sums, draw, dz_out=dres)
        elif ctx.relu:
            ops.bn_backward(raw, dout, None, mean, invstd, gamma, sums, draw, relu_scale=scale, relu_shift=shift)
        else:
            ops.bn_backward(raw, dout, None, mean, invstd, gamma, sums, draw)
        dw = torch.empty_like(weight)
        ops.conv3d_wgrad(d, x, draw, dw, cout, cin)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if conv.strides != (1, 1, 1):
                table = ops.PackTable(x.device)
                plan = ops.DgradPlan(d, weight.permute(0, 2, 3, 4, 1).contiguous(), table)
                table.run()
                plan.run(draw, dx)
            else:
                dd = ops.dgrad_desc(d)
                ops.conv3d_fwd(dd, draw, ops.pack_conv_weight_dgrad(dd, weight), out=dx)
        dbias = torch.zeros(cout, dtype=torch.float32, device=raw.device) if ctx.has_bias else None
        return dx, dw, dbias, sums[:cout].clone(), sums[cout_s:cout_s + cout].clone(), dres, None, None, None


class Conv3D(torch.nn.Module):
    """nn.Conv3D(channels, kernel_size, strides, padding, use_bias=False) on K1.  The trunk never has a bias
    (R2Plus1.py:31,38,70 pass use_bias=False); the multi-task heads use gluon's default use_bias=True
    (multi_taskR3d.py:171,179), supported in eval mode by folding the bias into the epilogue's shift."""

    def __init__(self, in_channels, channels, kernel_size, strides=(1, 1, 1), padding=(0, 0, 0), use_bias=False):
        super().__init__()
        self.in_channels, self.channels = in_channels, channels
        self.kernel, self.strides, self.padding = tuple(kernel_size), tuple(strides), tuple(padding)
        self.weight = torch.nn.Parameter(_xavier((channels, in_channels) + self.kernel))
        self.bias = torch.nn.Parameter(torch.zeros(channels)) if use_bias else None

    def run(self, x, bn=None, relu=False, residual=None, training=False):
        """x: NDHWC bf16.  bn: BatchNorm module applied to the conv output (folded in eval mode)."""
        n, t, h, w, cs = x.shape
        cout_s = pad16(self.channels)
        flags = (FVT_CONV_RELU if relu else 0) | (FVT_CONV_RESIDUAL if residual is not None else 0)
        if bn is None or not training:
            # eval mode: packed bf16 weights and the folded (scale, shift) are cached until a parameter changes
            ver = (self.weight._version, self.weight.data_ptr(), None if self.bias is None else self.bias._version,
                   None if bn is None else (bn.gamma._version, bn.beta._version, bn.running_mean._version,
                                            bn.running_var._version, bn.gamma.data_ptr()), cs, flags, (n, t, h, w))
            hit = getattr(self, "_eval_cache", None)
            if hit is not None and hit[0] == ver:
                d, wp, scale, shift = hit[1]
                return ops.conv3d_fwd(d, x, wp, scale, shift, residual)
            d = ops.conv_desc(n, t, h, w, cs, cout_s, self.kernel, self.strides, self.padding, flags)
            wp = ops.pack_conv_weight(d, self.weight)
            scale = shift = None
            if bn is not None:
                scale, shift = bn.folded(cout_s)
            if self.bias is not None:                  # BN(conv + b) = scale * conv + (shift + scale * b)
                b = torch.zeros(cout_s, dtype=torch.float32, device=x.device)
                b[: self.channels] = self.bias.detach().float()
                if scale is None:
                    scale = torch.zeros(cout_s, dtype=torch.float32, device=x.device)
                    scale[: self.channels] = 1.0
                    shift = b
                else:
                    shift = shift + scale * b
            self._eval_cache = (ver, (d, wp, scale, shift))
            return ops.conv3d_fwd(d, x, wp, scale, shift, residual)
        # training mode: batch statistics (biased variance), running-stat update with the MXNet convention; one autograd node
        return _ConvBnFn.apply(x, self.weight, self.bias, bn.gamma, bn.beta, residual, self, bn, relu)


class BatchNorm(torch.nn.Module):
    """nn.BatchNorm() over the channel axis (R2Plus1.py:32,59,62,71): gamma=1, beta=0, running mean 0 / var 1."""

    def __init__(self, channels, eps=BN_EPS, momentum=BN_MOMENTUM):
        super().__init__()
        self.channels, self.eps, self.momentum = channels, eps, momentum
        self.gamma = torch.nn.Parameter(torch.ones(channels))
        self.beta = torch.nn.Parameter(torch.zeros(channels))
        self.register_buffer("running_mean", torch.zeros(channels))
        self.register_buffer("running_var", torch.ones(channels))

    def folded(self, c_store):
        scale = self.gamma.detach().float() / torch.sqrt(self.running_var.float() + self.eps)
        shift = self.beta.detach().float() - self.running_mean.float() * scale
        s = torch.zeros(c_store, dtype=torch.float32, device=scale.device)
        b = torch.zeros(c_store, dtype=torch.float32, device=scale.device)
        s[: self.channels] = scale
        b[: self.channels] = shift
        return s, b


class SpatialTemporalConv(torch.nn.Module):
    """The (2+1)D unit (R2Plus1.py:19-40): Conv3D(mid,(1,3,3),s=(1,sH,sW),p=(0,1,1)) -> BN -> ReLU ->
    Conv3D(out,(3,1,1),s=(sT,1,1),p=(1,0,0)).  As in the Gluon reference the spatial strides are stride[0], stride[1]
    and the temporal stride is stride[0] (identical to the symbol version for the strides ever passed)."""

    def __init__(self, in_filters, out_filter, stride, use_bias=False):
        super().__init__()
        mid = middle_filters(in_filters, out_filter)
        self.middle_filters = mid
        self.conv_middle = Conv3D(in_filters, mid, (1, 3, 3), (1, stride[0], stride[1]), (0, 1, 1), use_bias)
        self.bn_middle = BatchNorm(mid)
        self.conv = Conv3D(mid, out_filter, (3, 1, 1), (stride[0], 1, 1), (1, 0, 0), use_bias)
        self.out_filter = out_filter

    def run(self, x, bn=None, relu=False, residual=None, training=False):
        y = self.conv_middle.run(x, self.bn_middle, relu=True, training=training)
        return self.conv.run(y, bn, relu=relu, residual=residual, training=training)

    def forward(self, x):
        _require_cuda(x)
        return _from_ndhwc(self.run(_to_ndhwc(x), training=self.training and torch.is_grad_enabled()), self.out_filter)


def get_spatial_temporal_conv(in_filters, out_filter, stride, use_bias=False):
    return SpatialTemporalConv(in_filters, out_filter, stride, use_bias)


class R3DBlock(torch.nn.Module):
    """Residual block (R2Plus1.py:42-82): unit1 -> bn1 -> ReLU -> unit2 -> bn2; shortcut = x or BN(Conv3D 1x1x1 stride s);
    out = relu(y + shortcut)."""

    def __init__(self, input_filter, num_filter, comp_index=-1, downsampling=False, spation_batch_norm=True,
                 only_spatial_downsampling=False, use_bias=False):
        super().__init__()
        if comp_index == -1:
            print("error construct a residual block")          # the reference prints and carries on (:52-53)
        if downsampling:
            self.use_striding = [1, 2, 2] if only_spatial_downsampling else [2, 2, 2]
        else:
            self.use_striding = [1, 1, 1]
        self.spatial_temporal_conv1 = get_spatial_temporal_conv(input_filter, num_filter, self.use_striding, use_bias)
        self.bn1 = BatchNorm(num_filter)
        self.spatial_temporal_conv2 = get_spatial_temporal_conv(num_filter, num_filter, [1, 1, 1], use_bias)
        self.bn2 = BatchNorm(num_filter)
        self.num_filter, self.input_filter, self.downsampling = num_filter, input_filter, downsampling
        self.comp_index = comp_index
        if num_filter != input_filter or downsampling:
            self.branch_conv = Conv3D(input_filter, num_filter, (1, 1, 1), self.use_striding, (0, 0, 0), use_bias)
            self.branch_bn = BatchNorm(num_filter)

    def run(self, x, training=False):
        y = self.spatial_temporal_conv1.run(x, self.bn1, relu=True, training=training)
        if self.num_filter != self.input_filter or self.downsampling:
            sc = self.branch_conv.run(x, self.branch_bn, relu=False, training=training)
        else:
            sc = x
        return self.spatial_temporal_conv2.run(y, self.bn2, relu=True, residual=sc, training=training)

    def forward(self, x):
        _require_cuda(x)
        return _from_ndhwc(self.run(_to_ndhwc(x), training=self.training and torch.is_grad_enabled()), self.num_filter)


class _Stem(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.conv_middle = Conv3D(3, 45, (1, 7, 7), (1, 2, 2), (0, 3, 3))
        self.bn_middle = BatchNorm(45)
        self.conv = Conv3D(45, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0))
        self.bn = BatchNorm(64)

    def run(self, x_ncdhw, training=False):
        # the 3-channel input goes through the W-unfold transform so the 1x7x7 conv runs on the tensor-core kernel
        from ..engine import STEM_UNFOLD_CH, stem_equivalent_weight
        u = ops.stem_unfold(x_ncdhw.contiguous())
        n, t, h, wo, cu = u.shape
        d = ops.conv_desc(n, t, h, wo, STEM_UNFOLD_CH, pad16(45), (1, 7, 1), (1, 2, 1), (0, 3, 0), FVT_CONV_RELU)
        wp = ops.pack_conv_weight(d, stem_equivalent_weight(self.conv_middle.weight.detach()))
        if training:
            raise NotImplementedError("the stand-alone stem is forward/eval only; train through R2Plus2D")
        scale, shift = self.bn_middle.folded(pad16(45))
        y = ops.conv3d_fwd(d, u, wp, scale, shift)
        return self.conv.run(y, self.bn, relu=True, training=False)


class R2Plus1DSequential(torch.nn.Module):
    """get_R2plus1d (R2Plus1.py:295-347): the nn.Sequential variant — stem, residual blocks, AvgPool3D, Dense with a
    **sigmoid** activation (:346).  The reference passes no comp_index to some blocks (:335, it only prints a warning);
    the block structure is unaffected."""

    def __init__(self, num_class, model_depth, final_spatial_kernel, final_temporal_kernel):
        super().__init__()
        self.stem = _Stem()
        n1, n2, n3, n4 = BLOCK_CONFIG[model_depth]
        blocks, comp = [], 0
        for cin, cout, nb, down in ((64, 64, n1, False), (64, 128, n2, True), (128, 256, n3, True), (256, 512, n4, True)):
            for b in range(nb):
                blocks.append(R3DBlock(cin if b == 0 else cout, cout, comp_index=comp, downsampling=bool(down and b == 0)))
                comp += 1
        self.blocks = torch.nn.ModuleList(blocks)
        self.pool = (final_temporal_kernel, final_spatial_kernel, final_spatial_kernel)
        self.dense_weight = torch.nn.Parameter(torch.empty(num_class, 512).uniform_(-0.07, 0.07))
        self.dense_bias = torch.nn.Parameter(torch.zeros(num_class))
        self.num_class = num_class

    def forward(self, x):
        _require_cuda(x)
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError("get_R2plus1d() is forward/eval only here; train through R2Plus2D + Trainer")
        y = self.stem.run(x)
        for blk in self.blocks:
            y = blk.run(y)
        tp, hp, wp = y.shape[1] - self.pool[0] + 1, y.shape[2] - self.pool[1] + 1, y.shape[3] - self.pool[2] + 1
        if (tp, hp, wp) != (1, 1, 1):
            raise ValueError("AvgPool3D%s over a %s map: only a global pool is supported" % (self.pool, tuple(y.shape[1:4])))
        logits = ops.pool_fc_fwd(y, 512, self.dense_weight.detach().float().contiguous(), self.dense_bias.detach().float().contiguous())
        return torch.sigmoid(logits)


def get_R2plus1d(num_class=101, no_bias=0, model_depth=18, final_spatial_kernel=7, final_temporal_kernel=4):
    return R2Plus1DSequential(num_class, model_depth, final_spatial_kernel, final_temporal_kernel)


def _require_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("the R(2+1)D blocks run on sm_100a only: move the input to a CUDA device (no CPU fallback)")
