"""Other heads on the same convolution kernels: eval-mode forward on K1, and training-mode forward +
backward (module.train() with gradients enabled) through the same BatchNorm / weight-gradient / data-gradient kernels the
trunk trains with (blocks._ConvBnFn per Conv3D -> BatchNorm -> ReLU group; the trunk's own backward through
R2Plus2D.conv5_features):

* `R2Plus2D_MT` — the multi-task scene/action network, reference model/multi_taskR3d.py:93-185 (ctor), :246-267
  (forward): the R(2+1)D trunk, then
      scene  = Dense(num_scenes)(Dropout(0.3)(flatten(ReLU(BN(Conv3D(256,(1,3,3),s(1,2,2))(x))))))
      action = Dense(num_actions)(AvgPool3D(ReLU(BN(Conv3D(512,(1,3,3),p(0,1,1))(x)))))
  Both head convs keep gluon's default bias.  Dropout is the identity in eval mode.  The flatten + Dense of the scene
  branch runs as ONE convolution whose window is the whole (T', H', W') map: the Dense weight reshaped
  (num_scenes, 256, T', H', W') is exactly the NCDHW flatten order `reshape(0, -1)` produces (:254).
* `Decision_thresh` — reference model/decision_model.py:4-14 (per-class threshold subtracted from the confidences).
* `ECOLite3DHead` — the 3D-ResNet18 tail of ECO-Lite (BASELINE configs[4]).  **The reference holds no code for it**
  (model/ECO.py:1-3 is two import lines; README.md:10 links the paper), so this follows the ECO paper's description:
  the 96-channel 28x28 feature maps of N_frames 2D-trunk outputs are stacked in time and go through 3x3x3 residual
  stages 128 -> 256 -> 512 (conv3_x..conv5_x of 3D-ResNet18), global average pool, Dense.  Parity is therefore
  unpinned against the reference; tests compare with a torch fp32 restatement of the same definition.
"""
import torch

from .. import ops
from ..ops import pad16
from .blocks import BatchNorm, Conv3D, to_ndhwc, _require_cuda, _xavier
from .R2Plus1 import R2Plus2D


class Decision_thresh(torch.nn.Module):
    def __init__(self, thresh_size=63):
        super().__init__()
        self.thresh = torch.nn.Parameter(torch.zeros(1, thresh_size))

    def forward(self, x):
        """x: (N, thresh_size) confidences; returns x - thresh (decision_model.py:11-14)."""
        return x - self.thresh


class _Dense(torch.nn.Module):
    def __init__(self, in_units, units):
        super().__init__()
        bound = (6.0 / (in_units + units)) ** 0.5
        self.weight = torch.nn.Parameter(torch.empty(units, in_units).uniform_(-bound, bound))
        self.bias = torch.nn.Parameter(torch.zeros(units))


class _PoolFcFn(torch.autograd.Function):
    """AvgPool3D (global) + Dense on an NDHWC bf16 map: fvt_pool_fc_fwd / fvt_pool_fc_bwd (the trunk's own head kernels)."""

    @staticmethod
    def forward(ctx, x, weight, bias, c_real):
        logits, pooled = ops.pool_fc_fwd(x, c_real, weight.detach().float().contiguous(), bias.detach().float().contiguous(), want_pooled=True)
        ctx.save_for_backward(pooled, weight.detach().float().contiguous())
        ctx.shape = tuple(x.shape)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        pooled, w = ctx.saved_tensors
        dw, db = torch.empty_like(w), torch.empty(w.shape[0], dtype=torch.float32, device=w.device)
        dx = torch.empty(ctx.shape, dtype=torch.bfloat16, device=w.device)
        ops.pool_fc_bwd(dlogits.float().contiguous(), pooled, w, dw, db, dx)
        return dx, dw, db, None


class _FlattenDenseFn(torch.autograd.Function):
    """flatten (NCDHW order, multi_taskR3d.py:254) + Dense as ONE convolution whose window is the whole (T', H', W') map:
    forward fvt_conv3d_fwd, backward the weight-gradient kernel and the data gradient (a convolution of dY, padded by k-1)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        n, ts, hs, ws, cs = x.shape
        units = weight.shape[0]
        up = pad16(units)
        d = ops.conv_desc(n, ts, hs, ws, cs, up, (ts, hs, ws), (1, 1, 1), (0, 0, 0), 0)
        w5 = weight.detach().float().reshape(units, -1, ts, hs, ws).contiguous()
        one = torch.zeros(up, dtype=torch.float32, device=x.device)
        one[:units] = 1.0
        b = torch.zeros_like(one)
        b[:units] = bias.detach().float()
        y = ops.conv3d_fwd(d, x, ops.pack_conv_weight(d, w5), one, b)
        ctx.save_for_backward(x, w5)
        ctx.d, ctx.units = d, units
        return y.reshape(n, -1)[:, :units].float()

    @staticmethod
    def backward(ctx, dy):
        x, w5 = ctx.saved_tensors
        d, units = ctx.d, ctx.units
        n, ts, hs, ws, cs = x.shape
        up = pad16(units)
        dyp = torch.zeros((n, 1, 1, 1, up), dtype=torch.bfloat16, device=x.device)
        dyp[:, 0, 0, 0, :units] = dy.to(torch.bfloat16)
        dw5 = torch.empty_like(w5)
        ops.conv3d_wgrad(d, x, dyp, dw5, units, w5.shape[1])
        dd = ops.conv_desc(n, 1, 1, 1, up, cs, (ts, hs, ws), (1, 1, 1), (ts - 1, hs - 1, ws - 1), 0)
        dx = ops.conv3d_fwd(dd, dyp, ops.pack_conv_weight_dgrad(dd, w5))
        return dx, dw5.reshape(units, -1), dy.float().sum(0)


class R2Plus2D_MT(torch.nn.Module):
    """forward(x) -> (scene, action), reference multi_taskR3d.py:246-267.  `trunk` is an R2Plus2D whose pooled/dense
    tail is unused; its parameters carry the same canonical names."""

    def __init__(self, num_scenes, num_actions, model_depth, final_spatial_kernel=7, final_temporal_kernel=2,
                 with_bias=False, scene_map=(1, 3, 3)):
        super().__init__()
        self.trunk = R2Plus2D(num_actions, model_depth, final_spatial_kernel, final_temporal_kernel, with_bias)
        self.scene_conv = Conv3D(512, 256, (1, 3, 3), (1, 2, 2), (0, 0, 0), use_bias=True)
        self.scene_bn = BatchNorm(256)
        self.scene_map = tuple(scene_map)                 # (T', H', W') of the scene conv output that the Dense flattens
        self.scene_output = _Dense(256 * scene_map[0] * scene_map[1] * scene_map[2], num_scenes)
        self.action_conv = Conv3D(512, 512, (1, 3, 3), (1, 1, 1), (0, 1, 1), use_bias=True)
        self.action_bn = BatchNorm(512)
        self.action_output = _Dense(512, num_actions)
        self.pool = (final_temporal_kernel, final_spatial_kernel, final_spatial_kernel)
        self.num_scenes, self.num_actions = num_scenes, num_actions

    dropout = 0.3          # nn.Dropout(0.3) on the scene branch (multi_taskR3d.py:173)

    def _forward_train(self, x):
        """multi_taskR3d.py:246-267 inside autograd.record(): batch-statistics BatchNorm in the trunk and both heads, Dropout
        active; backward() fills the trunk's flat gradient buffer and the head parameters' .grad."""
        self.trunk.train()
        feat = self.trunk.conv5_features(x)                                   # autograd node: the trunk's full backward
        s = self.scene_conv.run(feat, self.scene_bn, relu=True, training=True)
        if tuple(s.shape[1:4]) != self.scene_map:
            raise ValueError("scene feature map is %s but the Dense layer was sized for %s" % (tuple(s.shape[1:4]), self.scene_map))
        if self.dropout > 0:                                                  # MXNet Dropout: keep with p = 1 - rate, scale kept values by 1/p
            keep = (torch.rand(s.shape, device=s.device) >= self.dropout).to(s.dtype) / (1.0 - self.dropout)
            s = s * keep
        scene = _FlattenDenseFn.apply(s, self.scene_output.weight, self.scene_output.bias)
        a = self.action_conv.run(feat, self.action_bn, relu=True, training=True)
        tp, hp, wpool = a.shape[1] - self.pool[0] + 1, a.shape[2] - self.pool[1] + 1, a.shape[3] - self.pool[2] + 1
        if (tp, hp, wpool) != (1, 1, 1):
            raise ValueError("AvgPool3D%s over a %s map: only a global pool is supported" % (self.pool, tuple(a.shape[1:4])))
        action = _PoolFcFn.apply(a, self.action_output.weight, self.action_output.bias, 512)
        return scene, action

    def forward(self, x):
        _require_cuda(x)
        if self.training and torch.is_grad_enabled():
            return self._forward_train(x)
        feat = self.trunk.conv5_features(x)                                   # (N, T/8, H/16, W/16, 512) bf16
        # ---- scene branch
        s = self.scene_conv.run(feat, self.scene_bn, relu=True)
        n, ts, hs, ws, cs = s.shape
        if (ts, hs, ws) != self.scene_map:
            raise ValueError("scene feature map is %s but the Dense layer was sized for %s" % ((ts, hs, ws), self.scene_map))
        d = ops.conv_desc(n, ts, hs, ws, cs, pad16(self.num_scenes), (ts, hs, ws), (1, 1, 1), (0, 0, 0), 0)
        w5 = self.scene_output.weight.detach().float().reshape(self.num_scenes, 256, ts, hs, ws)
        wp = ops.pack_conv_weight(d, w5.contiguous())
        one = torch.zeros(pad16(self.num_scenes), dtype=torch.float32, device=x.device)
        one[: self.num_scenes] = 1.0
        b = torch.zeros_like(one)
        b[: self.num_scenes] = self.scene_output.bias.detach().float()
        scene = ops.conv3d_fwd(d, s, wp, one, b).reshape(n, -1)[:, : self.num_scenes].float()
        # ---- action branch
        a = self.action_conv.run(feat, self.action_bn, relu=True)
        tp, hp, wpool = a.shape[1] - self.pool[0] + 1, a.shape[2] - self.pool[1] + 1, a.shape[3] - self.pool[2] + 1
        if (tp, hp, wpool) != (1, 1, 1):
            raise ValueError("AvgPool3D%s over a %s map: only a global pool is supported" % (self.pool, tuple(a.shape[1:4])))
        action = ops.pool_fc_fwd(a, 512, self.action_output.weight.detach().float().contiguous(),
                                 self.action_output.bias.detach().float().contiguous())
        return scene, action


class _Basic3D(torch.nn.Module):
    """3x3x3 basic residual block: conv-BN-ReLU-conv-BN (+ projection conv-BN when the shape changes) -> add -> ReLU."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        st = (stride, stride, stride)
        self.conv1 = Conv3D(cin, cout, (3, 3, 3), st, (1, 1, 1))
        self.bn1 = BatchNorm(cout)
        self.conv2 = Conv3D(cout, cout, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        self.bn2 = BatchNorm(cout)
        self.project = cin != cout or stride != 1
        if self.project:
            self.down = Conv3D(cin, cout, (3, 3, 3), st, (1, 1, 1))
            self.down_bn = BatchNorm(cout)

    def run(self, x, training=False):
        y = self.conv1.run(x, self.bn1, relu=True, training=training)
        sc = self.down.run(x, self.down_bn, relu=False, training=training) if self.project else x
        return self.conv2.run(y, self.bn2, relu=True, residual=sc, training=training)


class ECOLite3DHead(torch.nn.Module):
    """(N, 96, T, 28, 28) stacked 2D-trunk features -> (N, num_class) logits.  See the module docstring: defined from the
    ECO paper, not from reference code."""

    STAGES = ((96, 128, 1), (128, 128, 1), (128, 256, 2), (256, 256, 1), (256, 512, 2), (512, 512, 1))

    def __init__(self, num_class=101):
        super().__init__()
        self.blocks = torch.nn.ModuleList([_Basic3D(ci, co, s) for ci, co, s in self.STAGES])
        self.dense = _Dense(512, num_class)
        self.num_class = num_class

    def forward(self, x):
        _require_cuda(x)
        y = x if (x.dim() == 5 and x.dtype == torch.bfloat16) else to_ndhwc(x)     # NDHWC bf16 is accepted as is
        if self.training and torch.is_grad_enabled():
            for blk in self.blocks:                                               # batch-statistics BatchNorm, autograd nodes
                y = blk.run(y, training=True)
            return _PoolFcFn.apply(y, self.dense.weight, self.dense.bias, 512)
        for blk in self.blocks:
            y = blk.run(y)
        return ops.pool_fc_fwd(y, 512, self.dense.weight.detach().float().contiguous(), self.dense.bias.detach().float().contiguous())

    @staticmethod
    def conv_gflop_per_clip(t=16, hw=28):
        """2*M*N*K over the 3x3x3 convolutions (incl. the two projection convs), per clip."""
        total, tt, ss = 0.0, t, hw
        for ci, co, s in ECOLite3DHead.STAGES:
            to, so = (tt + 2 - 3) // s + 1, (ss + 2 - 3) // s + 1
            m = to * so * so
            total += 2.0 * m * co * ci * 27 + 2.0 * m * co * co * 27
            if ci != co or s != 1:
                total += 2.0 * m * co * ci * 27
            tt, ss = to, so
        return total / 1e9
