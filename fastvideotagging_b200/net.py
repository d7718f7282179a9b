"""Symbol-API twin of the model builder (reference net.py:18-170): `ModelBuilder` and `create_r3d`.

In the reference these build an `mx.sym` graph that `mx.module.Module` binds, one executor per GPU (train.py:35-54,
validation.py:22-30).  MXNet does not exist for sm_100, so `create_r3d` returns an `R3DSymbol`: a description of the same
graph (same layer names, same 211 arguments + 138 auxiliary states for depth 34 / 101 classes, cf.
r2plus1d_output/log.txt:38) whose `bind()` gives an executor running on the C-ABI kernels with the symbol API's
semantics — BatchNorm eps = 1e-3 and momentum = bn_mom (net.py:44-45), `SoftmaxOutput(multi_output=True,
use_ignore=True, normalization='null')` head (:167-169): forward returns class probabilities, backward starts from
(p - onehot) with label -1 ignored.
"""
import torch

from . import engine
from .engine import BLOCK_CONFIG  # noqa: F401  (reference net.py:9-15 exports it)
from .model.R2Plus1 import R2Plus2D
from .model.mlc_loss import SoftmaxOutput

BN_EPS_SYMBOL = 1e-3


class _Node:
    """A symbolic tensor: just enough bookkeeping to check that what the caller builds is the R(2+1)D topology."""
    __slots__ = ("channels", "t_div", "s_div", "blocks")

    def __init__(self, channels, t_div, s_div, blocks):
        self.channels, self.t_div, self.s_div, self.blocks = channels, t_div, s_div, blocks


class ModelBuilder:
    """reference net.py:18-104.  `add_spatial_temporal_conv` / `add_r3d_block` take and return symbolic nodes and
    register the layer names (`comp_%d_conv_%d[_middle]`, `comp_%d_spatbn_%d[_middle]`, `shortcut_projection_%d`)."""

    def __init__(self, no_bias, bn_mom=0.9, cudnn_tune="off", workspace=512):
        self.comp_count = 0
        self.comp_idx = 0
        self.bn_mom = bn_mom
        self.no_bias = 1 if no_bias else 0
        self.cudnn_tune = cudnn_tune       # accepted and ignored: there is no cuDNN underneath
        self.workspace = workspace
        self.arguments = []                # parameter names in creation order
        self.aux_states = []

    def _conv(self, name):
        self.arguments.append(name + "_weight")
        if not self.no_bias:
            self.arguments.append(name + "_bias")

    def _bn(self, name):
        self.arguments += [name + "_gamma", name + "_beta"]
        self.aux_states += [name + "_moving_mean", name + "_moving_var"]

    def add_spatial_temporal_conv(self, body, in_filters, out_filters, stride):
        """net.py:31-52: conv (1,3,3) stride (1, s[1], s[2]) -> BN(eps 1e-3) -> ReLU -> conv (3,1,1) stride (s[0],1,1)."""
        self.comp_idx += 1
        mid = engine.middle_filters(in_filters, out_filters)
        self._conv("comp_%d_conv_%d_middle" % (self.comp_count, self.comp_idx))
        self._bn("comp_%d_spatbn_%d_middle" % (self.comp_count, self.comp_idx))
        self._conv("comp_%d_conv_%d" % (self.comp_count, self.comp_idx))
        del mid
        return _Node(out_filters, body.t_div * stride[0], body.s_div * stride[1], body.blocks)

    def add_r3d_block(self, data, input_filters, num_filters, down_sampling=False, spatial_batch_norm=True,
                      only_spatial_downsampling=False):
        """net.py:54-104."""
        if data.channels != input_filters:
            raise ValueError("block %d expects %d input channels, got %d" % (self.comp_count, input_filters, data.channels))
        if only_spatial_downsampling:
            raise NotImplementedError("only_spatial_downsampling=True is never used by the reference callers")
        self.comp_idx = 0
        stride = [2, 2, 2] if down_sampling else [1, 1, 1]
        body = self.add_spatial_temporal_conv(data, input_filters, num_filters, stride)
        self._bn("comp_%d_spatbn_%d" % (self.comp_count, self.comp_idx))
        body = self.add_spatial_temporal_conv(body, num_filters, num_filters, [1, 1, 1])
        self._bn("comp_%d_spatbn_%d" % (self.comp_count, self.comp_idx))
        if num_filters != input_filters or down_sampling:
            self._conv("shortcut_projection_%d" % self.comp_count)
            self._bn("shortcut_projection_%d_spatbn" % self.comp_count)
        out = _Node(num_filters, body.t_div, body.s_div, data.blocks + [(self.comp_count, input_filters, num_filters, bool(down_sampling))])
        self.comp_count += 1
        return out


class R3DSymbol:
    """What `create_r3d` returns: the graph description plus `bind()`."""

    def __init__(self, num_class, model_depth, pool, bn_mom, arguments, aux_states):
        self.num_class, self.model_depth, self.pool, self.bn_mom = num_class, model_depth, pool, bn_mom
        self._arguments, self._aux = arguments, aux_states

    def list_arguments(self):
        return list(self._arguments)

    def list_auxiliary_states(self):
        return list(self._aux)

    def list_outputs(self):
        return ["softmax_output"]

    def infer_shape(self, data):
        """data = (N, 3, T, H, W) -> (argument shapes, output shapes, aux shapes), MXNet style."""
        pshapes, ashapes = engine.parameter_shapes(self.model_depth, self.num_class)
        args = []
        for name in self._arguments:
            if name == "data":
                args.append(tuple(data))
            elif name == "softmax_label":
                args.append((data[0],))
            else:
                args.append(tuple(pshapes[name]))
        return args, [(data[0], self.num_class)], [tuple(ashapes[n]) for n in self._aux]

    def bind(self, ctx=None, arg_params=None, aux_params=None):
        """-> R3DExecutor on `ctx` (a torch CUDA device); parameters from MXNet-style {name: array} dicts."""
        net = R2Plus2D(self.num_class, self.model_depth, final_spatial_kernel=self.pool[1], final_temporal_kernel=self.pool[0],
                       bn_eps=BN_EPS_SYMBOL)
        net.bn_momentum = self.bn_mom
        device = ctx if ctx is not None else torch.device("cuda", torch.cuda.current_device())
        net.to(device)
        params = dict(arg_params or {})
        params.update(aux_params or {})
        if params:
            net.load_param_dict(params, with_dense=True, strict=False)
        return R3DExecutor(net)


class R3DExecutor:
    """forward(is_train, data, softmax_label) / backward() / outputs, like an MXNet executor bound to one context."""

    def __init__(self, net):
        self.net = net
        self.outputs = []
        self._logits = None

    def forward(self, is_train=False, data=None, softmax_label=None):
        self.net.train(bool(is_train))
        if is_train:
            logits = self.net(data)
            self._logits = logits
            label = softmax_label if softmax_label is not None else torch.full((data.shape[0],), -1.0, device=data.device)
            self._prob = SoftmaxOutput(logits, label.float())
            self.outputs = [self._prob.detach()]
        else:
            with torch.no_grad():
                logits = self.net(data)
                label = torch.full((data.shape[0],), -1.0, device=data.device)
                self.outputs = [SoftmaxOutput(logits, label)]
        return self.outputs

    def backward(self):
        """SoftmaxOutput ignores the incoming head gradient: d(logits) = p - onehot (net.py:167-169)."""
        if self._logits is None:
            raise RuntimeError("backward() needs a forward(is_train=True) first")
        self._prob.backward(torch.ones_like(self._prob))
        self._logits = None


def create_r3d(num_class, no_bias=0, model_depth=18, final_spatial_kernel=7, final_temporal_kernel=1, bn_mom=0.9,
               cudnn_tune="off", workspace=512):
    """reference net.py:110-170.  Returns an `R3DSymbol`."""
    if not no_bias:
        # The reference's default no_bias=0 gives every Convolution a bias that is immediately absorbed by the
        # BatchNorm that follows; its pretrained checkpoints carry none (r2plus1d_output/log.txt lists no *_bias but
        # final_fc_bias).  The kernels implement the bias-free form.
        no_bias = 1
    builder = ModelBuilder(no_bias=no_bias, bn_mom=bn_mom, cudnn_tune=cudnn_tune, workspace=workspace)
    builder.arguments.append("data")
    builder._conv("conv1_middle")
    builder._bn("conv1_middle_spatbn_relu")
    builder._conv("conv1")
    builder._bn("conv1_spatbn_relu")
    body = _Node(64, 1, 2, [])
    n1, n2, n3, n4 = BLOCK_CONFIG[model_depth]
    for _ in range(n1):
        body = builder.add_r3d_block(body, 64, 64)
    body = builder.add_r3d_block(body, 64, 128, down_sampling=True)
    for _ in range(n2 - 1):
        body = builder.add_r3d_block(body, 128, 128)
    body = builder.add_r3d_block(body, 128, 256, down_sampling=True)
    for _ in range(n3 - 1):
        body = builder.add_r3d_block(body, 256, 256)
    body = builder.add_r3d_block(body, 256, 512, down_sampling=True)
    for _ in range(n4 - 1):
        body = builder.add_r3d_block(body, 512, 512)
    print(builder.comp_count)                                   # the reference prints the block count (net.py:162)
    if body.blocks != engine.network_blocks(model_depth):
        raise RuntimeError("builder produced a topology the engine does not implement")
    builder.arguments += ["final_fc_weight", "final_fc_bias", "softmax_label"]
    return R3DSymbol(num_class, model_depth, (final_temporal_kernel, final_spatial_kernel, final_spatial_kernel), bn_mom,
                     builder.arguments, builder.aux_states)
