"""Host-side mirror of reference model/R2Plus1.py: the R(2+1)D block builder API on top of the sm_100a kernels.

Same public names and argument meaning as the reference (`get_spatial_temporal_conv`, `R3DBlock`, `BLOCK_CONFIG`,
`R2Plus2D`, `get_R2plus1d`); MXNet is not installable on this platform, so the containers are torch.nn.Modules
(the "PyTorch extension" branch of the boundary) and tensors are torch CUDA tensors in the reference's NCDHW fp32
layout.  All arithmetic happens in libfvt_b200.so; there is no eager/CPU fallback.
"""
import math
import os
import pickle

import torch

from .. import engine, params_io
from ..engine import BLOCK_CONFIG, middle_filters  # noqa: F401  (BLOCK_CONFIG re-exported like the reference)

BN_EPS_GLUON = 1e-5      # nn.BatchNorm() default used throughout model/R2Plus1.py
BN_EPS_SYMBOL = 1e-3     # mx.sym.BatchNorm(eps=1e-3) in net.py
BN_MOMENTUM = 0.9


def xavier_uniform_(tensor, factor_type="avg", magnitude=3.0, generator=None):
    """mx.init.Xavier (uniform): U(-s, s), s = sqrt(magnitude / factor), fan = channels * prod(kernel)
    (train_simple_r3d.py:81 uses the defaults; train.py:88 uses factor_type='in', magnitude=2.34)."""
    shape = tensor.shape
    hw = 1
    for s in shape[2:]:
        hw *= s
    fan_in, fan_out = shape[1] * hw, shape[0] * hw
    factor = {"avg": (fan_in + fan_out) / 2.0, "in": fan_in, "out": fan_out}[factor_type]
    s = math.sqrt(magnitude / factor)
    with torch.no_grad():
        tensor.uniform_(-s, s, generator=generator)
    return tensor


class _TrainStep(torch.autograd.Function):
    """One training-mode forward/backward of the whole network on the C-ABI kernels.  Parameter gradients are written
    into the flat gradient buffer (the .grad views of the parameters), not returned through autograd."""

    @staticmethod
    def forward(ctx, x, net, anchor):
        plan = net._train_plan(x)
        ctx.plan = plan
        out = plan.forward(x, net._weights_signature())
        # Activations, BatchNorm statistics and the pooled features live in plan-owned buffers that the next same-shape
        # forward overwrites: remember which forward this graph node belongs to.
        ctx.generation = plan.fwd_generation
        for p in net._plans.values():   # a training forward moves the running statistics: folded-BN inference plans are stale
            p.stale = True
        return out

    @staticmethod
    def backward(ctx, dlogits):
        plan = ctx.plan
        if ctx.generation != plan.fwd_generation:
            raise RuntimeError(
                "backward() of a training forward whose saved activations were overwritten by a later forward of the same "
                "shape on this network (forward #%d, current #%d).  Run forward -> backward per batch shard (the one-process-"
                "per-GPU form of train_simple_r3d.py:116-123), or use different networks per shard." % (ctx.generation, plan.fwd_generation))
        plan.backward(dlogits.float())       # every gradient slot is overwritten (MXNet grad_req='write')
        return None, None, None


class _TrunkStep(torch.autograd.Function):
    """Training-mode forward/backward of the trunk alone: x -> conv5_x output (N, T/8, H/16, W/16, 512) bf16."""

    @staticmethod
    def forward(ctx, x, net, anchor):
        plan = net._train_plan(x)
        ctx.plan = plan
        fmap = plan.forward_map(x, net._weights_signature())
        ctx.generation = plan.fwd_generation
        for p in net._plans.values():
            p.stale = True
        return fmap.clone()

    @staticmethod
    def backward(ctx, dmap):
        plan = ctx.plan
        if ctx.generation != plan.fwd_generation:
            raise RuntimeError("backward() of a trunk forward whose saved activations were overwritten by a later forward")
        plan.backward_map(dmap.to(torch.bfloat16))
        return None, None, None


class R2Plus2D(torch.nn.Module):
    """R(2+1)D-{10,16,18,26,34} (reference model/R2Plus1.py:93-254).

    forward(x): x is (N, 3, T, H, W) fp32 on a CUDA device; returns (N, num_class) fp32 logits (no activation,
    :171).  precision='fp32' runs eval-mode forwards on the plain-fp32 kernels (the reference's own number type; slow,
    for checking against fp32 references at 1e-4); training always uses the bf16-storage tcgen05 path.  `final_temporal_kernel` / `final_spatial_kernel` are the AvgPool3D window and must cover the whole
    conv5 map (T/8, H/16), as every reference caller arranges (train.py:39-40, R2Plus1.py:370).
    """

    def __init__(self, num_class, model_depth, final_spatial_kernel=7, final_temporal_kernel=2, with_bias=False,
                 bn_eps=BN_EPS_GLUON, precision="bf16"):
        super().__init__()
        if with_bias:
            raise NotImplementedError("with_bias=True is never used by the reference callers; conv bias is not built")
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' (tcgen05 kernels, the fast path) or 'fp32' (CUDA-core inference path)")
        self.precision = precision
        self.num_class = num_class
        self.model_depth = model_depth
        self.pool = (final_temporal_kernel, final_spatial_kernel, final_spatial_kernel)
        self.bn_eps = bn_eps
        pshapes, ashapes = engine.parameter_shapes(model_depth, num_class)
        self._param_names = list(pshapes)
        self._aux_names = list(ashapes)
        for name, shape in pshapes.items():
            if name.endswith("_gamma"):
                init = torch.ones(shape)
            elif name.endswith("_beta") or name.endswith("_bias"):
                init = torch.zeros(shape)
            else:
                init = torch.zeros(shape)
            self.register_parameter(name, torch.nn.Parameter(init))
        for name, shape in ashapes.items():
            self.register_buffer(name, torch.ones(shape) if name.endswith("_var") else torch.zeros(shape))
        # name lists kept for load_from_sym_params compatibility (model/R2Plus1.py:174-227)
        self.base_name = self.set_base_name()
        self.dense0_name = ["final_fc_weight", "final_fc_bias"]
        self._plans = {}
        self._train_plans = {}
        self._weights_version = 0
        self._flat = None
        self._trainer = None
        self._grad_anchor = None
        self.bn_momentum = BN_MOMENTUM
        self._input_norm = None

    # ------------------------------------------------------------------ reference helpers
    @staticmethod
    def set_base_name():
        return ["conv1_middle_weight",
                "conv1_middle_spatbn_relu_gamma", "conv1_middle_spatbn_relu_beta",
                "conv1_middle_spatbn_relu_moving_mean", "conv1_middle_spatbn_relu_moving_var",
                "conv1_weight",
                "conv1_spatbn_relu_gamma", "conv1_spatbn_relu_beta",
                "conv1_spatbn_relu_moving_mean", "conv1_spatbn_relu_moving_var"]

    @staticmethod
    def add_comp_count_index(change_channels=False, downsampling=False, comp_index=-1, prefix=None):
        names = []
        for conv_idx in (1, 2):
            names += ["comp_%d_conv_%d_middle_weight" % (comp_index, conv_idx)]
            names += ["comp_%d_spatbn_%d_middle_%s" % (comp_index, conv_idx, s) for s in ("gamma", "beta", "moving_mean", "moving_var")]
            names += ["comp_%d_conv_%d_weight" % (comp_index, conv_idx)]
            names += ["comp_%d_spatbn_%d_%s" % (comp_index, conv_idx, s) for s in ("gamma", "beta", "moving_mean", "moving_var")]
        if change_channels or downsampling:
            names += ["shortcut_projection_%d_weight" % comp_index]
            names += ["shortcut_projection_%d_spatbn_%s" % (comp_index, s) for s in ("gamma", "beta", "moving_mean", "moving_var")]
        return names

    # ------------------------------------------------------------------ gluon-style parameter protocol
    def initialize(self, init=None, ctx=None, seed=None, factor_type="avg", magnitude=3.0):
        """net.initialize(init.Xavier(), ctx) (train_simple_r3d.py:81): Xavier-uniform weights, gamma=1, beta=0,
        running mean 0 / var 1, dense bias 0."""
        device = ctx if ctx is not None else torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if isinstance(device, (list, tuple)):
            device = device[0]
        if device is not None:
            self.to(device)
        gen = None
        if seed is not None:
            gen = torch.Generator(device=self.final_fc_weight.device)
            gen.manual_seed(seed)
        for name in self._param_names:
            p = getattr(self, name)
            if name.endswith("_weight"):
                xavier_uniform_(p.data, factor_type, magnitude, gen)
            elif name.endswith("_gamma"):
                p.data.fill_(1.0)
            else:
                p.data.zero_()
        for name in self._aux_names:
            b = getattr(self, name)
            b.fill_(1.0) if name.endswith("_var") else b.zero_()
        self.invalidate()
        return self

    def collect_params(self):
        return {n: getattr(self, n) for n in self._param_names + self._aux_names}

    def load_param_dict(self, arrays, with_dense=True, strict=True):
        """Load {symbol-API name: array-like}.  `arg:`/`aux:` prefixes of MXNet checkpoints are stripped
        (model/R2Plus1.py:262-265)."""
        clean = {k.split(":")[-1]: v for k, v in arrays.items()}
        missing = []
        with torch.no_grad():
            for name in self._param_names + self._aux_names:
                if name.startswith("final_fc") and not with_dense:
                    continue
                if name not in clean:
                    missing.append(name)
                    continue
                dst = getattr(self, name)
                src = torch.as_tensor(clean[name]).to(dst.device, dst.dtype)
                if tuple(src.shape) != tuple(dst.shape):
                    raise ValueError("shape mismatch for %s: %s vs %s" % (name, tuple(src.shape), tuple(dst.shape)))
                dst.copy_(src)
        if strict and missing:
            raise KeyError("missing parameters: %s" % missing[:5])
        self.invalidate()
        return missing

    def save_parameters(self, filename):
        """gluon `save_parameters` (train_simple_r3d.py:137): an MXNet NDArray-dict file keyed by the symbol-API names."""
        params_io.nd_save(filename, {k: v.detach().float().cpu().numpy() for k, v in self.collect_params().items()})

    def load_parameters(self, filename, ctx=None, allow_missing=False):
        """gluon `load_parameters` / deprecated `load_params` (train_simple_r3d.py:92,225).  Reads an MXNet NDArray file
        (plain, or `arg:`/`aux:`-prefixed as written by `do_checkpoint`) or a torch-saved dict."""
        self.load_param_dict(params_io.load_any(filename), strict=not allow_missing)
        if ctx is not None:
            self.to(ctx[0] if isinstance(ctx, (list, tuple)) else ctx)

    load_params = load_parameters   # deprecated gluon alias used at train_simple_r3d.py:92

    def load_from_sym_params(self, f, ctx=None, with_dense=False):
        """Reference model/R2Plus1.py:256-279: load a symbol-API checkpoint (`prefix-0001.params`), skipping the dense
        layer unless asked."""
        if not os.path.exists(f):
            print("parameter file is not exist", f)
            return
        self.load_param_dict(params_io.load_any(f), with_dense=with_dense, strict=True)

    def load_from_caffe2_pickle(self, f):
        """Reference model/R2Plus1.py:285-290 opens the Caffe2 R(2+1)D pickle and stops (an unfinished stub); the mapping
        it was heading for is the one `utils.load_from_caffe2_pkl` applies (utils.py:21-33), used here.  The Kinetics
        head (`last_out_L400_*`) has no counterpart and the dense layer keeps its initialisation, exactly the
        "not loaded / not used" outcome the reference logs (r2plus1d_output/log.txt:41-47)."""
        if not os.path.exists(f):
            print("parameter file is not exist", f)
            return None
        with open(f, "rb") as fh:
            blobs = pickle.load(fh, encoding="latin1")["blobs"]
        args_loaded, auxs_loaded = params_io.caffe2_blobs_to_params(blobs)
        merged = dict(args_loaded)
        merged.update(auxs_loaded)
        own = set(self._param_names + self._aux_names)
        missing = self.load_param_dict({k: v for k, v in merged.items() if k in own}, with_dense=False, strict=False)
        return {"not_loaded": [m for m in missing], "not_used": sorted(k for k in merged if k not in own)}

    def invalidate(self):
        """Call after changing parameters outside of this module's own optimiser hooks."""
        self._weights_version += 1
        self._plans.clear()

    def _weights_changed(self):
        """Optimiser hook: the training plan re-packs its operand copies lazily; inference plans are marked stale and
        re-pack / re-fold in place at their next use (their buffers and captured graphs are kept)."""
        self._weights_version += 1
        for plan in self._plans.values():
            plan.stale = True

    def _weights_signature(self):
        """Changes whenever any parameter or running statistic changes: the module's own counter (bumped by Trainer.step,
        load_param_dict, initialize — they write through raw pointers) plus torch's in-place version counters, so that a
        torch.optim step, load_state_dict or any other in-place edit also invalidates packed weights and folded plans."""
        if self._flat is not None:
            v = self._flat.w._version          # the parameters are views of one buffer: one shared counter
        else:
            v = sum(getattr(self, n)._version for n in self._param_names)
        return (self._weights_version, v, sum(getattr(self, n)._version for n in self._aux_names))

    def zero_grad(self, set_to_none=False):
        """The .grad tensors are views of the flat gradient buffer (what the NCCL buckets and the fused SGD launch read):
        they are zeroed in place, never detached."""
        if self._flat is not None:
            self._flat.g.zero_()
        else:
            super().zero_grad(set_to_none=set_to_none)

    def _attach_trainer(self, trainer):
        self._trainer = trainer
        for plan in self._train_plans.values():
            plan.set_hooks(trainer.on_grads_ready, trainer.allreduce_grads)

    def _ensure_flat(self, device):
        """Move every trainable tensor into one flat fp32 buffer (engine.FlatParams); the nn.Parameters become views
        of it and their .grad views of the flat gradient buffer."""
        if self._flat is not None and self._flat.w.device == device:
            return self._flat
        if self._flat is not None:
            raise RuntimeError("this network already trains on %s; one process per GPU: build one network per device "
                               "(torch.distributed), do not move a training network between devices" % self._flat.w.device)
        flat = engine.FlatParams(self.model_depth, self.num_class, device)
        with torch.no_grad():
            for name in self._param_names:
                p = getattr(self, name)
                flat.view(flat.w, name).copy_(p.data.to(device))
                p.data = flat.view(flat.w, name)
                p.grad = flat.view(flat.g, name)
        for name in self._aux_names:
            b = getattr(self, name)
            if b.device != device:
                setattr(self, name, b.to(device))
        self._flat = flat
        self._grad_anchor = torch.zeros(1, device=device, requires_grad=True)
        self._train_plans.clear()
        return flat

    @staticmethod
    def _clip_dims(x):
        """(N, T, H, W) of a clip batch in either accepted form: the reference's (N, 3, T, H, W) fp32, or decoded uint8
        frames (N, T, H, W, 3)."""
        if x.dtype == torch.uint8:
            if x.dim() != 5 or x.shape[-1] != 3:
                raise ValueError("uint8 clips must be (N, T, H, W, 3) decoded frames")
            return x.shape[0], x.shape[1], x.shape[2], x.shape[3]
        if x.dim() != 5 or x.shape[1] != 3:
            raise ValueError("clips must be (N, 3, T, H, W)")
        return x.shape[0], x.shape[2], x.shape[3], x.shape[4]

    def set_input_normalization(self, mean, std, scale=1.0 / 255.0, std_eps=0.0):
        """Constants for uint8 clip batches (N, T, H, W, 3): value = (v*scale - mean[c]) / (std[c] + std_eps), applied
        inside the stem's input transform (one pass from bytes to the stem operand).  ImageNet statistics with
        scale = 1/255 is data/ucf101.py:124-128; the per-batch statistics of videos_reader.py:93-97 are
        evaluate.batch_statistics(clips) with scale = 1, std_eps = 1e-3."""
        self._input_norm = (float(scale), tuple(float(v) for v in mean), tuple(1.0 / (float(v) + std_eps) for v in std))
        for plan in list(self._plans.values()) + list(self._train_plans.values()):
            plan.input_norm = self._input_norm

    def _train_plan(self, x):
        flat = self._ensure_flat(x.device)
        n, t, h, w = self._clip_dims(x)
        key = ((n, t, h, w), x.device.index)
        plan = self._train_plans.get(key)
        if plan is None:
            aux = {k: getattr(self, k) for k in self._aux_names}
            plan = engine.TrainPlan(flat, aux, self.model_depth, self.num_class, self.pool, self.bn_eps, n, t, h, w,
                                    x.device, momentum=self.bn_momentum)
            if self._trainer is not None:
                plan.set_hooks(self._trainer.on_grads_ready, self._trainer.allreduce_grads)
            plan.input_norm = self._input_norm
            self._train_plans[key] = plan
        return plan

    # ------------------------------------------------------------------ forward
    def _inference_plan(self, x):
        sig = self._weights_signature()
        if sig != getattr(self, "_plans_sig", None):       # parameters changed (also in place, behind our back): re-fold / re-pack
            for p in self._plans.values():
                p.stale = True
            self._plans_sig = sig
        n, t, h, w = self._clip_dims(x)
        key = ((n, t, h, w), x.device.index)
        plan = self._plans.get(key)
        params = {k: getattr(self, k) for k in self._param_names}
        aux = {k: getattr(self, k) for k in self._aux_names}
        if plan is not None and getattr(plan, "stale", False):
            if not (hasattr(plan, "refresh") and plan.refresh(params, aux)):
                plan = None                                # parameter storage moved: rebuild
        if plan is None:
            cls = engine.InferencePlanF32 if self.precision == "fp32" else engine.InferencePlan
            plan = cls(params, aux, self.model_depth, self.num_class, self.pool, self.bn_eps, n, t, h, w, x.device)
            plan.input_norm = self._input_norm
            self._plans[key] = plan
        return plan

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("R2Plus2D runs on sm_100a only: move the clip batch to a CUDA device (no CPU fallback)")
        if self.training and torch.is_grad_enabled():
            # inside autograd.record() in the reference: batch-statistics BatchNorm, gradients on backward()
            return _TrainStep.apply(x, self, self._ensure_anchor(x.device))
        return self._inference_plan(x).forward(x)

    def _ensure_anchor(self, device):
        self._ensure_flat(device)
        return self._grad_anchor

    def conv5_features(self, x):
        """conv5_x output in the kernels' layout, (N, T/8, H/16, W/16, 512) bf16 — the input of the multi-task heads
        (reference multi_taskR3d.py:246-251 runs the same trunk).  In training mode with gradients enabled the result is
        an autograd node: its backward runs the trunk's full backward pass (batch-statistics BatchNorm, all weight
        gradients into the flat gradient buffer)."""
        if not x.is_cuda:
            raise RuntimeError("R2Plus2D runs on sm_100a only: move the clip batch to a CUDA device (no CPU fallback)")
        if self.training and torch.is_grad_enabled():
            return _TrunkStep.apply(x, self, self._ensure_anchor(x.device))
        return self._inference_plan(x).forward(x, want_map=True)

    def extract_features(self, x):
        """Reference :247-254 — the AvgPool3D output, shape (N, 512, 1, 1, 1)."""
        _, pooled = self._inference_plan(x).forward(x, want_features=True)
        return pooled.reshape(x.shape[0], 512, 1, 1, 1)
