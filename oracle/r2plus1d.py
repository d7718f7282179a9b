"""CPU ORACLE (test infrastructure, not product code) for the R(2+1)D hot path of bruceyang2012/FastVideoTagging.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module.
The product path (fastvideotagging_b200/) never does.

What is restated here, with the reference lines each function follows:
  * mid-filter formula, block plan, parameter names/order  — model/R2Plus1.py:19-40, 84-90, 174-227; net.py:31-52
  * Conv3D / BatchNorm / Activation / AvgPool3D / Dense semantics of MXNet 1.x that the reference invokes
    (model/R2Plus1.py:27-38,59-62,67-71,99-114,168-171; net.py:40-51,79-98,122-133,164-169)
  * residual block and whole-network forward — model/R2Plus1.py:42-82, 93-172, 232-254; net.py:54-104, 110-170
  * Xavier initialisation — train_simple_r3d.py:81, train.py:88
  * SGD-momentum update and Trainer.step rescale — train_simple_r3d.py:95-97,124; train.py:69-73

PARITY STATUS: the arithmetic itself lives in third-party MXNet (`mxnet-cu90`, unpinned, requirements.txt:6) which is
not installable here and is not vendored under /root/reference, and the reference has no tests or recorded outputs for
it.  Parity for conv/BN values is therefore pinned only (a) structurally against r2plus1d_output/log.txt and
(b) against golden vectors produced by executing the reference's own model/R2Plus1.py + model/mlc_loss.py source on a
numpy/torch stand-in for the `mxnet` namespace (tests/golden/make_golden.py).  Numerically this is
"parity unpinned against real MXNet" — see DESIGN.md.

Two engines: `np_*` functions are the definition (plain numpy, fp32 or fp64); `Net` uses torch-CPU ops (oneDNN) as
a fast second engine for full-size clips and is itself checked against the numpy definition in tests/test_oracle.py.
"""
import math

import numpy as np

BLOCK_CONFIG = {10: (1, 1, 1, 1), 16: (2, 2, 2, 1), 18: (2, 2, 2, 2), 26: (2, 3, 4, 3), 34: (3, 4, 6, 3)}

# MXNet BatchNorm eps: gluon nn.BatchNorm() default (model/R2Plus1.py:32) vs symbol eps=1e-3 (net.py:44)
EPS_GLUON = 1e-5
EPS_SYMBOL = 1e-3
BN_MOMENTUM = 0.9


# ----------------------------------------------------------------------------------------------------------------
# structure
# ----------------------------------------------------------------------------------------------------------------
def mid_filters(cin, cout):
    """model/R2Plus1.py:22-24 / net.py:34-36 — true division then int()."""
    i = 3 * cin * cout * 3 * 3
    i /= cin * 3 * 3 + 3 * cout
    return int(i)


def conv_out(x, k, s, p):
    """MXNet Convolution / Pooling('valid') output extent."""
    return (x + 2 * p - k) // s + 1


def layer_plan(model_depth=34):
    """Ordered list of layers exactly as R2Plus2D.__init__ builds them (model/R2Plus1.py:93-172).

    Each entry: dict(kind='conv'|'bn', name=<symbol-API name, net.py / set_base_name / add_comp_count_index>, ...).
    Convs carry cin, cout, kernel, stride, pad; 'block' markers carry the block structure.
    """
    plan = []
    plan.append(dict(kind="conv", name="conv1_middle", cin=3, cout=45, kernel=(1, 7, 7), stride=(1, 2, 2), pad=(0, 3, 3)))
    plan.append(dict(kind="bn", name="conv1_middle_spatbn_relu", c=45, relu=True))
    plan.append(dict(kind="conv", name="conv1", cin=45, cout=64, kernel=(3, 1, 1), stride=(1, 1, 1), pad=(1, 0, 0)))
    plan.append(dict(kind="bn", name="conv1_spatbn_relu", c=64, relu=True))
    n2, n3, n4, n5 = BLOCK_CONFIG[model_depth]
    comp = 0
    stages = [(64, 64, n2, False), (64, 128, n3, True), (128, 256, n4, True), (256, 512, n5, True)]
    for cin, cout, nblocks, down in stages:
        for b in range(nblocks):
            first = b == 0
            plan.append(dict(kind="block", comp=comp, cin=cin if first else cout, cout=cout,
                             downsampling=bool(down and first)))
            comp += 1
    return plan


def block_layers(comp, cin, cout, downsampling):
    """Layers of one R3DBlock (model/R2Plus1.py:42-82; net.py:54-104) with the symbol-API names."""
    s = 2 if downsampling else 1
    layers = []
    mid1 = mid_filters(cin, cout)
    mid2 = mid_filters(cout, cout)
    layers.append(dict(kind="conv", name="comp_%d_conv_1_middle" % comp, cin=cin, cout=mid1, kernel=(1, 3, 3), stride=(1, s, s), pad=(0, 1, 1)))
    layers.append(dict(kind="bn", name="comp_%d_spatbn_1_middle" % comp, c=mid1, relu=True))
    layers.append(dict(kind="conv", name="comp_%d_conv_1" % comp, cin=mid1, cout=cout, kernel=(3, 1, 1), stride=(s, 1, 1), pad=(1, 0, 0)))
    layers.append(dict(kind="bn", name="comp_%d_spatbn_1" % comp, c=cout, relu=True))
    layers.append(dict(kind="conv", name="comp_%d_conv_2_middle" % comp, cin=cout, cout=mid2, kernel=(1, 3, 3), stride=(1, 1, 1), pad=(0, 1, 1)))
    layers.append(dict(kind="bn", name="comp_%d_spatbn_2_middle" % comp, c=mid2, relu=True))
    layers.append(dict(kind="conv", name="comp_%d_conv_2" % comp, cin=mid2, cout=cout, kernel=(3, 1, 1), stride=(1, 1, 1), pad=(1, 0, 0)))
    layers.append(dict(kind="bn", name="comp_%d_spatbn_2" % comp, c=cout, relu=False))
    if cin != cout or downsampling:
        layers.append(dict(kind="conv", name="shortcut_projection_%d" % comp, cin=cin, cout=cout, kernel=(1, 1, 1), stride=(s, s, s), pad=(0, 0, 0), shortcut=True))
        layers.append(dict(kind="bn", name="shortcut_projection_%d_spatbn" % comp, c=cout, relu=False, shortcut=True))
    return layers


def all_convs(model_depth=34):
    """Flat list of every conv (dict) in forward order."""
    out = []
    for item in layer_plan(model_depth):
        if item["kind"] == "conv":
            out.append(item)
        elif item["kind"] == "block":
            out.extend(l for l in block_layers(item["comp"], item["cin"], item["cout"], item["downsampling"]) if l["kind"] == "conv")
    return out


def param_names(model_depth=34, num_class=101):
    """(arg_names, aux_names) of the symbol built by create_r3d (net.py:110-170): the known-answer in
    r2plus1d_output/log.txt:38 is 211 args (incl. data + softmax_label) and 138 aux for depth 34."""
    args, aux = ["data"], []

    def add(layers):
        for l in layers:
            if l["kind"] == "conv":
                args.append(l["name"] + "_weight")
            else:
                args.extend([l["name"] + "_gamma", l["name"] + "_beta"])
                aux.extend([l["name"] + "_moving_mean", l["name"] + "_moving_var"])

    for item in layer_plan(model_depth):
        if item["kind"] == "block":
            add(block_layers(item["comp"], item["cin"], item["cout"], item["downsampling"]))
        else:
            add([item])
    args.extend(["final_fc_weight", "final_fc_bias", "softmax_label"])
    return args, aux


def conv_flops(model_depth, t, h, w):
    """2*M*N*K over unpadded conv dims per clip (forward), plus the list of per-conv (name, M, N, K)."""
    total, rows = 0, []
    shapes = {}

    def run(layers, shp_in, shortcut_in=None):
        nonlocal total
        cur = shp_in
        for l in layers:
            if l["kind"] != "conv":
                continue
            src = shortcut_in if l.get("shortcut") else cur
            to = conv_out(src[0], l["kernel"][0], l["stride"][0], l["pad"][0])
            ho = conv_out(src[1], l["kernel"][1], l["stride"][1], l["pad"][1])
            wo = conv_out(src[2], l["kernel"][2], l["stride"][2], l["pad"][2])
            m, n, k = to * ho * wo, l["cout"], l["cin"] * l["kernel"][0] * l["kernel"][1] * l["kernel"][2]
            total += 2 * m * n * k
            rows.append((l["name"], m, n, k))
            if not l.get("shortcut"):
                cur = (to, ho, wo)
        return cur

    cur = (t, h, w)
    for item in layer_plan(model_depth):
        if item["kind"] == "conv":
            cur = run([item], cur)
        elif item["kind"] == "block":
            cur = run(block_layers(item["comp"], item["cin"], item["cout"], item["downsampling"]), cur, cur)
    shapes["final"] = cur
    return total, rows, cur


# ----------------------------------------------------------------------------------------------------------------
# numpy definitions of the operators (MXNet 1.x semantics)
# ----------------------------------------------------------------------------------------------------------------
def _windows(xp, k, s, out):
    """Strided view xw[n, c, to, ho, wo, kt, kh, kw] over a padded NCDHW array."""
    n, c = xp.shape[:2]
    sn, sc, st, sh, sw = xp.strides
    shape = (n, c, out[0], out[1], out[2], k[0], k[1], k[2])
    strides = (sn, sc, st * s[0], sh * s[1], sw * s[2], st, sh, sw)
    return np.lib.stride_tricks.as_strided(xp, shape=shape, strides=strides, writeable=False)


def np_conv3d(x, w, stride, pad):
    """Cross-correlation, NCDHW x (O,I,kT,kH,kW), zero padding, no bias, out = floor((x+2p-k)/s)+1
    (nn.Conv3D / mx.sym.Convolution as used at model/R2Plus1.py:27-38, net.py:40-51)."""
    k = w.shape[2:]
    out = tuple(conv_out(x.shape[2 + i], k[i], stride[i], pad[i]) for i in range(3))
    xp = np.pad(x, ((0, 0), (0, 0), (pad[0],) * 2, (pad[1],) * 2, (pad[2],) * 2))
    xw = _windows(xp, k, stride, out)
    return np.einsum("nithwdef,oidef->nothw", xw, w, optimize=True)


def np_conv3d_wgrad(x, dy, kernel, stride, pad):
    xp = np.pad(x, ((0, 0), (0, 0), (pad[0],) * 2, (pad[1],) * 2, (pad[2],) * 2))
    xw = _windows(xp, kernel, stride, dy.shape[2:])
    return np.einsum("nothw,nithwdef->oidef", dy, xw, optimize=True)


def np_conv3d_dgrad(dy, w, x_shape, stride, pad):
    n, c, t, h, ww = x_shape
    k = w.shape[2:]
    dxp = np.zeros((n, c, t + 2 * pad[0], h + 2 * pad[1], ww + 2 * pad[2]), dtype=dy.dtype)
    to, ho, wo = dy.shape[2:]
    for dt in range(k[0]):
        for dh in range(k[1]):
            for dw in range(k[2]):
                contrib = np.einsum("nothw,oi->nithw", dy, w[:, :, dt, dh, dw], optimize=True)
                dxp[:, :, dt:dt + stride[0] * to:stride[0], dh:dh + stride[1] * ho:stride[1],
                    dw:dw + stride[2] * wo:stride[2]] += contrib
    return dxp[:, :, pad[0]:pad[0] + t, pad[1]:pad[1] + h, pad[2]:pad[2] + ww]


def np_batchnorm_train(x, gamma, beta, running_mean, running_var, eps=EPS_GLUON, momentum=BN_MOMENTUM):
    """MXNet BatchNorm(axis=1) in training mode: batch mean, BIASED batch variance;
    running = momentum*running + (1-momentum)*batch, using the biased variance (model/R2Plus1.py:32; net.py:44-45).
    Returns (y, new_running_mean, new_running_var, mean, inv_std)."""
    axes = (0, 2, 3, 4)
    mean = x.mean(axis=axes, dtype=np.float64)
    var = x.var(axis=axes, dtype=np.float64)          # biased (ddof=0)
    inv_std = 1.0 / np.sqrt(var + eps)
    sh = (1, -1, 1, 1, 1)
    y = (x - mean.reshape(sh)) * (inv_std * gamma).reshape(sh) + beta.reshape(sh)
    new_rm = momentum * running_mean + (1 - momentum) * mean
    new_rv = momentum * running_var + (1 - momentum) * var
    return y.astype(x.dtype), new_rm.astype(x.dtype), new_rv.astype(x.dtype), mean, inv_std


def np_batchnorm_eval(x, gamma, beta, running_mean, running_var, eps=EPS_GLUON):
    sh = (1, -1, 1, 1, 1)
    scale = gamma / np.sqrt(running_var + eps)
    return x * scale.reshape(sh) + (beta - running_mean * scale).reshape(sh)


def np_batchnorm_backward(x, dy, gamma, mean, inv_std):
    """Gradient of training-mode BatchNorm w.r.t. x, gamma, beta (batch statistics participate)."""
    axes = (0, 2, 3, 4)
    sh = (1, -1, 1, 1, 1)
    m = x.size // x.shape[1]
    xhat = (x - mean.reshape(sh)) * inv_std.reshape(sh)
    dbeta = dy.sum(axis=axes, dtype=np.float64)
    dgamma = (dy * xhat).sum(axis=axes, dtype=np.float64)
    dx = (gamma * inv_std).reshape(sh) * (dy - dbeta.reshape(sh) / m - xhat * dgamma.reshape(sh) / m)
    return dx.astype(x.dtype), dgamma, dbeta


def np_avgpool_dense(x, w, b, pool):
    """AvgPool3D(pool, stride 1, 'valid', divisor = full window) then Dense on the flattened result
    (model/R2Plus1.py:168-171,243-245; net.py:164-166).  For the configs used the pooled extent is 1x1x1."""
    out = tuple(conv_out(x.shape[2 + i], pool[i], 1, 0) for i in range(3))
    xw = _windows(np.ascontiguousarray(x), pool, (1, 1, 1), out)
    pooled = xw.mean(axis=(5, 6, 7))
    flat = pooled.reshape(x.shape[0], -1)
    return flat @ w.T + b, pooled


# ----------------------------------------------------------------------------------------------------------------
# initialisation / optimiser
# ----------------------------------------------------------------------------------------------------------------
def xavier_uniform(rng, shape, factor_type="avg", magnitude=3.0):
    """mx.init.Xavier(rnd_type='uniform'): fan_in = shape[1]*prod(shape[2:]), fan_out = shape[0]*prod(shape[2:]);
    U(-s, s) with s = sqrt(magnitude / factor).  train_simple_r3d.py:81 uses the defaults (avg, 3);
    train.py:88 uses factor_type='in', magnitude=2.34."""
    hw = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    fan_in, fan_out = shape[1] * hw, shape[0] * hw
    factor = {"avg": (fan_in + fan_out) / 2.0, "in": fan_in, "out": fan_out}[factor_type]
    s = math.sqrt(magnitude / factor)
    return rng.uniform(-s, s, size=shape).astype(np.float32)


def init_params(model_depth=34, num_class=101, seed=0, factor_type="avg", magnitude=3.0):
    """Random-init parameter dict keyed by the symbol-API names: Xavier-uniform conv/fc weights, gamma=1, beta=0,
    moving_mean=0, moving_var=1, fc bias 0 (gluon defaults; train_simple_r3d.py:81)."""
    rng = np.random.default_rng(seed)
    p = {}

    def add(layers):
        for l in layers:
            if l["kind"] == "conv":
                p[l["name"] + "_weight"] = xavier_uniform(rng, (l["cout"], l["cin"]) + tuple(l["kernel"]), factor_type, magnitude)
            else:
                c = l["c"]
                p[l["name"] + "_gamma"] = np.ones(c, np.float32)
                p[l["name"] + "_beta"] = np.zeros(c, np.float32)
                p[l["name"] + "_moving_mean"] = np.zeros(c, np.float32)
                p[l["name"] + "_moving_var"] = np.ones(c, np.float32)

    for item in layer_plan(model_depth):
        if item["kind"] == "block":
            add(block_layers(item["comp"], item["cin"], item["cout"], item["downsampling"]))
        else:
            add([item])
    p["final_fc_weight"] = xavier_uniform(rng, (num_class, 512), factor_type, magnitude)
    p["final_fc_bias"] = np.zeros(num_class, np.float32)
    return p


def randomize_bn(params, seed=1):
    """Give every BatchNorm non-trivial gamma/beta/moving stats so parity tests exercise the affine path."""
    rng = np.random.default_rng(seed)
    for k in list(params):
        if k.endswith("_gamma"):
            params[k] = rng.uniform(0.5, 1.5, params[k].shape).astype(np.float32)
        elif k.endswith("_beta"):
            params[k] = rng.normal(0, 0.1, params[k].shape).astype(np.float32)
        elif k.endswith("_moving_mean"):
            params[k] = rng.normal(0, 0.1, params[k].shape).astype(np.float32)
        elif k.endswith("_moving_var"):
            params[k] = rng.uniform(0.5, 1.5, params[k].shape).astype(np.float32)
    return params


def sgd_momentum_step(w, g, mom, lr, momentum=0.9, wd=0.0, rescale=1.0):
    """MXNet sgd_mom_update: g' = rescale*g + wd*w; mom = momentum*mom - lr*g'; w += mom
    (gluon.Trainer 'sgd', train_simple_r3d.py:95-97; step(batch) sets rescale = 1/batch, :124)."""
    gp = rescale * g + wd * w
    mom_new = momentum * mom - lr * gp
    return w + mom_new, mom_new


# ----------------------------------------------------------------------------------------------------------------
# whole network — numpy engine (definition; small inputs) and torch engine (fast; full-size clips)
# ----------------------------------------------------------------------------------------------------------------
def np_forward(params, x, model_depth=18, pool=(1, 7, 7), eps=EPS_GLUON, train=False, dtype=np.float32):
    """R2Plus2D.forward (model/R2Plus1.py:232-245) on numpy.  `train` selects batch statistics
    (inside autograd.record) vs moving statistics.  Returns (logits, features) where features is avg-pool output."""
    x = x.astype(dtype)
    P = {k: v.astype(dtype) for k, v in params.items()}

    def bn(name, h, relu):
        if train:
            y = np_batchnorm_train(h, P[name + "_gamma"], P[name + "_beta"], P[name + "_moving_mean"], P[name + "_moving_var"], eps)[0]
        else:
            y = np_batchnorm_eval(h, P[name + "_gamma"], P[name + "_beta"], P[name + "_moving_mean"], P[name + "_moving_var"], eps)
        return np.maximum(y, 0) if relu else y

    def run(layers, h):
        for l in layers:
            if l["kind"] == "conv":
                h = np_conv3d(h, P[l["name"] + "_weight"], l["stride"], l["pad"])
            else:
                h = bn(l["name"], h, l["relu"])
        return h

    for item in layer_plan(model_depth):
        if item["kind"] != "block":
            x = run([item], x)
        else:
            layers = block_layers(item["comp"], item["cin"], item["cout"], item["downsampling"])
            main = [l for l in layers if not l.get("shortcut")]
            short = [l for l in layers if l.get("shortcut")]
            y = run(main, x)
            sc = run(short, x) if short else x
            x = np.maximum(y + sc, 0)                        # nd.relu(y + x), R2Plus1.py:81
    logits, pooled = np_avgpool_dense(x, P["final_fc_weight"], P["final_fc_bias"], pool)
    return logits, pooled


class Net:
    """torch-CPU engine of the same network (fp32 or fp64), used where numpy einsum is too slow and for
    autograd-derived gradients.  BatchNorm follows MXNet (biased running variance, momentum multiplies the old
    value) — NOT torch.nn.BatchNorm3d's conventions."""

    def __init__(self, params, model_depth=18, pool=(1, 7, 7), eps=EPS_GLUON, dtype=None, bf16_storage=False):
        import torch
        self.torch = torch
        self.dtype = dtype or torch.float32
        self.depth = model_depth
        self.pool = pool
        self.eps = eps
        # bf16_storage: round activations/weights to bf16 at the points where the CUDA path stores bf16, so the
        # comparison isolates kernel errors from the (intended) storage precision.
        self.bf16 = bf16_storage
        self.p = {k: torch.tensor(np.asarray(v), dtype=self.dtype) for k, v in params.items()}
        self.running = {k: v.clone() for k, v in self.p.items() if "moving" in k}

    def _q(self, t):
        return t.to(self.torch.bfloat16).to(self.dtype) if self.bf16 else t

    def require_grad(self):
        for k, v in self.p.items():
            if "moving" not in k:
                v.requires_grad_(True)

    def _bn(self, name, h, relu, train):
        torch = self.torch
        g, b = self.p[name + "_gamma"], self.p[name + "_beta"]
        sh = (1, -1, 1, 1, 1)
        if train:
            mean = h.mean(dim=(0, 2, 3, 4))
            var = h.var(dim=(0, 2, 3, 4), unbiased=False)
            with torch.no_grad():
                self.running[name + "_moving_mean"] = BN_MOMENTUM * self.running[name + "_moving_mean"] + (1 - BN_MOMENTUM) * mean
                self.running[name + "_moving_var"] = BN_MOMENTUM * self.running[name + "_moving_var"] + (1 - BN_MOMENTUM) * var
        else:
            mean, var = self.running[name + "_moving_mean"], self.running[name + "_moving_var"]
        scale = g / torch.sqrt(var + self.eps)
        y = h * scale.reshape(sh) + (b - mean * scale).reshape(sh)
        return torch.relu(y) if relu else y

    def _run(self, layers, h, train, taps=None):
        F = self.torch.nn.functional
        for l in layers:
            if l["kind"] == "conv":
                h = F.conv3d(h, self._q(self.p[l["name"] + "_weight"]), stride=l["stride"], padding=l["pad"])
                if train:
                    h = self._q(h)              # training path stores the raw conv output in bf16
            else:
                h = self._bn(l["name"], h, l["relu"], train)
                if not l.get("last_in_block"):
                    h = self._q(h)
            if taps is not None:
                taps[l["name"]] = h
        return h

    def forward(self, x, train=False, taps=None):
        torch = self.torch
        x = self._q(torch.as_tensor(x, dtype=self.dtype))
        for item in layer_plan(self.depth):
            if item["kind"] != "block":
                x = self._run([item], x, train, taps)
            else:
                layers = block_layers(item["comp"], item["cin"], item["cout"], item["downsampling"])
                main = [dict(l) for l in layers if not l.get("shortcut")]
                main[-1]["last_in_block"] = True
                short = [dict(l) for l in layers if l.get("shortcut")]
                if short and train:
                    # training path: the projection's BatchNorm is applied inside the fused residual kernel in fp32
                    # (no bf16 round trip); the eval path stores the folded shortcut in bf16.
                    short[-1]["last_in_block"] = True
                y = self._run(main, x, train, taps)
                sc = self._run(short, x, train, taps) if short else x
                x = self._q(torch.relu(y + sc))
                if taps is not None:
                    taps["comp_%d_out" % item["comp"]] = x
        pooled = torch.nn.functional.avg_pool3d(x, self.pool, stride=1)
        flat = pooled.reshape(x.shape[0], -1)
        logits = flat @ self.p["final_fc_weight"].t() + self.p["final_fc_bias"]
        return logits, pooled
