"""ORACLE (test infrastructure only — never imported by the product path): torch-CPU fp32 restatement of the heads
next to the R(2+1)D trunk.

* `multitask_heads` follows reference model/multi_taskR3d.py:169-185 (layers) and :246-267 (forward):
  scene  = Dense(flatten_NCDHW(ReLU(BN(Conv3D(256,(1,3,3),s(1,2,2),bias)(x)))))   [Dropout = identity in eval]
  action = Dense(AvgPool3D(ReLU(BN(Conv3D(512,(1,3,3),p(0,1,1),bias)(x)))))
  BatchNorm in eval mode with the gluon default eps 1e-5 (MXNet third-party semantics, as oracle/r2plus1d.py).
* `eco_lite_3d_head` has NO reference code to follow (reference model/ECO.py:1-3 holds two import lines): it restates
  the definition in fastvideotagging_b200/model/heads.py (3x3x3 basic residual blocks 96->128->256->512 after the ECO
  paper) independently with torch.nn.functional — **parity unpinned** against the reference.
* `decision_thresh` follows model/decision_model.py:11-14.
"""
import torch
import torch.nn.functional as F

EPS = 1e-5


def _bn_eval(x, g, b, m, v, eps=EPS):
    sh = (1, -1, 1, 1, 1)
    scale = g / torch.sqrt(v + eps)
    return x * scale.reshape(sh) + (b - m * scale).reshape(sh)


def _q(t, bf16):
    return t.to(torch.bfloat16).to(torch.float32) if bf16 else t


def multitask_heads(feat, p, pool, bf16_storage=False):
    """feat: (N, 512, T', H', W') conv5_x output; p: dict of head tensors.  -> (scene, action)."""
    s = F.conv3d(feat, _q(p["scene_conv_weight"], bf16_storage), p["scene_conv_bias"], stride=(1, 2, 2))
    s = _q(torch.relu(_bn_eval(s, p["scene_bn_gamma"], p["scene_bn_beta"], p["scene_bn_mean"], p["scene_bn_var"])), bf16_storage)
    scene = s.reshape(s.shape[0], -1) @ _q(p["scene_dense_weight"], bf16_storage).t() + p["scene_dense_bias"]
    a = F.conv3d(feat, _q(p["action_conv_weight"], bf16_storage), p["action_conv_bias"], padding=(0, 1, 1))
    a = _q(torch.relu(_bn_eval(a, p["action_bn_gamma"], p["action_bn_beta"], p["action_bn_mean"], p["action_bn_var"])), bf16_storage)
    a = F.avg_pool3d(a, pool, stride=1).reshape(a.shape[0], -1)
    action = a @ p["action_dense_weight"].t() + p["action_dense_bias"]
    return scene, action


def decision_thresh(x, thresh):
    return x - thresh


def eco_lite_3d_head(x, blocks, dense_w, dense_b, bf16_storage=False):
    """x: (N, 96, T, 28, 28).  blocks: list of dicts {w1,bn1,w2,bn2[,wd,bnd],stride}, bn* = (gamma,beta,mean,var)."""
    h = _q(x, bf16_storage)
    for blk in blocks:
        s = blk["stride"]
        y = F.conv3d(h, _q(blk["w1"], bf16_storage), stride=s, padding=1)
        y = _q(torch.relu(_bn_eval(y, *blk["bn1"])), bf16_storage)
        y = _bn_eval(F.conv3d(y, _q(blk["w2"], bf16_storage), padding=1), *blk["bn2"])
        if "wd" in blk:
            sc = _q(_bn_eval(F.conv3d(h, _q(blk["wd"], bf16_storage), stride=s, padding=1), *blk["bnd"]), bf16_storage)
        else:
            sc = h
        h = _q(torch.relu(y + sc), bf16_storage)
    pooled = h.mean(dim=(2, 3, 4))
    return pooled @ dense_w.t() + dense_b
