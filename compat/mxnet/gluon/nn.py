"""`mxnet.gluon.nn` is imported by train_simple_r3d.py:13 and not used there.  The reference's model files build their
networks from nn.Conv3D / nn.BatchNorm / ...; the equivalents on the sm_100a kernels live in model/ (R2Plus2D, R3DBlock,
get_spatial_temporal_conv) — the layer classes themselves are not re-exported as free-standing gluon layers."""
from .block import Block, HybridBlock      # noqa: F401


def __getattr__(name):
    raise NotImplementedError("gluon.nn.%s: build networks with model.R2Plus2D / R3DBlock / get_spatial_temporal_conv "
                              "(the reference's builders), which run on the sm_100a kernels" % name)
