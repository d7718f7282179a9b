"""fastvideotagging_b200 — B200-native R(2+1)D hot path behind FastVideoTagging's model and loss API.

Layout: csrc/ (sm_100a CUDA kernels + C ABI), _lib.py (ctypes binding), ops.py (tensor-level wrappers),
model/ (host-side mirror of reference model/R2Plus1.py, net.py, model/mlc_loss.py).
"""
__version__ = "0.1.0"
