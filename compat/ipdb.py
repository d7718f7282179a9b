"""`import ipdb` (train_simple_r3d.py:22, never called on the training path)."""


def set_trace(*args, **kwargs):
    raise RuntimeError("ipdb is not installed here (compat stub)")
