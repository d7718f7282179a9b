"""`mx.init.Xavier` (train_simple_r3d.py:81: defaults; train.py:88: factor_type='in', magnitude=2.34)."""


class Initializer:
    pass


class Xavier(Initializer):
    def __init__(self, rnd_type="uniform", factor_type="avg", magnitude=3):
        if rnd_type != "uniform":
            raise NotImplementedError("the reference uses the uniform Xavier initialiser only")
        self.rnd_type, self.factor_type, self.magnitude = rnd_type, factor_type, float(magnitude)


class Uniform(Initializer):
    def __init__(self, scale=0.07):
        self.scale = scale
