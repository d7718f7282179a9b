"""`mx.initializer` alias of mx.init."""
from .init import *  # noqa: F401,F403
from .init import Xavier, Uniform, Initializer  # noqa: F401
