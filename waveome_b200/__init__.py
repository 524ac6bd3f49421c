"""waveome_b200 — B200-native batched Gaussian-process model fitting behind waveome's API.

Hot path (BASELINE.json north_star): many independent exact-GPR models (outcomes x candidate kernel
structures), each fitted by repeated fp64 log-marginal-likelihood + gradient evaluations, executed by
hand-written sm_100a CUDA through the C ABI in ``include/waveome_b200.h``.
"""
from . import kernels  # noqa: F401
from .kernels import (Categorical, Constant, Empty, Horseshoe, Laplace, Lin, Linear, Matern12, Matern32,  # noqa: F401
                      Matern52, Parameter, Periodic, Poly, Polynomial, Product, SquaredExponential, Sum, Uniform,
                      deepcopy, set_trainable)
from .models import GPR, ConstantMean, Gaussian, ZeroMean  # noqa: F401
from .kernel_search import full_kernel_search, kernel_test  # noqa: F401

__version__ = "0.1.0"
