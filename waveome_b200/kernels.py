"""Kernel tree of the B200 engine: the reference's kernel vocabulary without GPflow/TensorFlow.

Mirrors the objects the reference builds its models from:

* waveome/kernels.py:5-39 ``Lin``, :42-83 ``Poly``, :86-124 ``Categorical``, :127-142 ``Empty``
* the GPflow kernels waveome uses (waveome/model_search.py:1071-1076, waveome/regularization.py:23):
  ``SquaredExponential``, ``Matern12/32/52``, ``Periodic``, ``Linear``, ``Constant``, ``Polynomial``,
  ``Sum``, ``Product``; and ``Parameter`` with GPflow's bijector/prior semantics
  (``positive()`` = softplus, priors on the constrained value, ``trainable`` flag).

Attribute names follow GPflow (``.variance``, ``.lengthscales``, ``.period``, ``.base_kernel``, ``.kernels``,
``.active_dims``, ``.name``, ``.trainable_parameters``) so that host logic ported from the reference
(naming, pruning, BIC) reads the same.  These classes only *describe* a model; every number on the
fitting path is produced by the CUDA engine (``waveome_b200.engine``).
"""
from __future__ import annotations

import copy
import math
from typing import Iterable, List, Optional

import numpy as np

__all__ = [
    "Parameter", "Prior", "Horseshoe", "Laplace", "Uniform", "Kernel", "SquaredExponential", "RBF", "Matern12",
    "Matern32", "Matern52", "Periodic", "Linear", "Lin", "Constant", "Categorical", "Polynomial", "Poly", "Empty",
    "Sum", "Product", "set_trainable", "deepcopy", "positive",
]


# ----------------------------------------------------------------------------------------------
# priors (tensorflow_probability.distributions stand-ins; evaluated on device)
# ----------------------------------------------------------------------------------------------
class Prior:
    type = "none"

    def to_spec(self):
        raise NotImplementedError


class Horseshoe(Prior):
    """tfd.Horseshoe(scale) — waveome/model_classes.py:857."""

    type = "horseshoe"

    def __init__(self, scale=1.0):
        self.scale = float(scale)

    def to_spec(self):
        return {"type": "horseshoe", "scale": self.scale}


class Laplace(Prior):
    """tfd.Laplace(loc, scale) — waveome/model_fitting.py:201,210."""

    type = "laplace"

    def __init__(self, loc=0.0, scale=1.0):
        self.loc, self.scale = float(loc), float(scale)

    def to_spec(self):
        return {"type": "laplace", "loc": self.loc, "scale": self.scale}


class Uniform(Prior):
    """tfd.Uniform(low, high) — waveome/model_fitting.py:242."""

    type = "uniform"

    def __init__(self, low=0.0, high=1.0):
        self.low, self.high = float(low), float(high)

    def to_spec(self):
        return {"type": "uniform", "low": self.low, "high": self.high}


def positive(lower: Optional[float] = None):
    """gpflow.utilities.positive(): softplus bijector, optionally shifted by ``lower``."""
    return "softplus" if not lower else ("softplus_shift", float(lower))


def _softplus(u):
    return max(u, 0.0) + math.log1p(math.exp(-abs(u)))


def _softplus_inv(y):
    if y <= 0.0:            # a variance that underflowed to 0 during a fit (TFP's softplus inverse gives -inf as well)
        return -math.inf if y == 0.0 else math.nan
    return y + math.log(-math.expm1(-y))


class Parameter:
    """gpflow.Parameter stand-in: constrained value + bijector + prior + trainable flag."""

    def __init__(self, value, transform="identity", prior: Optional[Prior] = None, trainable=True, name=None):
        if isinstance(transform, tuple):
            self.transform, self.shift = transform[0], float(transform[1])
        else:
            self.transform, self.shift = transform, 0.0
        self._value = np.asarray(value, dtype=np.float64).reshape(-1)[0].item()
        self.prior = prior
        self.trainable = bool(trainable)
        self.name = name or self.transform
        self.shape = ()

    # GPflow-like surface -------------------------------------------------------------------
    def numpy(self):
        return np.float64(self._value)

    def assign(self, value):
        self._value = float(np.asarray(value, dtype=np.float64).reshape(-1)[0])

    def __float__(self):
        return float(self._value)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._value, dtype=dtype or np.float64)

    def __ge__(self, o): return float(self) >= float(o)
    def __gt__(self, o): return float(self) > float(o)
    def __le__(self, o): return float(self) <= float(o)
    def __lt__(self, o): return float(self) < float(o)
    def __mul__(self, o): return float(self) * o
    __rmul__ = __mul__

    @property
    def unconstrained(self):
        v = self._value
        if self.transform == "softplus":
            return _softplus_inv(v)
        if self.transform == "softplus_shift":
            return _softplus_inv(v - self.shift)
        if self.transform == "exp":
            return math.log(v) if v > 0.0 else (-math.inf if v == 0.0 else math.nan)
        return v

    def transform_fn(self, u):
        """Forward bijector (unconstrained -> constrained), as ``p.transform_fn`` in
        waveome/model_fitting.py:252 and waveome/model_classes.py:233."""
        u = np.asarray(u, dtype=np.float64)
        if self.transform == "softplus":
            return np.maximum(u, 0) + np.log1p(np.exp(-np.abs(u)))
        if self.transform == "softplus_shift":
            return np.maximum(u, 0) + np.log1p(np.exp(-np.abs(u))) + self.shift
        if self.transform == "exp":
            return np.exp(u)
        return u

    def to_spec(self):
        d = {"value": float(self._value), "trainable": self.trainable, "transform": self.transform,
             "prior": self.prior.to_spec() if self.prior is not None else None}
        if self.transform == "softplus_shift":
            d["shift"] = self.shift
        return d

    def __repr__(self):
        return f"Parameter({self._value!r}, transform={self.transform!r}, trainable={self.trainable}, prior={self.prior and self.prior.type})"


def set_trainable(obj, flag: bool):
    """gpflow.utilities.set_trainable for a Parameter or a kernel (all its parameters)."""
    if isinstance(obj, Parameter):
        obj.trainable = bool(flag)
    else:
        for p in obj.parameters:
            p.trainable = bool(flag)


def _fast_copy(obj, memo):
    oid = id(obj)
    if oid in memo:
        return memo[oid]
    if obj is None or isinstance(obj, (bool, int, float, str, bytes, np.generic)):
        return obj
    if type(obj) is Parameter:           # the bulk of every kernel tree: scalar fields + an optional prior object
        new = object.__new__(Parameter)
        memo[oid] = new
        new.__dict__.update(obj.__dict__)
        if obj.prior is not None:
            new.prior = _fast_copy(obj.prior, memo)
        return new
    if isinstance(obj, np.ndarray):
        return obj                       # data arrays are read-only on this path: shared, not duplicated
    if isinstance(obj, list):
        new = []
        memo[oid] = new
        new.extend(_fast_copy(v, memo) for v in obj)
        return new
    if isinstance(obj, tuple):
        return tuple(_fast_copy(v, memo) for v in obj)
    if isinstance(obj, dict):
        new = {}
        memo[oid] = new
        for k, v in obj.items():
            new[k] = _fast_copy(v, memo)
        return new
    d = getattr(obj, "__dict__", None)
    if d is not None and type(obj).__module__.startswith(__name__.rsplit(".", 1)[0]):
        new = object.__new__(type(obj))
        memo[oid] = new
        nd = new.__dict__
        for k, v in d.items():
            nd[k] = _fast_copy(v, memo)
        return new
    return copy.deepcopy(obj, memo)


def deepcopy(obj):
    """gpflow.utilities.deepcopy for the objects of this package (kernel trees, parameters, priors, models): shared
    sub-objects stay shared inside the copy; numpy data arrays are referenced, not duplicated."""
    return _fast_copy(obj, {})


# ----------------------------------------------------------------------------------------------
# kernels
# ----------------------------------------------------------------------------------------------
class Kernel:
    name = "kernel"
    _param_names: tuple = ()

    def __init__(self, active_dims=None):
        self.active_dims = [0] if active_dims is None else [int(a) for a in np.atleast_1d(active_dims)]

    # composition (GPflow flattens same-type nesting) ---------------------------------------
    def __add__(self, other):
        return Sum([self, other])

    def __mul__(self, other):
        return Product([self, other])

    # parameters ----------------------------------------------------------------------------
    @property
    def parameters(self) -> List[Parameter]:
        return [getattr(self, n) for n in self._param_names]

    @property
    def trainable_parameters(self) -> List[Parameter]:
        return [p for p in self.parameters if p.trainable]

    def named_parameters(self, prefix="kernel"):
        """(path, Parameter) pairs in GPflow ``parameter_dict`` style (".kernel.variance", ...)."""
        return [(f".{prefix}.{n}", getattr(self, n)) for n in self._param_names]

    def to_spec(self):
        spec = {"type": self.name, "dim": int(self.active_dims[0]),
                "params": {n: getattr(self, n).to_spec() for n in self._param_names}}
        return spec

    def __repr__(self):
        ps = ", ".join(f"{n}={float(getattr(self, n)):.6g}" for n in self._param_names)
        return f"{type(self).__name__}(dims={self.active_dims}, {ps})"


class _Stationary(Kernel):
    _param_names = ("variance", "lengthscales")

    def __init__(self, variance=1.0, lengthscales=1.0, active_dims=None):
        super().__init__(active_dims)
        self.variance = Parameter(variance, transform="softplus")
        self.lengthscales = Parameter(lengthscales, transform="softplus")


class SquaredExponential(_Stationary):
    name = "squared_exponential"


RBF = SquaredExponential


class Matern12(_Stationary):
    name = "matern12"


class Matern32(_Stationary):
    name = "matern32"


class Matern52(_Stationary):
    name = "matern52"


class Periodic(Kernel):
    """gpflow.kernels.Periodic(base_kernel=SquaredExponential): parameters live on the base kernel
    (variance, lengthscales) plus ``period``; ``active_dims`` is delegated to the base."""

    name = "periodic"

    def __init__(self, base_kernel: Optional[Kernel] = None, period=1.0):
        self.base_kernel = base_kernel if base_kernel is not None else SquaredExponential()
        if not isinstance(self.base_kernel, SquaredExponential):
            raise NotImplementedError("Periodic: only a SquaredExponential base kernel is supported "
                                      "(the only one waveome uses, waveome/model_search.py:1075)")
        self.period = Parameter(period, transform="softplus")

    @property
    def active_dims(self):
        return self.base_kernel.active_dims

    @active_dims.setter
    def active_dims(self, v):
        self.base_kernel.active_dims = [int(a) for a in np.atleast_1d(v)]

    @property
    def parameters(self):
        return [self.base_kernel.variance, self.base_kernel.lengthscales, self.period]

    def named_parameters(self, prefix="kernel"):
        return [(f".{prefix}.base_kernel.variance", self.base_kernel.variance),
                (f".{prefix}.base_kernel.lengthscales", self.base_kernel.lengthscales),
                (f".{prefix}.period", self.period)]

    def to_spec(self):
        return {"type": "periodic", "dim": int(self.active_dims[0]),
                "params": {"variance": self.base_kernel.variance.to_spec(),
                           "lengthscales": self.base_kernel.lengthscales.to_spec(),
                           "period": self.period.to_spec()}}

    def __repr__(self):
        return (f"Periodic(dims={self.active_dims}, variance={float(self.base_kernel.variance):.6g}, "
                f"lengthscales={float(self.base_kernel.lengthscales):.6g}, period={float(self.period):.6g})")


class _VarOnly(Kernel):
    _param_names = ("variance",)

    def __init__(self, variance=1.0, active_dims=None):
        super().__init__(active_dims)
        self.variance = Parameter(variance, transform="softplus")


class Linear(_VarOnly):
    name = "linear"


class Lin(_VarOnly):
    """waveome/kernels.py:5-39."""
    name = "lin"

    def __init__(self, active_dims=None, variance=1.0):
        super().__init__(variance=variance, active_dims=active_dims)
        self.active_index = self.active_dims[0]


class Constant(_VarOnly):
    name = "constant"


class Categorical(_VarOnly):
    """waveome/kernels.py:86-124: variance * 1[round(x) == round(x')]."""
    name = "categorical"

    def __init__(self, active_dims=None, variance=1.0):
        super().__init__(variance=variance, active_dims=active_dims)
        self.active_index = self.active_dims[0]


class Polynomial(Kernel):
    name = "polynomial"
    _param_names = ("variance", "offset")

    def __init__(self, degree=3, variance=1.0, offset=1.0, active_dims=None):
        super().__init__(active_dims)
        self.degree = int(degree)
        self.variance = Parameter(variance, transform="softplus")
        self.offset = Parameter(offset, transform="softplus")

    def to_spec(self):
        s = super().to_spec()
        s["degree"] = self.degree
        return s


class Poly(Polynomial):
    """waveome/kernels.py:42-83."""
    name = "poly"

    def __init__(self, active_dims=None, variance=1.0, offset=1.0, degree=3):
        super().__init__(degree=degree, variance=variance, offset=offset, active_dims=active_dims)
        self.active_index = self.active_dims[0]


class Empty(Kernel):
    """waveome/kernels.py:127-142: K = 0, frozen 1e-6 variance."""
    name = "empty"
    _param_names = ("variance",)

    def __init__(self):
        super().__init__([0])
        self.variance = Parameter(1e-6, transform="softplus", trainable=False)


class _Combination(Kernel):
    def __init__(self, kernels: Iterable[Kernel]):
        self.kernels: List[Kernel] = []
        for k in kernels:
            if isinstance(k, type(self)):      # GPflow flattens nested Sum-in-Sum / Product-in-Product
                self.kernels.extend(k.kernels)
            else:
                self.kernels.append(k)

    @property
    def active_dims(self):
        return sorted({d for k in self.kernels for d in k.active_dims})

    @property
    def parameters(self):
        return [p for k in self.kernels for p in k.parameters]

    def named_parameters(self, prefix="kernel"):
        out = []
        for i, k in enumerate(self.kernels):
            out += k.named_parameters(prefix=f"{prefix}.kernels[{i}]")
        return out

    def to_spec(self):
        return {"type": self.name, "kernels": [k.to_spec() for k in self.kernels]}

    def __repr__(self):
        sep = " + " if self.name == "sum" else " * "
        return "(" + sep.join(repr(k) for k in self.kernels) + ")"


class Sum(_Combination):
    name = "sum"


class Product(_Combination):
    name = "product"
