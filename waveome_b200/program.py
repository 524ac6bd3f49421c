"""Kernel tree -> flat "kernel program" for the CUDA engine.

The engine evaluates a *sum of products of leaves*.  GPflow ``Sum``/``Product`` trees built by the
reference (waveome/regularization.py:14-189 ``full_kernel_build``; waveome/model_search.py:2408-2476,
2561-2664 sum / product / split-product expansions) are already in that form up to flattening; a
``Product`` that contains a ``Sum`` is expanded by distributivity (parameters stay shared, their
gradients accumulate in the same slot).

Packed parameter order (the optimiser's x vector): kernel parameters depth-first over the tree in
GPflow attribute order, then the likelihood variance, then the mean constant — trainable ones only.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import kernels as K

LEAF_CODE = {
    "squared_exponential": 0, "matern12": 1, "matern32": 2, "matern52": 3, "periodic": 4,
    "linear": 5, "lin": 5, "constant": 6, "categorical": 7, "polynomial": 8, "poly": 8, "empty": 9,
}
TRANSFORM_CODE = {"identity": 0, "softplus": 1, "softplus_shift": 2, "exp": 3}
PRIOR_CODE = {"none": 0, "horseshoe": 1, "laplace": 2, "uniform": 3}

MAX_COMP, MAX_LEAVES, MAX_SLOTS = 32, 64, 64


def expand_sum_of_products(kernel) -> List[List[K.Kernel]]:
    """Normal form: list of components, each a list of leaf kernels."""
    if isinstance(kernel, K.Sum):
        out = []
        for ch in kernel.kernels:
            out += expand_sum_of_products(ch)
        return out
    if isinstance(kernel, K.Product):
        parts = [expand_sum_of_products(ch) for ch in kernel.kernels]
        out = []
        for combo in itertools.product(*parts):
            comp = []
            for c in combo:
                comp += c
            out.append(comp)
        return out
    return [[kernel]]


def iter_leaves(kernel):
    if isinstance(kernel, (K.Sum, K.Product)):
        for ch in kernel.kernels:
            yield from iter_leaves(ch)
    else:
        yield kernel


@dataclass
class Program:
    """Flat arrays exactly as ``wv_program_desc`` (include/waveome_b200.h) wants them."""
    n_comp: int
    n_leaves: int
    n_slots: int
    n_x: int
    noise_slot: int
    mean_slot: int
    comp_start: np.ndarray
    leaf_type: np.ndarray
    leaf_dim: np.ndarray
    leaf_s_var: np.ndarray
    leaf_s_ls: np.ndarray
    leaf_s_aux: np.ndarray
    leaf_degree: np.ndarray
    slot_transform: np.ndarray
    slot_xindex: np.ndarray
    slot_prior: np.ndarray
    slot_fixed: np.ndarray
    slot_shift: np.ndarray
    slot_pa: np.ndarray
    slot_pb: np.ndarray
    params: List[K.Parameter] = field(default_factory=list)   # slot -> Parameter object
    x_params: List[K.Parameter] = field(default_factory=list)  # packed index -> Parameter object
    lik_slot2: int = -1            # slot of a second likelihood parameter (ZINB km), -1 if none

    def x0(self) -> np.ndarray:
        """Unconstrained start vector from the parameters' current values."""
        return np.array([p.unconstrained for p in self.x_params], dtype=np.float64)

    def assign(self, x: np.ndarray) -> None:
        """Write an unconstrained vector back into the Parameter objects."""
        for p, u in zip(self.x_params, np.asarray(x, dtype=np.float64)):
            p.assign(p.transform_fn(float(u)))

    def signature(self) -> tuple:
        """Hashable identity of everything the device sees (used to share programs in a batch).  The device reads
        ``slot_fixed`` of FROZEN slots only (trainable values travel in x), so the current values of trainable
        parameters do not enter: fitted models of one structure share one device program."""
        frozen_values = np.where(self.slot_xindex < 0, self.slot_fixed, 0.0)
        return tuple(a.tobytes() for a in (
            self.comp_start, self.leaf_type, self.leaf_dim, self.leaf_s_var, self.leaf_s_ls, self.leaf_s_aux,
            self.leaf_degree, self.slot_transform, self.slot_xindex, self.slot_prior, frozen_values,
            self.slot_shift, self.slot_pa, self.slot_pb)) + (self.noise_slot, self.mean_slot, self.lik_slot2)


def build_program(kernel, likelihood_variance: K.Parameter, mean_c: Optional[K.Parameter],
                  likelihood_aux: Optional[K.Parameter] = None) -> Program:
    comps = expand_sum_of_products(kernel)
    if len(comps) > MAX_COMP:
        raise ValueError(f"kernel has {len(comps)} additive components, engine limit is {MAX_COMP}")
    slot_of: Dict[int, int] = {}
    params: List[K.Parameter] = []

    def slot(p: K.Parameter) -> int:
        if id(p) not in slot_of:
            slot_of[id(p)] = len(params)
            params.append(p)
        return slot_of[id(p)]

    # register slots in packed order first: kernel depth-first, noise (or first likelihood parameter), second likelihood
    # parameter, mean
    for lf in iter_leaves(kernel):
        for p in lf.parameters:
            slot(p)
    noise_slot = slot(likelihood_variance)
    lik_slot2 = slot(likelihood_aux) if likelihood_aux is not None else -1
    mean_slot = slot(mean_c) if mean_c is not None else -1
    if len(params) > MAX_SLOTS:
        raise ValueError(f"model has {len(params)} parameters, engine limit is {MAX_SLOTS}")

    comp_start, ltype, ldim, lvar, lls, laux, ldeg = [0], [], [], [], [], [], []
    cheap = ("categorical", "constant", "linear", "lin", "empty")
    for comp in comps:
        # masks and other transcendental-free factors first: the device skips the expensive factors of a product
        # wherever the partial product is zero for a whole warp (stable order otherwise)
        comp = sorted(comp, key=lambda lf: 0 if lf.name in cheap else 1)
        for lf in comp:
            code = LEAF_CODE.get(lf.name)
            if code is None:
                raise ValueError(f"kernel '{lf.name}' is not supported by the engine")
            ltype.append(code)
            ldim.append(int(lf.active_dims[0]))
            if isinstance(lf, K.Periodic):
                lvar.append(slot(lf.base_kernel.variance)); lls.append(slot(lf.base_kernel.lengthscales))
                laux.append(slot(lf.period)); ldeg.append(0)
            elif isinstance(lf, K.Polynomial):
                lvar.append(slot(lf.variance)); lls.append(slot(lf.offset)); laux.append(-1); ldeg.append(lf.degree)
            elif isinstance(lf, K._Stationary):
                lvar.append(slot(lf.variance)); lls.append(slot(lf.lengthscales)); laux.append(-1); ldeg.append(0)
            elif isinstance(lf, K.Empty):
                lvar.append(slot(lf.variance)); lls.append(-1); laux.append(-1); ldeg.append(0)
            else:
                lvar.append(slot(lf.variance)); lls.append(-1); laux.append(-1); ldeg.append(0)
        comp_start.append(len(ltype))
    if len(ltype) > MAX_LEAVES:
        raise ValueError(f"kernel has {len(ltype)} leaves after expansion, engine limit is {MAX_LEAVES}")

    ns = len(params)
    tr = np.zeros(ns, np.int32); xi = np.full(ns, -1, np.int32); pr = np.zeros(ns, np.int32)
    fx = np.zeros(ns); sh = np.zeros(ns); pa = np.zeros(ns); pb = np.zeros(ns)
    x_params = []
    for s, p in enumerate(params):
        tr[s] = TRANSFORM_CODE[p.transform]
        sh[s] = p.shift
        fx[s] = float(p)
        if p.trainable:
            xi[s] = len(x_params)
            x_params.append(p)
            if p.prior is not None:
                pr[s] = PRIOR_CODE[p.prior.type]
                if p.prior.type == "horseshoe":
                    pa[s] = p.prior.scale
                elif p.prior.type == "laplace":
                    pa[s], pb[s] = p.prior.loc, p.prior.scale
                elif p.prior.type == "uniform":
                    pa[s], pb[s] = p.prior.low, p.prior.high
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    return Program(
        n_comp=len(comps), n_leaves=len(ltype), n_slots=ns, n_x=len(x_params), noise_slot=noise_slot,
        mean_slot=mean_slot, comp_start=i32(comp_start), leaf_type=i32(ltype), leaf_dim=i32(ldim),
        leaf_s_var=i32(lvar), leaf_s_ls=i32(lls), leaf_s_aux=i32(laux), leaf_degree=i32(ldeg),
        slot_transform=tr, slot_xindex=xi, slot_prior=pr, slot_fixed=fx, slot_shift=sh, slot_pa=pa, slot_pb=pb,
        params=params, x_params=x_params, lik_slot2=lik_slot2)
