"""Selection arithmetic of the hot path, host side — mirrors of waveome/utilities.py:
calc_bic (:77-95), check_if_model_exists (:281-307), print_kernel_names (:366-383),
freeze_variance_parameters (:977-986), find_variance_components (:1012-1062),
keep_kernel_lengthscale_ (:1136-1153), search_through_kernel_list_ (:1156-1184)."""
from __future__ import annotations

import functools

import numpy as np

from . import kernels as K


def calc_bic(loglik: float, n: int, k: int):
    """waveome's "BIC": 2k - 2 loglik (utilities.py:95)."""
    return 2 * k - 2 * loglik


@functools.lru_cache(maxsize=1 << 18)
def canonical_model_name(model_name: str) -> frozenset:
    """The form utilities.py:281-307 compares: the SET of '+'-separated terms, each with its characters sorted."""
    return frozenset("".join(sorted(x)) for x in model_name.split("+"))


def check_if_model_exists(model_name, model_list):
    """Dedup by canonicalised name: split on '+', sort the characters of every term, compare as sets
    (utilities.py:281-307: ``set() in [set(a) ^ set(b) for b in ...]``)."""
    canon = canonical_model_name(model_name)
    return any(canon == canonical_model_name(y) for y in model_list)


def print_kernel_names(kernel, with_idx=False):
    if kernel is None:
        return ""
    if not hasattr(kernel, "kernels"):
        if with_idx:
            return kernel.name + "[" + str(kernel.active_dims[0]) + "]"
        return kernel.name
    if kernel.name == "sum":
        return [print_kernel_names(x, with_idx) for x in kernel.kernels]
    if kernel.name == "product":
        return "*".join([print_kernel_names(x, with_idx) for x in kernel.kernels])
    return []


def kernel_name_string(kernel, with_idx=True):
    """BaseGP.update_kernel_name (waveome/model_classes.py:171-179)."""
    name = print_kernel_names(kernel, with_idx=with_idx)
    if not isinstance(name, str):
        name = "+".join(list(name))
    return name


def replace_kernel_variables(k_name, col_names):
    """utilities.replace_kernel_variables: "[i]" -> "[col_name]"."""
    for i, c in enumerate(col_names):
        k_name = k_name.replace("[" + str(i) + "]", "[" + c + "]")
    return k_name


def freeze_variance_parameters(kernel):
    if hasattr(kernel, "variance"):
        K.set_trainable(kernel.variance, False)
    elif kernel.name in ["sum", "product"]:
        for k in kernel.kernels:
            freeze_variance_parameters(k)
    elif kernel.name == "periodic":
        freeze_variance_parameters(kernel.base_kernel)


def find_variance_components(kern, sum_reduce=True, penalize_factor_prod=1):
    if kern.name == "sum":
        var_list = np.stack([find_variance_components(x, sum_reduce) for x in kern.kernels])
        return np.sum(var_list) if sum_reduce else var_list
    if kern.name == "product":
        return np.array([penalize_factor_prod * np.prod([find_variance_components(x, sum_reduce) for x in kern.kernels])])
    if kern.name == "periodic":
        return np.array([kern.base_kernel.variance.numpy()])
    if kern.name == "empty":
        return np.zeros(1)
    return np.array([kern.variance.numpy()])


def keep_kernel_lengthscale_(kernel_component, X):
    if kernel_component.name == "periodic":
        kernel_component = kernel_component.base_kernel
    if not hasattr(kernel_component, "lengthscales"):
        return True
    active_index = kernel_component.active_dims[0]
    var_range = 3 * np.ptp(X[:, active_index])
    return kernel_component.lengthscales.numpy() < var_range


def search_through_kernel_list_(kernel_list, list_type="sum", X=None):
    out_list = []
    for k in kernel_list:
        if k.name == "product":
            out_list.append(search_through_kernel_list_(k.kernels, list_type="product", X=X))
        elif keep_kernel_lengthscale_(k, X):
            out_list.append(k)
    if len(out_list) > 1:
        return K.Sum(out_list) if list_type == "sum" else K.Product(out_list)
    if len(out_list) == 1:
        return out_list[0]
    return K.Empty()


def individual_kernel_predictions(model, kernel_idx, data=None, X=None, predict_type="func", marginal=True, **unused):
    """Numerical part of waveome/utilities.py:710-974: (pred_mu [m, 1], pred_var [m], None, None) of additive component
    ``kernel_idx`` at the inputs X.  The reference also returns posterior function samples and the full covariance for
    its plots; those are not produced here.  ``predict_type="mean"`` maps through the likelihood's conditional moments at
    pred_mu like the reference (:967-971)."""
    from .postfit import component_predictions
    data = data if data is not None else model.data
    if data is None:
        raise ValueError("individual_kernel_predictions needs data=(X, Y)")
    Xtr, ytr = np.asarray(data[0]), np.asarray(data[1]).reshape(-1)
    X = Xtr if X is None else np.asarray(X)
    parts = component_predictions(model, Xtr, ytr, X, marginal=marginal)
    if kernel_idx >= len(parts):
        raise ValueError("Not enough kernel components for index requested!")
    mu, var = parts[kernel_idx]
    if predict_type == "mean" and getattr(model.likelihood, "name", "gaussian") != "gaussian":
        from .postfit import likelihood_predict_mean_and_var
        mu, var = likelihood_predict_mean_and_var(model.likelihood, mu, np.zeros_like(mu))
    return mu.reshape(-1, 1), var, None, None
