"""Post-fit quantities on the engine: posterior means and the per-component feature importances
that ``penalized_optimization`` attaches to every model (waveome/model_search.py:383-387 ->
waveome/model_classes.py:546-573 -> waveome/utilities.py:517-707).

For the Gaussian likelihood only the posterior MEAN of ``predict_y`` enters the importances
(``calc_deviance_explained`` evaluates ``gpflow.logdensities.gaussian(x, mu, var=np.var(y))``), and for the exact-GPR
model it is  c + K alpha = y - sigma^2 alpha  at the training inputs, alpha = (K + sigma^2 I)^{-1}(y - c).  Every
"model without component k" (the reference pops the component and predicts again with the same parameter values) is
therefore ONE more factorisation: all variants of all models are evaluated as one engine batch."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from . import kernels as K
from .models import GPR

LOG2PI = 1.8378770664093453


def gaussian_logdensity(x, mu, var):
    """gpflow.logdensities.gaussian"""
    return -0.5 * (LOG2PI + np.log(var) + np.square(mu - x) / var)


def nb_logpmf(m, Y, alpha):
    """waveome/likelihoods.py:68-79"""
    from scipy.special import gammaln
    k = 1.0 / alpha
    return gammaln(k + Y) - gammaln(Y + 1) - gammaln(k) + Y * np.log(m / (m + k)) - k * np.log1p(m * alpha)


def calc_deviance_loglik(y, model_mu, base_mu=None, likelihood="gaussian", alpha=1.0):
    """calc_deviance_explained(..., return_loglik=True) (utilities.py:517-583): (base_ll, mod_ll, sat_ll) per
    observation for the Gaussian (:544-552), Poisson (:553-558) and negative-binomial (:559-581) branches."""
    y = np.asarray(y, dtype=np.float64)
    model_mu = np.asarray(model_mu, dtype=np.float64)
    if likelihood == "poisson":
        from scipy.stats import poisson
        return (poisson.logpmf(y, np.mean(y) if base_mu is None else base_mu), poisson.logpmf(y, model_mu),
                poisson.logpmf(y, y))
    if likelihood == "bernoulli":       # gpflow.logdensities.bernoulli(x, p) = log(where(x == 1, p, 1 - p))
        def bern(x, p):
            with np.errstate(divide="ignore"):
                return np.log(np.where(x == 1, p, 1.0 - p))
        return bern(y, np.mean(y) if base_mu is None else base_mu), bern(y, model_mu), bern(y, y)
    if likelihood not in ("gaussian", "negative_binomial"):
        raise ValueError("Unknown likelihood to calculate deviance")
    if likelihood == "negative_binomial":
        base = max(1e-6, np.mean(y)) if base_mu is None else base_mu
        return nb_logpmf(base, y, alpha), nb_logpmf(model_mu, y, alpha), nb_logpmf(y + 1e-6, y, alpha)
    y_var = np.var(y)
    sat_ll = gaussian_logdensity(y, y, y_var)
    base_ll = gaussian_logdensity(y, np.mean(y) if base_mu is None else base_mu, y_var)
    mod_ll = gaussian_logdensity(y, np.asarray(model_mu, dtype=np.float64), y_var)
    return base_ll, mod_ll, sat_ll


def structure_key(model: GPR) -> tuple:
    """Everything of a model the device program depends on EXCEPT the values of its trainable parameters: the kernel tree
    with dims, per parameter its (trainable, transform, shift, prior, frozen value), likelihood, mean."""
    ids = {}

    def par(p):
        pr = p.prior
        return (ids.setdefault(id(p), len(ids)),        # sharing pattern: one Parameter object in several leaves
                p.trainable, p.transform, p.shift, None if pr is None else tuple(sorted(pr.to_spec().items())),
                None if p.trainable else float(p))

    def tree(k):
        if isinstance(k, (K.Sum, K.Product)):
            return (k.name,) + tuple(tree(c) for c in k.kernels)
        return (k.name, tuple(int(d) for d in k.active_dims), getattr(k, "degree", None)) + tuple(par(p) for p in k.parameters)
    lik = model.likelihood
    return (tree(model.kernel), getattr(lik, "name", "gaussian"), tuple(par(p) for p in getattr(lik, "parameters", [])),
            getattr(model.mean_function, "name", "zero"), tuple(par(p) for p in model.mean_function.parameters))


def _component_masks(model: GPR) -> List[int]:
    """Component masks of [full model] + [model without top-level additive component k for every k]: the reference
    pops the component and predicts again with the same parameter values (utilities.py:657-662).  A top-level
    component that is a product of sums expands into several program components; all of them are switched off."""
    from .program import expand_sum_of_products
    k = model.kernel
    full = 0xFFFFFFFF
    out = [full]
    if k.name == "sum":
        pos = 0
        for child in k.kernels:
            cnt = len(expand_sum_of_products(child))
            bits = ((1 << cnt) - 1) << pos
            out.append(full & ~bits)
            pos += cnt
    return out


def fitted_means(X: np.ndarray, Y: np.ndarray, models: Sequence[GPR], masks: Optional[Sequence[int]] = None, engine=None,
                 max_batch_bytes: float = 60e9):
    """predict_y mean of every model at the training inputs, [B, n], with the models' current parameter values (one
    engine evaluation).  Gaussian: y - sigma^2 alpha; count likelihoods: exp(m + v/2) at the converged sites.
    ``masks``: optional component mask per model."""
    from .engine import Batch
    from .model_fitting import get_engine
    engine = engine or get_engine()
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    B = len(models)
    # Device programs: one per distinct structure.  Fitted models of one penalized_optimization call differ in their
    # parameter VALUES and in which components survived the pruning; models that agree in structure_key() (component
    # names, frozen flags and values, priors, likelihood, mean) share the first one's Program -- building 2000 Program
    # objects cost more than the device evaluation -- and only contribute their packed parameter vector.
    from .model_fitting import packed_parameters
    by_model, by_key, table, uniq = {}, {}, [], {}
    prog_id = np.empty(B, np.int32)
    xs = [None] * B
    for b, m in enumerate(models):                       # the same model object may appear many times (once per mask)
        if id(m) not in by_model:
            key = structure_key(m)
            if key not in by_key:
                p = m.program()
                sig = p.signature()
                if sig not in uniq:
                    uniq[sig] = len(table)
                    table.append(p)
                by_key[key] = (uniq[sig], p.n_x)
            pid, n_x = by_key[key]
            x0 = np.array([q.unconstrained for q in packed_parameters(m)], dtype=np.float64)
            assert x0.size == n_x, "structure_key does not determine the program"
            by_model[id(m)] = (pid, x0)
        prog_id[b], xs[b] = by_model[id(m)]
    P = max(1, max(v.size for v in xs))
    x = np.zeros((B, P))
    for b, v in enumerate(xs):
        x[b, : v.size] = v
    n = X.shape[0]
    npad = ((n + 1 + 7) // 8 * 8 + 63) // 64 * 64
    chunk = max(1, int(max_batch_bytes // (2 * npad * npad * 8 + npad * 64 * 8)))
    from .model_fitting import likelihood_key
    mean = np.empty((B, n))
    status = np.empty(B, np.int32)
    groups = {}
    for b, m in enumerate(models):
        groups.setdefault(likelihood_key(m), []).append(b)
    masks = None if masks is None else np.asarray(masks, dtype=np.uint32)
    for (lik_name, lik_param), idx in groups.items():
        idx = np.asarray(idx)
        for lo in range(0, len(idx), chunk):
            sel = idx[lo: lo + chunk]
            batch = Batch(engine, X, Y[sel], table, prog_id[sel], P=P)
            try:
                if lik_name != "gaussian":
                    batch.set_likelihood(lik_name, lik_param)
                if masks is not None:
                    batch.set_component_mask(masks[sel])
                _f, _g, _lml, st = batch.eval(x[sel])
                if lik_name == "gaussian":
                    s2 = np.array([float(models[b].likelihood.variance) for b in sel])
                    mean[sel] = Y[sel] - s2[:, None] * batch.alpha()
                else:       # E[y] under q(f) = N(m, v) (predict_y of the non-Gaussian likelihoods)
                    fm, fv = batch.latent()
                    if lik_name == "bernoulli":       # closed form of gpflow's Bernoulli with the inv_probit link
                        from scipy.special import erfc
                        mean[sel] = 1e-3 + (1.0 - 2e-3) * 0.5 * erfc(-fm / np.sqrt(2.0 * (1.0 + fv)))
                    elif lik_name == "gamma":         # conditional mean shape * exp(f)
                        shape = np.array([float(models[b].likelihood.shape) for b in sel])[:, None]
                        mean[sel] = shape * np.exp(fm + 0.5 * fv)
                    elif lik_name == "negative_binomial":   # waveome's override plugs in Fmu (likelihoods.py:48-51)
                        mean[sel] = np.exp(fm)
                    elif lik_name == "zinb":          # quadrature of the conditional mean m^2 / (km + m)
                        mean[sel] = np.stack([likelihood_predict_mean_and_var(models[b].likelihood, fm[i], fv[i])[0]
                                              for i, b in enumerate(sel)])
                    else:                             # Poisson, exp link
                        mean[sel] = np.exp(fm + 0.5 * fv)
            finally:
                batch.close()
            status[sel] = st
    return mean, status


def train_predictive_variance(X, Y, models: Sequence[GPR], engine=None, max_batch_bytes: float = 60e9) -> np.ndarray:
    """``predict_y(X)[1]`` at the training inputs for B models, [B, n] (what ``penalization_factor=None`` iterates on,
    waveome/model_search.py:333).  Gaussian (gpflow GPR): var f_i = sigma^2 - sigma^4 diag((K + sigma^2 I)^-1)_i,
    var y_i = var f_i + sigma^2.  Other likelihoods: the likelihood's ``predict_mean_and_var`` of the latent posterior
    q(f_i) = N(m_i, v_i) under the converged sites (``likelihood_predict_mean_and_var``)."""
    from .engine import Batch
    from .model_fitting import get_engine, likelihood_key
    engine = engine or get_engine()
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    B, n = len(models), X.shape[0]
    progs = [m.program() for m in models]
    uniq, prog_id, table = {}, np.empty(B, np.int32), []
    for b, p in enumerate(progs):
        sig = p.signature()
        if sig not in uniq:
            uniq[sig] = len(table)
            table.append(p)
        prog_id[b] = uniq[sig]
    P = max(1, max(p.n_x for p in progs))
    x = np.zeros((B, P))
    for b, p in enumerate(progs):
        x[b, : p.n_x] = p.x0()
    npad = ((n + 1 + 7) // 8 * 8 + 63) // 64 * 64
    chunk = max(1, int(max_batch_bytes // (2 * npad * npad * 8 + npad * 64 * 8)))
    out = np.empty((B, n))
    groups = {}
    for b, m in enumerate(models):
        groups.setdefault(likelihood_key(m), []).append(b)
    for (lik_name, lik_param), idx in groups.items():
        idx = np.asarray(idx)
        for lo in range(0, len(idx), chunk):
            sel = idx[lo: lo + chunk]
            batch = Batch(engine, X, Y[sel], table, prog_id[sel], P=P)
            try:
                if lik_name != "gaussian":
                    batch.set_likelihood(lik_name, lik_param)
                batch.eval(x[sel])
                if lik_name == "gaussian":
                    d = batch.kinv_diag()
                    s2 = np.array([float(models[b].likelihood.variance) for b in sel])[:, None]
                    out[sel] = (s2 - s2 * s2 * d) + s2
                else:
                    fm, fv = batch.latent()
                    for i, b in enumerate(sel):
                        out[b] = likelihood_predict_mean_and_var(models[b].likelihood, fm[i], fv[i])[1]
            finally:
                batch.close()
    return out


def feature_importances_batch(X, Y, models: Sequence[GPR], return_value="log_bf", engine=None) -> List[list]:
    """calc_feature_importance_components (utilities.py:614-707) for B models at once: one list per model with one
    entry per additive component and a last entry for the residual (1 - deviance explained)."""
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    no_dev = [getattr(m.likelihood, "name", "gaussian") in ("gamma", "zinb") for m in models]
    if any(no_dev):
        # calc_deviance_explained has no branch for these likelihoods (utilities.py:544-581: "Unknown likelihood"):
        # no importances rather than a failed fit
        keep = [b for b in range(len(models)) if not no_dev[b]]
        sub = feature_importances_batch(X, Y[keep], [models[b] for b in keep], return_value, engine) if keep else []
        it = iter(sub)
        return [None if no_dev[b] else next(it) for b in range(len(models))]
    var_models, var_masks, rows = [], [], []
    for b, m in enumerate(models):
        ms = _component_masks(m)
        var_models += [m] * len(ms)
        var_masks += ms
        rows += [b] * len(ms)
    means, status = fitted_means(X, Y[rows], var_models, masks=var_masks, engine=engine)
    all_gaussian = all(getattr(m.likelihood, "name", "gaussian") == "gaussian" for m in models)
    if all_gaussian and len(models):
        # calc_deviance_loglik's Gaussian branch for every model and variant in three array operations
        rows_a = np.asarray(rows)
        y_var = np.var(Y, axis=1)
        g_null = np.sum(gaussian_logdensity(Y, np.mean(Y, axis=1, keepdims=True), y_var[:, None]), axis=1)
        g_sat = np.sum(gaussian_logdensity(Y, Y, y_var[:, None]), axis=1)
        g_mod = np.sum(gaussian_logdensity(Y[rows_a], means, y_var[rows_a][:, None]), axis=1)
    out, pos = [], 0
    for b, m in enumerate(models):
        nv = 1 + (len(m.kernel.kernels) if m.kernel.name == "sum" else 0)
        if all_gaussian:
            null_sum, sat_sum, mod_sums = g_null[b], g_sat[b], g_mod[pos: pos + nv]
            pos += nv
        else:
            mu = means[pos: pos + nv]
            pos += nv
            y = Y[b]
            lk = dict(likelihood=getattr(m.likelihood, "name", "gaussian"),
                      alpha=float(getattr(m.likelihood, "engine_param", 1.0)) or 1.0)
            # one vectorised call for the full model and every leave-one-component-out model (rows of mu)
            null_lls, all_mod_lls, sat_lls = calc_deviance_loglik(y, mu, **lk)
            null_sum, sat_sum = np.sum(null_lls), np.sum(sat_lls)
            mod_sums = np.sum(np.atleast_2d(all_mod_lls), axis=-1)
        mod_sum = mod_sums[0]
        if sat_sum >= mod_sum and mod_sum >= null_sum:
            full_de = 1 - (-2 * (mod_sum - sat_sum) / (-2 * (null_sum - sat_sum)))
            full_de = max(min(1, full_de), 0)
        else:
            full_de = 0
        de_list = []
        k = m.kernel
        if k.name == "sum":
            for k_idx in range(len(k.kernels)):
                sub_sum = mod_sums[1 + k_idx]
                if return_value == "statistic":
                    scaled = max(np.round(-2 * (sub_sum - mod_sum), 1), 0)
                elif return_value == "log_bf":
                    scaled = np.round(mod_sum - sub_sum, 1)
                else:
                    scaled = 1 - (-2 * (sub_sum - mod_sum) / (-2 * (null_sum - mod_sum)))
                    scaled = np.round(max(min(1, scaled), 0), 3)
                de_list.append(float(scaled))
        elif k.name == "constant":
            de_list.append(0.0)
        else:
            if return_value == "statistic":
                de_list.append(float(np.round(-2 * (null_sum - mod_sum), 1)))
            elif return_value == "log_bf":
                de_list.append(float(np.round(mod_sum - null_sum, 1)))
            else:
                de_list.append(float(np.round(full_de, 3)))
        de_list.append(float(np.round(1 - full_de, 3)))
        out.append(de_list)
    return out


def predict_f(model: GPR, X, y, Xnew, engine=None):
    """gpflow GPModel.predict_f(Xnew) (full_cov=False) for one model: ([m] mean, [m] variance of f).  Non-Gaussian
    likelihoods: the posterior of the latent GP under the converged Gaussian sites (the same cross-covariance kernels:
    alpha and (K + D)^-1 of the last sweep are what the Gaussian path leaves behind)."""
    from .engine import Batch
    from .model_fitting import get_engine, likelihood_key
    engine = engine or get_engine()
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(1, -1)
    batch = Batch(engine, X, y, [model.program()])
    try:
        lik_name, lik_param = likelihood_key(model)
        if lik_name != "gaussian":
            batch.set_likelihood(lik_name, lik_param)
        batch.eval(batch.x0())
        mean, var = batch.predict_f(np.asarray(Xnew, dtype=np.float64))
        return mean[0], var[0]
    finally:
        batch.close()


def component_predictions(model: GPR, X, y, Xnew, marginal: bool = False, engine=None):
    """Posterior of every top-level additive component at new inputs: list of ([m] mean, [m] variance), one entry per
    component of ``model.kernel`` (one entry = the whole kernel when it is not a sum).

    marginal=False — the component's share of the joint posterior (individual_kernel_predictions(..., marginal=False),
    waveome/utilities.py:829-935): mean_k = c + K*_k^T (K + S)^-1 (y - c), var_k = k**_k - K*_k^T (K + S)^-1 K*_k with
    K + S the FULL model's covariance (S = sigma^2 I, or the site covariance of a non-Gaussian model): one evaluation
    with every component on, then one cross-covariance pass per component with only that component's mask.
    marginal=True — the reference's default (:820-828): the model whose kernel IS the component, i.e. the mask is
    applied to the factorisation as well."""
    from .engine import Batch
    from .model_fitting import get_engine, likelihood_key
    engine = engine or get_engine()
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(1, -1)
    Xnew = np.asarray(Xnew, dtype=np.float64)
    masks = _component_masks(model)
    full = masks[0]
    only = [full & ~m for m in masks[1:]] or [full]          # masks[1 + k] switches component k off
    batch = Batch(engine, X, y, [model.program()])
    out = []
    try:
        lik_name, lik_param = likelihood_key(model)
        if lik_name != "gaussian":
            batch.set_likelihood(lik_name, lik_param)
        x = batch.x0()
        if not marginal:
            batch.eval(x)
        for mk in only:
            batch.set_component_mask(np.array([mk], dtype=np.uint32))
            if marginal:
                batch.eval(x)
            mean, var = batch.predict_f(Xnew)
            out.append((mean[0], var[0]))
    finally:
        batch.close()
    return out


def predict_mean(model: GPR, X, y, Xnew, engine=None) -> np.ndarray:
    """gpflow GPR.predict_f(Xnew)[0] for one model: [m] posterior mean at new inputs."""
    from .engine import Batch
    from .model_fitting import get_engine
    engine = engine or get_engine()
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(1, -1)
    p = model.program()
    batch = Batch(engine, X, y, [p])
    try:
        batch.eval(batch.x0())
        return batch.predict_mean(np.asarray(Xnew, dtype=np.float64))[0]
    finally:
        batch.close()


# ------------------------------------------------------------------------------------------------
# likelihood.predict_mean_and_var / predict_log_density of the non-Gaussian likelihoods (host side, [m] vectors)
# ------------------------------------------------------------------------------------------------
_GH_X, _GH_W = np.polynomial.hermite.hermgauss(20)         # gpflow's default quadrature


def _likelihood_log_prob(lik, f, y):
    """log p(y | f) of models.Poisson / NegativeBinomial / Bernoulli / Gamma, broadcast."""
    from scipy.special import erfc, gammaln
    name = lik.name
    if name == "poisson":
        return y * f - np.exp(f) - gammaln(y + 1.0)
    if name == "negative_binomial":
        return nb_logpmf(np.exp(f), y, float(lik.alpha))
    if name == "bernoulli":
        p = 1e-3 + (1.0 - 2e-3) * 0.5 * erfc(-f / np.sqrt(2.0))
        return np.log(np.where(y == 1, p, 1.0 - p))
    if name == "gamma":
        a = float(lik.shape)
        return -a * f - gammaln(a) + (a - 1.0) * np.log(y) - y * np.exp(-f)
    if name == "zinb":            # waveome/likelihoods.py:114-133
        a, km = float(lik.alpha), float(lik.km)
        m = np.exp(f)
        log_p_zero = np.log(km + m * np.exp(-np.log1p(a * m) / a)) - np.log(km + m)
        log_p_nonzero = f - np.log(km + m) + nb_logpmf(m, np.where(y == 0, 1.0, y), a)
        return np.where(y == 0, log_p_zero, log_p_nonzero)
    raise NotImplementedError(name)


def likelihood_predict_mean_and_var(lik, fm, fv):
    """(E[y], Var[y]) under f ~ N(fm, fv): gpflow's 20-point Gauss-Hermite of the conditional moments (Poisson, Gamma),
    its closed form for Bernoulli / inv_probit, waveome's plug-in override for the negative binomial
    (waveome/likelihoods.py:41-51)."""
    from scipy.special import erfc
    fm, fv = np.asarray(fm, dtype=np.float64), np.asarray(fv, dtype=np.float64)
    name = lik.name
    if name == "negative_binomial":
        m = np.exp(fm)
        return m, m + float(lik.alpha) * m * m
    if name == "bernoulli":
        p = 1e-3 + (1.0 - 2e-3) * 0.5 * erfc(-fm / np.sqrt(2.0 * (1.0 + fv)))
        return p, p - p * p
    f = fm[..., None] + np.sqrt(2.0 * fv)[..., None] * _GH_X
    w = _GH_W / np.sqrt(np.pi)
    if name == "poisson":
        cm, cv = np.exp(f), np.exp(f)
    elif name == "gamma":
        cm, cv = float(lik.shape) * np.exp(f), float(lik.shape) * np.exp(2.0 * f)
    elif name == "zinb":          # waveome/likelihoods.py:135-143
        a, km = float(lik.alpha), float(lik.km)
        m = np.exp(f)
        psi = 1.0 - m / (km + m)
        cm = m * (1.0 - psi)
        cv = m * (1.0 - psi) * (1.0 + m * (psi + a))
    else:
        raise NotImplementedError(name)
    ey = np.sum(w * cm, -1)
    return ey, np.sum(w * (cv + cm * cm), -1) - ey * ey


def likelihood_predict_log_density(lik, fm, fv, y):
    """log int p(y | f) N(f; fm, fv) df by the same quadrature in log space (gpflow ScalarLikelihood); Bernoulli: the
    log density at the predictive mean (gpflow.likelihoods.Bernoulli._predict_log_density)."""
    from scipy.special import logsumexp
    fm, fv, y = (np.asarray(a, dtype=np.float64) for a in (fm, fv, y))
    if lik.name == "bernoulli":
        p, _ = likelihood_predict_mean_and_var(lik, fm, fv)
        return np.log(np.where(y == 1, p, 1.0 - p))
    f = fm[..., None] + np.sqrt(2.0 * fv)[..., None] * _GH_X
    return logsumexp(_likelihood_log_prob(lik, f, y[..., None]) + np.log(_GH_W / np.sqrt(np.pi)), axis=-1)
