"""Cross-validated choice of the penalisation factor on the batch engine.

Reference: ``PenalizedGP.penalization_search`` (waveome/model_classes.py:866-998) with ``make_folds``
(waveome/regularization.py:245-276): for every (factor, fold) fit the penalised model on the training rows with
``num_restart`` random restarts, score the held-out rows with the mean ``predict_log_density``, pick the factor with
the best mean score (minus one standard error with ``selection_type="se"``), refit on all rows.

There the (factor x fold) grid is a joblib pool of independent fits of ONE outcome.  Here all outcomes of a GPSearch
share X and therefore the folds, so a fold is ONE engine batch of outcomes x factors x restarts models, followed by
one ``wv_batch_predict_f`` on the held-out rows."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from . import kernels as K
from .models import ConstantMean, PenalizedGPR


def make_folds(X, unit_col, k_fold=5, random_seed=None) -> List[np.ndarray]:
    """waveome/regularization.py:245-276 — row indices of the k folds, sampled at the unit level when there is one."""
    if random_seed is not None:
        np.random.seed(random_seed)
    if unit_col is None:
        sample_idx = np.arange(0, X.shape[0])
    else:
        sample_idx = np.unique(X[:, unit_col])
        assert len(sample_idx) >= k_fold, (
            "Not enough unique units for number of folds requested, " f"{len(sample_idx)} unit(s) < {k_fold} fold(s)")
    np.random.shuffle(sample_idx)
    div, mod = divmod(len(sample_idx), k_fold)
    folds = [sample_idx[(i * div + min(i, mod)):((i + 1) * div + min(i + 1, mod))] for i in range(k_fold)]
    if unit_col is not None:
        folds = [np.where(np.isin(X[:, unit_col], f))[0] for f in folds]
    return folds


def _restart_models(template_kernel, mean_function, factor, n_models, num_restart, random_seed):
    """n_models x num_restart penalised models; restart r of every model starts at N(0, 1) draws in the unconstrained
    space from RandomState(seed_r) (random_restart_optimize, waveome/model_classes.py:472-524: seeds random_seed + 1 + r,
    or r when no seed is given)."""
    models = []
    R = max(1, int(num_restart))
    for _b in range(n_models):
        for r in range(R):
            m = PenalizedGPR(K.deepcopy(template_kernel), mean_function=K.deepcopy(mean_function), penalization_factor=factor)
            rs = np.random.RandomState(r if random_seed is None else random_seed + 1 + r)
            for p in m.trainable_parameters:
                p.assign(p.transform_fn(rs.normal(loc=0.0, scale=1.0)))
            models.append(m)
    return models


def _best_restart(f, status, B, R):
    """index of the restart with the best objective per model (failed restarts lose), [B]"""
    f = np.where((status & 1) | ~np.isfinite(f), np.inf, f).reshape(B, R)
    return np.argmin(f, axis=1)


def penalization_search_batch(X, Y, kernel, mean_function=None, penalization_factor_list=(0.0, 1.0, 10.0, 100.0), k_fold=3,
                              unit_col=None, fit_best=True, random_seed=None, num_restart=5, selection_type="se",
                              engine=None, max_iter=50000):
    """The search of :866-998 for B outcomes at once.  X [n, D], Y [B, n].  Returns dict(best_factor [B],
    results [B, F, k_fold] held-out mean log densities, folds, models (refitted on all rows when ``fit_best``))."""
    from .engine import Batch
    from .model_fitting import fit_models, get_engine
    engine = engine or get_engine()
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    B, n = Y.shape
    mean_function = mean_function if mean_function is not None else ConstantMean()
    factors = [float(f) for f in penalization_factor_list]
    F, R = len(factors), max(1, int(num_restart))
    folds = make_folds(X, unit_col, k_fold, random_seed)
    results = np.full((B, F, len(folds)), np.nan)
    selected = {}            # (outcome, factor index, fold) -> the model whose held-out score was recorded
    for k, hold in enumerate(folds):
        train = np.setdiff1d(np.arange(n), hold)
        Xt, Xh = X[train], X[hold]
        models, rows = [], []
        for fi, pf in enumerate(factors):
            models += _restart_models(kernel, mean_function, pf, B, R, random_seed)
            rows += [b for b in range(B) for _ in range(R)]
        Yt = Y[rows][:, train]
        # fit and predict with the SAME batch object: the held-out prediction uses the state of the final evaluation
        progs = [m.program() for m in models]
        uniq, prog_id, table = {}, np.empty(len(models), np.int32), []
        for i, p in enumerate(progs):
            sig = p.signature()
            if sig not in uniq:
                uniq[sig] = len(table)
                table.append(p)
            prog_id[i] = uniq[sig]
        P = max(1, max(p.n_x for p in progs))
        x0 = np.zeros((len(models), P))
        for i, p in enumerate(progs):
            x0[i, : p.n_x] = p.x0()
        batch = Batch(engine, Xt, Yt, table, prog_id, P=P)
        try:
            r = batch.fit(x0, maxiter=max_iter, maxfun=max_iter)
            mu, var = batch.predict_f(Xh)
        finally:
            batch.close()
        for fi in range(F):
            sl = slice(fi * B * R, (fi + 1) * B * R)
            best = _best_restart(r["f"][sl], r["status"][sl], B, R)
            for b in range(B):
                i = fi * B * R + b * R + best[b]
                progs[i].assign(r["x"][i, : progs[i].n_x])
                s2 = float(models[i].likelihood.variance)
                vy = var[i] + s2
                results[b, fi, k] = np.mean(-0.5 * (np.log(2 * np.pi) + np.log(vy) + (Y[b, hold] - mu[i]) ** 2 / vy))
                models[i].log_posterior_density_value = float(-r["f"][i])
                models[i].log_marginal_likelihood_value = float(r["lml"][i])
                models[i].fit_info = dict(n_iter=int(r["n_iter"][i]), n_eval=int(r["n_eval"][i]), status=int(r["status"][i]))
                selected[(b, fi, k)] = models[i]
    # best factor per outcome (:961-975): mean over folds, minus one standard error when selection_type == "se"
    best_factor = np.empty(B)
    for b in range(B):
        max_val, max_factor = -np.inf, -np.inf
        for fi, pf in enumerate(factors):
            cur_val = results[b, fi].mean()
            if selection_type == "se":
                cur_val -= results[b, fi].std() / np.sqrt(k_fold)
            if cur_val > max_val:
                max_factor, max_val = pf, cur_val
        best_factor[b] = max_factor
    out = dict(best_factor=best_factor, results=results, folds=folds, factors=factors, models=None, fold_models=selected)
    if fit_best:
        final = []
        for b in range(B):
            pf = best_factor[b] if np.isfinite(best_factor[b]) else 0.0
            final += _restart_models(kernel, mean_function, pf, 1, R, random_seed)
        rows = [b for b in range(B) for _ in range(R)]
        r = fit_models(X, Y[rows], final, engine=engine, maxiter=max_iter, maxfun=max_iter)
        best = _best_restart(r["f"], r["status"], B, R)
        out["models"] = [final[b * R + best[b]] for b in range(B)]
    return out
