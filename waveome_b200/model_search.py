"""GPSearch — drop-in surface of waveome/model_search.py:47 for the model-fitting hot path.

``__init__`` reproduces the reference's data preparation (:73-195: factorise categoricals, z-score the
continuous X columns with pandas' ddof=1 std, optional Y transforms).  ``penalized_optimization``
(:197-517) keeps its signature; instead of one Ray worker + TensorFlow graph + SciPy loop per outcome it
packs all outcomes into engine batches (one per GPU when launched under torchrun) and runs the device
L-BFGS-B on the exact-GPR objective (objective A).
"""
from __future__ import annotations

import os
import time
import warnings
from typing import Dict, List, Optional

import numpy as np
import pandas as pd

from . import kernels as K
from .model_fitting import fit_models, fit_replicated, get_engine
from .models import ConstantMean, PenalizedGPR, make_likelihood
from .postfit import feature_importances_batch, train_predictive_variance
from .regularization import full_kernel_build


def _rank_world():
    """(rank, world) of the current torch.distributed job, (0, 1) otherwise."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def shard_bounds(n_items: int, rank: int, world: int):
    """Static contiguous split of the model list over ranks: [lo, hi) of this rank."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GPSearch:
    def __init__(self, X, Y, unit_col=None, standardize_X=True, Y_transform=None, categorical_vars=None,
                 outcome_likelihood="gaussian"):
        X = X.copy()
        if not isinstance(X, pd.DataFrame):
            raise TypeError("X is not a Pandas DataFrame")
        if not isinstance(Y, pd.DataFrame):
            raise TypeError("Y is not a Pandas DataFrame")
        categorical_vars = list(categorical_vars or [])       # (the reference mutates a shared default list)
        if unit_col is not None and unit_col not in categorical_vars:
            categorical_vars += [unit_col]
        self.categorical_dict = {}
        for c in categorical_vars:
            if X[c].dtype == object or str(X[c].dtype).startswith(("string", "str", "category")):
                factor_out = pd.factorize(X[c])
                self.categorical_dict[c] = factor_out
                X[c] = factor_out[0].astype(float)
        try:
            X = X.astype(float)
        except Exception:
            raise TypeError("X columns must all be float type. Perhaps use pandas.factorize().")
        try:
            Y = Y.astype(float)
        except Exception:
            raise TypeError("Y columns must all be float type.")
        assert X.isna().sum().sum() == 0, "NAs in X, waveome cannot currently handle missing values!"
        assert Y.isna().sum().sum() == 0, "NAs in Y, waveome cannot currently handle missing values!"
        self.X = X.copy()
        self.Y = Y.copy()
        self.feat_names = X.columns.tolist()
        self.out_names = Y.columns.tolist()
        self.cat_idx = [self.feat_names.index(x) for x in categorical_vars]
        self.unit_idx = self.feat_names.index(unit_col) if unit_col is not None else None
        self.likelihood = outcome_likelihood
        self.cont_idx = np.where(~np.isin(np.arange(X.shape[1]), self.cat_idx))[0].tolist()
        if standardize_X:
            self.X_means = self.X.iloc[:, self.cont_idx].mean(axis=0)
            self.X_stds = self.X.iloc[:, self.cont_idx].std(axis=0)
            self.X_original = self.X.copy()
            for c in self.cont_idx:
                name = self.feat_names[c]
                self.X[name] = (self.X[name] - self.X_means[name]) / self.X_stds[name]
        if Y_transform == "standardize":
            if self.likelihood != "gaussian":
                warnings.warn("Standardizing Y without a gaussian likelihood is not advised!")
            self.Y_means = self.Y.mean(axis=0)
            self.Y_stds = self.Y.std(axis=0)
            self.Y_original = self.Y.copy()
            self.Y = (self.Y - self.Y_means) / self.Y_stds
        elif Y_transform == "scale":
            self.Y_stds = self.Y.std(axis=0)
            self.Y_original = self.Y.copy()
            self.Y = self.Y / self.Y_stds
        self.models: Dict[str, PenalizedGPR] = {}
        self.fit_report: Optional[dict] = None

    # ------------------------------------------------------------------------------------------
    def penalized_optimization(self, full_kernel=None, num_jobs=-1, verbose=False, mean_function=None,
                               kernel_options=None, penalization_factor=1.0, num_factor_iter=5, num_restart=0,
                               sparse_options=None, variational_options=None, optimization_options=None,
                               random_seed=None, ray_dashboard=False, ray_logging=False, gather=True):
        """Fit the saturated penalised kernel to every outcome (waveome/model_search.py:197-517).

        ``num_jobs``, ``ray_*``, ``sparse_options`` and ``variational_options`` are accepted for signature
        compatibility; the engine always fits the exact model on the GPU.  Under ``torch.distributed`` the
        outcomes are sharded over ranks with no collective on the data path; with ``gather=True`` the fitted
        models are exchanged afterwards so every rank holds ``self.models`` for all outcomes."""
        make_likelihood(self.likelihood)              # raises for likelihoods the engine does not cover
        self.model_selection_type = "penalized"
        if random_seed is not None:
            np.random.seed(random_seed)
        kernel_options = kernel_options or {"second_order_numeric": False, "categorical_numeric_interactions": True,
                                            "unit_numeric_interactions": False, "kerns": [K.SquaredExponential()]}
        optimization_options = dict(optimization_options or {"optimizer": "scipy"})
        optimization_options.pop("optimizer", None)
        num_opt_iter = int(optimization_options.pop("num_opt_iter", 50000))   # maxiter = maxfun (model_classes.py:310-315)
        if full_kernel is None:
            full_kernel, _ = full_kernel_build(cat_vars=self.cat_idx, num_vars=self.cont_idx, unit_idx=self.unit_idx,
                                               var_names=self.feat_names, return_sum=True, **kernel_options)
        mean_function = mean_function if mean_function is not None else ConstantMean()
        rank, world = _rank_world()
        lo, hi = shard_bounds(len(self.out_names), rank, world)
        names = self.out_names[lo:hi]
        t0 = time.time()
        template = PenalizedGPR(K.deepcopy(full_kernel), mean_function=K.deepcopy(mean_function),
                                penalization_factor=1.0 if penalization_factor is None else penalization_factor,
                                likelihood=make_likelihood(self.likelihood))

        def make_models() -> List[PenalizedGPR]:      # every outcome owns its copy (model_search.py:305-306)
            return [K.deepcopy(template) for _ in names]
        Xn = self.X.to_numpy(dtype=np.float64)
        Yn = np.ascontiguousarray(self.Y[names].to_numpy(dtype=np.float64).T)
        n_fits = max(1, int(num_restart)) if num_restart else 1
        if verbose and rank == 0:
            print(f"Building {len(self.out_names)} models on {world} GPU(s)...")
        # one kernel structure for every outcome: worth run-time specialised Gram / gradient kernels when the JOB (all
        # outcomes, not this rank's shard) is large enough to pay for their compilation
        from .engine import SPECIALIZE_MIN_MODELS
        self._specialize = len(self.out_names) >= SPECIALIZE_MIN_MODELS
        post_done = np.zeros(len(names), bool)        # outcomes whose post-fit work was done behind the fit
        if penalization_factor is None:
            res, models = self._iterated_factor_fit(full_kernel, mean_function, names, Xn, Yn, num_factor_iter, num_opt_iter,
                                                    verbose and rank == 0)
        elif num_restart and num_restart > 0:
            models = make_models()
            # random_restart_optimize (model_classes.py:472-524): keep the restart with the best objective
            best = None
            for r in range(num_restart):
                seed = r if random_seed is None else random_seed + r + 1
                rs = np.random.RandomState(seed)
                for m in models:
                    for p in m.trainable_parameters:
                        p.assign(p.transform_fn(rs.normal(loc=0.0, scale=1.0)))
                res = fit_models(Xn, Yn, models, maxiter=num_opt_iter, maxfun=num_opt_iter, specialize=self._specialize)
                if best is None:
                    best = {k: np.array(v) for k, v in res.items() if isinstance(v, np.ndarray)}
                else:
                    better = -res["f"] > -best["f"]
                    for k in best:
                        best[k][better] = res[k][better]
            res = best
            for m, xb in zip(models, res["x"]):
                m.program().assign(xb[: len(m.trainable_parameters)])
        else:
            # one structure for every outcome: the device fit overlaps the construction of the model objects, and the
            # post-fit work of a piece of outcomes (pruning, importances -- on an engine of its own) runs behind the
            # pieces that are still being fitted
            from .model_fitting import lease_engine, release_engine
            side = {}

            def post(lo, hi, ms):
                if not side:                                   # an engine of its own, taken when the first piece is done
                    side["engine"], side["lease"] = lease_engine()
                for m in ms:
                    m.cut_kernel_components(Xn)
                    m.update_kernel_name()
                for m, fi in zip(ms, feature_importances_batch(Xn, Yn[lo:hi], ms, engine=side["engine"])):
                    m.feature_importances = fi
                post_done[lo:hi] = True
            try:
                res, models = fit_replicated(Xn, Yn, template, make_models=make_models, maxiter=num_opt_iter,
                                             maxfun=num_opt_iter, specialize=self._specialize, post=post)
            finally:
                if side:
                    release_engine(side["lease"])
        todo = [b for b in range(len(models)) if not post_done[b]]
        if todo:
            rest = [models[b] for b in todo]
            for m in rest:
                m.cut_kernel_components(Xn)
                m.update_kernel_name()
            # get_feature_importances of every model (model_search.py:383-387) as one more engine batch
            for m, fi in zip(rest, feature_importances_batch(Xn, Yn[todo], rest)):
                m.feature_importances = fi
        local = dict(zip(names, models))
        report = dict(n_models=len(names), seconds=time.time() - t0, n_eval=int(np.sum(res["n_eval"])),
                      status=np.asarray(res["status"]).copy(), n_fits_per_model=n_fits)
        if world > 1 and gather:
            import torch.distributed as dist
            parts = [None] * world
            dist.all_gather_object(parts, local)
            local = {}
            for p in parts:
                local.update(p)
        self.models = local
        self.fit_report = report
        return None

    # ------------------------------------------------------------------------------------------
    def _iterated_factor_fit(self, full_kernel, mean_function, names, Xn, Yn, num_factor_iter, num_opt_iter, verbose):
        """penalization_factor=None (waveome/model_search.py:271-375): start every outcome at
        2 * 1.1 * sd(y) * sqrt(n) * Phi^-1(1 - 0.1 / (2 p)), p = number of additive components, then up to
        ``num_factor_iter`` times re-estimate the residual sd from the predictive y-variance, recompute the factor and
        continue the optimisation while it keeps decreasing.  All outcomes iterate together: one engine batch per step
        for the models that are still moving."""
        from scipy.stats import norm
        from .utilities import find_variance_components
        n = Xn.shape[0]
        num_params = len(find_variance_components(full_kernel, sum_reduce=False))
        z = norm().ppf(1 - (0.1 / (2 * num_params)))
        models = []
        for b in range(len(names)):
            sigma_hat = 1 if num_factor_iter == 0 else float(np.std(Yn[b]))
            models.append(PenalizedGPR(K.deepcopy(full_kernel), mean_function=K.deepcopy(mean_function),
                                       penalization_factor=2 * 1.1 * sigma_hat * np.sqrt(n) * z,
                                       likelihood=make_likelihood(self.likelihood)))
        spec = getattr(self, "_specialize", False)
        res = fit_models(Xn, Yn, models, maxiter=num_opt_iter, maxfun=num_opt_iter, specialize=spec)
        active = list(range(len(models)))
        for _ in range(int(num_factor_iter)):
            if not active:
                break
            var_y = train_predictive_variance(Xn, Yn[active], [models[b] for b in active])
            new_pf = 2 * 1.1 * np.sqrt(np.mean(var_y, axis=1)) * np.sqrt(n) * z
            nxt = []
            for b, pf in zip(active, new_pf):
                cur = models[b].penalization_factor
                if abs(pf - cur) <= 1e-3 or pf > cur:      # similar, or larger (the reference then stops with the current fit)
                    continue
                models[b].set_penalization_factor(float(pf))
                nxt.append(b)
            active = nxt
            if verbose:
                print(f"penalization factor iteration: {len(active)} outcomes continue")
            if active:
                r = fit_models(Xn, Yn[active], [models[b] for b in active], maxiter=num_opt_iter, maxfun=num_opt_iter,
                               specialize=spec)
                for key in ("f", "lml", "n_iter", "n_eval", "status"):
                    res[key][active] = r[key]
        self.iterating_penalization_factor = True
        return res, models

    # ------------------------------------------------------------------------------------------
    def run_search(self, kernels=None, max_depth=5, early_stopping=True, prune=True, keep_all=False, metric_diff=6,
                   num_restart=1, random_seed=None, num_jobs=-1, verbose=False, debug=False, gather=True, fit=None,
                   pipeline_groups=None, optimizer="lbfgs"):
        """Greedy compositional kernel search per outcome (waveome/model_search.py:1069-1250 -> full_kernel_search
        :2987-3272).  The searches of all outcomes (of this rank's shard) advance in lock-step; at every step the
        candidate kernels they ask for are fitted as ONE engine batch (kernel_search.run_lockstep).
        ``self.models[outcome]`` = best model, ``self.search_info[outcome]`` = {"models", "edges", "best_model"}.
        ``num_jobs`` is accepted for signature compatibility.  ``fit`` replaces the engine fitter (tests).
        ``optimizer``: "lbfgs" (default) or "adam" (upstream's schedule for the candidate fits, ``kernel_search.kernel_test``).
        ``pipeline_groups``: outcome groups of the lock-step driver (default one; with more, one group's device batch
        overlaps the other's host work -- same result, measured slower on config 2, see ``kernel_search.run_lockstep``)."""
        from . import kernel_search as ks
        make_likelihood(self.likelihood)              # raises for likelihoods the engine does not cover
        self.model_selection_type = "stepwise"
        self.verbose = verbose
        if kernels is None:
            kernels = [K.SquaredExponential(), K.Matern12(), K.Lin(), K.Periodic(K.SquaredExponential())]
        if random_seed is not None:
            np.random.seed(random_seed)
        rank, world = _rank_world()
        lo, hi = shard_bounds(len(self.out_names), rank, world)
        names = self.out_names[lo:hi]
        t0 = time.time()
        Xn = self.X.to_numpy(dtype=np.float64)
        ys = {o: np.ascontiguousarray(self.Y[o].to_numpy(dtype=np.float64)) for o in names}
        if verbose and rank == 0:
            print(f"Building {len(self.out_names)} models on {world} GPU(s)...")
        counters = dict(fits=0, batches=0)
        inner = fit or ks.engine_fitter(Xn, num_restart=num_restart, random_seed=random_seed, likelihood=self.likelihood,
                                        optimizer=optimizer)

        def counted(requests, **kw):
            counters["fits"] += len(requests) * max(1, int(num_restart))
            counters["batches"] += 1
            return inner(requests, **kw)

        counted.supports_tail = getattr(inner, "supports_tail", False)      # run_lockstep: stragglers finish in the background

        gens = {o: ks.full_kernel_search_gen(Xn.shape[1], kernels, cat_vars=self.cat_idx, max_depth=max_depth,
                                             keep_all=keep_all, metric_diff=metric_diff,
                                             early_stopping=early_stopping, prune=prune) for o in names}
        info = ks.run_lockstep(gens, ys, counted, groups=pipeline_groups)
        local_models, local_info = {}, {}
        for o in names:
            best = info[o]["models"][info[o]["best_model"]]["model"]
            best.update_kernel_name() if hasattr(best, "update_kernel_name") else None
            from .utilities import kernel_name_string
            best.kernel_name = kernel_name_string(best.kernel, with_idx=True)
            best.search_name = info[o]["best_model"]
            local_models[o] = best
            local_info[o] = info[o]
        report = dict(n_models=len(names), seconds=time.time() - t0, n_fits=counters["fits"], batches=counters["batches"])
        if world > 1 and gather:
            import torch.distributed as dist
            parts = [None] * world
            dist.all_gather_object(parts, (local_models, {o: dict(best_model=v["best_model"], edges=v["edges"])
                                                          for o, v in local_info.items()}))
            local_models, local_info = {}, {}
            for pm, pi in parts:
                local_models.update(pm)
                local_info.update(pi)
        self.models = local_models
        self.search_info = local_info
        self.fit_report = report
        return None

    # ------------------------------------------------------------------------------------------
    def multioutput_penalized_optimization(self, latent_kernels=None, penalization_factor=1.0, num_opt_iter=2000,
                                           adam_learning_rate=0.01, nat_gradient_gamma=0.1, constraint_weight=1.0,
                                           sparse_options=None, variational_options=None, verbose=False, random_seed=None,
                                           kernel_options=None, device=None):
        """Fit ONE linear-coregionalisation model to all outcomes (waveome/model_search.py:519-573):
        ``self.models["multioutput"]`` = the fitted ``multioutput.MultiOutputPSVGP``."""
        from .multioutput import MultiOutputPSVGP
        if random_seed is not None:
            np.random.seed(random_seed)
        variational_options = dict(variational_options or {})
        variational_options["likelihood"] = self.likelihood
        model = MultiOutputPSVGP(X=self.X.to_numpy(dtype=np.float64), Y=self.Y.to_numpy(dtype=np.float64),
                                 latent_kernels=latent_kernels, penalization_factor=penalization_factor, verbose=verbose,
                                 sparse_options=sparse_options or {}, variational_options=variational_options,
                                 kernel_options=kernel_options if kernel_options is not None else {},
                                 cat_vars=self.cat_idx, num_vars=self.cont_idx, unit_idx=self.unit_idx,
                                 var_names=self.feat_names, device=device)
        model.optimize_params(num_opt_iter=num_opt_iter, adam_learning_rate=adam_learning_rate,
                              nat_gradient_gamma=nat_gradient_gamma, constraint_weight=constraint_weight)
        self.models = {"multioutput": model}
        return None

    def run_penalized_search(self, *args, **kwargs):
        """model_search.py:933-958: deprecated upstream, raises there as well."""
        raise NotImplementedError("run_penalized_search is deprecated, use penalized_optimization instead.")

    def reverse_transform(self, array, feature_name=None, input_type="X", round_digits=1):
        """Input values back on the original scale (model_search.py:1677-1715)."""
        if input_type == "X":
            assert hasattr(self, "X_stds"), "Standardize_X wasn't called in GPSearch()"
            scale_vals = self.X_stds.values if feature_name is None else self.X_stds[feature_name]
            shift_vals = self.X_means.values if feature_name is None else self.X_means[feature_name]
        elif input_type == "Y":
            assert hasattr(self, "Y_stds"), "Y_transform wasn't called in GPSearch()"
            scale_vals = self.Y_stds.values if feature_name is None else self.Y_stds[feature_name]
            if hasattr(self, "Y_means"):
                shift_vals = self.Y_means.values if feature_name is None else self.Y_means[feature_name]
            else:
                shift_vals = np.zeros_like(scale_vals)
        else:
            raise ValueError("Unknown type requested for transform!")
        return np.round(scale_vals * np.array(array) + shift_vals, decimals=round_digits)
