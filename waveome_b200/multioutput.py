"""Multi-output penalised model: the counterpart of ``MultiOutputPSVGP`` (waveome/model_classes.py:1129-1612) and of
``GPSearch.multioutput_penalized_optimization`` (waveome/model_search.py:519-573) — SURVEY §8(f) row 4.

Model (as upstream): P outputs are mixtures  f_p = sum_q W[p, q] g_q + c  of Q independent latent GPs (GPflow's
``LinearCoregionalization`` over the latent kernels, ``SeparateIndependentInducingVariables``); the latent kernels'
variances are frozen (the scale lives in W), W carries a horseshoe prior of scale 1 / (penalization_factor sqrt(Q)); the
bound is the whitened SVGP ELBO with one (q_mu_q, q_sqrt_q) per latent,

    g_q | u_q:  mean A_q^T q_mu_q,  var k_q(x, x) - |A_q|^2 + |q_sqrt_q^T A_q|^2,   A_q = chol(Kuu_q)^-1 Kuf_q
    elbo = sum_{i,p} E_q log p(y_ip | f_ip) - sum_q KL[N(q_mu_q, q_sqrt_q q_sqrt_q^T) || N(0, I)]

and the optimiser alternates a natural-gradient step on the variational parameters with an Adam step (per-variable
gradient clipping, weak sign constraint on W[0, :]) on everything else, ``optimize_params`` (:1473-1612).

This objective is NOT the engine's hot path (it has no n x n factorisation per model: Q Cholesky factorisations of
M x M with M <= 100 inducing points and thin M x n products), so it is written on torch tensor operations — cuSOLVER /
cuBLAS through PyTorch on the GPU, autograd for the gradients — rather than on hand-written kernels; on a CUDA machine
it runs on ``cuda:LOCAL_RANK``, ``device="cpu"`` exists for the CPU test-suite.  The hand-written engine is what fits the
single-output models of ``penalized_optimization`` / ``run_search``.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np

from . import kernels as K
from .regularization import full_kernel_build

JITTER = 1e-6           # gpflow.config.default_jitter()


def calculate_rank_estimate(Y, threshold=0.90, transform_counts=True) -> int:
    """waveome/utilities.py:1393-1422: number of principal components explaining ``threshold`` of the variance."""
    Y = np.asarray(Y, dtype=np.float64)
    if transform_counts:
        Y = np.log1p(Y)
    Ys = (Y - Y.mean(axis=0)) / (Y.std(axis=0) + 1e-6)
    s = np.linalg.svd(Ys, full_matrices=False, compute_uv=False)
    cum = np.cumsum(s ** 2 / np.sum(s ** 2))
    return int(np.argmax(cum >= threshold) + 1)


def _default_device():
    import os
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("waveome_b200.multioutput: no CUDA device (pass device='cpu' explicitly to run on the host)")
    return torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))


def _softplus(u):
    import torch
    return torch.clamp(u, min=0) + torch.log1p(torch.exp(-torch.abs(u)))


def _horseshoe_logp(x, s):
    """tfd.Horseshoe(scale=s).log_prob(x) — TFP's closed-form approximation (the engine's wv_horseshoe, in torch)"""
    import torch
    g, b, h_inf, pw = 0.5614594835668851, 1.0420764938351215, 1.0801359952503342, 1.0919284281983377
    xx = (x / s) ** 2 / 2
    q = 20.0 / 47.0 * xx ** pw
    h = 1.0 / (1 + xx ** 1.5) + h_inf * q / (1 + q)
    c = -0.5 * math.log(2 * math.pi ** 3) - math.log(g * s)
    z = math.log1p(-g) - math.log(g)
    return -torch.nn.functional.softplus(z - xx / (1 - g)) + torch.log(torch.log1p(g / xx - (1 - g) / (h + b * xx) ** 2)) + c


class _Params:
    """Trainable Parameter objects of the latent kernels -> one unconstrained torch vector"""

    def __init__(self, kernels, dtype, device):
        import torch
        self.objs: List[K.Parameter] = []
        seen = set()
        for k in kernels:
            for p in k.parameters:
                if p.trainable and id(p) not in seen:
                    seen.add(id(p))
                    self.objs.append(p)
        self.u = torch.tensor([p.unconstrained for p in self.objs], dtype=dtype, device=device)

    def values(self, u):
        """{id(Parameter): constrained torch scalar} for the trainable ones (frozen ones are read from the object)"""
        import torch
        out = {}
        for i, p in enumerate(self.objs):
            if p.transform == "softplus":
                out[id(p)] = _softplus(u[i])
            elif p.transform == "softplus_shift":
                out[id(p)] = _softplus(u[i]) + p.shift
            elif p.transform == "exp":
                out[id(p)] = torch.exp(u[i])
            else:
                out[id(p)] = u[i]
        return out

    def write_back(self, u):
        for p, v in zip(self.objs, u.detach().cpu().numpy()):
            p.assign(p.transform_fn(float(v)))


def _kernel_matrix(kern, X1, X2, vals):
    """k(X1, X2) [n1, n2] of a kernel tree of this package in torch (GPflow semantics, waveome/kernels.py for Lin /
    Categorical); ``vals`` maps trainable Parameter ids to torch scalars."""
    import torch

    def val(p):
        return vals[id(p)] if id(p) in vals else float(p)
    if isinstance(kern, K.Sum):
        return sum(_kernel_matrix(c, X1, X2, vals) for c in kern.kernels)
    if isinstance(kern, K.Product):
        out = None
        for c in kern.kernels:
            k = _kernel_matrix(c, X1, X2, vals)
            out = k if out is None else out * k
        return out
    d = int(kern.active_dims[0])
    a, b = X1[:, d], X2[:, d]
    if isinstance(kern, K.Periodic):
        base = kern.base_kernel
        s = torch.sin(math.pi * (a[:, None] - b[None, :]) / val(kern.period)) / val(base.lengthscales)
        return val(base.variance) * torch.exp(-0.5 * s * s)
    if isinstance(kern, K._Stationary):
        ell = val(kern.lengthscales)
        pa, pb = a / ell, b / ell
        r2 = torch.clamp(-2 * torch.outer(pa, pb) + (pa * pa)[:, None] + (pb * pb)[None, :], min=0)
        if kern.name == "squared_exponential":
            return val(kern.variance) * torch.exp(-0.5 * r2)
        r = torch.sqrt(torch.clamp(r2, min=1e-36))
        if kern.name == "matern12":
            return val(kern.variance) * torch.exp(-r)
        if kern.name == "matern32":
            return val(kern.variance) * (1 + math.sqrt(3) * r) * torch.exp(-math.sqrt(3) * r)
        return val(kern.variance) * (1 + math.sqrt(5) * r + 5.0 / 3.0 * r * r) * torch.exp(-math.sqrt(5) * r)
    if kern.name in ("linear", "lin"):
        return val(kern.variance) * torch.outer(a, b)
    if kern.name == "constant":
        return val(kern.variance) * torch.ones(len(a), len(b), dtype=X1.dtype, device=X1.device)
    if kern.name == "categorical":
        return val(kern.variance) * (torch.round(a)[:, None] == torch.round(b)[None, :]).to(X1.dtype)
    raise NotImplementedError(f"latent kernel '{kern.name}' is not covered by the multi-output model")


def _kernel_diag(kern, X, vals):
    import torch

    def val(p):
        return vals[id(p)] if id(p) in vals else float(p)
    if isinstance(kern, K.Sum):
        return sum(_kernel_diag(c, X, vals) for c in kern.kernels)
    if isinstance(kern, K.Product):
        out = None
        for c in kern.kernels:
            k = _kernel_diag(c, X, vals)
            out = k if out is None else out * k
        return out
    x = X[:, int(kern.active_dims[0])]
    one = torch.ones(len(x), dtype=X.dtype, device=X.device)
    if isinstance(kern, K.Periodic):
        return val(kern.base_kernel.variance) * one
    if kern.name in ("linear", "lin"):
        return val(kern.variance) * x * x
    return val(kern.variance) * one


class _Coregion:
    """``model.kernel`` of the reference: the latent kernels and the mixing matrix W [P, Q]"""

    def __init__(self, kernels, W):
        self.kernels = list(kernels)
        self.W = np.array(W, dtype=np.float64)
        self.name = "linear_coregionalization"


class MultiOutputPSVGP:
    """Drop-in for waveome/model_classes.py:1129-1612 (constructor arguments, ``optimize_params``,
    ``prune_latent_factors``, ``predict_f``, ``.kernel.W`` / ``.kernel.kernels``, ``.likelihood_variance``)."""

    def __init__(self, X, Y, latent_kernels=None, mean_function=None, verbose=False, num_latent_gps=None,
                 penalization_factor=1.0, dtype=None, kernel_options=None, cat_vars=(), num_vars=(), unit_idx=None,
                 var_names=None, sparse_options=None, variational_options=None, device=None, **unused):
        import torch
        kernel_options = dict(kernel_options or {})
        sparse_options = dict(sparse_options or {})
        variational_options = dict(variational_options or {})
        Xn = X.to_numpy() if hasattr(X, "to_numpy") else np.asarray(X)
        Yn = Y.to_numpy() if hasattr(Y, "to_numpy") else np.asarray(Y)
        Xn, Yn = np.asarray(Xn, dtype=np.float64), np.asarray(Yn, dtype=np.float64)
        self.likelihood_name = variational_options.get("likelihood", "gaussian")
        if self.likelihood_name not in ("gaussian", "poisson"):
            raise NotImplementedError("multi-output likelihoods covered: gaussian, poisson")
        n, P = Yn.shape
        cat_vars, num_vars = list(cat_vars), list(num_vars)
        if latent_kernels is None:                                   # :1156-1241
            if "ranks" not in kernel_options:
                kernel_options["ranks"] = calculate_rank_estimate(Yn, 0.90, transform_counts=self.likelihood_name == "poisson")
                if verbose:
                    print(f"No rank provided. Estimated rank Q={kernel_options['ranks']} (explains 90% variance).")
            k_opts = {"second_order_numeric": False, "categorical_numeric_interactions": True,
                      "unit_numeric_interactions": False, "kerns": [K.SquaredExponential()], **kernel_options,
                      "num_outputs": P}
            if not num_vars and not cat_vars:
                num_vars = list(range(Xn.shape[1]))
            elif not num_vars:
                num_vars = sorted(set(range(Xn.shape[1])) - set(cat_vars))
            built = full_kernel_build(cat_vars=cat_vars, num_vars=num_vars, unit_idx=unit_idx, var_names=var_names,
                                      return_sum=False, **k_opts)
            latent_kernels = built[0] if isinstance(built, tuple) else built
            if verbose:
                print(f"Built {len(latent_kernels)} latent kernels.")
        latent_kernels = [K.deepcopy(k) for k in latent_kernels]
        Q = len(latent_kernels) if num_latent_gps is None else int(num_latent_gps)
        if Q != len(latent_kernels):
            raise ValueError("num_latent_gps must equal the number of latent kernels")
        for k in latent_kernels:                                     # freeze_variance_parameters (:1383-1386)
            for path, p in k.named_parameters():
                if "variance" in path:
                    p.trainable = False
        self.kernel = _Coregion(latent_kernels, np.random.normal(scale=0.01, size=(P, Q)))      # :1246
        self.verbose = verbose
        self.penalization_factor = float(penalization_factor)
        adj = self.penalization_factor * math.sqrt(Q)
        self.w_prior_scale = 1.0 / adj if adj > 0 else 1.0           # :1363-1378
        self.mean_c = 0.0 if mean_function is None else float(getattr(mean_function, "c", 0.0))
        self.likelihood_variance = 1.0
        self.device = torch.device(device) if device is not None else _default_device()
        self.dtype = torch.float64
        self.X = torch.tensor(Xn, dtype=self.dtype, device=self.device)
        self.Y = torch.tensor(Yn, dtype=self.dtype, device=self.device)
        self.data = (Xn, Yn)
        # inducing points (:1257-1339): all rows when num_inducing_points >= n, else per latent kernel a grid over its
        # active dimension (unique values for a categorical one), the other columns at their means
        M = int(sparse_options.get("num_inducing_points", min(n, 100)))
        self.Z = []
        for k in latent_kernels:
            if M >= n:
                Zq = Xn.copy()
            else:
                dims = getattr(k, "active_dims", None)
                if dims is not None and len(dims) == 1 and not isinstance(k, (K.Sum, K.Product)):
                    d = int(dims[0])
                    Zq = np.repeat(Xn.mean(axis=0, keepdims=True), M, axis=0)
                    if isinstance(k, K.Categorical):
                        uniq = np.unique(Xn[:, d])
                        if len(uniq) >= M:
                            np.random.seed(sparse_options.get("random_seed"))
                            grid = np.random.choice(uniq, M, replace=False)
                        else:
                            grid = np.tile(uniq, int(np.ceil(M / len(uniq))))[:M]
                    else:
                        grid = np.linspace(Xn[:, d].min(), Xn[:, d].max(), M)
                    Zq[:, d] = grid
                else:
                    np.random.seed(sparse_options.get("random_seed"))
                    Zq = Xn[np.random.choice(n, M, replace=False)].copy()
            self.Z.append(torch.tensor(Zq, dtype=self.dtype, device=self.device))
        Ms = [z.shape[0] for z in self.Z]
        self.q_mu = [torch.zeros(m, dtype=self.dtype, device=self.device) for m in Ms]
        self.q_sqrt = [torch.eye(m, dtype=self.dtype, device=self.device) for m in Ms]
        self._params = _Params(latent_kernels, self.dtype, self.device)
        self.optimizer = None
        self.kernel_name = ""
        self.update_kernel_name()

    # ------------------------------------------------------------------------------------------ objective
    def _latent_moments(self, u, q_mu, q_cov, X):
        """[(mean [n], var [n])] of every latent GP at X under q = N(q_mu, q_cov) (whitened; q_cov = q_sqrt q_sqrt^T)"""
        import torch
        vals = self._params.values(u)
        out = []
        for q, k in enumerate(self.kernel.kernels):
            Z = self.Z[q]
            Kuu = _kernel_matrix(k, Z, Z, vals) + JITTER * torch.eye(Z.shape[0], dtype=self.dtype, device=self.device)
            L = torch.linalg.cholesky(Kuu)
            A = torch.linalg.solve_triangular(L, _kernel_matrix(k, Z, X, vals), upper=False)          # [M, n]
            mean = A.T @ q_mu[q]
            var = _kernel_diag(k, X, vals) - (A * A).sum(0) + (A * (q_cov[q] @ A)).sum(0)
            out.append((mean, var))
        return out

    def _loss(self, u, W, raw_noise, c, q_mu, q_cov):
        """training_loss = -(elbo + log prior of W)"""
        import torch
        mom = self._latent_moments(u, q_mu, q_cov, self.X)
        G = torch.stack([m for m, _ in mom], 1)                    # [n, Q]
        S = torch.stack([v for _, v in mom], 1)
        fm = G @ W.T + c                                           # [n, P]
        fv = S @ (W * W).T
        if self.likelihood_name == "gaussian":
            s2 = _softplus(raw_noise) + 1e-6
            ve = -0.5 * math.log(2 * math.pi) - 0.5 * torch.log(s2) - 0.5 * ((self.Y - fm) ** 2 + fv) / s2
        else:                                                      # gpflow.likelihoods.Poisson, exp link
            ve = self.Y * fm - torch.exp(fm + 0.5 * fv) - torch.lgamma(self.Y + 1.0)
        kl = 0.0
        for q in range(len(q_mu)):
            kl = kl + 0.5 * ((q_mu[q] ** 2).sum() + torch.diagonal(q_cov[q]).sum() - len(q_mu[q])
                             - torch.linalg.slogdet(q_cov[q])[1])
        log_prior = _horseshoe_logp(W, self.w_prior_scale).sum()
        return -(ve.sum() - kl + log_prior)

    def _state(self):
        import torch
        W = torch.tensor(self.kernel.W, dtype=self.dtype, device=self.device)
        s2 = max(self.likelihood_variance - 1e-6, 1e-300)
        raw = torch.tensor(s2 + math.log(-math.expm1(-s2)), dtype=self.dtype, device=self.device)
        c = torch.tensor(self.mean_c, dtype=self.dtype, device=self.device)
        return self._params.u.clone(), W, raw, c

    def _q_cov(self):
        import torch
        return [torch.tril(sq) @ torch.tril(sq).T for sq in self.q_sqrt]

    def training_loss(self) -> float:
        import torch
        with torch.no_grad():
            u, W, raw, c = self._state()
            return float(self._loss(u, W, raw, c, self.q_mu, self._q_cov()))

    def elbo(self) -> float:
        import torch
        with torch.no_grad():
            W = torch.tensor(self.kernel.W, dtype=self.dtype, device=self.device)
            return -self.training_loss() - float(_horseshoe_logp(W, self.w_prior_scale).sum())

    # ------------------------------------------------------------------------------------------ optimiser
    def _natgrad_step(self, u, W, raw, c, gamma):
        """gpflow.optimizers.NaturalGradient(gamma) on every (q_mu_q, q_sqrt_q): theta <- theta - gamma dLoss/d eta with
        theta = (S^-1 m, -S^-1 / 2) the natural and eta = (m, S + m m^T) the expectation parameters.  The loss depends on
        q through (m, S) only, so  dLoss/d eta1 = dLoss/dm - 2 (dLoss/dS) m,  dLoss/d eta2 = dLoss/dS  (autograd gives the
        right-hand sides)."""
        import torch
        ms = [m.clone().requires_grad_(True) for m in self.q_mu]
        Ss = [S.clone().requires_grad_(True) for S in self._q_cov()]
        loss = self._loss(u, W, raw, c, ms, Ss)
        grads = torch.autograd.grad(loss, ms + Ss)
        Q = len(ms)
        new_mu, new_sqrt = [], []
        for q in range(Q):
            m, S = ms[q].detach(), Ss[q].detach()
            gS = 0.5 * (grads[Q + q] + grads[Q + q].T)
            g1, g2 = grads[q] - 2.0 * gS @ m, gS
            Sinv = torch.linalg.inv(S)
            t1, t2 = Sinv @ m - gamma * g1, -0.5 * Sinv - gamma * g2
            Sn = torch.linalg.inv(-2.0 * t2)
            Sn = 0.5 * (Sn + Sn.T)
            new_mu.append(Sn @ t1)
            new_sqrt.append(torch.linalg.cholesky(Sn))
        return new_mu, new_sqrt

    def optimize_params(self, adam_learning_rate=0.01, nat_gradient_gamma=0.1, num_opt_iter=2000, constraint_weight=1.0,
                        **unused):
        """:1473-1612 — per step a natural-gradient update of the variational parameters, then a legacy-Keras Adam step
        (gradients clipped to norm 1 per variable) on W, the lengthscales, the likelihood variance and the mean of
        loss + constraint_weight * sum(relu(-W[0, :])); checkpoint every 100 steps, restore on a failed factorisation or a
        non-finite loss, stop after 500 steps without improvement."""
        import torch
        u, W, raw, c = self._state()
        variables = [u, W, c] + ([raw] if self.likelihood_name == "gaussian" else [])
        ms = [torch.zeros_like(v) for v in variables]
        vs = [torch.zeros_like(v) for v in variables]
        best, no_improve, it = float("inf"), 0, 0
        snap = ([v.clone() for v in variables], [m.clone() for m in self.q_mu], [s.clone() for s in self.q_sqrt])
        history = []

        def restore():
            for v, sv in zip(variables, snap[0]):
                v.copy_(sv)
            self.q_mu, self.q_sqrt = [m.clone() for m in snap[1]], [s.clone() for s in snap[2]]
        for i in range(int(num_opt_iter)):
            try:
                nm, ns = self._natgrad_step(u, W, raw, c, nat_gradient_gamma)
                self.q_mu, self.q_sqrt = [t.detach() for t in nm], [t.detach() for t in ns]
                leaves = [v.clone().requires_grad_(True) for v in variables]
                lu, lW, lc = leaves[0], leaves[1], leaves[2]
                lraw = leaves[3] if self.likelihood_name == "gaussian" else raw
                loss = self._loss(lu, lW, lraw, lc, self.q_mu, self._q_cov())
                total = loss + constraint_weight * torch.relu(-lW[0, :]).sum()
                grads = torch.autograd.grad(total, leaves, allow_unused=True)
            except Exception as e:           # torch.linalg.LinAlgError: upstream's InvalidArgumentError branch
                if self.verbose:
                    print(f"Optimization failed at step {i} with error: {e}\nRestoring previous parameter values and stopping.")
                restore()
                break
            t = i + 1
            for j, (v, g) in enumerate(zip(variables, grads)):
                if g is None:
                    continue
                nrm = torch.linalg.norm(g)
                if nrm > 1.0:
                    g = g / nrm
                ms[j] = 0.9 * ms[j] + 0.1 * g
                vs[j] = 0.999 * vs[j] + 0.001 * g * g
                v -= adam_learning_rate * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * ms[j] / (torch.sqrt(vs[j]) + 1e-7)
            loss_val = float(loss.detach())
            history.append(loss_val)
            it = t
            if self.verbose and i % 500 == 0:
                print(f"Iteration {i}: Loss = {loss_val}, Total = {float(total.detach())}")
            if i % 100 == 0:
                snap = ([v.clone() for v in variables], [m.clone() for m in self.q_mu], [s.clone() for s in self.q_sqrt])
            if not math.isfinite(loss_val):
                restore()
                break
            if loss_val < best:
                best, no_improve = loss_val, 0
            else:
                no_improve += 1
                if no_improve >= 500:
                    break
        self._params.u = variables[0].detach().clone()
        self._params.write_back(self._params.u)
        self.kernel.W = variables[1].detach().cpu().numpy().copy()
        self.mean_c = float(variables[2])
        if self.likelihood_name == "gaussian":
            self.likelihood_variance = float(_softplus(variables[3]) + 1e-6)
        self.optimizer = "custom_multioutput"
        self.fit_info = dict(n_iter=it, loss=history[-1] if history else float("nan"), best_loss=best)
        self.update_kernel_name()
        return None

    # ------------------------------------------------------------------------------------------ post-fit
    def predict_f(self, Xnew):
        """(mean [m, P], var [m, P]) of f at new inputs (full_cov = False, full_output_cov = False)"""
        import torch
        with torch.no_grad():
            Xn = torch.tensor(np.asarray(Xnew, dtype=np.float64), dtype=self.dtype, device=self.device)
            mom = self._latent_moments(self._params.u, self.q_mu, self._q_cov(), Xn)
            W = torch.tensor(self.kernel.W, dtype=self.dtype, device=self.device)
            G = torch.stack([m for m, _ in mom], 1)
            S = torch.stack([v for _, v in mom], 1)
            return (G @ W.T + self.mean_c).cpu().numpy(), (S @ (W * W).T).cpu().numpy()

    def predict_y(self, Xnew):
        mu, var = self.predict_f(Xnew)
        if self.likelihood_name == "gaussian":
            return mu, var + self.likelihood_variance
        m = np.exp(mu + 0.5 * var)
        return m, m + m * m * np.expm1(var)

    def prune_latent_factors(self, threshold=0.1, variance_threshold=None, optimize_after_prune=True, optimize_kwargs=None):
        """:1388-1471 — drop latent factors whose largest |W| entry is below ``threshold`` (or whose kernel variance is
        below ``variance_threshold``), keep at least one, optionally re-optimise from the pruned state."""
        W = self.kernel.W
        importance = np.max(np.abs(W), axis=0)
        prune = importance < threshold
        if variance_threshold is not None:
            prune = np.logical_or(prune, np.array([float(k.variance) if hasattr(k, "variance") else 1.0
                                                   for k in self.kernel.kernels]) < variance_threshold)
        keep = np.where(~prune)[0]
        if len(keep) == 0:
            print("Warning: All latent factors would be pruned! Keeping the one with max weight.")
            keep = np.array([int(np.argmax(importance))])
        if len(keep) == W.shape[1]:
            if self.verbose:
                print("No latent factors pruned.")
            return
        if self.verbose:
            print(f"Pruning {W.shape[1] - len(keep)} latent factors. Keeping {len(keep)}.")
        self.kernel = _Coregion([self.kernel.kernels[i] for i in keep], W[:, keep])
        self.q_mu = [self.q_mu[i] for i in keep]
        self.q_sqrt = [self.q_sqrt[i] for i in keep]
        self.Z = [self.Z[i] for i in keep]
        self._params = _Params(self.kernel.kernels, self.dtype, self.device)
        self.update_kernel_name()
        if optimize_after_prune:
            kw = {"adam_learning_rate": 1e-3, "nat_gradient_gamma": 0.05, "num_opt_iter": 1000, "constraint_weight": 0.1}
            kw.update(optimize_kwargs or {})
            try:
                self.optimize_params(**kw)
            except Exception as e:          # upstream swallows the failure of the warm-start re-optimisation as well
                if self.verbose:
                    print(f"Warning: re-optimization after pruning failed: {e}")

    def update_kernel_name(self):
        from .utilities import kernel_name_string
        self.kernel_name = "+".join(kernel_name_string(k, with_idx=True) for k in self.kernel.kernels)
