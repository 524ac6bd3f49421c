"""CPU oracle of objective (B): the whitened SVGP bound with Z = X that the live reference API fits, and the
reference's default optimiser for it.  TEST INFRASTRUCTURE ONLY (same rules as gp_oracle.py: nothing under
``waveome_b200/`` imports it).  torch fp64 on the CPU, gradients by autograd.

PARITY: unpinned by reference-held values except through gp_oracle (for a Gaussian likelihood max_q B = the exact-GPR
LML with sigma^2 + 1e-6 jitter on K, checked in tests/test_vgp_oracle.py, and gp_oracle is pinned on GPflow-recorded
notebook outputs).  TensorFlow / GPflow cannot be installed here; restated from

* the model: ``PSVGP`` -> ``PenalizedGP`` -> ``VarGP`` -> ``SparseGP`` -> ``BaseGP(gpflow.models.SVGP)``,
  waveome/model_classes.py:33-169, 622-676, 690-815, 1082-1126 (``inducing_variable = InducingPoints(X)`` :100, frozen
  :169, whitened, ``q_mu = 0``, ``q_sqrt = I``), bound: the in-repo mirror waveome/model_types_DEPR.py:126-158 of
  ``gpflow.models.SVGP.elbo``  = sum_i E_q[log p(y_i | f_i)] - KL[q(v) || N(0, I)],  f = L v + c,  L L^T = K + 1e-6 I;
* the objective handed to the optimiser: ``training_loss`` = -(elbo + log prior) with the horseshoe prior on the
  trainable kernel variances (waveome/model_classes.py:817-864);
* the optimiser: ``BaseGP.optimize_params`` "adam/gradient" branch, waveome/model_classes.py:344-462 — per step ONE
  Keras-Adam update (learning_rate 0.1) of the unconstrained hyper-parameters, then ONE
  ``gpflow.optimizers.NaturalGradient(gamma=0.1)`` update of (q_mu, q_sqrt); every 100 steps a checkpoint (loss, snapshot,
  every 500 steps learning_rate = 0.1 * 0.96^(i/500)); stop when the loss fell by < 1e-9 between checkpoints.

For the Gaussian likelihood the natural gradient of the bound with respect to the expectation parameters of q is
theta* - theta (theta = natural parameters of q, theta* those of the optimal q for the current hyper-parameters), so
NaturalGradient(gamma) is the relaxation  theta <- theta + gamma (theta* - theta)  with
theta1* = L^T (y - c) / sigma^2,  theta2* = -1/2 (I + L^T L / sigma^2)   (whitened parameterisation).
"""
from __future__ import annotations

import copy
import math

import numpy as np
import torch

import gp_oracle as go

JITTER = 1e-6
DT = torch.float64


# ---------------------------------------------------------------------------------------------- kernel tree in torch
def t_softplus(u):
    return torch.clamp(u, min=0) + torch.log1p(torch.exp(-torch.abs(u)))


def t_param(p, u):
    t = p["transform"]
    if t == "softplus":
        return t_softplus(u)
    if t == "softplus_shift":
        return t_softplus(u) + p["shift"]
    if t == "exp":
        return torch.exp(u)
    return u


def t_horseshoe(x, s):
    """tfd.Horseshoe(scale=s).log_prob(x): TFP's closed-form approximation"""
    g, b, h_inf, pw = 0.5614594835668851, 1.0420764938351215, 1.0801359952503342, 1.0919284281983377
    xx = (x / s) ** 2 / 2
    q = 20.0 / 47.0 * xx ** pw
    h = 1.0 / (1 + xx ** 1.5) + h_inf * q / (1 + q)
    c = -0.5 * math.log(2 * math.pi ** 3) - math.log(g * s)
    z = math.log1p(-g) - math.log(g)
    return -torch.nn.functional.softplus(z - xx / (1 - g)) + torch.log(torch.log1p(g / xx - (1 - g) / (h + b * xx) ** 2)) + c


def t_kernel(node, X, vals):
    typ = node["type"]
    if typ == "sum":
        return sum(t_kernel(c, X, vals) for c in node["kernels"])
    if typ == "product":
        out = None
        for c in node["kernels"]:
            k = t_kernel(c, X, vals)
            out = k if out is None else out * k
        return out
    x = X[:, node["dim"]]
    P = {name: vals[id(p)] for name, p in node["params"].items()}
    if typ == "empty":
        return torch.zeros(len(x), len(x), dtype=DT)
    if typ in ("squared_exponential", "matern12", "matern32", "matern52"):
        a = x / P["lengthscales"]
        r2 = -2 * torch.outer(a, a) + (a * a)[:, None] + (a * a)[None, :]
        if typ == "squared_exponential":
            return P["variance"] * torch.exp(-0.5 * r2)
        r = torch.sqrt(torch.clamp(r2, min=1e-36))
        if typ == "matern12":
            return P["variance"] * torch.exp(-r)
        if typ == "matern32":
            return P["variance"] * (1 + math.sqrt(3) * r) * torch.exp(-math.sqrt(3) * r)
        return P["variance"] * (1 + math.sqrt(5) * r + 5.0 / 3.0 * r * r) * torch.exp(-math.sqrt(5) * r)
    if typ == "periodic":
        d = x[:, None] - x[None, :]
        s = torch.sin(math.pi * d / P["period"]) / P["lengthscales"]
        return P["variance"] * torch.exp(-0.5 * s * s)
    if typ in ("linear", "lin"):
        return P["variance"] * torch.outer(x, x)
    if typ == "constant":
        return P["variance"] * torch.ones(len(x), len(x), dtype=DT)
    if typ == "categorical":
        c = torch.round(x)
        return P["variance"] * (c[:, None] == c[None, :]).to(DT)
    if typ in ("poly", "polynomial"):
        return (P["variance"] * torch.outer(x, x) + P["offset"]) ** node.get("degree", 3)
    raise ValueError(typ)


def _values(spec, u):
    """{id(param dict): torch value} for every parameter of the model, trainable ones as functions of u"""
    vals = {id(p): torch.tensor(p["value"], dtype=DT) for p in go.iter_model_params(spec)}
    for i, p in enumerate(go.trainable_params(spec)):
        vals[id(p)] = t_param(p, u[i])
    return vals


def _log_prior(spec, vals):
    lp = torch.zeros((), dtype=DT)
    for p in go.trainable_params(spec):
        pr = p.get("prior")
        if pr is None:
            continue
        v = vals[id(p)]
        if pr["type"] == "horseshoe":
            lp = lp + t_horseshoe(v, pr["scale"])
        elif pr["type"] == "laplace":
            lp = lp - torch.abs(v - pr["loc"]) / pr["scale"] - math.log(2 * pr["scale"])
        elif pr["type"] == "uniform":
            lp = lp - math.log(pr["high"] - pr["low"])
    return lp


# ---------------------------------------------------------------------------------------------- objective (B)
def svgp_loss(spec, Xt, yt, u, q_mu, q_sqrt):
    """training_loss = -(ELBO + log prior) of the whitened SVGP with Z = X and a Gaussian likelihood (torch scalar)."""
    vals = _values(spec, u)
    n = yt.shape[0]
    K = t_kernel(spec["kernel"], Xt, vals) + JITTER * torch.eye(n, dtype=DT)
    L = torch.linalg.cholesky(K)
    c = vals[id(spec["mean"]["c"])] if spec["mean"]["type"] == "constant" else 0.0
    s2 = vals[id(spec["likelihood_variance"])]
    fm = L @ q_mu + c
    LS = L @ torch.tril(q_sqrt)
    fv = (LS * LS).sum(1)
    ve = -0.5 * math.log(2 * math.pi) - 0.5 * torch.log(s2) - 0.5 * ((yt - fm) ** 2 + fv) / s2
    kl = 0.5 * ((q_mu ** 2).sum() + (torch.tril(q_sqrt) ** 2).sum() - n - 2.0 * torch.log(torch.abs(torch.diagonal(q_sqrt))).sum())
    return -(ve.sum() - kl + _log_prior(spec, vals))


def optimal_natural_parameters(spec, Xt, yt, u):
    """theta1*, theta2* of the optimal whitened q for the hyper-parameters u (Gaussian likelihood)"""
    with torch.no_grad():
        vals = _values(spec, u)
        n = yt.shape[0]
        K = t_kernel(spec["kernel"], Xt, vals) + JITTER * torch.eye(n, dtype=DT)
        L = torch.linalg.cholesky(K)
        c = vals[id(spec["mean"]["c"])] if spec["mean"]["type"] == "constant" else 0.0
        s2 = vals[id(spec["likelihood_variance"])]
        return L.T @ (yt - c) / s2, -0.5 * (torch.eye(n, dtype=DT) + L.T @ L / s2)


def q_from_natural(t1, t2):
    S = torch.linalg.inv(-2.0 * t2)
    S = 0.5 * (S + S.T)
    return S @ t1, torch.linalg.cholesky(S)


def fit_adam_natgrad(spec, X, y, learning_rate=0.1, decay_rate=0.96, gamma=0.1, max_iter=50000, threshold=1e-9,
                     check_every=100, decay_every=500, beta1=0.9, beta2=0.999, eps=1e-7, verbose=False):
    """BaseGP.optimize_params(optimizer="adam/gradient") on objective (B).  Returns dict(x, loss, n_iter, why, q_mu,
    q_sqrt, model) with x the unconstrained hyper-parameters and ``model`` the spec at the fitted values."""
    spec = copy.deepcopy(spec)
    Xt, yt = torch.tensor(np.asarray(X), dtype=DT), torch.tensor(np.asarray(y).reshape(-1), dtype=DT)
    n = len(yt)
    u = torch.tensor(go.pack(spec), dtype=DT)
    t1, t2 = torch.zeros(n, dtype=DT), -0.5 * torch.eye(n, dtype=DT)        # q_mu = 0, q_sqrt = I
    q_mu, q_sqrt = torch.zeros(n, dtype=DT), torch.eye(n, dtype=DT)
    m, v = torch.zeros_like(u), torch.zeros_like(u)
    lr = learning_rate
    losses = []
    prev = (u.clone(), t1.clone(), t2.clone())
    why, i = "maxiter", -1
    for i in range(max_iter):
        try:
            uu = u.clone().requires_grad_(True)
            loss = svgp_loss(spec, Xt, yt, uu, q_mu, q_sqrt)
            (g,) = torch.autograd.grad(loss, uu)
            t = i + 1
            m = beta1 * m + (1 - beta1) * g
            v = beta2 * v + (1 - beta2) * g * g
            u = u - lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t) * m / (torch.sqrt(v) + eps)
            o1, o2 = optimal_natural_parameters(spec, Xt, yt, u)
            t1, t2 = t1 + gamma * (o1 - t1), t2 + gamma * (o2 - t2)
            q_mu, q_sqrt = q_from_natural(t1, t2)
        except Exception:                       # torch.linalg.LinAlgError: TensorFlow's InvalidArgumentError
            u, t1, t2 = prev
            q_mu, q_sqrt = q_from_natural(t1, t2)
            why = "restored"
            break
        if i % check_every == 0:
            prev = (u.clone(), t1.clone(), t2.clone())
            with torch.no_grad():
                cur = float(svgp_loss(spec, Xt, yt, u, q_mu, q_sqrt))
            if math.isnan(cur):
                why = "nan"
                break
            losses.append(cur)
            if i % decay_every == 0:
                lr = learning_rate * decay_rate ** (i / decay_every)
            if verbose:
                print(i, cur, lr, flush=True)
            if len(losses) > 1 and losses[-2] - losses[-1] < threshold:
                why = "converged"
                break
    x = u.numpy().copy()
    go.unpack(spec, x)
    return dict(x=x, loss=losses[-1] if losses else float("nan"), n_iter=i + 1, why=why, q_mu=q_mu.numpy(),
                q_sqrt=q_sqrt.numpy(), model=spec)


# ---------------------------------------------------------------------------------------------- Adam on objective (A)
def fit_adam_collapsed(spec, X, y, learning_rate=0.1, decay_rate=0.96, max_iter=50000, threshold=1e-9, check_every=100,
                       decay_every=500, beta1=0.9, beta2=0.999, eps=1e-7):
    """The same schedule on the collapsed objective (gp_oracle.objective, NumPy): what ``wv_batch_fit_adam`` runs on the
    device.  A Cholesky failure or a NaN checkpoint loss restores the last checkpoint.  Returns dict(x, f, n_iter, why)."""
    spec = copy.deepcopy(spec)
    x = go.pack(spec)
    m, v = np.zeros_like(x), np.zeros_like(x)
    lr, prev_loss, n_loss, xprev = learning_rate, 0.0, 0, x.copy()
    why, it = "maxiter", 0
    while True:
        try:
            f, g, _, _ = go.objective(copy.deepcopy(spec), X, y, x)
        except go.CholeskyFailure:
            x, why = xprev.copy(), "restored"
            break
        i = it
        if i >= 1 and (i - 1) % check_every == 0:
            if math.isnan(f):           # engine policy: return the last finite checkpoint (upstream keeps the NaN values)
                x, why = xprev.copy(), "nan"
                break
            xprev = x.copy()
            if (i - 1) % decay_every == 0:
                lr = learning_rate * decay_rate ** ((i - 1) / decay_every)
            if n_loss >= 1 and prev_loss - f < threshold:
                why = "converged"
                break
            prev_loss, n_loss = f, n_loss + 1
        if i >= max_iter:
            break
        t = i + 1
        m = beta1 * m + (1 - beta1) * g
        v = beta2 * v + (1 - beta2) * g * g
        x = x - lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t) * m / (np.sqrt(v) + eps)
        it = i + 1
    f, _, lml, _ = go.objective(copy.deepcopy(spec), X, y, x, want_grad=False)
    return dict(x=x, f=float(f), lml=float(lml), n_iter=it, why=why)
