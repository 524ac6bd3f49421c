"""The oracle against independent restatements: torch-fp64 autograd of the same objective, finite differences,
the exact horseshoe density, and the committed golden vectors."""
import copy
import json
import math
import os

import numpy as np
import pytest
import torch

import gp_oracle as oracle
import helpers
import waveome_b200 as wb

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---------------------------------------------------------------- independent torch restatement (autograd)
def _t_softplus(u):
    return torch.clamp(u, min=0) + torch.log1p(torch.exp(-torch.abs(u)))


def _t_param(p, u):
    t = p["transform"]
    if t == "softplus":
        return _t_softplus(u)
    if t == "softplus_shift":
        return _t_softplus(u) + p["shift"]
    if t == "exp":
        return torch.exp(u)
    return u


def _t_horseshoe(x, s):
    g, b, h_inf, pw = 0.5614594835668851, 1.0420764938351215, 1.0801359952503342, 1.0919284281983377
    xx = (x / s) ** 2 / 2
    q = 20.0 / 47.0 * xx ** pw
    h = 1.0 / (1 + xx ** 1.5) + h_inf * q / (1 + q)
    c = -0.5 * math.log(2 * math.pi ** 3) - math.log(g * s)
    z = math.log1p(-g) - math.log(g)
    return -torch.nn.functional.softplus(z - xx / (1 - g)) + torch.log(torch.log1p(g / xx - (1 - g) / (h + b * xx) ** 2)) + c


def _t_kernel(node, X, vals):
    typ = node["type"]
    if typ == "sum":
        return sum(_t_kernel(c, X, vals) for c in node["kernels"])
    if typ == "product":
        out = None
        for c in node["kernels"]:
            k = _t_kernel(c, X, vals)
            out = k if out is None else out * k
        return out
    x = X[:, node["dim"]]
    P = {name: vals[id(p)] for name, p in node["params"].items()}
    if typ == "empty":
        return torch.zeros(len(x), len(x), dtype=torch.float64)
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
        return P["variance"] * torch.ones(len(x), len(x), dtype=torch.float64)
    if typ == "categorical":
        c = torch.round(x)
        return P["variance"] * (c[:, None] == c[None, :]).to(torch.float64)
    if typ in ("poly", "polynomial"):
        return (P["variance"] * torch.outer(x, x) + P["offset"]) ** node.get("degree", 3)
    raise ValueError(typ)


def torch_objective(spec, X, y, x):
    spec = copy.deepcopy(spec)
    tp = oracle.trainable_params(spec)
    u = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    vals = {}
    for p in oracle.iter_model_params(spec):
        vals[id(p)] = torch.tensor(p["value"], dtype=torch.float64)
    for i, p in enumerate(tp):
        vals[id(p)] = _t_param(p, u[i])
    Xt = torch.tensor(X, dtype=torch.float64)
    yt = torch.tensor(y, dtype=torch.float64)
    K = _t_kernel(spec["kernel"], Xt, vals)
    n = len(y)
    Ks = K + vals[id(spec["likelihood_variance"])] * torch.eye(n, dtype=torch.float64)
    L = torch.linalg.cholesky(Ks)
    c = vals[id(spec["mean"]["c"])] if spec["mean"]["type"] == "constant" else 0.0
    a = torch.linalg.solve_triangular(L, (yt - c)[:, None], upper=False)
    lml = -0.5 * (a * a).sum() - 0.5 * n * math.log(2 * math.pi) - torch.log(torch.diagonal(L)).sum()
    lp = 0.0
    for p in tp:
        pr = p.get("prior")
        if pr is None:
            continue
        v = vals[id(p)]
        if pr["type"] == "horseshoe":
            lp = lp + _t_horseshoe(v, pr["scale"])
        elif pr["type"] == "laplace":
            lp = lp - torch.abs(v - pr["loc"]) / pr["scale"] - math.log(2 * pr["scale"])
        elif pr["type"] == "uniform":
            lp = lp - math.log(pr["high"] - pr["low"])
    f = -(lml + lp)
    f.backward()
    return float(f), u.grad.numpy().copy(), float(lml)


@pytest.mark.parametrize("n,kern", [(40, "all"), (75, "all"), (120, "sat")])
def test_oracle_vs_torch_autograd(n, kern):
    X, y = helpers.make_data(n, seed=n)
    k = helpers.all_leaf_kernel() if kern == "all" else helpers.saturated_kernel()
    spec = wb.GPR(k, mean_function=wb.ConstantMean(0.2), noise_variance=0.4).to_spec()
    x = oracle.pack(spec) + 0.25 * np.random.default_rng(n).normal(size=len(oracle.pack(spec)))
    f, g, lml, _ = oracle.objective(copy.deepcopy(spec), X, y, x)
    ft, gt, lt = torch_objective(spec, X, y, x)
    assert abs(f - ft) <= 1e-11 * abs(ft)
    assert abs(lml - lt) <= 1e-11 * abs(lt)
    np.testing.assert_allclose(g, gt, rtol=1e-9, atol=1e-11 * np.max(np.abs(gt)))


def test_oracle_priors_laplace_uniform_and_zero_mean():
    X, y = helpers.make_data(50, seed=5)
    k = wb.SquaredExponential(active_dims=[1]) + wb.Categorical(active_dims=[0])
    k.kernels[0].variance.prior = wb.Laplace(0.0, 0.5)
    k.kernels[0].lengthscales.prior = wb.Uniform(0.0, 10.0)
    spec = wb.GPR(k).to_spec()
    assert spec["mean"]["type"] == "zero"
    x = oracle.pack(spec) + 0.1
    f, g, lml, lp = oracle.objective(copy.deepcopy(spec), X, y, x)
    ft, gt, lt = torch_objective(spec, X, y, x)
    assert abs(f - ft) <= 1e-11 * abs(ft)
    np.testing.assert_allclose(g, gt, rtol=1e-9)


def test_horseshoe_matches_exact_density():
    """TFP's closed form is an approximation of -0.5 log(2 pi^3) - log s + xx + log E1(xx)."""
    from scipy.special import exp1
    for s in (0.1, 1.0, 10.0):
        for x in np.logspace(-4, 2, 25):
            xx = (x / s) ** 2 / 2
            exact = -0.5 * np.log(2 * np.pi ** 3) - np.log(s) + xx + np.log(exp1(xx)) if xx < 500 else None
            lp, dlp = oracle.horseshoe_logp_and_grad(x, s)
            if exact is not None and np.isfinite(exact):
                assert abs(lp - exact) < 1e-3
            h = 1e-6 * x
            fd = (oracle.horseshoe_logp_and_grad(x + h, s)[0] - oracle.horseshoe_logp_and_grad(x - h, s)[0]) / (2 * h)
            assert abs(dlp - fd) <= 1e-5 * abs(fd)


def test_bic_and_selection_arithmetic():
    assert oracle.calc_bic(-10.0, 100, 4) == 28.0
    k = wb.SquaredExponential(active_dims=[1]) + wb.Categorical(active_dims=[0])
    m = wb.GPR(k, mean_function=wb.ConstantMean())
    assert oracle.count_trainable_parameter_objects(m.to_spec()) == len(m.trainable_parameters) == 5


def test_golden_vectors():
    """Committed oracle outputs (tests/golden/make_golden.py) — guards the oracle against silent drift."""
    with open(os.path.join(GOLDEN, "eval_cases.json")) as fh:
        cases = json.load(fh)
    assert len(cases) >= 4
    for c in cases:
        X, y, x = np.array(c["X"]), np.array(c["y"]), np.array(c["x"])
        f, g, lml, _ = oracle.objective(copy.deepcopy(c["spec"]), X, y, x)
        assert abs(f - c["f"]) <= 1e-12 * abs(c["f"])
        assert abs(lml - c["lml"]) <= 1e-12 * abs(c["lml"])
        np.testing.assert_allclose(g, np.array(c["grad"]), rtol=1e-10, atol=1e-12)
