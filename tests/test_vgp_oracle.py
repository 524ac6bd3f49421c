"""oracle/vgp_oracle.py pinned against itself: the collapsed site form equals the whitened gpflow-VGP ELBO
 at the q defined by the sites, that q is a maximiser, and dF/dtheta matches finite differences."""
import copy

import numpy as np
import pytest

import gp_oracle as go
import vgp_oracle as vo


def _setup(seed=0, n=40):
    rng = np.random.default_rng(seed)
    subj = np.repeat(np.arange(8), n // 8).astype(float)
    t = rng.normal(size=n)
    X = np.stack([subj, t], 1)
    f = 0.5 * rng.normal(size=8)[subj.astype(int)] + np.sin(2 * t) + 1.0
    y = rng.poisson(np.exp(f)).astype(float)
    kern = {"type": "sum", "kernels": [go.leaf("categorical", 0, variance=0.7),
                                       go.leaf("squared_exponential", 1, variance=1.3, lengthscales=0.8)]}
    model = go.gpr_model(kern, noise=1.0, mean="constant", c=0.3)
    model["likelihood_variance"]["trainable"] = False
    return model, X, y, go.pack(model), rng


@pytest.mark.parametrize("lik", [{"type": "poisson"}, {"type": "negative_binomial", "alpha": 0.7}])
def test_collapsed_equals_max_of_whitened_elbo(lik):
    model, X, y, x, rng = _setup()
    r = vo.vgp_collapsed(model, lik, X, y, x)
    q_mu, q_sqrt = vo.q_from_sites(model, X, y, x, r["sites"])
    e = vo.vgp_elbo(model, lik, X, y, x, q_mu, q_sqrt)
    assert abs(e - r["F"]) <= 1e-10 * abs(e)
    n = len(y)
    for _ in range(8):                                   # any perturbation of q lowers the bound
        dq, dS = 1e-3 * rng.normal(size=n), 1e-3 * np.tril(rng.normal(size=(n, n)))
        assert vo.vgp_elbo(model, lik, X, y, x, q_mu + dq, q_sqrt + dS) < e
    h, fd = 1e-5, []
    for i in range(len(x)):
        xp, xm = x.copy(), x.copy()
        xp[i] += h; xm[i] -= h
        fd.append((vo.vgp_collapsed(model, lik, X, y, xp, want_grad=False)["F"]
                   - vo.vgp_collapsed(model, lik, X, y, xm, want_grad=False)["F"]) / (2 * h))
    np.testing.assert_allclose(r["grad"], fd, rtol=1e-7)


def test_damping_does_not_change_the_fixed_point():
    model, X, y, x, _ = _setup(seed=3)
    a = vo.vgp_collapsed(model, {"type": "poisson"}, X, y, x, rho=1.0, want_grad=False)
    b = vo.vgp_collapsed(model, {"type": "poisson"}, X, y, x, rho=0.4, want_grad=False, maxit=2000)
    assert abs(a["F"] - b["F"]) <= 1e-10 * abs(a["F"])


@pytest.mark.parametrize("kind", ["bernoulli", "gamma"])
def test_bernoulli_and_gamma_collapsed_bounds(kind):
    """the same three checks for gp_likelihood_crosswalk's Bernoulli (inv_probit) and Gamma (exp link, trainable shape
    in the noise slot) entries"""
    model, X, y, x, rng = _setup(seed=5)
    if kind == "bernoulli":
        lik, yy = {"type": "bernoulli"}, (y > np.median(y)).astype(float)
    else:
        lik, yy = {"type": "gamma", "shape": 2.0}, rng.gamma(2.0, np.exp(0.3 * np.sin(X[:, 1])))
        model["likelihood_variance"] = {"value": 2.0, "trainable": True, "transform": "softplus", "prior": None}
        x = go.pack(model)
    kw = dict(rho=0.7, maxit=3000)
    r = vo.vgp_collapsed(model, lik, X, yy, x, **kw)
    q_mu, q_sqrt = vo.q_from_sites(model, X, yy, x, r["sites"])
    e = vo.vgp_elbo(model, lik, X, yy, x, q_mu, q_sqrt)
    assert abs(e - r["F"]) <= 1e-10 * abs(e)
    n = len(yy)
    for _ in range(8):
        dq, dS = 1e-3 * rng.normal(size=n), 1e-3 * np.tril(rng.normal(size=(n, n)))
        assert vo.vgp_elbo(model, lik, X, yy, x, q_mu + dq, q_sqrt + dS) < e
    h, fd = 1e-5, []
    for i in range(len(x)):
        xp, xm = x.copy(), x.copy()
        xp[i] += h; xm[i] -= h
        fd.append((vo.vgp_collapsed(model, lik, X, yy, xp, want_grad=False, **kw)["F"]
                   - vo.vgp_collapsed(model, lik, X, yy, xm, want_grad=False, **kw)["F"]) / (2 * h))
    np.testing.assert_allclose(r["grad"], fd, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("lik", [{"type": "poisson"}, {"type": "bernoulli"}])
def test_site_form_prediction_equals_whitened_conditional(lik):
    """predict_f at new inputs: the heteroscedastic-GPR form the engine evaluates == gpflow's whitened conditional"""
    model, X, y, x, rng = _setup(seed=7)
    yy = y if lik["type"] == "poisson" else (y > np.median(y)).astype(float)
    r = vo.vgp_collapsed(model, lik, X, yy, x, rho=0.7, maxit=3000, want_grad=False)
    Xnew = np.stack([rng.integers(0, 9, size=15).astype(float), rng.normal(size=15)], 1)      # incl. an unseen subject
    m1, v1 = vo.predict_f(model, X, yy, x, r["sites"], Xnew)
    m2, v2 = vo.predict_f(model, X, yy, x, r["sites"], Xnew, whitened=True)
    np.testing.assert_allclose(m1, m2, rtol=0, atol=1e-9)
    np.testing.assert_allclose(v1, v2, rtol=0, atol=1e-9)
    assert np.all(v1 > 0)


def zinb_sample(rng, f, alpha=0.5, km=1.0):
    """waveome/likelihoods.py:96-139 as a sampler: structural zero with probability km / (km + m), else NB(m, alpha)."""
    m = np.exp(f)
    k = 1.0 / alpha
    keep = rng.uniform(size=len(m)) < m / (km + m)
    return (keep * rng.negative_binomial(k, k / (k + m))).astype(float)


def test_zinb_collapsed_bound():
    """Zero-inflated negative binomial: the site form still equals the whitened ELBO at the q of the sites, and the
    gradient (incl. d/d alpha and d/d km through the second likelihood slot) matches finite differences.  The zero
    branch is not log-concave: max d2 log p / df2 > 0."""
    model, X, _, _, rng = _setup(seed=9)
    yz = zinb_sample(rng, 0.5 * np.sin(2 * X[:, 1]) + 1.0)
    assert np.sum(yz == 0) >= 8
    f = np.linspace(-3, 4, 29)
    assert vo.zinb_terms(f, np.zeros_like(f), 0.3, 0.5)[2].max() > 0.05
    model["likelihood_variance"] = {"value": 0.5, "trainable": True, "transform": "softplus", "prior": None}
    model["likelihood_aux"] = {"value": 1.2, "trainable": True, "transform": "softplus", "prior": None}
    x = go.pack(model)
    assert len(x) == 6
    lik = {"type": "zinb"}
    kw = dict(rho=0.5, maxit=5000)
    r = vo.vgp_collapsed(model, lik, X, yz, x, **kw)
    q_mu, q_sqrt = vo.q_from_sites(model, X, yz, x, r["sites"])
    e = vo.vgp_elbo(model, lik, X, yz, x, q_mu, q_sqrt)
    assert abs(e - r["F"]) <= 1e-10 * abs(e)
    h, fd = 1e-5, []
    for i in range(len(x)):
        xp, xm = x.copy(), x.copy()
        xp[i] += h; xm[i] -= h
        fd.append((vo.vgp_collapsed(model, lik, X, yz, xp, want_grad=False, **kw)["F"]
                   - vo.vgp_collapsed(model, lik, X, yz, xm, want_grad=False, **kw)["F"]) / (2 * h))
    np.testing.assert_allclose(r["grad"], fd, rtol=1e-6, atol=1e-7)


def test_objective_b_collapses_to_objective_a_for_the_gaussian_likelihood():
    """the reference's penalised path maximises the whitened SVGP ELBO with Z = X (objective B) over
    (theta, q_mu, q_sqrt).  For the Gaussian likelihood its maximum over q is the exact log marginal likelihood
    (objective A) with the noise variance increased by the 1e-6 jitter -- which is why the engine evaluates A."""
    model, X, y, _, rng = _setup(seed=2)
    y = np.log1p(y) + 0.1 * rng.normal(size=len(y))
    s2 = 0.37
    model["likelihood_variance"]["value"] = s2
    x = go.pack(model)
    lik = {"type": "gaussian", "variance": s2}
    r = vo.vgp_collapsed(model, lik, X, y, x, want_grad=False)
    assert r["iters"] <= 2                                   # the sites are exact after one sweep: lam = 1 / s2
    np.testing.assert_allclose(r["sites"][0], 1.0 / s2, rtol=1e-12)
    a_model = copy.deepcopy(model)
    a_model["likelihood_variance"]["value"] = s2 + vo.JITTER
    fa = go.objective(a_model, X, y, go.pack(a_model), want_grad=False)
    lml = fa[2] if isinstance(fa, tuple) else -fa
    assert abs(r["F"] - lml) <= 1e-10 * abs(lml), (r["F"], lml)
    q_mu, q_sqrt = vo.q_from_sites(model, X, y, x, r["sites"])
    e = vo.vgp_elbo(model, lik, X, y, x, q_mu, q_sqrt)
    assert abs(e - lml) <= 1e-10 * abs(lml)
    n = len(y)
    for _ in range(5):
        dq, dS = 1e-3 * rng.normal(size=n), 1e-3 * np.tril(rng.normal(size=(n, n)))
        assert vo.vgp_elbo(model, lik, X, y, x, q_mu + dq, q_sqrt + dS) < e


def test_collapsed_objective_with_hyperparameter_priors():
    """MAP objective of the VGP branches (waveome/model_fitting.py:236-242 puts Uniform(0, 10) on the non-variance kernel
    parameters; penalised models carry horseshoe priors on the variances): f = -(F + log prior), gradient checked by
    finite differences, and equal to the whitened ELBO + log prior at the optimal q."""
    model, X, y, x, rng = _setup(seed=3)
    kerns = model["kernel"]["kernels"]
    kerns[0]["params"]["variance"]["prior"] = {"type": "horseshoe", "scale": 0.5}
    kerns[1]["params"]["variance"]["prior"] = {"type": "laplace", "loc": 0.0, "scale": 2.0}
    kerns[1]["params"]["lengthscales"]["prior"] = {"type": "uniform", "low": 0.0, "high": 10.0}
    lik = {"type": "poisson"}
    r = vo.vgp_collapsed(model, lik, X, y, x)
    assert r["log_prior"] != 0.0 and abs(r["f"] + r["F"] + r["log_prior"]) <= 1e-12 * abs(r["f"])
    q_mu, q_sqrt = vo.q_from_sites(model, X, y, x, r["sites"])
    assert abs(vo.vgp_elbo(model, lik, X, y, x, q_mu, q_sqrt) + r["log_prior"] + r["f"]) <= 1e-10 * abs(r["f"])
    h, fd = 1e-5, []
    for i in range(len(x)):
        xp, xm = x.copy(), x.copy()
        xp[i] += h; xm[i] -= h
        fd.append((vo.vgp_collapsed(model, lik, X, y, xm, want_grad=False)["f"]
                   - vo.vgp_collapsed(model, lik, X, y, xp, want_grad=False)["f"]) / (2 * h))
    np.testing.assert_allclose(r["grad"], fd, rtol=2e-7)
    # the fit minimises f: it ends at a point where the prior terms are part of the stationarity condition
    ro = vo.fit(model, lik, X, y)
    assert abs(ro["f"] + ro["F"] + ro["log_prior"]) <= 1e-12 * abs(ro["f"]) and ro["f"] <= r["f"]
