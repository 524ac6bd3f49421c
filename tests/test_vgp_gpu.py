"""Count likelihoods on the engine (BASELINE configs[4]): the collapsed variational bound F(theta) = max_q ELBO and its
gradient vs oracle/vgp_oracle.py (which is itself checked against the whitened gpflow-VGP ELBO in test_vgp_oracle.py)."""
import copy

import numpy as np
import pytest

import gp_oracle as go
import vgp_oracle as vo
import waveome_b200 as wb

pytestmark = pytest.mark.gpu


def count_data(n, n_subj, seed, scale=1.0, offset=1.0):
    rng = np.random.default_rng(seed)
    subj = rng.integers(0, n_subj, size=n).astype(float)
    t = rng.normal(size=n)
    X = np.stack([subj, t], 1)
    f = scale * (0.6 * rng.normal(size=n_subj)[subj.astype(int)] + np.sin(2 * t)) + offset
    return X, rng.poisson(np.exp(f)).astype(float)


def count_model(var=1.0, ls=1.0, c=0.2):
    k = wb.Sum([wb.Categorical(active_dims=[0], variance=0.7 * var), wb.SquaredExponential(active_dims=[1], variance=var, lengthscales=ls)])
    m = wb.GPR(k, mean_function=wb.ConstantMean(c))
    wb.set_trainable(m.likelihood.variance, False)          # no Gaussian noise on this path
    return m


@pytest.mark.parametrize("lik,param", [("poisson", 0.0), ("negative_binomial", 0.7)])
@pytest.mark.parametrize("n", [40, 130, 333])
def test_collapsed_bound_and_gradient_match_oracle(engine, lik, param, n):
    from waveome_b200.engine import Batch
    X, y = count_data(n, max(4, n // 6), seed=n)
    model = count_model()
    rng = np.random.default_rng(n + 1)
    Y = np.stack([y, rng.poisson(3.0, size=n).astype(float)])
    batch = Batch(engine, X, Y, [model.program()])
    batch.set_likelihood(lik, param)
    x = batch.x0() + 0.2 * rng.normal(size=(2, batch.P))
    f, g, lml, st = batch.eval(x)
    fm, fv = batch.latent()
    olik = {"type": lik, "alpha": param}
    for b in range(2):
        r = vo.vgp_collapsed(copy.deepcopy(model.to_spec()), olik, X, Y[b], x[b], rho=0.5, tol=1e-12, maxit=2000)
        assert st[b] == 0
        assert abs(lml[b] - r["F"]) <= 1e-8 * abs(r["F"]), (lml[b], r["F"])
        assert abs(f[b] + r["F"]) <= 1e-8 * abs(r["F"])
        np.testing.assert_allclose(g[b], -r["grad"], rtol=0, atol=1e-6 * np.max(np.abs(r["grad"])))
        np.testing.assert_allclose(fm[b], r["m"], rtol=0, atol=1e-7 * (1 + np.max(np.abs(r["m"]))))
        np.testing.assert_allclose(fv[b], r["v"], rtol=1e-6, atol=1e-9)
    # warm start: a second evaluation at the same point is already converged after one sweep and gives the same numbers
    c0 = batch.counters()
    f2, g2, lml2, _ = batch.eval(x)
    np.testing.assert_allclose(lml2, lml, rtol=1e-10)
    np.testing.assert_allclose(g2, g, rtol=0, atol=1e-7 * np.max(np.abs(g)))
    batch.close()


def test_trainable_dispersion_gradient(engine):
    """waveome's NegativeBinomial trains alpha (Exp bijector, likelihoods.py:24-28): it rides in the program's noise slot
    and receives d(bound)/d(alpha) = sum_i dE_i/d(alpha)."""
    from waveome_b200.engine import Batch
    from waveome_b200.models import NegativeBinomial
    X, y = count_data(90, 15, seed=21)
    model = count_model()
    model.likelihood = NegativeBinomial(alpha=0.6)
    p = model.program()
    assert p.n_x == 5 and p.slot_xindex[p.noise_slot] >= 0
    batch = Batch(engine, X, y[None, :], [p])
    batch.set_likelihood("negative_binomial", 123.0)          # ignored: the slot is trainable
    x = batch.x0() + 0.1
    f, g, lml, st = batch.eval(x)
    r = vo.vgp_collapsed(copy.deepcopy(model.to_spec()), {"type": "negative_binomial", "alpha": 123.0}, X, y, x[0],
                         rho=0.5, tol=1e-12, maxit=2000)
    assert st[0] == 0 and abs(lml[0] - r["F"]) <= 1e-8 * abs(r["F"])
    np.testing.assert_allclose(g[0], -r["grad"], rtol=0, atol=1e-6 * np.max(np.abs(r["grad"])))
    batch.close()


def test_hard_regimes_converge(engine):
    """huge counts under a tight prior (the plain fixed point overflows on its first move) and low counts under a wide
    prior (it oscillates): the safeguarded iteration must still reach the oracle's optimum"""
    from waveome_b200.engine import Batch
    for (scale, off, var, ls) in [(2.5, 5.0, 0.05, 2.0), (1.5, 3.0, 5.0, 0.3), (1.0, -2.0, 5.0, 0.3)]:
        X, y = count_data(100, 20, seed=7, scale=scale, offset=off)
        model = count_model(var=var, ls=ls, c=0.0)
        batch = Batch(engine, X, y[None, :], [model.program()])
        batch.set_likelihood("poisson")
        x = batch.x0()
        f, g, lml, st = batch.eval(x)
        r = vo.vgp_collapsed(copy.deepcopy(model.to_spec()), {"type": "poisson"}, X, y, x[0], rho=0.3, tol=1e-12, maxit=5000)
        assert st[0] == 0, (scale, off, var, ls, st)
        assert abs(lml[0] - r["F"]) <= 1e-7 * abs(r["F"]), (scale, off, var, ls, lml[0], r["F"])
        batch.close()


def test_sweep_cap_is_flagged_and_bound_stays_valid(engine):
    """independent-looking latent values with a huge prior variance and counts in {0, 1, 2}: the site map has a
    strongly oscillatory mode and the iteration creeps.  At the sweep cap the evaluation must come back finite, with a
    value that is a lower bound of (and close to) the optimum, and with status bit 16 unless it got close."""
    from waveome_b200.engine import Batch
    X, y = count_data(100, 20, seed=7, scale=1.0, offset=-2.0)
    model = count_model(var=30.0, ls=0.1, c=0.0)
    batch = Batch(engine, X, y[None, :], [model.program()])
    batch.set_likelihood("poisson")
    f, g, lml, st = batch.eval(batch.x0())
    r = vo.vgp_collapsed(copy.deepcopy(model.to_spec()), {"type": "poisson"}, X, y, batch.x0()[0], rho=0.2, tol=1e-10, maxit=20000)
    assert st[0] in (0, 16) and np.isfinite(f[0]) and np.all(np.isfinite(g))
    assert lml[0] <= r["F"] + 1e-8 and r["F"] - lml[0] < 1e-3 * abs(r["F"])
    batch.close()


def test_fit_matches_oracle_lbfgs(engine):
    """MAP fit of the hyper-parameters on the collapsed bound: device L-BFGS-B vs SciPy L-BFGS-B on the oracle."""
    from waveome_b200.model_fitting import fit_models
    from waveome_b200.models import make_likelihood
    X, y = count_data(80, 16, seed=11)
    rng = np.random.default_rng(5)
    Y = np.stack([y, rng.poisson(np.exp(0.8 + np.sin(X[:, 1]))).astype(float)])
    for lik, olik in (("poisson", {"type": "poisson"}), ("negative_binomial", {"type": "negative_binomial", "alpha": 1.0})):
        models = []
        for _ in range(2):
            m = count_model()
            m.likelihood = make_likelihood(lik)
            models.append(m)
        spec = copy.deepcopy(models[0].to_spec())
        res = fit_models(X, Y, models, engine=engine)
        for b in range(2):
            ro = vo.fit(copy.deepcopy(spec), olik, X, Y[b])
            assert res["status"][b] in (0, 8), res["status"]
            assert abs(res["lml"][b] - ro["F"]) <= 1e-6 * abs(ro["F"]), (lik, b, res["lml"][b], ro["F"], res["n_iter"][b], ro["nit"])
            np.testing.assert_allclose(res["x"][b][: len(ro["x"])], ro["x"], rtol=2e-3, atol=2e-3)


def test_penalized_optimization_poisson_config5_shape():
    """BASELINE configs[4] shape at a reduced outcome count: n = 500 (50 subjects x 10 times), Poisson outcomes, model
    categorical[subject] + squared_exponential[time] (+ product) with horseshoe penalties."""
    from waveome_b200 import datasets
    from waveome_b200.model_search import GPSearch
    X, Y = datasets.count_microbiome(n_outcomes=12)
    gps = GPSearch(X, Y, unit_col="subject", outcome_likelihood="poisson")
    gps.penalized_optimization()
    assert len(gps.models) == 12
    for o, m in gps.models.items():
        assert m.likelihood.name == "poisson" and np.isfinite(m.log_posterior_density_value)
        assert "categorical[0]" in m.kernel_name            # every taxon has a subject effect of sd 0.5
    assert np.mean(["squared_exponential[1]" in m.kernel_name for m in gps.models.values()]) >= 0.5
    # feature importances with the Poisson deviance (utilities.py:553-558): one entry per component + residual
    for m in gps.models.values():
        ncomp = len(m.kernel.kernels) if m.kernel.name == "sum" else 1
        assert len(m.feature_importances) == ncomp + 1 and 0.0 <= m.feature_importances[-1] <= 1.0
        assert all(np.isfinite(v) for v in m.feature_importances)


def test_predictive_variance_and_iterated_factor_for_counts(engine):
    """penalization_factor=None (waveome/model_search.py:271-375) with a count likelihood: the factor iteration reads
    sqrt(mean(predict_y(X)[1])); predict_y's variance = the likelihood's predict_mean_and_var of the latent posterior,
    checked against the oracle's q(f) + 20-point Gauss-Hermite moments."""
    from scipy.stats import norm
    from waveome_b200 import datasets, postfit
    from waveome_b200.model_search import GPSearch
    from waveome_b200.models import make_likelihood
    X, y = count_data(70, 10, seed=23)
    for lik, olik in (("poisson", {"type": "poisson"}), ("negative_binomial", {"type": "negative_binomial", "alpha": 1.0})):
        m = count_model()
        m.likelihood = make_likelihood(lik)
        spec = copy.deepcopy(m.to_spec())
        var_y = postfit.train_predictive_variance(X, y[None, :], [m], engine=engine)[0]
        r = vo.vgp_collapsed(copy.deepcopy(spec), olik, X, y, go.pack(spec), tol=1e-13, maxit=2000, want_grad=False)
        ref = vo.predict_y_moments(vo.lik_of(spec, olik), r["m"], r["v"])[1]
        np.testing.assert_allclose(var_y, ref, rtol=1e-6, atol=1e-9)
    Xd, Yd = datasets.count_microbiome(n_outcomes=4, n_subjects=12, n_times=6)
    gps = GPSearch(Xd, Yd, unit_col="subject", outcome_likelihood="poisson")
    gps.penalized_optimization(penalization_factor=None, num_factor_iter=2)
    assert gps.iterating_penalization_factor is True
    from waveome_b200.regularization import full_kernel_build
    from waveome_b200.utilities import find_variance_components
    fk, _ = full_kernel_build(cat_vars=gps.cat_idx, num_vars=gps.cont_idx, unit_idx=gps.unit_idx, var_names=gps.feat_names,
                              return_sum=True, second_order_numeric=False, categorical_numeric_interactions=True,
                              unit_numeric_interactions=False, kerns=[wb.SquaredExponential()])
    p = len(find_variance_components(fk, sum_reduce=False))
    for o, m in gps.models.items():
        start = 2 * 1.1 * np.std(gps.Y[o].to_numpy()) * np.sqrt(len(gps.X)) * norm().ppf(1 - 0.1 / (2 * p))
        assert m.likelihood.name == "poisson" and 0 < m.penalization_factor <= start + 1e-9
        assert np.isfinite(m.log_posterior_density_value)


def other_data(n, n_subj, seed, lik):
    rng = np.random.default_rng(seed)
    subj = rng.integers(0, n_subj, size=n).astype(float)
    t = rng.normal(size=n)
    X = np.stack([subj, t], 1)
    f = 0.6 * rng.normal(size=n_subj)[subj.astype(int)] + np.sin(2 * t)
    if lik == "bernoulli":
        return X, (rng.uniform(size=n) < 0.5 * (1 + np.tanh(f))).astype(float)
    return X, rng.gamma(2.0, np.exp(f))


@pytest.mark.parametrize("lik,param", [("bernoulli", 0.0), ("gamma", 1.7)])
@pytest.mark.parametrize("n", [50, 200])
def test_bernoulli_and_gamma_bounds_match_oracle(engine, lik, param, n):
    """gp_likelihood_crosswalk's 'binomial'/'bernoulli' and 'gamma' entries (utilities.py:989-1009) on the same site
    iteration: Bernoulli with gpflow's inv_probit link (Gauss-Hermite), Gamma with the exp link (closed form)."""
    from waveome_b200.engine import Batch
    X, y = other_data(n, max(4, n // 6), seed=n + 3, lik=lik)
    X2, y2 = other_data(n, max(4, n // 6), seed=n + 4, lik=lik)
    model = count_model(c=0.0)
    Y = np.stack([y, y2])
    batch = Batch(engine, X, Y, [model.program()])
    batch.set_likelihood(lik, param)
    rng = np.random.default_rng(n)
    x = batch.x0() + 0.2 * rng.normal(size=(2, batch.P))
    f, g, lml, st = batch.eval(x)
    fm, fv = batch.latent()
    olik = {"type": lik, "shape": param}
    for b in range(2):
        r = vo.vgp_collapsed(copy.deepcopy(model.to_spec()), olik, X, Y[b], x[b], rho=0.5, tol=1e-12, maxit=3000)
        assert st[b] == 0
        assert abs(lml[b] - r["F"]) <= 1e-8 * abs(r["F"]), (lml[b], r["F"])
        np.testing.assert_allclose(g[b], -r["grad"], rtol=0, atol=1e-6 * np.max(np.abs(r["grad"])))
        np.testing.assert_allclose(fm[b], r["m"], rtol=0, atol=1e-7 * (1 + np.max(np.abs(r["m"]))))
        np.testing.assert_allclose(fv[b], r["v"], rtol=1e-6, atol=1e-9)
    batch.close()


def test_gamma_shape_is_trained_and_bernoulli_fits(engine):
    """fit through the host API: Gamma's trainable shape rides in the noise slot (softplus); Bernoulli has none."""
    from waveome_b200.model_fitting import fit_models
    from waveome_b200.models import make_likelihood
    from waveome_b200.postfit import fitted_means
    for lik, olik in (("gamma", {"type": "gamma", "shape": 1.0}), ("bernoulli", {"type": "bernoulli"})):
        X, y = other_data(90, 15, seed=31, lik=lik)
        m = count_model(c=0.0)
        m.likelihood = make_likelihood(lik)
        spec = copy.deepcopy(m.to_spec())
        res = fit_models(X, y[None, :], [m], engine=engine)
        ro = vo.fit(copy.deepcopy(spec), olik, X, y)
        assert res["status"][0] in (0, 8)
        assert abs(res["lml"][0] - ro["F"]) <= 1e-5 * abs(ro["F"]), (lik, res["lml"][0], ro["F"])
        mu, st = fitted_means(X, y[None, :], [m], engine=engine)
        assert np.all(np.isfinite(mu)) and (lik != "bernoulli" or (mu.min() > 0 and mu.max() < 1))
        if lik == "gamma":
            assert abs(float(m.likelihood.shape) - 2.0) < 1.0 and float(m.likelihood.shape) != 1.0


@pytest.mark.parametrize("lik", ["poisson", "negative_binomial", "bernoulli", "gamma"])
def test_predictions_at_new_inputs_non_gaussian(engine, lik):
    """predict_f / predict_y / predict_log_density of a fitted non-Gaussian model (gpflow VGP.predict_f + the
    likelihood's predict_mean_and_var / predict_log_density) vs the oracle at the same hyper-parameters."""
    from scipy.special import logsumexp
    from waveome_b200.models import make_likelihood
    from waveome_b200 import postfit
    if lik in ("poisson", "negative_binomial"):
        X, y = count_data(70, 12, seed=5)
    else:
        X, y = other_data(70, 12, seed=5, lik=lik)
    m = count_model(c=0.1)
    m.likelihood = make_likelihood(lik)
    olik = {"type": lik, "alpha": 1.0, "shape": 1.0}
    spec = copy.deepcopy(m.to_spec())
    x = m.program().x0()
    r = vo.vgp_collapsed(spec, olik, X, y, x, rho=0.5, tol=1e-12, maxit=4000, want_grad=False)
    rng = np.random.default_rng(1)
    Xnew = np.stack([rng.integers(0, 14, size=23).astype(float), rng.normal(size=23)], 1)
    fm_o, fv_o = vo.predict_f(spec, X, y, x, r["sites"], Xnew)
    fm, fv = m.predict_f(Xnew, data=(X, y))
    np.testing.assert_allclose(fm[:, 0], fm_o, rtol=0, atol=1e-6 * (1 + np.max(np.abs(fm_o))))
    np.testing.assert_allclose(fv[:, 0], fv_o, rtol=1e-5, atol=1e-8)
    ym_o, yv_o = vo.predict_y_moments(vo.lik_of(spec, olik), fm_o, fv_o)
    ym, yv = m.predict_y(Xnew, data=(X, y))
    np.testing.assert_allclose(ym[:, 0], ym_o, rtol=1e-5)
    np.testing.assert_allclose(yv[:, 0], yv_o, rtol=1e-5)
    # predictive log density of plausible new observations: log-space quadrature of the oracle's log p(y | f)
    ynew = np.round(ym_o) if lik in ("poisson", "negative_binomial") else ((ym_o > 0.5).astype(float) if lik == "bernoulli" else ym_o)
    ld = m.predict_log_density((Xnew, ynew), data=(X, y))
    if lik == "bernoulli":
        want = np.log(np.where(ynew == 1, ym_o, 1 - ym_o))
    else:
        f = fm_o[:, None] + np.sqrt(2 * fv_o)[:, None] * vo.GH_X[None, :]
        lp = np.stack([postfit._likelihood_log_prob(m.likelihood, f[:, k], ynew) for k in range(20)], 1)
        want = logsumexp(lp + np.log(vo.GH_W / np.sqrt(np.pi))[None, :], axis=1)
    np.testing.assert_allclose(ld, want, rtol=1e-5, atol=1e-7)
    assert np.all(np.isfinite(ld))


def test_zinb_bound_gradient_and_fit(engine):
    """Zero-inflated negative binomial (waveome/likelihoods.py:96-139; 'zeroinflated_negativebinomial' in
    gp_likelihood_crosswalk): alpha in the noise slot, km in the second likelihood slot, both softplus."""
    from test_vgp_oracle import zinb_sample
    from waveome_b200.engine import Batch
    from waveome_b200.model_fitting import fit_models
    from waveome_b200.models import make_likelihood
    rng = np.random.default_rng(12)
    n = 120
    subj = rng.integers(0, 15, size=n).astype(float)
    t = rng.normal(size=n)
    X = np.stack([subj, t], 1)
    f = 0.5 * rng.normal(size=15)[subj.astype(int)] + 0.7 * np.sin(2 * t) + 1.2
    Y = np.stack([zinb_sample(rng, f), zinb_sample(rng, f - 0.5, alpha=1.0, km=2.0)])
    assert (Y == 0).sum() > 30
    model = count_model(c=0.3)
    model.likelihood = make_likelihood("zeroinflated_negativebinomial", alpha=0.7, km=1.5)
    p = model.program()
    assert p.n_x == 6 and p.lik_slot2 >= 0
    batch = Batch(engine, X, Y, [p])
    batch.set_likelihood("zinb", (9.0, 9.0))                 # ignored: both slots are trainable
    x = batch.x0() + 0.15 * rng.normal(size=(2, batch.P))
    fv_, g, lml, st = batch.eval(x)
    fm, fvar = batch.latent()
    spec = model.to_spec()
    for b in range(2):
        r = vo.vgp_collapsed(copy.deepcopy(spec), {"type": "zinb"}, X, Y[b], x[b], rho=0.5, tol=1e-12, maxit=5000)
        assert st[b] == 0
        assert abs(lml[b] - r["F"]) <= 1e-8 * abs(r["F"]), (lml[b], r["F"])
        np.testing.assert_allclose(g[b], -r["grad"], rtol=0, atol=1e-6 * np.max(np.abs(r["grad"])))
        np.testing.assert_allclose(fm[b], r["m"], rtol=0, atol=1e-6 * (1 + np.max(np.abs(r["m"]))))
        np.testing.assert_allclose(fvar[b], r["v"], rtol=1e-5, atol=1e-8)
    batch.close()
    # fit through the host API and predictions with the fitted likelihood
    m2 = count_model(c=0.0)
    m2.likelihood = make_likelihood("zinb")
    spec2 = copy.deepcopy(m2.to_spec())
    res = fit_models(X, Y[:1], [m2], engine=engine)
    ro = vo.fit(spec2, {"type": "zinb"}, X, Y[0])
    assert res["status"][0] in (0, 8)
    assert abs(res["lml"][0] - ro["F"]) <= 1e-5 * abs(ro["F"]), (res["lml"][0], ro["F"])
    assert float(m2.likelihood.alpha) != 1.0 and float(m2.likelihood.km) != 1.0
    ym, yv = m2.predict_y(X[:9], data=(X, Y[0]))
    ld = m2.predict_log_density((X[:9], Y[0][:9]), data=(X, Y[0]))
    assert np.all(ym > 0) and np.all(yv > 0) and np.all(np.isfinite(ld))


def test_zinb_sites_at_the_precision_bound(engine):
    """Zeros at a high latent mean under a tight prior: the ZINB zero branch is not log-concave there, the optimal site
    precisions would be negative and sit at the bound 1e-6 instead.  The value must still be the oracle's (a valid
    lower bound), the evaluation carries status bit 32, and the gradient -- which omits the non-stationarity term of
    the bounded sites -- stays within a few percent of the oracle's finite differences."""
    from test_vgp_oracle import zinb_sample
    from waveome_b200.engine import Batch
    from waveome_b200.models import make_likelihood
    rng = np.random.default_rng(4)
    n = 64
    subj = np.repeat(np.arange(8), 8).astype(float)
    t = rng.normal(size=n)
    X = np.stack([subj, t], 1)
    y = zinb_sample(rng, 0.3 * np.sin(2 * t) + 3.0, alpha=0.3, km=0.5)
    y[::7] = 0.0
    model = count_model(var=0.05, c=3.0)
    model.likelihood = make_likelihood("zinb", alpha=0.3, km=0.5)
    batch = Batch(engine, X, y[None, :], [model.program()])
    batch.set_likelihood("zinb", (1.0, 1.0))          # placeholders: both parameters ride in trainable slots
    x = batch.x0()
    f, g, lml, st = batch.eval(x)
    batch.close()
    spec = model.to_spec()
    kw = dict(rho=0.3, tol=1e-11, maxit=20000)
    r = vo.vgp_collapsed(copy.deepcopy(spec), {"type": "zinb"}, X, y, x[0], **kw)
    assert np.sum(r["sites"][0] <= 1.0001e-6) >= 5
    assert st[0] == 32
    assert abs(lml[0] - r["F"]) <= 1e-7 * abs(r["F"]), (lml[0], r["F"])
    np.testing.assert_allclose(g[0], -r["grad"], rtol=0, atol=1e-5 * np.max(np.abs(r["grad"])))
    h, fd = 1e-5, []
    for i in range(len(x[0])):
        xp, xm = x[0].copy(), x[0].copy()
        xp[i] += h; xm[i] -= h
        fd.append((vo.vgp_collapsed(copy.deepcopy(spec), {"type": "zinb"}, X, y, xp, want_grad=False, **kw)["F"]
                   - vo.vgp_collapsed(copy.deepcopy(spec), {"type": "zinb"}, X, y, xm, want_grad=False, **kw)["F"]) / (2 * h))
    assert np.max(np.abs(-g[0] - np.array(fd))) <= 0.05 * np.max(np.abs(fd))


def test_penalized_optimization_zinb_outcomes():
    """GPSearch with outcome_likelihood='zeroinflated_negativebinomial' (the crosswalk name): fits run on the engine;
    the reference's deviance has no ZINB branch (utilities.py:544-581), so the importances are None."""
    from test_vgp_oracle import zinb_sample
    from waveome_b200 import datasets
    from waveome_b200.model_search import GPSearch
    X, Y = datasets.count_microbiome(n_subjects=15, n_times=6, n_outcomes=3, seed=5)
    rng = np.random.default_rng(3)
    for c in Y.columns:
        Y[c] = zinb_sample(rng, np.log(Y[c].to_numpy() + 1.0) * 0.6 + 0.5, alpha=0.5, km=1.0)
    gps = GPSearch(X, Y, unit_col="subject", outcome_likelihood="zeroinflated_negativebinomial")
    gps.penalized_optimization()
    assert len(gps.models) == 3
    for m in gps.models.values():
        assert m.likelihood.name == "zinb" and np.isfinite(m.log_posterior_density_value)
        assert m.feature_importances is None
        assert float(m.likelihood.alpha) > 0 and float(m.likelihood.km) > 0
