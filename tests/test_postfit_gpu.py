"""Post-fit quantities on the engine vs the oracle's arithmetic: alpha, posterior means at training and new inputs,
feature importances (waveome/utilities.py:614-707)."""
import copy

import numpy as np
import pytest

import gp_oracle as oracle
import helpers
import waveome_b200 as wb
from waveome_b200 import postfit

pytestmark = pytest.mark.gpu


def _oracle_alpha_mean(model, X, y, Xnew=None):
    spec = copy.deepcopy(model.to_spec())
    K, _ = oracle.kernel_K_and_grads(spec["kernel"], X, want_grads=False)
    s2 = spec["likelihood_variance"]["value"]
    c = spec["mean"]["c"]["value"] if spec["mean"]["type"] == "constant" else 0.0
    alpha = np.linalg.solve(K + s2 * np.eye(len(y)), y - c)
    if Xnew is None:
        return alpha, c + K @ alpha
    Kall, _ = oracle.kernel_K_and_grads(spec["kernel"], np.vstack([Xnew, X]), want_grads=False)
    return alpha, c + Kall[: len(Xnew), len(Xnew):] @ alpha


@pytest.mark.parametrize("n,m", [(50, 7), (150, 64), (333, 130)])
def test_alpha_and_predict_mean(engine, n, m):
    from waveome_b200.engine import Batch
    X, y = helpers.make_data(n, seed=n)
    rng = np.random.default_rng(n)
    Xnew = X[rng.integers(0, n, size=m)].copy()
    Xnew[:, 1:3] += 0.3 * rng.normal(size=(m, 2))
    for kern in (helpers.all_leaf_kernel(), helpers.saturated_kernel()):
        model = wb.GPR(kern, mean_function=wb.ConstantMean(0.2), noise_variance=0.3)
        batch = Batch(engine, X, np.stack([y, -y]), [model.program()])
        batch.eval(batch.x0())
        a = batch.alpha()
        mu = batch.predict_mean(Xnew)
        batch.close()
        ao, muo = _oracle_alpha_mean(model, X, y, Xnew)
        np.testing.assert_allclose(a[0], ao, rtol=0, atol=1e-9 * np.max(np.abs(ao)))
        np.testing.assert_allclose(mu[0], muo, rtol=0, atol=1e-9 * max(1.0, np.max(np.abs(muo))))
        mu2, var2 = model.predict_f(Xnew, data=(X, y))
        np.testing.assert_array_equal(mu2[:, 0], mu[0])
        # predictive variance: k** - k*^T (K + s2 I)^-1 k*
        spec = copy.deepcopy(model.to_spec())
        Kall, _ = oracle.kernel_K_and_grads(spec["kernel"], np.vstack([Xnew, X]), want_grads=False)
        Kss, Ksx, Kxx = Kall[:m, :m], Kall[:m, m:], Kall[m:, m:]
        vref = np.diag(Kss) - np.einsum("ij,ij->i", Ksx, np.linalg.solve(Kxx + 0.3 * np.eye(n), Ksx.T).T)
        np.testing.assert_allclose(var2[:, 0], vref, rtol=0, atol=1e-9 * max(1.0, np.max(np.abs(vref))))
        _, vy = model.predict_y(Xnew, data=(X, y))
        np.testing.assert_allclose(vy[:, 0], vref + 0.3, rtol=0, atol=1e-9 * max(1.0, np.max(np.abs(vref))))
        lp = model.predict_log_density((Xnew[:5], np.zeros(5)), data=(X, y))
        assert lp.shape == (5,) and np.all(np.isfinite(lp))


def test_feature_importances_match_reference_arithmetic(engine):
    n = 120
    X, y = helpers.make_data(n, seed=4)
    k = wb.Sum([wb.Categorical(active_dims=[0], variance=0.4), wb.SquaredExponential(active_dims=[1], lengthscales=0.6),
                wb.Product([wb.Categorical(active_dims=[3]), wb.SquaredExponential(active_dims=[1])])])
    model = wb.GPR(k, mean_function=wb.ConstantMean(0.1), noise_variance=0.05)
    single = wb.GPR(wb.SquaredExponential(active_dims=[1], lengthscales=0.5), mean_function=wb.ConstantMean(), noise_variance=0.1)
    for rv in ("log_bf", "statistic", "de"):
        got = postfit.feature_importances_batch(X, np.stack([y, y]), [model, single], return_value=rv)
        # reference arithmetic with oracle means
        exp = []
        for mdl in (model, single):
            _, mu_full = _oracle_alpha_mean(mdl, X, y)
            null, mod, sat = postfit.calc_deviance_loglik(y, mu_full)
            full_de = max(min(1, 1 - np.sum(mod - sat) / np.sum(null - sat)), 0) \
                if np.sum(sat) >= np.sum(mod) >= np.sum(null) else 0
            lst = []
            if mdl.kernel.name == "sum":
                for i in range(len(mdl.kernel.kernels)):
                    mc = wb.deepcopy(mdl)
                    mc.kernel.kernels.pop(i)
                    _, mu_sub = _oracle_alpha_mean(mc, X, y)
                    n2, sub, _ = postfit.calc_deviance_loglik(y, mu_sub)
                    if rv == "statistic":
                        lst.append(max(np.round(-2 * (np.sum(sub) - np.sum(mod)), 1), 0))
                    elif rv == "log_bf":
                        lst.append(np.round(np.sum(mod) - np.sum(sub), 1))
                    else:
                        lst.append(np.round(max(min(1, 1 - np.sum(sub - mod) / np.sum(n2 - mod)), 0), 3))
            else:
                lst.append({"statistic": np.round(-2 * (np.sum(null) - np.sum(mod)), 1),
                            "log_bf": np.round(np.sum(mod) - np.sum(null), 1), "de": np.round(full_de, 3)}[rv])
            lst.append(np.round(1 - full_de, 3))
            exp.append(lst)
        for g, e in zip(got, exp):
            assert len(g) == len(e)
            np.testing.assert_allclose(g, e, rtol=0, atol=0.11 if rv != "de" else 1.1e-3)   # one rounding step


def test_penalized_optimization_sets_feature_importances():
    from waveome_b200 import datasets
    from waveome_b200.model_search import GPSearch
    X, Y = datasets.overview_notebook(n_people=30, n_observations=5)
    gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
    gps.penalized_optimization(random_seed=1)
    for o, m in gps.models.items():
        ncomp = len(m.kernel.kernels) if m.kernel.name == "sum" else 1
        assert len(m.feature_importances) == ncomp + 1
        assert 0.0 <= m.feature_importances[-1] <= 1.0
    # outcome1 = sin(time) + noise: almost nothing is left for the residual
    assert gps.models["outcome1"].feature_importances[-1] < 0.1


def test_train_predictive_variance_and_iterated_factor(engine):
    """predict_y variances at the training inputs vs the oracle's linear algebra, and penalization_factor=None
    (waveome/model_search.py:271-375): factors start at the formula value and only ever decrease."""
    from scipy.stats import norm
    from waveome_b200 import datasets
    from waveome_b200.model_search import GPSearch
    n = 90
    X, y = helpers.make_data(n, seed=8)
    model = wb.GPR(helpers.saturated_kernel(hs=0.0), mean_function=wb.ConstantMean(0.1), noise_variance=0.3)
    var_y = postfit.train_predictive_variance(X, y[None, :], [model])[0]
    spec = copy.deepcopy(model.to_spec())
    K, _ = oracle.kernel_K_and_grads(spec["kernel"], X, want_grads=False)
    s2 = 0.3
    ref = np.diag(K - K @ np.linalg.solve(K + s2 * np.eye(n), K)) + s2
    np.testing.assert_allclose(var_y, ref, rtol=1e-8, atol=1e-10)
    Xd, Yd = datasets.overview_notebook(n_people=30, n_observations=5)
    gps = GPSearch(Xd, Yd, unit_col="person_id", categorical_vars=["female"])
    gps.penalized_optimization(penalization_factor=None, num_factor_iter=3)
    p = 5                                                   # unit + female + SE[time] + unit... components of the saturated kernel
    for o, m in gps.models.items():
        start = 2 * 1.1 * np.std(gps.Y[o].to_numpy()) * np.sqrt(len(gps.X)) * norm().ppf(1 - 0.1 / (2 * 4))
        assert m.penalization_factor <= start + 1e-9 and m.penalization_factor > 0
        assert np.isfinite(m.log_posterior_density_value)


def test_component_predictions(engine):
    """individual_kernel_predictions (waveome/utilities.py:710-974): per-component posterior mean and variance at new
    inputs, joint (marginal=False) and as a stand-alone model (marginal=True, the reference's default)."""
    from waveome_b200.utilities import individual_kernel_predictions
    n, m = 140, 33
    X, y = helpers.make_data(n, seed=8)
    rng = np.random.default_rng(2)
    Xnew = X[rng.integers(0, n, size=m)].copy()
    Xnew[:, 1:3] += 0.2 * rng.normal(size=(m, 2))
    kern = wb.Sum([wb.Categorical(active_dims=[0], variance=0.6),
                   wb.SquaredExponential(active_dims=[1], variance=1.2, lengthscales=0.7),
                   wb.Product([wb.Categorical(active_dims=[3]), wb.Matern32(active_dims=[2], lengthscales=1.3)])])
    model = wb.GPR(kern, mean_function=wb.ConstantMean(0.15), noise_variance=0.25)
    spec = copy.deepcopy(model.to_spec())
    c, s2 = 0.15, 0.25
    Kfull, _ = oracle.kernel_K_and_grads(spec["kernel"], X, want_grads=False)
    for marginal in (False, True):
        parts = postfit.component_predictions(model, X, y, Xnew, marginal=marginal)
        assert len(parts) == 3
        for k, sub in enumerate(spec["kernel"]["kernels"]):
            Kall, _ = oracle.kernel_K_and_grads(sub, np.vstack([Xnew, X]), want_grads=False)
            Kss, Ksx, Kxx = Kall[:m, :m], Kall[:m, m:], Kall[m:, m:]
            A = (Kxx if marginal else Kfull) + s2 * np.eye(n)
            mu_ref = c + Ksx @ np.linalg.solve(A, y - c)
            var_ref = np.diag(Kss) - np.einsum("ij,ij->i", Ksx, np.linalg.solve(A, Ksx.T).T)
            np.testing.assert_allclose(parts[k][0], mu_ref, rtol=0, atol=1e-9 * max(1.0, np.max(np.abs(mu_ref))))
            np.testing.assert_allclose(parts[k][1], var_ref, rtol=0, atol=1e-9 * max(1.0, np.max(np.abs(var_ref))))
    mu, var, samples, cov = individual_kernel_predictions(model, 1, data=(X, y), X=Xnew, marginal=False)
    parts = postfit.component_predictions(model, X, y, Xnew, marginal=False)
    assert mu.shape == (m, 1) and samples is None and cov is None
    np.testing.assert_array_equal(mu[:, 0], parts[1][0])
    with pytest.raises(ValueError):
        individual_kernel_predictions(model, 5, data=(X, y), X=Xnew)


def test_post_fit_work_behind_the_fit_gives_the_same_models():
    """penalized_optimization fits a large job as concurrent pieces and does the post-fit work of a piece (pruning, feature
    importances on an engine of its own) while the others are still on the device: names, fitted values and importances must
    be those of the single-piece run."""
    from waveome_b200 import datasets, model_fitting as mf
    from waveome_b200.model_search import GPSearch
    X, Y = datasets.overview_synthetic(n_people=10, n_observations=5, n_outcomes=600)
    out = {}
    for streams in (1, 4):
        old, mf.FIT_STREAMS = mf.FIT_STREAMS, streams
        try:
            g = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
            g.penalized_optimization(num_restart=0, optimization_options={"num_opt_iter": 30})
        finally:
            mf.FIT_STREAMS = old
        out[streams] = g
    a, b = out[1], out[4]
    assert all(not v for v in mf._LEASED.values())
    for o in a.out_names:
        ma, mb = a.models[o], b.models[o]
        assert ma.kernel_name == mb.kernel_name
        assert [float(p) for p in ma.trainable_parameters] == [float(p) for p in mb.trainable_parameters]
        np.testing.assert_array_equal(np.asarray(ma.feature_importances, dtype=float),
                                      np.asarray(mb.feature_importances, dtype=float))
