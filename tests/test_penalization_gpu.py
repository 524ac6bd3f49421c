"""Cross-validated penalisation search (waveome/model_classes.py:866-998, regularization.py:245-276) on the engine."""
import copy

import numpy as np
import pytest

import gp_oracle as oracle
import helpers
import waveome_b200 as wb
from waveome_b200 import penalization

pytestmark = pytest.mark.gpu


def test_make_folds_unit_level_and_reproducible():
    X, _ = helpers.make_data(100, seed=1, n_subj=12)
    f1 = penalization.make_folds(X, 0, k_fold=3, random_seed=7)
    f2 = penalization.make_folds(X, 0, k_fold=3, random_seed=7)
    assert all(np.array_equal(a, b) for a, b in zip(f1, f2))
    assert sorted(np.concatenate(f1).tolist()) == list(range(100))
    units = [set(X[f, 0]) for f in f1]
    assert not (units[0] & units[1]) and not (units[1] & units[2])          # a subject never straddles folds
    f3 = penalization.make_folds(X, None, k_fold=4, random_seed=7)
    assert sorted(len(f) for f in f3) == [25, 25, 25, 25]


def test_search_scores_match_oracle_and_rule(engine):
    n = 90
    X, y = helpers.make_data(n, seed=5, n_subj=9)
    rng = np.random.default_rng(2)
    Y = np.stack([y, rng.normal(size=n)])
    kern = helpers.saturated_kernel(hs=0.0)
    factors = [0.0, 1.0, 50.0]
    out = penalization.penalization_search_batch(X, Y, kern, penalization_factor_list=factors, k_fold=3, unit_col=0,
                                                 random_seed=3, num_restart=2, fit_best=True, engine=engine)
    res = out["results"]
    assert res.shape == (2, 3, 3) and np.all(np.isfinite(res))
    # the selection rule (:961-975) on the scores
    for b in range(2):
        vals = [res[b, fi].mean() - res[b, fi].std() / np.sqrt(3) for fi in range(3)]
        assert out["best_factor"][b] == factors[int(np.argmax(vals))]
    # pure noise must not score better held-out than the structured outcome
    assert res[1].mean() < res[0].mean()
    # EVERY held-out score against the oracle's predictive density (gpflow GPR.predict_log_density arithmetic) at the
    # parameters the engine fitted for that (outcome, factor, fold): the engine's fit state and wv_batch_predict_f
    folds = out["folds"]
    for (b, fi, k), m in out["fold_models"].items():
        train = np.setdiff1d(np.arange(n), folds[k])
        ref = np.mean(oracle.predict_log_density(m.to_spec(), X[train], Y[b, train], X[folds[k]], Y[b, folds[k]]))
        assert abs(res[b, fi, k] - ref) <= 1e-8 * max(1.0, abs(ref)), (b, fi, k, res[b, fi, k], ref)
    # ... and the fits behind them: the un-penalised factor's best restart is a stationary point of the ORACLE's objective
    # on the training rows (gradient below the optimiser's tolerance scale), with the objective value the engine reported
    for (b, fi, k), m in out["fold_models"].items():
        if fi != 0 or m.fit_info["status"] != 0:
            continue
        train = np.setdiff1d(np.arange(n), folds[k])
        spec = m.to_spec()
        fo, go, _, _ = oracle.objective(copy.deepcopy(spec), X[train], Y[b, train], oracle.pack(spec))
        assert abs(fo + m.log_posterior_density_value) <= 1e-8 * max(1.0, abs(fo))
        assert np.max(np.abs(go)) <= 5e-3 * max(1.0, abs(fo)) ** 0.5, (b, k, go)
    assert len(out["models"]) == 2 and all(np.isfinite(m.log_posterior_density_value) for m in out["models"])
    assert out["models"][0].penalization_factor == out["best_factor"][0]


def test_model_method(engine):
    n = 80
    X, y = helpers.make_data(n, seed=9, n_subj=8)
    m = wb.models.PenalizedGPR(helpers.saturated_kernel(hs=0.0), mean_function=wb.ConstantMean())
    m.penalization_search(data=(X, y), penalization_factor_list=[0.0, 10.0], k_fold=2, random_seed=1, num_restart=2,
                          unit_col=0)
    assert m.penalization_search_results.shape == (4, 3)
    assert m.penalization_factor in (0.0, 10.0) and np.isfinite(m.log_posterior_density_value)
