"""run_search / full_kernel_search on the engine (BASELINE configs[1]: overview synthetic data, 50 subjects x 10 time
points, Gaussian outcomes, kernels [SE, Matern12, Lin, Periodic(SE)]): the engine-fitted search must select the same
structure, with the same rounded BICs, as the identical host logic driven by the CPU oracle."""
import numpy as np
import pytest

import waveome_b200 as wb
from waveome_b200 import datasets, kernel_search as ks
from waveome_b200.model_search import GPSearch
from oracle_fitter import oracle_fitter
from test_kernel_search_cpu import _toy

pytestmark = pytest.mark.gpu


def test_search_engine_vs_oracle_small(engine):
    X, y = _toy()
    kl = [wb.SquaredExponential(), wb.Lin()]
    a = ks.full_kernel_search(X, y, kl, cat_vars=[0, 2], max_depth=3, engine=engine, num_restart=1, keep_only_best=False)
    b = ks.full_kernel_search(X, y, kl, cat_vars=[0, 2], max_depth=3, fit=oracle_fitter(X), keep_only_best=False)
    assert a["best_model"] == b["best_model"]
    assert set(a["models"]) == set(b["models"]) and a["edges"] == b["edges"]
    for k in a["models"]:
        assert abs(a["models"][k]["bic"] - b["models"][k]["bic"]) <= 0.011, (k, a["models"][k]["bic"], b["models"][k]["bic"])
        assert a["models"][k]["try_next"] == b["models"][k]["try_next"]


def test_run_search_overview_outcomes_vs_oracle():
    """Archetype outcomes of the config-2 generator at a size the oracle finishes quickly (20 subjects x 6):
    lock-step batched search on the GPU == per-outcome search with oracle fits.  (Without the periodic kernel here:
    its fits start with line searches that run into failed factorisations, where the next step depends on the last
    bit of a pivot on either side; the periodic kernel is covered by the config-2 test below.)"""
    X, Y = datasets.overview_synthetic(n_people=20, n_observations=6, n_outcomes=8)
    kl = lambda: [wb.SquaredExponential(), wb.Matern12(), wb.Lin()]
    gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
    gps.run_search(kernels=kl(), max_depth=3, random_seed=0)
    assert gps.fit_report["batches"] < gps.fit_report["n_fits"] / 4          # requests of all outcomes share batches
    ref = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
    ref.run_search(kernels=kl(), max_depth=3, random_seed=0, fit=oracle_fitter(ref.X.to_numpy(dtype=np.float64)))
    for o in gps.out_names:
        assert gps.search_info[o]["best_model"] == ref.search_info[o]["best_model"], o
        a = gps.search_info[o]["models"][gps.search_info[o]["best_model"]]
        b = ref.search_info[o]["models"][ref.search_info[o]["best_model"]]
        xa = gps.models[o].program().x0()
        xb = ref.models[o].program().x0()
        if max(np.max(np.abs(xa)), np.max(np.abs(xb))) > 20:
            # a parameter ran off along a flat direction (variance -> 0 with its lengthscale -> inf): the optimum is a
            # ridge, trajectories are sensitive to the last bit, only the objective value is comparable
            assert abs(a["bic"] - b["bic"]) <= 0.5
            continue
        assert abs(a["bic"] - b["bic"]) <= 0.011
        np.testing.assert_allclose(xa, xb, rtol=1e-4, atol=1e-4)


def test_run_search_config2_shape_lockstep():
    """n = 500 (50 x 10), 12 outcomes cycling the four archetypes: structure recovered per archetype."""
    X, Y = datasets.overview_synthetic(n_outcomes=12)
    gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
    gps.run_search(max_depth=3)
    names = {o: gps.search_info[o]["best_model"] for o in gps.out_names}
    for j, o in enumerate(gps.out_names):
        k = j % 4
        if k == 0:
            assert "[1]" in names[o] and "categorical[0]" not in names[o], (o, names[o])      # smooth function of time
        elif k == 1:
            assert "categorical[2]" in names[o] and "[1]" in names[o], (o, names[o])          # female x f(time)
        elif k == 2:
            assert "categorical[0]" in names[o] and "[1]" in names[o], (o, names[o])          # unit effect + trend
        else:
            assert names[o] == "constant", (o, names[o])                                        # pure noise
    assert gps.fit_report["batches"] <= 12              # one candidate batch + one pruning batch per depth


def test_run_search_poisson_counts():
    """The search with a count likelihood (model_fitting.py:158-185 builds a VGP per candidate; here every candidate is a
    collapsed-bound fit): Poisson outcomes with a subject effect and a smooth time trend must select both, and an
    outcome that only has the subject effect must not select time."""
    X, Y = datasets.count_microbiome(n_subjects=20, n_times=6, n_outcomes=3, seed=3)
    rng = np.random.default_rng(0)
    subj = X["subject"].to_numpy().astype(int)
    Y["flat"] = rng.poisson(np.exp(1.0 + 0.8 * rng.normal(size=20)[subj])).astype(float)
    gps = GPSearch(X, Y, unit_col="subject", outcome_likelihood="poisson")
    gps.run_search(kernels=[wb.SquaredExponential(), wb.Lin()], max_depth=2, random_seed=0)
    for o in gps.out_names:
        best = gps.search_info[o]["best_model"]
        assert gps.models[o].likelihood.name == "poisson" and np.isfinite(gps.search_info[o]["models"][best]["bic"])
        assert "categorical[0]" in best, (o, best)
    assert "[1]" not in gps.search_info["flat"]["best_model"]
    assert np.mean(["[1]" in gps.search_info[o]["best_model"] for o in gps.out_names[:3]]) >= 2 / 3
    # kernel_test / full_kernel_search drop-ins take the likelihood as well
    m, bic = ks.kernel_test(X.to_numpy(), Y["flat"].to_numpy(), wb.Categorical(active_dims=[0]), likelihood="poisson",
                            num_restart=1)
    assert m.likelihood.name == "poisson" and np.isfinite(bic)


def test_pipelined_run_search_matches_single_group():
    """run_search with two outcome groups (device batch of one group behind the host work of the other, fits on a
    worker thread) against the single-group driver: fits do not depend on the composition of their batch, so every
    candidate's BIC and the selected structures are identical."""
    X, Y = datasets.overview_synthetic(n_people=12, n_observations=5, n_outcomes=36)
    out = {}
    for groups in (1, 2):
        gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
        gps.run_search(kernels=[wb.SquaredExponential(), wb.Lin()], max_depth=3, pipeline_groups=groups)
        out[groups] = gps
    a, b = out[1], out[2]
    assert a.fit_report["n_fits"] == b.fit_report["n_fits"]
    for o in a.out_names:
        assert a.search_info[o]["best_model"] == b.search_info[o]["best_model"], o
        ma, mb = a.search_info[o]["models"], b.search_info[o]["models"]
        assert list(ma.keys()) == list(mb.keys())
        assert [v["bic"] for v in ma.values()] == [v["bic"] for v in mb.values()], o
        np.testing.assert_array_equal(a.models[o].program().x0(), b.models[o].program().x0())


def test_deferred_stragglers_run_search_matches_level_synchronous(monkeypatch):
    """run_search with the stragglers of a level finishing in the background (kernel_search.SEARCH_TAIL > 0: the fit
    returns once at most that many models are still iterating, wv_batch_fit_lbfgs_begin / _run / _report) against the
    level-synchronous driver: bit-identical fits, identical candidate sets, BICs and selected structures."""
    X, Y = datasets.overview_synthetic(n_people=12, n_observations=5, n_outcomes=36)
    out = {}
    for tail in (0, 24):
        monkeypatch.setattr(ks, "SEARCH_TAIL", tail)
        gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
        gps.run_search(max_depth=3)
        out[tail] = gps
    a, b = out[0], out[24]
    assert a.fit_report["n_fits"] == b.fit_report["n_fits"]
    assert b.fit_report["batches"] > a.fit_report["batches"]             # late outcomes did rejoin in later batches
    for o in a.out_names:
        assert a.search_info[o]["best_model"] == b.search_info[o]["best_model"], o
        ma, mb = a.search_info[o]["models"], b.search_info[o]["models"]
        assert list(ma.keys()) == list(mb.keys())
        assert [v["bic"] for v in ma.values()] == [v["bic"] for v in mb.values()], o
        np.testing.assert_array_equal(a.models[o].program().x0(), b.models[o].program().x0())
    from waveome_b200 import model_fitting as mf
    assert all(not v for v in mf._LEASED.values())                       # every engine lease was returned


def test_split_kernel_search_engine_vs_oracle(engine):
    """split_kernel_search (waveome/model_search.py:3275-3532) on the engine -- training fits by the device L-BFGS-B, hold-out
    scores through wv_batch_predict_f -- against the same search with the oracle's fits and the oracle's predictive
    density: same split, same candidates, same hold-out criteria (to their two printed decimals), same selection."""
    import gp_oracle as oracle
    X, y = _toy()
    kl = lambda: [wb.SquaredExponential(), wb.Lin()]
    a = ks.split_kernel_search(X, y, kl(), unit_idx=0, cat_vars=[0, 2], max_depth=2, random_seed=4, engine=engine,
                               num_restart=1, keep_only_best=False)
    tr = np.isin(X[:, 0], np.unique(a["X"][:, 0]))
    b = ks.split_kernel_search(X, y, kl(), unit_idx=0, cat_vars=[0, 2], max_depth=2, random_seed=4, keep_only_best=False,
                               fit=oracle_fitter(X[tr]),
                               log_density=lambda m, Xt, yt, Xh, yh: oracle.predict_log_density(m.to_spec(), Xt, yt, Xh, yh))
    np.testing.assert_array_equal(a["X"], b["X"])
    np.testing.assert_array_equal(a["X_holdout"], b["X_holdout"])
    assert a["best_model"] == b["best_model"] and set(a["models"]) == set(b["models"]) and a["edges"] == b["edges"]
    for k in a["models"]:
        assert abs(a["models"][k]["bic"] - b["models"][k]["bic"]) <= 0.011, (k, a["models"][k]["bic"], b["models"][k]["bic"])
