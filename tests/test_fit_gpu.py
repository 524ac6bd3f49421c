"""Device L-BFGS-B fits vs the oracle's SciPy L-BFGS-B run on the same objective and start: optimised
hyper-parameters to 1e-5, same objective value, same selected kernel structure."""
import copy
import json
import os

import numpy as np
import pytest

import gp_oracle as oracle
import helpers
import waveome_b200 as wb

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_fit_matches_golden_scipy_runs(engine):
    from waveome_b200.engine import Batch
    with open(os.path.join(GOLDEN, "fit_cases.json")) as fh:
        cases = [c for c in json.load(fh) if c["status"] == 0]
    assert len(cases) >= 2
    for c in cases:
        X, y = helpers.make_data(c["n"], seed=c["seed"])
        model = wb.GPR(helpers.saturated_kernel(), mean_function=wb.ConstantMean(0.0))
        batch = Batch(engine, X, y[None, :], [model.program()])
        r = batch.fit()
        assert r["status"][0] == 0
        assert r["n_iter"][0] == c["nit"] and r["n_eval"][0] == c["nfev"]
        assert abs(r["f"][0] - c["f"]) <= 1e-8 * abs(c["f"])
        np.testing.assert_allclose(r["x"][0], np.array(c["x"]), rtol=1e-5, atol=1e-5)   # north star: 1e-5
        batch.close()


def test_batched_fit_independent_of_batch_composition(engine):
    """Fitting a model alone or inside a batch of others gives bit-identical results (per-model state machines)."""
    from waveome_b200.engine import Batch
    n = 120
    X, y = helpers.make_data(n, seed=21)
    rng = np.random.default_rng(4)
    Y = np.stack([y, y + 0.3 * rng.normal(size=n), rng.normal(size=n), np.sin(3 * X[:, 1])])
    model = wb.GPR(helpers.saturated_kernel(hs=0.0), mean_function=wb.ConstantMean(0.0))
    batch = Batch(engine, X, Y, [model.program()])
    r = batch.fit()
    batch.close()
    for b in (0, 3):
        single = Batch(engine, X, Y[b:b + 1], [model.program()])
        rs = single.fit()
        single.close()
        assert np.array_equal(rs["x"][0], r["x"][b]) and rs["n_eval"][0] == r["n_eval"][b]


def test_fit_vs_oracle_structure(engine):
    """Penalised fits on the overview notebook data (soft pin: waveome_overview.ipynb text — outcome1 -> SE[time],
    outcome2 -> female x SE[time], outcome3 -> unit effect + linear time): GPU fit == oracle fit, and the pruned
    structure is the notebook's."""
    from waveome_b200 import datasets
    from waveome_b200.model_search import GPSearch
    X, Y = datasets.overview_notebook()
    gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
    gps.penalized_optimization(random_seed=9102, kernel_options={
        "second_order_numeric": False, "unit_numeric_interactions": False, "categorical_numeric_interactions": True,
        "kerns": [wb.SquaredExponential(), wb.Lin()]})
    names = {k: m.kernel_name for k, m in gps.models.items()}
    assert names["outcome1"] == "squared_exponential[1]"
    assert names["outcome2"] == "categorical[2]*squared_exponential[1]"
    assert names["outcome3"] == "categorical[0]+lin[1]"
    # noise variance soft pin (notebook cell 11, SVGP path: 0.010672; exact-GPR optimum per SURVEY App. C: 0.01075)
    assert abs(float(gps.models["outcome1"].likelihood.variance) - 0.0107) < 5e-4


def test_readme_quickstart_iris_vs_oracle():
    """BASELINE configs[0] — README quick-start: iris, X = petal_length, petal_width, species; Y = sepal_length,
    sepal_width; GPSearch(...).penalized_optimization().  Every outcome: engine fit == oracle's SciPy L-BFGS-B fit of the
    same saturated penalised model (objective, pruned structure)."""
    from waveome_b200 import datasets
    from waveome_b200.model_search import GPSearch
    from waveome_b200.regularization import full_kernel_build
    X, Y = datasets.iris()
    gps = GPSearch(X, Y, categorical_vars=["species"])
    gps.penalized_optimization()
    Xn = gps.X.to_numpy(dtype=np.float64)
    k, names = full_kernel_build(cat_vars=gps.cat_idx, num_vars=gps.cont_idx, unit_idx=None, var_names=gps.feat_names,
                                 return_sum=True)
    assert len(names) == 5                      # cat[species] + SE[pl] + SE[pw] + cat x SE[pl] + cat x SE[pw]
    for o in gps.out_names:
        m = wb.models.PenalizedGPR(wb.deepcopy(k), mean_function=wb.ConstantMean(), penalization_factor=1.0)
        ro = oracle.fit(m.to_spec(), Xn, gps.Y[o].to_numpy(dtype=np.float64), maxiter=50000, maxfun=50000)
        got = gps.models[o]
        assert abs(-got.log_posterior_density_value - ro["f"]) <= 1e-6 * max(1.0, abs(ro["f"])), (o, got.log_posterior_density_value, ro["f"])
        m.program().assign(ro["x"])
        m.cut_kernel_components(Xn)
        m.update_kernel_name()
        assert got.kernel_name == m.kernel_name, (o, got.kernel_name, m.kernel_name)
        assert len(got.feature_importances) >= 2


def test_concurrent_sub_batches_are_bit_identical():
    """fit_models / fit_replicated split a large group over FIT_STREAMS engines (one CUDA stream and host thread each,
    one buffer cache per device): the models are independent, so every result is bit-identical to the single-stream fit."""
    from waveome_b200 import model_fitting as mf
    X, y = helpers.make_data(100, seed=5)
    rng = np.random.default_rng(12)
    B = 600
    Y = y[None, :] + 0.3 * rng.normal(size=(B, 100))
    assert mf.split_for_streams(B, 10 ** 6, 4) == [(0, 150), (150, 300), (300, 450), (450, 600)]

    def run(streams, replicated):
        old, mf.FIT_STREAMS = mf.FIT_STREAMS, streams
        try:
            if replicated:
                template = wb.GPR(helpers.saturated_kernel(hs=0.0), mean_function=wb.ConstantMean(0.0))
                res, models = mf.fit_replicated(X, Y, template, maxiter=25)
            else:
                models = [wb.GPR(helpers.saturated_kernel(hs=0.0) if b % 2 else helpers.all_leaf_kernel(hs=0.0),
                                 mean_function=wb.ConstantMean(0.0)) for b in range(B)]
                res = mf.fit_models(X, Y, models, maxiter=25, streams=streams)
            return res, models
        finally:
            mf.FIT_STREAMS = old

    for replicated in (False, True):
        r1, m1 = run(1, replicated)
        r4, m4 = run(4, replicated)
        for key in ("x", "f", "lml", "n_iter", "n_eval", "status"):
            assert np.array_equal(r1[key], r4[key], equal_nan=True), key
        assert r1["launches"] > 0 and r4["rounds"] >= r1["rounds"]
        for a, b in zip(m1[:5], m4[:5]):
            assert [float(p) for p in a.trainable_parameters] == [float(p) for p in b.trainable_parameters]
    assert len(mf.get_engine_pool(4)) == 4 and mf.get_engine_pool(4)[0] is mf.get_engine()


def test_fit_in_three_calls_matches_single_call(engine):
    """wv_batch_fit_lbfgs_begin / _run(min_active) / _report: the models finished when control comes back already carry
    their final results; running the stragglers to the end afterwards gives, for every model, exactly what the
    single-call fit gives."""
    from waveome_b200.engine import Batch
    n = 120
    X, y = helpers.make_data(n, seed=21)
    rng = np.random.default_rng(4)
    Y = np.stack([y + s * rng.normal(size=n) for s in np.linspace(0.0, 2.0, 24)])
    model = wb.GPR(helpers.saturated_kernel(hs=0.0), mean_function=wb.ConstantMean(0.0))
    batch = Batch(engine, X, Y, [model.program()])
    full = batch.fit()
    batch.fit_begin()
    left = batch.fit_run(6)
    part = batch.fit_report()
    assert 0 < left <= 6 and int((~part["finished"]).sum()) == left
    done = part["finished"]
    for key in ("x", "f", "lml", "n_iter", "n_eval", "status"):
        assert np.array_equal(part[key][done], full[key][done]), key
    assert np.all(part["n_eval"][~done] < full["n_eval"][~done])
    assert batch.fit_run(0) == 0
    rest = batch.fit_report()
    assert rest["finished"].all()
    for key in ("x", "f", "lml", "n_iter", "n_eval", "status"):
        assert np.array_equal(rest[key], full[key]), key
    batch.close()
