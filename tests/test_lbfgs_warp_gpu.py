"""The warp-parallel form of the device L-BFGS-B step (csrc/wv_lbfgsb.h, policy WvExWarp) against the thread-per-model
form (policy WvExSerial, the code the CPU suite checks against SciPy): every value is computed by one lane with the serial
code's operations in the serial code's order, so whole fits must agree BIT FOR BIT -- iterates, objective values, iteration
and evaluation counts, exit states (reference: scipy L-BFGS-B behind gpflow.optimizers.Scipy,
waveome/model_fitting.py:276-281)."""
import os

import numpy as np
import pytest

import helpers
import waveome_b200 as wb

pytestmark = pytest.mark.gpu


def _fit_both(engine, X, Y, models, **opts):
    from waveome_b200.engine import Batch
    out = []
    for serial in ("1", "0"):
        os.environ["WV_LB_SERIAL"] = serial
        try:
            batch = Batch(engine, X, Y, [m.program() for m in models]) if len(models) > 1 else \
                Batch(engine, X, Y, [models[0].program()])
            out.append(batch.fit(**opts))
            batch.close()
        finally:
            os.environ.pop("WV_LB_SERIAL", None)
    return out


def _assert_identical(a, b):
    for key in ("x", "f", "n_iter", "n_eval", "status"):
        assert np.array_equal(np.asarray(a[key]), np.asarray(b[key]), equal_nan=True), key


def test_saturated_kernel_fits_identical(engine):
    n = 150
    X, y = helpers.make_data(n, seed=3)
    rng = np.random.default_rng(8)
    Y = np.stack([y + s * rng.normal(size=n) for s in (0.0, 0.1, 0.5, 2.0)] + [rng.normal(size=n) for _ in range(28)])
    for hs, maxcor in ((0.0, 10), (1.0, 10), (1.0, 3), (0.0, 20)):
        model = wb.GPR(helpers.saturated_kernel(hs=hs), mean_function=wb.ConstantMean(0.0))
        ser, par = _fit_both(engine, X, Y, [model], maxcor=maxcor)
        _assert_identical(ser, par)
        assert ser["n_iter"].max() > maxcor          # the memory wrapped (shift of SS / SY / YY exercised)


def test_mixed_structures_and_failed_line_searches_identical(engine):
    """One model per structure (different parameter counts), outcomes that end in ABNORMAL line searches included."""
    n = 90
    X, y = helpers.make_data(n, seed=11)
    rng = np.random.default_rng(1)
    kerns = [wb.SquaredExponential(active_dims=[1]), wb.Matern12(active_dims=[2]),
             wb.Lin(active_dims=[1]) + wb.Periodic(wb.SquaredExponential(active_dims=[2])),
             wb.Categorical(active_dims=[0]) * wb.SquaredExponential(active_dims=[1]) + wb.Lin(active_dims=[2]),
             wb.Periodic(wb.SquaredExponential(active_dims=[1])) * wb.SquaredExponential(active_dims=[1])]
    models, ys = [], []
    for k in kerns:
        for rep in range(6):
            models.append(wb.GPR(wb.deepcopy(k), mean_function=wb.ConstantMean(0.0)))
            ys.append(y * (rep % 3) + rng.normal(size=n) * (0.05 + rep))
    from waveome_b200.model_fitting import fit_models
    res = []
    for serial in ("1", "0"):
        os.environ["WV_LB_SERIAL"] = serial
        try:
            res.append(fit_models(X, np.stack(ys), [wb.deepcopy(m) for m in models], engine=engine))
        finally:
            os.environ.pop("WV_LB_SERIAL", None)
    _assert_identical(res[0], res[1])
    assert len(set(int(v) for v in res[0]["n_iter"])) > 5
