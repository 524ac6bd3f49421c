"""The device L-BFGS-B state machine (compiled for the host) must follow SciPy's L-BFGS-B trajectory:
same iterates, same number of iterations / evaluations, same stopping reason."""
import copy

import numpy as np
import pytest
import scipy.optimize as so

import gp_oracle as oracle
import helpers
import lbfgsb_host
import waveome_b200 as wb


def rosen(x):
    return so.rosen(x), so.rosen_der(x)


def _scipy(fun, x0, **kw):
    tr = []

    def f(x):
        v, g = fun(x)
        tr.append((x.copy(), float(v)))
        return v, g
    opts = dict(maxcor=10, maxiter=15000, maxfun=15000, maxls=20, ftol=2.220446049250313e-09, gtol=1e-5)
    opts.update(kw)
    res = so.minimize(f, x0, jac=True, method="L-BFGS-B", options=opts)
    return res, tr


@pytest.mark.parametrize("dim,seed", [(2, 0), (5, 1), (10, 2), (20, 3)])
def test_rosenbrock_trajectory(dim, seed):
    x0 = np.random.default_rng(seed).normal(size=dim)
    res, tr_ref = _scipy(rosen, x0)
    tr = []
    out = lbfgsb_host.minimize(rosen, x0, trace=tr)
    if dim <= 10:
        assert out["nit"] == res.nit and out["nfev"] == res.nfev
        assert len(tr) == len(tr_ref)
        for (xa, fa), (xb, fb) in zip(tr, tr_ref):
            np.testing.assert_allclose(xa, xb, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(out["x"], res.x, rtol=1e-8, atol=1e-10)
    else:
        # 20-d Rosenbrock amplifies last-bit differences (dot-product order) by ~10x every 10 iterations;
        # the first 40 evaluations must still coincide and both runs must reach the minimum.
        for (xa, fa), (xb, fb) in list(zip(tr, tr_ref))[:40]:
            np.testing.assert_allclose(xa, xb, rtol=1e-8, atol=1e-9)
        assert abs(out["nit"] - res.nit) <= 3 and out["f"] < 1e-7


def test_maxiter_and_first_point_convergence():
    x0 = np.array([-1.2, 1.0])
    res, _ = _scipy(rosen, x0, maxiter=5)
    out = lbfgsb_host.minimize(rosen, x0, maxiter=5)
    assert out["task"] == "MAXITER" and out["nit"] == res.nit == 5
    np.testing.assert_allclose(out["x"], res.x, rtol=1e-10)
    out = lbfgsb_host.minimize(rosen, np.ones(4))
    assert out["task"] == "CONV_PG" and out["nfev"] == 1


@pytest.mark.parametrize("seed", [1, 2])
def test_gp_objective_matches_scipy(seed):
    """Penalised saturated kernel (horseshoe on variances) on synthetic longitudinal data; seeds whose
    SciPy run stays finite (the horseshoe's inf/NaN regime is covered in test_nonfinite_policy)."""
    X, y = helpers.make_data(150, seed=seed)
    model = wb.GPR(helpers.saturated_kernel(), mean_function=wb.ConstantMean(0.0))
    spec = model.to_spec()

    def fun(x):
        f, g, _, _ = oracle.objective(copy.deepcopy(spec), X, y, x)
        return f, g
    x0 = oracle.pack(spec)
    res, tr_ref = _scipy(fun, x0)
    out = lbfgsb_host.minimize(fun, x0)
    assert out["nit"] == res.nit and out["nfev"] == res.nfev
    # the optimiser itself stops at a relative decrease of 2.2e-9; north-star tolerance on hyper-parameters: 1e-5
    np.testing.assert_allclose(out["f"], res.fun, rtol=1e-8)
    np.testing.assert_allclose(out["x"], res.x, rtol=1e-5, atol=1e-5)


def test_nonfinite_policy():
    """Variance underflow under the horseshoe makes TFP's log_prob gradient inf/NaN (SURVEY §7 "hard parts").
    L-BFGS-B then burns one line search on non-finite trials, restores the last finite iterate, drops its
    memory and restarts.  SciPy's C port gives up that line search at the first trial at stpmax (18 trials),
    L-BFGS-B 3.0's logic (ours) after maxls=20; both must come back to the same restored iterate."""
    def fun(x):
        if x[0] > 3.0:
            return np.nan, np.full_like(x, np.nan)
        return -x[0] + 0.5 * x[1] ** 2, np.array([-1.0, x[1]])
    tr_ref, tr = [], []

    def wrap(t):
        def f(x):
            v, g = fun(x)
            t.append(x.copy())
            return v, g
        return f
    res = so.minimize(wrap(tr_ref), np.array([0.0, 1.0]), jac=True, method="L-BFGS-B")
    out = lbfgsb_host.minimize(wrap(tr), np.array([0.0, 1.0]))
    assert out["task"] == "ABNORMAL" and res.status == 2
    np.testing.assert_allclose(out["x"], res.x, rtol=1e-12)
    finite_ref = [x for x in tr_ref if x[0] <= 3.0]
    finite = [x for x in tr if x[0] <= 3.0]
    assert len(finite) == len(finite_ref)
    for a, b in zip(finite, finite_ref):
        np.testing.assert_allclose(a, b, rtol=1e-12)
