"""Run-time specialised Gram / gradient kernels (specialize.py -> NVRTC) against the oracle (rel 1e-9, the north star's
bar) and against the interpreter kernels on identical inputs: structures with every covered leaf type, products of squared
exponentials, frozen parameters, component masks, the variational (site) path, tile-boundary sizes."""
import copy

import numpy as np
import pytest

import gp_oracle as oracle
import helpers
import waveome_b200 as wb
from test_specialize_cpu import _mixed_kernel

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _check(engine, kernel, n, seed, nb=3):
    from waveome_b200.engine import Batch
    X, y = helpers.make_data(n, seed=seed)
    rng = np.random.default_rng(seed)
    Y = np.stack([y] + [rng.normal(size=n) for _ in range(nb - 1)])
    model = wb.GPR(kernel, mean_function=wb.ConstantMean(0.1), noise_variance=0.3)
    batch = Batch(engine, X, Y, [model.program()], specialize=True)
    assert batch.specialized
    x = batch.x0() + 0.3 * rng.normal(size=(nb, batch.P))
    f, g, lml, st = batch.eval(x)
    batch.unspecialize()
    f2, g2, lml2, st2 = batch.eval(x)
    batch.close()
    assert np.all(st == 0) and np.all(st2 == 0)
    np.testing.assert_allclose(lml, lml2, rtol=1e-12)
    np.testing.assert_allclose(g, g2, rtol=0, atol=1e-11 * np.max(np.abs(g2)))
    for b in range(nb):
        fo, go, lo, _ = oracle.objective(copy.deepcopy(model.to_spec()), X, Y[b], x[b])
        assert abs(lml[b] - lo) <= RTOL * abs(lo), (b, lml[b], lo)
        assert abs(f[b] - fo) <= RTOL * max(1.0, abs(fo))
        assert np.max(np.abs(g[b] - go)) <= RTOL * np.max(np.abs(go)), (b, g[b], go)


@pytest.mark.parametrize("n", [1, 7, 63, 64, 65, 127, 128, 200, 333])
def test_saturated_kernel_sizes(engine, n):
    _check(engine, helpers.saturated_kernel(), n, seed=500 + n)


@pytest.mark.parametrize("n", [50, 130, 257])
def test_every_covered_leaf_type(engine, n):
    _check(engine, _mixed_kernel(), n, seed=600 + n)


def test_config3_size_and_structure(engine):
    """n = 600, D = 5, the benchmark's saturated kernel (9 components, P = 17)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_c3_golden import c3_setup
    from waveome_b200.engine import Batch
    gps, model = c3_setup(4)
    Xn = gps.X.to_numpy(dtype=np.float64)
    Yn = np.ascontiguousarray(gps.Y.to_numpy(dtype=np.float64).T)
    batch = Batch(engine, Xn, Yn, [model.program()], specialize=True)
    assert batch.specialized
    x = batch.x0() + 0.2 * np.random.default_rng(3).normal(size=(4, batch.P))
    f, g, lml, st = batch.eval(x)
    batch.close()
    for b in range(4):
        fo, go, lo, _ = oracle.objective(copy.deepcopy(model.to_spec()), Xn, Yn[b], x[b])
        assert st[b] == 0
        assert abs(lml[b] - lo) <= RTOL * abs(lo)
        assert abs(f[b] - fo) <= RTOL * abs(fo)
        assert np.max(np.abs(g[b] - go)) <= RTOL * np.max(np.abs(go))


def test_component_masks_and_fit_identical_structure(engine):
    """feature-importance batches (component masks) and a complete fit run on the specialised kernels"""
    from waveome_b200.engine import Batch
    n = 150
    X, y = helpers.make_data(n, seed=9)
    model = wb.GPR(helpers.saturated_kernel(hs=0.0), mean_function=wb.ConstantMean(0.0))
    prog = model.program()
    Y = np.stack([y] * 4)
    mask = np.array([0xffffffff, 0xffffffff & ~1, 0xffffffff & ~(1 << 3), 0b101], dtype=np.uint32)
    out = {}
    for spec in (True, False):
        batch = Batch(engine, X, Y, [prog], specialize=spec)
        assert batch.specialized == spec
        batch.set_component_mask(mask)
        out[spec] = batch.eval(batch.x0())
        batch.close()
    np.testing.assert_allclose(out[True][2], out[False][2], rtol=1e-12)
    np.testing.assert_allclose(out[True][1], out[False][1], rtol=0, atol=1e-11 * np.max(np.abs(out[False][1])))
    fits = {}
    for spec in (True, False):
        batch = Batch(engine, X, Y[:1], [prog], specialize=spec)
        fits[spec] = batch.fit()
        batch.close()
    ref = oracle.fit(model.to_spec(), X, y)
    for spec in (True, False):
        assert fits[spec]["status"][0] == 0
        assert abs(fits[spec]["f"][0] - ref["f"]) <= 1e-8 * max(1.0, abs(ref["f"]))
        np.testing.assert_allclose(fits[spec]["x"][0], ref["x"], rtol=1e-5, atol=1e-5)


def test_variational_path_on_specialised_kernels(engine):
    """Poisson counts: the site iteration (heteroscedastic Gram with per-row noise) through both kernel families"""
    from waveome_b200.engine import Batch
    n = 120
    X, y = helpers.make_data(n, seed=4)
    rng = np.random.default_rng(0)
    counts = rng.poisson(np.exp(0.5 + 0.5 * y)).astype(float)
    k = wb.Sum([wb.Categorical(active_dims=[0]), wb.SquaredExponential(active_dims=[1])])
    model = wb.GPR(k, mean_function=wb.ConstantMean(0.0), likelihood=wb.models.Poisson())
    out = {}
    for spec in (True, False):
        batch = Batch(engine, X, counts[None, :], [model.program()], specialize=spec)
        batch.set_likelihood("poisson")
        out[spec] = batch.eval(batch.x0())
        batch.close()
    assert out[True][3][0] == 0 and out[False][3][0] == 0
    np.testing.assert_allclose(out[True][2], out[False][2], rtol=1e-9)
    np.testing.assert_allclose(out[True][1], out[False][1], rtol=0, atol=1e-7 * np.max(np.abs(out[False][1])))
