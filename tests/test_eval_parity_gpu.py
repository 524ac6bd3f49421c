"""CUDA path vs the oracle through the C ABI: LML and gradients to rel 1e-9 (north-star tolerance) on identical
inputs, edge cases, status bits, and size-independent properties at the full BASELINE size."""
import copy
import json
import os

import numpy as np
import pytest

import gp_oracle as oracle
import helpers
import waveome_b200 as wb

pytestmark = pytest.mark.gpu
RTOL = 1e-9          # BASELINE.json north_star: "LML and gradients to rel 1e-9"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _check(batch, model, X, Y, x, idx=None):
    f, g, lml, st = batch.eval(x)
    for b in (range(len(Y)) if idx is None else idx):
        fo, go, lo, _ = oracle.objective(copy.deepcopy(model.to_spec()), X, Y[b], x[b][: len(oracle.pack(model.to_spec()))])
        assert st[b] == 0
        assert abs(lml[b] - lo) <= RTOL * abs(lo), (b, lml[b], lo)
        assert abs(f[b] - fo) <= RTOL * abs(fo), (b, f[b], fo)
        scale = np.max(np.abs(go))
        assert np.max(np.abs(g[b][: len(go)] - go)) <= RTOL * scale, (b, g[b], go)


@pytest.mark.parametrize("n", [1, 2, 7, 31, 62, 63, 64, 65, 126, 127, 128, 150, 191, 200, 333])
@pytest.mark.parametrize("kern", ["all", "sat"])
def test_eval_matches_oracle(engine, n, kern):
    from waveome_b200.engine import Batch
    X, y = helpers.make_data(n, seed=100 + n)
    k = helpers.all_leaf_kernel() if kern == "all" else helpers.saturated_kernel()
    model = wb.GPR(k, mean_function=wb.ConstantMean(0.1), noise_variance=0.5)
    rng = np.random.default_rng(n)
    Y = np.stack([y, 0.5 * y + 0.1, rng.normal(size=n)])
    batch = Batch(engine, X, Y, [model.program()])
    x = batch.x0() + 0.3 * rng.normal(size=(3, batch.P))
    _check(batch, model, X, Y, x)
    batch.close()


def test_golden_vectors(engine):
    from waveome_b200.engine import Batch
    with open(os.path.join(GOLDEN, "eval_cases.json")) as fh:
        cases = json.load(fh)
    for c in cases:
        X, y, x = np.array(c["X"]), np.array(c["y"]), np.array(c["x"])
        n = len(y)
        k = helpers.all_leaf_kernel() if c["name"].startswith("all") else helpers.saturated_kernel()
        model = wb.GPR(k, mean_function=wb.ConstantMean(0.1), noise_variance=0.6)
        assert model.to_spec() == c["spec"]
        batch = Batch(engine, X, y[None, :], [model.program()])
        f, g, lml, st = batch.eval(x[None, :])
        assert st[0] == 0
        assert abs(lml[0] - c["lml"]) <= RTOL * abs(c["lml"])
        assert abs(f[0] - c["f"]) <= RTOL * abs(c["f"])
        np.testing.assert_allclose(g[0], np.array(c["grad"]), rtol=0, atol=RTOL * np.max(np.abs(c["grad"])))
        batch.close()


def test_heterogeneous_programs_zero_mean_and_frozen(engine):
    """One batch, different kernel programs per model (the search path's level batches)."""
    from waveome_b200.engine import Batch
    n = 97
    X, y = helpers.make_data(n, seed=7)
    k1 = wb.SquaredExponential(active_dims=[1]) + wb.Categorical(active_dims=[0])
    k2 = wb.Periodic(wb.SquaredExponential(active_dims=[1])) * wb.Matern12(active_dims=[2])
    wb.set_trainable(k2.kernels[1].variance, False)
    k3 = wb.Constant(variance=1e-6); wb.set_trainable(k3.variance, False)     # the search's frozen "empty" candidate
    models = [wb.GPR(k1), wb.GPR(k2, mean_function=wb.ConstantMean(0.3)), wb.GPR(k3, mean_function=wb.ConstantMean())]
    progs = [m.program() for m in models]
    Y = np.stack([y, y, y])
    batch = Batch(engine, X, Y, progs, prog_id=[0, 1, 2])
    x = batch.x0()
    x[:, :2] += 0.2
    f, g, lml, st = batch.eval(x)
    for b, m in enumerate(models):
        spec = m.to_spec()
        nx = len(oracle.pack(spec))
        fo, go, lo, _ = oracle.objective(copy.deepcopy(spec), X, y, x[b, :nx])
        assert st[b] == 0 and abs(lml[b] - lo) <= RTOL * abs(lo)
        assert np.max(np.abs(g[b, :nx] - go)) <= RTOL * np.max(np.abs(go))
        assert np.all(g[b, nx:] == 0.0)
    batch.close()


def test_status_bits(engine):
    from waveome_b200.engine import Batch
    n = 40
    X, y = helpers.make_data(n, seed=3)
    # (1) not positive definite: duplicated rows with a noise floor of 1e-6 and a huge linear variance
    Xd = np.repeat(X[:20], 2, axis=0)
    k = wb.Lin(active_dims=[1], variance=1e12)
    m = wb.GPR(k, noise_variance=1e-6 + 1e-12)
    batch = Batch(engine, Xd, y[None, :], [m.program()])
    f, g, lml, st = batch.eval(batch.x0())
    try:
        oracle.objective(copy.deepcopy(m.to_spec()), Xd, y, oracle.pack(m.to_spec()))
        oracle_failed = False
    except oracle.CholeskyFailure:
        oracle_failed = True
    assert bool(st[0] & 1) == oracle_failed
    batch.close()
    # (2) horseshoe underflow: variance ~ exp(-400) -> log prior +inf -> non-finite objective, no Cholesky failure
    k = wb.SquaredExponential(active_dims=[1]); k.variance.prior = wb.Horseshoe(1.0)
    m = wb.GPR(k)
    batch = Batch(engine, X, y[None, :], [m.program()])
    x = batch.x0(); x[0, 0] = -400.0
    f, g, lml, st = batch.eval(x)
    fo, go, lo, _ = oracle.objective(copy.deepcopy(m.to_spec()), X, y, x[0])
    assert st[0] == 2 and not np.isfinite(f[0]) and not np.isfinite(fo)
    assert abs(lml[0] - lo) <= RTOL * abs(lo)          # the likelihood part itself stays finite and equal
    batch.close()


def test_full_size_properties(engine):
    """BASELINE configs[2] shape (n = 600, saturated 9-component kernel): properties that need no oracle run at
    size, plus a few models checked against the oracle."""
    from waveome_b200 import datasets
    from waveome_b200.engine import Batch
    from waveome_b200.model_search import GPSearch
    from waveome_b200.regularization import full_kernel_build
    B = 64
    X, Y = datasets.ihmp_scale(n_outcomes=B)
    gps = GPSearch(X, Y, unit_col="participant", categorical_vars=["participant", "sex", "site"], Y_transform="standardize")
    k = full_kernel_build(cat_vars=gps.cat_idx, num_vars=gps.cont_idx, unit_idx=gps.unit_idx, return_sum=True)
    model = wb.models.PenalizedGPR(k, mean_function=wb.ConstantMean())
    Xn, Yn = gps.X.to_numpy(), np.ascontiguousarray(gps.Y.to_numpy().T)
    batch = Batch(engine, Xn, Yn, [model.program()])
    rng = np.random.default_rng(0)
    x = batch.x0() + 0.2 * rng.normal(size=(B, batch.P))
    f, g, lml, st = batch.eval(x)
    assert np.all(st == 0)
    _check(batch, model, Xn, Yn, x, idx=[0, 17, 63])
    # determinism: bit-identical on repetition (fixed-order reductions, no atomics on the data path)
    f2, g2, lml2, _ = batch.eval(x)
    assert np.array_equal(f, f2) and np.array_equal(g, g2) and np.array_equal(lml, lml2)
    # batch independence: a model's result does not depend on its neighbours or its slot
    perm = rng.permutation(B)
    b2 = Batch(engine, Xn, Yn[perm], [model.program()])
    f3, g3, lml3, _ = b2.eval(x[perm])
    assert np.array_equal(f3, f[perm]) and np.array_equal(g3, g[perm])
    b2.close()
    # ... nor on the schedule: alone, a model takes the fused one-launch Cholesky step; among 64 its early steps run as
    # a diagonal launch + the panel-only kernel -- same arithmetic, bit-identical results
    b1 = Batch(engine, Xn, Yn[17:18], [model.program()])
    f1, g1, lml1, _ = b1.eval(x[17:18])
    assert f1[0] == f[17] and np.array_equal(g1[0], g[17]) and lml1[0] == lml[17]
    b1.close()
    # row-permutation invariance of the marginal likelihood (up to summation order)
    rp = rng.permutation(Xn.shape[0])
    b3 = Batch(engine, Xn[rp], Yn[:, rp], [model.program()])
    f4, g4, lml4, _ = b3.eval(x)
    np.testing.assert_allclose(lml4, lml, rtol=1e-10)
    np.testing.assert_allclose(g4, g, rtol=0, atol=1e-9 * np.max(np.abs(g)))
    b3.close()
    # directional derivative vs central differences of the device objective itself
    d = rng.normal(size=x.shape)
    h = 1e-6
    fp, _, _, _ = batch.eval(x + h * d)
    fm, _, _, _ = batch.eval(x - h * d)
    num = (fp - fm) / (2 * h)
    ana = np.sum(g * d, axis=1)
    np.testing.assert_allclose(num, ana, rtol=2e-6, atol=1e-6)
    batch.close()


def test_device_pointer_entry(engine):
    """wv_batch_eval_device with torch CUDA tensors (DLPack-compatible buffers), no host staging."""
    import torch
    from waveome_b200.engine import Batch
    n = 80
    X, y = helpers.make_data(n, seed=9)
    model = wb.GPR(helpers.saturated_kernel(), mean_function=wb.ConstantMean())
    batch = Batch(engine, X, np.stack([y, -y]), [model.program()])
    x = batch.x0()
    f_h, g_h, lml_h, st_h = batch.eval(x)
    dev = torch.device("cuda", engine.device)
    xd = torch.tensor(x, device=dev)
    fd = torch.empty(2, dtype=torch.float64, device=dev); gd = torch.empty_like(xd)
    ld = torch.empty(2, dtype=torch.float64, device=dev); sd = torch.empty(2, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    batch.eval_device(xd, fd, gd, ld, sd)
    torch.cuda.ExternalStream(engine.stream).synchronize()
    assert np.array_equal(fd.cpu().numpy(), f_h) and np.array_equal(gd.cpu().numpy(), g_h)
    batch.close()


def test_one_launch_factorisation_paths_are_bit_identical(engine):
    """A batch that declares itself alone on the device (wv_batch_set_solo) factorises mid-size batches with the persistent
    one-launch kernels (wv_chol_all_kernel, wv_trtri_all_kernel) instead of a launch pair per tile column and a launch per
    inverse step: same tile bodies, same arithmetic -- every value, gradient and status must agree BIT FOR BIT, for two- and
    ten-tile models, with and without the full active list."""
    from waveome_b200.engine import Batch
    rng = np.random.default_rng(3)
    for n, B in ((100, 320), (600, 60)):
        X, y = helpers.make_data(n, seed=n)
        Y = y[None, :] + 0.3 * rng.normal(size=(B, n))
        model = wb.GPR(helpers.saturated_kernel(hs=0.0), mean_function=wb.ConstantMean(0.0))
        out = []
        for solo in (False, True):
            batch = Batch(engine, X, Y, [model.program()])
            batch.set_solo(solo)
            x = batch.x0() + 0.1 * np.random.default_rng(1).normal(size=(B, batch.P))
            out.append(batch.eval(x))
            fit = batch.fit(maxiter=12)              # the active list shrinks: both forms see ragged lists too
            out[-1] = out[-1] + (fit["x"], fit["f"], fit["n_eval"])
            batch.close()
        for a, b in zip(*out):
            assert np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)
