"""The ENGINE against values recorded by real GPflow runs in the reference's notebooks (tests/reference_pins.py): the
same pins as tests/test_reference_pins_cpu.py, through the C-ABI on the GPU."""
import numpy as np
import pytest

import reference_pins as rp
import waveome_b200 as wb
from waveome_b200.model_fitting import fit_models

pytestmark = pytest.mark.gpu


def _fit(engine, kernel, X, Y):
    m = wb.GPR(kernel)                     # gpflow.models.GPR(data, kernel, mean_function=None), noise variance 1.0
    res = fit_models(X, Y[None, :], [m], engine=engine, maxiter=100)
    assert res["status"][0] == 0
    return m, res


def test_se_fit_reproduces_recorded_lml(engine):
    X, Y = rp.colab_sine10()
    m, res = _fit(engine, wb.SquaredExponential(active_dims=[0]), X, Y)
    assert abs(res["lml"][0] - rp.COLAB_SE["lml"]) < 1e-10
    assert abs(float(m.kernel.variance) - rp.COLAB_SE["variance"]) < 1e-6
    assert abs(float(m.kernel.lengthscales) - rp.COLAB_SE["lengthscales"]) < 1e-6
    assert abs(float(m.likelihood.variance) - rp.COLAB_SE["noise"]) < 1e-7


def test_matern12_periodic_recorded_comment(engine):
    X, Y = rp.colab_sine10()
    for kern, want in ((wb.Matern12(active_dims=[0]), rp.COLAB_LML_COMMENT["matern12"]),
                       (wb.Periodic(wb.SquaredExponential(active_dims=[0])), rp.COLAB_LML_COMMENT["periodic"])):
        _m, res = _fit(engine, kern, X, Y)
        digits = len(str(want).split(".")[1])
        assert abs(res["lml"][0] - want) < 10.0 ** (-digits), (res["lml"][0], want)


def test_matern52_recorded_summary(engine):
    X, Y = rp.basic_inline12()
    m, _res = _fit(engine, wb.Matern52(active_dims=[0]), X, Y)
    assert abs(float(m.kernel.variance) - rp.BASIC_M52["variance"]) < 5e-6
    assert abs(float(m.kernel.lengthscales) - rp.BASIC_M52["lengthscales"]) < 5e-7
    assert abs(float(m.likelihood.variance) - rp.BASIC_M52["noise"]) < 5e-8


def test_constant_mean_and_bernoulli_recorded_losses(engine):
    from waveome_b200.engine import Batch
    from waveome_b200.models import make_likelihood
    X, Y, Yb = rp.simple_regression()
    for c, var, ls, noise, loss in rp.SIMPLE_GAUSSIAN:
        m = wb.GPR(wb.SquaredExponential(active_dims=[0], variance=var, lengthscales=ls), mean_function=wb.ConstantMean(c),
                   noise_variance=noise)
        batch = Batch(engine, X, Y[None, :], [m.program()])
        _f, _g, lml, st = batch.eval(batch.x0())
        batch.close()
        assert st[0] == 0 and lml[0] >= -loss - 1e-9 and lml[0] + loss < 0.05, (lml[0], loss)
    b = rp.SIMPLE_BERNOULLI
    m = wb.GPR(wb.SquaredExponential(active_dims=[0], variance=b["variance"], lengthscales=b["lengthscales"]),
               mean_function=wb.ConstantMean(b["c"]))
    m.likelihood = make_likelihood("bernoulli")
    batch = Batch(engine, X, Yb[None, :], [m.program()])
    batch.set_likelihood("bernoulli")
    _f, _g, lml, st = batch.eval(batch.x0())
    batch.close()
    assert st[0] == 0 and lml[0] >= -b["loss"] - 1e-9 and lml[0] + b["loss"] < 2e-3, (lml[0], b["loss"])


def test_categorical_plus_matern12_density_recorded(engine):
    from waveome_b200.engine import Batch
    X, y, loglik, noise = rp.simulated_y1()
    k = wb.Sum([wb.Matern12(active_dims=[2]), wb.Categorical(active_dims=[0], variance=2.0)])
    m = wb.GPR(k, noise_variance=noise)
    batch = Batch(engine, X, y[None, :], [m.program()])
    _f, _g, lml, st = batch.eval(batch.x0())
    batch.close()
    assert st[0] == 0 and abs(lml[0] - loglik) < 1e-9 * abs(loglik), (lml[0], loglik)
