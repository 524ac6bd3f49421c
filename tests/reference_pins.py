"""Numbers RECORDED BY THE REFERENCE'S OWN NOTEBOOKS (outputs of real GPflow runs committed upstream), with the seeded
data generators of the same cells restated.  They are the only reference-computed golden vectors that exist for this
path (the reference ships no tests), and they pin the oracle and the engine to GPflow itself:

* kernel_learning/gpflow_basic_colab.ipynb cells 3-7: GPR(SquaredExponential), 10 points, Scipy L-BFGS-B maxiter=100
  from the defaults: print_summary values and ``m.log_marginal_likelihood()`` = -9.914289155637; the cell's comment
  records the same quantity for Matern12 (-10.68) and Periodic(SE) (-12.242);
* kernel_learning/gpflow_basic.ipynb cells 28-31: GPR(Matern52), 12 inline points, fitted values of the summary table;
* examples/simulations/simple_regression_different_models.ipynb cells 1-5: BaseGP (whitened SVGP with Z = X, Gaussian
  likelihood, Constant mean) trained with Adam + natural gradients: recorded training losses at the recorded
  hyper-parameters (an ELBO, hence <= the exact log marginal likelihood at the same hyper-parameters);
* same notebook, cells 9-11: VarGP(likelihood='bernoulli'): recorded loss, hyper-parameters, q_mu[0], q_sqrt[0, 0];
* kernel_learning/data_generation.ipynb cell 14: log density of simulated_data.csv's y1_obs under
  Matern12[time] + Categorical(variance=2)[unit] + 1e-2 I (fixture tests/golden/ref_simulated_y1.json, extracted by
  tests/golden/make_reference_pins.py): 3 log(126) - 2 loglik = 169.71614261744028.
"""
import json
import math
import os

import numpy as np


def simulated_y1():
    """-> X [126, 3] (unit, treatment, time), y, recorded log density"""
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_simulated_y1.json")) as fh:
        d = json.load(fh)
    X, y = np.array(d["X"]), np.array(d["y1_obs"])
    loglik = (d["k"] * math.log(len(y)) - d["recorded_k_log_n_minus_2_loglik"]) / 2.0
    return X, y, loglik, d["noise_variance"]


def colab_sine10():
    """gpflow_basic_colab.ipynb cell 3"""
    np.random.seed(9102)
    X = np.random.uniform(low=0, high=10, size=10)
    Y = np.sin(X) + np.random.normal(scale=0.5, size=10)
    return X.reshape(-1, 1), Y


COLAB_SE = dict(variance=0.749872, lengthscales=0.740194, noise=0.0379764, lml=-9.914289155637)
COLAB_LML_COMMENT = {"matern12": -10.68, "periodic": -12.242, "squared_exponential": -9.91}


def basic_inline12():
    """gpflow_basic.ipynb cell 28"""
    X = np.array([0, 0.05, 0.05, 0.15, 0.17, 0.21, 0.61, 0.79, 0.8, 0.9, 0.9, 0.95]).reshape(-1, 1)
    Y = np.array([3.6, 3.7, 3.5, 3.1, 3.2, 3.8, 3.5, 3.6, 3.0, 1.8, 1.6, 1.4])
    return X, Y


BASIC_M52 = dict(variance=7.76607, lengthscales=0.492535, noise=0.0969323)


def simple_regression():
    """simple_regression_different_models.ipynb cells 1 and 9: (X, Y gaussian, Y bernoulli)"""
    import scipy.special
    np.random.seed(9102)
    N = 100
    X = np.random.uniform(low=-5, high=5, size=(N, 1))
    Y = np.sin(X)
    Y[X >= 3] = 0.5
    Y += np.random.normal(scale=.1, size=(N, 1))
    np.random.seed(9102)
    Yb = np.random.binomial(n=1, p=scipy.special.expit(Y))
    return X, Y[:, 0], Yb[:, 0].astype(float)


SIMPLE_Z0 = -4.02011          # first row of the printed inducing_variable.Z (= X)
# (mean c, kernel variance, lengthscale, noise variance, recorded final training loss = -ELBO)
SIMPLE_GAUSSIAN = [(0.32965214640943175, 0.24707589611807843, 1.37048, 0.00846385650648054, -72.4484598846995),
                   (0.4093631338560377, 0.39237571168426433, 1.37526, 0.008144713086723682, -73.22063598606498)]
SIMPLE_BERNOULLI = dict(c=-0.8448049208938503, variance=4.60216, lengthscales=1.02164, loss=22.81370224103085,
                        q_mu0=1.69420000, q_sqrt00=3.84843436e-01)
