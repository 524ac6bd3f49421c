"""CPU oracle of the multi-output bound: NumPy restatement of the whitened SVGP ELBO of GPflow's
``LinearCoregionalization`` kernel with ``SeparateIndependentInducingVariables`` — the objective of ``MultiOutputPSVGP``
(waveome/model_classes.py:1129-1386).  TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (the reference holds no recorded value
of this objective and GPflow cannot be installed here): restated from gpflow.conditionals (independent latents, mixed
by W, full_output_cov = False) and gpflow.kullback_leiblers.gauss_kl (white).  Kernel matrices come from gp_oracle.

    for every latent q:  Luu Luu^T = k_q(Z_q, Z_q) + 1e-6 I,  A = Luu^-1 k_q(Z_q, X)
                         mean_q = A^T q_mu_q,  var_q = k_q(x, x) - colsum(A^2) + colsum((tril(q_sqrt_q)^T A)^2)
    f mean = [mean_q] W^T + c,  f var = [var_q] (W o W)^T
    elbo = sum_{i,p} E log p(y_ip | f_ip) - sum_q 1/2 (|q_mu_q|^2 + |tril q_sqrt_q|_F^2 - M - 2 sum log |diag q_sqrt_q|)
"""
import numpy as np
import scipy.linalg as sla
from scipy.special import gammaln

import gp_oracle as go


def lmc_elbo(kernel_specs, Zs, X, Y, W, c, noise_variance, q_mus, q_sqrts, likelihood="gaussian"):
    X, Y, W = np.asarray(X, float), np.asarray(Y, float), np.asarray(W, float)
    n = X.shape[0]
    means, vars_ = [], []
    kl = 0.0
    for spec, Z, q_mu, q_sqrt in zip(kernel_specs, Zs, q_mus, q_sqrts):
        Z = np.asarray(Z, float)
        M = Z.shape[0]
        Kall, _ = go.kernel_K_and_grads(spec, np.vstack([Z, X]), want_grads=False)
        Kuu, Kuf, kff = Kall[:M, :M] + 1e-6 * np.eye(M), Kall[:M, M:], np.diag(Kall[M:, M:])
        L = np.linalg.cholesky(Kuu)
        A = sla.solve_triangular(L, Kuf, lower=True)
        S = np.tril(q_sqrt)
        SA = S.T @ A
        means.append(A.T @ q_mu)
        vars_.append(kff - np.sum(A * A, 0) + np.sum(SA * SA, 0))
        kl += 0.5 * (np.sum(q_mu ** 2) + np.sum(S ** 2) - M - 2.0 * np.sum(np.log(np.abs(np.diag(S)))))
    G, V = np.stack(means, 1), np.stack(vars_, 1)
    fm, fv = G @ W.T + c, V @ (W * W).T
    if likelihood == "gaussian":
        ve = -0.5 * np.log(2 * np.pi * noise_variance) - ((Y - fm) ** 2 + fv) / (2 * noise_variance)
    else:
        ve = Y * fm - np.exp(fm + 0.5 * fv) - gammaln(Y + 1.0)
    return float(np.sum(ve) - kl)
