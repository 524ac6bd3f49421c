"""Multi-output penalised model (waveome_b200/multioutput.py; reference: MultiOutputPSVGP, waveome/model_classes.py:1129-1612,
GPSearch.multioutput_penalized_optimization, waveome/model_search.py:519-573) on device="cpu": the bound against the NumPy
oracle (oracle/lmc_oracle.py), the natural-gradient step against its closed form for the Gaussian likelihood, and a small
end-to-end fit with latent-factor pruning."""
import numpy as np
import pytest
import torch

import lmc_oracle
import waveome_b200 as wb
from waveome_b200.multioutput import MultiOutputPSVGP, calculate_rank_estimate


def _data(n=70, P=4, seed=0):
    rng = np.random.default_rng(seed)
    subj = rng.integers(0, 7, size=n).astype(float)
    t = np.sort(rng.uniform(-2, 2, size=n))
    z = rng.normal(size=n)
    X = np.stack([subj, t, z], 1)
    g1, g2 = np.sin(2 * t), 0.8 * rng.normal(size=7)[subj.astype(int)]
    W = np.array([[1.0, 0.0], [0.7, 0.7], [0.0, 1.2], [-0.9, 0.4]])[:P]
    Y = np.stack([g1, g2], 1) @ W.T + 0.1 * rng.normal(size=(n, P))
    return X, Y


def _latents():
    return [wb.SquaredExponential(active_dims=[1], lengthscales=0.7), wb.Categorical(active_dims=[0]),
            wb.SquaredExponential(active_dims=[2])]


@pytest.mark.parametrize("lik,M", [("gaussian", 25), ("gaussian", 1000), ("poisson", 30)])
def test_bound_matches_the_numpy_oracle(lik, M):
    X, Y = _data()
    if lik == "poisson":
        Y = np.random.default_rng(1).poisson(np.exp(0.3 * Y)).astype(float)
    np.random.seed(3)
    m = MultiOutputPSVGP(X, Y, latent_kernels=_latents(), penalization_factor=2.0, device="cpu",
                         sparse_options={"num_inducing_points": M, "random_seed": 0}, variational_options={"likelihood": lik})
    rng = np.random.default_rng(5)
    m.kernel.W = rng.normal(scale=0.5, size=m.kernel.W.shape)
    m.likelihood_variance, m.mean_c = 0.37, 0.2
    m.q_mu = [torch.tensor(0.3 * rng.normal(size=len(q)), dtype=torch.float64) for q in m.q_mu]
    m.q_sqrt = [torch.tensor(np.tril(0.1 * rng.normal(size=q.shape)) + np.diag(0.5 + rng.random(len(q))), dtype=torch.float64)
                for q in m.q_sqrt]
    ref = lmc_oracle.lmc_elbo([k.to_spec() for k in m.kernel.kernels], [z.numpy() for z in m.Z], X, Y, m.kernel.W, m.mean_c,
                              m.likelihood_variance, [q.numpy() for q in m.q_mu], [q.numpy() for q in m.q_sqrt], likelihood=lik)
    assert abs(m.elbo() - ref) <= 1e-9 * max(1.0, abs(ref)), (m.elbo(), ref)
    assert all(len(z) == min(M, len(X)) for z in m.Z)
    # the variances of the latent kernels are frozen, the scale lives in W (freeze_variance_parameters, :1383-1386)
    assert all(not p.trainable for k in m.kernel.kernels for path, p in k.named_parameters() if "variance" in path)


def test_natural_gradient_step_is_exact_for_the_gaussian_likelihood():
    """One latent: gamma = 1 lands on the optimal q of the current hyper-parameters (a second step does not move).
    Several latents: the steps of all latents are taken simultaneously (as GPflow does), so gamma = 1 is a Jacobi
    iteration on coupled means; damped steps converge to the joint optimum, where the bound is stationary."""
    X, Y = _data(n=50)
    np.random.seed(0)
    one = MultiOutputPSVGP(X, Y, latent_kernels=_latents()[:1], device="cpu", sparse_options={"num_inducing_points": 20})
    one.kernel.W = np.array([[1.0], [0.7], [0.1], [-0.9]])
    u, W, raw, c = one._state()
    mu, sq = one._natgrad_step(u, W, raw, c, 1.0)
    one.q_mu, one.q_sqrt = [t.detach() for t in mu], [t.detach() for t in sq]
    mu2, sq2 = one._natgrad_step(u, W, raw, c, 1.0)
    assert float(torch.max(torch.abs(mu2[0] - one.q_mu[0]))) < 1e-8
    assert float(torch.max(torch.abs(sq2[0] - one.q_sqrt[0]))) < 1e-8
    m = MultiOutputPSVGP(X, Y, latent_kernels=_latents(), device="cpu", sparse_options={"num_inducing_points": 20})
    m.kernel.W = np.random.default_rng(2).normal(scale=0.6, size=m.kernel.W.shape)
    u, W, raw, c = m._state()
    for _ in range(900):
        mu, sq = m._natgrad_step(u, W, raw, c, 0.2)
        m.q_mu, m.q_sqrt = [t.detach() for t in mu], [t.detach() for t in sq]
    before = m.elbo()
    mu, sq = m._natgrad_step(u, W, raw, c, 0.2)
    assert max(float(torch.max(torch.abs(a.detach() - b))) for a, b in zip(mu, m.q_mu)) < 1e-5
    for delta in (0.01, -0.01):
        keep = m.q_mu
        m.q_mu = [q + delta for q in keep]
        assert m.elbo() < before
        m.q_mu = keep


def test_fit_recovers_two_factors_and_prunes_the_third():
    X, Y = _data(n=80, seed=4)
    np.random.seed(1)
    m = MultiOutputPSVGP(X, Y, latent_kernels=_latents(), penalization_factor=1.0, device="cpu",
                         sparse_options={"num_inducing_points": 30})
    l0 = m.training_loss()
    m.optimize_params(num_opt_iter=700, adam_learning_rate=0.02)
    assert m.training_loss() < l0 - 50 and m.fit_info["n_iter"] >= 100
    imp = np.max(np.abs(m.kernel.W), axis=0)
    assert imp[0] > 0.3 and imp[1] > 0.3 and imp[2] < 0.1, imp              # the z-latent explains nothing
    m.prune_latent_factors(threshold=0.1, optimize_after_prune=False)
    assert len(m.kernel.kernels) == 2 and m.kernel.W.shape == (4, 2)
    assert m.kernel_name == "squared_exponential[1]+categorical[0]"
    mu, var = m.predict_f(X)
    assert mu.shape == Y.shape and np.all(var > 0)
    assert np.corrcoef(mu.ravel(), Y.ravel())[0, 1] > 0.95
    ym, yv = m.predict_y(X)
    assert np.all(yv > var)


def test_gpsearch_entry_point_and_rank_estimate():
    import pandas as pd
    from waveome_b200.model_search import GPSearch
    X, Y = _data(n=40, seed=2)
    assert calculate_rank_estimate(Y, 0.90, transform_counts=False) in (1, 2)
    gps = GPSearch(pd.DataFrame(X, columns=["id", "t", "z"]), pd.DataFrame(Y, columns=[f"y{i}" for i in range(4)]),
                   unit_col="id", categorical_vars=["id"])
    gps.multioutput_penalized_optimization(num_opt_iter=30, device="cpu", random_seed=0,
                                           kernel_options={"ranks": 1, "categorical_numeric_interactions": False})
    m = gps.models["multioutput"]
    assert m.kernel.W.shape == (4, 3) and m.optimizer == "custom_multioutput" and np.isfinite(m.training_loss())
    # ranks replicate every term of the saturated kernel (reference regularization.py:27-47)
    from waveome_b200.regularization import full_kernel_build
    ks, names = full_kernel_build(cat_vars=[0], num_vars=[1], unit_idx=0, var_names=["id", "t"], ranks=2)
    assert names == ["categorical[id]_0", "categorical[id]_1", "squared_exponential[t]_0", "squared_exponential[t]_1"]
