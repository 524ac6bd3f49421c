"""Objective (B) on the engine: ``wv_batch_eval_elbo`` evaluates the whitened VGP / SVGP-with-Z = X bound at GIVEN
(q_mu, q_sqrt) (the objective of PSVGP, waveome/model_classes.py:1082-1126; gpflow SVGP.elbo, mirror
waveome/model_types_DEPR.py:126-158) against oracle/vgp_oracle.vgp_elbo, and ties it to what the engine optimises: at the
optimal q the bound equals the collapsed objective (A) the fit entry points evaluate."""
import copy

import numpy as np
import pytest

import gp_oracle as go
import helpers
import vgp_oracle as vo
import waveome_b200 as wb

pytestmark = pytest.mark.gpu


def _random_q(rng, B, n):
    q_mu = 0.3 * rng.normal(size=(B, n))
    q_sqrt = np.tril(0.1 * rng.normal(size=(B, n, n)))
    for b in range(B):
        q_sqrt[b][np.diag_indices(n)] = 0.5 + rng.random(n)
    return q_mu, q_sqrt


@pytest.mark.parametrize("n", [30, 64, 150, 333])
@pytest.mark.parametrize("lik", ["gaussian", "poisson", "negative_binomial"])
def test_elbo_at_given_q_matches_the_oracle(engine, n, lik):
    from waveome_b200.engine import Batch
    X, y = helpers.make_data(n, seed=700 + n)
    rng = np.random.default_rng(n)
    B = 2
    kern = wb.Sum([wb.Categorical(active_dims=[0]), wb.SquaredExponential(active_dims=[1]),
                   wb.Product([wb.Categorical(active_dims=[3]), wb.Matern32(active_dims=[2])])])
    for path, p in kern.named_parameters():
        if "variance" in path and p.trainable:
            p.prior = wb.Horseshoe(1.0)
    if lik == "gaussian":
        Y = np.stack([y, rng.normal(size=n)])
        model = wb.GPR(kern, mean_function=wb.ConstantMean(0.1), noise_variance=0.4)
    else:
        Y = np.stack([rng.poisson(np.exp(0.5 + 0.5 * y)).astype(float), rng.poisson(2.0, size=n).astype(float)])
        likelihood = wb.models.Poisson() if lik == "poisson" else wb.models.NegativeBinomial(alpha=0.7)
        model = wb.GPR(kern, mean_function=wb.ConstantMean(0.1), likelihood=likelihood)
    batch = Batch(engine, X, Y, [model.program()], keep_row_order=True)
    if lik != "gaussian":
        batch.set_likelihood(lik, model.likelihood.engine_param)
    x = batch.x0() + 0.2 * rng.normal(size=(B, batch.P))
    q_mu, q_sqrt = _random_q(rng, B, n)
    elbo, f, st = batch.eval_elbo(x, q_mu, q_sqrt)
    batch.close()
    assert np.all(st == 0)
    spec = model.to_spec()
    for b in range(B):
        m = go.unpack(copy.deepcopy(spec), x[b])
        if lik == "gaussian":
            ld = {"type": "gaussian", "variance": m["likelihood_variance"]["value"]}
        else:
            ld = dict(spec["likelihood"])
        ref = vo.vgp_elbo(spec, ld, X, Y[b], x[b], q_mu[b], q_sqrt[b])
        lp = sum(go.prior_logp_and_grad(p.get("prior"), p["value"])[0] for p in go.trainable_params(m))
        assert abs(elbo[b] - ref) <= 1e-9 * max(1.0, abs(ref)), (b, elbo[b], ref)
        assert abs(f[b] + ref + lp) <= 1e-9 * max(1.0, abs(ref + lp))


def test_bound_at_the_optimal_q_is_the_collapsed_objective(engine):
    """Gaussian likelihood: max_q B(theta, q) = log N(y; c, K + 1e-6 I + sigma^2 I) = objective (A) with the jitter folded
    into the noise — the statement behind fitting (A) in place of (B), checked on the engine's two entry
    points.  The optimal whitened q comes from oracle/svgp_oracle (natural parameters of the exact posterior)."""
    import torch
    import svgp_oracle as so
    from waveome_b200.engine import Batch
    n = 120
    X, y = helpers.make_data(n, seed=77)
    model = wb.GPR(helpers.saturated_kernel(hs=0.0), mean_function=wb.ConstantMean(0.2), noise_variance=0.3)
    spec = model.to_spec()
    batch = Batch(engine, X, y[None, :], [model.program()], keep_row_order=True)
    x = batch.x0() + 0.1 * np.random.default_rng(2).normal(size=(1, batch.P))
    t1, t2 = so.optimal_natural_parameters(spec, torch.tensor(X, dtype=so.DT), torch.tensor(y, dtype=so.DT),
                                           torch.tensor(x[0], dtype=so.DT))
    q_mu, q_sqrt = so.q_from_natural(t1, t2)
    elbo, f, st = batch.eval_elbo(x, q_mu.numpy()[None], q_sqrt.numpy()[None])
    # objective (A) at noise + jitter: shift the unconstrained noise parameter so that its constrained value grows by 1e-6
    prog = model.program()
    k = [i for i, p in enumerate(prog.x_params) if p is model.likelihood.variance][0]
    s2 = np.logaddexp(0.0, x[0, k]) + 1e-6                       # softplus + lower bound 1e-6
    xa = x.copy()
    xa[0, k] = np.log(np.expm1(s2 + 1e-6 - 1e-6))                # softplus^-1(s2 + jitter - shift)
    fa, _, lml, sta = batch.eval(xa)
    batch.close()
    assert st[0] == 0 and sta[0] == 0
    assert abs(elbo[0] - lml[0]) <= 1e-8 * abs(lml[0]), (elbo[0], lml[0])
    # any other q is worse
    q2 = q_sqrt.numpy().copy()
    q2[np.diag_indices(n)] *= 1.05
    b2 = Batch(engine, X, y[None, :], [model.program()], keep_row_order=True)
    worse, _, _ = b2.eval_elbo(x, q_mu.numpy()[None], q2[None])
    b2.close()
    assert worse[0] < elbo[0]


def test_row_order_is_required(engine):
    from waveome_b200.engine import Batch, EngineError
    X, y = helpers.make_data(20, seed=1)
    model = wb.GPR(wb.SquaredExponential(active_dims=[1]), mean_function=wb.ConstantMean(0.0))
    batch = Batch(engine, X, y[None, :], [model.program()])
    with pytest.raises(EngineError):
        batch.eval_elbo(batch.x0(), np.zeros((1, 20)), np.eye(20)[None])
    batch.close()
