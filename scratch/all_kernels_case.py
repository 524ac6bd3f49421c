"""Small invocations of every kernel family in one process (written for compute-sanitizer, which is closed on this
pool; still a quick all-kernels check): batched schedule (evaluation, a short L-BFGS fit, alpha / K^-1 diagonal / predictions, component masks), the large-n
schedule forced at a small size, and the site iteration of the count likelihoods.  NumPy only (no torch import)."""
import sys
import numpy as np
sys.path[:0] = [".", "oracle", "tests"]
import helpers
import waveome_b200 as wb
from waveome_b200.engine import Batch, Engine
from waveome_b200.models import make_likelihood

which = sys.argv[1] if len(sys.argv) > 1 else "all"
eng = Engine(0)
rng = np.random.default_rng(0)
if which in ("all", "batched"):
    X, y = helpers.make_data(150, seed=3)                       # nt = 3 tiles, ragged last tile
    Y = np.stack([y, 0.5 * y + 0.2, rng.normal(size=150)])
    m = wb.GPR(helpers.all_leaf_kernel(), mean_function=wb.ConstantMean(0.0))
    b = Batch(eng, X, Y, [m.program()])
    x = b.x0()
    f, g, lml, st = b.eval(x)
    assert np.all(st == 0) and np.all(np.isfinite(g)), (st, f)
    b.alpha(); b.kinv_diag()
    b.set_component_mask(np.array([0xFFFFFFFF, 0xFFFFFFFE, 0x5], np.uint32))
    b.eval(x)
    b.set_component_mask(np.full(3, 0xFFFFFFFF, np.uint32))
    r = b.fit(x, maxiter=4, maxfun=12)
    print("batched ok", f, r["n_eval"], flush=True)
    b.close()
    m2 = wb.GPR(helpers.saturated_kernel(), mean_function=wb.ConstantMean(0.1), noise_variance=0.4)
    b = Batch(eng, X, Y[:2], [m2.program()])
    b.eval(b.x0())
    from waveome_b200 import postfit
    mu, var = postfit.predict_f(m2, X, Y[0], X[:37] + 0.1, engine=eng)
    assert np.all(var > 0)
    print("predict ok", flush=True)
    b.close()
if which in ("all", "large"):
    eng2 = Engine(0, large_n_tiles=2)
    X, y = helpers.make_data(333, seed=9)
    m = wb.GPR(helpers.saturated_kernel(), mean_function=wb.ConstantMean(0.1), noise_variance=0.5)
    b = Batch(eng2, X, np.stack([y, rng.normal(size=333)]), [m.program()])
    f, g, lml, st = b.eval(b.x0())
    assert np.all(st == 0) and np.all(np.isfinite(g))
    print("large-n ok", f, flush=True)
    b.close(); eng2.close()
if which in ("all", "counts"):
    n = 130
    subj = rng.integers(0, 10, size=n).astype(float)
    t = rng.normal(size=n)
    X = np.stack([subj, t], 1)
    yc = rng.poisson(np.exp(0.5 + np.sin(2 * t))).astype(float)
    k = wb.Sum([wb.Categorical(active_dims=[0]), wb.SquaredExponential(active_dims=[1])])
    for lik in ("poisson", "negative_binomial", "bernoulli", "gamma", "zinb"):
        m = wb.GPR(k, mean_function=wb.ConstantMean(0.2), likelihood=make_likelihood(lik))
        yy = (yc > 1).astype(float) if lik == "bernoulli" else (yc + 0.5 if lik == "gamma" else yc)
        b = Batch(eng, X, np.stack([yy, yy[::-1].copy()]), [m.program()])
        from waveome_b200.model_fitting import likelihood_key
        name, par = likelihood_key(m)
        b.set_likelihood(name, par)
        f, g, lml, st = b.eval(b.x0())
        b.latent()
        assert np.all(np.isfinite(f)) and np.all(np.isfinite(g)), (lik, f, g, st)
        print("counts ok", lik, f, st, flush=True)
        b.close()
eng.close()
print("sanitize_case done")
