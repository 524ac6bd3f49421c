"""CPU-only profile of the host side of GPSearch.penalized_optimization (engine calls stubbed)."""
import sys, time, cProfile, pstats, io
sys.path[:0] = [".", "oracle", "tests"]
import numpy as np
import bench
from waveome_b200 import model_search, postfit, model_fitting

def fake_fit_replicated(X, Y, template, make_models=None, **kw):
    models = make_models()
    B = len(models)
    rng = np.random.default_rng(0)
    P = template.program().n_x
    x = np.tile(template.program().x0(), (B, 1)) + 0.5 * rng.normal(size=(B, P))
    for m, xb in zip(models, x):
        m.program().assign(xb)
        m.fit_info = dict(status=0)
    return dict(x=x, f=np.zeros(B), lml=np.zeros(B), n_iter=np.zeros(B, np.int32), n_eval=np.ones(B, np.int32), status=np.zeros(B, np.int32)), models
model_search.fit_replicated = fake_fit_replicated
def fake_fitted_means(X, Y, models, masks=None, engine=None, **kw):
    # keep the host part of the real function: program building and grouping
    progs = [m.program() for m in models]
    sigs = [p.signature() for p in progs]
    return np.random.default_rng(1).normal(size=Y.shape), np.zeros(len(models), np.int32)
postfit.fitted_means = fake_fitted_means
X, Y = bench.make_workload(2000, seed=2024)
g = bench.make_search(X, Y); g.penalized_optimization(penalization_factor=1.0, gather=False)
pr = cProfile.Profile(); t0 = time.time(); pr.enable()
g = bench.make_search(X, Y); g.penalized_optimization(penalization_factor=1.0, gather=False)
pr.disable(); print("host step %.2f s" % (time.time() - t0))
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(30); print(s.getvalue()[:6000])
