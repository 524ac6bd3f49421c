"""Where the e2e step (GPSearch from pandas -> fitted, pruned models with importances) spends its wall-clock."""
import sys, time
sys.path[:0] = [".", "oracle", "tests"]
import bench
from waveome_b200 import model_search as ms, models as M, postfit
acc = {}
def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
    return w
ms.fit_replicated = timed("fit_replicated (device fit || model construction)", ms.fit_replicated)
import cProfile, pstats, io
_fib = ms.feature_importances_batch
def _fib_prof(*a, **k):
    pr = cProfile.Profile(); pr.enable()
    try:
        return _fib(*a, **k)
    finally:
        pr.disable()
        st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(22); acc["_prof"] = st.getvalue()
ms.feature_importances_batch = timed("feature_importances_batch", _fib_prof)
M.PenalizedGPR.cut_kernel_components = timed("cut_kernel_components (2000 x host)", M.PenalizedGPR.cut_kernel_components)
M.PenalizedGPR.update_kernel_name = timed("update_kernel_name", M.PenalizedGPR.update_kernel_name)
postfit.fitted_means = timed("  fitted_means (inside importances)", postfit.fitted_means)
from waveome_b200 import engine as E
E.Batch.eval = timed("    Batch.eval (inside fitted_means)", E.Batch.eval)
E.Batch.alpha = timed("    Batch.alpha", E.Batch.alpha)
_binit = E.Batch.__init__
def _binit_logged(self, engine, X, Y, *a, **k):
    t0 = time.perf_counter()
    _binit(self, engine, X, Y, *a, **k)
    print("      Batch(B=%d, programs=%d): %.3f s" % (self.B, len(self.programs), time.perf_counter() - t0), flush=True)
E.Batch.__init__ = _binit_logged
_lib = E.load_library()
_create, _destroy = _lib.wv_batch_create, _lib.wv_batch_destroy
_lib.wv_batch_create = timed("    wv_batch_create (C, all)", _create)
_lib.wv_batch_destroy = timed("    wv_batch_destroy (C, all)", _destroy)
M.GPR.program = timed("    GPR.program (all)", M.GPR.program)
postfit.calc_deviance_loglik = timed("    calc_deviance_loglik", postfit.calc_deviance_loglik)
X, Y = bench.make_workload(2000, seed=2024)
for rep in range(3):
    acc.clear()
    t0 = time.perf_counter()
    g = bench.make_search(X, Y)
    t1 = time.perf_counter()
    g.penalized_optimization(penalization_factor=1.0, gather=False)
    t2 = time.perf_counter()
    print("rep %d: GPSearch() %.3f s, penalized_optimization %.3f s, total %.3f s" % (rep, t1 - t0, t2 - t1, t2 - t0))
    prof = acc.pop("_prof", "")
    for k, v in acc.items():
        print("    %-55s %.3f s" % (k, v))
    if rep == 2:
        print(prof[:4500])
    print("    unaccounted inside penalized_optimization: %.3f s" % (t2 - t1 - sum(v for k, v in acc.items() if not k.startswith("  "))), flush=True)
