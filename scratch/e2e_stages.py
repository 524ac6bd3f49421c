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
ms.feature_importances_batch = timed("feature_importances_batch", ms.feature_importances_batch)
M.PenalizedGPR.cut_kernel_components = timed("cut_kernel_components (2000 x host)", M.PenalizedGPR.cut_kernel_components)
M.PenalizedGPR.update_kernel_name = timed("update_kernel_name", M.PenalizedGPR.update_kernel_name)
postfit.fitted_means = timed("  fitted_means (inside importances)", postfit.fitted_means)
X, Y = bench.make_workload(2000, seed=2024)
for rep in range(3):
    acc.clear()
    t0 = time.perf_counter()
    g = bench.make_search(X, Y)
    t1 = time.perf_counter()
    g.penalized_optimization(penalization_factor=1.0, gather=False)
    t2 = time.perf_counter()
    print("rep %d: GPSearch() %.3f s, penalized_optimization %.3f s, total %.3f s" % (rep, t1 - t0, t2 - t1, t2 - t0))
    for k, v in acc.items():
        print("    %-55s %.3f s" % (k, v))
    print("    unaccounted inside penalized_optimization: %.3f s" % (t2 - t1 - sum(v for k, v in acc.items() if not k.startswith("  "))), flush=True)
