import sys, time, cProfile, pstats, io
sys.path[:0] = [".", "oracle", "tests"]
from waveome_b200 import datasets, engine as E
from waveome_b200.model_search import GPSearch
acc = {}
_fit = E.Batch.fit
def fit_timed(self, *a, **k):
    t0 = time.perf_counter()
    try:
        return _fit(self, *a, **k)
    finally:
        acc["fit"] = acc.get("fit", 0.0) + time.perf_counter() - t0
        acc.setdefault("B", []).append((self.B, len(self.programs), round(time.perf_counter() - t0, 2)))
E.Batch.fit = fit_timed
X, Y = datasets.overview_synthetic(n_outcomes=200)
gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
gps.run_search(max_depth=2)     # warm-up (buffers, attributes)
acc.clear()
pr = cProfile.Profile()
t0 = time.time(); pr.enable()
gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
gps.run_search(max_depth=5)
pr.disable(); dt = time.time() - t0
print("search %.1f s (under cProfile), device fits %.1f s, batches (B, programs, s): %s" % (dt, acc["fit"], acc["B"]))
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25); print(s.getvalue()[:5500])
