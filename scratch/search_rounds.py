import sys, time, numpy as np
sys.path[:0] = [".", "oracle", "tests"]
from waveome_b200 import datasets, engine as E
from waveome_b200.model_search import GPSearch
_fit = E.Batch.fit
def fit_timed(self, *a, **k):
    c0 = self.counters()
    t0 = time.perf_counter()
    res = _fit(self, *a, **k)
    dt = time.perf_counter() - t0
    c1 = self.counters()
    ne = np.asarray(res["n_eval"])
    srt = np.sort(ne)[::-1]
    # rounds with fewer than 64 active models = the 64th largest n_eval .. max
    k64 = srt[min(63, len(srt) - 1)]
    print("B %5d progs %4d n %d P %d: %.2f s, rounds %d, n_eval median %d p90 %d p99 %d max %d; rounds with <64 active: %d; top5 %s" % (
        self.B, len(self.programs), getattr(self, "n", -1), self.P, dt, c1["rounds"] - c0["rounds"], np.median(ne), np.percentile(ne, 90), np.percentile(ne, 99),
        ne.max(), ne.max() - k64, srt[:5]), flush=True)
    return res
E.Batch.fit = fit_timed
X, Y = datasets.overview_synthetic(n_outcomes=200)
gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
gps.run_search(max_depth=2)
print("---- timed")
t0 = time.time()
gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
gps.run_search(max_depth=5)
print("search %.1f s" % (time.time() - t0))
