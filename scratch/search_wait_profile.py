"""How much of the search's wall-clock does the main thread spend waiting for fits (device-bound) vs advancing generators?"""
import sys, time, threading, numpy as np
sys.path[:0] = [".", "oracle", "tests"]
import concurrent.futures as cf
from waveome_b200 import datasets, kernel_search as ks, model_fitting as mf, engine as E
from waveome_b200.model_search import GPSearch
acc = {"wait": 0.0, "fit_c": 0.0, "fit_py": 0.0}
_wait = cf.wait
def wait_timed(*a, **k):
    t0 = time.perf_counter()
    try: return _wait(*a, **k)
    finally: acc["wait"] += time.perf_counter() - t0
cf.wait = wait_timed
lock = threading.Lock()
for name in ("fit_begin", "fit_run", "fit_report", "fit"):
    fn = getattr(E.Batch, name)
    def mk(fn):
        def w(self, *a, **k):
            t0 = time.perf_counter()
            try: return fn(self, *a, **k)
            finally:
                with lock: acc["fit_c"] += time.perf_counter() - t0
        return w
    setattr(E.Batch, name, mk(fn))
_fm = mf.fit_models
def fm_timed(*a, **k):
    t0 = time.perf_counter()
    try: return _fm(*a, **k)
    finally:
        with lock: acc["fit_py"] += time.perf_counter() - t0
mf.fit_models = fm_timed
X, Y = datasets.overview_synthetic(n_outcomes=200)
gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"]); gps.run_search(max_depth=2)
for k in acc: acc[k] = 0.0
t0 = time.time()
gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"]); gps.run_search(max_depth=5)
dt = time.time() - t0
print("search %.2f s; main thread waiting %.2f s; fit_models total (all threads) %.2f s of which inside C calls %.2f s" % (dt, acc["wait"], acc["fit_py"], acc["fit_c"]))
