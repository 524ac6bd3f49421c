import sys, time, numpy as np
sys.path[:0] = [".", "oracle", "tests"]
from waveome_b200 import datasets, kernel_search as ks
from waveome_b200.model_search import GPSearch
X, Y = datasets.overview_synthetic(n_outcomes=200)
gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
gps.run_search(max_depth=2)
ref = None
for lanes, tail in [(1, 0), (1, 32), (2, 32), (3, 32), (2, 64), (4, 32)]:
    ks.SEARCH_TAIL, ks.SEARCH_LANES = tail, lanes
    for rep in range(2):
        gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
        t0 = time.time()
        gps.run_search(max_depth=5)
        dt = time.time() - t0
        r = gps.fit_report
        best = {o: gps.search_info[o]["best_model"] for o in gps.out_names}
        if ref is None: ref = best
        print("lanes %d tail %3d: %.2f s, %d fits in %d batches; same structures: %s" % (lanes, tail, dt, r["n_fits"], r["batches"], best == ref), flush=True)
