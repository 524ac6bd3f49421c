import sys, time, numpy as np
sys.path[:0] = [".", "oracle", "tests"]
import waveome_b200 as wb
from waveome_b200 import datasets
from waveome_b200.model_search import GPSearch
n_out = int(sys.argv[1]) if len(sys.argv) > 1 else 200
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 5
X, Y = datasets.overview_synthetic(n_outcomes=n_out)
gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
t0 = time.time()
gps.run_search(max_depth=depth)
dt = time.time() - t0
r = gps.fit_report
print("config 2: %d outcomes, depth %d: %.1f s, %d fits in %d batches -> %.1f fits/s, %.2f outcomes/s" % (n_out, depth, dt, r["n_fits"], r["batches"], r["n_fits"] / dt, n_out / dt))
import collections
print(collections.Counter(gps.search_info[o]["best_model"] for o in gps.out_names).most_common(12))
