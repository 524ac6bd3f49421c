"""Config 5: n = 500 (50 x 10), count outcomes, penalized_optimization with the Poisson / NB variational bound."""
import sys, time, numpy as np
sys.path[:0] = [".", "oracle", "tests"]
from waveome_b200 import datasets
from waveome_b200.model_search import GPSearch
n_out = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
fam = sys.argv[2] if len(sys.argv) > 2 else "poisson"
X, Y = datasets.count_microbiome(n_outcomes=n_out, family=fam)
gps = GPSearch(X, Y, unit_col="subject", outcome_likelihood=fam)
gps.penalized_optimization()          # warm-up (context, module load)
t0 = time.time()
gps = GPSearch(X, Y, unit_col="subject", outcome_likelihood=fam)
gps.penalized_optimization()
dt = time.time() - t0
r = gps.fit_report
import collections
print("config 5 (%s): %d outcomes n=%d: %.2f s -> %.1f fits/s; outer evaluations %d (%.0f/s); status %s" % (
    fam, n_out, len(X), dt, n_out / dt, r["n_eval"], r["n_eval"] / dt, dict(collections.Counter(r["status"].tolist()))))
print(collections.Counter(m.kernel_name for m in gps.models.values()).most_common(5))
