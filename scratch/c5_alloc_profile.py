import sys, time, collections
sys.path[:0] = [".", "oracle", "tests"]
from waveome_b200 import datasets, engine as E
from waveome_b200.model_search import GPSearch
fam = sys.argv[2] if len(sys.argv) > 2 else "negative_binomial"
n_out = int(sys.argv[1]) if len(sys.argv) > 1 else 200
_init, _close, _fit, _eval = E.Batch.__init__, E.Batch.close, E.Batch.fit, E.Batch.eval
agg = collections.defaultdict(lambda: [0, 0.0])
def wrap(name, fn):
    def w(self, *a, **k):
        t = time.time(); r = fn(self, *a, **k); agg[name][0] += 1; agg[name][1] += time.time() - t; return r
    return w
E.Batch.__init__ = wrap("init", _init); E.Batch.close = wrap("close", _close); E.Batch.fit = wrap("fit", _fit); E.Batch.eval = wrap("eval", _eval)
X, Y = datasets.count_microbiome(n_outcomes=n_out, family=fam)
for it in range(2):
    agg.clear()
    t0 = time.time()
    gps = GPSearch(X, Y, unit_col="subject", outcome_likelihood=fam)
    gps.penalized_optimization()
    print("step %d: %.2f s" % (it, time.time() - t0), {k: (v[0], round(v[1], 3)) for k, v in agg.items()}, flush=True)
