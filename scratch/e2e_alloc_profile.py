import sys, time
sys.path[:0] = [".", "oracle", "tests"]
import bench
from waveome_b200 import engine as E
_init, _close, _fit, _eval = E.Batch.__init__, E.Batch.close, E.Batch.fit, E.Batch.eval
log = []
def wrap(name, fn):
    def w(self, *a, **k):
        t = time.time(); r = fn(self, *a, **k); log.append((name, getattr(self, "B", -1), time.time() - t)); return r
    return w
E.Batch.__init__ = wrap("init", _init); E.Batch.close = wrap("close", _close); E.Batch.fit = wrap("fit", _fit); E.Batch.eval = wrap("eval", _eval)
X, Y = bench.make_workload(2000, seed=2024)
for it in range(3):
    log.clear()
    t0 = time.time()
    g = bench.make_search(X, Y); g.penalized_optimization(penalization_factor=1.0, gather=False)
    print("step %d: %.2f s" % (it, time.time() - t0))
    for l in log: print("   %-6s B=%-6d %.3f s" % l)
