import sys, time, cProfile, pstats, io
sys.path[:0] = [".", "oracle", "tests"]
import bench
X, Y = bench.make_workload(2000, seed=2024)
g = bench.make_search(X, Y); g.penalized_optimization(penalization_factor=1.0, gather=False)   # warm-up
pr = cProfile.Profile()
t0 = time.time()
pr.enable()
g = bench.make_search(X, Y)
g.penalized_optimization(penalization_factor=1.0, gather=False)
pr.disable()
print("e2e step %.2f s; fit report %s" % (time.time() - t0, {k: v for k, v in g.fit_report.items() if k != "status"}))
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
