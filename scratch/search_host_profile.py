"""Host-side cost of the config-2 kernel search without a GPU: the engine fitter is replaced by a fake that does the
same host work per request (candidate_model, program build) and returns a deterministic pseudo-BIC."""
import sys, time, cProfile, pstats, io, zlib
sys.path[:0] = [".", "oracle", "tests"]
import numpy as np
import torch  # outside the timed region (model_search imports it lazily for the rank lookup)
from waveome_b200 import datasets, kernel_search as ks
from waveome_b200.model_search import GPSearch

def fake_fit(requests):
    out = []
    for y, name, kernel in requests:
        m = ks.candidate_model(kernel)
        m.program().signature()
        h = zlib.crc32((name + str(float(y[0]))).encode()) % 10000 / 50.0
        bic = 300.0 - 12.0 * min(name.count("+") + name.count("*"), 2) + h
        out.append((m, round(bic, 2)))
    return out

X, Y = datasets.overview_synthetic(n_outcomes=int(sys.argv[1]) if len(sys.argv) > 1 else 200)
gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
pr = cProfile.Profile() if "--prof" in sys.argv else None
t0 = time.time()
if pr: pr.enable()
gps.run_search(max_depth=5, fit=fake_fit)
if pr: pr.disable()
print("host-only search: %.2f s, %d fits in %d batches" % (time.time() - t0, gps.fit_report["n_fits"], gps.fit_report["batches"]))
if pr:
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:4000])
