"""Golden fixture of BASELINE configs[1] (waveome_overview synthetic longitudinal data: 50 subjects x 10 time points, n = 500;
kernels [SE, Matern12, Lin, Periodic(SE)], max_depth 5, metric_diff 6, early stopping + pruning, num_restart 1): the
compositional search of the first N outcomes with every candidate fitted by the CPU oracle (oracle/gp_oracle.py + SciPy
L-BFGS-B) through the product's own host search logic (``GPSearch.run_search(fit=oracle_fitter)``; reference:
waveome/model_search.py:1069-1250, 2987-3272).

    python tests/golden/make_c2_search_golden.py [N=8] [procs=8]      ->  tests/golden/c2_search.json

tests/test_c2_search_parity_gpu.py runs the same search on the engine and compares the selected structure per outcome.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

N_TOTAL = 200       # the generator's outcome columns depend on the total count: always generate configs[1]'s 200


def c2_search(outcomes, fit=None, **kw):
    """GPSearch over the listed outcome columns of the 200-outcome configs[1] workload, default search options."""
    from waveome_b200 import datasets
    from waveome_b200.model_search import GPSearch
    X, Y = datasets.overview_synthetic(n_people=50, n_observations=10, n_outcomes=N_TOTAL, seed=9102)
    gps = GPSearch(X, Y[list(outcomes)], unit_col="person_id", categorical_vars=["female"])
    if fit == "oracle":
        from oracle_fitter import oracle_fitter
        import numpy as np
        kw["fit"] = oracle_fitter(gps.X.to_numpy(dtype=np.float64))
    gps.run_search(max_depth=5, random_seed=0, **kw)
    return gps


def _one(col):
    os.environ["OMP_NUM_THREADS"] = "1"
    from threadpoolctl import threadpool_limits
    with threadpool_limits(1):
        t0 = time.perf_counter()
        gps = c2_search([col], fit="oracle")
    info = gps.search_info[col]
    return dict(outcome=col, best_model=info["best_model"], kernel_name=gps.models[col].kernel_name,
                bic={k: float(v["bic"]) for k, v in info["models"].items()},
                x=[float(v) for v in gps.models[col].program().x0()],
                n_fits=gps.fit_report["n_fits"], seconds=time.perf_counter() - t0)


def main():
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    from waveome_b200 import datasets
    n_out = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    _, Y = datasets.overview_synthetic(n_people=50, n_observations=10, n_outcomes=N_TOTAL, seed=9102)
    cols = list(Y.columns[:n_out])
    with ProcessPoolExecutor(max_workers=procs, mp_context=mp.get_context("spawn")) as ex:
        res = list(ex.map(_one, cols))
    with open(os.path.join(ROOT, "tests", "golden", "c2_search.json"), "w") as fh:
        json.dump({"config": "BASELINE configs[1]: datasets.overview_synthetic(50, 10, 200, seed=9102), run_search defaults "
                             "(4 kernels incl. Periodic, max_depth 5, metric_diff 6), first N outcomes",
                   "searches": res}, fh, indent=0)
    for r in res:
        print(r["outcome"], r["best_model"], r["n_fits"], "%.0fs" % r["seconds"])


if __name__ == "__main__":
    main()
