"""Golden fixture of BASELINE configs[2] (iHMP-scale synthetic metabolome, n = 600, D = 5, saturated horseshoe-penalised
kernel, P = 17): the workload generator of bench.py at N outcomes, every outcome fitted by the CPU oracle (oracle/gp_oracle.py + SciPy
L-BFGS-B, the restatement of gpflow.optimizers.Scipy().minimize at waveome/model_fitting.py:276-281), then pruned with
``cut_kernel_components`` (waveome/model_classes.py:1029-1079).

    python tests/golden/make_c3_golden.py [N=64] [procs=8]            ->  tests/golden/c3_fits.json
    python tests/golden/make_c3_golden.py 64 8 perturb                ->  tests/golden/c3_fits_perturbed.json

The second file is the ORACLE'S OWN sensitivity: the same fits with every y multiplied by (1 +- 2^-50) (a last-bits change
of the input).  Most fits of this workload run into the horseshoe's singular regime (its log-density grows without bound as
a variance goes to 0), where the stopping point of L-BFGS-B depends on the last bits of every evaluation; the spread between
the two oracle runs is the yardstick for the spread between the engine and the oracle.

tests/test_c3_parity_gpu.py fits the same outcomes on the engine and compares objective, parameters and the pruned
structure; tests/test_c3_golden_cpu.py re-runs a few entries on the oracle so that the fixture cannot drift.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def c3_setup(n_outcomes):
    """(GPSearch, PenalizedGPR) of bench.py's workload generator at ``n_outcomes`` outcomes (same covariates as the
    2000-outcome benchmark; the outcome columns depend on n_outcomes, so fixture and tests both use 64)."""
    import waveome_b200 as wb
    from waveome_b200 import datasets
    from waveome_b200.model_search import GPSearch
    from waveome_b200.regularization import full_kernel_build
    X, Y = datasets.ihmp_scale(n_subjects=120, n_visits=5, n_outcomes=n_outcomes, seed=2024)
    gps = GPSearch(X, Y, unit_col="participant", categorical_vars=["participant", "sex", "site"],
                   Y_transform="standardize")
    k = full_kernel_build(cat_vars=gps.cat_idx, num_vars=gps.cont_idx, unit_idx=gps.unit_idx, return_sum=True)
    model = wb.models.PenalizedGPR(k, mean_function=wb.ConstantMean(), penalization_factor=1.0)
    return gps, model


def pruned_name(model, x, Xn):
    """kernel_name of a deep copy of ``model`` at the packed unconstrained vector ``x`` after cut_kernel_components."""
    import waveome_b200 as wb
    m = wb.kernels.deepcopy(model)
    m.program().assign(x)
    m.cut_kernel_components(Xn)
    m.update_kernel_name()
    return m.kernel_name


def _fit_one(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    from threadpoolctl import threadpool_limits
    import gp_oracle
    spec, Xn, y = args
    with threadpool_limits(1):
        t0 = time.perf_counter()
        r = gp_oracle.fit(spec, Xn, y, maxiter=50000, maxfun=50000)
    return dict(x=[float(v) for v in r["x"]], f=float(r["f"]), lml=float(r["lml"]), nit=r["nit"], nfev=r["nfev"],
                status=r["status"], seconds=time.perf_counter() - t0)


def main():
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    import numpy as np
    n_out = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    perturb = len(sys.argv) > 3 and sys.argv[3] == "perturb"
    gps, model = c3_setup(n_out)
    spec = model.to_spec()
    Xn, Yn = gps.X.to_numpy(dtype=np.float64), gps.Y.to_numpy(dtype=np.float64)
    if perturb:
        sign = np.where(np.arange(Yn.shape[0]) % 2 == 0, 1.0, -1.0)[:, None]
        Yn = Yn * (1.0 + sign * 2.0 ** -50)
    with ProcessPoolExecutor(max_workers=procs, mp_context=mp.get_context("spawn")) as ex:
        fits = list(ex.map(_fit_one, [(spec, Xn, Yn[:, c].copy()) for c in range(n_out)]))
    for c, r in enumerate(fits):
        r["outcome"] = c
        r["kernel_name"] = pruned_name(model, np.array(r["x"]), Xn)
        r["y_checksum"] = float(np.sum(Yn[:, c] * np.arange(1, Yn.shape[0] + 1)))
    out = {"config": "BASELINE configs[2]: datasets.ihmp_scale(120, 5, n_outcomes, seed=2024), standardised, saturated kernel "
                     "(9 components, P = 17), horseshoe pf = 1.0, L-BFGS-B maxiter = maxfun = 50000",
           "n": int(Xn.shape[0]), "n_outcomes": n_out, "fits": fits}
    if perturb:
        out["config"] += "; y perturbed by a factor (1 +- 2^-50)"
        for r in fits:
            del r["y_checksum"]
    with open(os.path.join(ROOT, "tests", "golden", "c3_fits_perturbed.json" if perturb else "c3_fits.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    st = [r["status"] for r in fits]
    print("statuses", {s: st.count(s) for s in set(st)}, "mean nfev", np.mean([r["nfev"] for r in fits]),
          "seconds", sum(r["seconds"] for r in fits))


if __name__ == "__main__":
    main()
