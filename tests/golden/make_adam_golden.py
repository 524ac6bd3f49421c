"""Golden fixture for the reference's DEFAULT optimiser on its LIVE objective: Adam + natural gradient on the whitened
SVGP bound with Z = X (objective B; oracle/svgp_oracle.py restates BaseGP.optimize_params "adam/gradient",
waveome/model_classes.py:344-462, and the PSVGP bound, :1082-1126) for the penalised saturated kernel on

* the outcome of examples/simulations/penalized_regression.ipynb cell 1 (np.random.seed(1), N = 100; the notebook's own
  search selects categorical[4]+squared_exponential[0], cell 4), and
* the three outcomes of waveome_overview.ipynb cell 4 (datasets.overview_notebook; the notebook text gives
  SE[time], female x SE[time], unit + linear time),

    python tests/golden/make_adam_golden.py      ->  tests/golden/adam_natgrad_fits.json

tests/test_adam_gpu.py fits the same models on the engine (L-BFGS-B and Adam on the collapsed objective A) and compares
the selected structures: SURVEY 0.3 / VERDICT item 7 — (B)'s selected kernel structure is reproduced by (A).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def cases():
    """[(name, X [n, D], y [n], PenalizedGPR template)] -- shared by the generator and the GPU test"""
    import numpy as np
    import pandas as pd
    import waveome_b200 as wb
    from waveome_b200 import datasets
    from waveome_b200.model_search import GPSearch
    from waveome_b200.regularization import full_kernel_build
    out = []
    np.random.seed(1)
    N, nc, nid = 100, 5, 50
    X = np.random.uniform(low=-5, high=5, size=(N, nc - 1)).reshape(-1, nc - 1)
    X = np.hstack((X, np.random.choice(a=np.arange(nid), size=(N, 1)).astype(int)))
    Y = (np.sin(X[:, 0]) + ((2 / nid) * X[:, nc - 1] - 1) + np.random.uniform(low=-1, high=1, size=N)).reshape(-1, 1)
    gps = GPSearch(X=pd.DataFrame(X, columns=["X" + str(i) for i in range(nc)]), Y=pd.DataFrame(Y, columns=["Y"]),
                   unit_col="X4", categorical_vars=["X4"])
    k = full_kernel_build(cat_vars=gps.cat_idx, num_vars=gps.cont_idx, unit_idx=gps.unit_idx, return_sum=True)
    out.append(("penalized_regression", gps.X.to_numpy(dtype=float), gps.Y.to_numpy(dtype=float)[:, 0],
                wb.models.PenalizedGPR(k, mean_function=wb.ConstantMean(), penalization_factor=1.0)))
    Xo, Yo = datasets.overview_notebook()
    gps = GPSearch(Xo, Yo, unit_col="person_id", categorical_vars=["female"])
    for name in gps.out_names:
        k = full_kernel_build(cat_vars=gps.cat_idx, num_vars=gps.cont_idx, unit_idx=gps.unit_idx, return_sum=True,
                              kerns=[wb.SquaredExponential(), wb.Lin()])
        out.append(("overview_" + name, gps.X.to_numpy(dtype=float), gps.Y[name].to_numpy(dtype=float),
                    wb.models.PenalizedGPR(k, mean_function=wb.ConstantMean(), penalization_factor=1.0)))
    return out


def pruned_name(model, x, Xn):
    import waveome_b200 as wb
    m = wb.kernels.deepcopy(model)
    m.program().assign(x)
    m.cut_kernel_components(Xn)
    m.update_kernel_name()
    return m.kernel_name


def _one(i):
    import torch
    torch.set_num_threads(2)
    import svgp_oracle as so
    name, X, y, model = cases()[i]
    t0 = time.perf_counter()
    r = so.fit_adam_natgrad(model.to_spec(), X, y)
    return dict(case=name, n=int(len(y)), x=[float(v) for v in r["x"]], loss=float(r["loss"]), n_iter=int(r["n_iter"]),
                why=r["why"], kernel_name=pruned_name(model, r["x"], X), seconds=time.perf_counter() - t0)


def main():
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    n = len(cases())
    with ProcessPoolExecutor(max_workers=n, mp_context=mp.get_context("spawn")) as ex:
        res = list(ex.map(_one, range(n)))
    with open(os.path.join(ROOT, "tests", "golden", "adam_natgrad_fits.json"), "w") as fh:
        json.dump({"optimizer": "Adam(0.1, decay 0.96/500) + NaturalGradient(gamma=0.1) on the whitened SVGP bound, Z = X "
                                "(oracle/svgp_oracle.fit_adam_natgrad)", "fits": res}, fh, indent=0)
    for r in res:
        print(r["case"], r["n_iter"], r["why"], r["loss"], r["kernel_name"], "%.0fs" % r["seconds"])


if __name__ == "__main__":
    main()
