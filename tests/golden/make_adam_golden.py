"""Golden fixture for the reference's DEFAULT optimiser on its LIVE objective: Adam + natural gradient on the whitened
SVGP bound with Z = X (objective B; oracle/svgp_oracle.py restates BaseGP.optimize_params "adam/gradient",
waveome/model_classes.py:344-462, and the bound of PSVGP, :1082-1126) -- the fits ``kernel_test`` runs for the candidates
of a kernel search (waveome/model_search.py:2239-2334), on

* the data of examples/simulations/penalized_regression.ipynb cell 1 (np.random.seed(1), N = 100; the notebook's search
  selects categorical[4]+squared_exponential[0], cell 4), and
* the three outcomes of waveome_overview.ipynb cell 4 (datasets.overview_notebook, n = 500; the notebook text gives
  SE[time], female x SE[time], unit effect + linear time),

each with the documented structure and a few competitors.

    python tests/golden/make_adam_golden.py      ->  tests/golden/adam_natgrad_fits.json

tests/test_adam_gpu.py fits the same candidates on the engine (L-BFGS-B and Adam on the collapsed objective A) and
compares the BIC ranking: SURVEY 0.3 / VERDICT r01 item 7 -- (B)'s selected kernel structure is reproduced by (A).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _kernel(name):
    """'categorical[0]+lin[1]' / 'categorical[2]*squared_exponential[1]' -> kernel tree (fresh default parameters)"""
    import waveome_b200 as wb
    leaf = {"categorical": wb.Categorical, "squared_exponential": wb.SquaredExponential, "lin": wb.Lin}

    def one(tok):
        cls, dim = tok[:-1].split("[")
        return leaf[cls](active_dims=[int(dim)])
    comps = []
    for comp in name.split("+"):
        fs = [one(t) for t in comp.split("*")]
        comps.append(fs[0] if len(fs) == 1 else wb.Product(fs))
    return comps[0] if len(comps) == 1 else wb.Sum(comps)


CANDIDATES = {
    # examples/simulations/penalized_regression.ipynb cell 4: the search selects the first one
    "penalized_regression": ["categorical[4]+squared_exponential[0]", "squared_exponential[0]", "categorical[4]",
                             "categorical[4]+squared_exponential[1]", "categorical[4]+lin[0]"],
    # waveome_overview.ipynb text: SE[time]; female x SE[time]; unit effect + linear time
    "overview_outcome1": ["squared_exponential[1]", "lin[1]", "categorical[2]*squared_exponential[1]", "categorical[0]+lin[1]"],
    "overview_outcome2": ["categorical[2]*squared_exponential[1]", "squared_exponential[1]", "categorical[2]*lin[1]",
                          "categorical[0]+squared_exponential[1]"],
    "overview_outcome3": ["categorical[0]+lin[1]", "categorical[0]", "lin[1]", "squared_exponential[1]"],
}


def datasets_():
    """{case: (X [n, D], y [n])}"""
    import numpy as np
    import pandas as pd
    from waveome_b200 import datasets
    from waveome_b200.model_search import GPSearch
    out = {}
    np.random.seed(1)
    N, nc, nid = 100, 5, 50
    X = np.random.uniform(low=-5, high=5, size=(N, nc - 1)).reshape(-1, nc - 1)
    X = np.hstack((X, np.random.choice(a=np.arange(nid), size=(N, 1)).astype(int)))
    Y = (np.sin(X[:, 0]) + ((2 / nid) * X[:, nc - 1] - 1) + np.random.uniform(low=-1, high=1, size=N)).reshape(-1, 1)
    gps = GPSearch(X=pd.DataFrame(X, columns=["X" + str(i) for i in range(nc)]), Y=pd.DataFrame(Y, columns=["Y"]),
                   unit_col="X4", categorical_vars=["X4"])
    out["penalized_regression"] = (gps.X.to_numpy(dtype=float), gps.Y.to_numpy(dtype=float)[:, 0])
    Xo, Yo = datasets.overview_notebook()
    gps = GPSearch(Xo, Yo, unit_col="person_id", categorical_vars=["female"])
    for name in gps.out_names:
        out["overview_" + name] = (gps.X.to_numpy(dtype=float), gps.Y[name].to_numpy(dtype=float))
    return out


def cases():
    """[(case, candidate name, X, y, model)]: the candidate fits a search would run through kernel_test
    (waveome/model_search.py:2239-2334: PSVGP with penalization_factor 0, constant mean) -- shared with the GPU test"""
    import waveome_b200 as wb
    data = datasets_()
    out = []
    for case, names in CANDIDATES.items():
        X, y = data[case]
        for nm in names:
            out.append((case, nm, X, y, wb.GPR(_kernel(nm), mean_function=wb.ConstantMean())))
    return out


def _one(i):
    import torch
    torch.set_num_threads(1)
    import svgp_oracle as so
    case, nm, X, y, model = cases()[i]
    t0 = time.perf_counter()
    r = so.fit_adam_natgrad(model.to_spec(), X, y)
    k = len(model.trainable_parameters)
    # kernel_test: bic = round(calc_bic(log_posterior_density, n, k), 2) with waveome's calc_bic = 2 k - 2 loglik
    # (utilities.py:77-95); the variational parameters add the same two Parameter objects to every candidate's k
    return dict(case=case, candidate=nm, n=int(len(y)), x=[float(v) for v in r["x"]], loss=float(r["loss"]),
                n_iter=int(r["n_iter"]), why=r["why"], k=k, bic=round(2 * (k + 2) + 2 * float(r["loss"]), 2),
                seconds=time.perf_counter() - t0)


def main():
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    n = len(cases())
    with ProcessPoolExecutor(max_workers=min(n, os.cpu_count() or 1), mp_context=mp.get_context("spawn")) as ex:
        res = list(ex.map(_one, range(n)))
    with open(os.path.join(ROOT, "tests", "golden", "adam_natgrad_fits.json"), "w") as fh:
        json.dump({"optimizer": "Adam(0.1, decay 0.96/500) + NaturalGradient(gamma=0.1) on the whitened SVGP bound, Z = X "
                                "(oracle/svgp_oracle.fit_adam_natgrad)", "fits": res}, fh, indent=0)
    for r in res:
        print(r["case"], r["candidate"], r["n_iter"], r["why"], r["loss"], r["bic"], "%.0fs" % r["seconds"])


if __name__ == "__main__":
    main()
