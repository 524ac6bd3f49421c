"""Generates tests/golden/eval_cases.json and fit_cases.json from the oracle (run from the repo root):

    python tests/golden/make_golden.py

The reference ships no golden vectors and cannot be imported here, so these pin the *oracle*:
fixed inputs -> (f, lml, grad) and L-BFGS-B optima.  The GPU tests compare the CUDA path against them."""
import copy
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import gp_oracle as oracle  # noqa: E402
import helpers  # noqa: E402
import waveome_b200 as wb  # noqa: E402


def main():
    out = []
    for n, kern, seed in [(24, "all", 1), (37, "sat", 2), (64, "sat", 3), (90, "all", 4), (130, "sat", 5)]:
        X, y = helpers.make_data(n, seed=seed)
        k = helpers.all_leaf_kernel() if kern == "all" else helpers.saturated_kernel()
        m = wb.GPR(k, mean_function=wb.ConstantMean(0.1), noise_variance=0.6)
        spec = m.to_spec()
        x = oracle.pack(spec) + 0.2 * np.random.default_rng(seed).normal(size=len(oracle.pack(spec)))
        f, g, lml, lp = oracle.objective(copy.deepcopy(spec), X, y, x)
        out.append(dict(name=f"{kern}_n{n}", spec=spec, X=X.tolist(), y=y.tolist(), x=x.tolist(), f=f, lml=lml,
                        log_prior=lp, grad=g.tolist()))
    with open(os.path.join(os.path.dirname(__file__), "eval_cases.json"), "w") as fh:
        json.dump(out, fh)
    fits = []
    for n, seed in [(150, 1), (150, 2), (100, 11)]:
        X, y = helpers.make_data(n, seed=seed)
        m = wb.GPR(helpers.saturated_kernel(), mean_function=wb.ConstantMean(0.0))
        spec = m.to_spec()
        r = oracle.fit(spec, X, y)
        fits.append(dict(name=f"sat_n{n}_s{seed}", n=n, seed=seed, x=r["x"].tolist(), f=r["f"], lml=r["lml"],
                         nit=r["nit"], nfev=r["nfev"], status=r["status"], message=r["message"]))
        print(fits[-1]["name"], r["message"], r["nit"], r["nfev"], r["f"])
    with open(os.path.join(os.path.dirname(__file__), "fit_cases.json"), "w") as fh:
        json.dump(fits, fh)


if __name__ == "__main__":
    main()
