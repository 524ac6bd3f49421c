"""The reference's default optimiser (BaseGP.optimize_params "adam/gradient", waveome/model_classes.py:344-462; what
kernel_test calls, waveome/model_search.py:2284-2297) on the engine: ``wv_batch_fit_adam`` runs the same Adam schedule
on the collapsed objective (A) — the natural-gradient half of the upstream step is the engine's exact inner maximisation.

* against oracle/svgp_oracle.fit_adam_collapsed (NumPy, same schedule, same objective): iteration counts and optimum;
* against the committed fixture of the REAL upstream algorithm — Adam + NaturalGradient(0.1) on the whitened SVGP bound
  (B), tests/golden/adam_natgrad_fits.json — the selected kernel structure of (B) is reproduced by (A), with either
  optimiser, on the notebook outcomes the reference documents."""
import json
import os
import sys

import numpy as np
import pytest

import helpers
import waveome_b200 as wb

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)


def test_adam_matches_the_numpy_restatement(engine):
    import svgp_oracle as so
    from waveome_b200.engine import Batch
    n = 70
    X, y = helpers.make_data(n, seed=12)
    rng = np.random.default_rng(1)
    Y = np.stack([y, np.sin(2 * X[:, 1]) + 0.2 * rng.normal(size=n)])
    model = wb.GPR(helpers.saturated_kernel(hs=0.0), mean_function=wb.ConstantMean(0.0))
    # Identical trajectories while rounding has had no time to act (learning rate 0.1: Adam's iterates on this objective
    # separate exponentially -- 2e-13 after 450 steps, 2e-7 after 1050, O(0.1) after 1150, measured): every prefix of the
    # schedule, incl. the checkpoints at steps 0, 100, ... with their learning-rate decay, must reproduce the NumPy run.
    for steps in (50, 150, 250, 450):
        batch = Batch(engine, X, Y, [model.program()])
        r = batch.fit_adam(max_iter=steps)
        batch.close()
        for b in range(2):
            ref = so.fit_adam_collapsed(model.to_spec(), X, Y[b], max_iter=steps)
            assert int(r["n_iter"][b]) == steps == ref["n_iter"] and (int(r["status"][b]) & 4) and ref["why"] == "maxiter"
            np.testing.assert_allclose(r["x"][b], ref["x"], rtol=1e-5, atol=1e-5)      # 1e-12 at 50 steps, 6e-7 at 450 (measured)
            assert abs(r["f"][b] - ref["f"]) <= 1e-9 * max(1.0, abs(ref["f"]))
    # The stopping rule (loss fell by < 1e-9 between two checkpoints 100 steps apart) fires on the first checkpoint at
    # which Adam's oscillating loss happens to be higher than at the previous one, so WHERE a run stops is a property of
    # its last bits; both runs must stop at a checkpoint (or the step limit) in the flat region of the same optimum.
    batch = Batch(engine, X, Y, [model.program()])
    r = batch.fit_adam(max_iter=3000)
    batch.close()
    for b in range(2):
        ref = so.fit_adam_collapsed(model.to_spec(), X, Y[b], max_iter=3000)
        assert int(r["n_iter"][b]) == 3000 or int(r["n_iter"][b]) % 100 == 1, r["n_iter"][b]
        assert ref["n_iter"] == 3000 or ref["n_iter"] % 100 == 1
        assert abs(r["f"][b] - ref["f"]) <= 5e-2 * max(1.0, abs(ref["f"])), (r["f"][b], ref["f"])


def test_objective_b_structures_reproduced_by_the_collapsed_objective(engine):
    """kernel_test's candidate fits: the BIC ranking of the upstream algorithm on objective (B) (fixture) == the ranking
    of the engine on objective (A) with L-BFGS-B and with Adam; the winner is the structure the reference's notebooks
    report; the BICs themselves differ by the constant 4 (two variational Parameter objects in upstream's k), the 1e-6
    jitter, and what Adam's stopping rule leaves on the table."""
    from make_adam_golden import CANDIDATES, cases
    from waveome_b200.engine import Batch
    with open(os.path.join(GOLDEN, "adam_natgrad_fits.json")) as fh:
        gold = {(g["case"], g["candidate"]): g for g in json.load(fh)["fits"]}
    by_case = {}
    for case, nm, X, y, model in cases():
        batch = Batch(engine, X, y[None, :], [model.program()])
        lb = batch.fit(maxiter=50000, maxfun=50000)
        ad = batch.fit_adam()
        batch.close()
        k = len(model.trainable_parameters)
        row = dict(b=gold[(case, nm)]["bic"] - 4.0, lbfgs=round(2 * k + 2 * float(lb["f"][0]), 2),
                   adam=round(2 * k + 2 * float(ad["f"][0]), 2), adam_iter=int(ad["n_iter"][0]), b_iter=gold[(case, nm)]["n_iter"])
        by_case.setdefault(case, {})[nm] = row
        print(case, nm, row)
    for case, rows in by_case.items():
        documented = CANDIDATES[case][0]
        for key in ("b", "lbfgs", "adam"):
            best = min(rows, key=lambda nm: rows[nm][key])
            assert best == documented, (case, key, best, rows)
        for nm, row in rows.items():
            # (A) is the collapsed form of (B): max_q B = A, so an optimiser of (A) may not end ABOVE upstream's run on (B)
            # by more than a local-optimum's worth (measured: one candidate with an irrelevant SE term, 304.89 vs 304.78);
            # upstream's Adam often stops early on (B) (e.g. 301 steps, BIC -753.65 against the optimum's -768.58)
            assert row["lbfgs"] <= row["b"] + 0.5 and row["adam"] <= row["b"] + 0.5, (case, nm, row)
            assert abs(row["lbfgs"] - row["adam"]) <= 0.5, (case, nm, row)
