"""Fit-level parity on BASELINE configs[2] (the benchmark's workload): 64 outcomes of the iHMP-scale generator, n = 600,
saturated horseshoe-penalised kernel (P = 17), fitted by the engine's device L-BFGS-B and — in the committed fixture
tests/golden/c3_fits.json (tests/golden/make_c3_golden.py) — by the oracle's SciPy L-BFGS-B.

North-star criteria checked per outcome: identical selected kernel structure after ``cut_kernel_components``
(waveome/model_classes.py:1029-1079), objective value, optimised hyper-parameters to 1e-5 where both optimisers converge
along the same trajectory.  The horseshoe log-density has no minimum in an unused variance (it grows like log log 1/v),
so most fits of this workload end in the underflow regime where both sides stop ABNORMAL at a last-bits-dependent point:
those are compared by objective value and structure, and the counts are printed."""
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)


@pytest.fixture(scope="module")
def c3(engine):
    from make_c3_golden import c3_setup, pruned_name
    from waveome_b200.engine import Batch
    with open(os.path.join(GOLDEN, "c3_fits.json")) as fh:
        gold = json.load(fh)
    B = gold["n_outcomes"]
    gps, model = c3_setup(B)
    Xn = gps.X.to_numpy(dtype=np.float64)
    Yn = np.ascontiguousarray(gps.Y.to_numpy(dtype=np.float64).T)
    for r in gold["fits"]:      # the fixture belongs to exactly this data
        assert abs(float(np.sum(Yn[r["outcome"]] * np.arange(1, Yn.shape[1] + 1))) - r["y_checksum"]) < 1e-9
    batch = Batch(engine, Xn, Yn, [model.program()])
    res = batch.fit(maxiter=50000, maxfun=50000)
    batch.close()
    names = [pruned_name(model, res["x"][b], Xn) for b in range(B)]
    return gold["fits"], res, names


def test_c3_selected_structure_identical_for_every_outcome(c3):
    gold, res, names = c3
    diff = [(r["outcome"], r["kernel_name"], names[r["outcome"]]) for r in gold if r["kernel_name"] != names[r["outcome"]]]
    print("structure_identical", len(gold) - len(diff), "of", len(gold))
    assert not diff, diff


def test_c3_objective_matches_for_every_outcome(c3):
    gold, res, _ = c3
    rel = np.array([abs(res["f"][r["outcome"]] - r["f"]) / max(1.0, abs(r["f"])) for r in gold])
    print("objective rel diff: median %.2e max %.2e" % (np.median(rel), rel.max()))
    # every outcome, whatever the termination: the optimum value of the same objective
    assert rel.max() <= 1e-6, sorted(zip(rel, range(len(rel))))[-5:]


def test_c3_parameters_match_where_both_converge(c3):
    gold, res, _ = c3
    both = [r for r in gold if r["status"] == 0 and res["status"][r["outcome"]] == 0]
    same_traj = [r for r in both if r["nit"] == res["n_iter"][r["outcome"]] and r["nfev"] == res["n_eval"][r["outcome"]]]
    agree = sum(1 for r in gold if (r["status"] == 0) == (res["status"][r["outcome"]] == 0))
    print("status_agree", agree, "of", len(gold), "| converged on both", len(both), "| identical nit/nfev", len(same_traj))
    assert len(both) >= 5
    for r in both:
        b = r["outcome"]
        # north star: optimised hyper-parameters to 1e-5 (constrained values; the unconstrained softplus argument of a
        # variance pushed to ~0 is ill-conditioned by construction)
        np.testing.assert_allclose(_constrained(res["x"][b]), _constrained(np.array(r["x"])), rtol=1e-5, atol=1e-5,
                                   err_msg=f"outcome {b}")


def _constrained(x):
    """softplus of the 16 positive parameters (noise: + 1e-6), identity for the mean constant (last entry)."""
    v = np.logaddexp(0.0, x)
    v[-1] = x[-1]
    return v
