"""The committed oracle fixtures of the BASELINE configs cannot drift from the oracle: a sample of their entries is
recomputed here (CPU) with oracle/gp_oracle.py."""
import json
import os
import sys

import numpy as np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import gp_oracle as oracle


def test_c3_fixture_entries_reproduce_on_the_oracle():
    from make_c3_golden import c3_setup, pruned_name
    with open(os.path.join(GOLDEN, "c3_fits.json")) as fh:
        gold = json.load(fh)
    gps, model = c3_setup(gold["n_outcomes"])
    Xn, Yn = gps.X.to_numpy(dtype=np.float64), gps.Y.to_numpy(dtype=np.float64)
    spec = model.to_spec()
    # objective value and pruned structure of every entry at its recorded optimum (one evaluation each) ...
    for r in gold["fits"][:12]:
        f, _, lml, _ = oracle.objective(json.loads(json.dumps(spec)), Xn, Yn[:, r["outcome"]], np.array(r["x"]), want_grad=False)
        assert abs(f - r["f"]) <= 1e-10 * max(1.0, abs(r["f"])) and abs(lml - r["lml"]) <= 1e-10 * max(1.0, abs(r["lml"]))
        assert pruned_name(model, np.array(r["x"]), Xn) == r["kernel_name"]
    # ... and one complete fit (the converged entry with the fewest evaluations)
    r = min((q for q in gold["fits"] if q["status"] == 0), key=lambda q: q["nfev"])
    fit = oracle.fit(spec, Xn, Yn[:, r["outcome"]], maxiter=50000, maxfun=50000)
    assert fit["nit"] == r["nit"] and fit["nfev"] == r["nfev"] and fit["status"] == 0
    np.testing.assert_allclose(fit["x"], np.array(r["x"]), rtol=1e-9, atol=1e-9)
