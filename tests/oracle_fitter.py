"""Fitter for waveome_b200.kernel_search driven by the CPU oracle (SciPy L-BFGS-B on oracle/gp_oracle.py) instead of
the engine: the same host-side search logic then runs once per backend and the selected structures are compared."""
import numpy as np

import gp_oracle as oracle
from waveome_b200 import kernel_search as ks


def oracle_fitter(X, max_iter=50000):
    def fit(requests):
        out = []
        for y, _name, kernel in requests:
            m = ks.candidate_model(kernel)
            r = oracle.fit(m.to_spec(), X, np.asarray(y), maxiter=max_iter, maxfun=max_iter)
            if r["status"] & oracle.STATUS_CHOL_FAIL or not np.isfinite(r["f"]):
                out.append((None, np.inf))
                continue
            m.program().assign(r["x"])
            m.log_posterior_density_value = -r["f"]
            m.log_marginal_likelihood_value = r["lml"]
            out.append((m, ks.candidate_bic(m, -r["f"])))
        return out
    return fit
