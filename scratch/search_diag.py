import sys, time, copy, numpy as np
sys.path[:0] = [".", "oracle", "tests"]
import waveome_b200 as wb, gp_oracle as oracle
from waveome_b200 import datasets, kernel_search as ks
from waveome_b200.model_search import GPSearch
from waveome_b200.model_fitting import fit_models
from oracle_fitter import oracle_fitter
X, Y = datasets.overview_synthetic(n_people=20, n_observations=6, n_outcomes=4)
kl = lambda: [wb.SquaredExponential(), wb.Matern12(), wb.Lin(), wb.Periodic(wb.SquaredExponential())]
gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
gps.run_search(kernels=kl(), max_depth=3, random_seed=0)
Xn = gps.X.to_numpy(dtype=np.float64)
for o in gps.out_names:
    info = gps.search_info[o]
    name = info["best_model"]
    kern = info["models"][name]["kernel"]
    kern = ks._reset(wb.deepcopy(kern)) if info["models"][name]["parent"] != "None" else wb.deepcopy(kern)
    y = gps.Y[o].to_numpy()
    m = ks.candidate_model(kern)
    r = fit_models(Xn, y[None, :], [m])
    m2 = ks.candidate_model(kern)
    ro = oracle.fit(m2.to_spec(), Xn, y, maxiter=50000, maxfun=50000)
    print(o, name, "bic(search)", info["models"][name]["bic"])
    print("   gpu   : f %.8f nit %d nfev %d status %d x %s" % (r["f"][0], r["n_iter"][0], r["n_eval"][0], r["status"][0], np.round(r["x"][0][:m.program().n_x], 5)))
    print("   oracle: f %.8f nit %d nfev %d status %d x %s %s" % (ro["f"], ro["nit"], ro["nfev"], ro["status"], np.round(ro["x"], 5), ro["message"]))
