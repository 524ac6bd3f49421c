import sys, numpy as np
sys.path[:0] = [".", "oracle", "tests"]
import waveome_b200 as wb
from waveome_b200 import datasets
from waveome_b200.model_search import GPSearch
from waveome_b200.regularization import full_kernel_build
from waveome_b200.model_fitting import fit_models
X, Y = datasets.overview_notebook()
gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
k = full_kernel_build(cat_vars=gps.cat_idx, num_vars=gps.cont_idx, unit_idx=gps.unit_idx, return_sum=True, kerns=[wb.SquaredExponential(), wb.Lin()])
models = [wb.models.PenalizedGPR(wb.deepcopy(k), mean_function=wb.ConstantMean(), penalization_factor=1.0) for _ in range(3)]
r = fit_models(gps.X.to_numpy(), gps.Y.to_numpy().T.copy(), models, maxiter=50000, maxfun=50000)
np.set_printoptions(linewidth=200, precision=3, suppress=True)
for b in range(3):
    print(b, "status", r["status"][b], "nit", r["n_iter"][b], "nfev", r["n_eval"][b], "f", r["f"][b])
    print("   x", r["x"][b])
