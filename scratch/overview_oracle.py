import sys, copy, numpy as np
sys.path[:0] = [".", "oracle", "tests"]
import gp_oracle as o
import waveome_b200 as wb
from waveome_b200 import datasets
from waveome_b200.model_search import GPSearch
from waveome_b200.regularization import full_kernel_build
X, Y = datasets.overview_notebook()
gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
k, names = full_kernel_build(cat_vars=gps.cat_idx, num_vars=gps.cont_idx, unit_idx=gps.unit_idx, var_names=gps.feat_names, return_sum=True,
                      kerns=[wb.SquaredExponential(), wb.Lin()])
print(names)
m = wb.models.PenalizedGPR(k, mean_function=wb.ConstantMean(), penalization_factor=1.0)
spec = m.to_spec()
for out in ("outcome1", "outcome2", "outcome3"):
    r = o.fit(spec, gps.X.to_numpy(), gps.Y[out].to_numpy(), maxiter=50000, maxfun=50000)
    sp = r["model"]
    vs = []
    for kk in sp["kernel"]["kernels"]:
        if kk["type"] == "product": vs.append(np.prod([c["params"]["variance"]["value"] for c in kk["kernels"]]))
        else: vs.append(kk["params"]["variance"]["value"])
    print(out, r["message"][:20], r["nit"], r["nfev"], "f=%.5f" % r["f"], "noise=%.5f" % sp["likelihood_variance"]["value"], "vars", np.round(vs, 4))
    print("   x", np.round(r["x"], 3))
