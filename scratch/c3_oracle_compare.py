import sys, time, copy, numpy as np
sys.path[:0] = [".", "oracle", "tests"]
import gp_oracle as o
from waveome_b200 import datasets, regularization as R
import waveome_b200 as wb
B = 2000
X, Y = datasets.ihmp_scale(n_outcomes=B)
Xs = X.copy()
for c in ("age", "study_day"):
    Xs[c] = (X[c] - X[c].mean()) / X[c].std()
Ys = (Y - Y.mean(axis=0)) / Y.std(axis=0)
k = R.full_kernel_build(cat_vars=[0, 3, 4], num_vars=[1, 2], unit_idx=0, return_sum=True)
for path, p in k.named_parameters():
    if "variance" in path and p.trainable: p.prior = wb.Horseshoe(1.0)
m = wb.GPR(k, mean_function=wb.ConstantMean())
spec = m.to_spec()
xg = np.load("gpurun_out/c3_fit_x.npy"); fg = np.load("gpurun_out/c3_fit_f.npy")
ne = np.load("gpurun_out/c3_fit_neval.npy"); stg = np.load("gpurun_out/c3_fit_status.npy")
Xn = Xs.to_numpy()
def variances(x):
    sp = copy.deepcopy(spec); o.unpack(sp, x)
    out = []
    for kk in sp["kernel"]["kernels"]:
        if kk["type"] == "product":
            out.append(np.prod([c["params"]["variance"]["value"] for c in kk["kernels"]]))
        else:
            out.append(kk["params"]["variance"]["value"])
    return np.array(out)
for b in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    t0 = time.time()
    r = o.fit(spec, Xn, Ys.iloc[:, b].to_numpy(), maxiter=50000, maxfun=50000)
    vo = variances(r["x"]) if r["x"] is not None else None
    vg = variances(xg[b])
    same = None if vo is None else bool(np.all((vo >= 0.1) == (vg >= 0.1)))
    print(b, "oracle: f=%.6f nit=%d nfev=%d st=%d %s | gpu: f=%.6f nfev=%d st=%d | dx=%.2e same_structure=%s (%.1fs)" % (
        r["f"], r["nit"], r["nfev"], r["status"], r["message"][:12], fg[b], ne[b], stg[b],
        np.max(np.abs(r["x"] - xg[b])) if r["x"] is not None else np.nan, same, time.time() - t0), flush=True)
    print("    keep oracle", None if vo is None else np.where(vo >= 0.1)[0], "gpu", np.where(vg >= 0.1)[0])
