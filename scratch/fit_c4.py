"""Config 4: one n = 8192 GPR (SE x Categorical + Periodic), full L-BFGS-B fit."""
import sys, time, numpy as np
sys.path[:0] = [".", "oracle", "tests"]
import waveome_b200 as wb
from waveome_b200 import datasets
from waveome_b200.model_fitting import fit_models
X, Y = datasets.large_gpr(512, 16)
Xn = X.to_numpy().copy(); Xn[:, 1] = (Xn[:, 1] - Xn[:, 1].mean()) / Xn[:, 1].std()
y = Y.to_numpy()[:, 0]
def model():
    cat = wb.Categorical(active_dims=[0]); wb.set_trainable(cat.variance, False)
    k = wb.Sum([wb.Product([cat, wb.SquaredExponential(active_dims=[1], lengthscales=0.5)]),
                wb.Periodic(wb.SquaredExponential(active_dims=[1]), period=0.9)])
    return wb.GPR(k, mean_function=wb.ConstantMean(), noise_variance=0.1)
for B in (1, 8):
    ms = [model() for _ in range(B)]
    rng = np.random.default_rng(0)
    Ys = np.stack([y + 0.02 * b * rng.normal(size=len(y)) for b in range(B)])
    t0 = time.time()
    r = fit_models(Xn, Ys, ms)
    dt = time.time() - t0
    print("config 4, B=%d: fit %.2f s, n_iter %s n_eval %s status %s -> %.1f evals/s, lml[0] %.3f, params %s" % (
        B, dt, r["n_iter"].tolist(), r["n_eval"].tolist(), r["status"].tolist(), r["n_eval"].sum() / dt, r["lml"][0],
        np.round([float(p) for p in ms[0].trainable_parameters], 4).tolist()), flush=True)
