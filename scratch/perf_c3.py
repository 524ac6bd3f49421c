import sys, time, copy, numpy as np, torch
sys.path[:0] = [".", "oracle", "tests"]
from waveome_b200 import datasets, regularization as R
import waveome_b200 as wb
from waveome_b200.engine import Engine, Batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
do_fit = len(sys.argv) > 2 and sys.argv[2] == "fit"
X, Y = datasets.ihmp_scale(n_outcomes=B)
Xs = X.copy()
for c in ("age", "study_day"):
    Xs[c] = (X[c] - X[c].mean()) / X[c].std()
Ys = (Y - Y.mean(0)) / Y.std(0)
k = R.full_kernel_build(cat_vars=[0, 3, 4], num_vars=[1, 2], unit_idx=0, return_sum=True)
for path, p in k.named_parameters():
    if "variance" in path and p.trainable: p.prior = wb.Horseshoe(1.0)
m = wb.GPR(k, mean_function=wb.ConstantMean())
eng = Engine(0)
t0 = time.time()
bt = Batch(eng, Xs.to_numpy(), Ys.to_numpy().T.copy(), [m.program()], specialize=True)
print("batch create %.2fs workspace %.2f GB" % (time.time() - t0, bt.workspace_bytes / 1e9), flush=True)
import os
if os.environ.get("SOLO"): bt.set_solo(True)
x = bt.x0()
bt.profile(True)
st = torch.cuda.ExternalStream(eng.stream)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    with torch.cuda.stream(st):
        e0.record(st)
        f, g, lml, s = bt.eval(x)
        e1.record(st)
    torch.cuda.synchronize()
    print("eval %d: wall %.2f ms, device %.2f ms; status!=0: %d; f[0]=%.6f" % (it, (time.time() - t0) * 1e3, e0.elapsed_time(e1), int((s != 0).sum()), f[0]), flush=True)
print("per-class ms/eval:", {k_: round(v[0] / 3, 2) for k_, v in bt.profile_read().items() if v[0] > 0})
bt.profile(False)
n = 600
print("per-eval flops (n^3) GF: %.1f -> %.2f TFLOP/s" % (B * n**3 / 1e9, B * n**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12))
if do_fit:
    t0 = time.time()
    res = bt.fit(maxiter=50000, maxfun=50000)
    dt = time.time() - t0
    c = bt.counters()
    print("fit: %.2fs  fits/s %.1f  evals %d  evals/s %.0f rounds %d launches %d" % (dt, B / dt, res["n_eval"].sum(), res["n_eval"].sum() / dt, c["rounds"], c["launches"]))
    print("status hist", np.unique(res["status"], return_counts=True), "n_eval median/max", np.median(res["n_eval"]), res["n_eval"].max(), "nit median", np.median(res["n_iter"]))
    np.save("gpurun_out/c3_fit_x.npy", res["x"][:64]); np.save("gpurun_out/c3_fit_f.npy", res["f"][:64])
    np.save("gpurun_out/c3_fit_neval.npy", res["n_eval"]); np.save("gpurun_out/c3_fit_status.npy", res["status"])
