"""Config-4 style probe: single (or few) large GPR models, SE[t] x Cat[subject] + Periodic[t].
usage: python scratch/perf_large.py n_subjects n_times B [check]"""
import sys, time, copy, numpy as np, torch
sys.path[:0] = [".", "oracle", "tests"]
import waveome_b200 as wb
from waveome_b200 import datasets
from waveome_b200.engine import Engine, Batch

ns, nt_, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
check = len(sys.argv) > 4 and sys.argv[4] == "check"
X, Y = datasets.large_gpr(ns, nt_)
Xn = X.to_numpy().copy()
Xn[:, 1] = (Xn[:, 1] - Xn[:, 1].mean()) / Xn[:, 1].std()
y = Y.to_numpy()[:, 0]
n = len(y)
cat = wb.Categorical(active_dims=[0]); wb.set_trainable(cat.variance, False)
k = wb.Sum([wb.Product([cat, wb.SquaredExponential(active_dims=[1], lengthscales=0.5)]),
            wb.Periodic(wb.SquaredExponential(active_dims=[1]), period=0.9)])
m = wb.GPR(k, mean_function=wb.ConstantMean(), noise_variance=0.1)
eng = Engine(0)
rng = np.random.default_rng(0)
Ys = np.stack([y + 0.01 * b * rng.normal(size=n) for b in range(B)])
t0 = time.time()
bt = Batch(eng, Xn, Ys, [m.program()])
print("n=%d B=%d batch create %.2fs workspace %.2f GB" % (n, B, time.time() - t0, bt.workspace_bytes / 1e9), flush=True)
x = bt.x0()
st = torch.cuda.ExternalStream(eng.stream)
bt.profile(True)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        f, g, lml, s = bt.eval(x)
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("eval %d: device %.2f ms; status %s; f[0]=%.9f lml[0]=%.9f" % (it, ms, s.tolist()[:4], f[0], lml[0]), flush=True)
prof = bt.profile_read()
print("per-class ms over 3 evals:", {k_: round(v[0], 2) for k_, v in prof.items()})
chol_ms = (prof["chol_diag"][0] + prof["chol_panel"][0] + prof["chol_syrk"][0]) / 3
print("n^3 flops/eval %.1f GF -> %.2f TFLOP/s; cholesky n^3/3: %.2f ms -> %.2f TFLOP/s" % (
    B * n**3 / 1e9, B * n**3 / (ms * 1e-3) / 1e12, chol_ms, B * n**3 / 3 / (chol_ms * 1e-3) / 1e12))
if check:
    import gp_oracle as oracle
    t0 = time.time()
    fo, go, lo, _ = oracle.objective(copy.deepcopy(m.to_spec()), Xn, Ys[0], x[0])
    print("oracle %.1fs: lml rel err %.3e  grad rel err %.3e" % (time.time() - t0, abs(lml[0] - lo) / abs(lo),
                                                              np.max(np.abs(g[0] - go)) / np.max(np.abs(go))))
