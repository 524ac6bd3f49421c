"""Does fitting the 2000 models of config 3 as k concurrent sub-batches (own engine/stream + host thread each) hide the
L-BFGS tail (rounds with few active models are latency bound)?"""
import sys, time, threading, numpy as np, torch
sys.path[:0] = [".", "oracle", "tests"]
from waveome_b200 import datasets, regularization as R
import waveome_b200 as wb
from waveome_b200.engine import Engine, Batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
X, Y = datasets.ihmp_scale(n_outcomes=B)
Xs = X.copy()
for c in ("age", "study_day"):
    Xs[c] = (X[c] - X[c].mean()) / X[c].std()
Ys = ((Y - Y.mean(0)) / Y.std(0)).to_numpy().T.copy()
k = R.full_kernel_build(cat_vars=[0, 3, 4], num_vars=[1, 2], unit_idx=0, return_sum=True)
m = wb.models.PenalizedGPR(k, mean_function=wb.ConstantMean(), penalization_factor=1.0)
ref = {}
def run(nparts, reps=2, interleave=False):
    engs = [Engine(0) for _ in range(nparts)]
    if interleave:
        idx = [np.arange(i, B, nparts) for i in range(nparts)]
    else:
        idx = [np.arange(i * B // nparts, (i + 1) * B // nparts) for i in range(nparts)]
    bts = [Batch(engs[i], Xs.to_numpy(), np.ascontiguousarray(Ys[idx[i]]), [m.program()]) for i in range(nparts)]
    xs = [b.x0() for b in bts]
    out = [None] * nparts
    def work(i):
        out[i] = bts[i].fit(xs[i], maxiter=50000, maxfun=50000)
    for rep in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.time()
        th = [threading.Thread(target=work, args=(i,)) for i in range(nparts)]
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize()
        dt = time.time() - t0
        if rep:
            print("parts %d interleave %d: %.3f s per fit of %d models = %.1f fits/s" % (nparts, interleave, dt, B, B / dt), flush=True)
    f = np.empty(B)
    for i in range(nparts):
        f[idx[i]] = out[i]["f"]
    if not ref:
        ref["f"] = f
    else:
        print("   max |f - f(1 part)| =", np.nanmax(np.abs(f - ref["f"])), "bit-identical:", np.array_equal(f, ref["f"], equal_nan=True), flush=True)
    for b in bts: b.close()
run(1); run(2); run(4); run(4, interleave=True); run(8)
