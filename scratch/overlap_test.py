"""Do two half-batches on two streams overlap FP64-ALU (gram/grad) and DMMA (factorisation) kernels?"""
import sys, time, threading, numpy as np, torch
sys.path[:0] = [".", "oracle", "tests"]
from waveome_b200 import datasets, regularization as R
import waveome_b200 as wb
from waveome_b200.engine import Engine, Batch
B = 2000
X, Y = datasets.ihmp_scale(n_outcomes=B)
Xs = X.copy()
for c in ("age", "study_day"):
    Xs[c] = (X[c] - X[c].mean()) / X[c].std()
Ys = ((Y - Y.mean(0)) / Y.std(0)).to_numpy().T.copy()
k = R.full_kernel_build(cat_vars=[0, 3, 4], num_vars=[1, 2], unit_idx=0, return_sum=True)
for path, p in k.named_parameters():
    if "variance" in path and p.trainable: p.prior = wb.Horseshoe(1.0)
m = wb.GPR(k, mean_function=wb.ConstantMean())
def run(nparts, reps=6, stagger=0.0):
    engs = [Engine(0) for _ in range(nparts)]
    bts = [Batch(engs[i], Xs.to_numpy(), Ys[i * B // nparts:(i + 1) * B // nparts], [m.program()]) for i in range(nparts)]
    xs = [b.x0() for b in bts]
    for b, x in zip(bts, xs): b.eval(x)
    torch.cuda.synchronize()
    def work(i):
        if stagger: time.sleep(stagger * i)
        for _ in range(reps): bts[i].eval(xs[i])
    t0 = time.time()
    th = [threading.Thread(target=work, args=(i,)) for i in range(nparts)]
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    dt = time.time() - t0
    print("parts %d stagger %.3f: %.2f ms per full-batch evaluation" % (nparts, stagger, dt / reps * 1e3), flush=True)
    for b in bts: b.close()
run(1); run(2); run(2, stagger=0.008); run(4); run(4, stagger=0.004)
