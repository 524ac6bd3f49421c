import sys, copy, time, numpy as np
sys.path[:0] = [".", "oracle", "tests"]
import gp_oracle as o, helpers as H, waveome_b200 as wb
from waveome_b200.engine import Engine, Batch
eng = Engine(0)
for n in (40, 63, 64, 150, 200):
    X, y = H.make_data(n, seed=n)
    for kname, kern in (("all", H.all_leaf_kernel()), ("sat", H.saturated_kernel())):
        m = wb.GPR(kern, mean_function=wb.ConstantMean(0.1), noise_variance=0.5)
        prog = m.program()
        B = 3
        rng = np.random.default_rng(1)
        Y = np.stack([y, y * 0.5 + 0.1, rng.normal(size=n)])
        bt = Batch(eng, X, Y, [prog])
        x = bt.x0() + 0.3 * rng.normal(size=(B, bt.P))
        f, g, lml, st = bt.eval(x)
        for b in range(B):
            fo, go, lo, _ = o.objective(copy.deepcopy(m.to_spec()), X, Y[b], x[b])
            print(n, kname, b, "status", st[b], "f rel", abs(f[b]-fo)/abs(fo), "lml rel", abs(lml[b]-lo)/abs(lo), "g rel", np.max(np.abs(g[b]-go))/np.max(np.abs(go)))
        bt.close()
