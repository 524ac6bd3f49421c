import os, sys, numpy as np
sys.path[:0] = [".", "oracle", "tests"]
import helpers
import waveome_b200 as wb
from waveome_b200.engine import Engine, Batch
n, B = int(sys.argv[1]), int(sys.argv[2])
X, y = helpers.make_data(n, seed=5)
rng = np.random.default_rng(12)
Y = y[None, :] + 0.3 * rng.normal(size=(B, n))
m = wb.GPR(helpers.saturated_kernel(hs=0.0), mean_function=wb.ConstantMean(0.0))
out = {}
for mode in ("0", "1", "1b"):
    os.environ["WV_CHOL_ALL"] = mode[0]
    eng = Engine(0)
    bt = Batch(eng, X, Y, [m.program()])
    x = bt.x0()
    f, g, lml, s = bt.eval(x)
    f2, g2, _, _ = bt.eval(x)
    out[mode] = (f.copy(), g.copy())
    print(mode, "repeatable:", np.array_equal(f, f2) and np.array_equal(g, g2), "f[0] %.17g" % f[0], "fails", int((s != 0).sum()))
    bt.close()
for mode in ("1", "1b"):
    df = np.abs(out[mode][0] - out["0"][0]); dg = np.abs(out[mode][1] - out["0"][1])
    print(mode, "vs 0: f identical:", np.array_equal(out[mode][0], out["0"][0]), "n differing", int((df > 0).sum()), "max |df|", df.max(),
          "g identical:", np.array_equal(out[mode][1], out["0"][1]), "max |dg|", dg.max(), "first differing models", np.nonzero(df > 0)[0][:10])
