import sys, copy, numpy as np
sys.path[:0] = [".", "oracle", "tests"]
import gp_oracle as oracle, helpers, waveome_b200 as wb
from waveome_b200.engine import Batch, Engine
eng = Engine(0)
for seed in (3,):
    X, y = helpers.make_data(150, seed=seed)
    model = wb.GPR(helpers.saturated_kernel(hs=float(sys.argv[1])), mean_function=wb.ConstantMean(0.0))
    Y = np.stack([y, 0.5 * y + 0.2])
    batch = Batch(eng, X, Y, [model.program()])
    res = batch.fit()
    for b in range(2):
        ref = oracle.fit(model.to_spec(), X, Y[b])
        print("seed", seed, "b", b, "gpu f %.6f nit %d nfev %d st %d | oracle f %.6f nit %d nfev %d st %d %s" % (
            res["f"][b], res["n_iter"][b], res["n_eval"][b], res["status"][b], ref["f"], ref["nit"], ref["nfev"], ref["status"], ref["message"][:30]))
        pass
