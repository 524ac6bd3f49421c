import sys; sys.path[:0]=['.','oracle','tests']
import numpy as np
import helpers, waveome_b200 as wb
from waveome_b200.engine import Batch, Engine
import svgp_oracle as so
eng=Engine(0)
n=70
X,y=helpers.make_data(n,seed=12)
rng=np.random.default_rng(1)
Y=np.stack([y, np.sin(2*X[:,1])+0.2*rng.normal(size=n)])
model=wb.GPR(helpers.saturated_kernel(hs=0.0), mean_function=wb.ConstantMean(0.0))
for mi in (50, 150, 250, 450, 1050, 1150, 1250, 3000):
    b=Batch(eng,X,Y,[model.program()]); r=b.fit_adam(max_iter=mi); b.close()
    ref=so.fit_adam_collapsed(model.to_spec(),X,Y[0],max_iter=mi)
    print(mi, r["n_iter"], r["status"], r["f"], "| oracle", ref["n_iter"], ref["why"], ref["f"], np.abs(r["x"][0]-ref["x"]).max(), flush=True)
