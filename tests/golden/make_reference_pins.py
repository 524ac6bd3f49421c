"""Extracts the inputs of the one recorded GPflow value whose data is a file of the reference rather than a seeded
generator: kernel_learning/data_generation.ipynb cell 14 evaluates, on kernel_learning/simulated_data.csv (columns unit,
treatment, time, y1_obs), scipy's multivariate_normal.logpdf under cov = (Matern12[time] + Categorical(variance=2)[unit])(X)
+ 1e-2 I and prints k*log(n) - 2*loglik = 169.71614261744028 with k = 3, n = 126.  Run in the build container
(needs /root/reference); writes tests/golden/ref_simulated_y1.json."""
import json
import os

import pandas as pd

df = pd.read_csv("/root/reference/kernel_learning/simulated_data.csv")
out = {"source": "kernel_learning/simulated_data.csv + data_generation.ipynb cell 14 output",
       "X": df[["unit", "treatment", "time"]].to_numpy().tolist(), "y1_obs": df["y1_obs"].tolist(),
       "recorded_k_log_n_minus_2_loglik": 169.71614261744028, "k": 3, "noise_variance": 1e-2}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_simulated_y1.json"), "w") as fh:
    json.dump(out, fh)
