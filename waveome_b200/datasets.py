"""Synthetic inputs of the BASELINE.json configs.  All float64; no files, no network."""
from __future__ import annotations

import numpy as np
import pandas as pd


def ihmp_scale(n_subjects=120, n_visits=5, n_outcomes=2000, seed=2024, null_frac=0.3):
    """Config 3: iHMP-scale synthetic metabolome.  n = 600 samples x 5 covariates
    (participant[unit], age, study_day, sex, site), ``n_outcomes`` log1p-transformed intensities."""
    rng = np.random.default_rng(seed)
    n = n_subjects * n_visits
    pid = np.repeat(np.arange(n_subjects), n_visits)
    day = np.sort(rng.uniform(0, 365, size=(n_subjects, n_visits)), axis=1).reshape(-1)
    age0 = rng.uniform(6, 76, size=n_subjects)
    age = age0[pid] + day / 365.0
    sex = rng.integers(0, 2, size=n_subjects)[pid]
    site = rng.integers(0, 5, size=n_subjects)[pid]
    X = pd.DataFrame({"participant": pid.astype(float), "age": age, "study_day": day, "sex": sex.astype(float),
                      "site": site.astype(float)})
    az = (age - age.mean()) / age.std()
    dz = (day - day.mean()) / day.std()
    Y = np.empty((n, n_outcomes))
    kinds = rng.integers(0, 5, size=n_outcomes)
    is_null = rng.uniform(size=n_outcomes) < null_frac
    for j in range(n_outcomes):
        a = rng.uniform(0.5, 2.0)
        ph = rng.uniform(0, 2 * np.pi)
        if is_null[j]:
            f = np.zeros(n)
        elif kinds[j] == 0:
            f = np.sin(2.0 * dz + ph)
        elif kinds[j] == 1:
            f = (sex - 0.5) * 2 * np.cos(1.5 * az + ph)
        elif kinds[j] == 2:
            f = 0.8 * rng.normal(size=n_subjects)[pid] + 0.5 * dz
        elif kinds[j] == 3:
            f = 0.7 * rng.normal(size=5)[site] + 0.6 * np.tanh(az)
        else:
            f = np.sin(1.2 * az + ph) + 0.5 * rng.normal(size=n_subjects)[pid]
        mu = a * f + rng.normal(scale=0.5, size=n)
        Y[:, j] = np.log1p(np.exp(4.0 + 0.6 * mu))       # log-normal-like intensities through log1p
    Yd = pd.DataFrame(Y, columns=[f"metabolite_{j}" for j in range(n_outcomes)])
    return X, Yd


def overview_notebook(n_people=100, n_observations=5, seed=9102):
    """Exact regeneration of waveome_overview.ipynb cell 4 (legacy np.random.seed RNG): X = person_id, time, female
    ("N"/"Y" strings), Y = outcome1..3, rows sorted by (person_id, time)."""
    total_obs = n_people * n_observations
    np.random.seed(seed)
    id_vec = np.repeat(np.arange(n_people), repeats=n_observations)
    time_vec = np.random.uniform(low=0, high=12, size=total_obs)
    female_vec = np.repeat(np.random.choice(a=["N", "Y"], size=n_people), repeats=n_observations)
    out1 = np.sin(time_vec)
    out2 = (female_vec == "Y") * np.cos(time_vec)
    out3 = 0.5 * time_vec + np.repeat(np.random.normal(scale=1.0, size=n_people), repeats=n_observations)
    out1 = out1 + np.random.normal(scale=0.1, size=total_obs)
    out2 = out2 + np.random.normal(scale=0.1, size=total_obs)
    out3 = out3 + np.random.normal(scale=0.1, size=total_obs)
    df = pd.DataFrame({"person_id": id_vec, "time": time_vec, "female": female_vec, "outcome1": out1,
                       "outcome2": out2, "outcome3": out3}).sort_values(["person_id", "time"])
    return df[["person_id", "time", "female"]], df[["outcome1", "outcome2", "outcome3"]]


def overview_synthetic(n_people=50, n_observations=10, n_outcomes=200, seed=9102):
    """Config 2: covariates from the generator of waveome_overview.ipynb cell 4 (50 subjects x 10 time points), and
    ``n_outcomes`` Gaussian outcomes cycling the notebook's archetypes {sin(t), female*cos(t), 0.5 t + subject offset,
    pure noise} with per-outcome amplitude U(0.5, 2) and phase U(0, 2 pi) from default_rng(seed + j), noise sd 0.1."""
    X, _ = overview_notebook(n_people, n_observations, seed)
    X = X.reset_index(drop=True)
    n = len(X)
    t = X["time"].to_numpy()
    fem = (X["female"] == "Y").to_numpy().astype(float)
    pid = X["person_id"].to_numpy().astype(int)
    Y = np.empty((n, n_outcomes))
    for j in range(n_outcomes):
        r = np.random.default_rng(seed + j)
        a, ph = r.uniform(0.5, 2.0), r.uniform(0, 2 * np.pi)
        k = j % 4
        if k == 0:
            f = a * np.sin(t + ph)
        elif k == 1:
            f = a * fem * np.cos(t + ph)
        elif k == 2:
            f = 0.5 * t + r.normal(size=n_people)[pid]
        else:
            f = r.normal(size=n)
        Y[:, j] = f + r.normal(scale=0.1, size=n)
    return X, pd.DataFrame(Y, columns=[f"outcome{j + 1}" for j in range(n_outcomes)])


def iris():
    """Config 1 (README quick-start): X = petal_length, petal_width, species; Y = sepal_length, sepal_width."""
    from sklearn.datasets import load_iris
    d = load_iris()
    df = pd.DataFrame(d.data, columns=["sepal_length", "sepal_width", "petal_length", "petal_width"])
    df["species"] = np.array(["setosa", "versicolor", "virginica"])[d.target]
    return df[["petal_length", "petal_width", "species"]], df[["sepal_length", "sepal_width"]]


def large_gpr(n_subjects=512, n_times=16, seed=11):
    """Config 4: n = 8192 longitudinal samples, truth SE(0.5)[t] x Cat[subject] + Periodic(period 3)[t], noise 0.1."""
    rng = np.random.default_rng(seed)
    n = n_subjects * n_times
    subj = np.repeat(np.arange(n_subjects), n_times)
    t = rng.uniform(0, 12, size=n)
    f = np.sin(2 * np.pi * t / 3.0) + 0.7 * np.sin(t / 0.8 + rng.uniform(0, 2 * np.pi, size=n_subjects)[subj])
    y = f + rng.normal(scale=0.1 ** 0.5, size=n)
    X = pd.DataFrame({"subject": subj.astype(float), "t": t})
    return X, pd.DataFrame({"y": y})


def count_microbiome(n_subjects=50, n_times=10, n_outcomes=1000, seed=7, family="poisson", alpha=1.0):
    """Config 5: n = 500 (50 x 10) longitudinal samples, ``n_outcomes`` count outcomes y ~ Poisson(exp(f)) or
    NB(mean exp(f), alpha), f = subject effect (sd 0.5) + smooth function of time + outcome-specific log abundance."""
    rng = np.random.default_rng(seed)
    n = n_subjects * n_times
    subj = np.repeat(np.arange(n_subjects), n_times)
    t = np.tile(np.linspace(0.0, 9.0, n_times), n_subjects) + rng.uniform(-0.3, 0.3, size=n)
    X = pd.DataFrame({"subject": subj.astype(float), "time": t})
    tz = (t - t.mean()) / t.std()
    Y = np.empty((n, n_outcomes))
    for j in range(n_outcomes):
        amp, ph, base = rng.uniform(0.2, 1.0), rng.uniform(0, 2 * np.pi), rng.uniform(0.0, 3.0)
        f = base + 0.5 * rng.normal(size=n_subjects)[subj] + amp * np.sin(1.5 * tz + ph)
        mu = np.exp(f)
        if family == "poisson":
            Y[:, j] = rng.poisson(mu)
        else:
            k = 1.0 / alpha
            Y[:, j] = rng.negative_binomial(k, k / (k + mu))
    return X, pd.DataFrame(Y, columns=[f"taxon_{j}" for j in range(n_outcomes)])
