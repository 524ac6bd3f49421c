"""Shared builders for the parity tests: random data and kernel trees in the product's classes."""
import numpy as np

import waveome_b200 as wb


def make_data(n, seed=0, n_subj=None):
    rng = np.random.default_rng(seed)
    n_subj = n_subj or max(2, n // 5)
    subj = rng.integers(0, n_subj, size=n).astype(float)
    t = rng.normal(size=n)
    z = rng.normal(size=n)
    sex = rng.integers(0, 2, size=n).astype(float)
    X = np.stack([subj, t, z, sex], 1)
    y = (np.sin(2 * t) + 0.5 * sex * np.cos(t) + 0.3 * rng.normal(size=n_subj)[subj.astype(int)]
         + 0.1 * rng.normal(size=n))
    return X, y


def all_leaf_kernel(hs=1.0):
    """Every leaf type, sums and products, one frozen factor, horseshoe on variances."""
    def V(k):
        for path, p in k.named_parameters():
            if "variance" in path and p.trainable:
                p.prior = wb.Horseshoe(hs) if hs else None
        return k
    cat_frozen = wb.Categorical(active_dims=[3]); wb.set_trainable(cat_frozen.variance, False)
    ks = [
        V(wb.Categorical(active_dims=[0])),
        V(wb.SquaredExponential(active_dims=[1])),
        V(wb.Matern12(active_dims=[2])),
        V(wb.Matern32(active_dims=[1], lengthscales=0.7)),
        V(wb.Matern52(active_dims=[2], lengthscales=1.3)),
        V(wb.Periodic(wb.SquaredExponential(active_dims=[1]), period=1.7)),
        V(wb.Lin(active_dims=[2], variance=0.5)),
        V(wb.Constant(variance=0.3)),
        V(wb.Poly(active_dims=[1], variance=0.2, offset=0.5, degree=3)),
        wb.Product([cat_frozen, V(wb.SquaredExponential(active_dims=[1], lengthscales=0.8))]),
        wb.Product([V(wb.Lin(active_dims=[1])), V(wb.Matern32(active_dims=[2]))]),
    ]
    return wb.Sum(ks)


def saturated_kernel(cat=(0, 3), num=(1, 2), unit=0, hs=1.0):
    """full_kernel_build-like: unit cat + cats + SE per numeric + cat x SE (cat variance frozen)."""
    ks = [wb.Categorical(active_dims=[unit])]
    cats = [c for c in cat if c != unit]
    ks += [wb.Categorical(active_dims=[c]) for c in cats]
    ks += [wb.SquaredExponential(active_dims=[d]) for d in num]
    for c in cats:
        for d in num:
            k1 = wb.Categorical(active_dims=[c]); wb.set_trainable(k1.variance, False)
            ks.append(wb.Product([k1, wb.SquaredExponential(active_dims=[d])]))
    k = wb.Sum(ks)
    if hs:
        for path, p in k.named_parameters():
            if "variance" in path and p.trainable:
                p.prior = wb.Horseshoe(hs)
    return k
