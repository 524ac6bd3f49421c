"""Saturated additive kernel builder — host-side mirror of waveome/regularization.py:14-189
(``full_kernel_build``).  Component order and the frozen categorical variance inside
categorical x numeric products (reference :131-132) follow the reference exactly, because the order
defines the packed parameter vector and the component names."""
from __future__ import annotations

from . import kernels as K


def full_kernel_build(cat_vars=[], num_vars=[], unit_idx=None, var_names=None, second_order_numeric=False,
                      categorical_numeric_interactions=True, unit_numeric_interactions=False, return_sum=False,
                      kerns=None, num_outputs=None, ranks=None):
    if kerns is None:
        kerns = [K.SquaredExponential()]
    # multi-output ranks (reference :27-47): every term is repeated `rank` times (names get a _r suffix); an int applies
    # to all terms, a dict maps a covariate index to its rank, the default is num_outputs (1 for single-output models)
    default_rank = ranks if isinstance(ranks, int) else (num_outputs if num_outputs is not None else 1)

    def rank_of(idx):
        return ranks.get(idx, default_rank) if isinstance(ranks, dict) else default_rank

    def suffix(r, rk):
        return f"_{r}" if rk > 1 else ""
    kernel_list, var_list = [], []
    cat_vars = list(cat_vars)
    named = var_names is not None

    def nm(i):
        return var_names[i]

    if unit_idx is not None:
        cat_vars = [x for x in cat_vars if x != unit_idx]
        rk = rank_of(unit_idx)
        for r in range(rk):
            kernel_list.append(K.Categorical(active_dims=[unit_idx]))
            if named:
                var_list.append("categorical[" + nm(unit_idx) + "]" + suffix(r, rk))
    for c in cat_vars:
        rk = rank_of(c)
        for r in range(rk):
            kernel_list.append(K.Categorical(active_dims=[c]))
            if named:
                var_list.append("categorical[" + nm(c) + "]" + suffix(r, rk))
    for n in num_vars:
        rk = rank_of(n)
        for k in kerns:
            for r in range(rk):
                kc = K.deepcopy(k)
                kc.active_dims = [n]
                kernel_list.append(kc)
                if named:
                    var_list.append(kc.name + "[" + nm(n) + "]" + suffix(r, rk))
    rk = default_rank                  # interactions: the default rank (reference :100-105)
    if unit_numeric_interactions and unit_idx is not None:
        for n in num_vars:
            for k in kerns:
                for r in range(rk):
                    k1 = K.Categorical(active_dims=[unit_idx])
                    K.set_trainable(k1.variance, False)
                    k2 = K.deepcopy(k)
                    k2.active_dims = [n]
                    kernel_list.append(K.Product([k1, k2]))
                    if named:
                        var_list.append(f"{k1.name}[{nm(unit_idx)}]*{k2.name}[{nm(n)}]{suffix(r, rk)}")
    if categorical_numeric_interactions:
        for c in cat_vars:
            for n in num_vars:
                for k in kerns:
                    for r in range(rk):
                        k1 = K.Categorical(active_dims=[c])
                        K.set_trainable(k1.variance, False)
                        k2 = K.deepcopy(k)
                        k2.active_dims = [n]
                        kernel_list.append(K.Product([k1, k2]))
                        if named:
                            var_list.append(f"{k1.name}[{nm(c)}]*{k2.name}[{nm(n)}]{suffix(r, rk)}")
    if second_order_numeric:
        n_count = 0
        for n_first in num_vars:
            for k_first in kerns:
                for n_second in num_vars[n_count:]:
                    for k_second in kerns:
                        for r in range(rk):
                            k1 = K.deepcopy(k_first); k1.active_dims = [n_first]
                            k2 = K.deepcopy(k_second); k2.active_dims = [n_second]
                            kernel_list.append(K.Product([k1, k2]))
                            if named:
                                var_list.append(f"{k1.name}[{nm(n_first)}]*{k2.name}[{nm(n_second)}]{suffix(r, rk)}")
            n_count += 1
    out = K.Sum(kernel_list) if return_sum else kernel_list
    return (out, var_list) if named else out
