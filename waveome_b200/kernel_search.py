"""Compositional kernel search (``GPSearch.run_search``) on the batch engine — BASELINE configs[1].

Host-side restatement of the reference's greedy search (waveome/model_search.py):

    kernel_test              :2239-2334   one candidate fit -> (model, "BIC" = round(2k - 2 log p, 2))
    set_feature_kernels      :2337-2344
    loc_kernel_search        :2347-2558   candidates of one base kernel: first level / sum / product / split product
    prod_kernel_creation     :2561-2664   product of a new factor with each additive component of the base
    check_if_better_metric   :2667-2681
    keep_top_k               :2684-2710   BIC window (metric_diff) -> ``try_next`` flags
    prune_best_model2        :2778-2885   leave-one-component-out refits of the level's best model
    prune_prod_kernel        :2888-2984   ... and of the factors of its product components
    full_kernel_search       :2987-3272   depth loop, early stopping, final arg-min by (bic, depth, name)

The selection arithmetic (names, ordering by string comparison, de-duplication by canonicalised name, windows,
what is frozen in a product) follows the reference decision by decision, including its quirks, because it determines
*which fits exist* and which structure is selected.  What changes is the execution: every function here is a
GENERATOR that yields the list of candidate kernels it needs fitted and receives ``(model, bic)`` pairs back, so a
driver can run the searches of many outcomes in lock-step and hand the union of their requests to the engine as ONE
batch ("outcomes x candidate kernel structures") instead of one process per outcome and one TensorFlow
graph per candidate.  ``run_lockstep`` is that driver; the plain functions of the reference's names
(``full_kernel_search`` ...) run a single search with one engine batch per request.

Objective: the exact-GPR log marginal likelihood (objective A) maximised by the device L-BFGS-B; the
reference's live path scores candidates with an Adam/NatGrad-fitted SVGP bound of the same quantity.
"""
from __future__ import annotations

import re
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import os

import numpy as np

from . import kernels as K
from .models import GPR, ConstantMean
from .utilities import calc_bic, canonical_model_name, check_if_model_exists

Candidate = Tuple[str, K.Kernel]                 # (name key, kernel to fit)
Fitted = Tuple[Optional[GPR], float]             # (model or None if the fit failed, bic)


# ------------------------------------------------------------------------------------------------
# candidate construction
# ------------------------------------------------------------------------------------------------
def kernel_info(k: K.Kernel) -> str:
    """``k.name + str(k.active_dims)`` (reference :2395, :2588); the frozen first-level constant is "constant"."""
    return k.name if k.name == "constant" else k.name + "[" + str(int(k.active_dims[0])) + "]"


def set_feature_kernels(f, kern_list, cat_vars):
    """:2337-2344 — a categorical column gets the categorical kernel, any other column every kernel of the list."""
    if f in cat_vars:
        return [K.Categorical(active_dims=[f])]
    k_list = list(kern_list)
    for k_ in k_list:
        k_.active_dims = [int(f)]
    return k_list


def _freeze_new_factor(k):
    """:2464-2468 / :2513-2516 — the factor added by a product keeps variance 1 (Periodic: on its base kernel)."""
    K.set_trainable(k.base_kernel.variance if k.name == "periodic" else k.variance, False)


def _entry(kernel, model, bic, depth, parent):
    return {"kernel": kernel, "model": model, "bic": bic, "depth": depth, "parent": parent, "try_next": True}


def _feature_kernel_names(f, kern_list, cat_vars, depth):
    """Names of ``set_feature_kernels(f, ...)`` (+ the frozen first-level constant, :2387-2392) without building them."""
    names = ["categorical"] if f in cat_vars else [x.name for x in kern_list]
    if f == 0 and depth == 1:
        names = names + ["constant"]
    return names


def loc_candidates(n_features, kern_list, base_kern=None, base_name=None, cat_vars=(), depth=0, operation="sum",
                   prev_models=None, build=True) -> List[Candidate]:
    """Candidate kernels of one ``loc_kernel_search`` call (:2347-2558), in the reference's order.

    Which candidates exist depends on names only, so kernels are copied and composed only for the candidates that
    survive the name checks (the reference deep-copies the base for every feature x kernel pair first);
    ``build=False`` returns the names alone (kernel = None)."""
    prev_set = {canonical_model_name(y) for y in prev_models} if prev_models is not None else set()
    cat_vars = list(cat_vars)
    cands: List[Candidate] = []
    reset = []

    def base_copy():
        """deepcopy(base_kern) with every trainable parameter back at 1.0 (:2414-2416)"""
        if not reset:
            b = K.deepcopy(base_kern)
            for p in b.trainable_parameters:
                p.assign(1.0)
            reset.append(b)
        return K.deepcopy(reset[0])

    for f in range(n_features):
        k_list: list = []

        def leaf(j):
            if not k_list:
                k_list.extend(set_feature_kernels(f, [K.deepcopy(x) for x in kern_list], cat_vars))
                if f == 0 and depth == 1:
                    empty_kernel = K.Constant(variance=1e-6)
                    K.set_trainable(empty_kernel.variance, False)
                    k_list.append(empty_kernel)
            return k_list[j]

        for j, kname in enumerate(_feature_kernel_names(f, kern_list, cat_vars, depth)):
            k_info = kname if kname == "constant" else kname + "[" + str(int(f)) + "]"
            if base_kern is None:
                cands.append((k_info, leaf(j) if build else None))
                continue
            if operation in ("sum", "product"):
                if "categorical[" + str(f) + "]" in base_name:
                    continue
                if operation == "product" and "*" in base_name:               # two-way interactions only
                    continue
                sep, node = ("+", K.Sum) if operation == "sum" else ("*", K.Product)
                base_first = base_name < k_info
                name = base_name + sep + k_info if base_first else k_info + sep + base_name
                if canonical_model_name(name) in prev_set:
                    continue
                if not build:
                    cands.append((name, None))
                    continue
                k = leaf(j)
                if operation == "product":
                    _freeze_new_factor(k)
                cands.append((name, node([base_copy(), k]) if base_first else node([k, base_copy()])))
            elif operation == "split_product":
                def new_factor(j=j):
                    k = leaf(j)
                    _freeze_new_factor(k)
                    return k
                cands += _prod_candidates(len(base_kern.kernels), base_name, k_info, f, prev_set,
                                          base_copy if build else None, new_factor)
            else:
                raise ValueError(f"unknown operation {operation!r}")
    return cands


def loc_kernel_search(n_features, kern_list, base_kern=None, base_name=None, cat_vars=(), depth=0, operation="sum",
                      prev_models=None, cache=None):
    """:2347-2558 as a generator: yields ONE request with every candidate of this call that is not in ``cache``
    (name -> (kernel, (model, bic)) fitted ahead of time), returns the result dict (failed fits drop out, as the
    reference's ``except Exception: None`` does)."""
    parents = "None" if base_kern is None else base_name
    cache = cache if cache is not None else {}
    cands = loc_candidates(n_features, kern_list, base_kern, base_name, cat_vars, depth, operation, prev_models,
                           build=not cache)
    if cache and any(c[0] not in cache for c in cands):       # a fit of the level failed: build what is still missing
        cands = loc_candidates(n_features, kern_list, base_kern, base_name, cat_vars, depth, operation, prev_models)
    missing = [c for c in cands if c[0] not in cache]
    if missing:
        got = yield missing
        for c, r in zip(missing, got):
            cache[c[0]] = (c[1], r)
    cands = [(name, cache[name][0]) for name, _k in cands]
    fitted = [cache[name][1] for name, _k in cands]
    out = {}
    for (name, k), (m, bic) in zip(cands, fitted):
        if m is not None:
            out[name] = _entry(k, m, bic, depth, parents)
    return out


def _prod_candidates(n_components, base_name, k_info, f, prev_set, base_copy, new_factor) -> List[Candidate]:
    """:2561-2664 — the new factor ``k_info`` (on column ``f``) times each additive component of the base.  When the new
    factor sorts before the component, the reference moves the component's NAME inside the '+'-joined key but leaves
    the kernel in place; later stages index names and kernels in parallel, so that is kept as it is.
    ``base_copy()`` -> a fresh copy of the base sum, ``new_factor()`` -> the new factor kernel; ``base_copy=None``
    returns names only."""
    out: List[Candidate] = []
    for feat in range(n_components):
        temp_name = base_name.split("+")
        if "categorical[" + str(int(f)) + "]" in temp_name[feat]:
            continue
        if "*" in temp_name[feat]:
            continue
        comp_first = temp_name[feat] < k_info
        if comp_first:
            temp_name[feat] = temp_name[feat] + "*" + k_info
        else:
            later = [i for i, x in enumerate(temp_name) if k_info < x]
            new_idx = later[0] if later else len(temp_name) - 1
            cur_component_name = temp_name.pop(feat)
            temp_name.insert(new_idx, k_info + "*" + cur_component_name)
        name = "+".join(temp_name)
        if canonical_model_name(name) in prev_set:
            continue
        if base_copy is None:
            out.append((name, None))
            continue
        temp_kernel, new_kernel = base_copy(), new_factor()
        temp_kernel.kernels[feat] = K.Product([temp_kernel.kernels[feat], new_kernel] if comp_first
                                              else [new_kernel, temp_kernel.kernels[feat]])
        out.append((name, temp_kernel))
    return out


def prod_kernel_candidates(base_kernel, base_name, new_kernel, prev_models) -> List[Candidate]:
    """:2561-2664 with the reference's arguments (``base_kernel``: a sum kernel; ``new_kernel``: the factor)."""
    return _prod_candidates(len(base_kernel.kernels), base_name, kernel_info(new_kernel), int(new_kernel.active_dims[0]),
                            {canonical_model_name(y) for y in prev_models}, lambda: K.deepcopy(base_kernel),
                            lambda: new_kernel)


def prod_kernel_creation(base_kernel, base_name, new_kernel, depth, prev_models=()):
    """Generator form of :2561-2664 for callers that use it on its own."""
    cands = prod_kernel_candidates(base_kernel, base_name, new_kernel, list(prev_models))
    fitted = yield cands
    return {name: _entry(k, m, bic, depth, base_name) for (name, k), (m, bic) in zip(cands, fitted) if m is not None}


# ------------------------------------------------------------------------------------------------
# selection
# ------------------------------------------------------------------------------------------------
def check_if_better_metric(model_dict, depth):
    """:2667-2681"""
    prev_vals = [x["bic"] for x in model_dict.values() if x["depth"] == depth - 1]
    new_vals = [x["bic"] for x in model_dict.values() if x["depth"] == depth]
    if len(prev_vals) > 0 and len(new_vals) > 0:
        return min(new_vals) < min(prev_vals)
    return False


def keep_top_k(res_dict, depth, metric_diff=6, split=False):
    """:2684-2710 — entries of this depth further than ``metric_diff`` from its best are not expanded."""
    t = np.log(metric_diff) if split else metric_diff
    best_bic = min(v["bic"] for v in res_dict.values() if v["depth"] == depth)
    for v in res_dict.values():
        if v["depth"] == depth and v["bic"] - best_bic > t:
            v["try_next"] = False
    return res_dict.copy()


def _reset(k):
    for p in k.trainable_parameters:
        p.assign(1.0)
    return k


def _prune_prod_candidates(prod_kernel, prod_name, other_kernel=None, other_name="") -> List[Candidate]:
    """Candidate list of :2888-2984 — each factor of a product component on its own (plus the other components)."""
    other_kernel = K.deepcopy(other_kernel)
    prod_kernel = K.deepcopy(prod_kernel)
    kernel_parts = prod_name.split("*")
    out: List[Candidate] = []
    if prod_kernel.name != "product":
        return out
    for i in range(len(prod_kernel.kernels)):
        if i >= len(kernel_parts):
            break
        new_piece = kernel_parts[i]
        if other_name == "":
            k_info, k = new_piece, prod_kernel.kernels[i]
        else:
            names = [other_name, new_piece]
            order_set = sorted(range(2), key=lambda t: names[t])          # np.argsort of the two strings
            k_info = "+".join(names[t] for t in order_set)
            if not isinstance(other_kernel, list):
                other_kernel = [other_kernel]
            # the reference indexes ``np.array(other_kernel + [factor])`` with the length-2 permutation: with two or
            # more other components only the first two entries survive (the factor itself is then not in the kernel)
            pool = other_kernel + [prod_kernel.kernels[i]]
            k = K.Sum([pool[t] for t in order_set])
        out.append((k_info, _reset(K.deepcopy(k))))
    return out


def prune_prod_kernel(prod_kernel, prod_name, res_dict, best_bic, best_model_name, depth, other_kernel=None,
                      other_name=""):
    """:2888-2984 as a generator (one request with the factors that have not been fitted yet)."""
    out_dict = res_dict.copy()
    cands = [c for c in _prune_prod_candidates(prod_kernel, prod_name, other_kernel, other_name)
             if not check_if_model_exists(c[0], list(res_dict.keys()))]
    fitted = (yield cands) if cands else []
    for (k_info, _k), (m, bic) in zip(cands, fitted):
        if m is not None and bic < best_bic:
            out_dict[k_info] = _entry(m.kernel, m, bic, depth, best_model_name)
    return out_dict


def prune_best_model2(res_dict, depth):
    """:2778-2885 — leave-one-component-out refits of the best model of ``depth``; better ones join the dict.

    All refits of one call are independent fits, so they are requested together; the bookkeeping is then replayed in
    the reference's order (a product component's duplicate check sees what earlier components added)."""
    best_bic, best_model_name, best_model = min(
        ((i["bic"], k, i["model"]) for k, i in res_dict.items() if i["depth"] == depth), key=lambda t: (t[0], t[1]))
    best_model = K.deepcopy(best_model)
    kernel_names = re.split(r"\+", best_model_name)
    if len(kernel_names) <= 1 and "*" not in kernel_names[0]:
        return res_dict
    model_kernels = list(getattr(best_model.kernel, "kernels", []))
    plan = []           # (kind, candidates) per component, in order
    for i in range(len(kernel_names)):
        k_info = "+".join(x_ for i_, x_ in enumerate(kernel_names) if i_ != i)
        kerns = [k_ for i_, k_ in enumerate(model_kernels) if i_ != i]
        if "*" in kernel_names[i]:
            if len(kernel_names) == 1:
                prod_kernel, other_kernel, other_name = best_model.kernel, None, ""
            elif i < len(model_kernels):
                prod_kernel, other_kernel, other_name = model_kernels[i], kerns, k_info
            else:
                continue
            plan.append(("prod", _prune_prod_candidates(prod_kernel, kernel_names[i], other_kernel, other_name)))
            continue
        if not kerns:
            continue
        k = K.Sum(kerns) if len(kerns) > 1 else kerns[0]
        plan.append(("drop", [(k_info, _reset(K.deepcopy(k)))]))
    known = list(res_dict.keys())
    request = [c for _kind, cands in plan for c in cands if not check_if_model_exists(c[0], known)]
    fitted = (yield request) if request else []
    result = {id(c[1]): r for c, r in zip(request, fitted)}
    out_dict = res_dict.copy()
    for kind, cands in plan:
        seen = list(res_dict.keys()) if kind == "drop" else list(out_dict.keys())
        for k_info, k in cands:
            if check_if_model_exists(k_info, seen) or id(k) not in result:
                continue
            m, bic = result[id(k)]
            if m is not None and bic < best_bic:
                out_dict[k_info] = _entry(m.kernel, m, bic, depth, best_model_name)
    return out_dict


def _best_of_depth(search_dict, d):
    return min((i["bic"], i["depth"], k) for k, i in search_dict.items() if i["depth"] == d)[2]


def softmax_kernel_selection(bic_list, name_list):
    """:3535-3567 — draw ONE model name with probability softmax of the min-max normalised negated criteria (models whose
    criterion is inf are left out; a single candidate is returned as it is).  Draws from ``np.random`` like the reference."""
    name_list = [name_list[x] for x in range(len(bic_list)) if bic_list[x] != np.inf]
    bic_list = [x for x in bic_list if x != np.inf]
    if len(bic_list) == 1:
        return name_list[0]
    neg = np.array([-x for x in bic_list])
    norm = (neg - min(neg)) / (max(neg) - min(neg))
    prob = np.exp(norm) / sum(np.exp(norm))
    return name_list[np.random.choice(a=np.arange(len(prob)), p=prob)]


class BicRequests(list):
    """Fit requests whose results must be scored by BIC on the training data even in a hold-out (split) search: the
    reference prunes with ``prune_best_model2(...)`` WITHOUT the hold-out arguments (:3441-3449, :3483-3491)."""


def _scored_by_bic(gen):
    """Re-yield the requests of ``gen`` as BicRequests."""
    try:
        req = next(gen)
        while True:
            req = gen.send((yield BicRequests(req)))
    except StopIteration as e:
        return e.value


def full_kernel_search_gen(n_features, kern_list, cat_vars=(), max_depth=5, keep_all=False, metric_diff=6,
                           early_stopping=True, prune=True, keep_only_best=True, softmax_select=False, split=False):
    """:2987-3272 as a generator over fit requests; returns {"models", "edges", "best_model", "var_exp"}.

    ``split`` = True gives the level logic of ``split_kernel_search`` (:3338-3500) instead: the criterion the fitter returns
    is the negated hold-out log density, ``keep_top_k`` uses its log window, there is no stop at a constant best model,
    and pruning happens once -- when the search stops early or reaches ``max_depth`` -- scored by BIC (BicRequests).
    ``softmax_select`` (:3190-3204, :3467-3481): after ``keep_top_k`` one model of the whole dictionary is drawn by
    ``softmax_kernel_selection`` and only it stays expandable at this depth (draws from ``np.random``: per-outcome
    results then depend on the order in which the generators are advanced)."""
    search_dict: Dict[str, dict] = {}
    edge_list = []
    for d in range(1, max_depth + 1):
        if d == 1:
            search_dict = yield from loc_kernel_search(n_features, kern_list, cat_vars=cat_vars, depth=d)
        else:
            bases = [k for k in search_dict.keys()
                     if search_dict[k]["depth"] == d - 1 and search_dict[k]["try_next"] is not False and k != "constant"]
            # The reference fits the expansions of one base kernel at a time and de-duplicates each against everything
            # fitted so far.  Which candidates exist depends only on names unless a fit fails, so the whole level is
            # generated ahead under the assumption that every fit succeeds and requested as ONE batch; the exact
            # sequential bookkeeping below then finds its fits in the cache (and asks for more only after a failure).
            cache, ahead, names_ahead = {}, [], list(search_dict.keys())
            for k in bases:
                cur_kern = search_dict[k]["kernel"]
                for op in ("sum", "split_product" if cur_kern.name == "sum" else "product"):
                    cs = loc_candidates(n_features, kern_list, cur_kern, k, cat_vars, d, op, names_ahead)
                    ahead += cs
                    names_ahead += [c[0] for c in cs]
            if ahead:
                got = yield ahead
                for c, r in zip(ahead, got):
                    cache.setdefault(c[0], (c[1], r))
            temp_dict = search_dict.copy()
            for k in bases:
                cur_kern = search_dict[k]["kernel"]
                new_res = yield from loc_kernel_search(n_features, kern_list, base_kern=cur_kern, base_name=k,
                                                       cat_vars=cat_vars, depth=d, operation="sum",
                                                       prev_models=temp_dict.keys(), cache=cache)
                temp_dict.update(new_res)
                edge_list += [(k, k_) for k_ in new_res.keys()]
                op = "split_product" if cur_kern.name == "sum" else "product"
                new_res = yield from loc_kernel_search(n_features, kern_list, base_kern=cur_kern, base_name=k,
                                                       cat_vars=cat_vars, depth=d, operation=op,
                                                       prev_models=temp_dict.keys(), cache=cache)
                temp_dict.update(new_res)
                edge_list += [(k, k_) for k_ in new_res.keys()]
            search_dict = temp_dict
        if not any(i["depth"] == d for i in search_dict.values()):
            break          # (the reference raises on the empty min(); nothing left to expand)
        best_model_name = _best_of_depth(search_dict, d)
        if best_model_name == "constant" and not split:
            break
        if early_stopping and d > 1:
            if not check_if_better_metric(search_dict, d):
                if prune:
                    search_dict = yield from (_scored_by_bic(prune_best_model2(search_dict, depth=d)) if split
                                              else prune_best_model2(search_dict, depth=d))
                break
        if d != max_depth:
            if not keep_all:
                search_dict = keep_top_k(search_dict, depth=d, metric_diff=metric_diff, split=split)
            if softmax_select:
                info = [(i["bic"], k) for k, i in search_dict.items()]
                chosen = softmax_kernel_selection([x[0] for x in info], [x[1] for x in info])
                for k, v in search_dict.items():
                    if v["depth"] == d and k != chosen:
                        v["try_next"] = False
        if prune and not split:
            search_dict = yield from prune_best_model2(search_dict, depth=d)
        elif prune and d == max_depth:
            search_dict = yield from _scored_by_bic(prune_best_model2(search_dict, depth=d))
    best_model_name = min((i["bic"], i["depth"], k) for k, i in search_dict.items())[2]
    if keep_only_best:
        search_dict = {best_model_name: search_dict[best_model_name]}
    return {"models": search_dict, "edges": edge_list, "best_model": best_model_name, "var_exp": None}


# ------------------------------------------------------------------------------------------------
# execution: fitters and the lock-step driver
# ------------------------------------------------------------------------------------------------
def candidate_model(kernel, mean_function=None, likelihood="gaussian") -> GPR:
    """The model ``kernel_test`` builds (:2269-2282): penalisation 0 (no prior), Gaussian noise 1.0 (or the count
    likelihood of the search), constant mean.  The model owns deep copies (BaseGP.__init__,
    waveome/model_classes.py:110-111)."""
    from .models import make_likelihood
    return GPR(K.deepcopy(kernel), mean_function=K.deepcopy(mean_function) if mean_function is not None else ConstantMean(),
               likelihood=make_likelihood(likelihood))


def candidate_bic(model: GPR, log_posterior_density: float) -> float:
    """:2311-2321 — round(2 k - 2 log p, 2), k = number of trainable Parameter objects."""
    return round(calc_bic(loglik=log_posterior_density, n=0, k=len(model.trainable_parameters)), 2)


#: run_lockstep: a level's fit hands back control once at most this many models are still iterating; the stragglers finish
#: in the background while the other outcomes go on (0 / WV_SEARCH_TAIL=0: every level waits for its slowest model)
SEARCH_TAIL = max(0, int(os.environ.get("WV_SEARCH_TAIL", "32")))
#: ... and this many level fits are in flight at a time (one group of outcomes each, on fitter threads with their own
#: engines): the device works on one group while the host advances the other's generators (WV_SEARCH_LANES)
SEARCH_LANES = max(1, int(os.environ.get("WV_SEARCH_LANES", "2")))
_FITTER_SLOT = __import__("threading").local()      # .slot = index of a run_lockstep fitter thread (unset elsewhere)


def _thread_engine():
    """The process engine, except on a ``run_lockstep`` fitter thread: fitter i drives engine i + 1 of the device pool
    (an engine is driven by one host thread at a time; the pool is reused by later searches)."""
    from .model_fitting import get_engine, get_engine_pool
    slot = getattr(_FITTER_SLOT, "slot", None)
    if slot is None:
        return get_engine()
    return get_engine_pool(slot + 2)[slot + 1]


def engine_fitter(X: np.ndarray, engine=None, num_restart=1, random_seed=None, max_iter=50000,
                  likelihood="gaussian", optimizer="lbfgs") -> Callable:
    """Returns ``fit(requests) -> results`` where requests is a list of (y [n], name, kernel): all of them become one
    engine batch (restarts included: ``num_restart`` > 1 adds randomised starts as extra models of the batch,
    waveome/model_classes.py:472-524, seeds ``random_seed + 1 + r`` or ``r``)."""
    from .model_fitting import fit_models

    def collect(requests, models, res, R, which=None):
        out = []
        for i in range(len(requests)):
            if which is not None and not which[i]:
                out.append(None)
                continue
            best, best_lpd = None, -np.inf
            for r in range(R):
                b = i * R + r
                ok = not (int(res["status"][b]) & 1) and np.isfinite(res["f"][b])
                if ok and -float(res["f"][b]) > best_lpd:
                    best, best_lpd = models[b], -float(res["f"][b])
            out.append((None, np.inf) if best is None else (best, candidate_bic(best, best_lpd)))
        return out

    def fit(requests, tail=0):
        """results = [(model, bic)] per request.  ``tail`` > 0 (run_lockstep): returns (results, pending) as soon as at
        most ``tail`` models are still iterating -- results[i] is None for a request with an unfinished model, and
        ``pending()`` (callable from a worker thread; None when nothing is left) finishes them and returns the complete
        list."""
        if not requests:
            return ([], None) if tail else []
        R = max(1, int(num_restart))
        models, ys = [], []
        for y, _name, kernel in requests:
            for r in range(R):
                m = candidate_model(kernel, likelihood=likelihood)
                if R > 1:
                    rs = np.random.RandomState(r if random_seed is None else random_seed + 1 + r)
                    for p in m.trainable_parameters:
                        p.assign(p.transform_fn(rs.normal(loc=0.0, scale=1.0)))
                models.append(m)
                ys.append(y)
        if not tail:
            res = fit_models(X, np.stack(ys), models, engine=engine or _thread_engine(), maxiter=max_iter, maxfun=max_iter,
                             optimizer=optimizer)
            return collect(requests, models, res, R)
        res = fit_models(X, np.stack(ys), models, engine=engine or _thread_engine(), maxiter=max_iter, maxfun=max_iter,
                         optimizer=optimizer, tail=tail)
        if res["pending"] is None:
            return collect(requests, models, res, R), None
        done = res["finished"].reshape(len(requests), R).all(axis=1)
        finish = res["pending"]
        return collect(requests, models, res, R, which=done), lambda: collect(requests, models, finish(), R)

    fit.supports_tail = optimizer == "lbfgs"
    return fit


def run_lockstep(searches: Dict[str, object], ys: Dict[str, np.ndarray], fit: Callable,
                 groups: Optional[int] = None) -> Dict[str, dict]:
    """Advance the search generators of many outcomes together: the requests they are waiting on are fitted as one
    batch per round.  ``searches``: outcome -> generator; ``ys``: outcome -> y [n].

    ``groups`` > 1 splits the outcomes into that many contiguous groups and runs the fits on worker threads (env
    WV_SEARCH_FITTERS of them, default 1, one engine each; the engine call releases the GIL): while group A's batch is
    on the device, the host advances group B's generators with the results it already has.  Every outcome sees exactly
    the results it would see alone -- fits do not depend on the composition of their batch -- so the outcome of the
    search is the same for any ``groups`` (tested).  Default 1, on measurement (config 2, 200 outcomes): a level batch's
    duration is set by its slowest models, so two half batches cost more than one (18.3 s with one fitter thread,
    17.1 s with two, against 16.0 s for the single group)."""
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor
    names = list(searches.keys())
    if groups is None:
        groups = 1
    groups = max(1, min(int(groups), len(names) or 1))
    bounds = [len(names) * g // groups for g in range(groups + 1)]
    members = [names[bounds[g]: bounds[g + 1]] for g in range(groups)]
    done: Dict[str, dict] = {}
    waiting: List[Dict[str, list]] = [{} for _ in range(groups)]
    for g in range(groups):
        for o in members[g]:
            try:
                waiting[g][o] = next(searches[o])
            except StopIteration as e:
                done[o] = e.value
    rounds = 0

    def flatten(g):
        return [(ys[o], name, k) for o, cands in waiting[g].items() for name, k in cands]

    def advance(g, results):
        pos, nxt = 0, {}
        for o, cands in waiting[g].items():
            r = results[pos: pos + len(cands)]
            pos += len(cands)
            try:
                nxt[o] = searches[o].send(r)
            except StopIteration as e:
                done[o] = e.value
        waiting[g] = nxt

    tail = SEARCH_TAIL if getattr(fit, "supports_tail", False) else 0
    if groups == 1 and tail > 0:
        # A level's batch lasts as long as its slowest model (config 2: one candidate of 1931 needs 5019 evaluations, the
        # others at most 386), and while the host advances the generators the device has nothing to do.  So:
        #  * a fit returns once at most `tail` models are still iterating; the outcomes whose candidates are all fitted move
        #    on at once, the stragglers finish on a worker thread (their batch re-homed onto a high-priority engine) and
        #    their outcomes rejoin whatever batch is formed next;
        #  * SEARCH_LANES fits are in flight at a time, each on a fitter thread with its own engine (the C call releases
        #    the GIL): the outcomes start as that many groups, and while one group's batch is on the device the host
        #    advances the generators of the group that has just come back.
        # Every outcome still sees exactly the results it would see alone, so the search result does not depend on the
        # grouping (tested).
        import itertools
        import threading
        from concurrent.futures import FIRST_COMPLETED, wait
        slots, slot_lock = itertools.count(), threading.Lock()

        def take_lane_slot():
            with slot_lock:
                _FITTER_SLOT.slot = next(slots)

        lanes = max(1, min(SEARCH_LANES, len(waiting[0]) or 1))
        with ThreadPoolExecutor(max_workers=lanes, initializer=take_lane_slot) as main_pool, \
                ThreadPoolExecutor(max_workers=8) as late_pool:
            ready = dict(waiting[0])
            inflight = {}                           # future -> ("main", {outcome: candidates}) | ("late", {outcome: (pos, count)})
            n_main = 0

            def step(o, results):
                try:
                    ready[o] = searches[o].send(results)
                except StopIteration as e:
                    done[o] = e.value

            def launch(first=False):
                nonlocal n_main
                while ready and n_main < lanes:
                    names_now = list(ready)
                    if first:                       # the initial split: one group per lane
                        names_now = names_now[: -(-len(names_now) // (lanes - n_main))]
                    cur = {o: ready.pop(o) for o in names_now}
                    reqs = [(ys[o], name, k) for o, cands in cur.items() for name, k in cands]
                    inflight[main_pool.submit(fit, reqs, tail=tail)] = ("main", cur)
                    n_main += 1

            launch(first=True)
            while inflight:
                finished, _ = wait(list(inflight), return_when=FIRST_COMPLETED)
                for fut in finished:
                    kind, info = inflight.pop(fut)
                    if kind == "main":
                        n_main -= 1
                        results, pending = fut.result()
                        rounds += 1
                        pos, held = 0, {}
                        for o, cands in info.items():
                            r = results[pos: pos + len(cands)]
                            if any(x is None for x in r):
                                held[o] = (pos, len(cands))
                            else:
                                step(o, r)
                            pos += len(cands)
                        if pending is not None:
                            inflight[late_pool.submit(pending)] = ("late", held)
                    else:
                        full = fut.result()
                        for o, (pos, k) in info.items():
                            step(o, full[pos: pos + k])
                    launch()
    elif groups == 1:
        while waiting[0]:
            results = fit(flatten(0))
            rounds += 1
            advance(0, results)
    else:
        import itertools
        import threading
        slots, slot_lock = itertools.count(), threading.Lock()

        def take_slot():
            with slot_lock:
                _FITTER_SLOT.slot = next(slots)

        with ThreadPoolExecutor(max_workers=max(1, int(os.environ.get("WV_SEARCH_FITTERS", "1"))),
                                initializer=take_slot) as pool:
            pending = deque()
            for g in range(groups):
                if waiting[g]:
                    pending.append((g, pool.submit(fit, flatten(g))))
            while pending:
                g, fut = pending.popleft()
                results = fut.result()
                rounds += 1
                advance(g, results)                              # host work, behind the next group's batch
                if waiting[g]:
                    pending.append((g, pool.submit(fit, flatten(g))))
    for v in done.values():
        v["batches"] = rounds
    return done


def _drive_single(gen, y, fit):
    return run_lockstep({"_": gen}, {"_": y}, fit)["_"]


def kernel_test(X, Y, k, mean_function=None, num_restart=5, random_init=True, random_seed=None, verbose=False,
                likelihood="gaussian", engine=None, keep_data=False, optimizer="lbfgs", **unused):
    """Drop-in for :2239-2334: (fitted model, bic).  ``likelihood``: any name ``models.make_likelihood`` covers (it
    raises NotImplementedError for the others).  ``optimizer``: "lbfgs" (default: L-BFGS-B reaches the optimum of the
    collapsed objective) or "adam" -- upstream's own choice here, ``optimize_params()`` with its "adam/gradient" default
    (:2297), run on the device with the same schedule."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(Y, dtype=np.float64).reshape(-1)
    fit = engine_fitter(X, engine=engine, num_restart=num_restart if random_init else 1, random_seed=random_seed,
                        likelihood=likelihood, optimizer=optimizer)
    (m, bic), = fit([(y, "", k)])
    if m is None:
        raise RuntimeError("kernel_test: the fit failed (Cholesky failure or non-finite objective at the start point)")
    if verbose:
        from .utilities import print_kernel_names
        print(f"Model: {print_kernel_names(k)}, BIC: {bic}")
    m.data = (X, y.reshape(-1, 1)) if keep_data else None
    return m, bic


def full_kernel_search(X, Y, kern_list, cat_vars=(), max_depth=5, keep_all=False, metric_diff=6, early_stopping=True,
                       prune=True, num_restart=5, lik="gaussian", verbose=False, debug=False, keep_only_best=True,
                       softmax_select=False, random_seed=None, feature_name=None, engine=None, fit=None, **unused):
    """Drop-in for :2987-3272 (one outcome).  ``fit`` overrides the engine fitter (the tests pass the CPU oracle)."""
    if random_seed is not None:
        np.random.seed(random_seed)
    Xn = X.to_numpy() if hasattr(X, "to_numpy") else np.asarray(X)
    Xn = np.asarray(Xn, dtype=np.float64).reshape(len(Xn), -1)
    if hasattr(Y, "to_numpy"):
        Yn = (Y if feature_name is None else Y[feature_name]).to_numpy()
    else:
        Yn = np.asarray(Y)
    y = np.asarray(Yn, dtype=np.float64).reshape(-1)
    ok = ~np.isnan(Xn).any(axis=1) & ~np.isnan(y)
    Xn, y = Xn[ok], y[ok]
    fit = fit or engine_fitter(Xn, engine=engine, num_restart=num_restart, random_seed=random_seed, likelihood=lik)
    gen = full_kernel_search_gen(Xn.shape[1], kern_list, cat_vars=cat_vars, max_depth=max_depth, keep_all=keep_all,
                                 metric_diff=metric_diff, early_stopping=early_stopping, prune=prune,
                                 keep_only_best=keep_only_best, softmax_select=softmax_select)
    return _drive_single(gen, y, fit)



def holdout_fitter(fit: Callable, X_train: np.ndarray, X_holdout: np.ndarray, y_holdout: np.ndarray,
                   log_density: Optional[Callable] = None) -> Callable:
    """Wrap a fitter ``fit(requests) -> [(model, bic)]`` (models fitted on the training rows) so that the criterion becomes
    what ``kernel_test(split=True)`` reports (:2299-2308): round(-sum of the predictive log density of the hold-out rows, 2).
    Requests that arrive as BicRequests keep their BIC.  ``log_density(model, X_train, y_train, X_holdout, y_holdout)`` ->
    [m] overrides the engine's ``model.predict_log_density`` (the CPU tests score with the oracle)."""
    if log_density is None:
        def log_density(m, Xt, yt, Xh, yh):
            return m.predict_log_density((Xh, yh), data=(Xt, np.asarray(yt).reshape(-1, 1)))

    def scored(requests, **kw):
        res = fit(requests, **kw)
        if isinstance(requests, BicRequests):
            return res
        out = []
        for (y, _name, _k), (m, bic) in zip(requests, res):
            if m is None:
                out.append((None, np.inf))
                continue
            lp = log_density(m, X_train, np.asarray(y), X_holdout, y_holdout)
            out.append((m, round(-float(np.sum(lp)), 2)))
        return out
    return scored


def split_kernel_search(X, Y, kern_list, unit_idx, training_percent=0.7, cat_vars=(), max_depth=5, keep_all=False,
                        metric_diff=1, early_stopping=True, prune=True, num_restart=5, lik="gaussian", scale_value=None,
                        verbose=False, debug=False, keep_only_best=True, softmax_select=False, random_seed=None,
                        engine=None, fit=None, log_density=None, **unused):
    """Drop-in for :3275-3532: the kernel search of one outcome with the units split into a training part
    (``training_percent`` of the unit ids, drawn with ``np.random.choice`` after ``np.random.seed(random_seed)``) and a
    hold-out part; candidates are fitted on the training rows and ranked by the negated hold-out log density.  Returns the
    reference's dictionary (models, edges, best_model, var_exp, X_holdout, Y_holdout, X, Y).  ``fit`` overrides the engine
    fitter for the TRAINING fits and ``log_density`` the hold-out scorer (the CPU tests pass the oracle for both)."""
    if random_seed is not None:
        np.random.seed(random_seed)
    Xn = X.to_numpy() if hasattr(X, "to_numpy") else np.asarray(X)
    Xn = np.asarray(Xn, dtype=np.float64).reshape(len(Xn), -1)
    Yn = Y.to_numpy() if hasattr(Y, "to_numpy") else np.asarray(Y)
    y = np.asarray(Yn, dtype=np.float64).reshape(-1)
    ok = ~np.isnan(Xn).any(axis=1) & ~np.isnan(y)
    Xn, y = Xn[ok], y[ok]
    unique_ids = np.unique(Xn[:, unit_idx])
    train_ids = np.random.choice(unique_ids, size=round(training_percent * len(unique_ids)), replace=False)
    tr = np.isin(Xn[:, unit_idx], train_ids)
    X_tr, y_tr, X_ho, y_ho = Xn[tr], y[tr], Xn[~tr], y[~tr]
    base = fit or engine_fitter(X_tr, engine=engine, num_restart=num_restart, random_seed=random_seed, likelihood=lik)
    scored = holdout_fitter(base, X_tr, X_ho, y_ho, log_density=log_density)
    gen = full_kernel_search_gen(Xn.shape[1], kern_list, cat_vars=cat_vars, max_depth=max_depth, keep_all=keep_all,
                                 metric_diff=metric_diff, early_stopping=early_stopping, prune=prune,
                                 keep_only_best=keep_only_best, softmax_select=softmax_select, split=True)
    # one outcome: a plain request / reply loop (the BicRequests marker must reach the fitter unflattened)
    try:
        req = next(gen)
        while True:
            tagged = type(req)((y_tr, name, k) for name, k in req)
            req = gen.send(scored(tagged))
    except StopIteration as e:
        out = e.value
    out.update(X_holdout=X_ho, Y_holdout=y_ho.reshape(-1, 1), X=X_tr, Y=y_tr.reshape(-1, 1))
    return out


def softmax_kernel_search(X, Y, kern_list, num_trials=5, cat_vars=(), max_depth=5, lik="gaussian", verbose=False,
                          engine=None, fit=None, **unused):
    """:3570-3627 — ``num_trials`` full searches with the softmax exploration step (``keep_all=True``, no early stopping,
    no pruning), the trial whose best model has the lowest BIC wins.  Returns the reference's 5-tuple (models of the best
    trial, its edges, its best model's name, var_exp, {trial: models}).  (Upstream unpacks ``full_kernel_search``'s
    dictionary as a tuple there and cannot run as written; this is what the code sets out to do.)"""
    best_bic, best = np.inf, ({}, [], "", None)
    search_book = {}
    for i in range(num_trials):
        out = full_kernel_search(X, Y, kern_list, cat_vars=cat_vars, max_depth=max_depth, keep_all=True,
                                 early_stopping=False, prune=False, lik=lik, verbose=verbose, keep_only_best=False,
                                 softmax_select=True, engine=engine, fit=fit)
        search_book[i] = out["models"]
        bic = out["models"][out["best_model"]]["bic"]
        if verbose:
            print(out["best_model"], bic)
        if bic < best_bic:
            best_bic, best = bic, (out["models"], out["edges"], out["best_model"], out["var_exp"])
    return best + (search_book,)
