"""Host logic of the compositional kernel search (waveome/model_search.py:2239-3272) — candidate generation, naming,
de-duplication, windows, pruning — checked against hand-derived expectations from the reference's rules, and one full
search driven by the CPU oracle."""
import numpy as np

import helpers
import waveome_b200 as wb
from waveome_b200 import kernel_search as ks
from oracle_fitter import oracle_fitter

KL = lambda: [wb.SquaredExponential(), wb.Matern12(), wb.Lin(), wb.Periodic(wb.SquaredExponential())]


def _request(cands):
    return cands


def test_first_level_candidates():
    cands = _request(ks.loc_candidates(3, KL(), cat_vars=[0, 2], depth=1))
    assert [c[0] for c in cands] == ["categorical[0]", "constant", "squared_exponential[1]", "matern12[1]", "lin[1]",
                                     "periodic[1]", "categorical[2]"]
    const = dict(cands)["constant"]
    assert float(const.variance) == 1e-6 and not const.variance.trainable     # frozen "empty" candidate (:2385-2389)


def test_sum_candidates_order_skip_and_dedup():
    base = wb.SquaredExponential(active_dims=[1], lengthscales=0.3, variance=2.0)
    cands = _request(ks.loc_candidates(3, KL(), base_kern=base, base_name="squared_exponential[1]", cat_vars=[0, 2],
                                          depth=2, operation="sum", prev_models=["categorical[0]+squared_exponential[1]"]))
    names = [c[0] for c in cands]
    # categorical[0]+SE[1] already exists (any term order); names are joined in string order (:2415-2420)
    assert names == ["squared_exponential[1]+squared_exponential[1]", "matern12[1]+squared_exponential[1]",
                     "lin[1]+squared_exponential[1]", "periodic[1]+squared_exponential[1]",
                     "categorical[2]+squared_exponential[1]"]
    k = dict(cands)["lin[1]+squared_exponential[1]"]
    assert [x.name for x in k.kernels] == ["lin", "squared_exponential"]
    assert float(k.kernels[1].lengthscales) == 1.0 and float(k.kernels[1].variance) == 1.0    # base reset to 1 (:2405)
    # a categorical feature already in the base is never added again (:2410)
    cands = _request(ks.loc_candidates(3, KL(), base_kern=wb.Categorical(active_dims=[0]), base_name="categorical[0]",
                                          cat_vars=[0, 2], depth=2, operation="sum", prev_models=[]))
    assert all("categorical[0]+categorical[0]" != c[0] for c in cands)


def test_product_candidates_freeze_new_factor():
    base = wb.SquaredExponential(active_dims=[1])
    cands = _request(ks.loc_candidates(3, KL(), base_kern=base, base_name="squared_exponential[1]", cat_vars=[0, 2],
                                          depth=2, operation="product", prev_models=[]))
    d = dict(cands)
    assert "categorical[0]*squared_exponential[1]" in d and "periodic[1]*squared_exponential[1]" in d
    k = d["categorical[0]*squared_exponential[1]"]
    assert [x.name for x in k.kernels] == ["categorical", "squared_exponential"]
    assert not k.kernels[0].variance.trainable and k.kernels[1].variance.trainable            # :2464-2468
    kp = d["periodic[1]*squared_exponential[1]"]
    assert not kp.kernels[0].base_kernel.variance.trainable
    # products of products are not built (:2460)
    cands = _request(ks.loc_candidates(3, KL(), base_kern=k, base_name="categorical[0]*squared_exponential[1]",
                                          cat_vars=[0, 2], depth=3, operation="product", prev_models=[]))
    assert cands == []


def test_split_product_name_moves_but_kernel_stays():
    base = wb.Sum([wb.Lin(active_dims=[1]), wb.SquaredExponential(active_dims=[1])])
    new = wb.Categorical(active_dims=[0])
    cands = ks.prod_kernel_candidates(base, "lin[1]+squared_exponential[1]", new, [])
    names = [c[0] for c in cands]
    # "categorical[0]" < "lin[1]": the product is named categorical[0]*lin[1] and its NAME is re-inserted before the
    # first name that sorts after "categorical[0]" (:2607-2624); the kernel list keeps its positions
    assert names == ["categorical[0]*lin[1]+squared_exponential[1]", "categorical[0]*squared_exponential[1]+lin[1]"]
    k2 = cands[1][1]
    assert [x.name for x in k2.kernels] == ["lin", "product"]
    # a component that already holds the categorical of the new factor, or is a product, is skipped (:2591-2596)
    base = wb.Sum([wb.Categorical(active_dims=[0]), wb.Product([wb.Categorical(active_dims=[2]), wb.Lin(active_dims=[1])])])
    assert ks.prod_kernel_candidates(base, "categorical[0]+categorical[2]*lin[1]", new, []) == []


def test_window_and_better_metric():
    d = {"a": dict(bic=10.0, depth=1, try_next=True), "b": dict(bic=15.9, depth=1, try_next=True),
         "c": dict(bic=16.01, depth=1, try_next=True), "a+b": dict(bic=9.5, depth=2, try_next=True)}
    ks.keep_top_k(d, depth=1, metric_diff=6)
    assert d["a"]["try_next"] and d["b"]["try_next"] and not d["c"]["try_next"]
    assert ks.check_if_better_metric(d, 2) and not ks.check_if_better_metric(d, 3)


def test_prune_requests():
    m = ks.candidate_model(wb.Sum([wb.Categorical(active_dims=[0]),
                                   wb.Product([wb.Categorical(active_dims=[2]), wb.SquaredExponential(active_dims=[1])])]))
    res = {"categorical[0]+categorical[2]*squared_exponential[1]": dict(kernel=m.kernel, model=m, bic=5.0, depth=2,
                                                                        parent="x", try_next=True),
           "categorical[0]": dict(kernel=None, model=None, bic=9.0, depth=1, parent="None", try_next=True)}
    gen = ks.prune_best_model2(res, depth=2)
    req = next(gen)
    # dropping categorical[0] leaves the product; the product's factors are tried next to the other component;
    # "categorical[0]+categorical[2]" etc. are new, nothing collides with the existing "categorical[0]"
    assert [r[0] for r in req] == ["categorical[2]*squared_exponential[1]", "categorical[0]+categorical[2]",
                                   "categorical[0]+squared_exponential[1]"]
    fitted = [(ks.candidate_model(k), b) for (n, k), b in zip(req, [4.0, 6.0, 4.5])]
    try:
        gen.send(fitted)
    except StopIteration as e:
        out = e.value
    assert set(out) == set(res) | {"categorical[2]*squared_exponential[1]", "categorical[0]+squared_exponential[1]"}
    assert out["categorical[0]+squared_exponential[1]"]["depth"] == 2


def _toy():
    rng = np.random.default_rng(5)
    n = 60
    subj = np.repeat(np.arange(12), 5).astype(float)
    t = rng.normal(size=n)
    grp = rng.integers(0, 2, size=n).astype(float)
    X = np.stack([subj, t, grp], 1)
    y = 1.5 * np.sin(2.0 * t) + rng.normal(size=12)[subj.astype(int)] + 0.1 * rng.normal(size=n)
    return X, y


def test_full_search_with_oracle_fits():
    """y = smooth f(t) + subject offset + noise: the search must pick up both components; every rule of the depth loop
    is visible in the result dictionary."""
    X, y = _toy()
    out = ks.full_kernel_search(X, y, [wb.SquaredExponential(), wb.Lin()], cat_vars=[0, 2], max_depth=3,
                                fit=oracle_fitter(X), keep_only_best=False)
    models = out["models"]
    d1 = {k: v for k, v in models.items() if v["depth"] == 1}
    assert set(d1) == {"categorical[0]", "constant", "squared_exponential[1]", "lin[1]", "categorical[2]"}
    best2 = min((v["bic"], k) for k, v in models.items() if v["depth"] == 2)[1]
    assert best2 == "categorical[0]+squared_exponential[1]"
    best = out["best_model"]
    assert "categorical[0]" in best and "squared_exponential[1]" in best
    assert models[best]["bic"] == min(v["bic"] for v in models.values())
    # window: first-level entries further than 6 from the level's best were not expanded (:2705-2707)
    b1 = min(v["bic"] for v in d1.values())
    for k, v in d1.items():
        assert v["try_next"] == (v["bic"] - b1 <= 6)
        if not v["try_next"]:
            assert not any(p == k for p, _c in out["edges"])
    # BIC = round(2 k - 2 log p, 2) with k = trainable Parameter objects (kernel + noise + mean)
    m = models[best2]["model"]
    assert len(m.trainable_parameters) == 5
    assert models[best2]["bic"] == round(2 * 5 - 2 * m.log_posterior_density_value, 2)


def test_pipelined_lockstep_gives_the_same_search():
    """run_lockstep with outcome groups (one group's batch on the fitter thread while the host advances the other):
    same candidates, same results, same selected structures as the single-group driver."""
    import zlib
    from waveome_b200 import datasets
    from waveome_b200.model_search import GPSearch

    calls = []

    def fake_fit(requests):
        calls.append(len(requests))
        out = []
        for y, name, kernel in requests:
            m = ks.candidate_model(kernel)
            h = zlib.crc32((name + repr(float(y[0]))).encode()) % 10000 / 50.0
            out.append((m, round(300.0 - 12.0 * min(name.count("+") + name.count("*"), 2) + h, 2)))
        return out

    X, Y = datasets.overview_synthetic(n_people=6, n_observations=4, n_outcomes=40)
    res = {}
    for groups in (1, 2, 3, None):
        calls.clear()
        gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
        gps.run_search(max_depth=4, fit=fake_fit, pipeline_groups=groups)
        res[groups] = ({o: gps.search_info[o]["best_model"] for o in gps.out_names},
                       {o: sorted((k, v["bic"], v["depth"], v["try_next"]) for k, v in gps.search_info[o]["models"].items())
                        for o in gps.out_names}, gps.fit_report["n_fits"], sum(calls))
    for groups in (2, 3, None):
        assert res[groups] == res[1], groups
    assert len(set(res[1][0].values())) > 3          # the fake criterion does spread the outcomes over structures


def test_fitter_threads_get_their_own_engines(monkeypatch):
    """run_lockstep(groups > 1) with WV_SEARCH_FITTERS fitter threads: every fitter thread resolves to its own engine of
    the device pool (never the process engine, never a shared one), and the pool does not grow from search to search."""
    import threading
    import time
    from waveome_b200 import model_fitting as mf
    monkeypatch.setenv("WV_SEARCH_FITTERS", "2")
    monkeypatch.setattr(mf, "get_engine", lambda device=None: "engine0")
    monkeypatch.setattr(mf, "get_engine_pool", lambda k, device=None: ["engine%d" % i for i in range(k)])
    used = []

    def fit(requests):
        used.append((threading.get_ident(), ks._thread_engine()))
        time.sleep(0.02)
        return [(None, np.inf)] * len(requests)

    def gen():
        for _ in range(3):
            yield [("a", None)]
        return {}

    assert ks._thread_engine() == "engine0"                     # the calling thread keeps the process engine
    for _ in range(2):                                          # (thread ids may be reused from one search to the next)
        used.clear()
        ks.run_lockstep({o: gen() for o in "abcd"}, {o: np.zeros(3) for o in "abcd"}, fit, groups=4)
        by_thread = {}
        for tid, eng in used:
            by_thread.setdefault(tid, set()).add(eng)
        assert len(used) == 12 and all(len(v) == 1 for v in by_thread.values())
        assert {e for _t, e in used} == {"engine1", "engine2"}
    assert ks._thread_engine() == "engine0"


def test_deferred_stragglers_give_the_same_search():
    """run_lockstep with a fitter that hands back control while some requests are unfinished (the engine fitter with
    ``tail`` > 0): outcomes without stragglers move on, the others rejoin later -- same candidates, same results, same
    selected structures as the level-synchronous driver, every request fitted exactly once."""
    import threading
    import time
    import zlib
    from waveome_b200 import datasets
    from waveome_b200.model_search import GPSearch

    def result(y, name, kernel):
        m = ks.candidate_model(kernel)
        h = zlib.crc32((name + repr(float(y[0]))).encode()) % 10000 / 50.0
        return (m, round(300.0 - 12.0 * min(name.count("+") + name.count("*"), 2) + h, 2))

    def sync_fit(requests):
        return [result(*r) for r in requests]

    seen, threads, shapes = [], set(), []

    def deferring_fit(requests, tail=0):
        assert tail > 0
        seen.extend((float(y[0]), name) for y, name, _k in requests)
        full = [result(*r) for r in requests]
        # every 7th request (by a hash of its identity) is a straggler
        slow = [zlib.crc32((name + repr(float(y[0]))).encode()) % 7 == 0 for y, name, _k in requests]
        shapes.append((len(requests), sum(slow)))
        if not any(slow):
            return full, None

        def pending():
            threads.add(threading.get_ident())
            time.sleep(0.01)
            return full

        return [None if s else r for r, s in zip(full, slow)], pending

    deferring_fit.supports_tail = True
    X, Y = datasets.overview_synthetic(n_people=6, n_observations=4, n_outcomes=40)
    res = {}
    for label, fit in (("sync", sync_fit), ("deferred", deferring_fit)):
        gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
        gps.run_search(max_depth=4, fit=fit)
        res[label] = ({o: gps.search_info[o]["best_model"] for o in gps.out_names},
                      {o: sorted((k, v["bic"], v["depth"], v["try_next"]) for k, v in gps.search_info[o]["models"].items())
                       for o in gps.out_names}, gps.fit_report["n_fits"])
    assert res["deferred"] == res["sync"]
    assert len(seen) == len(set(seen)) == res["sync"][2]            # nothing fitted twice, nothing dropped
    assert threads and threading.get_ident() not in threads         # the stragglers finished on worker threads
    assert sum(1 for _n, k in shapes if k) > 3                      # ... at several levels


def test_engine_lease_hands_out_distinct_engines(monkeypatch):
    from waveome_b200 import engine as E, model_fitting as mf
    made = []

    class FakeEngine:
        def __init__(self, device, high_priority=False):
            assert high_priority
            made.append(self)

    monkeypatch.setattr(E, "Engine", FakeEngine)
    monkeypatch.setattr(mf, "_LEASED", {})
    monkeypatch.setattr(mf, "_HP_POOLS", {})
    a, ia = mf.lease_engine(0)
    b, ib = mf.lease_engine(0)
    assert a is not b and len(made) == 2
    mf.release_engine(ia, 0)
    c, ic = mf.lease_engine(0)
    assert c is a and ic == ia and len(made) == 2          # returned engines are reused, the pool does not grow
    mf.release_engine(ib, 0)
    mf.release_engine(ic, 0)
    assert mf._LEASED[0] == set()


def test_softmax_kernel_selection_rule():
    """waveome/model_search.py:3535-3567: inf criteria are dropped, one survivor is returned as it is, otherwise one draw
    from softmax((-bic - min) / (max - min)) through np.random."""
    assert ks.softmax_kernel_selection([np.inf, 4.0], ["a", "b"]) == "b"
    bics, names = [10.0, 12.0, np.inf, 7.0], ["a", "b", "c", "d"]
    neg = np.array([-10.0, -12.0, -7.0])
    p = np.exp((neg - neg.min()) / (neg.max() - neg.min()))
    p /= p.sum()
    np.random.seed(5)
    want = [["a", "b", "d"][np.random.choice(np.arange(3), p=p)] for _ in range(20)]
    np.random.seed(5)
    got = [ks.softmax_kernel_selection(bics, names) for _ in range(20)]
    assert got == want and set(got) <= {"a", "b", "d"} and len(set(got)) > 1


def test_split_kernel_search_with_oracle():
    """split_kernel_search (:3275-3532) driven by the CPU oracle: units never straddle the split, candidates are ranked by
    the negated hold-out log density of fits on the training rows, pruning refits are scored by BIC (the reference prunes
    without the hold-out arguments), and the reference's dictionary comes back."""
    import gp_oracle as oracle
    import waveome_b200 as wb
    from oracle_fitter import oracle_fitter
    X, y = _toy()
    seen = {"bic": 0, "holdout": 0}

    def make_fit(Xt):
        inner = oracle_fitter(Xt)

        def fit(requests, **kw):
            seen["bic" if isinstance(requests, ks.BicRequests) else "holdout"] += len(requests)
            return inner(requests)
        return fit

    def log_density(m, Xt, yt, Xh, yh):
        return oracle.predict_log_density(m.to_spec(), Xt, yt, Xh, yh)

    np.random.seed(11)
    ids = np.unique(X[:, 0])
    train_ids = np.random.choice(ids, size=round(0.7 * len(ids)), replace=False)       # what the search must draw
    tr = np.isin(X[:, 0], train_ids)
    out = ks.split_kernel_search(X, y, [wb.SquaredExponential(), wb.Lin()], unit_idx=0, cat_vars=[0, 2], max_depth=2,
                                 random_seed=11, fit=make_fit(X[tr]), log_density=log_density, keep_only_best=False)
    assert set(out) >= {"models", "edges", "best_model", "var_exp", "X_holdout", "Y_holdout", "X", "Y"}
    np.testing.assert_array_equal(out["X"], X[tr])
    np.testing.assert_array_equal(out["X_holdout"], X[~tr])
    assert not set(out["X"][:, 0]) & set(out["X_holdout"][:, 0])
    assert seen["holdout"] > 5 and seen["bic"] >= 0
    best = out["models"][out["best_model"]]
    searched = {k: v for k, v in out["models"].items() if v["model"] is not None and v["parent"] != "prune"}
    # every searched candidate's criterion is the rounded negated hold-out log density at its fitted parameters
    checked = 0
    for k, v in list(searched.items())[:6]:
        lp = oracle.predict_log_density(v["model"].to_spec(), X[tr], y[tr], X[~tr], y[~tr])
        if abs(v["bic"] - round(-float(np.sum(lp)), 2)) <= 0.011:
            checked += 1
    assert checked >= 4                       # (pruning refits carry BICs instead: not all entries are hold-out scores)
    assert best["bic"] == min(v["bic"] for v in out["models"].values())


def test_softmax_kernel_search_keeps_the_best_trial():
    """softmax_kernel_search (:3570-3627): every trial expands ONE softmax-drawn model per depth; the best trial wins."""
    import zlib
    import waveome_b200 as wb
    X, y = _toy()
    calls = []

    def fake_fit(requests, **kw):
        calls.append(len(requests))
        return [(ks.candidate_model(k), round(100.0 - 5.0 * (name.count("+") + name.count("*")) + 50.0 * (name == "constant")
                                             + zlib.crc32(name.encode()) % 100 / 25.0, 2)) for _y, name, k in requests]

    np.random.seed(3)
    models, edges, best, var_exp, book = ks.softmax_kernel_search(X, y, [wb.SquaredExponential(), wb.Lin()], num_trials=3,
                                                                 cat_vars=[0, 2], max_depth=3, fit=fake_fit)
    assert len(book) == 3 and best in models and models is book[[i for i in book if book[i] is models][0]]
    assert models[best]["bic"] == min(min(v["bic"] for v in b.values()) for b in book.values())
    for trial in book.values():
        for d in (1, 2):                                   # one expandable model per depth below the last
            assert sum(1 for v in trial.values() if v["depth"] == d and v["try_next"] is not False) <= 1
