"""CPU-only checks of the boundary and the host logic: the C-ABI library loads and exports every symbol the header
declares, program encoding, kernel naming / pruning arithmetic, data preparation, sharding."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

import waveome_b200 as wb
from waveome_b200 import datasets, program, regularization, utilities
from waveome_b200.model_search import GPSearch, shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from waveome_b200 import engine
    hdr = open(os.path.join(ROOT, "include", "waveome_b200.h")).read()
    declared = set(re.findall(r"\b(wv_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(engine.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(os.path.join(ROOT, "waveome_b200", "_lib", "libwaveome_b200.so"))
    for s in declared:
        assert hasattr(lib, s), s
    lib.wv_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.wv_version()


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from waveome_b200.engine import Engine, EngineError
    with pytest.raises(EngineError, match="no CPU fallback"):
        Engine(0)


def test_program_encoding_saturated_kernel():
    k, names = regularization.full_kernel_build(cat_vars=[0, 3, 4], num_vars=[1, 2], unit_idx=0,
                                                var_names=["participant", "age", "study_day", "sex", "site"],
                                                return_sum=True)
    assert names == ["categorical[participant]", "categorical[sex]", "categorical[site]",
                     "squared_exponential[age]", "squared_exponential[study_day]",
                     "categorical[sex]*squared_exponential[age]", "categorical[sex]*squared_exponential[study_day]",
                     "categorical[site]*squared_exponential[age]", "categorical[site]*squared_exponential[study_day]"]
    m = wb.models.PenalizedGPR(k, mean_function=wb.ConstantMean(), penalization_factor=2.0)
    p = m.program()
    assert (p.n_comp, p.n_leaves, p.n_x) == (9, 13, 17)
    # frozen categorical variances inside products carry no prior and no packed index (regularization.py:131-132)
    frozen = [s for s in range(p.n_slots) if p.slot_xindex[s] < 0]
    assert len(frozen) == 4 and all(p.slot_prior[s] == 0 for s in frozen)
    hs = [s for s in range(p.n_slots) if p.slot_prior[s] == program.PRIOR_CODE["horseshoe"]]
    assert len(hs) == 9 and np.allclose(p.slot_pa[hs], 0.5)          # scale = 1 / penalization_factor
    assert p.slot_transform[p.noise_slot] == program.TRANSFORM_CODE["softplus_shift"] and p.slot_shift[p.noise_slot] == 1e-6
    assert len(m.trainable_parameters) == 17
    x0 = p.x0()
    p.assign(x0 + 0.5)
    np.testing.assert_allclose(p.x0(), x0 + 0.5, rtol=1e-12)


def test_product_of_sum_is_distributed():
    a, b, c = wb.SquaredExponential(active_dims=[0]), wb.Lin(active_dims=[1]), wb.Categorical(active_dims=[2])
    comps = program.expand_sum_of_products(wb.Product([wb.Sum([a, b]), c]))
    assert [[l.name for l in comp] for comp in comps] == [["squared_exponential", "categorical"], ["lin", "categorical"]]
    p = wb.GPR(wb.Product([wb.Sum([a, b]), c])).program()
    # the shared categorical variance is one slot (the encoder moves transcendental-free leaves to the front)
    cat = [l for l in range(p.n_leaves) if p.leaf_type[l] == program.LEAF_CODE["categorical"]]
    assert len(cat) == 2 and p.leaf_s_var[cat[0]] == p.leaf_s_var[cat[1]]


def test_names_dedup_and_pruning():
    k = wb.Sum([wb.Categorical(active_dims=[0]), wb.Product([wb.Categorical(active_dims=[2]), wb.SquaredExponential(active_dims=[1])])])
    assert utilities.kernel_name_string(k) == "categorical[0]+categorical[2]*squared_exponential[1]"
    assert utilities.check_if_model_exists("squared_exponential[1]*categorical[2]+categorical[0]", [utilities.kernel_name_string(k)])
    assert not utilities.check_if_model_exists("categorical[0]", [utilities.kernel_name_string(k)])
    X = np.random.default_rng(0).normal(size=(30, 3))
    m = wb.models.PenalizedGPR(k)
    m.kernel.kernels[0].variance.assign(0.05)
    m.cut_kernel_components(X)
    assert m.kernel.name == "product"
    m2 = wb.models.PenalizedGPR(wb.Sum([wb.SquaredExponential(active_dims=[1], lengthscales=100.0), wb.Lin(active_dims=[0])]))
    m2.cut_kernel_components(X)                        # lengthscale >= 3 * range -> dropped (utilities.py:1150-1153)
    assert m2.kernel.name == "lin"
    m3 = wb.models.PenalizedGPR(wb.Sum([wb.Lin(active_dims=[0], variance=0.01), wb.Lin(active_dims=[1], variance=0.02)]))
    m3.cut_kernel_components(X)
    assert m3.kernel.name == "constant"
    assert utilities.calc_bic(-12.5, 100, 6) == 37.0


def test_gpsearch_data_preparation():
    X, Y = datasets.iris()
    gps = GPSearch(X, Y, categorical_vars=["species"])
    assert gps.cat_idx == [2] and gps.cont_idx == [0, 1] and gps.unit_idx is None
    np.testing.assert_allclose(gps.X["petal_length"].std(), 1.0)           # pandas std, ddof = 1
    assert sorted(gps.X["species"].unique()) == [0.0, 1.0, 2.0]
    Xo, Yo = datasets.overview_notebook()
    # soft pins recorded in waveome_overview.ipynb: cell 4 head() and cell 11 Z[0] = [0., -1.44041, 0.]
    np.testing.assert_allclose(Xo["time"].iloc[:2].to_numpy(), [1.175864, 1.843311], atol=5e-7)
    np.testing.assert_allclose(Yo.iloc[0].to_numpy(), [0.889715, 0.381912, 1.314904], atol=5e-7)
    g2 = GPSearch(Xo, Yo, unit_col="person_id", categorical_vars=["female"])
    np.testing.assert_allclose(g2.X.iloc[0].to_numpy(), [0.0, -1.44041, 0.0], atol=5e-6)
    with pytest.raises(TypeError):
        GPSearch(X.to_numpy(), Y)
    # reverse_transform (model_search.py:1677-1715) undoes the standardisation; run_penalized_search is deprecated upstream
    np.testing.assert_allclose(g2.reverse_transform(g2.X["time"].iloc[:2], feature_name="time", round_digits=6),
                               [1.175864, 1.843311], atol=1e-6)
    g3 = GPSearch(Xo, Yo, unit_col="person_id", categorical_vars=["female"], Y_transform="standardize")
    np.testing.assert_allclose(g3.reverse_transform(g3.Y.iloc[0], input_type="Y", round_digits=6),
                               [0.889715, 0.381912, 1.314904], atol=1e-6)
    with pytest.raises(NotImplementedError):
        g2.run_penalized_search()


def test_shard_bounds_cover_everything():
    for n in (1, 7, 250, 2000):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_models_survive_pickle_and_deepcopy():
    """fitted models travel between processes (Ray in the reference, all_gather_object here) and are
    deep-copied by callers; every likelihood of the crosswalk must survive both with the same device program."""
    import copy
    import pickle
    import waveome_b200 as wb
    from waveome_b200 import kernels as K
    from waveome_b200.models import PenalizedGPR, make_likelihood
    k = wb.Sum([wb.Categorical(active_dims=[0]),
                wb.Product([wb.Categorical(active_dims=[2]), wb.SquaredExponential(active_dims=[1])]),
                wb.Periodic(wb.SquaredExponential(active_dims=[1]))])
    for lik, n_x in (("gaussian", 9), ("poisson", 8), ("negative_binomial", 9), ("bernoulli", 8), ("gamma", 9),
                     ("zeroinflated_negativebinomial", 10)):
        m = PenalizedGPR(k, mean_function=wb.ConstantMean(0.1), likelihood=None if lik == "gaussian" else make_likelihood(lik))
        assert m.program().n_x == n_x
        for m2 in (pickle.loads(pickle.dumps(m)), copy.deepcopy(m)):
            assert m2.kernel_name == m.kernel_name and m2.program().signature() == m.program().signature()
            assert m2.kernel is not m.kernel and m2.likelihood is not m.likelihood
        assert K.deepcopy(m.kernel).to_spec() == m.kernel.to_spec()


def test_likelihood_key_ignores_trained_parameters():
    """Fitted models with individual (trainable) dispersions must share one engine batch in the post-fit passes; a frozen
    likelihood parameter is a batch-level value and does enter the key."""
    from waveome_b200.model_fitting import likelihood_key
    from waveome_b200.models import NegativeBinomial, ZeroInflatedNegativeBinomial, make_likelihood
    import waveome_b200 as wb
    k = wb.SquaredExponential(active_dims=[0])
    a = wb.GPR(k, likelihood=NegativeBinomial(alpha=0.3))
    b = wb.GPR(k, likelihood=NegativeBinomial(alpha=2.5))
    assert likelihood_key(a) == likelihood_key(b) == ("negative_binomial", 1.0)
    c = wb.GPR(k, likelihood=NegativeBinomial(alpha=2.5, trainable=False))
    assert likelihood_key(c) == ("negative_binomial", 2.5)
    z1 = wb.GPR(k, likelihood=ZeroInflatedNegativeBinomial(alpha=0.3, km=2.0))
    z2 = wb.GPR(k, likelihood=ZeroInflatedNegativeBinomial(alpha=0.9, km=0.5))
    assert likelihood_key(z1) == likelihood_key(z2)
    assert likelihood_key(wb.GPR(k, likelihood=make_likelihood("poisson"))) == ("poisson", 0.0)
    assert likelihood_key(wb.GPR(k)) == ("gaussian", 0.0)


def test_split_for_streams():
    """Sub-batches of one fit (model_fitting.split_for_streams): contiguous cover, at most `chunk` models in flight over
    all streams, no split below MIN_MODELS_PER_STREAM models per piece."""
    from waveome_b200.model_fitting import MIN_MODELS_PER_STREAM, split_for_streams
    for n, chunk, streams in [(2000, 9000, 4), (2000, 900, 4), (255, 9000, 4), (256, 9000, 4), (1, 5, 4), (1000, 9000, 1),
                              (10254, 9000, 4), (513, 100, 8)]:
        pieces = split_for_streams(n, chunk, streams)
        assert pieces[0][0] == 0 and pieces[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
        sizes = [hi - lo for lo, hi in pieces]
        k = max(1, min(streams, n // MIN_MODELS_PER_STREAM))
        assert max(sizes) * k <= max(chunk, k) or max(sizes) == 1          # k concurrent pieces stay within the chunk
        if n < 2 * MIN_MODELS_PER_STREAM and n <= chunk:
            assert pieces == [(0, n)]
    assert split_for_streams(2000, 9000, 4) == [(0, 500), (500, 1000), (1000, 1500), (1500, 2000)]


def test_fit_models_reassembles_concurrent_sub_batches(monkeypatch):
    """Host logic of the concurrent sub-batches without a GPU: ``engine.Batch`` and the engine pool are replaced by
    fakes; fit_models(streams=4) must hand every model its own row of results whatever piece and thread fitted it,
    group the models by likelihood, and re-raise a worker's exception on the calling thread."""
    import threading
    import time
    import numpy as np
    import waveome_b200 as wb
    from waveome_b200 import engine as E, model_fitting as mf
    from waveome_b200.models import make_likelihood

    seen = []
    solo_flags = []

    class FakeBatch:
        def __init__(self, eng, X, Y, table, prog_id=None, P=None, specialize=False):
            self.eng, self.Y, self.P, self.lik = eng, np.asarray(Y), P, ("gaussian", 0.0)
            self.B = self.Y.shape[0]

        def set_likelihood(self, name, param):
            self.lik = (name, param)

        def set_solo(self, solo=True):
            solo_flags.append(bool(solo))

        def fit(self, x0=None, **opts):
            if self.Y[0, 0] < 0:
                raise RuntimeError("boom")
            seen.append((self.eng, threading.get_ident(), self.B, self.lik[0]))
            time.sleep(0.05)                                          # a real fit leaves the other workers time to start
            tag = self.Y[:, 0]                                        # every model is recognisable by its first y value
            return dict(x=np.asarray(x0) + tag[:, None], f=-tag, lml=tag.copy(), n_iter=tag.astype(np.int32),
                        n_eval=(2 * tag).astype(np.int32), status=np.zeros(self.B, np.int32))

        def counters(self):
            return dict(launches=10, rounds=1, model_evals=self.B)

        def close(self):
            pass

    monkeypatch.setattr(E, "Batch", FakeBatch)
    monkeypatch.setattr(mf, "get_engine", lambda device=None: "engine0")
    monkeypatch.setattr(mf, "get_engine_pool", lambda k, device=None: ["engine%d" % i for i in range(k)])
    B, n = 700, 12
    X = np.zeros((n, 2))
    Y = np.zeros((B, n))
    Y[:, 0] = np.arange(B)
    models = []
    for b in range(B):
        k = wb.SquaredExponential(active_dims=[1]) if b % 3 else wb.Sum([wb.SquaredExponential(active_dims=[1]), wb.Lin(active_dims=[1])])
        lik = make_likelihood("poisson") if b >= 600 else None
        models.append(wb.GPR(k, mean_function=wb.ConstantMean(0.0), likelihood=lik))
    res = mf.fit_models(X, Y, models, streams=4)
    np.testing.assert_array_equal(res["lml"], np.arange(B))
    np.testing.assert_array_equal(res["n_eval"], 2 * np.arange(B))
    for b in (0, 1, 299, 650, 699):
        assert models[b].log_marginal_likelihood_value == b and models[b].fit_info["n_iter"] == b
    # 600 Gaussian models -> 4 pieces of 150 on the pool; 100 Poisson models -> one piece (too few to split)
    assert sorted(s[2] for s in seen) == [100, 150, 150, 150, 150]
    assert {s[3] for s in seen if s[2] == 100} == {"poisson"} and res["rounds"] == 5 and res["launches"] == 50
    assert len({s[0] for s in seen}) > 1 or len({s[1] for s in seen}) > 1          # more than one engine / thread took part
    assert solo_flags == []                                            # concurrent pieces never declare themselves alone
    # default: one stream, everything on the process engine -- and the batch says it has the device to itself
    seen.clear()
    mf.fit_models(X, Y[:600], models[:600])
    assert [s[2] for s in seen] == [600] and seen[0][0] == "engine0" and solo_flags == [True]
    # a failing piece surfaces as the caller's exception
    Y[300, 0] = -1.0
    with pytest.raises(RuntimeError, match="boom"):
        mf.fit_models(X, Y, models, streams=4)


def test_program_signature_ignores_trainable_values():
    """Models of one structure share one device program whatever their current (trainable) values -- the values travel
    in x --, while frozen values, priors and structure are part of the identity."""
    import helpers
    import waveome_b200 as wb
    a = wb.GPR(helpers.saturated_kernel(), mean_function=wb.ConstantMean(0.1))
    b = wb.GPR(helpers.saturated_kernel(), mean_function=wb.ConstantMean(0.3))
    b.kernel.kernels[2].lengthscales.assign(0.3)
    b.likelihood.variance.assign(0.2)
    assert a.program().signature() == b.program().signature()
    assert not np.array_equal(a.program().x0(), b.program().x0())
    wb.set_trainable(b.kernel.kernels[2].lengthscales, False)                    # frozen at 0.3
    assert a.program().signature() != b.program().signature()
    c = wb.GPR(helpers.saturated_kernel(), mean_function=wb.ConstantMean(0.1))
    wb.set_trainable(c.kernel.kernels[2].lengthscales, False)
    c.kernel.kernels[2].lengthscales.assign(0.3)
    assert b.program().signature() == c.program().signature()
    c.kernel.kernels[2].lengthscales.assign(0.4)                                 # another frozen value
    assert b.program().signature() != c.program().signature()
    d = wb.GPR(helpers.saturated_kernel(hs=2.0), mean_function=wb.ConstantMean(0.1))   # another prior scale
    assert a.program().signature() != d.program().signature()


def test_fit_models_deferred_tail_bookkeeping(monkeypatch):
    """fit_models(tail=...): the finished models are written back when control returns, the batch moves to a leased
    high-priority engine, ``pending()`` completes arrays and write-back, closes the batch and returns the lease; small
    batches and multi-job fits are not deferred.  (Host logic on a fake batch; the device side is tests/test_fit_gpu.py.)"""
    import waveome_b200 as wb
    from waveome_b200 import engine as E, model_fitting as mf
    log = []

    class FakeEngine:
        def __init__(self, device=0, high_priority=False):
            self.device, self.high_priority = device, high_priority

    class FakeBatch:
        def __init__(self, eng, X, Y, table, prog_id=None, P=None, specialize=False):
            self.engine, self.B, self.P, self.tag = eng, np.asarray(Y).shape[0], P, np.asarray(Y)[:, 0].copy()
            self.slow = self.tag % 10 == 0                     # every tenth model is a straggler
            self.done_all = False

        def fit_begin(self, x0=None, **opts):
            self.x0 = np.asarray(x0).copy()
            log.append(("begin", self.engine.high_priority))

        def fit_run(self, min_active=0):
            if min_active == 0:
                self.done_all = True
                log.append(("finish", self.engine.high_priority))
                return 0
            return int(self.slow.sum())

        def fit_report(self):
            fin = np.ones(self.B, bool) if self.done_all else ~self.slow
            return dict(x=self.x0 + self.tag[:, None], f=np.where(fin, -self.tag, np.nan), lml=np.where(fin, self.tag, np.nan),
                        n_iter=self.tag.astype(np.int32), n_eval=(2 * self.tag).astype(np.int32),
                        status=np.zeros(self.B, np.int32), finished=fin)

        def fit(self, x0=None, **opts):
            log.append(("sync", self.B))
            return dict(x=np.asarray(x0) + self.tag[:, None], f=-self.tag, lml=self.tag.copy(), n_iter=self.tag.astype(np.int32),
                        n_eval=(2 * self.tag).astype(np.int32), status=np.zeros(self.B, np.int32))

        def move_to(self, eng):
            self.engine = eng
            log.append(("move", eng.high_priority))

        def set_solo(self, solo=True):
            log.append(("solo", bool(solo)))

        def counters(self):
            return dict(launches=7, rounds=3, model_evals=self.B)

        def close(self):
            log.append(("close",))

    monkeypatch.setattr(E, "Batch", FakeBatch)
    monkeypatch.setattr(E, "Engine", FakeEngine)
    monkeypatch.setattr(mf, "_LEASED", {})
    monkeypatch.setattr(mf, "_HP_POOLS", {})
    B, n = 200, 12
    X = np.zeros((n, 2))
    Y = np.zeros((B, n))
    Y[:, 0] = np.arange(B)
    models = [wb.GPR(wb.SquaredExponential(active_dims=[1]), mean_function=wb.ConstantMean(0.0)) for _ in range(B)]
    main = FakeEngine()
    res = mf.fit_models(X, Y, models, engine=main, tail=16)
    slow = np.arange(B) % 10 == 0
    assert np.array_equal(res["finished"], ~slow) and callable(res["pending"])
    assert log == [("begin", False), ("move", True)] and mf._LEASED[0] == {0}
    assert models[7].fit_info["n_iter"] == 7 and not getattr(models[10], "fit_info", None)      # stragglers: not written yet
    out = res["pending"]()
    assert out is res and res["finished"].all() and res["pending"] is None
    assert log[2:] == [("finish", True), ("close",)] and mf._LEASED[0] == set()
    np.testing.assert_array_equal(res["lml"], np.arange(B))
    assert models[10].fit_info["n_eval"] == 20 and models[190].log_marginal_likelihood_value == 190.0
    assert res["launches"] == 7 and res["rounds"] == 3
    # a batch that is not several times the tail is fitted in one go
    log.clear()
    small = mf.fit_models(X, Y[:40], models[:40], engine=main, tail=16)
    assert small["pending"] is None and small["finished"].all() and log[:2] == [("solo", True), ("sync", 40)]
