"""Batched model fitting on the engine — the replacement of the reference's per-model hot loop.

* ``fit_models``       packs B models that share X into one ``engine.Batch`` and runs the device L-BFGS-B
                       (replaces one Ray task + one ``gpflow.optimizers.Scipy().minimize`` per model:
                       waveome/model_search.py:250-393, waveome/model_fitting.py:276-281)
* ``kernel_test_reg``  drop-in for waveome/model_fitting.py:16-373 (exact GPR for "gaussian", the VGP branches for
                       "exponential" / "poisson" / "gamma" / "bernoulli"): best-of-restarts MAP fit of one kernel, returns
                       ``(model, bic)``; failure -> ``(None, inf)``.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import kernels as K
from .models import GPR, make_likelihood
from .utilities import calc_bic, print_kernel_names

_ENGINES: Dict[int, "object"] = {}
DEFAULT_LBFGS_KEYS = ("maxcor", "maxiter", "maxfun", "maxls", "ftol", "gtol", "on_chol_fail")


_ENGINE_LOCK = __import__("threading").RLock()      # engines are created lazily, possibly from several fitter threads at once


def get_engine(device: Optional[int] = None):
    """One engine per GPU per process (LOCAL_RANK picks the GPU under torchrun)."""
    import os
    from .engine import Engine
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    with _ENGINE_LOCK:
        if device not in _ENGINES:
            _ENGINES[device] = Engine(device)
        return _ENGINES[device]


_POOLS: Dict[int, list] = {}
#: sub-batches of one fit that run concurrently, one engine (= CUDA stream) and host thread each: the mid-size and tail
#: rounds of one sub-batch's L-BFGS fill the wave-quantisation gaps of the others (+2.7 % on BASELINE configs[2],
#: bit-identical results: models are independent).  WV_FIT_STREAMS=1 restores the single-stream fit.
FIT_STREAMS = max(1, int(__import__("os").environ.get("WV_FIT_STREAMS", "4")))
MIN_MODELS_PER_STREAM = 128


def get_engine_pool(k: int, device: Optional[int] = None) -> list:
    """k engines on one GPU (pool[0] is ``get_engine(device)``); they share the device's buffer cache."""
    import os
    from .engine import Engine
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    with _ENGINE_LOCK:
        if device not in _POOLS:
            _POOLS[device] = [get_engine(device)]
        pool = _POOLS[device]
        while len(pool) < k:
            pool.append(Engine(device))
        return pool[:k]


def split_for_streams(n_models: int, chunk: int, streams: Optional[int] = None) -> List[tuple]:
    """[lo, hi) pieces of a group of ``n_models`` models: at most ``chunk`` models in flight over all streams, pieces of
    at least MIN_MODELS_PER_STREAM models (one piece when the group is too small to be worth splitting)."""
    streams = FIT_STREAMS if streams is None else max(1, int(streams))
    k = max(1, min(streams, n_models // MIN_MODELS_PER_STREAM))
    piece = max(1, min(-(-n_models // k), max(1, chunk // k)))
    return [(lo, min(n_models, lo + piece)) for lo in range(0, n_models, piece)]


def _open_batch(eng, job):
    """The engine batch of one fit job (dict X, Y, table, prog_id, P, lik_name, lik_param, specialize)."""
    from .engine import Batch
    batch = Batch(eng, job["X"], job["Y"], job["table"], job.get("prog_id"), P=job["P"],
                  specialize=job.get("specialize", False))
    try:
        if job.get("solo"):
            batch.set_solo(True)
        if job["lik_name"] != "gaussian":
            batch.set_likelihood(job["lik_name"], job["lik_param"])
    except BaseException:
        batch.close()
        raise
    return batch


# High-priority engines for the stragglers of a deferred fit (fit_models(tail=...)): the batch moves onto one of them when
# control goes back to the caller, so the caller's engine is free again and the stragglers' tiny launches are scheduled
# ahead of whatever batch runs next.  An engine is driven by ONE host thread at a time: leased, then returned.
_LEASE_LOCK = __import__("threading").Lock()
_HP_POOLS: Dict[int, list] = {}
_LEASED: Dict[int, set] = {}


def lease_engine(device: Optional[int] = None):
    """(engine, index): the first high-priority engine of the device nobody has leased (the pool grows on demand)."""
    import os
    from .engine import Engine
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    with _LEASE_LOCK:
        used = _LEASED.setdefault(device, set())
        pool = _HP_POOLS.setdefault(device, [])
        idx = 0
        while idx in used:
            idx += 1
        while len(pool) <= idx:
            pool.append(Engine(device, high_priority=True))
        used.add(idx)
        return pool[idx], idx


def release_engine(idx: int, device: Optional[int] = None):
    import os
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    with _LEASE_LOCK:
        _LEASED.setdefault(device, set()).discard(idx)


def run_fit_jobs(jobs: List[dict], engine=None, streams: Optional[int] = None, on_done: Optional[Callable] = None,
                 **lbfgs_opts) -> List[tuple]:
    """Fit every job (dict X, Y, table, prog_id, P, lik_name, lik_param, starts) as one engine batch; up to ``streams``
    jobs at a time, each on its own engine and host thread (the C call releases the GIL).  Returns [(result dict,
    counters)] in job order.  With a caller-supplied ``engine`` or a single job everything runs on that one engine.
    ``on_done(i, (result, counters))`` is called, on the thread that ran it, as soon as job i has finished."""
    import threading
    from .engine import Batch
    streams = FIT_STREAMS if streams is None else max(1, int(streams))

    def run_one(eng, job):
        batch = _open_batch(eng, job)
        try:
            if job.get("optimizer", "lbfgs") in ("adam", "adam/gradient"):
                # the reference's default optimiser for kernel_test (waveome/model_classes.py:344-462); num_opt_iter
                # arrives as maxiter
                adam = {k: v for k, v in lbfgs_opts.items() if k in ("learning_rate", "decay_rate", "convergence_threshold",
                                                                    "check_every", "decay_every", "beta1", "beta2", "epsilon")}
                if "maxiter" in lbfgs_opts:
                    adam["max_iter"] = lbfgs_opts["maxiter"]
                r = batch.fit_adam(job.get("starts"), **adam)
            else:
                r = batch.fit(job.get("starts"), **{k: v for k, v in lbfgs_opts.items() if k in DEFAULT_LBFGS_KEYS})
            return r, batch.counters()
        finally:
            batch.close()

    if engine is not None or streams == 1 or len(jobs) <= 1:
        eng = engine or get_engine()
        # one batch at a time on this process's device -- unless this is one of several fitter threads of a search
        from .kernel_search import _FITTER_SLOT
        alone = getattr(_FITTER_SLOT, "slot", None) is None
        res = []
        for i, j in enumerate(jobs):
            res.append(run_one(eng, dict(j, solo=alone)))
            if on_done is not None:
                on_done(i, res[-1])
        return res
    engines = get_engine_pool(min(streams, len(jobs)))
    out: List[Optional[tuple]] = [None] * len(jobs)
    err: list = []
    nxt = [0]
    lock = threading.Lock()

    def worker(eng):
        while not err:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= len(jobs):
                return
            try:
                out[i] = run_one(eng, jobs[i])
                if on_done is not None:
                    on_done(i, out[i])
            except BaseException as e:       # re-raised on the calling thread
                err.append(e)

    threads = [threading.Thread(target=worker, args=(e,)) for e in engines]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if err:
        raise err[0]
    return out


def likelihood_key(model) -> tuple:
    """(engine likelihood name, parameter) of a model: ("gaussian", 0.0), ("poisson", 0.0), ("negative_binomial", alpha),
    ("gamma", shape), ("zinb", (alpha, km)).  Models that share a key can share an engine batch.  A TRAINABLE likelihood
    parameter rides in the model's program (noise slot / second likelihood slot) and the batch-level value is ignored,
    so it does not enter the key: fitted models with individual dispersions still form one batch."""
    lik = model.likelihood
    name = getattr(lik, "name", "gaussian")
    if name == "gaussian":
        return ("gaussian", 0.0)
    p = getattr(lik, "engine_param", 0.0)
    if all(q.trainable for q in getattr(lik, "parameters", [])):
        return (name, (1.0, 1.0) if isinstance(p, tuple) else (1.0 if name in ("negative_binomial", "gamma") else 0.0))
    return (name, tuple(float(v) for v in p) if isinstance(p, tuple) else float(p))


def fit_models(X: np.ndarray, Y: np.ndarray, models: Sequence[GPR], x0: Optional[np.ndarray] = None,
               engine=None, max_batch_bytes: float = 60e9, streams: int = 1, specialize: bool = False,
               optimizer: str = "lbfgs", tail: int = 0, **lbfgs_opts) -> dict:
    """MAP-fit ``models[b]`` to outcome ``Y[b]`` (Y is [B, n]); all models share X [n, D].

    Models with identical kernel programs share one device program.  Fitted values are written back into the
    models' Parameter objects; per-model ``fit_info`` / ``log_marginal_likelihood_value`` /
    ``log_posterior_density_value`` are set.  Returns the raw arrays (x, f, lml, n_iter, n_eval, status).

    ``streams``: concurrent sub-batches per group (``run_fit_jobs``).  Default 1: the mixed batches of the kernel search
    (hundreds of structures, a few hundred models per piece) measured 5 % SLOWER on 4 streams (config 2: 24.8 s against
    23.5 s); ``fit_replicated`` -- one structure, thousands of models -- is where the split pays.
    ``specialize``: ask for run-time specialised element-wise kernels (engine.Batch); only pieces whose models share one
    program structure get them.
    ``tail`` > 0 (needs ``engine``; L-BFGS, one batch): return as soon as at most ``tail`` models are still iterating.
    The result then carries ``finished`` [B] (bool; results and write-back are complete for those models) and
    ``pending``: a callable that runs the stragglers to the end (from any ONE thread; the batch has moved to a leased
    high-priority engine, ``engine`` itself is free again), completes the arrays and the write-back and returns the
    same dict.
    Otherwise ``finished`` is all True and ``pending`` None.
    ``optimizer``: "lbfgs" (SciPy-compatible L-BFGS-B, default) or "adam" / "adam/gradient" (the schedule of
    BaseGP.optimize_params, waveome/model_classes.py:344-462: ``Batch.fit_adam``; maxiter = num_opt_iter)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    B = len(models)
    if Y.shape != (B, X.shape[0]):
        raise ValueError("Y must be [len(models), n]")
    progs = [m.program() for m in models]
    uniq, prog_id, table = {}, np.empty(B, np.int32), []
    for b, p in enumerate(progs):
        sig = p.signature()
        if sig not in uniq:
            uniq[sig] = len(table)
            table.append(p)
        prog_id[b] = uniq[sig]
    P = max(1, max(p.n_x for p in progs))
    starts = np.zeros((B, P))
    for b, p in enumerate(progs):
        starts[b, : p.n_x] = p.x0()
    if x0 is not None:
        starts = np.array(x0, dtype=np.float64).reshape(B, P)
    # bound the device workspace: two padded n x n matrices per model in flight
    n = X.shape[0]
    npad = ((n + 1 + 7) // 8 * 8 + 63) // 64 * 64
    per_model = 2 * npad * npad * 8 + npad * 64 * 8
    chunk = max(1, int(max_batch_bytes // per_model))
    out = dict(x=np.empty((B, P)), f=np.empty(B), lml=np.empty(B), n_iter=np.empty(B, np.int32),
               n_eval=np.empty(B, np.int32), status=np.empty(B, np.int32), launches=0, rounds=0)
    # one engine batch holds one likelihood: group the models by (likelihood, parameter), keep the caller's order
    groups: Dict[tuple, list] = {}
    for b, m in enumerate(models):
        groups.setdefault(likelihood_key(m), []).append(b)
    jobs, sels = [], []
    for (lik_name, lik_param), idx in groups.items():
        idx = np.asarray(idx)
        for lo, hi in split_for_streams(len(idx), chunk, 1 if engine is not None else streams):
            sel = idx[lo:hi]
            sels.append(sel)
            jobs.append(dict(X=X, Y=Y[sel], table=table, prog_id=prog_id[sel], P=P, lik_name=lik_name,
                             lik_param=lik_param, starts=starts[sel], specialize=specialize, optimizer=optimizer))
    def write_back(which):
        for b in which:
            m, p = models[b], progs[b]
            p.assign(out["x"][b, : p.n_x])
            m.log_marginal_likelihood_value = float(out["lml"][b])
            m.log_posterior_density_value = float(-out["f"][b])
            m.fit_info = dict(n_iter=int(out["n_iter"][b]), n_eval=int(out["n_eval"][b]), status=int(out["status"][b]))

    out["finished"] = np.ones(B, bool)
    out["pending"] = None
    # (a batch that is not several times larger than `tail` has no bulk to separate from its stragglers: fitted in one go)
    if int(tail) > 0 and B >= 4 * int(tail) and engine is not None and optimizer == "lbfgs" and len(jobs) == 1:
        sel, job = sels[0], jobs[0]
        opts = {k: v for k, v in lbfgs_opts.items() if k in DEFAULT_LBFGS_KEYS}
        batch = _open_batch(engine, job)
        try:
            batch.fit_begin(job.get("starts"), **opts)
            left = batch.fit_run(int(tail))
            r = batch.fit_report()
        except BaseException:
            batch.close()
            raise
        for key in ("x", "f", "lml", "n_iter", "n_eval", "status"):
            out[key][sel] = r[key]
        out["finished"][sel] = r["finished"]
        write_back(sel[r["finished"]])
        if left == 0:
            c = batch.counters()
            batch.close()
            out["launches"] += c["launches"]
            out["rounds"] += c["rounds"]
            return out
        late = sel[~r["finished"]]
        hp, lease = lease_engine(engine.device)
        try:
            batch.move_to(hp)
        except BaseException:
            release_engine(lease, engine.device)
            batch.close()
            raise

        def pending():
            try:
                batch.fit_run(0)
                r2 = batch.fit_report()
                c = batch.counters()
            finally:
                batch.close()
                release_engine(lease, engine.device)
            keep = ~r["finished"]
            for key in ("x", "f", "lml", "n_iter", "n_eval", "status"):
                out[key][late] = r2[key][keep]
            out["launches"] += c["launches"]
            out["rounds"] += c["rounds"]
            write_back(late)
            out["finished"][:] = True
            out["pending"] = None
            return out

        out["pending"] = pending
        return out
    for sel, (r, c) in zip(sels, run_fit_jobs(jobs, engine=engine, streams=streams, **lbfgs_opts)):
        for key in ("x", "f", "lml", "n_iter", "n_eval", "status"):
            out[key][sel] = r[key]
        out["launches"] += c["launches"]
        out["rounds"] += c["rounds"]
    write_back(range(B))
    return out


def packed_parameters(model: GPR):
    """The model's trainable Parameter objects in the engine's packed order (program.x_params of ``model.program()``
    without building the program): kernel leaves depth-first, likelihood variance, mean constant; shared objects once."""
    from .program import iter_leaves
    seen, out = set(), []
    ps = [p for lf in iter_leaves(model.kernel) for p in lf.parameters] + list(model.likelihood.parameters) + \
        list(model.mean_function.parameters)
    for p in ps:
        if id(p) not in seen:
            seen.add(id(p))
            if p.trainable:
                out.append(p)
    return out


def fit_replicated(X: np.ndarray, Y: np.ndarray, template: GPR, make_models=None, engine=None,
                   max_batch_bytes: float = 60e9, specialize: bool = False, post: Optional[Callable] = None,
                   **lbfgs_opts):
    """MAP-fit B copies of ONE model structure (same kernel tree, priors and start values) to the B outcomes Y[b]:
    what GPSearch.penalized_optimization does (waveome/model_search.py:302-329 builds the same PSVGP for every
    outcome).  The device fit runs on worker threads (the C calls release the GIL) while ``make_models()`` -- the
    per-outcome model objects the caller wants back -- is evaluated on the calling thread.  Returns (raw result dict,
    models); the fitted values are written into the models' parameters.

    The outcomes are fitted as concurrent pieces (``split_for_streams``).  As each piece finishes, the calling thread
    writes its values into its models and calls ``post(lo, hi, models[lo:hi])`` -- the caller's post-fit work on that
    piece (pruning, feature importances) then runs behind the pieces that are still on the device."""
    import queue
    import threading
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    B = Y.shape[0]
    prog = template.program()
    lik_name, lik_param = likelihood_key(template)
    P = max(1, prog.n_x)
    n = X.shape[0]
    npad = ((n + 1 + 7) // 8 * 8 + 63) // 64 * 64
    chunk = max(1, int(max_batch_bytes // (2 * npad * npad * 8 + npad * 64 * 8)))
    out = dict(x=np.empty((B, P)), f=np.empty(B), lml=np.empty(B), n_iter=np.empty(B, np.int32),
               n_eval=np.empty(B, np.int32), status=np.empty(B, np.int32), launches=0, rounds=0)
    err = []
    pieces = split_for_streams(B, chunk, 1 if engine is not None else None)
    finished = queue.Queue()
    lock = threading.Lock()

    def piece_done(i, rc):
        (lo, hi), (r, c) = pieces[i], rc
        for key in ("x", "f", "lml", "n_iter", "n_eval", "status"):
            out[key][lo:hi] = r[key]
        with lock:
            out["launches"] += c["launches"]
            out["rounds"] += c["rounds"]
        finished.put(i)

    def work():
        try:
            jobs = [dict(X=X, Y=Y[lo:hi], table=[prog], P=P, lik_name=lik_name, lik_param=lik_param,
                         specialize=specialize) for lo, hi in pieces]
            run_fit_jobs(jobs, engine=engine, on_done=piece_done, **lbfgs_opts)
        except BaseException as e:      # re-raised on the calling thread
            err.append(e)
        finally:
            finished.put(None)          # no more pieces will arrive

    t = threading.Thread(target=work)
    t.start()
    try:
        models = make_models() if make_models is not None else [K.deepcopy(template) for _ in range(B)]
        packed = [packed_parameters(m) for m in models]        # still behind the device fit
        template_params = packed_parameters(template)
        while True:
            i = finished.get()
            if i is None:
                break
            lo, hi = pieces[i]
            # fitted values: the bijector of every packed position applied to its whole column (all copies share the
            # template's transforms), then plain stores into the models' Parameter objects
            values = np.empty((hi - lo, len(template_params)))
            for j, p in enumerate(template_params):
                values[:, j] = p.transform_fn(np.ascontiguousarray(out["x"][lo:hi, j]))
            lml, lpd = out["lml"][lo:hi].tolist(), (-out["f"][lo:hi]).tolist()
            nit, nev, st = out["n_iter"][lo:hi].tolist(), out["n_eval"][lo:hi].tolist(), out["status"][lo:hi].tolist()
            for b, (m, ps, row) in enumerate(zip(models[lo:hi], packed[lo:hi], values.tolist())):
                for p, v in zip(ps, row):
                    p._value = v
                m.log_marginal_likelihood_value = lml[b]
                m.log_posterior_density_value = lpd[b]
                m.fit_info = dict(n_iter=nit[b], n_eval=nev[b], status=st[b])
            if post is not None and not err:
                post(lo, hi, models[lo:hi])
    finally:
        t.join()
    if err:
        raise err[0]
    return out, models


def kernel_test_reg(X, Y, k, num_restarts=5, random_init=True, verbose=False, likelihood="gaussian", lasso=False,
                    lam=0, use_priors=True, max_iter=50000, keep_data=False, X_holdout=None, Y_holdout=None,
                    split=False, freeze_variances=False, random_seed=None, engine=None, **unused):
    """waveome/model_fitting.py:16-373 without the lasso (SVPGPR) branch.  ``likelihood="gaussian"`` is the exact GPR
    (:150-155); "exponential" / "poisson" / "gamma" / "bernoulli" are the reference's ``gpflow.models.VGP`` branches
    (:156-185, zero mean), fitted on the engine's collapsed bound max_q ELBO.  The ``num_restarts`` restarts are
    one device batch (same y, different starts) instead of a Python loop.  ``split=True`` scores the best model on
    (``X_holdout``, ``Y_holdout``) instead: bic = round(-sum predict_log_density, 2) (:337-347).  The remaining
    reference arguments (``gam``, ``base_variances``, ``freeze_inducing``, ``num_inducing_points``) belong to the lasso /
    sparse branches and are ignored."""
    vgp_likelihoods = ("exponential", "poisson", "gamma", "bernoulli")
    if lasso or likelihood not in ("gaussian",) + vgp_likelihoods:
        raise NotImplementedError("kernel_test_reg on the B200 engine covers likelihood='gaussian' (exact GPR) and "
                                  "'exponential' / 'poisson' / 'gamma' / 'bernoulli' (VGP) with lasso=False; the "
                                  "SVPGPR (lasso) branch is not on the engine")
    from .utilities import freeze_variance_parameters
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64).reshape(-1)
    if random_seed is not None:
        np.random.seed(random_seed)
    models = []
    for _ in range(num_restarts):
        if likelihood == "gaussian":
            m = GPR(K.deepcopy(k))                  # Zero mean, noise variance 1.0 (reference :151-155)
        else:
            m = GPR(K.deepcopy(k), likelihood=make_likelihood(likelihood))   # VGP, zero mean (:163-185)
        if freeze_variances:
            freeze_variance_parameters(m.kernel)
        if lam > 0:
            for name, p in m.parameter_dict().items():
                if "kernel" in name and "variance" in name:
                    p.prior = K.Laplace(0.0, 1.0 / lam)
        if use_priors:
            for name, p in m.parameter_dict().items():
                if "kernel" in name and "variance" not in name and "W" not in name:
                    p.prior = K.Uniform(0.0, 10.0)
        if random_init:                              # reference :245-259: unconstrained ~ N(0, 1)
            for p in m.kernel.trainable_parameters:
                p.assign(p.transform_fn(np.random.normal(size=1)))
            for p in m.likelihood.parameters:
                if p.trainable:
                    p.assign(p.transform_fn(np.random.normal(size=1)))
        models.append(m)
    res = fit_models(X, np.tile(Y, (num_restarts, 1)), models, engine=engine, maxiter=max_iter)
    best_model, best_loglik = None, -np.inf
    for b, m in enumerate(models):
        st = int(res["status"][b])
        if st & 1 or not np.isfinite(res["f"][b]):   # exception / not invertible -> restart skipped (:290-309)
            continue
        cur = -float(res["f"][b])
        if cur > best_loglik:
            best_loglik, best_model = cur, m
    if best_model is None:
        return None, -1 * best_loglik
    # :353-361: k = number of trainable Parameter objects; a gpflow VGP also carries q_mu and q_sqrt, which the
    # collapsed bound has maximised out
    if split:
        if X_holdout is None or Y_holdout is None:
            raise ValueError("kernel_test_reg(split=True) needs X_holdout and Y_holdout")
        held_out = best_model.predict_log_density((np.asarray(X_holdout, dtype=np.float64), Y_holdout),
                                                  data=(X, Y.reshape(-1, 1)))
        bic = round(-1 * float(np.sum(held_out)), 2)
    else:
        n_par = len(best_model.trainable_parameters) + (2 if likelihood in vgp_likelihoods else 0)
        bic = round(calc_bic(loglik=best_loglik, n=X.shape[0], k=n_par), 2)
    if verbose:
        print(f"Model: {print_kernel_names(k)}, BIC: {bic}")
    best_model.data = (X, Y.reshape(-1, 1)) if keep_data else None
    return best_model, bic
