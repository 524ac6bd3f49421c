#!/usr/bin/env python
"""Benchmark of the model-fitting hot path (BASELINE.json: "GP model fits/sec & LML+grad evals/sec, fp64, batched").

Workload (config.workload): BASELINE.json configs[2] — iHMP-scale synthetic metabolome, n = 600 samples x 5
covariates, 2000 Gaussian outcomes PER GPU, saturated horseshoe-penalised kernel (9 additive components, 17
trainable parameters), L-BFGS-B with SciPy/GPflow options (maxiter = maxfun = 50000).  It is the configuration the
north star quotes its target on and it fits one GPU (13.8 GB of workspace).  The other configs are parity-test cases.

A "step" is one complete batched MAP fit of the rank's 2000 models (about 90 LML+gradient evaluations per model).
  value  fits/s with X, Y resident in HBM (engine batch built outside the timed region)
  e2e    fits/s through the public API (GPSearch.penalized_optimization on pandas/host inputs: standardisation,
         kernel build, program encoding, H2D of X/Y/x0, fit, D2H of results, structure pruning)
Weak scaling: every rank fits its own 2000 outcomes; there is no collective on the data path.

`--impl reference` times the CPU restatement of the reference path (oracle/gp_oracle.py: NumPy/SciPy objective +
scipy L-BFGS-B — GPflow/TensorFlow are not installable here) on all host cores, one outcome per core.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

N_SUBJECTS, N_VISITS = 120, 5
FP64_PEAK_TFLOPS = 35.5     # cuBLAS DGEMM 8192^3 measured on this pool's B200 (profiles/r01_library_fp64_context.log);
                            # MEASURED_PEAKS.json carries no fp64 entry.  Raw DMMA issue peak: 37.1 (r01_fp64_peak_microbench.log)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def make_workload(n_outcomes, seed):
    from waveome_b200 import datasets
    X, Y = datasets.ihmp_scale(n_subjects=N_SUBJECTS, n_visits=N_VISITS, n_outcomes=n_outcomes, seed=seed)
    return X, Y


def make_search(X, Y):
    from waveome_b200.model_search import GPSearch
    return GPSearch(X, Y, unit_col="participant", categorical_vars=["participant", "sex", "site"],
                    Y_transform="standardize")


def build_model(gps):
    import waveome_b200 as wb
    from waveome_b200.regularization import full_kernel_build
    k = full_kernel_build(cat_vars=gps.cat_idx, num_vars=gps.cont_idx, unit_idx=gps.unit_idx, return_sum=True)
    return wb.models.PenalizedGPR(k, mean_function=wb.ConstantMean(), penalization_factor=1.0)


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle restatement, one outcome per host core
# ------------------------------------------------------------------------------------------------
def _cpu_fit_one(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    spec, X, y = args
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(1)
    except Exception:
        ctx = None
    import gp_oracle
    t0 = time.perf_counter()
    r = gp_oracle.fit(spec, X, y, maxiter=50000, maxfun=50000)
    if ctx is not None:
        ctx.restore_original_limits()
    return r["nfev"], r["status"], time.perf_counter() - t0, float(r["f"]), int(r["nit"]), [float(v) for v in r["x"]]


def cpu_reference_step(spec, Xn, Yn, cols, cores):
    """Fit outcomes `cols` on `cores` processes; returns (fits, evals, seconds)."""
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    t0 = time.perf_counter()
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn")) as ex:
        out = list(ex.map(_cpu_fit_one, [(spec, Xn, Yn[:, c].copy()) for c in cols]))
    dt = time.perf_counter() - t0
    cpu_reference_step.last = out          # per-outcome (nfev, status, seconds, f, nit) for the parity sample
    return len(cols), sum(o[0] for o in out), dt


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.stop_flag = gpu_index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in o.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# measurements of the other BASELINE configs and of the strong-scaling point, reported next to the headline
# ------------------------------------------------------------------------------------------------
def measure_extras(args, rank, world, local_rank, eng, barrier, reduce_max, torch):
    """Each entry carries its own device-clock time: CUDA events on the engine stream, recorded after a barrier +
    synchronize and read after the call returned (every entry point below synchronises its streams before returning),
    max over ranks.  A failing entry reports its error instead of taking the headline line down."""
    import numpy as np
    import waveome_b200 as wb
    from waveome_b200 import datasets
    from waveome_b200.engine import Batch
    from waveome_b200.model_search import GPSearch, shard_bounds
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local_rank))
    out = {}

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r = fn()
        e1.record(stream)
        barrier()
        return reduce_max(e0.elapsed_time(e1)) * 1e-3, r

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as e:                       # noqa: BLE001 -- reported in the line
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            barrier()

    # ---- configs[2] as BASELINE.json words it: 2000 outcomes IN TOTAL, sharded over the ranks (strong scaling)
    def strong():
        total = args.outcomes
        X, Y = make_workload(total, seed=2024)
        gps = make_search(X, Y)
        model = build_model(gps)
        lo, hi = shard_bounds(total, rank, world)
        Xn = gps.X.to_numpy(dtype=np.float64)
        Yn = np.ascontiguousarray(gps.Y.iloc[:, lo:hi].to_numpy(dtype=np.float64).T)
        batch = Batch(eng, Xn, Yn, [model.program()], specialize=True)
        batch.set_solo(True)            # one batch per GPU, nothing beside it (what fit_models does for a single job)
        x0 = batch.x0()
        batch.fit(x0, maxiter=50000, maxfun=50000)
        steps = max(1, min(args.steps, 3))
        secs, _ = timed(lambda: [batch.fit(x0, maxiter=50000, maxfun=50000) for _ in range(steps)])
        batch.close()
        return {"outcomes_total": total, "outcomes_per_gpu": hi - lo, "steps": steps, "value": total * steps / secs,
                "unit": "fits/s", "ms_per_step": 1e3 * secs / steps, "scaling": "strong"}
    if world > 1:
        guarded("strong_scaling", strong)

    # ---- configs[0]: the README quick-start (iris: 150 rows, 2 outcomes) -- a latency-bound job; single-GPU runs only
    def config1():
        X, Y = datasets.iris()
        def run():
            gps = GPSearch(X, Y, categorical_vars=["species"])
            gps.penalized_optimization(gather=False)
            return gps
        run()                                    # warm-up (buffers)
        secs, gps = timed(run)
        return {"outcomes": int(Y.shape[1]), "n": int(X.shape[0]), "seconds": secs,
                "structures": {o: m.kernel_name for o, m in gps.models.items()}}
    if world == 1:
        guarded("config1_iris", config1)

    # ---- configs[1]: full kernel search, 200 outcomes x depth 5 (outcomes sharded over the ranks by run_search)
    def config2():
        X, Y = datasets.overview_synthetic(n_people=50, n_observations=10, n_outcomes=200)
        gps = GPSearch(X, Y, unit_col="person_id", categorical_vars=["female"])
        secs, _ = timed(lambda: gps.run_search(max_depth=5, gather=False))
        return {"outcomes": 200, "max_depth": 5, "search_s": secs, "fits_this_rank": int(gps.fit_report["n_fits"]),
                "batches_this_rank": int(gps.fit_report["batches"]), "outcomes_per_s": 200 / secs}
    guarded("config2_search", config2)

    # ---- configs[3]: one n = 8192 GPR (large-n schedule); replicas only -- every rank runs its own copy
    def config4():
        X, Y = datasets.large_gpr(512, 16)
        Xn = X.to_numpy().copy()
        Xn[:, 1] = (Xn[:, 1] - Xn[:, 1].mean()) / Xn[:, 1].std()
        y = Y.to_numpy()[:, 0]
        cat = wb.Categorical(active_dims=[0]); wb.set_trainable(cat.variance, False)
        k = wb.Sum([wb.Product([cat, wb.SquaredExponential(active_dims=[1], lengthscales=0.5)]),
                    wb.Periodic(wb.SquaredExponential(active_dims=[1]), period=0.9)])
        model = wb.GPR(k, mean_function=wb.ConstantMean(), noise_variance=0.1)
        batch = Batch(eng, Xn, y[None, :], [model.program()])
        x = batch.x0()
        batch.eval(x)
        batch.profile(True)
        reps = 3
        secs, _ = timed(lambda: [batch.eval(x) for _ in range(reps)])
        prof = batch.profile_read()
        batch.close()
        n = float(len(y))
        chol_ms = (prof["chol_diag"][0] + prof["chol_panel"][0] + prof["chol_syrk"][0]) / reps
        tf = n ** 3 / 3 / (chol_ms * 1e-3) / 1e12
        return {"n": int(n), "eval_ms": 1e3 * secs / reps, "evals_per_s": reps / secs, "cholesky_ms": chol_ms,
                "cholesky_tflops": tf, "cholesky_frac_of_fp64_peak": tf / FP64_PEAK_TFLOPS,
                "class_ms": {k_: round(v[0] / reps, 3) for k_, v in prof.items() if v[0] > 0}, "parallelism": "replicas only"}
    guarded("config4_large_gpr", config4)

    # ---- configs[4]: 1000 count outcomes (n = 500), variational bound, through the public API
    def config5():
        res = {}
        for fam in ("poisson", "negative_binomial"):
            X, Y = datasets.count_microbiome(n_outcomes=1000, family=fam)
            def run():
                g = GPSearch(X, Y, unit_col="subject", outcome_likelihood=fam)
                g.penalized_optimization(gather=False)
                return g
            secs, g = timed(run)
            res[fam] = {"outcomes": 1000, "n": int(len(X)), "seconds": secs, "fits_per_s": 1000 / secs}
        return res
    guarded("config5_counts", config5)
    return out


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--outcomes", type=int, default=2000, help="outcomes (models) per GPU")
    ap.add_argument("--cpu-sample", type=int, default=0, help="outcomes in the CPU baseline sample (0 = one per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-config measurements next to the headline")
    args = ap.parse_args()

    # stdout carries exactly ONE line (the JSON): libraries that print there (NCCL's version banner under torchrun,
    # joblib workers) are sent to stderr; the line itself is written to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    n = N_SUBJECTS * N_VISITS
    workload = (f"BASELINE configs[2] iHMP-scale synthetic metabolome: n={n}, D=5, {args.outcomes} Gaussian outcomes/GPU, "
                "saturated kernel (9 components, P=17), horseshoe pf=1.0, L-BFGS-B maxiter=maxfun=50000")
    config = {"workload": workload, "n": n, "D": 5, "outcomes_per_gpu": args.outcomes, "P": 17,
              "parallelism": f"outcome-sharded x{world}, no collective", "l2_policy": "inputs >> L2 (13.8 GB workspace/GPU)"}

    # -------------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        X, Y = make_workload(max(cores, args.cpu_sample or cores), seed=2024)
        gps = make_search(X, Y)
        spec = build_model(gps).to_spec()
        Xn, Yn = gps.X.to_numpy(), gps.Y.to_numpy()
        ncols = args.cpu_sample or cores
        cols = list(range(ncols))
        for _ in range(min(args.warmup, 1)):
            cpu_reference_step(spec, Xn, Yn, cols[: max(1, ncols // 4)], cores)
        fits = evals = 0
        secs = 0.0
        for _ in range(args.steps):
            f, e, dt = cpu_reference_step(spec, Xn, Yn, cols, cores)
            fits += f; evals += e; secs += dt
        v = fits / secs
        sample = f"{ncols} outcomes of the same workload per step, one per host core, BLAS threads pinned to 1"
        line = {"impl": "reference", "metric": "gp_model_fits_per_sec", "value": v, "unit": "fits/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "lml_grad_evals_per_sec": evals / secs,
                "cpu_baseline": {"value": v, "unit": "fits/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": v, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # -------------------------------------------------------------------------------- B200 arm
    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        t_ = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item())

    from waveome_b200.engine import Batch
    from waveome_b200.model_fitting import get_engine
    # ONE global workload of outcomes x world columns (identical on every rank); rank r owns the contiguous shard
    # [r * outcomes, (r + 1) * outcomes) -- exactly the split GPSearch.penalized_optimization makes under torchrun
    X, Y = make_workload(args.outcomes * world, seed=2024)
    gps = make_search(X, Y)
    model = build_model(gps)
    eng = get_engine(local_rank)
    from waveome_b200.model_search import shard_bounds
    lo, hi = shard_bounds(args.outcomes * world, rank, world)
    Xn = gps.X.to_numpy(dtype=np.float64)
    Yn = np.ascontiguousarray(gps.Y.iloc[:, lo:hi].to_numpy(dtype=np.float64).T)
    batch = Batch(eng, Xn, Yn, [model.program()], specialize=True)     # as penalized_optimization asks for this job size
    batch.set_solo(True)                # one batch per GPU, nothing beside it (what fit_models does for a single job)
    x0 = batch.x0()
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local_rank))
    opts = dict(maxiter=50000, maxfun=50000)

    for _ in range(args.warmup):
        batch.fit(x0, **opts)
    barrier()
    c0 = batch.counters()
    batch.profile(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    n_eval = 0
    status_hist = {}
    for _ in range(args.steps):
        r = batch.fit(x0, **opts)
        n_eval += int(r["n_eval"].sum()) + len(r["n_eval"])      # + the closing evaluation at the optimum
        for s in r["status"]:
            status_hist[int(s)] = status_hist.get(int(s), 0) + 1
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.summary()
    prof = batch.profile_read()
    batch.profile(False)
    c1 = batch.counters()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    tot = torch.tensor([args.steps * args.outcomes, n_eval, c1["launches"] - c0["launches"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    fits_total, evals_total, launches_total = [float(v) for v in tot.tolist()]
    value = fits_total / (ms_max * 1e-3)

    # ---------------- e2e through the public API: host pandas in (all outcomes of the job), fitted model objects out;
    # under torchrun every rank builds the same GPSearch and fits its own shard (no gather: results stay per rank)
    def e2e_step():
        g = make_search(X, Y)
        g.penalized_optimization(penalization_factor=1.0, gather=False)
        return g
    batch.close()
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = fits_total / float(t.item())
    # the API default gather=True on top (every rank ends with the fitted models of ALL outcomes: one all_gather_object
    # of pickled models); one step
    e2e_gather_value = None
    if world > 1:
        barrier()
        t0 = time.perf_counter()
        g = make_search(X, Y)
        g.penalized_optimization(penalization_factor=1.0, gather=True)
        barrier()
        e2e_gather_value = args.outcomes * world / reduce_max(time.perf_counter() - t0)
    h2d = Xn.nbytes + Yn.nbytes + x0.nbytes
    d2h = x0.nbytes + args.outcomes * (8 + 8 + 4 + 4 + 4)

    # ---------------- roofline of the dominant kernel class (live CUDA-event times over the timed region)
    peaks, peak_kind = load_peaks()
    model_evals = c1["model_evals"] - c0["model_evals"]
    dom = max(prof, key=lambda k: prof[k][0])
    dom_ms, dom_launches = prof[dom]
    step_ms = sum(v[0] for v in prof.values())
    if dom in ("gram", "grad"):
        alg = 8.0 * n * n * model_evals + 8.0 * n * 5 * model_evals          # bytes: K written / W read once + X
        achieved = alg / (dom_ms * 1e-3) / 1e9
        roof = {"kernel": f"wv_{dom}_kernel", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": {"gram": 3.397e9, "grad": 3.470e9}[dom], "peak_kind": peak_kind + " copy bandwidth",
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch with 2000 models in flight "
                                "(profiles/r02t_eval_dram_per_launch.csv); algorithmic bytes of that launch: 5.8e9 "
                                "(8 n^2 per model, the kernel touches the lower tiles only)",
                "note": "issue bound, not HBM bound: up to 6 squared-exponential leaves per matrix element for this "
                        "kernel tree, each a table-driven 2^u of 10 FP64 + 8 integer instructions (DESIGN.md section 4a')"}
    else:
        share = {"chol_diag": 1.0 / 3, "chol_panel": 1.0 / 3, "trtri": 1.0 / 3, "kinv": 1.0 / 3}.get(dom, 0.0)
        if dom in ("chol_diag", "chol_panel"):
            dom_ms = prof["chol_diag"][0] + prof["chol_panel"][0]
            dom = "chol_diag+chol_panel"
        alg = share * float(n) ** 3 * model_evals
        achieved = alg / (dom_ms * 1e-3) / 1e12
        # dram__bytes_read.sum + dram__bytes_write.sum over ALL launches of the class in one evaluation of 2000 models
        # (ncu --metrics pass, profiles/r02t_eval_dram_per_launch.csv)
        traffic = {"trtri": 16.86e9, "kinv": 6.31e9, "chol_diag+chol_panel": 20.82e9}.get(dom)
        roof = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": achieved / FP64_PEAK_TFLOPS, "traffic": traffic,
                "traffic_note": "CLASS TOTAL per evaluation of 2000 models: dram__bytes_read.sum + dram__bytes_write.sum summed over "
                                "all launches of the class in one evaluation (Cholesky: 10 diagonal + 9 panel launches, trtri: 9, "
                                "kinv: 1), profiles/r02t_eval_dram_per_launch.csv; the lower tiles of the 2000 matrices are 3.5e9 "
                                "bytes: the left-looking panel / trtri steps re-read the earlier tile columns (13.6 / 14.4e9 read). "
                                "achieved counts the algorithmic n^3/3 flops of the class, the kernels execute 1.06-1.25x that",
                "peak_kind": "measured cuBLAS DGEMM fp64 on this pool (not in MEASURED_PEAKS.json)"}
    roof["share_of_step"] = dom_ms / step_ms if step_ms else None
    roof["launches"] = dom_launches
    fact_ms = sum(prof[k][0] for k in ("chol_diag", "chol_panel", "trtri", "kinv"))
    groups = {
        "factorisation_tflops": float(n) ** 3 * model_evals / (fact_ms * 1e-3) / 1e12 if fact_ms else None,
        "factorisation_frac_of_fp64_peak": float(n) ** 3 * model_evals / (fact_ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS if fact_ms else None,
        "class_ms": {k: round(v[0], 3) for k, v in prof.items()},
        # algorithmic n^3/3 flops of each factorisation class over its own time (Cholesky = diagonal + panel kernels)
        "class_tflops": {
            "cholesky": float(n) ** 3 / 3 * model_evals / ((prof["chol_diag"][0] + prof["chol_panel"][0]) * 1e-3) / 1e12
            if prof["chol_diag"][0] + prof["chol_panel"][0] else None,
            "trtri": float(n) ** 3 / 3 * model_evals / (prof["trtri"][0] * 1e-3) / 1e12 if prof["trtri"][0] else None,
            "kinv": float(n) ** 3 / 3 * model_evals / (prof["kinv"][0] * 1e-3) / 1e12 if prof["kinv"][0] else None},
        # element-wise passes against HBM: algorithmic 8 n^2 (K written / W read once) + 8 n D bytes per model evaluation
        # (the kernels touch the lower tiles only, i.e. move ~0.6x of it), over the class's own time
        "class_hbm_gbs": {k: (8.0 * n * n + 8.0 * n * 5) * model_evals / (prof[k][0] * 1e-3) / 1e9 if prof[k][0] else None
                          for k in ("gram", "grad")},
        "hbm_peak_gbs": peaks["hbm_gbs"],
    }
    groups["class_hbm_frac"] = {k: (v / peaks["hbm_gbs"] if v else None) for k, v in groups["class_hbm_gbs"].items()}

    extras = {} if args.no_extras else measure_extras(args, rank, world, local_rank, eng, barrier, reduce_max, torch)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu = None
    if not args.no_cpu_baseline and world == 1:      # the CPU baseline is reported at N = 1 only
        ncols = args.cpu_sample or cores
        spec = model.to_spec()
        f, e, dt = cpu_reference_step(spec, Xn, gps.Y.to_numpy(dtype=np.float64)[:, lo:hi], list(range(ncols)), cores)
        cpu = {"value": f / dt, "unit": "fits/s", "cores": cores, "kind": "port",
               "sample": f"first {ncols} outcomes of rank 0's workload, one per host core, BLAS threads pinned to 1 "
                         f"({dt:.1f} s wall); oracle/gp_oracle.py + scipy L-BFGS-B",
               "lml_grad_evals_per_sec": e / dt}
        # the same outcomes as fitted by the engine in the last timed step: the oracle is the checker here
        same = [i for i, o in enumerate(cpu_reference_step.last)
                if o[1] == 0 and int(r["status"][i]) == 0 and o[0] == int(r["n_eval"][i]) and o[4] == int(r["n_iter"][i])]
        fin = [i for i, o in enumerate(cpu_reference_step.last) if np.isfinite(o[3]) and np.isfinite(r["f"][i])]
        rel = [abs(float(r["f"][i]) - cpu_reference_step.last[i][3]) / max(1.0, abs(cpu_reference_step.last[i][3])) for i in fin]
        def pruned(xv):          # selected structure: cut_kernel_components at the fitted values
            m = wb_kernels.deepcopy(model)
            m.program().assign(np.asarray(xv, dtype=np.float64))
            m.cut_kernel_components(Xn)
            m.update_kernel_name()
            return m.kernel_name
        from waveome_b200 import kernels as wb_kernels
        last = cpu_reference_step.last
        structure_identical = sum(1 for i, o in enumerate(last) if pruned(o[5]) == pruned(r["x"][i]))
        status_agree = sum(1 for i, o in enumerate(last) if (o[1] == 0) == (int(r["status"][i]) == 0))
        cpu["parity_sample"] = {
            "outcomes": ncols, "structure_identical": structure_identical, "status_agree": status_agree,
            "converged_on_both_with_identical_nit_nfev": len(same),
            "max_rel_objective_diff_identical_trajectories":
                max([abs(float(r["f"][i]) - cpu_reference_step.last[i][3]) / max(1.0, abs(cpu_reference_step.last[i][3]))
                     for i in same], default=None),
            "median_rel_objective_diff_all": float(np.median(rel)) if rel else None,
            "max_rel_objective_diff_all": max(rel, default=None),
            "note": "fits that enter the horseshoe's non-finite regime end ABNORMAL on both sides and are chaotic in the "
                    "last bits (DESIGN.md section 5): they agree in selected structure and objective value, not iteration by "
                    "iteration; tests/test_c3_parity_gpu.py compares the spread with the oracle's own last-bit sensitivity"}

    line = {"metric": "gp_model_fits_per_sec", "value": value, "unit": "fits/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "lml_grad_evals_per_sec": evals_total / (ms_max * 1e-3),
            "e2e": {"value": e2e_value, "unit": "fits/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "GPSearch.penalized_optimization (pandas in, fitted models out; gather=False)",
                    "value_gather_true": e2e_gather_value},
            "gpu_launches": int(launches_total), "clocks": clocks, "roofline": roof, "roofline_groups": groups,
            "cpu_baseline": cpu, "fit_status_hist": status_hist, "other_configs": extras}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
