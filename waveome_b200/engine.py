"""ctypes binding of the C ABI in include/waveome_b200.h (libwaveome_b200.so, built in-tree by
``__graft_entry__.build()``).

There is deliberately no CPU fallback: importing this module without the built library, or creating
an ``Engine`` without a CUDA device, raises.  PyTorch is used only as plumbing (device tensors for the
DLPack / device-pointer entry point, ``torch.distributed`` for sharding in ``model_search``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from .program import Program

_LIB_PATH = os.environ.get("WAVEOME_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib",
                                                              "libwaveome_b200.so")

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class _ProgramDesc(C.Structure):
    _fields_ = [
        ("n_comp", C.c_int32), ("n_leaves", C.c_int32), ("n_slots", C.c_int32), ("noise_slot", C.c_int32),
        ("mean_slot", C.c_int32),
        ("comp_start", _i32p), ("leaf_type", _i32p), ("leaf_dim", _i32p), ("leaf_s_var", _i32p),
        ("leaf_s_ls", _i32p), ("leaf_s_aux", _i32p), ("leaf_degree", _i32p),
        ("slot_transform", _i32p), ("slot_xindex", _i32p), ("slot_prior", _i32p),
        ("slot_fixed", _f64p), ("slot_shift", _f64p), ("slot_pa", _f64p), ("slot_pb", _f64p),
        ("lik_slot2", C.c_int32),
    ]


class _BatchDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("D", C.c_int32), ("B", C.c_int32), ("P", C.c_int32),
        ("X", _f64p), ("Y", _f64p), ("n_programs", C.c_int32), ("programs", C.POINTER(_ProgramDesc)),
        ("prog_id", _i32p),
    ]


class _LbfgsOpts(C.Structure):
    _fields_ = [("maxcor", C.c_int32), ("maxiter", C.c_int32), ("maxfun", C.c_int32), ("maxls", C.c_int32),
                ("ftol", C.c_double), ("gtol", C.c_double), ("chol_fail_policy", C.c_int32), ("reserved", C.c_int32)]


class _AdamOpts(C.Structure):
    _fields_ = [("learning_rate", C.c_double), ("decay_rate", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("epsilon", C.c_double), ("convergence_threshold", C.c_double), ("max_iter", C.c_int32),
                ("check_every", C.c_int32), ("decay_every", C.c_int32), ("reserved", C.c_int32)]


EXPORTED_SYMBOLS = [
    "wv_engine_create", "wv_engine_create2", "wv_batch_set_engine", "wv_engine_destroy", "wv_engine_stream", "wv_engine_set_large_n_tiles", "wv_batch_create", "wv_batch_destroy",
    "wv_batch_workspace_bytes", "wv_batch_set_y", "wv_batch_set_component_mask", "wv_batch_set_solo", "wv_batch_set_likelihood", "wv_batch_set_likelihood2", "wv_batch_get_latent", "wv_batch_eval", "wv_batch_eval_device", "wv_batch_fit_lbfgs",
    "wv_batch_fit_lbfgs_begin", "wv_batch_fit_lbfgs_run", "wv_batch_fit_lbfgs_report",
    "wv_batch_fit_adam", "wv_batch_create2", "wv_batch_eval_elbo", "wv_batch_specialize", "wv_rtc_check", "wv_rtc_set_cache", "wv_rtc_precompile_text",
    "wv_batch_counters", "wv_batch_get_alpha", "wv_batch_get_kinv_diag", "wv_batch_predict_mean", "wv_batch_predict_f", "wv_batch_profile_enable", "wv_batch_profile_read", "wv_last_error", "wv_version",
]
KERNEL_CLASSES = ["gram", "chol_diag", "chol_panel", "trtri", "extract", "kinv", "grad", "finalize", "lbfgs", "chol_syrk", "sites"]

_lib = None


def load_library():
    """Load libwaveome_b200.so or raise — never falls back to a CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"waveome_b200: CUDA extension not built ({_LIB_PATH} missing). Run `python -c 'import "
            "__graft_entry__ as g; g.build()'` from the repository root. There is no CPU fallback.")
    lib = C.CDLL(_LIB_PATH)
    vp = C.c_void_p
    lib.wv_engine_create.argtypes = [C.c_int, C.POINTER(vp)]; lib.wv_engine_create.restype = C.c_int
    lib.wv_engine_create2.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]; lib.wv_engine_create2.restype = C.c_int
    lib.wv_batch_set_engine.argtypes = [vp, vp]; lib.wv_batch_set_engine.restype = C.c_int
    lib.wv_batch_set_solo.argtypes = [vp, C.c_int]; lib.wv_batch_set_solo.restype = C.c_int
    lib.wv_engine_destroy.argtypes = [vp]; lib.wv_engine_destroy.restype = None
    lib.wv_engine_stream.argtypes = [vp]; lib.wv_engine_stream.restype = vp
    lib.wv_engine_set_large_n_tiles.argtypes = [vp, C.c_int]; lib.wv_engine_set_large_n_tiles.restype = C.c_int
    lib.wv_batch_create.argtypes = [vp, C.POINTER(_BatchDesc), C.POINTER(vp)]; lib.wv_batch_create.restype = C.c_int
    lib.wv_batch_create2.argtypes = [vp, C.POINTER(_BatchDesc), C.c_int32, C.POINTER(vp)]; lib.wv_batch_create2.restype = C.c_int
    lib.wv_batch_eval_elbo.argtypes = [vp, _f64p, _f64p, _f64p, C.c_double, _f64p, _f64p, _i32p]
    lib.wv_batch_eval_elbo.restype = C.c_int
    lib.wv_batch_destroy.argtypes = [vp]; lib.wv_batch_destroy.restype = None
    lib.wv_batch_workspace_bytes.argtypes = [vp]; lib.wv_batch_workspace_bytes.restype = C.c_int64
    lib.wv_batch_set_y.argtypes = [vp, _f64p]; lib.wv_batch_set_y.restype = C.c_int
    lib.wv_batch_set_component_mask.argtypes = [vp, C.POINTER(C.c_uint32)]; lib.wv_batch_set_component_mask.restype = C.c_int
    lib.wv_batch_set_likelihood.argtypes = [vp, C.c_int32, C.c_double]; lib.wv_batch_set_likelihood.restype = C.c_int
    lib.wv_batch_set_likelihood2.argtypes = [vp, C.c_int32, C.c_double, C.c_double]; lib.wv_batch_set_likelihood2.restype = C.c_int
    lib.wv_batch_get_latent.argtypes = [vp, _f64p, _f64p]; lib.wv_batch_get_latent.restype = C.c_int
    lib.wv_batch_eval.argtypes = [vp, _f64p, _f64p, _f64p, _f64p, _i32p]; lib.wv_batch_eval.restype = C.c_int
    lib.wv_batch_eval_device.argtypes = [vp, vp, vp, vp, vp, vp]; lib.wv_batch_eval_device.restype = C.c_int
    lib.wv_batch_fit_lbfgs.argtypes = [vp, _f64p, C.POINTER(_LbfgsOpts), _f64p, _f64p, _i32p, _i32p, _i32p]
    lib.wv_batch_fit_lbfgs.restype = C.c_int
    lib.wv_batch_fit_lbfgs_begin.argtypes = [vp, _f64p, C.POINTER(_LbfgsOpts)]
    lib.wv_batch_fit_lbfgs_begin.restype = C.c_int
    lib.wv_batch_fit_lbfgs_run.argtypes = [vp, C.c_int32, C.POINTER(C.c_int32)]
    lib.wv_batch_fit_lbfgs_run.restype = C.c_int
    lib.wv_batch_fit_lbfgs_report.argtypes = [vp, _f64p, _f64p, _f64p, _i32p, _i32p, _i32p, _i32p]
    lib.wv_batch_fit_lbfgs_report.restype = C.c_int
    lib.wv_batch_fit_adam.argtypes = [vp, _f64p, C.POINTER(_AdamOpts), _f64p, _f64p, _i32p, _i32p]
    lib.wv_batch_fit_adam.restype = C.c_int
    lib.wv_batch_counters.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.wv_batch_counters.restype = None
    lib.wv_batch_get_alpha.argtypes = [vp, _f64p]; lib.wv_batch_get_alpha.restype = C.c_int
    lib.wv_batch_get_kinv_diag.argtypes = [vp, _f64p]; lib.wv_batch_get_kinv_diag.restype = C.c_int
    lib.wv_batch_predict_mean.argtypes = [vp, _f64p, C.c_int32, _f64p]; lib.wv_batch_predict_mean.restype = C.c_int
    lib.wv_batch_predict_f.argtypes = [vp, _f64p, C.c_int32, _f64p, _f64p]; lib.wv_batch_predict_f.restype = C.c_int
    lib.wv_batch_profile_enable.argtypes = [vp, C.c_int]; lib.wv_batch_profile_enable.restype = None
    lib.wv_batch_profile_read.argtypes = [vp, _f64p, C.POINTER(C.c_int64), C.c_int]; lib.wv_batch_profile_read.restype = C.c_int
    lib.wv_batch_specialize.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_int32]
    lib.wv_batch_specialize.restype = C.c_int
    lib.wv_rtc_check.argtypes = [C.c_char_p, C.c_char_p, C.c_int]; lib.wv_rtc_check.restype = C.c_int
    lib.wv_rtc_set_cache.argtypes = [C.c_char_p]; lib.wv_rtc_set_cache.restype = None
    lib.wv_rtc_precompile_text.argtypes = [C.c_char_p, C.c_char_p]; lib.wv_rtc_precompile_text.restype = C.c_int
    # compiled specialisations are kept next to the library (in-tree: they travel with it), WV_RTC_CACHE overrides
    cache = os.environ.get("WV_RTC_CACHE", os.path.join(os.path.dirname(_LIB_PATH), "rtc_cache"))
    if cache:
        try:
            os.makedirs(cache, exist_ok=True)
        except OSError:
            cache = ""
    lib.wv_rtc_set_cache(cache.encode() if cache else None)
    lib.wv_last_error.argtypes = []; lib.wv_last_error.restype = C.c_char_p
    lib.wv_version.argtypes = []; lib.wv_version.restype = C.c_char_p
    _lib = lib
    return lib


class EngineError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise EngineError(f"{what} failed: {load_library().wv_last_error().decode()}")


def _f64(a):
    return a.ctypes.data_as(_f64p)


def _i32(a):
    return a.ctypes.data_as(_i32p)


#: SciPy L-BFGS-B defaults (scipy.optimize._lbfgsb_py._minimize_lbfgsb); waveome overrides maxiter/maxfun
#: with 50000 at waveome/model_classes.py:310-315 and maxiter at waveome/model_fitting.py:280.
DEFAULT_LBFGS = dict(maxcor=10, maxiter=15000, maxfun=15000, maxls=20, ftol=2.220446049250313e-09, gtol=1e-05,
                     on_chol_fail="nan")


#: Jobs of at least this many models that share ONE program structure (GPSearch.penalized_optimization: every outcome x
#: the saturated kernel) ask for run-time specialised Gram / gradient kernels (specialize.py -> NVRTC, ~4 s once per
#: structure and machine, then a disk cache); the decision is taken on the WHOLE job, not per shard or sub-batch.
#: WV_SPECIALIZE=0 turns the specialisation off, WV_SPECIALIZE=1 forces it for every single-structure batch.
SPECIALIZE_MIN_MODELS = 64
_spec_warned = False


def rtc_check(source: str) -> int:
    """Compile CUDA text with NVRTC for sm_100a (no GPU needed); returns the cubin size or raises with the log."""
    lib = load_library()
    log = C.create_string_buffer(1 << 16)
    rc = lib.wv_rtc_check(source.encode(), log, len(log))
    if rc < 0:
        raise EngineError("NVRTC: " + (log.value.decode(errors="replace") or lib.wv_last_error().decode()))
    return rc


def rtc_precompile(programs) -> int:
    """Compile the specialised kernels of the given programs into the disk cache (no GPU needed); returns how many texts
    were compiled now.  ``__graft_entry__.build()`` does this for the benchmark's kernel structure."""
    from . import specialize as sp
    lib = load_library()
    done = 0
    for p in programs:
        s = sp.generate(p)
        if s is None:
            continue
        rc = lib.wv_rtc_precompile_text(s.key.encode(), s.source.encode())
        if rc < 0:
            raise EngineError(lib.wv_last_error().decode())
        done += rc == 0
    return done


#: BaseGP.optimize_params defaults (waveome/model_classes.py:236-246) + Keras Adam's
DEFAULT_ADAM = dict(learning_rate=0.1, decay_rate=0.96, beta1=0.9, beta2=0.999, epsilon=1e-7, convergence_threshold=1e-9,
                    max_iter=50000, check_every=100, decay_every=500)


class Engine:
    """One per GPU / host thread; owns the CUDA stream the batches launch on."""

    def __init__(self, device: int = 0, large_n_tiles: Optional[int] = None, high_priority: bool = False):
        """``high_priority``: the stream is scheduled ahead of the default engines' streams (stragglers of a search
        level that finish while the next level's batch runs, ``Batch.move_to``)."""
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib.wv_engine_create2(int(device), 1 if high_priority else 0, C.byref(h)), "wv_engine_create2")
        self.handle = h
        self.device = int(device)
        if large_n_tiles is not None:
            self.set_large_n_tiles(large_n_tiles)

    def set_large_n_tiles(self, nt: int):
        """Tile count ((n + 1 + 63) // 64) from which batches use the large-n schedule (default 16)."""
        _check(self.lib.wv_engine_set_large_n_tiles(self.handle, int(nt)), "wv_engine_set_large_n_tiles")

    @property
    def stream(self) -> int:
        return int(self.lib.wv_engine_stream(self.handle) or 0)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.wv_engine_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """B independent GP models on shared covariates X: y_b ~ GP(mean_b, k_b) + noise."""

    def __init__(self, engine: Engine, X: np.ndarray, Y: np.ndarray, programs: Sequence[Program],
                 prog_id: Optional[Sequence[int]] = None, P: Optional[int] = None, specialize=None,
                 keep_row_order: bool = False):
        self.engine = engine
        self.lib = engine.lib
        X = np.ascontiguousarray(X, dtype=np.float64)
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        if X.ndim != 2 or Y.ndim != 2 or Y.shape[1] != X.shape[0]:
            raise ValueError("X must be [n, D] and Y must be [B, n]")
        self.n, self.D = X.shape
        self.B = Y.shape[0]
        self.programs: List[Program] = list(programs)
        pid = np.zeros(self.B, np.int32) if prog_id is None else np.ascontiguousarray(prog_id, dtype=np.int32)
        if pid.shape != (self.B,):
            raise ValueError("prog_id must have one entry per model")
        self.prog_id = pid
        self.P = int(P) if P is not None else max(1, max(p.n_x for p in self.programs))
        descs = (_ProgramDesc * len(self.programs))()
        for d, p in zip(descs, self.programs):
            d.n_comp, d.n_leaves, d.n_slots = p.n_comp, p.n_leaves, p.n_slots
            d.noise_slot, d.mean_slot = p.noise_slot, p.mean_slot
            d.comp_start, d.leaf_type, d.leaf_dim = _i32(p.comp_start), _i32(p.leaf_type), _i32(p.leaf_dim)
            d.leaf_s_var, d.leaf_s_ls, d.leaf_s_aux = _i32(p.leaf_s_var), _i32(p.leaf_s_ls), _i32(p.leaf_s_aux)
            d.leaf_degree = _i32(p.leaf_degree)
            d.slot_transform, d.slot_xindex, d.slot_prior = _i32(p.slot_transform), _i32(p.slot_xindex), _i32(p.slot_prior)
            d.slot_fixed, d.slot_shift, d.slot_pa, d.slot_pb = _f64(p.slot_fixed), _f64(p.slot_shift), _f64(p.slot_pa), _f64(p.slot_pb)
            d.lik_slot2 = getattr(p, "lik_slot2", -1)
        bd = _BatchDesc(self.n, self.D, self.B, self.P, _f64(X), _f64(Y), len(self.programs), descs, _i32(pid))
        h = C.c_void_p()
        _check(self.lib.wv_batch_create2(engine.handle, C.byref(bd), 1 if keep_row_order else 0, C.byref(h)), "wv_batch_create2")
        self.handle = h
        # Run-time specialised element-wise kernels are the CALLER's choice, never a function of the batch size: a model's
        # result must not depend on how many neighbours share its batch (specialised and interpreter kernels agree to
        # ~1e-13, not bit for bit).  WV_SPECIALIZE=0 / 1 overrides every caller.
        self.specialized = False
        env = os.environ.get("WV_SPECIALIZE", "")
        want = bool(specialize) and env != "0" or env == "1"
        if want:
            self.specialize(X)

    # ------------------------------------------------------------------------------------------
    def specialize(self, X: Optional[np.ndarray] = None, strict: bool = False) -> bool:
        """Route the batch's Gram / gradient passes to run-time specialised kernels (specialize.generate -> NVRTC) when
        all its programs share one structure the generator covers; returns whether it happened.  ``strict`` raises
        instead of staying on the interpreter kernels."""
        global _spec_warned
        from . import specialize as sp
        srcs = [sp.generate(p) for p in self.programs]
        why = None
        if any(s is None for s in srcs):
            why = "a program uses a leaf the generator does not cover"
        elif len({s.key for s in srcs}) != 1:
            why = "the batch mixes program structures"
        elif X is not None:
            cat_dims = {int(d) for p in self.programs for t, d in zip(p.leaf_type, p.leaf_dim) if int(t) == sp.CAT}
            if cat_dims and not np.all(np.abs(np.rint(np.asarray(X)[:, sorted(cat_dims)])) < 2 ** 31 - 1):
                why = "categorical codes do not fit int32"
        if why is None:
            s = srcs[0]
            rc = self.lib.wv_batch_specialize(self.handle, s.key.encode(), s.source.encode(), s.gram_name.encode(),
                                              s.grad_name.encode(), s.gram_smem, s.grad_smem)
            if rc == 0:
                self.specialized = True
                return True
            why = self.lib.wv_last_error().decode()
            if not _spec_warned:
                _spec_warned = True
                import sys
                print(f"waveome_b200: element-wise kernels not specialised ({why}); interpreter kernels stay in charge",
                      file=sys.stderr)
        if strict:
            raise EngineError(f"Batch.specialize: {why}")
        return False

    def unspecialize(self):
        """Back to the interpreter kernels (tests compare the two paths on identical inputs)."""
        _check(self.lib.wv_batch_specialize(self.handle, None, None, None, None, 0, 0), "wv_batch_specialize")
        self.specialized = False

    def x0(self) -> np.ndarray:
        """[B, P] unconstrained start vectors from the programs' current parameter values."""
        x = np.zeros((self.B, self.P))
        starts = [p.x0() for p in self.programs]
        for b in range(self.B):
            s = starts[self.prog_id[b]]
            x[b, : s.size] = s
        return x

    def set_y(self, Y: np.ndarray):
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        if Y.shape != (self.B, self.n):
            raise ValueError("Y must be [B, n]")
        _check(self.lib.wv_batch_set_y(self.handle, _f64(Y)), "wv_batch_set_y")

    def set_component_mask(self, mask):
        """[B] uint32: bit c enables additive component c of the model's program (default: all)."""
        mask = np.ascontiguousarray(mask, dtype=np.uint32)
        if mask.shape != (self.B,):
            raise ValueError("mask must have one entry per model")
        _check(self.lib.wv_batch_set_component_mask(self.handle, mask.ctypes.data_as(C.POINTER(C.c_uint32))),
               "wv_batch_set_component_mask")

    LIKELIHOODS = {"gaussian": 0, "poisson": 1, "negative_binomial": 2, "bernoulli": 3, "gamma": 4, "zinb": 5}

    def set_likelihood(self, kind, param=0.0, param2: float = 1.0):
        """"gaussian" (default), "poisson", "negative_binomial" (param = alpha), "bernoulli", "gamma" (param = shape) or
        "zinb" (param = alpha, param2 = km; ``param`` may be the pair): switches the objective to the variational bound
        maximised over q (include/waveome_b200.h)."""
        code = self.LIKELIHOODS[kind] if isinstance(kind, str) else int(kind)
        if isinstance(param, (tuple, list)):
            param, param2 = param
        _check(self.lib.wv_batch_set_likelihood2(self.handle, code, float(param), float(param2)), "wv_batch_set_likelihood2")

    def latent(self):
        """(mean, var) of f at the training inputs after the last evaluation of a non-Gaussian batch, [B, n] each."""
        m = np.empty((self.B, self.n)); v = np.empty((self.B, self.n))
        _check(self.lib.wv_batch_get_latent(self.handle, _f64(m), _f64(v)), "wv_batch_get_latent")
        return m, v

    def eval(self, x: np.ndarray):
        """One LML+gradient evaluation from HOST buffers.  Returns (f, grad, lml, status)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape != (self.B, self.P):
            raise ValueError(f"x must be [{self.B}, {self.P}]")
        f = np.empty(self.B); g = np.empty((self.B, self.P)); lml = np.empty(self.B)
        st = np.empty(self.B, np.int32)
        _check(self.lib.wv_batch_eval(self.handle, _f64(x), _f64(f), _f64(g), _f64(lml), _i32(st)), "wv_batch_eval")
        return f, g, lml, st

    def eval_device(self, x, f, grad, lml, status):
        """Same with DEVICE buffers (torch CUDA tensors or anything exposing ``data_ptr()`` /
        ``__dlpack__``), enqueued on the engine stream without host synchronisation."""
        ptrs = []
        for t in (x, f, grad, lml, status):
            if not hasattr(t, "data_ptr"):
                import torch
                t = torch.from_dlpack(t)
            ptrs.append(C.c_void_p(t.data_ptr()))
        _check(self.lib.wv_batch_eval_device(self.handle, *ptrs), "wv_batch_eval_device")

    def fit(self, x0: Optional[np.ndarray] = None, **opts):
        """Batched L-BFGS-B MAP fit.  Returns dict(x, f, lml, n_iter, n_eval, status)."""
        o = dict(DEFAULT_LBFGS); o.update(opts)
        x = self.x0() if x0 is None else np.array(x0, dtype=np.float64, order="C", copy=True)
        if x.shape != (self.B, self.P):
            raise ValueError(f"x0 must be [{self.B}, {self.P}]")
        if o["on_chol_fail"] not in ("nan", "abort"):
            raise ValueError("on_chol_fail must be 'nan' (failed trial = non-finite value) or 'abort'")
        co = _LbfgsOpts(int(o["maxcor"]), int(o["maxiter"]), int(o["maxfun"]), int(o["maxls"]), float(o["ftol"]),
                        float(o["gtol"]), 1 if o["on_chol_fail"] == "abort" else 0, 0)
        f = np.empty(self.B); lml = np.empty(self.B)
        nit = np.empty(self.B, np.int32); nev = np.empty(self.B, np.int32); st = np.empty(self.B, np.int32)
        _check(self.lib.wv_batch_fit_lbfgs(self.handle, _f64(x), C.byref(co), _f64(f), _f64(lml), _i32(nit), _i32(nev),
                                           _i32(st)), "wv_batch_fit_lbfgs")
        return dict(x=x, f=f, lml=lml, n_iter=nit, n_eval=nev, status=st)

    def set_solo(self, solo: bool = True):
        """Scheduling hint (wv_batch_set_solo): this batch has the device to itself while its calls run."""
        _check(self.lib.wv_batch_set_solo(self.handle, 1 if solo else 0), "wv_batch_set_solo")

    def move_to(self, engine: "Engine"):
        """Re-home the batch: later calls run on ``engine``'s stream (same device; no call may be in progress)."""
        _check(self.lib.wv_batch_set_engine(self.handle, engine.handle), "wv_batch_set_engine")
        self.engine = engine

    def fit_begin(self, x0: Optional[np.ndarray] = None, **opts):
        """First of the three calls ``fit`` consists of (include/waveome_b200.h: wv_batch_fit_lbfgs_begin / _run /
        _report): upload the starts.  Then ``fit_run(min_active)`` iterates while more than ``min_active`` models are
        unfinished and ``fit_report()`` returns ``fit``'s dict plus ``finished`` [B] (bool) at any point in between."""
        o = dict(DEFAULT_LBFGS); o.update(opts)
        x = self.x0() if x0 is None else np.array(x0, dtype=np.float64, order="C", copy=True)
        if x.shape != (self.B, self.P):
            raise ValueError(f"x0 must be [{self.B}, {self.P}]")
        if o["on_chol_fail"] not in ("nan", "abort"):
            raise ValueError("on_chol_fail must be 'nan' (failed trial = non-finite value) or 'abort'")
        co = _LbfgsOpts(int(o["maxcor"]), int(o["maxiter"]), int(o["maxfun"]), int(o["maxls"]), float(o["ftol"]),
                        float(o["gtol"]), 1 if o["on_chol_fail"] == "abort" else 0, 0)
        _check(self.lib.wv_batch_fit_lbfgs_begin(self.handle, _f64(x), C.byref(co)), "wv_batch_fit_lbfgs_begin")

    def fit_run(self, min_active: int = 0) -> int:
        """Optimiser rounds while more than ``min_active`` models are still iterating; returns how many are left."""
        left = C.c_int32(0)
        _check(self.lib.wv_batch_fit_lbfgs_run(self.handle, int(min_active), C.byref(left)), "wv_batch_fit_lbfgs_run")
        return int(left.value)

    def fit_report(self):
        """dict(x, f, lml, n_iter, n_eval, status, finished) of the fit in progress; f / lml / status are meaningful for
        the finished models only."""
        x = np.empty((self.B, self.P)); f = np.empty(self.B); lml = np.empty(self.B)
        nit = np.empty(self.B, np.int32); nev = np.empty(self.B, np.int32); st = np.empty(self.B, np.int32)
        fin = np.empty(self.B, np.int32)
        _check(self.lib.wv_batch_fit_lbfgs_report(self.handle, _f64(x), _f64(f), _f64(lml), _i32(nit), _i32(nev), _i32(st),
                                                  _i32(fin)), "wv_batch_fit_lbfgs_report")
        return dict(x=x, f=f, lml=lml, n_iter=nit, n_eval=nev, status=st, finished=fin != 0)

    def eval_elbo(self, x: np.ndarray, q_mu: np.ndarray, q_sqrt: np.ndarray, jitter: float = 1e-6):
        """Objective (B) at given variational parameters: (elbo [B], f = -(elbo + log prior) [B], status [B]) of the
        whitened VGP / SVGP-with-Z = X bound (include/waveome_b200.h: wv_batch_eval_elbo).  q_mu [B, n], q_sqrt [B, n, n]
        (lower triangles).  The batch must have been created with ``keep_row_order=True``."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        q_mu = np.ascontiguousarray(q_mu, dtype=np.float64)
        q_sqrt = np.ascontiguousarray(q_sqrt, dtype=np.float64)
        if x.shape != (self.B, self.P) or q_mu.shape != (self.B, self.n) or q_sqrt.shape != (self.B, self.n, self.n):
            raise ValueError("x [B, P], q_mu [B, n], q_sqrt [B, n, n] expected")
        elbo = np.empty(self.B); f = np.empty(self.B); st = np.empty(self.B, np.int32)
        _check(self.lib.wv_batch_eval_elbo(self.handle, _f64(x), _f64(q_mu), _f64(q_sqrt), float(jitter), _f64(elbo), _f64(f),
                                           _i32(st)), "wv_batch_eval_elbo")
        return elbo, f, st

    def fit_adam(self, x0: Optional[np.ndarray] = None, **opts):
        """Batched Adam with the reference's schedule (BaseGP.optimize_params, waveome/model_classes.py:344-462; the
        natural-gradient half is the engine's exact inner maximisation).  Options: learning_rate 0.1, decay_rate 0.96,
        beta1 0.9, beta2 0.999, epsilon 1e-7, convergence_threshold 1e-9, max_iter 50000, check_every 100, decay_every 500.
        Returns dict(x, f, lml, n_iter, n_eval, status)."""
        o = dict(DEFAULT_ADAM); o.update(opts)
        x = self.x0() if x0 is None else np.array(x0, dtype=np.float64, order="C", copy=True)
        if x.shape != (self.B, self.P):
            raise ValueError(f"x0 must be [{self.B}, {self.P}]")
        co = _AdamOpts(float(o["learning_rate"]), float(o["decay_rate"]), float(o["beta1"]), float(o["beta2"]),
                       float(o["epsilon"]), float(o["convergence_threshold"]), int(o["max_iter"]), int(o["check_every"]),
                       int(o["decay_every"]), 0)
        f = np.empty(self.B); lml = np.empty(self.B)
        nit = np.empty(self.B, np.int32); st = np.empty(self.B, np.int32)
        _check(self.lib.wv_batch_fit_adam(self.handle, _f64(x), C.byref(co), _f64(f), _f64(lml), _i32(nit), _i32(st)),
               "wv_batch_fit_adam")
        return dict(x=x, f=f, lml=lml, n_iter=nit, n_eval=nit + 1, status=st)

    def alpha(self) -> np.ndarray:
        """[B, n] alpha = (K + sigma^2 I)^{-1} (y - c) of the last evaluation, in the caller's row order."""
        a = np.empty((self.B, self.n))
        _check(self.lib.wv_batch_get_alpha(self.handle, _f64(a)), "wv_batch_get_alpha")
        return a

    def kinv_diag(self) -> np.ndarray:
        """[B, n] diag((K + sigma^2 I)^-1) of the last evaluation, in the caller's row order."""
        d = np.empty((self.B, self.n))
        _check(self.lib.wv_batch_get_kinv_diag(self.handle, _f64(d)), "wv_batch_get_kinv_diag")
        return d

    def predict_mean(self, Xnew: np.ndarray) -> np.ndarray:
        """[B, m] posterior means at new inputs [m, D] with the parameters of the last evaluation."""
        Xnew = np.ascontiguousarray(Xnew, dtype=np.float64)
        if Xnew.ndim != 2 or Xnew.shape[1] != self.D:
            raise ValueError(f"Xnew must be [m, {self.D}]")
        out = np.empty((self.B, Xnew.shape[0]))
        _check(self.lib.wv_batch_predict_mean(self.handle, _f64(Xnew), int(Xnew.shape[0]), _f64(out)),
               "wv_batch_predict_mean")
        return out

    def predict_f(self, Xnew: np.ndarray):
        """([B, m] mean, [B, m] variance) of f at new inputs [m, D] with the parameters of the last evaluation."""
        Xnew = np.ascontiguousarray(Xnew, dtype=np.float64)
        if Xnew.ndim != 2 or Xnew.shape[1] != self.D:
            raise ValueError(f"Xnew must be [m, {self.D}]")
        mean = np.empty((self.B, Xnew.shape[0])); var = np.empty((self.B, Xnew.shape[0]))
        _check(self.lib.wv_batch_predict_f(self.handle, _f64(Xnew), int(Xnew.shape[0]), _f64(mean), _f64(var)),
               "wv_batch_predict_f")
        return mean, var

    def counters(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.lib.wv_batch_counters(self.handle, C.byref(a), C.byref(b), C.byref(c))
        return dict(launches=a.value, rounds=b.value, model_evals=c.value)

    def profile(self, on: bool = True):
        """Enable/disable CUDA-event timing per kernel class (resets the accumulators when enabling)."""
        self.lib.wv_batch_profile_enable(self.handle, 1 if on else 0)

    def profile_read(self):
        """{class: (total_ms, launches)} since profile(True)."""
        n = len(KERNEL_CLASSES)
        ms = np.zeros(n); cnt = np.zeros(n, np.int64)
        self.lib.wv_batch_profile_read(self.handle, _f64(ms), cnt.ctypes.data_as(C.POINTER(C.c_int64)), n)
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(KERNEL_CLASSES)}

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.wv_batch_workspace_bytes(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.wv_batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
