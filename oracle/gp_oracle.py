"""CPU oracle for the waveome model-fitting hot path (objective A: exact GPR MAP fit).

TEST INFRASTRUCTURE ONLY.  Nothing under ``waveome_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs use it, and only as the checker / reported baseline.

PARITY PINNED ON REFERENCE-RECORDED VALUES for the Gaussian objective (tests/reference_pins.py,
tests/test_reference_pins_cpu.py): outputs of real GPflow runs committed in the reference's notebooks - log marginal
likelihood -9.914289155637 reproduced to all 13 printed digits by this file's fit (SE; also Matern12, Periodic,
Matern52 summaries, Categorical + Matern12 log density to 1e-10, Constant-mean SVGP ELBOs as bounds).
PARITY UNPINNED for the priors (tfd.Horseshoe / Laplace / Uniform: no recorded value exists) - those follow the
published algorithms only.  The reference (omicsEye/waveome v0.1.3) ships no tests or fixtures for this path, and
its arithmetic lives in third-party packages that are not vendored
in /root/reference and cannot be installed here (no network, Python 3.12 > requires-python):

    gpflow==2.9.1, tensorflow>=2.12,<2.16, tensorflow_probability>=0.20,<0.24, scipy>=1.11,<1.13
    (/root/reference/pyproject.toml:29-47)

This file therefore restates the published algorithms of those packages, anchored on the
reference's own call sites and in-repo mirrors:

* GPR log marginal likelihood  - waveome/model_types_DEPR.py:49-56 (PGPR mirror of
  gpflow.models.GPR.log_marginal_likelihood, constructed at waveome/model_fitting.py:150-155)
* custom kernels               - waveome/kernels.py:19-31 (Lin), :56-73 (Poly), :95-117
  (Categorical), :136-139 (Empty)
* GPflow kernels               - SquaredExponential / Matern12/32/52 / Periodic / Linear /
  Constant / Polynomial / Sum / Product (used at waveome/model_search.py:1071-1076,
  waveome/regularization.py:23)
* priors                       - tfd.Horseshoe (waveome/model_classes.py:857), tfd.Laplace
  (waveome/model_fitting.py:201,210), tfd.Uniform (waveome/model_fitting.py:242)
* training loss / optimiser    - gpflow.optimizers.Scipy().minimize(m.training_loss, ...)
  (waveome/model_fitting.py:276-281, waveome/model_classes.py:309-334) -> scipy L-BFGS-B
* BIC                          - waveome/utilities.py:77-95, rounded at model_fitting.py:352-360

Soft pins (notebook cell outputs recorded by the reference, tolerance ~1e-3) are checked in
tests/test_oracle_softpins.py.

Model description (plain JSON-able dicts, produced by ``waveome_b200`` kernels' ``to_spec()``
but defined here independently):

    param  := {"value": float, "trainable": bool,
               "transform": "softplus" | "softplus_shift" | "identity" | "exp",
               "shift": float (softplus_shift only),
               "prior": None | {"type": "horseshoe", "scale": s}
                             | {"type": "laplace", "loc": m, "scale": b}
                             | {"type": "uniform", "low": lo, "high": hi}}
    kernel := {"type": "sum" | "product", "kernels": [kernel, ...]}
            | {"type": <leaf>, "dim": int, "params": {name: param}, ["degree": int]}
    leaf   in squared_exponential, matern12, matern32, matern52, periodic, linear, lin,
              constant, categorical, polynomial, poly, empty
    model  := {"kernel": kernel,
               "likelihood_variance": param,
               "mean": {"type": "zero"} | {"type": "constant", "c": param}}

Trainable parameters are packed depth-first over the kernel tree (per leaf in the order listed
in ``LEAF_PARAMS``), then the likelihood variance, then the mean constant.
"""
from __future__ import annotations

import copy
import math

import numpy as np

LOG2PI = math.log(2.0 * math.pi)

LEAF_PARAMS = {
    "squared_exponential": ("variance", "lengthscales"),
    "matern12": ("variance", "lengthscales"),
    "matern32": ("variance", "lengthscales"),
    "matern52": ("variance", "lengthscales"),
    "periodic": ("variance", "lengthscales", "period"),
    "linear": ("variance",),
    "lin": ("variance",),
    "constant": ("variance",),
    "categorical": ("variance",),
    "polynomial": ("variance", "offset"),
    "poly": ("variance", "offset"),
    "empty": ("variance",),
}


# --------------------------------------------------------------------------------------
# transforms  (GPflow positive() = tfp Softplus; Gaussian likelihood variance = Shift(1e-6)∘Softplus;
#              SURVEY Appendix A.1)
# --------------------------------------------------------------------------------------
def softplus(u):
    u = np.asarray(u, dtype=np.float64)
    return np.maximum(u, 0.0) + np.log1p(np.exp(-np.abs(u)))


def sigmoid(u):
    u = np.asarray(u, dtype=np.float64)
    e = np.exp(-np.abs(u))
    return np.where(u >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


def softplus_inverse(y):
    """tfp.math.softplus_inverse: log(expm1(y)) evaluated stably."""
    y = np.asarray(y, dtype=np.float64)
    return y + np.log(-np.expm1(-y))


def transform_forward(param, u):
    t = param["transform"]
    if t == "softplus":
        return float(softplus(u))
    if t == "softplus_shift":
        return float(softplus(u)) + float(param.get("shift", 1e-6))
    if t == "identity":
        return float(u)
    if t == "exp":
        return float(np.exp(u))
    raise ValueError(f"unknown transform {t}")


def transform_inverse(param, v):
    t = param["transform"]
    if t == "softplus":
        return float(softplus_inverse(v))
    if t == "softplus_shift":
        return float(softplus_inverse(v - float(param.get("shift", 1e-6))))
    if t == "identity":
        return float(v)
    if t == "exp":
        return float(np.log(v))
    raise ValueError(f"unknown transform {t}")


def transform_dtheta_du(param, u):
    t = param["transform"]
    if t in ("softplus", "softplus_shift"):
        return float(sigmoid(u))
    if t == "identity":
        return 1.0
    if t == "exp":
        return float(np.exp(u))
    raise ValueError(f"unknown transform {t}")


# --------------------------------------------------------------------------------------
# priors: log density of the CONSTRAINED value, no Jacobian (GPflow prior_on=CONSTRAINED;
# SURVEY Appendix A.1 / A.6)
# --------------------------------------------------------------------------------------
HS_G = 0.5614594835668851
HS_B = 1.0420764938351215
HS_HINF = 1.0801359952503342
HS_P = 1.0919284281983377


def horseshoe_logp_and_grad(x, scale):
    """tfd.Horseshoe(scale).log_prob(x) (TFP's closed-form approximation) and d/dx."""
    with np.errstate(all="ignore"):
        x = np.float64(x)
        s = np.float64(scale)
        xx = (x / s) ** 2 / 2.0
        g, b, h_inf = HS_G, HS_B, HS_HINF
        q = 20.0 / 47.0 * xx**HS_P
        h = 1.0 / (1.0 + xx**1.5) + h_inf * q / (1.0 + q)
        c = -0.5 * np.log(2.0 * np.pi**3) - np.log(g * s)
        z = np.log1p(-g) - np.log(g)
        t = z - xx / (1.0 - g)
        hb = h + b * xx
        u = g / xx - (1.0 - g) / hb**2
        l1 = np.log1p(u)
        logp = -softplus(t) + np.log(l1) + c
        # derivative wrt xx, then chain dxx/dx = x / s^2
        dA = sigmoid(t) / (1.0 - g)
        dq = HS_P * q / xx
        dh = -1.5 * np.sqrt(xx) / (1.0 + xx**1.5) ** 2 + h_inf * dq / (1.0 + q) ** 2
        du = -g / xx**2 + 2.0 * (1.0 - g) * (dh + b) / hb**3
        dB = du / ((1.0 + u) * l1)
        dlogp = (dA + dB) * x / s**2
    return float(logp), float(dlogp)


def prior_logp_and_grad(prior, v):
    if prior is None:
        return 0.0, 0.0
    t = prior["type"]
    if t == "horseshoe":
        return horseshoe_logp_and_grad(v, prior["scale"])
    if t == "laplace":
        loc, b = float(prior.get("loc", 0.0)), float(prior["scale"])
        return -abs(v - loc) / b - math.log(2.0 * b), -math.copysign(1.0, v - loc) / b if v != loc else 0.0
    if t == "uniform":
        lo, hi = float(prior["low"]), float(prior["high"])
        if lo <= v <= hi:
            return -math.log(hi - lo), 0.0
        return -math.inf, 0.0
    raise ValueError(f"unknown prior {t}")


# --------------------------------------------------------------------------------------
# kernels  
# --------------------------------------------------------------------------------------
def _col(X, d):
    return np.ascontiguousarray(X[:, int(d)], dtype=np.float64)


def _scaled_sqdist(x, ell):
    """gpflow.utilities.ops.square_distance(X/ell, None) for one column:
    -2 a a^T + |a_i|^2 + |a_j|^2 (not clamped)."""
    a = x / ell
    a2 = a * a
    return -2.0 * np.outer(a, a) + (a2[:, None] + a2[None, :])


def _leaf_K_and_grads(node, X, want_grads):
    """Returns K (n x n) and {param_name: dK/dtheta} for a leaf."""
    typ = node["type"]
    P = node["params"]
    x = _col(X, node.get("dim", 0))
    n = x.shape[0]
    g = {}
    if typ == "empty":
        K = np.zeros((n, n))
        if want_grads:
            g["variance"] = np.zeros((n, n))
        return K, g
    var = P["variance"]["value"]
    if typ == "squared_exponential":
        ell = P["lengthscales"]["value"]
        r2 = _scaled_sqdist(x, ell)
        E = np.exp(-0.5 * r2)
        K = var * E
        if want_grads:
            g["variance"] = E
            g["lengthscales"] = K * r2 / ell
    elif typ in ("matern12", "matern32", "matern52"):
        ell = P["lengthscales"]["value"]
        r2 = _scaled_sqdist(x, ell)
        r = np.sqrt(np.maximum(r2, 1e-36))
        # dr/dell = -r/ell (where not clamped; clamped entries have zero gradient in TF's
        # maximum(), which picks the constant branch) -- those entries have r ~ 1e-18 anyway.
        live = r2 > 1e-36
        if typ == "matern12":
            E = np.exp(-r)
            dE_dr = -E
        elif typ == "matern32":
            s3 = math.sqrt(3.0)
            E = (1.0 + s3 * r) * np.exp(-s3 * r)
            dE_dr = -3.0 * r * np.exp(-s3 * r)
        else:
            s5 = math.sqrt(5.0)
            E = (1.0 + s5 * r + 5.0 / 3.0 * r * r) * np.exp(-s5 * r)
            dE_dr = -(5.0 / 3.0) * r * (1.0 + s5 * r) * np.exp(-s5 * r)
        K = var * E
        if want_grads:
            g["variance"] = E
            g["lengthscales"] = np.where(live, var * dE_dr * (-r / ell), 0.0)
    elif typ == "periodic":
        # gpflow.kernels.Periodic(base=SquaredExponential): var*exp(-0.5*(sin(pi (x-x')/p)/ell)^2)
        ell = P["lengthscales"]["value"]
        per = P["period"]["value"]
        diff = x[:, None] - x[None, :]
        arg = np.pi * diff / per
        sn = np.sin(arg)
        ss = sn / ell
        r2 = ss * ss
        E = np.exp(-0.5 * r2)
        K = var * E
        if want_grads:
            g["variance"] = E
            g["lengthscales"] = K * r2 / ell
            # d r2/d per = 2 ss * cos(arg)/ell * (-arg/per)
            g["period"] = K * (ss * np.cos(arg) / ell) * (arg / per)
    elif typ in ("linear", "lin"):
        xx = np.outer(x, x)
        K = var * xx
        if want_grads:
            g["variance"] = xx
    elif typ == "constant":
        K = np.full((n, n), float(var))
        if want_grads:
            g["variance"] = np.ones((n, n))
    elif typ == "categorical":
        # waveome/kernels.py:109-117: equality of int64(round(x)); tf.round = half-to-even = np.rint
        c = np.rint(x).astype(np.int64)
        M = (c[:, None] == c[None, :]).astype(np.float64)
        K = var * M
        if want_grads:
            g["variance"] = M
    elif typ in ("polynomial", "poly"):
        deg = int(node.get("degree", 3))
        off = P["offset"]["value"]
        xx = np.outer(x, x)
        base = var * xx + off
        K = base**deg
        if want_grads:
            dbase = deg * base ** (deg - 1)
            g["variance"] = dbase * xx
            g["offset"] = dbase
    else:
        raise ValueError(f"unknown kernel type {typ}")
    return K, g


def kernel_K_and_grads(node, X, want_grads=True):
    """Evaluate a kernel tree.  Returns (K, [(param_dict, dK/dtheta), ...]) where the list
    covers every parameter (trainable or not) in depth-first order."""
    typ = node["type"]
    if typ == "sum":
        K = None
        grads = []
        for ch in node["kernels"]:
            Kc, gc = kernel_K_and_grads(ch, X, want_grads)
            K = Kc if K is None else K + Kc
            grads += gc
        return K, grads
    if typ == "product":
        parts = [kernel_K_and_grads(ch, X, want_grads) for ch in node["kernels"]]
        K = None
        for Kc, _ in parts:
            K = Kc if K is None else K * Kc
        grads = []
        if want_grads:
            for i, (_, gc) in enumerate(parts):
                others = None
                for j, (Kj, _) in enumerate(parts):
                    if j != i:
                        others = Kj if others is None else others * Kj
                for (p, dK) in gc:
                    grads.append((p, dK if others is None else dK * others))
        return K, grads
    K, g = _leaf_K_and_grads(node, X, want_grads)
    grads = [(node["params"][name], g[name]) for name in LEAF_PARAMS[typ]] if want_grads else []
    return K, grads


def iter_kernel_params(node):
    """Depth-first parameter dicts of a kernel tree (same order as kernel_K_and_grads)."""
    if node["type"] in ("sum", "product"):
        for ch in node["kernels"]:
            yield from iter_kernel_params(ch)
    else:
        for name in LEAF_PARAMS[node["type"]]:
            yield node["params"][name]


def iter_model_params(model):
    yield from iter_kernel_params(model["kernel"])
    yield model["likelihood_variance"]
    if "likelihood_aux" in model:          # second likelihood parameter (ZINB km), packed right after the first
        yield model["likelihood_aux"]
    if model["mean"]["type"] == "constant":
        yield model["mean"]["c"]


def trainable_params(model):
    return [p for p in iter_model_params(model) if p.get("trainable", True)]


def pack(model):
    """Unconstrained start vector of the trainable parameters."""
    return np.array([transform_inverse(p, p["value"]) for p in trainable_params(model)], dtype=np.float64)


def unpack(model, x):
    """Write unconstrained vector x into the model's parameter values (in place)."""
    for p, u in zip(trainable_params(model), x):
        p["value"] = transform_forward(p, float(u))
    return model


class CholeskyFailure(Exception):
    """K + sigma^2 I was not numerically positive definite (TF raises InvalidArgumentError)."""


def objective(model, X, y, x, want_grad=True):
    """GPflow ``GPR.training_loss`` at unconstrained x.

    Returns (f, grad, lml, log_prior) with f = -(lml + log_prior); grad is d f / d x.
    Follows waveome/model_types_DEPR.py:49-56 + gpflow.logdensities.multivariate_normal.
    """
    with np.errstate(all="ignore"):
        return _objective(model, X, y, x, want_grad)


def _objective(model, X, y, x, want_grad):
    model = unpack(model, x)
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n = y.shape[0]
    K, kgrads = kernel_K_and_grads(model["kernel"], X, want_grads=want_grad)
    s2 = model["likelihood_variance"]["value"]
    Ks = K + s2 * np.eye(n)
    tp = trainable_params(model)
    if not np.all(np.isfinite(Ks)):
        # TensorFlow's CPU Cholesky (Eigen LLT) only fails on a pivot <= 0; NaN/inf entries flow through
        # silently and the loss and gradients come back non-finite.  Same here.
        nan = float("nan")
        return nan, (np.full(len(tp), nan) if want_grad else None), nan, nan
    try:
        L = np.linalg.cholesky(Ks)
    except np.linalg.LinAlgError as e:
        raise CholeskyFailure(str(e))
    c = model["mean"]["c"]["value"] if model["mean"]["type"] == "constant" else 0.0
    d = y - c
    import scipy.linalg as sla

    a = sla.solve_triangular(L, d, lower=True)
    lml = -0.5 * float(a @ a) - 0.5 * n * LOG2PI - float(np.sum(np.log(np.diag(L))))
    log_prior = 0.0
    prior_grads = []
    for p in tp:
        lp, dlp = prior_logp_and_grad(p.get("prior"), p["value"])
        log_prior += lp
        prior_grads.append(dlp)
    f = -(lml + log_prior)
    if not want_grad:
        return f, None, lml, log_prior
    alpha = sla.solve_triangular(L, a, lower=True, trans="T")
    Linv = sla.solve_triangular(L, np.eye(n), lower=True)
    Kinv = Linv.T @ Linv
    W = np.outer(alpha, alpha) - Kinv
    # d lml / d theta for every parameter object (by identity)
    dl = {}
    for (p, dK) in kgrads:
        dl[id(p)] = dl.get(id(p), 0.0) + 0.5 * float(np.sum(W * dK))
    dl[id(model["likelihood_variance"])] = 0.5 * float(np.trace(W))
    if model["mean"]["type"] == "constant":
        dl[id(model["mean"]["c"])] = float(np.sum(alpha))
    grad = np.empty(len(tp))
    for i, (p, u) in enumerate(zip(tp, x)):
        grad[i] = -(dl[id(p)] + prior_grads[i]) * transform_dtheta_du(p, float(u))
    return f, grad, lml, log_prior


# --------------------------------------------------------------------------------------
# fit: gpflow.optimizers.Scipy().minimize(...) == scipy.optimize.minimize(method="L-BFGS-B", jac=True)
# (SURVEY Appendix A.7).  SciPy's L-BFGS-B is the optimiser oracle.
# --------------------------------------------------------------------------------------
STATUS_OK = 0
STATUS_CHOL_FAIL = 1
STATUS_NONFINITE = 2
STATUS_MAXITER = 4
STATUS_LINESEARCH = 8


def fit(model, X, y, maxiter=50000, maxfun=None, x0=None, maxcor=10, ftol=2.220446049250313e-09, gtol=1e-05, maxls=20,
        on_chol_fail="nan"):
    """L-BFGS-B MAP fit.  Returns dict(x, f, lml, nit, nfev, status, message, model).

    on_chol_fail: what a Cholesky failure at a line-search trial point does.
      "abort" - the reference's behaviour: TensorFlow raises, GPflow's minimize is abandoned with the parameters left
                at the failing trial point (waveome/model_classes.py:323-340 then retries from there and fails again).
      "nan"   - (default, and the engine's default) the trial is reported to L-BFGS-B as a non-finite value, the line
                search backs out and the last finite iterate is kept.  Whether such a trial "fails" at all is a rounding
                coin-flip (e.g. variance 1e10 + noise 1e-6), so this is the only policy that is reproducible across
                LAPACK / Eigen / CUDA factorizations.  A failure at the start point always aborts."""
    import scipy.optimize as so

    model = copy.deepcopy(model)
    if x0 is None:
        x0 = pack(model)
    if maxfun is None:
        maxfun = 15000
    state = {"n": 0}

    def fun(x):
        state["n"] += 1
        try:
            f, g, _, _ = objective(model, X, y, x)
        except CholeskyFailure:
            if on_chol_fail == "abort" or state["n"] == 1:
                raise
            f, g = float("nan"), np.full(len(x), float("nan"))
        return f, g

    try:
        res = so.minimize(
            fun, np.array(x0, dtype=np.float64), jac=True, method="L-BFGS-B",
            options=dict(maxiter=maxiter, maxfun=maxfun, maxcor=maxcor, ftol=ftol, gtol=gtol, maxls=maxls),
        )
    except CholeskyFailure:
        return dict(x=None, f=math.inf, lml=-math.inf, nit=0, nfev=0, status=STATUS_CHOL_FAIL,
                    message="cholesky failed", model=None)
    unpack(model, res.x)
    try:
        f, _, lml, lp = objective(model, X, y, res.x, want_grad=False)
    except CholeskyFailure:
        return dict(x=res.x, f=math.inf, lml=-math.inf, nit=int(res.nit), nfev=int(res.nfev), status=STATUS_CHOL_FAIL,
                    message="cholesky failed at the returned point", model=model)
    # f is the objective AT THE RETURNED x (what log_posterior_density() reads after the optimiser returns,
    # waveome/model_search.py:2311, and what wv_batch_fit_lbfgs reports): after an ABNORMAL line-search exit SciPy's
    # res.fun still holds the last (possibly non-finite) trial value while res.x is the restored iterate.
    status = STATUS_OK
    if not np.isfinite(f):
        status |= STATUS_NONFINITE
    if res.status == 1:
        status |= STATUS_MAXITER
    if res.status == 2:
        status |= STATUS_LINESEARCH
    return dict(x=res.x, f=float(f), lml=lml, log_prior=lp, nit=int(res.nit), nfev=int(res.nfev),
                status=status, message=str(res.message), model=model)


def count_trainable_parameter_objects(model):
    """k in waveome's "BIC": len(model.trainable_parameters) (Parameter *objects*)."""
    return len(trainable_params(model))


def calc_bic(loglik, n, k):
    """waveome/utilities.py:77-95."""
    return 2 * k - 2 * loglik


def log_posterior_density(model, X, y):
    """gpflow GPModel.log_posterior_density() = lml + log_prior at the model's current values."""
    x = pack(model)
    _, _, lml, lp = objective(copy.deepcopy(model), X, y, x, want_grad=False)
    return lml + lp


def predict_f(model, X, y, Xnew):
    """gpflow GPR.predict_f(Xnew, full_cov=False) at the model's current parameter values: (mean [m], var [m])."""
    import scipy.linalg as sla
    X = np.asarray(X, dtype=np.float64)
    Xnew = np.asarray(Xnew, dtype=np.float64)
    n = X.shape[0]
    Kall, _ = kernel_K_and_grads(model["kernel"], np.vstack([Xnew, X]), want_grads=False)
    m = Xnew.shape[0]
    Kss, Ksx, Kxx = Kall[:m, :m], Kall[:m, m:], Kall[m:, m:]
    s2 = model["likelihood_variance"]["value"]
    c = model["mean"]["c"]["value"] if model["mean"]["type"] == "constant" else 0.0
    L = np.linalg.cholesky(Kxx + s2 * np.eye(n))
    A = sla.solve_triangular(L, Ksx.T, lower=True)                       # [n, m]
    v = sla.solve_triangular(L, np.asarray(y, dtype=np.float64).reshape(-1) - c, lower=True)
    return c + A.T @ v, np.diag(Kss) - np.sum(A * A, axis=0)


def predict_log_density(model, X, y, Xnew, ynew):
    """gpflow GPModel.predict_log_density((Xnew, ynew)) for the Gaussian likelihood: log N(ynew; mean, var_f + sigma^2)
    per point (what waveome/model_classes.py:930-945 averages over the held-out rows)."""
    mu, var = predict_f(model, X, y, Xnew)
    vy = var + model["likelihood_variance"]["value"]
    ynew = np.asarray(ynew, dtype=np.float64).reshape(-1)
    return -0.5 * (LOG2PI + np.log(vy) + (ynew - mu) ** 2 / vy)


# --------------------------------------------------------------------------------------
# small builders used by tests / bench
# --------------------------------------------------------------------------------------
def P(value=1.0, trainable=True, transform="softplus", prior=None, shift=None):
    d = {"value": float(value), "trainable": bool(trainable), "transform": transform, "prior": prior}
    if transform == "softplus_shift":
        d["shift"] = 1e-6 if shift is None else float(shift)
    return d


def leaf(typ, dim, **vals):
    params = {}
    for name in LEAF_PARAMS[typ]:
        v = vals.get(name, 1.0)
        params[name] = v if isinstance(v, dict) else P(v)
    if typ == "empty":
        params["variance"] = P(1e-6, trainable=False)
    node = {"type": typ, "dim": int(dim), "params": params}
    if typ in ("polynomial", "poly"):
        node["degree"] = int(vals.get("degree", 3))
    return node


def gpr_model(kernel, noise=1.0, mean="constant", c=0.0):
    m = {"kernel": kernel, "likelihood_variance": P(noise, transform="softplus_shift")}
    m["mean"] = {"type": "constant", "c": P(c, transform="identity")} if mean == "constant" else {"type": "zero"}
    return m
