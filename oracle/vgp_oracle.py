"""CPU oracle for the count-outcome path (BASELINE configs[4]: Poisson / negative-binomial outcomes via the
variational GP).  TEST INFRASTRUCTURE ONLY — same rules as gp_oracle.py (nothing under ``waveome_b200/`` imports it).

PARITY: soft-pinned for the Bernoulli likelihood on a reference-recorded VarGP run (tests/reference_pins.py:
examples/simulations/simple_regression_different_models.ipynb cells 9-11 - recorded loss 22.81370 vs this file's
max over q 22.81297 at the recorded hyper-parameters, q_mu[0] / q_sqrt[0,0] to 3 digits); UNPINNED for Poisson, negative
binomial and Gamma (no recorded value exists).  The reference builds ``gpflow.models.VGP`` (waveome/model_fitting.py:158-185) or the SVGP-with-Z=X
equivalent ``PSVGP`` (waveome/model_classes.py:1082-1126) with ``gpflow.likelihoods.Poisson`` or
``waveome.likelihoods.NegativeBinomial`` (waveome/likelihoods.py:16-79); GPflow 2.9.1 is not vendored and cannot be
installed here, and the reference has no tests.  Restated from the published algorithm:

    K = k(X) + 1e-6 I,  L = chol(K),  f_mean = L q_mu + c,  f_var = rowsum((L q_sqrt)^2)
    ELBO = sum_i E_{N(f_i; f_mean_i, f_var_i)}[log p(y_i | f_i)] - KL[N(q_mu, q_sqrt q_sqrt^T) || N(0, I)]
    Poisson (exp link): E = y mu - exp(mu + v/2) - lgamma(y + 1)
    NegativeBinomial:   20-point Gauss-Hermite of waveome/likelihoods.py:68-79 with m = exp(f)

``vgp_elbo`` is that objective as a function of (theta, q_mu, q_sqrt) — what the reference hands to L-BFGS.
``vgp_collapsed`` is the same objective maximised over (q_mu, q_sqrt) for fixed theta: for a factorising likelihood
the optimal q is the posterior of a GP with Gaussian pseudo-observations ytilde_i of precision lam_i ("sites"),

    S = (K^-1 + Lam)^-1,  m = c + K alpha,  alpha = (K + Lam^-1)^-1 (ytilde - c)
    fixed point:  lam_i = -2 dE_i/dv_i,   alpha_i = dE_i/dm_i
    F(theta) = log N(ytilde; c, K + Lam^-1) + sum_i [E_i + 1/2 log(2 pi / lam_i) + lam_i/2 ((ytilde_i - m_i)^2 + v_i)]
    dF/dtheta = 1/2 tr((alpha alpha^T - (K + Lam^-1)^-1) dK/dtheta),  dF/dc = sum(alpha)      (envelope theorem)

which is the form the engine evaluates (each fixed-point sweep is one heteroscedastic GPR factorisation).  The tests
check F == max_q ELBO (stationarity of vgp_elbo at the q built from the sites, by torch autograd) and dF/dtheta against
finite differences, before the engine is compared with F.
"""
from __future__ import annotations

import copy
import math

import numpy as np

import gp_oracle as go

JITTER = 1e-6            # gpflow.config.default_jitter()
GH_X, GH_W = np.polynomial.hermite.hermgauss(20)


def lgamma(x):
    from scipy.special import gammaln
    return gammaln(x)


# --------------------------------------------------------------------------------------------------
# likelihoods: E_q[log p(y|f)] for f ~ N(m, v), and its derivatives wrt m and v
# --------------------------------------------------------------------------------------------------
def nb_logpmf(f, y, alpha):
    """waveome/likelihoods.py:68-79 with m = exp(f) (log link), k = 1/alpha."""
    k = 1.0 / alpha
    m = np.exp(f)
    return lgamma(k + y) - lgamma(y + 1) - lgamma(k) + y * np.log(m / (m + k)) - k * np.log1p(m * alpha)


def nb_dalpha(y, m, v, a):
    """sum_i dE_i/d(alpha) for the negative binomial (20-point Gauss-Hermite of d log p / d alpha)."""
    from scipy.special import digamma
    y = np.asarray(y, dtype=np.float64)
    k = 1.0 / a
    f = m[:, None] + np.sqrt(2.0 * v)[:, None] * GH_X[None, :]
    w = GH_W[None, :] / math.sqrt(math.pi)
    ef = np.exp(f)
    dk = (digamma(k + y) - digamma(k))[:, None] - y[:, None] / (ef + k) - np.log1p(ef * a) + ef / (k + ef)
    return float(np.sum(w * dk) * (-k * k))


def zinb_terms(f, y, alpha, km):
    """waveome/likelihoods.py:96-139 (ZeroInflatedNegativeBinomial, exp link): log p(y | f) and its derivatives wrt f
    (first, second), alpha and km.  psi = km / (km + m) is the structural-zero probability, m = exp(f)."""
    from scipy.special import digamma
    m = np.exp(f)
    Q = km + m
    zero = (y == 0)
    # y > 0: log(1 - psi) + NB(m, y, alpha) = f - log(km + m) + NB
    k = 1.0 / alpha
    lp_nz = f - np.log(Q) + nb_logpmf(f, y, alpha)
    d1_nz = km / Q + y - (y + k) * alpha * m / (1.0 + alpha * m)
    d2_nz = -km * m / Q ** 2 - (y + k) * alpha * m / (1.0 + alpha * m) ** 2
    dk = digamma(k + y) - digamma(k) - y / (m + k) - np.log1p(m * alpha) + m / (k + m)
    da_nz = dk * (-k * k)
    dkm_nz = -1.0 / Q
    # y = 0: log(psi + (1 - psi) (1 + alpha m)^(-1/alpha)) = log(km + u) - log(km + m), u = m (1 + alpha m)^(-1/alpha)
    ls0 = -np.log1p(alpha * m) / alpha
    u = m * np.exp(ls0)
    P = km + u
    r = (1.0 + (alpha - 1.0) * m) / (1.0 + alpha * m)            # d log u / df
    u1 = u * r
    u2 = u * (r * r - m / (1.0 + alpha * m) ** 2)
    lp_z = np.log(P) - np.log(Q)
    d1_z = u1 / P - m / Q
    d2_z = u2 / P - (u1 / P) ** 2 - km * m / Q ** 2
    da_z = u * (np.log1p(alpha * m) / alpha ** 2 - m / (alpha * (1.0 + alpha * m))) / P
    dkm_z = 1.0 / P - 1.0 / Q
    pick = lambda a, b: np.where(zero, a, b)
    return pick(lp_z, lp_nz), pick(d1_z, d1_nz), pick(d2_z, d2_nz), pick(da_z, da_nz), pick(dkm_z, dkm_nz)


def zinb_dparams(y, m, v, alpha, km):
    """(sum_i dE_i/d alpha, sum_i dE_i/d km) by the same quadrature"""
    y = np.asarray(y, dtype=np.float64)
    f = m[:, None] + np.sqrt(2.0 * v)[:, None] * GH_X[None, :]
    w = GH_W[None, :] / math.sqrt(math.pi)
    _, _, _, da, dkm = zinb_terms(f, y[:, None], alpha, km)
    return float(np.sum(w * da)), float(np.sum(w * dkm))


# Smallest site precision: the ZINB zero branch is not log-concave in f, so the unconstrained optimum of q can have
# negative site precisions, which the heteroscedastic-GPR form (K + 1/lam) cannot carry.  The site iteration projects the
# precisions onto >= LAM_MIN: the value is the ELBO of the resulting q (a valid lower bound of the evidence, equal to
# the VGP optimum when no site sits at the bound); with sites at the bound q is not a stationary point of the ELBO and
# the envelope-theorem gradient is approximate.  The log-concave likelihoods never reach the bound.
LAM_MIN = {"zinb": 1e-6}


def lik_of(model, lik):
    """The likelihood dict with alpha taken from the model when it carries the dispersion in its (unused) Gaussian
    noise slot: spec["likelihood_variance"] with an exp transform (how the product's NegativeBinomial is encoded)."""
    lv = model["likelihood_variance"]
    if lik["type"] == "negative_binomial" and lv.get("transform") == "exp":
        return {"type": "negative_binomial", "alpha": lv["value"]}
    if lik["type"] == "gamma" and lv.get("transform") == "softplus":
        return {"type": "gamma", "shape": lv["value"]}
    if lik["type"] == "zinb" and "likelihood_aux" in model:
        return {"type": "zinb", "alpha": lv["value"], "km": model["likelihood_aux"]["value"]}
    return lik


def var_exp(lik, y, m, v):
    """(E, dE/dm, dE/dv) per observation."""
    y = np.asarray(y, dtype=np.float64)
    if lik["type"] == "gaussian":         # gpflow.likelihoods.Gaussian: used to show objective (B) == objective (A)
        s2 = lik["variance"]
        return -0.5 * np.log(2.0 * np.pi * s2) - ((y - m) ** 2 + v) / (2.0 * s2), (y - m) / s2, np.full_like(m, -0.5 / s2)
    if lik["type"] == "poisson":
        r = np.exp(m + 0.5 * v)
        return y * m - r - lgamma(y + 1.0), y - r, -0.5 * r
    if lik["type"] == "gamma":            # gpflow.likelihoods.Gamma, exp link, closed form
        a = lik["shape"]
        r = y * np.exp(-m + 0.5 * v)
        return -a * m - lgamma(a) + (a - 1.0) * np.log(y) - r, -a + r, -0.5 * r
    if lik["type"] == "bernoulli":        # gpflow.likelihoods.Bernoulli, inv_probit link, 20-point Gauss-Hermite
        from scipy.special import erfc
        f = m[:, None] + np.sqrt(2.0 * v)[:, None] * GH_X[None, :]
        w = GH_W[None, :] / math.sqrt(math.pi)
        p = 1e-3 + (1.0 - 2e-3) * 0.5 * erfc(-f / math.sqrt(2.0))
        ph = (1.0 - 2e-3) * np.exp(-0.5 * f * f) / math.sqrt(2.0 * math.pi)
        yy = y[:, None] > 0.5
        pp = np.where(yy, p, 1.0 - p)
        d1 = np.where(yy, ph, -ph) / pp
        d2 = -f * d1 - d1 * d1
        return np.sum(w * np.log(pp), 1), np.sum(w * d1, 1), 0.5 * np.sum(w * d2, 1)
    if lik["type"] == "zinb":
        f = m[:, None] + np.sqrt(2.0 * v)[:, None] * GH_X[None, :]
        w = GH_W[None, :] / math.sqrt(math.pi)
        lp, d1, d2, _, _ = zinb_terms(f, y[:, None], lik["alpha"], lik["km"])
        return np.sum(w * lp, 1), np.sum(w * d1, 1), 0.5 * np.sum(w * d2, 1)
    if lik["type"] == "negative_binomial":
        a = lik["alpha"]
        sd = np.sqrt(2.0 * v)
        f = m[:, None] + sd[:, None] * GH_X[None, :]
        w = GH_W[None, :] / math.sqrt(math.pi)
        lp = nb_logpmf(f, y[:, None], a)
        # d log p / df = y - (y + 1/a) a e^f / (1 + a e^f);  d2 = -(y + 1/a) a e^f / (1 + a e^f)^2
        ef = np.exp(f)
        d1 = y[:, None] - (y[:, None] + 1.0 / a) * a * ef / (1.0 + a * ef)
        d2 = -(y[:, None] + 1.0 / a) * a * ef / (1.0 + a * ef) ** 2
        return np.sum(w * lp, 1), np.sum(w * d1, 1), 0.5 * np.sum(w * d2, 1)      # Bonnet / Price
    raise ValueError(lik["type"])


# --------------------------------------------------------------------------------------------------
# the reference objective: whitened VGP ELBO
# --------------------------------------------------------------------------------------------------
def _kernel_and_mean(model, X, x):
    model = copy.deepcopy(model)
    go.unpack(model, x)
    K, _ = go.kernel_K_and_grads(model["kernel"], X, want_grads=False)
    c = model["mean"]["c"]["value"] if model["mean"]["type"] == "constant" else 0.0
    return model, K + JITTER * np.eye(len(X)), c


def vgp_elbo(model, lik, X, y, x, q_mu, q_sqrt):
    """gpflow.models.VGP.elbo (whitened), without the log prior of the hyper-parameters (``vgp_collapsed`` reports it)."""
    model, K, c = _kernel_and_mean(model, X, x)
    lik = lik_of(model, lik)
    L = np.linalg.cholesky(K)
    fm = L @ q_mu + c
    LS = L @ np.tril(q_sqrt)
    fv = np.sum(LS * LS, 1)
    E, _, _ = var_exp(lik, y, fm, fv)
    n = len(y)
    kl = 0.5 * (np.sum(q_mu ** 2) + np.sum(np.tril(q_sqrt) ** 2) - n - 2.0 * np.sum(np.log(np.abs(np.diag(q_sqrt)))))
    return float(np.sum(E) - kl)


# --------------------------------------------------------------------------------------------------
# the collapsed form the engine evaluates
# --------------------------------------------------------------------------------------------------
def vgp_collapsed(model, lik, X, y, x, sites=None, tol=1e-12, maxit=500, rho=1.0, want_grad=True,
                  exact_bound_gradient=False):
    """F(theta) = max_q ELBO, its gradient wrt the packed unconstrained hyper-parameters, and the converged sites.

    model: gp_oracle model dict WITHOUT a Gaussian likelihood variance being trainable (it is ignored);
    x: packed unconstrained (kernel params..., mean).  sites = (lam, lam * ytilde) to warm-start."""
    spec = copy.deepcopy(model)
    go.unpack(spec, x)
    lik = lik_of(spec, lik)
    n = len(y)
    K, dKs = go.kernel_K_and_grads(spec["kernel"], X, want_grads=want_grad)
    K = K + JITTER * np.eye(n)
    c = spec["mean"]["c"]["value"] if spec["mean"]["type"] == "constant" else 0.0
    y = np.asarray(y, dtype=np.float64)
    if sites is None:
        lam = np.ones(n)                       # unit-precision pseudo-observations at a link-scale guess of f
        if lik["type"] == "bernoulli":
            eta = np.where(y > 0.5, 1.0, -1.0)
        elif lik["type"] == "gamma":
            eta = np.log(np.maximum(y, 1e-12))
        else:
            eta = np.log(y + 1.0)
    else:
        lam, eta = [np.array(s, dtype=np.float64) for s in sites]
    it = 0
    for it in range(1, maxit + 1):
        D = 1.0 / lam
        yt = eta / lam
        A = K + np.diag(D)
        Ai = np.linalg.inv(A)
        alpha = Ai @ (yt - c)
        m = yt - D * alpha
        v = D - D * D * np.diag(Ai)
        E, g, h = var_exp(lik, y, m, v)
        lam_t = np.maximum(-2.0 * h, LAM_MIN.get(lik["type"], 1e-300))
        eta_t = g + lam_t * m
        lam_n = (1.0 - rho) * lam + rho * lam_t
        eta_n = (1.0 - rho) * eta + rho * eta_t
        delta = max(np.max(np.abs(lam_n - lam) / (np.abs(lam) + 1e-300)), np.max(np.abs(eta_n - eta)) / (1.0 + np.max(np.abs(eta))))
        lam, eta = lam_n, eta_n
        if delta < tol:
            break
    D = 1.0 / lam
    yt = eta / lam
    A = K + np.diag(D)
    Lc = np.linalg.cholesky(A)
    z = np.linalg.solve(Lc, yt - c)
    alpha = np.linalg.solve(Lc.T, z)
    Ai = np.linalg.inv(A)
    m = yt - D * alpha
    v = D - D * D * np.diag(Ai)
    E, g, h = var_exp(lik, y, m, v)
    logZ = -0.5 * z @ z - np.sum(np.log(np.diag(Lc))) - 0.5 * n * go.LOG2PI
    F = logZ + np.sum(E + 0.5 * np.log(2.0 * np.pi / lam) + 0.5 * lam * ((yt - m) ** 2 + v))
    # hyper-parameter priors (gpflow log_prior_density; waveome/model_fitting.py:236-242 puts Uniform(0, 10) on the
    # non-variance kernel parameters of the VGP branches): F stays the bound, f = -(F + log prior) is the MAP objective
    log_prior, prior_grads = 0.0, []
    for p in go.trainable_params(spec):
        lp, dlp = go.prior_logp_and_grad(p.get("prior"), p["value"])
        log_prior += lp
        prior_grads.append(dlp)
    out = dict(F=float(F), log_prior=float(log_prior), f=-(float(F) + float(log_prior)), sites=(lam, eta), m=m, v=v,
               iters=it, alpha=alpha)
    if want_grad:
        W = np.outer(alpha, alpha) - Ai
        # Sites at the lower precision bound (ZINB only) are not stationary in their variance: at fixed sites
        # dv_i/dtheta = D_i^2 (A^-1 dK A^-1)_ii contributes (h_i + lam_i / 2) dv_i/dtheta, i.e. W += 2 A^-1 C A^-1.
        # Optional and NOT what the engine evaluates: it removes most of the kernel-parameter error (30 % -> 0.3 % in
        # tests/test_vgp_gpu.py's bounded case) but the free sites then solve the unconstrained fixed-point equations
        # rather than the constrained stationarity conditions, so the envelope argument stays approximate either way.
        bound = (-2.0 * h < LAM_MIN.get(lik["type"], 0.0))
        if exact_bound_gradient and np.any(bound):
            C = np.where(bound, (h + 0.5 * lam) * D * D, 0.0)
            W = W + 2.0 * (Ai * C[None, :]) @ Ai
        grads = []
        # same packing as gp_oracle.pack: kernel params depth-first, (likelihood variance), mean
        tp = go.trainable_params(spec)
        dl = {}
        for (p, dK) in dKs:
            dl[id(p)] = dl.get(id(p), 0.0) + 0.5 * float(np.sum(W * dK))
        for p, u in zip(tp, x):
            if id(p) in dl:
                dth = dl[id(p)]
            elif spec["mean"]["type"] == "constant" and p is spec["mean"]["c"]:
                dth = np.sum(alpha)
            elif p is spec["likelihood_variance"] and lik["type"] == "negative_binomial":
                dth = nb_dalpha(y, m, v, lik["alpha"])          # the slot carries the NB dispersion
            elif p is spec["likelihood_variance"] and lik["type"] == "zinb":
                dth = zinb_dparams(y, m, v, lik["alpha"], lik["km"])[0]
            elif p is spec.get("likelihood_aux") and lik["type"] == "zinb":
                dth = zinb_dparams(y, m, v, lik["alpha"], lik["km"])[1]
            elif p is spec["likelihood_variance"] and lik["type"] == "gamma":
                from scipy.special import digamma
                dth = float(np.sum(-m - digamma(lik["shape"]) + np.log(y)))
            else:
                dth = 0.0                      # the Gaussian noise variance does not exist on this path
            grads.append((dth + prior_grads[len(grads)]) * go.transform_dtheta_du(p, u))
        out["grad"] = np.array(grads)          # d(F + log prior) / du
    return out


def q_from_sites(model, X, y, x, sites):
    """(q_mu, q_sqrt) of the whitened parameterisation for the Gaussian defined by the sites."""
    spec, K, c = _kernel_and_mean(model, X, x)
    lam, eta = sites
    L = np.linalg.cholesky(K)
    S = np.linalg.inv(np.linalg.inv(K) + np.diag(lam))
    A = K + np.diag(1.0 / lam)
    mf = c + K @ np.linalg.solve(A, eta / lam - c)
    q_mu = np.linalg.solve(L, mf - c)
    Sv = np.linalg.solve(L, np.linalg.solve(L, S).T)
    Sv = 0.5 * (Sv + Sv.T)
    return q_mu, np.linalg.cholesky(Sv)


def predict_f(model, X, y, x, sites, Xnew, whitened=False):
    """gpflow VGP.predict_f(Xnew) (full_cov=False) at the q defined by the sites: (mean [m], variance [m]).
    Default: the site form the engine uses, mean = c + K*^T (K + D)^-1 (ytilde - c), var = k** - K*^T (K + D)^-1 K*;
    ``whitened=True``: gpflow's conditional with (q_mu, q_sqrt), mean = K*^T L^-T q_mu + c,
    var = k** - |L^-1 K*|^2 + |q_sqrt^T L^-1 K*|^2 -- the two agree (tests/test_vgp_oracle.py)."""
    spec = copy.deepcopy(model)
    go.unpack(spec, x)
    n = len(X)
    Kall, _ = go.kernel_K_and_grads(spec["kernel"], np.vstack([X, Xnew]), want_grads=False)
    K = Kall[:n, :n] + JITTER * np.eye(n)
    Ks, kss = Kall[:n, n:], np.diag(Kall)[n:]
    c = spec["mean"]["c"]["value"] if spec["mean"]["type"] == "constant" else 0.0
    lam, eta = sites
    if not whitened:
        A = K + np.diag(1.0 / lam)
        return c + Ks.T @ np.linalg.solve(A, eta / lam - c), kss - np.sum(Ks * np.linalg.solve(A, Ks), 0)
    q_mu, q_sqrt = q_from_sites(model, X, y, x, sites)
    L = np.linalg.cholesky(K)
    B = np.linalg.solve(L, Ks)
    return B.T @ q_mu + c, kss - np.sum(B * B, 0) + np.sum((q_sqrt.T @ B) ** 2, 0)


def predict_y_moments(lik, fm, fv):
    """likelihood.predict_mean_and_var(Fmu, Fvar): gpflow's 20-point Gauss-Hermite of the conditional moments (Poisson,
    Gamma), its closed form for Bernoulli/inv_probit, and waveome's plug-in override for the negative binomial
    (waveome/likelihoods.py:48-51: mean exp(Fmu), variance m + alpha m^2 at m = exp(Fmu))."""
    from scipy.special import erfc
    f = fm[:, None] + np.sqrt(2.0 * fv)[:, None] * GH_X[None, :]
    w = GH_W[None, :] / math.sqrt(math.pi)
    t = lik["type"]
    if t == "negative_binomial":
        m = np.exp(fm)
        return m, m + lik["alpha"] * m * m
    if t == "bernoulli":
        p = 1e-3 + (1.0 - 2e-3) * 0.5 * erfc(-fm / np.sqrt(2.0 * (1.0 + fv)))
        return p, p - p * p
    if t == "poisson":
        cm, cv = np.exp(f), np.exp(f)
    elif t == "gamma":
        cm, cv = lik["shape"] * np.exp(f), lik["shape"] * np.exp(2.0 * f)
    else:
        raise ValueError(t)
    Ey = np.sum(w * cm, 1)
    return Ey, np.sum(w * (cv + cm * cm), 1) - Ey * Ey


def fit(model, lik, X, y, maxiter=50000, maxfun=50000, maxcor=10, ftol=2.220446049250313e-09, gtol=1e-05, maxls=20):
    """What the reference does with this objective: L-BFGS-B (waveome/model_fitting.py:267-281) -- here on the collapsed
    bound, i.e. over the hyper-parameters only, every evaluation at its optimal q.  Returns dict(x, F, f, log_prior, nit,
    nfev): F the bound at x, f = -(F + log prior) the minimised MAP objective."""
    import scipy.optimize as so
    state = {"sites": None}

    def fun(x):
        try:
            r = vgp_collapsed(model, lik, X, y, x, sites=state["sites"], rho=0.5, tol=1e-11, maxit=5000)
        except np.linalg.LinAlgError:
            return float("nan"), np.full(len(x), float("nan"))
        if np.isfinite(r["F"]):
            state["sites"] = r["sites"]
        return r["f"], -r["grad"]

    res = so.minimize(fun, go.pack(model), jac=True, method="L-BFGS-B",
                      options=dict(maxiter=maxiter, maxfun=maxfun, maxcor=maxcor, ftol=ftol, gtol=gtol, maxls=maxls))
    r = vgp_collapsed(model, lik, X, y, res.x, sites=state["sites"], rho=0.5, tol=1e-11, maxit=5000, want_grad=False)
    return dict(x=res.x, F=r["F"], f=r["f"], log_prior=r["log_prior"], nit=int(res.nit), nfev=int(res.nfev),
                message=str(res.message))
