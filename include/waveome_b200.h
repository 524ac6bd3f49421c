/* waveome_b200 — C ABI of the B200 batched GP model-fitting engine.
 *
 * The reference (omicsEye/waveome v0.1.3) has no FFI/plugin interface: its boundary on this path is the
 * Python closure handed to SciPy, f(x: float64[P]) -> (loss, grad), created at
 *   waveome/model_fitting.py:276-281   gpflow.optimizers.Scipy().minimize(m.training_loss, ...)
 *   waveome/model_classes.py:325-334   optimizer.minimize(closure=self.training_loss_closure(data), ...)
 * and the per-outcome fan-out around it (waveome/model_search.py:250-393, 474-489).
 * The entry points below are what a ctypes binding of that path binds:
 *
 *   wv_batch_create      <- model construction: gpflow.models.GPR(data=(X, Y), kernel=k)
 *                           (waveome/model_fitting.py:150-155) for B models that share X
 *   wv_batch_eval        <- m.training_loss + gradients, i.e. -(GPR.log_marginal_likelihood + log prior)
 *                           (mirror: waveome/model_types_DEPR.py:49-56)
 *   wv_batch_fit_lbfgs   <- gpflow.optimizers.Scipy().minimize(..., method="L-BFGS-B", options=...)
 *                           (waveome/model_fitting.py:276-281, waveome/model_classes.py:309-334)
 *
 * Plain pointers and sizes only; no torch types.  All floating point data is IEEE fp64.
 * Return codes: 0 ok, <0 API misuse or CUDA failure (wv_last_error() explains).  Numerical trouble is
 * reported per model in `status` bits, never as an error code.
 * Threading: an engine and its batches belong to one host thread / one GPU.
 */
#ifndef WAVEOME_B200_H
#define WAVEOME_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct wv_engine wv_engine;
typedef struct wv_batch wv_batch;

/* leaf kernel types (GPflow names: squared_exponential, matern12/32/52, periodic[SE base], linear|lin,
 * constant, categorical (waveome/kernels.py:86-124), polynomial|poly (waveome/kernels.py:42-83), empty) */
enum { WVK_SE = 0, WVK_M12 = 1, WVK_M32 = 2, WVK_M52 = 3, WVK_PERIODIC = 4, WVK_LINEAR = 5, WVK_CONST = 6,
       WVK_CAT = 7, WVK_POLY = 8, WVK_EMPTY = 9 };
/* parameter bijectors: identity, softplus (gpflow positive()), softplus + lower bound, exp */
enum { WVT_IDENTITY = 0, WVT_SOFTPLUS = 1, WVT_SOFTPLUS_SHIFT = 2, WVT_EXP = 3 };
/* priors on the constrained value (tfd.Horseshoe / tfd.Laplace / tfd.Uniform) */
enum { WVP_NONE = 0, WVP_HORSESHOE = 1, WVP_LAPLACE = 2, WVP_UNIFORM = 3 };
/* per-model status bits */
enum { WVS_OK = 0, WVS_CHOL_FAIL = 1, WVS_NONFINITE = 2, WVS_MAXITER = 4, WVS_LINESEARCH = 8, WVS_INNER_CAP = 16,
       WVS_RESTORED = 64 /* wv_batch_fit_adam: a step hit a failed factorisation, the last checkpoint was restored */,
       WVS_SITE_BOUND = 32 /* ZINB: a site precision sits at its lower bound (1e-6): the value is a valid lower bound,
                              the envelope-theorem gradient is approximate */ };

/* One kernel program = sum over components of products of leaves, plus the parameter slot table.
 * Arrays are flat; a batch passes `n_programs` of these back to back. */
typedef struct wv_program_desc {
  int32_t n_comp;              /* additive components */
  int32_t n_leaves;            /* leaves over all components */
  int32_t n_slots;             /* parameter slots (trainable + fixed), including noise variance and mean */
  int32_t noise_slot;          /* slot of the Gaussian likelihood variance */
  int32_t mean_slot;           /* slot of the constant mean, or -1 for a zero mean */
  const int32_t* comp_start;   /* [n_comp + 1] leaf ranges of the components */
  const int32_t* leaf_type;    /* [n_leaves] WVK_* */
  const int32_t* leaf_dim;     /* [n_leaves] covariate column (active_dims[0]) */
  const int32_t* leaf_s_var;   /* [n_leaves] slot of the variance, -1 if none */
  const int32_t* leaf_s_ls;    /* [n_leaves] slot of lengthscales (or polynomial offset), -1 if none */
  const int32_t* leaf_s_aux;   /* [n_leaves] slot of the period, -1 if none */
  const int32_t* leaf_degree;  /* [n_leaves] polynomial degree */
  const int32_t* slot_transform; /* [n_slots] WVT_* */
  const int32_t* slot_xindex;    /* [n_slots] index into the packed unconstrained vector, -1 = not trainable */
  const int32_t* slot_prior;     /* [n_slots] WVP_* */
  const double* slot_fixed;      /* [n_slots] constrained value used when not trainable */
  const double* slot_shift;      /* [n_slots] lower bound of WVT_SOFTPLUS_SHIFT */
  const double* slot_pa;         /* [n_slots] prior parameter a (scale | loc | low) */
  const double* slot_pb;         /* [n_slots] prior parameter b (  -   | scale | high) */
  int32_t lik_slot2;             /* slot of a second likelihood parameter (zero-inflated negative binomial: km,
                                    waveome/likelihoods.py:96-139), or -1 */
} wv_program_desc;

typedef struct wv_batch_desc {
  int32_t n;                 /* observations */
  int32_t D;                 /* covariate columns */
  int32_t B;                 /* models in the batch (all share X) */
  int32_t P;                 /* stride of the packed parameter vectors (>= max trainable count) */
  const double* X;           /* HOST [n, D] row-major */
  const double* Y;           /* HOST [B, n] row-major: outcome of each model */
  int32_t n_programs;
  const wv_program_desc* programs; /* HOST [n_programs] */
  const int32_t* prog_id;    /* HOST [B] program of each model */
} wv_batch_desc;

typedef struct wv_lbfgs_opts {
  int32_t maxcor;    /* SciPy default 10 */
  int32_t maxiter;   /* SciPy default 15000 (waveome passes 50000) */
  int32_t maxfun;    /* SciPy default 15000 */
  int32_t maxls;     /* SciPy default 20 */
  double ftol;       /* SciPy default 2.220446049250313e-09 (= factr * epsmch) */
  double gtol;       /* SciPy default 1e-5 (pgtol) */
  int32_t chol_fail_policy; /* Cholesky failure at a line-search trial point: 0 = treat the trial as non-finite and let
                               the search recover (default); 1 = abandon the fit there, as the TensorFlow exception does
                               in the reference (waveome/model_classes.py:323-340).  A failure at the start point always
                               abandons the fit. */
  int32_t reserved;
} wv_lbfgs_opts;

int wv_engine_create(int device, wv_engine** out);
/* flags: WV_ENGINE_HIGH_PRIORITY -- the engine's stream is scheduled ahead of the default engines' (for the few straggler
 * models of a batch that finish while the next batch already runs, see wv_batch_set_engine / wv_batch_fit_lbfgs_run). */
#define WV_ENGINE_HIGH_PRIORITY 1
int wv_engine_create2(int device, int flags, wv_engine** out);
void wv_engine_destroy(wv_engine* e);
/* CUDA stream the engine launches on (cudaStream_t as void*), for event timing by the caller */
void* wv_engine_stream(wv_engine* e);
/* Batches with at least `nt` 64-row tiles ((n + 1 + 63) / 64) take the large-n schedule (right-looking panel Cholesky
 * with look-ahead, recursive-doubling triangular inverse) instead of the batched left-looking one.  Default 16
 * (n >= 960); the environment variable WV_BIG_NT overrides the default at engine creation.  Both schedules compute the
 * same quantities; the switch only moves work between launch shapes. */
int wv_engine_set_large_n_tiles(wv_engine* e, int nt);

int wv_batch_create(wv_engine* e, const wv_batch_desc* desc, wv_batch** out);
/* Same with flags.  WV_BATCH_KEEP_ROW_ORDER: device rows stay in the caller's order (by default the rows are sorted by the
 * categorical columns the programs use -- the marginal likelihood is invariant, the element-wise kernels skip whole
 * zero blocks).  Needed by wv_batch_eval_elbo, whose whitened variational parameters refer to chol(K) in the caller's
 * row order. */
enum { WV_BATCH_KEEP_ROW_ORDER = 1 };
int wv_batch_create2(wv_engine* e, const wv_batch_desc* desc, int32_t flags, wv_batch** out);
void wv_batch_destroy(wv_batch* b);
/* bytes of device workspace held by the batch */
int64_t wv_batch_workspace_bytes(const wv_batch* b);
/* replace the outcomes (HOST [B, n]) without rebuilding programs / workspaces */
int wv_batch_set_y(wv_batch* b, const double* Y);

/* Likelihood of every model of the batch (waveome/utilities.py:989-1009 gp_likelihood_crosswalk): 0 Gaussian (default;
 * exact GPR marginal likelihood), 1 Poisson with exp link (gpflow.likelihoods.Poisson), 2 negative binomial with log
 * link and dispersion `param` = alpha (waveome/likelihoods.py:16-79), 3 Bernoulli with gpflow's inv_probit link (y in
 * {0, 1}), 4 Gamma with exp link and shape `param` (gpflow.likelihoods.Gamma), 5 zero-inflated negative binomial
 * (waveome/likelihoods.py:96-139: alpha = `param`, km = the second parameter, see wv_batch_set_likelihood2; its zero
 * branch is not log-concave: site precisions are projected onto >= 1e-6, status bit WVS_SITE_BOUND when that binds).  For 1-5 the objective is the variational bound of gpflow.models.VGP / PSVGP
 * with Z = X (waveome/model_fitting.py:158-185, waveome/model_classes.py:1082-1126) maximised over the variational
 * distribution for the given hyper-parameters: f = -(max_q ELBO + log prior), `lml` reports max_q ELBO, Y holds the
 * observations.  The programs' noise slot: ignored for Poisson / Bernoulli; for the negative binomial and the Gamma a
 * TRAINABLE noise slot is the likelihood parameter (alpha with an Exp bijector, shape with softplus) and gets
 * d(bound)/d(parameter), a frozen one means parameter = `param`.
 * Status bit 16: the inner iteration hit its sweep cap. */
int wv_batch_set_likelihood(wv_batch* b, int32_t kind, double param);
/* Same with a second parameter (kind 5: km, used when the programs have no lik_slot2); wv_batch_set_likelihood passes 1. */
int wv_batch_set_likelihood2(wv_batch* b, int32_t kind, double param, double param2);
/* Posterior mean and variance of the latent f at the training inputs after the last evaluation of a non-Gaussian
 * batch, HOST [B, n] each (predict_f at the training inputs). */
int wv_batch_get_latent(wv_batch* b, double* fmean, double* fvar);

/* Per-model mask over the additive components of its program (bit c = component c takes part in K; default all
 * ones).  "The model without component k, same parameter values" is what the reference's feature importances
 * evaluate (waveome/utilities.py:657-662 pops the component and predicts again): one program, B masks. HOST [B]. */
int wv_batch_set_component_mask(wv_batch* b, const uint32_t* mask);

/* Scheduling hint: solo != 0 says that nothing else runs on the device while this batch's calls run (one batch, one
 * stream).  Mid-size batches (n_active * tiles-per-side <= 10 000) then factorise in one persistent launch per evaluation
 * instead of a launch pair per tile column; results are bit-identical either way.  Default 0. */
int wv_batch_set_solo(wv_batch* b, int solo);

/* One LML+gradient evaluation of every model.  HOST buffers:
 *   x [B, P] unconstrained parameters; f [B] = -(lml + log prior); grad [B, P] = df/dx; lml [B]; status [B]. */
int wv_batch_eval(wv_batch* b, const double* x, double* f, double* grad, double* lml, int32_t* status);

/* Same, with DEVICE buffers, enqueued on the engine stream without host synchronisation. */
int wv_batch_eval_device(wv_batch* b, const double* d_x, double* d_f, double* d_grad, double* d_lml,
                         int32_t* d_status);

/* L-BFGS-B MAP fit of every model, starting from x (HOST [B, P], overwritten with the optimum).
 * f, lml [B] are evaluated at the returned x; n_iter, n_eval, status [B]. */
int wv_batch_fit_lbfgs(wv_batch* b, double* x, const wv_lbfgs_opts* opts, double* f, double* lml,
                       int32_t* n_iter, int32_t* n_eval, int32_t* status);

/* The same fit in three calls, for callers that want the results of the finished models while a few stragglers keep
 * iterating (kernel search: a level's batch lasts as long as its slowest model).
 *   begin   upload the starts x (HOST [B, P]) and initialise every model's optimiser state
 *   run     optimiser rounds while more than min_active models are still iterating; *n_active = how many are left
 *   report  as wv_batch_fit_lbfgs's outputs; finished [B] (may be NULL): 1 where the model's fit has ended.  While models
 *           are still iterating only the finished ones are evaluated: f / lml / status of the others are undefined.
 * wv_batch_fit_lbfgs(...) == begin, run(min_active = 0), report. */
int wv_batch_fit_lbfgs_begin(wv_batch* b, const double* x, const wv_lbfgs_opts* opts);
/* Re-home a batch onto another engine (= stream) of the same device; no call on the batch may be in progress. */
int wv_batch_set_engine(wv_batch* b, wv_engine* e);
int wv_batch_fit_lbfgs_run(wv_batch* b, int32_t min_active, int32_t* n_active);
int wv_batch_fit_lbfgs_report(wv_batch* b, double* x, double* f, double* lml, int32_t* n_iter, int32_t* n_eval,
                              int32_t* status, int32_t* finished);

/* Objective (B) at GIVEN variational parameters: the whitened bound of gpflow.models.VGP / SVGP with Z = X that the live
 * reference API optimises (PSVGP: waveome/model_classes.py:1082-1126; bound: gpflow SVGP.elbo, in-repo mirror
 * waveome/model_types_DEPR.py:126-158; VGP branches: waveome/model_fitting.py:158-185):
 *     L L^T = K + jitter I,  f_mean = c + L q_mu,  f_var_i = |(L tril(q_sqrt))_i|^2,
 *     elbo = sum_i E_{N(f_mean_i, f_var_i)} log p(y_i | f_i) - KL[N(q_mu, q_sqrt q_sqrt^T) || N(0, I)]
 * with the batch's likelihood (wv_batch_set_likelihood; Gaussian: the programs' noise slot).  f = -(elbo + log prior) is
 * what training_loss hands to the optimisers.  HOST buffers: x [B, P], q_mu [B, n], q_sqrt [B, n, n] row-major (the lower
 * triangle is read); jitter = gpflow's default_jitter() = 1e-6.  The batch must have been created with
 * WV_BATCH_KEEP_ROW_ORDER.  wv_batch_eval / wv_batch_fit_* evaluate the same bound MAXIMISED over (q_mu, q_sqrt); this
 * entry point exists to check that statement and to score externally optimised variational parameters. */
int wv_batch_eval_elbo(wv_batch* b, const double* x, const double* q_mu, const double* q_sqrt, double jitter,
                       double* elbo, double* f, int32_t* status);

/* Adam with the schedule of the reference's default optimiser (BaseGP.optimize_params "adam/gradient" branch,
 * waveome/model_classes.py:344-462; kernel_test calls it, waveome/model_search.py:2284-2297): Keras Adam steps on the
 * unconstrained hyper-parameters; every `check_every` steps the loss after the step is recorded, the parameters are
 * snapshotted and (every `decay_every` steps) the learning rate becomes learning_rate * decay_rate^(i / decay_every); a
 * model stops when the loss fell by less than `convergence_threshold` between two checkpoints, after `max_iter` steps
 * (WVS_MAXITER), on a NaN checkpoint loss (WVS_NONFINITE; upstream stops with the NaN values in place, the engine returns
 * the last finite snapshot), or when a step meets a failed factorisation (WVS_RESTORED: the last snapshot is returned,
 * upstream's InvalidArgumentError branch).
 * The natural-gradient half of the upstream step acts on (q_mu, q_sqrt); the engine's objective is the bound maximised
 * over q, so that half is exact here (its gamma = 1 limit for the Gaussian likelihood).  x HOST [B, P] in / out. */
typedef struct wv_adam_opts {
  double learning_rate;          /* upstream default 0.1 */
  double decay_rate;             /* 0.96 */
  double beta1, beta2, epsilon;  /* Keras Adam defaults 0.9, 0.999, 1e-7 */
  double convergence_threshold;  /* 1e-9 */
  int32_t max_iter;              /* num_opt_iter, 50000 */
  int32_t check_every;           /* 100 */
  int32_t decay_every;           /* 500 */
  int32_t reserved;
} wv_adam_opts;
int wv_batch_fit_adam(wv_batch* b, double* x, const wv_adam_opts* opts, double* f, double* lml, int32_t* n_iter,
                      int32_t* status);

/* Post-fit quantities of the last evaluation (wv_batch_eval / wv_batch_eval_device / the final evaluation of
 * wv_batch_fit_lbfgs): alpha = (K + sigma^2 I)^{-1} (y - c), HOST [B, n] in the caller's row order, and the posterior
 * mean of every model at new inputs, mean[b][i] = c_b + sum_j k_b(xnew_i, x_j) alpha_b[j]  (gpflow GPR.predict_f mean,
 * which waveome/utilities.py:614-707 calc_feature_importance_components and :710-974 consume).
 * Xnew HOST [m, D] row-major, mean HOST [B, m].  At the training inputs the mean is y - sigma^2 alpha. */
int wv_batch_get_alpha(wv_batch* b, double* alpha);
/* diag((K + sigma^2 I)^-1) of the last evaluation, HOST [B, n]: the predictive variance at the training inputs is
 * var f_i = sigma^2 - sigma^4 diag_i, var y_i = var f_i + sigma^2 (gpflow GPR.predict_f / predict_y, full_cov=False). */
int wv_batch_get_kinv_diag(wv_batch* b, double* diag);
int wv_batch_predict_mean(wv_batch* b, const double* Xnew, int32_t m, double* mean);
/* Same plus the predictive variance of f (gpflow GPR.predict_f, full_cov = False; predict_y adds sigma^2):
 * var[b][i] = k_b(xnew_i, xnew_i) - k*_i^T (K_b + sigma^2 I)^-1 k*_i, HOST [B, m]; var may be NULL.  With a non-Gaussian
 * likelihood (wv_batch_set_likelihood) both are the latent posterior under the converged sites, sigma^2 I replaced by
 * jitter I + diag(1 / lam) (gpflow VGP.predict_f at the optimal q). */
int wv_batch_predict_f(wv_batch* b, const double* Xnew, int32_t m, double* mean, double* var);

/* Run-time specialised element-wise kernels.  The Gram builder and the gradient reduction normally interpret the flat
 * kernel program; for a batch whose models all share one program structure the host (waveome_b200/specialize.py) can
 * generate straight-line CUDA text for that structure.  wv_batch_specialize compiles it with NVRTC for sm_100a (cached
 * per process under `key`, a hash of the text), loads it and routes the batch's Gram / gradient launches to the two
 * named kernels (dynamic shared memory sizes as given); src = NULL returns the batch to the interpreter kernels.  Both
 * paths compute the same quantities (reference: the GPflow kernel-tree ops behind waveome/regularization.py:14-189).
 * wv_rtc_check only compiles (no GPU needed) and returns the cubin size, or < 0 with the compiler log in `log`. */
int wv_batch_specialize(wv_batch* b, const char* key, const char* src, const char* gram_name, const char* grad_name,
                        int32_t gram_smem, int32_t grad_smem);
int wv_rtc_check(const char* src, char* log, int log_len);
/* Disk cache of compiled texts (<dir>/<key>.cubin; NULL or "" = none): looked up before NVRTC is invoked, written after.
 * wv_rtc_precompile_text fills it without a GPU (0 = compiled, 1 = already there). */
void wv_rtc_set_cache(const char* dir);
int wv_rtc_precompile_text(const char* key, const char* src);

/* counters since batch creation: kernels launched, batched evaluation rounds, model evaluations */
void wv_batch_counters(const wv_batch* b, int64_t* launches, int64_t* rounds, int64_t* model_evals);

/* Optional per-kernel-class device timing (CUDA events on the engine stream, resolved at the host syncs the fit
 * loop already has).  Classes, in order: gram, chol_diag, chol_panel, trtri, extract, kinv, grad, finalize, lbfgs,
 * chol_syrk (trailing updates of the large-n path), sites (site sweeps of the variational path).
 * wv_batch_profile_read fills ms[i] / launches[i] for i < n and returns the number of classes. */
void wv_batch_profile_enable(wv_batch* b, int on);
int wv_batch_profile_read(wv_batch* b, double* ms, int64_t* launches, int n);

const char* wv_last_error(void);
const char* wv_version(void);

#ifdef __cplusplus
}
#endif
#endif
