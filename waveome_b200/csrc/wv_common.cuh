// waveome_b200 — common device/host definitions for the batched GP fitting engine (sm_100a).
//
// Replaces, on the model-fitting hot path, what the reference delegates to GPflow/TensorFlow:
//   kernel tree evaluation   waveome/kernels.py:19-31,56-73,95-117,136-139 + gpflow.kernels.* [3P]
//   GPR log marginal lik.    waveome/model_types_DEPR.py:49-56 (mirror of gpflow.models.GPR)
//   priors / transforms      waveome/model_classes.py:837-864, waveome/model_fitting.py:198-242
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#include <cmath>
#include <cstring>
#else
// NVRTC (run-time specialised element-wise kernels, wv_spec.cuh): no host headers; the math functions are built in
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
#define INFINITY __longlong_as_double(0x7ff0000000000000LL)
#endif

#ifdef __CUDACC__
#define WV_HD __host__ __device__ __forceinline__
#else
#define WV_HD inline
#endif

// ---------------------------------------------------------------------------------------------
// limits of the flat "kernel program" (sum of products of leaves)
// ---------------------------------------------------------------------------------------------
#define WV_MAX_COMP 32     // additive components
#define WV_MAX_LEAVES 64   // leaves over all components
#define WV_MAX_FACT 4      // leaves per product component
#define WV_MAX_SLOTS 64    // parameter slots (trainable + fixed), incl. noise variance and mean
#define WV_MAX_DIMS 16     // covariate columns staged per tile

enum WvLeafType : int32_t {
  WV_LEAF_SE = 0, WV_LEAF_M12 = 1, WV_LEAF_M32 = 2, WV_LEAF_M52 = 3, WV_LEAF_PERIODIC = 4,
  WV_LEAF_LINEAR = 5, WV_LEAF_CONST = 6, WV_LEAF_CAT = 7, WV_LEAF_POLY = 8, WV_LEAF_EMPTY = 9
};
enum WvTransform : int32_t { WV_TR_IDENTITY = 0, WV_TR_SOFTPLUS = 1, WV_TR_SOFTPLUS_SHIFT = 2, WV_TR_EXP = 3 };
enum WvPrior : int32_t { WV_PRIOR_NONE = 0, WV_PRIOR_HORSESHOE = 1, WV_PRIOR_LAPLACE = 2, WV_PRIOR_UNIFORM = 3 };

// per-model status bits (replace TF exceptions, SURVEY §5 "failure detection")
#define WV_STATUS_CHOL_FAIL 1
#define WV_STATUS_NONFINITE 2
#define WV_STATUS_MAXITER 4
#define WV_STATUS_LINESEARCH 8
#define WV_STATUS_INNER_CAP 16    // variational path: the site iteration hit its sweep cap
#define WV_STATUS_RESTORED 64     // Adam: a step ran into a failed factorisation, the last checkpoint was restored
#define WV_STATUS_SITE_BOUND 32   // variational path (ZINB): a site precision sits at its lower bound; the value is a valid
                                 // bound, the gradient omits those sites' non-stationarity term

struct WvLeaf {
  int32_t type;    // WvLeafType
  int32_t dim;     // covariate column (active_dims[0])
  int32_t s_var;   // slot of variance (-1: none)
  int32_t s_ls;    // slot of lengthscales / poly offset (-1: none)
  int32_t s_aux;   // slot of period (-1: none)
  int32_t degree;  // polynomial degree
};

struct WvSlot {
  int32_t transform;  // WvTransform
  int32_t xindex;     // index into the unconstrained vector, -1 if not trainable
  int32_t prior;      // WvPrior (only applied when trainable)
  int32_t pad;
  double fixed;       // constrained value when not trainable
  double shift;       // softplus_shift lower bound
  double pa, pb;      // prior params: horseshoe(scale=pa), laplace(loc=pa, scale=pb), uniform(lo=pa, hi=pb)
};

struct WvProgram {
  int32_t n_comp, n_leaves, n_slots, n_x;   // n_x = number of trainable (packed) parameters
  int32_t noise_slot, mean_slot;            // mean_slot = -1 for a zero mean function
  int32_t lik_slot2;                        // second likelihood parameter (ZINB km), -1 if none
  int32_t n_dims, pad;                      // number of distinct covariate columns used
  int32_t comp_start[WV_MAX_COMP + 1];      // leaves of component c are [comp_start[c], comp_start[c+1])
  int32_t dims[WV_MAX_DIMS];                // distinct covariate columns; WvLeaf.dim indexes THIS table
  WvLeaf leaves[WV_MAX_LEAVES];
  WvSlot slots[WV_MAX_SLOTS];
};

// ---------------------------------------------------------------------------------------------
// scalar math shared by host (tests) and device
// ---------------------------------------------------------------------------------------------
WV_HD double wv_softplus(double u) { return fmax(u, 0.0) + log1p(exp(-fabs(u))); }
WV_HD double wv_sigmoid(double u) {
  double e = exp(-fabs(u));
  return u >= 0.0 ? 1.0 / (1.0 + e) : e / (1.0 + e);
}
WV_HD double wv_transform(int tr, double u, double shift) {
  switch (tr) {
    case WV_TR_SOFTPLUS: return wv_softplus(u);
    case WV_TR_SOFTPLUS_SHIFT: return wv_softplus(u) + shift;
    case WV_TR_EXP: return exp(u);
    default: return u;
  }
}
WV_HD double wv_transform_grad(int tr, double u) {
  switch (tr) {
    case WV_TR_SOFTPLUS:
    case WV_TR_SOFTPLUS_SHIFT: return wv_sigmoid(u);
    case WV_TR_EXP: return exp(u);
    default: return 1.0;
  }
}

// ---------------------------------------------------------------------------------------------
// 2^u for the squared-exponential leaves: table of 2^(j/64) (64 doubles, staged in shared memory by the callers) and
// a degree-5 polynomial on |f| <= 1/128.  9 FP64 operations instead of the ~16 + special-case handling of exp();
// relative error < 2 ulp (checked against long double in tests/test_exp2_host.py).  u may be any finite value or
// -inf; results below 2^-1021 are flushed to zero (irrelevant next to sigma^2 >= 1e-6 on the diagonal), above 2^1023
// they become +inf, NaN propagates.
// ---------------------------------------------------------------------------------------------
#define WV_EXP2_TAB 64
WV_HD void wv_exp2_table_entry(int j, double* out) { *out = exp2((double)j / WV_EXP2_TAB); }
// 2^u for finite u in (-1022, 1024): no range handling (the callers below add it)
WV_HD double wv_exp2_core(double u, const double* __restrict__ tab) {
  const double M = 1.5 * 70368744177664.0;                 // 1.5 * 2^46: adding it rounds u to a multiple of 1/64
  const double tb = u + M;
  const double f = u - (tb - M);                           // |f| <= 1/128
  // 2^f - 1 = f ln2 (1 + f ln2/2 (1 + f ln2/3 (1 + f ln2/4 (1 + f ln2/5))))   (Taylor; truncation 3.5e-17)
  double p = fma(f, 1.3333558146428443e-03, 9.6181291076284772e-03);   // ln2^5/120, ln2^4/24
  p = fma(f, p, 5.5504108664821580e-02);                                // ln2^3/6
  p = fma(f, p, 2.4022650695910071e-01);                                // ln2^2/2
  p = fma(f, p, 6.9314718055994531e-01);                                // ln2
  p = f * p;
#ifdef __CUDA_ARCH__
  const int ki = __double2loint(tb);
  const double T = tab[ki & (WV_EXP2_TAB - 1)];
  const double r = fma(T, p, T);
  return __hiloint2double(__double2hiint(r) + ((ki >> 6) << 20), __double2loint(r));
#else
  long long bits;
  memcpy(&bits, &tb, 8);
  const int ki = (int)(bits & 0xffffffffLL);
  const double T = tab[ki & (WV_EXP2_TAB - 1)];
  double r = fma(T, p, T);
  memcpy(&bits, &r, 8);
  bits += (long long)(ki >> 6) << 52;
  memcpy(&r, &bits, 8);
  return r;
#endif
}
// general argument: below 2^-1021 flushed to zero, above 2^1023 +inf, NaN propagates
WV_HD double wv_exp2_fast(double u, const double* __restrict__ tab) {
  double r = wv_exp2_core(u, tab);
  if (!(u > -1021.0)) r = (u != u) ? u : 0.0;
  if (u >= 1024.0) r = INFINITY;
  return r;
}
// u <= 0 (the squared-exponential argument -(s d)^2): one clamp instead of the range checks.  Below -1021 the result is
// 2^-1021 ~ 4e-308 instead of 0 (absolute error irrelevant next to sigma^2 >= 1e-6 on the diagonal); NaN propagates.
WV_HD double wv_exp2_neg(double u, const double* __restrict__ tab) {
  u = (u < -1021.0) ? -1021.0 : u;
  return wv_exp2_core(u, tab);
}
// u < 1024 (log2(variance) - (s d)^2 of a squared-exponential leaf): the clamp at -1021 as ONE unsigned integer minimum on
// the high word instead of a double compare and two selects.  Below -1021 (and for -inf) the result is ~2^-1021 instead
// of 0; a NaN with the sign bit clear -- what the FP64 units produce -- propagates.
WV_HD double wv_exp2_lo(double u, const double* __restrict__ tab) {
#ifdef __CUDA_ARCH__
  const unsigned hi = min((unsigned)__double2hiint(u), 0xC08FE800u);          // 0xC08FE800 00000000 = -1021.0
  return wv_exp2_core(__hiloint2double((int)hi, __double2loint(u)), tab);
#else
  unsigned long long bits;
  memcpy(&bits, &u, 8);
  unsigned hi = (unsigned)(bits >> 32);
  if (hi > 0xC08FE800u) hi = 0xC08FE800u;
  bits = ((unsigned long long)hi << 32) | (bits & 0xffffffffULL);
  memcpy(&u, &bits, 8);
  return wv_exp2_core(u, tab);
#endif
}

// The same 2^u with a 2048-entry table of 2^(j/2048) (16 KB, staged in shared memory from a per-device copy) and a
// degree-3 polynomial on |f| <= 2^-12 (truncation (f ln2)^4 / 24 < 3.4e-17): 7 FP64 operations instead of 9.  Used by the
// run-time specialised element-wise kernels (wv_spec.cuh), whose squared-exponential leaves are bound by the FP64 pipe.
// u < 1024; below -1021 (and -inf) the result is ~2^-1021 instead of 0; NaN propagates.
#define WV_EXP2_BIG_BITS 11
#define WV_EXP2_BIG_TAB (1 << WV_EXP2_BIG_BITS)
WV_HD double wv_exp2_big_core(double u, const double* __restrict__ tab) {
  const double M = 1.5 * 2199023255552.0;                  // 1.5 * 2^41: adding it rounds u to a multiple of 2^-11
  const double tb = u + M;
  const double f = u - (tb - M);                           // |f| <= 2^-12
  double p = fma(f, 5.5504108664821580e-02, 2.4022650695910071e-01);   // ln2^3/6, ln2^2/2
  p = fma(f, p, 6.9314718055994531e-01);                                // ln2
  p = f * p;
#ifdef __CUDA_ARCH__
  const int ki = __double2loint(tb);
  const double T = tab[ki & (WV_EXP2_BIG_TAB - 1)];
  const double r = fma(T, p, T);
  return __hiloint2double(__double2hiint(r) + ((ki >> WV_EXP2_BIG_BITS) << 20), __double2loint(r));
#else
  long long bits;
  memcpy(&bits, &tb, 8);
  const int ki = (int)(bits & 0xffffffffLL);
  const double T = tab[ki & (WV_EXP2_BIG_TAB - 1)];
  double r = fma(T, p, T);
  memcpy(&bits, &r, 8);
  bits += (long long)(ki >> WV_EXP2_BIG_BITS) << 52;
  memcpy(&r, &bits, 8);
  return r;
#endif
}
WV_HD double wv_exp2_big_lo(double u, const double* __restrict__ tab) {
#ifdef __CUDA_ARCH__
  const unsigned hi = min((unsigned)__double2hiint(u), 0xC08FE800u);          // clamp at -1021.0 (see wv_exp2_lo)
  return wv_exp2_big_core(__hiloint2double((int)hi, __double2loint(u)), tab);
#else
  unsigned long long bits;
  memcpy(&bits, &u, 8);
  unsigned hi = (unsigned)(bits >> 32);
  if (hi > 0xC08FE800u) hi = 0xC08FE800u;
  bits = ((unsigned long long)hi << 32) | (bits & 0xffffffffULL);
  memcpy(&u, &bits, 8);
  return wv_exp2_big_core(u, tab);
#endif
}

// tfd.Horseshoe(scale).log_prob(x) (TFP closed-form approximation) and d/dx.
WV_HD void wv_horseshoe(double x, double s, double* logp, double* dlogp) {
  const double g = 0.5614594835668851, b = 1.0420764938351215, h_inf = 1.0801359952503342;
  const double pw = 1.0919284281983377;
  double xs = x / s;
  double xx = xs * xs / 2.0;
  double q = 20.0 / 47.0 * pow(xx, pw);
  double x15 = pow(xx, 1.5);
  double h = 1.0 / (1.0 + x15) + h_inf * q / (1.0 + q);
  double c = -0.5 * log(2.0 * 3.141592653589793 * 3.141592653589793 * 3.141592653589793) - log(g * s);
  double z = log1p(-g) - log(g);
  double t = z - xx / (1.0 - g);
  double hb = h + b * xx;
  double u = g / xx - (1.0 - g) / (hb * hb);
  double l1 = log1p(u);
  *logp = -wv_softplus(t) + log(l1) + c;
  double dA = wv_sigmoid(t) / (1.0 - g);
  double dq = pw * q / xx;
  double dh = -1.5 * sqrt(xx) / ((1.0 + x15) * (1.0 + x15)) + h_inf * dq / ((1.0 + q) * (1.0 + q));
  double du = -g / (xx * xx) + 2.0 * (1.0 - g) * (dh + b) / (hb * hb * hb);
  double dB = du / ((1.0 + u) * l1);
  *dlogp = (dA + dB) * x / (s * s);
}

WV_HD void wv_prior(const WvSlot& sl, double v, double* logp, double* dlogp) {
  *logp = 0.0; *dlogp = 0.0;
  switch (sl.prior) {
    case WV_PRIOR_HORSESHOE: wv_horseshoe(v, sl.pa, logp, dlogp); break;
    case WV_PRIOR_LAPLACE: {
      double dv = v - sl.pa;
      *logp = -fabs(dv) / sl.pb - log(2.0 * sl.pb);
      *dlogp = dv == 0.0 ? 0.0 : (dv > 0.0 ? -1.0 / sl.pb : 1.0 / sl.pb);
    } break;
    case WV_PRIOR_UNIFORM:
      *logp = (v >= sl.pa && v <= sl.pb) ? -log(sl.pb - sl.pa) : -INFINITY;
      break;
    default: break;
  }
}
