// waveome_b200 — batched L-BFGS-B driver (unbounded case), one state machine per model.
//
// Replaces gpflow.optimizers.Scipy().minimize(..., method="L-BFGS-B") on the hot path
// (/root/reference waveome/model_fitting.py:276-281, waveome/model_classes.py:309-334), i.e.
// scipy.optimize.minimize -> L-BFGS-B 3.0 (Byrd, Lu, Nocedal, Zhu 1995; Morales & Nocedal 2011).
// With no bounds the algorithm reduces to:
//   * first iteration / after a memory reset: d = -g (generalised Cauchy point with theta = 1)
//   * otherwise the subspace-minimisation direction from the compact representation
//        d = (1/theta) r + (1/theta^2) W K^{-1} W^T r,   r = -g,  W = [Y, theta S]
//     with K factorised as in `formk` (LEL^T of the 2col x 2col middle matrix)
//   * More'-Thuente line search `dcsrch` (ftol 1e-3, gtol 0.9, xtol 0.1), first step 1/|d|
//   * stop on max|g| <= pgtol or (f_old - f) <= factr*epsmch*max(|f_old|,|f|,1); skip the update
//     when y's <= epsmch*(-g_old'd)*stp; reset the memory when a factorisation or the search fails.
// The same source is compiled for the device (one thread per model, wv_lbfgsb_kernel) and for the
// host (tests/ compare it against SciPy on CPU).  All arithmetic is fp64.
#pragma once
#include <cmath>
#include <cstdint>
#include "wv_common.cuh"

#define WV_LB_MAXCOR 20

enum WvLbTask : int32_t {
  WV_LB_FG = 0,            // evaluate f, g at x and call step again
  WV_LB_CONV_PG = 1,       // CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL
  WV_LB_CONV_F = 2,        // CONVERGENCE: REL_REDUCTION_OF_F <= FACTR*EPSMCH
  WV_LB_ABNORMAL = 3,      // ABNORMAL_TERMINATION_IN_LNSRCH
  WV_LB_MAXITER = 4,       // STOP: TOTAL NO. of ITERATIONS REACHED LIMIT
  WV_LB_MAXFUN = 5,        // STOP: TOTAL NO. of f AND g EVALUATIONS EXCEEDS LIMIT
  WV_LB_START = 6
};

struct WvLbOpts {
  int32_t m;        // maxcor
  int32_t maxiter;
  int32_t maxfun;
  int32_t maxls;
  double ftol;      // factr * epsmch
  double pgtol;
  int32_t chol_fail_policy;   // 0: failed trial = non-finite value, 1: abandon the fit
  int32_t reserved;
};

// scalar part of the per-model state
struct WvLbScalars {
  double f, fold, theta, stp, gd, gdold, dtd, dnorm, sbgnrm, stpmx;
  // dcsrch
  double finit, ginit, gtest, width, width1, stx, fx, gx, sty, fy, gy, stmin, stmax;
  int32_t brackt, stage, ls_started;
  int32_t col, head, itail, iupdat, iter, nfgv, ifun, iback, info, task, updatd, nskip, first;
  int32_t nit, neval;   // reported counters: accepted iterates (SciPy's nit) and objective evaluations actually made
};

// number of doubles of vector/matrix workspace per model
WV_HD size_t wv_lb_work_doubles(int P, int m) {
  return (size_t)4 * P + (size_t)2 * P * m + (size_t)4 * m * m + (size_t)4 * m * m + (size_t)2 * m;
}

struct WvLbState {
  int P, m;
  WvLbScalars* s;
  double *x, *g;                        // [P] iterate / gradient, owned by the caller (the evaluation buffers)
  double *t, *r, *d, *z;                // [P]
  double *ws, *wy;                      // [P x m] column-major (column c at c*P)
  double *sy, *ss, *wt;                 // [m x m] column-major, leading dimension m
  double *yy;                           // [m x m] y_i' y_j for i >= j (lower incl. diagonal), kept across iterations
  double *wn;                           // [2m x 2m] column-major, leading dimension 2m
  double *wv;                           // [2m]
  WV_HD void bind(WvLbScalars* sc, double* x_, double* g_, double* w, int P_, int m_) {
    P = P_; m = m_; s = sc;
    x = x_; g = g_; t = w; r = t + P; d = r + P; z = d + P;
    ws = z + P; wy = ws + (size_t)P * m;
    sy = wy + (size_t)P * m; ss = sy + m * m; wt = ss + m * m; yy = wt + m * m;
    wn = yy + m * m; wv = wn + 4 * m * m;
  }
};

// ---------------------------------------------------------------------------------------------
// More'-Thuente step (MINPACK-2 dcstep)
// ---------------------------------------------------------------------------------------------
WV_HD void wv_dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp,
                     double fp, double dp, int32_t& brackt, double stpmin, double stpmax) {
  double sgnd = dp * (dx / fabs(dx));
  double stpf, stpc, stpq, theta, s, gamma, p, q, r;
  if (fp > fx) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    p = (gamma - dx) + theta;
    q = ((gamma - dx) + gamma) + dp;
    r = p / q;
    stpc = stx + r * (stp - stx);
    stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    if (fabs(stpc - stx) < fabs(stpq - stx)) stpf = stpc;
    else stpf = stpc + (stpq - stpc) / 2.0;
    brackt = 1;
  } else if (sgnd < 0.0) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = ((gamma - dp) + gamma) + dx;
    r = p / q;
    stpc = stp + r * (stx - stp);
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
    else stpf = stpq;
    brackt = 1;
  } else if (fabs(dp) < fabs(dx)) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = (gamma + (dx - dp)) + gamma;
    r = p / q;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else if (stp > stx) stpc = stpmax;
    else stpc = stpmin;
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      if (fabs(stpc - stp) < fabs(stpq - stp)) stpf = stpc;
      else stpf = stpq;
      if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
      else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
    } else {
      if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
      else stpf = stpq;
      stpf = fmin(stpmax, stpf);
      stpf = fmax(stpmin, stpf);
    }
  } else {
    if (brackt) {
      theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
      gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      p = (gamma - dp) + theta;
      q = ((gamma - dp) + gamma) + dy;
      r = p / q;
      stpc = stp + r * (sty - stp);
      stpf = stpc;
    } else if (stp > stx) stpf = stpmax;
    else stpf = stpmin;
  }
  if (fp > fx) {
    sty = stp; fy = fp; dy = dp;
  } else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  stp = stpf;
}

// line-search verdicts
#define WV_LS_FG 0
#define WV_LS_CONV 1
#define WV_LS_WARN 2
#define WV_LS_ERROR 3

// MINPACK-2 dcsrch as used by L-BFGS-B (ftol=1e-3, gtol=0.9, xtol=0.1, stpmin=0)
WV_HD int wv_dcsrch(WvLbScalars& S, double f, double g, double& stp, double stpmax, bool start) {
  const double ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmin = 0.0;
  const double xtrapl = 1.1, xtrapu = 4.0;
  if (start) {
    if (stp < stpmin || stp > stpmax || g >= 0.0) return WV_LS_ERROR;
    S.brackt = 0; S.stage = 1; S.finit = f; S.ginit = g; S.gtest = ftol * g;
    S.width = stpmax - stpmin; S.width1 = S.width / 0.5;
    S.stx = 0.0; S.fx = f; S.gx = g; S.sty = 0.0; S.fy = f; S.gy = g;
    S.stmin = 0.0; S.stmax = stp + xtrapu * stp;
    return WV_LS_FG;
  }
  const double ftest = S.finit + stp * S.gtest;
  if (S.stage == 1 && f <= ftest && g >= 0.0) S.stage = 2;
  int verdict = WV_LS_FG;
  if (S.brackt && (stp <= S.stmin || stp >= S.stmax)) verdict = WV_LS_WARN;
  if (S.brackt && S.stmax - S.stmin <= xtol * S.stmax) verdict = WV_LS_WARN;
  if (stp == stpmax && f <= ftest && g <= S.gtest) verdict = WV_LS_WARN;
  if (stp == stpmin && (f > ftest || g >= S.gtest)) verdict = WV_LS_WARN;
  if (f <= ftest && fabs(g) <= gtol * (-S.ginit)) verdict = WV_LS_CONV;
  if (verdict != WV_LS_FG) return verdict;
  if (S.stage == 1 && f <= S.fx && f > ftest) {
    double fm = f - stp * S.gtest, fxm = S.fx - S.stx * S.gtest, fym = S.fy - S.sty * S.gtest;
    double gm = g - S.gtest, gxm = S.gx - S.gtest, gym = S.gy - S.gtest;
    wv_dcstep(S.stx, fxm, gxm, S.sty, fym, gym, stp, fm, gm, S.brackt, S.stmin, S.stmax);
    S.fx = fxm + S.stx * S.gtest; S.fy = fym + S.sty * S.gtest;
    S.gx = gxm + S.gtest; S.gy = gym + S.gtest;
  } else {
    wv_dcstep(S.stx, S.fx, S.gx, S.sty, S.fy, S.gy, stp, f, g, S.brackt, S.stmin, S.stmax);
  }
  if (S.brackt) {
    if (fabs(S.sty - S.stx) >= 0.66 * S.width1) stp = S.stx + 0.5 * (S.sty - S.stx);
    S.width1 = S.width;
    S.width = fabs(S.sty - S.stx);
  }
  if (S.brackt) {
    S.stmin = fmin(S.stx, S.sty);
    S.stmax = fmax(S.stx, S.sty);
  } else {
    S.stmin = stp + xtrapl * (stp - S.stx);
    S.stmax = stp + xtrapu * (stp - S.stx);
  }
  stp = fmax(stp, stpmin);
  stp = fmin(stp, stpmax);
  if ((S.brackt && (stp <= S.stmin || stp >= S.stmax)) || (S.brackt && S.stmax - S.stmin <= xtol * S.stmax))
    stp = S.stx;
  return WV_LS_FG;
}

// ---------------------------------------------------------------------------------------------
// Execution policy.  The state machine below is written once and runs either on ONE thread (host build, the
// thread-per-model kernel: WvExSerial) or on the 32 lanes of a warp that share the model's state in shared memory
// (wv_lb_step_warp_kernel: WvExWarp).  Rules of the warp form:
//   * scalars (WvLbScalars) are written by lane 0 only, between two sync(); what the other lanes branch on is either a
//     value lane 0 broadcasts (bcast) or a field read after a sync() that lane 0 does not write before the next one --
//     control flow is warp-uniform by construction;
//   * vectors / matrix entries are spread over the lanes, but EVERY value is computed by one lane with the operations and
//     the summation order of the serial code (independent dot products, columns, right-hand sides or matrix entries go
//     to different lanes; a recurrence is never split) -- results are bit-identical to the serial policy
//     (tests/test_lbfgs_warp_gpu.py), and the serial policy is the code the CPU suite checks against SciPy.
// ---------------------------------------------------------------------------------------------
struct WvExSerial {
  static constexpr int nl = 1;
  WV_HD int lane() const { return 0; }
  WV_HD void sync() const {}
  WV_HD int bcast(int v, int src = 0) const { (void)src; return v; }
  WV_HD bool any(bool p) const { return p; }
};
#ifdef __CUDACC__
struct WvExWarp {
  static constexpr int nl = 32;
  int ln;
  __device__ int lane() const { return ln; }
  __device__ void sync() const { __syncwarp(); }
  __device__ int bcast(int v, int src = 0) const { return __shfl_sync(0xffffffffu, v, src); }
  __device__ bool any(bool p) const { return __any_sync(0xffffffffu, p) != 0; }
};
#endif

// ---------------------------------------------------------------------------------------------
// small dense helpers (LINPACK dpofa / dtrsl restated for column-major upper-triangular factors)
// ---------------------------------------------------------------------------------------------
WV_HD double wv_dot(const double* a, const double* b, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

// Cholesky A = R^T R of the leading n x n block (upper triangle used/overwritten), n <= WV_LB_MAXCOR.  returns 0 or k>0.
// LINPACK computes column j from the columns before it; here the columns advance together, row k of all columns j > k at
// step k (a lane per column), every entry with LINPACK's operations in LINPACK's order.
template <class EX>
WV_HD int wv_dpofa(const EX& ex, double* a, int lda, int n) {
  constexpr int NOWN = EX::nl == 1 ? WV_LB_MAXCOR : 1;      // columns a lane owns: j = lane + q nl
  double s[NOWN];
  for (int q = 0; q < NOWN; ++q) s[q] = 0.0;
  for (int k = 0; k < n; ++k) {
    int fail = 0;
    if (k % EX::nl == ex.lane()) {          // rows 0..k-1 of column k are final: close it
      const double d = a[k + k * lda] - s[k / EX::nl];
      if (!(d > 0.0)) fail = k + 1;
      else a[k + k * lda] = sqrt(d);
    }
    fail = ex.bcast(fail, k % EX::nl);
    if (fail) return fail;
    ex.sync();
    for (int q = 0; q < NOWN; ++q) {
      const int j = ex.lane() + q * EX::nl;
      if (j > k && j < n) {
        double t = a[k + j * lda];
        for (int i = 0; i < k; ++i) t -= a[i + k * lda] * a[i + j * lda];
        t = t / a[k + k * lda];
        a[k + j * lda] = t;
        s[q] += t * t;
      }
    }
    ex.sync();
  }
  return 0;
}
// solve R^T x = b (job 11) in place, R upper triangular n x n: one thread
WV_HD int wv_dtrsl_t(const double* r, int ldr, int n, double* b) {
  for (int j = 0; j < n; ++j)
    if (r[j + j * ldr] == 0.0) return j + 1;
  for (int j = 0; j < n; ++j) {
    double t = b[j];
    for (int i = 0; i < j; ++i) t -= r[i + j * ldr] * b[i];
    b[j] = t / r[j + j * ldr];
  }
  return 0;
}
// the same solve for ONE right-hand side spread over the lanes, n <= 2 WV_LB_MAXCOR: lane j keeps t_j and subtracts
// r[i][j] b[i] as soon as b[i] is final (i ascending, as in the loop above)
template <class EX>
WV_HD int wv_dtrsl_t(const EX& ex, const double* r, int ldr, int n, double* b) {
  constexpr int NOWN = EX::nl == 1 ? 2 * WV_LB_MAXCOR : 2;
  double t[NOWN];
  bool zero = false;
  for (int q = 0; q < NOWN; ++q) {
    const int j = ex.lane() + q * EX::nl;
    t[q] = 0.0;
    if (j < n) { t[q] = b[j]; zero = zero || r[j + j * ldr] == 0.0; }
  }
  if (ex.any(zero)) return 1;
  for (int i = 0; i < n; ++i) {
    if (i % EX::nl == ex.lane()) b[i] = t[i / EX::nl] / r[i + i * ldr];
    ex.sync();
    const double bi = b[i];
    for (int q = 0; q < NOWN; ++q) {
      const int j = ex.lane() + q * EX::nl;
      if (j > i && j < n) t[q] -= r[i + j * ldr] * bi;
    }
  }
  ex.sync();
  return 0;
}
// solve R x = b (job 01) in place, R upper triangular n x n, one right-hand side spread over the lanes: lane i keeps b[i]
// and adds -b[j] r[i][j] as soon as b[j] is final (j descending, as LINPACK's column sweep does)
template <class EX>
WV_HD int wv_dtrsl_n(const EX& ex, const double* r, int ldr, int n, double* b) {
  constexpr int NOWN = EX::nl == 1 ? 2 * WV_LB_MAXCOR : 2;
  double t[NOWN];
  bool zero = false;
  for (int q = 0; q < NOWN; ++q) {
    const int j = ex.lane() + q * EX::nl;
    t[q] = 0.0;
    if (j < n) { t[q] = b[j]; zero = zero || r[j + j * ldr] == 0.0; }
  }
  if (ex.any(zero)) return 1;
  for (int j = n - 1; j >= 0; --j) {
    if (j % EX::nl == ex.lane()) b[j] = t[j / EX::nl] / r[j + j * ldr];
    ex.sync();
    const double tj = -b[j];
    for (int q = 0; q < NOWN; ++q) {
      const int i = ex.lane() + q * EX::nl;
      if (i < j) t[q] += tj * r[i + j * ldr];
    }
  }
  ex.sync();
  return 0;
}

// circular column index of the i-th stored correction pair (i = 0..col-1)
WV_HD int wv_lb_ptr(const WvLbScalars& S, int m, int i) { return (S.head + i) % m; }

WV_HD void wv_lb_reset_memory(WvLbScalars& S) {
  S.col = 0; S.head = 0; S.theta = 1.0; S.iupdat = 0; S.updatd = 0;
}

// formk for the all-free (unbounded) case: LEL^T factorisation of
//   [ D + Y'Y/theta      R_z'        ]      R_z = upper triangle (incl. diagonal) of S'Y
//   [ R_z                0           ]
// stored in the upper triangle of wn (2m x 2m).  returns 0, -1 or -2 (the same value on every lane).
// The inner products y_i'y_j (i >= j: yy) and s_i'y_j (i >= j: lower part of sy; i < j: its strict upper part) are formed
// ONCE, when a pair enters the memory (wv_lb_matupd, as the Fortran code keeps them in wn1), not recomputed from the
// vectors in every iteration: the same dot products, hence the same values, at a tenth of the work.
template <class EX>
WV_HD int wv_lb_formk(const EX& ex, WvLbState& L) {
  const WvLbScalars& S = *L.s;
  const int m = L.m, col = S.col, m2 = 2 * m;
  const double theta = S.theta;
  double* wn = L.wn;
  for (int e = ex.lane(); e < col * col; e += EX::nl) {
    const int iy = e / col, jy = e - iy * col, is = col + iy, js = col + jy;
    if (jy <= iy) {
      double v = L.yy[iy + jy * m] / theta;                 // Y'ZZ'Y / theta
      if (jy == iy) v += L.sy[iy + iy * m];                 // + D
      wn[jy + iy * m2] = v;
      wn[js + is * m2] = 0.0;                               // S'AA'S * theta (no active variables)
    }
    wn[jy + is * m2] = jy < iy ? 0.0 : L.sy[iy + jy * m];   // -L_a' (none) | R_z' = s_iy' y_jy, jy >= iy
  }
  ex.sync();
  if (wv_dpofa(ex, wn, m2, col) != 0) return -1;
  for (int js = col + ex.lane(); js < 2 * col; js += EX::nl) wv_dtrsl_t(wn, m2, col, wn + (size_t)js * m2);
  ex.sync();
  for (int e = ex.lane(); e < col * col; e += EX::nl) {
    const int is = col + e / col, js = col + e % col;
    if (js >= is) wn[is + js * m2] += wv_dot(wn + (size_t)is * m2, wn + (size_t)js * m2, col);
  }
  ex.sync();
  if (wv_dpofa(ex, wn + col + (size_t)col * m2, m2, col) != 0) return -2;
  return 0;
}

// formt: T = theta*SS + L*D^{-1}*L' (upper triangle of wt), then Cholesky.  returns 0 or -3.
template <class EX>
WV_HD int wv_lb_formt(const EX& ex, WvLbState& L) {
  const WvLbScalars& S = *L.s;
  const int m = L.m, col = S.col;
  const double theta = S.theta;
  for (int e = ex.lane(); e < col * col; e += EX::nl) {
    const int i = e / col, j = e - i * col;
    if (j < i) continue;
    if (i == 0) { L.wt[0 + j * m] = theta * L.ss[0 + j * m]; continue; }
    double ddum = 0.0;
    for (int k = 0; k < i; ++k) ddum += L.sy[i + k * m] * L.sy[j + k * m] / L.sy[k + k * m];
    L.wt[i + j * m] = ddum + theta * L.ss[i + j * m];
  }
  ex.sync();
  return wv_dpofa(ex, L.wt, m, col) != 0 ? -3 : 0;
}

// matupd: append the pair (s = d, y = r) to the limited memory and refresh SS, SY, YY.  rr, dr, stp, dtd are lane 0's.
template <class EX>
WV_HD void wv_lb_matupd(const EX& ex, WvLbState& L, double rr, double dr, double stp, double dtd) {
  WvLbScalars& S = *L.s;
  const int m = L.m, P = L.P;
  ex.sync();
  if (ex.lane() == 0) {
    if (S.iupdat <= m) {
      S.col = S.iupdat;
      S.itail = (S.head + S.iupdat - 1) % m;
    } else {
      S.itail = (S.itail + 1) % m;
      S.head = (S.head + 1) % m;
    }
    S.theta = rr / dr;
  }
  ex.sync();
  double* wsc = L.ws + (size_t)S.itail * P;
  double* wyc = L.wy + (size_t)S.itail * P;
  for (int i = ex.lane(); i < P; i += EX::nl) { wsc[i] = L.d[i]; wyc[i] = L.r[i]; }
  const int col = S.col;
  if (S.iupdat > m) {
    // shift the old part of SS (upper), SY (all of it) and YY (lower) one place up-left: an entry moves along its own
    // diagonal, so the diagonals are independent (one per lane, walked top-left to bottom-right)
    for (int dg = ex.lane() - (col - 2); dg <= col - 2; dg += EX::nl) {
      for (int j = dg > 0 ? dg : 0; j < col - 1; ++j) {
        const int i = j - dg;                      // entry (i, j), i = j - dg >= 0
        if (i >= col - 1) break;
        L.sy[i + j * m] = L.sy[(i + 1) + (j + 1) * m];
        if (dg >= 0) L.ss[i + j * m] = L.ss[(i + 1) + (j + 1) * m];
        if (dg <= 0) L.yy[i + j * m] = L.yy[(i + 1) + (j + 1) * m];
      }
    }
  }
  ex.sync();
  // the new row / column of inner products: 4 (col - 1) independent dot products
  for (int e = ex.lane(); e < 4 * (col - 1); e += EX::nl) {
    const int j = e >> 2, p = wv_lb_ptr(S, m, j);
    const double* wsp = L.ws + (size_t)p * P;
    const double* wyp = L.wy + (size_t)p * P;
    switch (e & 3) {
      case 0: L.sy[(col - 1) + j * m] = wv_dot(L.d, wyp, P); break;
      case 1: L.ss[j + (col - 1) * m] = wv_dot(wsp, L.d, P); break;
      case 2: L.sy[j + (col - 1) * m] = wv_dot(wsp, L.r, P); break;     // s_j' y_new: upper part of S'Y (wv_lb_formk)
      default: L.yy[(col - 1) + j * m] = wv_dot(L.r, wyp, P); break;    // y_new' y_j
    }
  }
  if (ex.lane() == (EX::nl > 1 ? EX::nl - 1 : 0)) {
    L.yy[(col - 1) + (col - 1) * m] = wv_dot(L.r, L.r, P);
  }
  if (ex.lane() == 0) {
    L.ss[(col - 1) + (col - 1) * m] = (stp == 1.0) ? dtd : stp * stp * dtd;
    L.sy[(col - 1) + (col - 1) * m] = dr;
  }
  ex.sync();
}

// subsm (all variables free, no bounds): on entry L.d = r = -g; on exit L.z = x + Newton step
template <class EX>
WV_HD void wv_lb_subsm(const EX& ex, WvLbState& L) {
  const WvLbScalars& S = *L.s;
  const int m = L.m, P = L.P, col = S.col, m2 = 2 * m;
  const double theta = S.theta;
  double* wv = L.wv;
  for (int e = ex.lane(); e < 2 * col; e += EX::nl) {
    const int i = e < col ? e : e - col, p = wv_lb_ptr(S, m, i);
    if (e < col) wv[i] = wv_dot(L.wy + (size_t)p * P, L.d, P);
    else wv[col + i] = theta * wv_dot(L.ws + (size_t)p * P, L.d, P);
  }
  ex.sync();
  // K^{-1} wv with the LEL^T factors: the leading 2col x 2col block of wn is the upper-triangular
  // [ R11 J ; 0 R22 ], so K^{-1} = R^{-1} diag(-I, I) R^{-T}.
  wv_dtrsl_t(ex, L.wn, m2, 2 * col, wv);
  for (int i = ex.lane(); i < col; i += EX::nl) wv[i] = -wv[i];
  ex.sync();
  wv_dtrsl_n(ex, L.wn, m2, 2 * col, wv);
  const double it = 1.0 / theta;
  for (int i = ex.lane(); i < P; i += EX::nl) {
    double di = L.d[i];
    for (int jy = 0; jy < col; ++jy) {
      const int p = wv_lb_ptr(S, m, jy);
      const double a = wv[jy] / theta, bcoef = wv[col + jy];
      di += L.wy[(size_t)p * P + i] * a + L.ws[(size_t)p * P + i] * bcoef;
    }
    di *= it;
    L.d[i] = di;
    L.z[i] = L.x[i] + di;
  }
}

WV_HD void wv_lb_start(WvLbState& L) {
  WvLbScalars& S = *L.s;
  S.col = 0; S.head = 0; S.theta = 1.0; S.iupdat = 0; S.updatd = 0; S.itail = 0;
  S.iter = 0; S.nfgv = 0; S.ifun = 0; S.iback = 0; S.info = 0; S.nskip = 0;
  S.fold = 0.0; S.dnorm = 0.0; S.gd = 0.0; S.gdold = 0.0; S.stp = 0.0; S.dtd = 0.0; S.sbgnrm = 0.0;
  S.stpmx = 1e10; S.first = 1; S.ls_started = 0; S.nit = 0; S.neval = 0;
  S.task = WV_LB_FG;
}

// Advance one model after an evaluation: `f`, L.g hold f(x), grad f(x) at the current L.x.
// Returns the new task (the same value on every lane); WV_LB_FG means "evaluate at L.x again".
template <class EX>
WV_HD int wv_lb_step(const EX& ex, WvLbState& L, const WvLbOpts& O, double f) {
  WvLbScalars& S = *L.s;
  const int P = L.P;
  const bool l0 = ex.lane() == 0;
  const double epsmch = 2.220446049250313e-16;
  // sections' verdicts, decided by lane 0 and broadcast
  enum { C_NONE = 0, C_CONV_START, C_NEW_ITER, C_RETURN_FG, C_RESTORE, C_NEW_X, C_UPDATE };
  int code = C_NONE;
  if (l0) {
    S.f = f;
    S.neval += 1;
    if (S.first) {
      S.first = 0;
      S.nfgv = 1;
      double sb = 0.0;
      for (int i = 0; i < P; ++i) sb = fmax(sb, fabs(L.g[i]));
      S.sbgnrm = sb;
      if (sb <= O.pgtol) { S.task = WV_LB_CONV_PG; code = C_CONV_START; }
      else code = C_NEW_ITER;
    }
  }
  code = ex.bcast(code);
  if (code == C_CONV_START) { ex.sync(); return WV_LB_CONV_PG; }
  bool in_linesearch = code != C_NEW_ITER;
  for (;;) {
    if (!in_linesearch) {
      // ---------------- new iteration: search direction ----------------
      ex.sync();                                  // the memory (col, head, theta, updatd) as lane 0 left it
      if (S.col == 0) {
        for (int i = ex.lane(); i < P; i += EX::nl) L.z[i] = L.x[i] - L.g[i];
      } else {
        if (S.updatd) {
          if (wv_lb_formk(ex, L) != 0) {
            ex.sync();
            if (l0) wv_lb_reset_memory(S);
            continue;
          }
        }
        for (int i = ex.lane(); i < P; i += EX::nl) L.d[i] = -L.g[i];
        ex.sync();
        wv_lb_subsm(ex, L);
      }
      // ---------------- lnsrlb: start ----------------  (element i stays with the lane that wrote z[i])
      for (int i = ex.lane(); i < P; i += EX::nl) {
        L.d[i] = L.z[i] - L.x[i];
        L.t[i] = L.x[i];
        L.r[i] = L.g[i];
      }
      ex.sync();
      if (l0) {
        S.dtd = wv_dot(L.d, L.d, P);
        S.dnorm = sqrt(S.dtd);
        S.stpmx = 1e10;
        S.stp = (S.iter == 0) ? fmin(1.0 / S.dnorm, S.stpmx) : 1.0;
        S.fold = S.f;
        S.ifun = 0; S.iback = 0;
        S.ls_started = 0;
      }
    }
    // ---------------- lnsrlb: continue (label 556) ----------------
    code = C_NONE;
    if (l0) {
      S.info = 0;
      S.gd = wv_dot(L.g, L.d, P);
      int verdict = WV_LS_FG;
      if (S.ifun == 0) {
        S.gdold = S.gd;
        if (S.gd >= 0.0) S.info = -4;   // ascent direction in projection
      }
      if (S.info == 0) {
        verdict = wv_dcsrch(S, S.f, S.gd, S.stp, S.stpmx, !S.ls_started);
        S.ls_started = 1;
        if (verdict == WV_LS_FG) { S.ifun += 1; S.nfgv += 1; S.iback = S.ifun - 1; }
        else if (verdict == WV_LS_ERROR) S.info = -4;
      }
      if (S.info != 0 || S.iback >= O.maxls) code = C_RESTORE;
      else if (verdict == WV_LS_FG) code = C_RETURN_FG;
      else code = C_NEW_X;
    }
    code = ex.bcast(code);
    ex.sync();                                    // S.stp for the trial point
    if (code == C_RETURN_FG) {
      const double stp = S.stp;
      if (stp == 1.0) { for (int i = ex.lane(); i < P; i += EX::nl) L.x[i] = L.z[i]; }
      else { for (int i = ex.lane(); i < P; i += EX::nl) L.x[i] = stp * L.d[i] + L.t[i]; }
      if (l0) S.task = WV_LB_FG;
      ex.sync();
      return WV_LB_FG;
    }
    if (code == C_RESTORE) {
      // restore the previous iterate
      for (int i = ex.lane(); i < P; i += EX::nl) { L.x[i] = L.t[i]; L.g[i] = L.r[i]; }
      int abnormal = 0;
      if (l0) {
        S.f = S.fold;
        if (S.col == 0) {
          if (S.info == 0) { S.info = -9; S.nfgv -= 1; S.ifun -= 1; S.iback -= 1; }
          S.iter += 1;
          S.task = WV_LB_ABNORMAL;
          abnormal = 1;
        } else {
          if (S.info == 0) S.nfgv -= 1;
          S.info = 0;
          wv_lb_reset_memory(S);
        }
      }
      abnormal = ex.bcast(abnormal);
      ex.sync();
      if (abnormal) return WV_LB_ABNORMAL;
      in_linesearch = false;
      continue;
    }
    // ---------------- line search finished: NEW_X ----------------
    code = -1;
    double rr = 0.0, dr = 0.0;
    if (l0) {
      S.iter += 1;
      S.nit += 1;
      double sb = 0.0;
      for (int i = 0; i < P; ++i) sb = fmax(sb, fabs(L.g[i]));
      S.sbgnrm = sb;
      // driver checks at NEW_X (scipy _minimize_lbfgsb), then label 777: convergence tests
      if (S.iter >= O.maxiter) code = WV_LB_MAXITER;
      else if (S.nfgv > O.maxfun) code = WV_LB_MAXFUN;
      else if (S.sbgnrm <= O.pgtol) code = WV_LB_CONV_PG;
      else {
        const double ddum = fmax(fmax(fabs(S.fold), fabs(S.f)), 1.0);
        if ((S.fold - S.f) <= O.ftol * ddum) code = WV_LB_CONV_F;
      }
      if (code >= 0) S.task = code;
    }
    code = ex.bcast(code);
    if (code >= 0) { ex.sync(); return code; }
    // BFGS update
    const double stp = S.stp;                     // not written again before the next sync
    for (int i = ex.lane(); i < P; i += EX::nl) {
      L.r[i] = L.g[i] - L.r[i];
      if (stp != 1.0) L.d[i] *= stp;
    }
    ex.sync();
    int upd = 0;
    if (l0) {
      for (int i = 0; i < P; ++i) rr += L.r[i] * L.r[i];
      double ddum;
      if (stp == 1.0) { dr = S.gd - S.gdold; ddum = -S.gdold; }
      else { dr = (S.gd - S.gdold) * stp; ddum = -S.gdold * stp; }
      if (dr <= epsmch * ddum) {
        S.nskip += 1; S.updatd = 0;
      } else {
        S.updatd = 1; S.iupdat += 1;
        upd = 1;
      }
    }
    upd = ex.bcast(upd);
    if (upd) {
      wv_lb_matupd(ex, L, rr, dr, stp, S.dtd);
      if (wv_lb_formt(ex, L) != 0) {
        ex.sync();
        if (l0) wv_lb_reset_memory(S);
      }
    }
    in_linesearch = false;
  }
}

WV_HD int wv_lb_step(WvLbState& L, const WvLbOpts& O, double f) { return wv_lb_step(WvExSerial(), L, O, f); }
