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
// small dense helpers (LINPACK dpofa / dtrsl restated for column-major upper-triangular factors)
// ---------------------------------------------------------------------------------------------
// Cholesky A = R^T R of the leading n x n block (upper triangle used/overwritten). returns 0 or k>0.
WV_HD int wv_dpofa(double* a, int lda, int n) {
  for (int j = 0; j < n; ++j) {
    double s = 0.0;
    for (int k = 0; k < j; ++k) {
      double t = a[k + j * lda];
      for (int i = 0; i < k; ++i) t -= a[i + k * lda] * a[i + j * lda];
      t = t / a[k + k * lda];
      a[k + j * lda] = t;
      s += t * t;
    }
    s = a[j + j * lda] - s;
    if (!(s > 0.0)) return j + 1;
    a[j + j * lda] = sqrt(s);
  }
  return 0;
}
// solve R^T x = b (job 11) in place, R upper triangular n x n
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
// solve R x = b (job 01) in place, R upper triangular n x n
WV_HD int wv_dtrsl_n(const double* r, int ldr, int n, double* b) {
  for (int j = 0; j < n; ++j)
    if (r[j + j * ldr] == 0.0) return j + 1;
  for (int j = n - 1; j >= 0; --j) {
    b[j] /= r[j + j * ldr];
    double t = -b[j];
    for (int i = 0; i < j; ++i) b[i] += t * r[i + j * ldr];
  }
  return 0;
}

WV_HD double wv_dot(const double* a, const double* b, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

// circular column index of the i-th stored correction pair (i = 0..col-1)
WV_HD int wv_lb_ptr(const WvLbScalars& S, int m, int i) { return (S.head + i) % m; }

WV_HD void wv_lb_reset_memory(WvLbScalars& S) {
  S.col = 0; S.head = 0; S.theta = 1.0; S.iupdat = 0; S.updatd = 0;
}

// formk for the all-free (unbounded) case: LEL^T factorisation of
//   [ D + Y'Y/theta      R_z'        ]      R_z = upper triangle (incl. diagonal) of S'Y
//   [ R_z                0           ]
// stored in the upper triangle of wn (2m x 2m).  returns 0, -1 or -2.
WV_HD int wv_lb_formk(WvLbState& L) {
  WvLbScalars& S = *L.s;
  const int m = L.m, P = L.P, col = S.col, m2 = 2 * m;
  double* wn = L.wn;
  (void)P;
  // The inner products y_i'y_j (i >= j: yy) and s_i'y_j (i >= j: lower part of sy; i < j: its strict upper part) are
  // formed ONCE, when a pair enters the memory (wv_lb_matupd, as the Fortran code keeps them in wn1), not recomputed from
  // the vectors in every iteration: the same dot products, hence the same values, at a tenth of the work.
  for (int iy = 0; iy < col; ++iy) {
    const int is = col + iy;
    for (int jy = 0; jy <= iy; ++jy) {
      const int js = col + jy;
      wn[jy + iy * m2] = L.yy[iy + jy * m] / S.theta;      // Y'ZZ'Y / theta
      wn[js + is * m2] = 0.0;                               // S'AA'S * theta (no active variables)
    }
    for (int jy = 0; jy < iy; ++jy) wn[jy + is * m2] = 0.0; // -L_a' (no active variables)
    for (int jy = iy; jy < col; ++jy) wn[jy + is * m2] = L.sy[iy + jy * m];      // R_z' = s_iy' y_jy, jy >= iy
    wn[iy + iy * m2] += L.sy[iy + iy * m];
  }
  if (wv_dpofa(wn, m2, col) != 0) return -1;
  for (int js = col; js < 2 * col; ++js) wv_dtrsl_t(wn, m2, col, wn + (size_t)js * m2);
  for (int is = col; is < 2 * col; ++is)
    for (int js = is; js < 2 * col; ++js)
      wn[is + js * m2] += wv_dot(wn + (size_t)is * m2, wn + (size_t)js * m2, col);
  if (wv_dpofa(wn + col + (size_t)col * m2, m2, col) != 0) return -2;
  return 0;
}

// formt: T = theta*SS + L*D^{-1}*L' (upper triangle of wt), then Cholesky.  returns 0 or -3.
WV_HD int wv_lb_formt(WvLbState& L) {
  WvLbScalars& S = *L.s;
  const int m = L.m, col = S.col;
  for (int j = 0; j < col; ++j) L.wt[0 + j * m] = S.theta * L.ss[0 + j * m];
  for (int i = 1; i < col; ++i)
    for (int j = i; j < col; ++j) {
      int k1 = (i < j ? i : j);
      double ddum = 0.0;
      for (int k = 0; k < k1; ++k) ddum += L.sy[i + k * m] * L.sy[j + k * m] / L.sy[k + k * m];
      L.wt[i + j * m] = ddum + S.theta * L.ss[i + j * m];
    }
  return wv_dpofa(L.wt, m, col) != 0 ? -3 : 0;
}

// matupd: append the pair (s = d, y = r) to the limited memory and refresh SS, SY.
WV_HD void wv_lb_matupd(WvLbState& L, double rr, double dr, double stp, double dtd) {
  WvLbScalars& S = *L.s;
  const int m = L.m, P = L.P;
  if (S.iupdat <= m) {
    S.col = S.iupdat;
    S.itail = (S.head + S.iupdat - 1) % m;
  } else {
    S.itail = (S.itail + 1) % m;
    S.head = (S.head + 1) % m;
  }
  double* wsc = L.ws + (size_t)S.itail * P;
  double* wyc = L.wy + (size_t)S.itail * P;
  for (int i = 0; i < P; ++i) { wsc[i] = L.d[i]; wyc[i] = L.r[i]; }
  S.theta = rr / dr;
  const int col = S.col;
  if (S.iupdat > m) {   // shift the old part of SS (upper), SY (all of it) and YY (lower) one place up-left
    for (int j = 0; j < col - 1; ++j) {
      for (int i = 0; i <= j; ++i) L.ss[i + j * m] = L.ss[(i + 1) + (j + 1) * m];
      for (int i = 0; i < col - 1; ++i) L.sy[i + j * m] = L.sy[(i + 1) + (j + 1) * m];
      for (int i = j; i < col - 1; ++i) L.yy[i + j * m] = L.yy[(i + 1) + (j + 1) * m];
    }
  }
  for (int j = 0; j < col - 1; ++j) {
    const int p = wv_lb_ptr(S, m, j);
    L.sy[(col - 1) + j * m] = wv_dot(L.d, L.wy + (size_t)p * P, P);
    L.ss[j + (col - 1) * m] = wv_dot(L.ws + (size_t)p * P, L.d, P);
    // for wv_lb_formk: s_j' y_new (upper part of S'Y) and y_new' y_j, operands in the order formk used to take them
    L.sy[j + (col - 1) * m] = wv_dot(L.ws + (size_t)p * P, L.r, P);
    L.yy[(col - 1) + j * m] = wv_dot(L.r, L.wy + (size_t)p * P, P);
  }
  L.ss[(col - 1) + (col - 1) * m] = (stp == 1.0) ? dtd : stp * stp * dtd;
  L.sy[(col - 1) + (col - 1) * m] = dr;
  L.yy[(col - 1) + (col - 1) * m] = wv_dot(L.r, L.r, P);
}

// subsm (all variables free, no bounds): on entry L.d = r = -g; on exit L.z = x + Newton step
WV_HD void wv_lb_subsm(WvLbState& L) {
  WvLbScalars& S = *L.s;
  const int m = L.m, P = L.P, col = S.col, m2 = 2 * m;
  double* wv = L.wv;
  for (int i = 0; i < col; ++i) {
    const int p = wv_lb_ptr(S, m, i);
    wv[i] = wv_dot(L.wy + (size_t)p * P, L.d, P);
    wv[col + i] = S.theta * wv_dot(L.ws + (size_t)p * P, L.d, P);
  }
  // K^{-1} wv with the LEL^T factors: the leading 2col x 2col block of wn is the upper-triangular
  // [ R11 J ; 0 R22 ], so K^{-1} = R^{-1} diag(-I, I) R^{-T}.
  wv_dtrsl_t(L.wn, m2, 2 * col, wv);
  for (int i = 0; i < col; ++i) wv[i] = -wv[i];
  wv_dtrsl_n(L.wn, m2, 2 * col, wv);
  for (int jy = 0; jy < col; ++jy) {
    const int p = wv_lb_ptr(S, m, jy);
    const double* wyc = L.wy + (size_t)p * P;
    const double* wsc = L.ws + (size_t)p * P;
    const double a = wv[jy] / S.theta, bcoef = wv[col + jy];
    for (int i = 0; i < P; ++i) L.d[i] += wyc[i] * a + wsc[i] * bcoef;
  }
  const double it = 1.0 / S.theta;
  for (int i = 0; i < P; ++i) {
    L.d[i] *= it;
    L.z[i] = L.x[i] + L.d[i];
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
// Returns the new task; WV_LB_FG means "evaluate at L.x again".
WV_HD int wv_lb_step(WvLbState& L, const WvLbOpts& O, double f) {
  WvLbScalars& S = *L.s;
  const int P = L.P;
  const double epsmch = 2.220446049250313e-16;
  S.f = f;
  S.neval += 1;
  bool in_linesearch;
  if (S.first) {
    S.first = 0;
    S.nfgv = 1;
    double sb = 0.0;
    for (int i = 0; i < P; ++i) sb = fmax(sb, fabs(L.g[i]));
    S.sbgnrm = sb;
    if (sb <= O.pgtol) { S.task = WV_LB_CONV_PG; return S.task; }
    in_linesearch = false;
  } else {
    in_linesearch = true;
  }
  for (;;) {
    if (!in_linesearch) {
      // ---------------- new iteration: search direction ----------------
      if (S.col == 0) {
        for (int i = 0; i < P; ++i) L.z[i] = L.x[i] - L.g[i];
      } else {
        if (S.updatd) {
          if (wv_lb_formk(L) != 0) { wv_lb_reset_memory(S); continue; }
        }
        for (int i = 0; i < P; ++i) L.d[i] = -L.g[i];
        wv_lb_subsm(L);
      }
      for (int i = 0; i < P; ++i) L.d[i] = L.z[i] - L.x[i];
      // ---------------- lnsrlb: start ----------------
      S.dtd = wv_dot(L.d, L.d, P);
      S.dnorm = sqrt(S.dtd);
      S.stpmx = 1e10;
      S.stp = (S.iter == 0) ? fmin(1.0 / S.dnorm, S.stpmx) : 1.0;
      for (int i = 0; i < P; ++i) { L.t[i] = L.x[i]; L.r[i] = L.g[i]; }
      S.fold = S.f;
      S.ifun = 0; S.iback = 0;
      S.ls_started = 0;
    }
    // ---------------- lnsrlb: continue (label 556) ----------------
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
      if (verdict == WV_LS_FG) {
        S.ifun += 1; S.nfgv += 1; S.iback = S.ifun - 1;
        if (S.stp == 1.0) { for (int i = 0; i < P; ++i) L.x[i] = L.z[i]; }
        else { for (int i = 0; i < P; ++i) L.x[i] = S.stp * L.d[i] + L.t[i]; }
      } else if (verdict == WV_LS_ERROR) {
        S.info = -4;
      }
    }
    if (S.info != 0 || S.iback >= O.maxls) {
      // restore the previous iterate
      for (int i = 0; i < P; ++i) { L.x[i] = L.t[i]; L.g[i] = L.r[i]; }
      S.f = S.fold;
      if (S.col == 0) {
        if (S.info == 0) { S.info = -9; S.nfgv -= 1; S.ifun -= 1; S.iback -= 1; }
        S.iter += 1;
        S.task = WV_LB_ABNORMAL;
        return S.task;
      }
      if (S.info == 0) S.nfgv -= 1;
      S.info = 0;
      wv_lb_reset_memory(S);
      in_linesearch = false;
      continue;
    }
    if (verdict == WV_LS_FG) { S.task = WV_LB_FG; return S.task; }
    // ---------------- line search finished: NEW_X ----------------
    S.iter += 1;
    S.nit += 1;
    double sb = 0.0;
    for (int i = 0; i < P; ++i) sb = fmax(sb, fabs(L.g[i]));
    S.sbgnrm = sb;
    // driver checks at NEW_X (scipy _minimize_lbfgsb)
    if (S.iter >= O.maxiter) { S.task = WV_LB_MAXITER; return S.task; }
    if (S.nfgv > O.maxfun) { S.task = WV_LB_MAXFUN; return S.task; }
    // label 777: convergence tests
    if (S.sbgnrm <= O.pgtol) { S.task = WV_LB_CONV_PG; return S.task; }
    {
      double ddum = fmax(fmax(fabs(S.fold), fabs(S.f)), 1.0);
      if ((S.fold - S.f) <= O.ftol * ddum) { S.task = WV_LB_CONV_F; return S.task; }
    }
    // BFGS update
    double rr = 0.0;
    for (int i = 0; i < P; ++i) { L.r[i] = L.g[i] - L.r[i]; rr += L.r[i] * L.r[i]; }
    double dr, ddum;
    if (S.stp == 1.0) { dr = S.gd - S.gdold; ddum = -S.gdold; }
    else {
      dr = (S.gd - S.gdold) * S.stp;
      for (int i = 0; i < P; ++i) L.d[i] *= S.stp;
      ddum = -S.gdold * S.stp;
    }
    if (dr <= epsmch * ddum) {
      S.nskip += 1; S.updatd = 0;
    } else {
      S.updatd = 1; S.iupdat += 1;
      wv_lb_matupd(L, rr, dr, S.stp, S.dtd);
      if (wv_lb_formt(L) != 0) wv_lb_reset_memory(S);
    }
    in_linesearch = false;
  }
}
