// waveome_b200 — prelude of the RUN-TIME SPECIALISED element-wise kernels (Gram builder, gradient reduction).
//
// waveome_b200/specialize.py turns one kernel program (sum of products of leaves; the trees come from
// waveome/regularization.py:14-189 and the search expansions) into straight-line CUDA: leaf types, covariate columns,
// parameter slots and the frozen / trainable split are compile-time constants, categorical leaves become integer
// predicates, the squared-exponential leaves of a product share ONE 2^u whose argument also carries log2 of the
// component's variances, and every gradient sum of a component is  sum_e G_e q_e  with  G = W o mask o prod(unit
// values)  formed once per component.  The text is compiled with NVRTC for sm_100a and loaded next to the interpreter
// kernels of wv_elem.cuh, which remain the path for batches with many different programs.
//
// This header is concatenated (with wv_common.cuh and wv_kernels.cuh) in front of the generated code: NVRTC-safe, no
// host headers.
#pragma once
#include "wv_kernels.cuh"

#define WVS_THREADS 256
#define WVS_TPC_MAX 8      // tiles per CTA (wv_elem_tpc)

extern __shared__ __align__(16) unsigned char wvs_smem_raw[];

// constrained parameter values of model b -> theta[0 .. ns)
__device__ __forceinline__ void wvs_stage_theta(const WvBatchDev& bd, int b, const double* __restrict__ xall,
                                                double* __restrict__ theta, int ns) {
  const WvProgram* gp = bd.programs + bd.prog_id[b];
  const double* x = xall + (size_t)b * bd.P;
  for (int s = threadIdx.x; s < ns; s += WVS_THREADS) {
    const WvSlot& sl = gp->slots[s];
    theta[s] = sl.xindex >= 0 ? wv_transform(sl.transform, x[sl.xindex], sl.shift) : sl.fixed;
  }
}

// 2^(j/2048) table: per-device copy in global memory (L2 resident) -> shared memory
__device__ __forceinline__ void wvs_load_tab(double* __restrict__ tab, const double* __restrict__ gtab) {
  const double2* src = reinterpret_cast<const double2*>(gtab);
  double2* dst = reinterpret_cast<double2*>(tab);
  for (int i = threadIdx.x; i < WV_EXP2_BIG_TAB / 2; i += WVS_THREADS) dst[i] = src[i];
}

// Work decomposition: a CTA owns `tpc` consecutive lower tiles of one model; inside every tile warp w owns the 16 x 32
// region at rows (w>>1) * 16, columns (w&1) * 32, lane l the 4 x 4 micro-tile at (+ (l>>3) 4, + (l&7) 4), walked row by
// row.  The warps are AUTONOMOUS: each stages the covariates of its own 16 rows and 32 columns (indices 0..15 and 16..47
// of its private arrays) and synchronises with __syncwarp only, so a warp whose categorical masks let it skip most of
// the squared exponentials moves on to its next tile instead of waiting at a CTA barrier (24 % of the stall samples of
// the barrier version, profiles/r02c).
#define WVS_STAGE 48

__device__ __forceinline__ void wvs_ld4(const double* __restrict__ p, double (&v)[4]) {
  const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void wvs_ld4i(const int* __restrict__ p, int (&v)[4]) {
  const int4 a = *reinterpret_cast<const int4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}

// ---- unit-variance values of the leaves that are not squared exponentials, and the multipliers q of their gradient sums
//      (dK/dparam = component value * q * scalar coefficient; see specialize.py)

// gpflow Matern: r = sqrt(max(r2, 1e-36)), r2 = |a|^2 + |b|^2 - 2ab on a = x / ell (exactly zero on the diagonal).
// q_ls = -(dE/dr) r / E  (0 where the clamp is active: TF's maximum() passes no gradient there)
template <int TYPE>
__device__ __forceinline__ void wvs_matern(double pa, double pb, double& E, double& q_ls) {
  const double r2 = __dadd_rn(__dmul_rn(-2.0 * pa, pb), __dadd_rn(__dmul_rn(pa, pa), __dmul_rn(pb, pb)));
  const double r = sqrt(fmax(r2, 1e-36));
  double q;
  if (TYPE == WV_LEAF_M12) { E = exp(-r); q = r; }
  else if (TYPE == WV_LEAF_M32) {
    const double s = 1.7320508075688772 * r;
    E = (1.0 + s) * exp(-s); q = 3.0 * r * r / (1.0 + s);
  } else {
    const double s = 2.23606797749979 * r, poly = 1.0 + s + 5.0 / 3.0 * r * r;
    E = poly * exp(-s); q = (5.0 / 3.0) * r * r * (1.0 + s) / poly;
  }
  q_ls = r2 > 1e-36 ? q : 0.0;
}
template <int TYPE>
__device__ __forceinline__ double wvs_matern_value(double pa, double pb) {
  double E, q;
  wvs_matern<TYPE>(pa, pb, E, q);
  return E;
}

// gpflow Periodic(SquaredExponential): exp(-0.5 (sin(pi (x - x') / p) / ell)^2); q_ls = (sin/ell)^2 (coefficient
// var / ell), q_per = (sin/ell) cos arg (coefficient var / (ell p))
__device__ __forceinline__ double wvs_periodic_value(double xi, double xj, double ell, double per) {
  const double arg = 3.141592653589793 * (xi - xj) / per;
  const double ss = sin(arg) / ell;
  return exp(-0.5 * (ss * ss));
}
__device__ __forceinline__ void wvs_periodic(double xi, double xj, double ell, double per, double& E, double& q_ls,
                                             double& q_per) {
  const double arg = 3.141592653589793 * (xi - xj) / per;
  double sn, cs;
  sincos(arg, &sn, &cs);
  const double ss = sn / ell;
  q_ls = ss * ss;
  E = exp(-0.5 * q_ls);
  q_per = (ss * cs) * arg;
}

// fixed-order reduction of the lanes' sums inside one warp: lane l holds sum k in red[k * 33 + l] (stride 33: the
// transposed read below is conflict free); lane k adds the 32 entries of row k in order
__device__ __forceinline__ double wvs_warp_row_sum(const double* __restrict__ red, int k) {
  const double* r = red + k * 33;
  double s = r[0];
#pragma unroll 8
  for (int j = 1; j < 32; ++j) s += r[j];
  return s;
}
