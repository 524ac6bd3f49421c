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
#define WVS_MR 4
#define WVS_NE 16

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

// 2^(j/4096) table: per-device copy in global memory (L2 resident) -> shared memory
__device__ __forceinline__ void wvs_load_tab(double* __restrict__ tab, const double* __restrict__ gtab) {
  const double2* src = reinterpret_cast<const double2*>(gtab);
  double2* dst = reinterpret_cast<double2*>(tab);
  for (int i = threadIdx.x; i < WV_EXP2_TAB12 / 2; i += WVS_THREADS) dst[i] = src[i];
}

// thread -> micro-tile (same map as wv_elem_coords): warp w owns rows (w>>1) * 16 .., cols (w&1) * 32 ..; lane l the
// 4 x 4 micro-tile at (+ (l>>3) 4, + (l&7) 4)
__device__ __forceinline__ void wvs_coords(int& r_off, int& c_off, bool& above_diag, bool diag_tile) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = (warp >> 1) * 16, wc = (warp & 1) * 32;
  r_off = wr + (lane >> 3) * 4;
  c_off = wc + (lane & 7) * 4;
  above_diag = diag_tile && wc > wr + 15;
}

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

// fixed-order reduction of per-thread sums over the CTA: red[k][tid] holds sum k of thread tid; warp w reduces the sums
// k = w, w + 8, ...: lane l adds red[k][l + 32 j], j = 0..7, in that order, then a shuffle tree; lane 0 writes out[k]
__device__ __forceinline__ void wvs_reduce_sums(const double* __restrict__ red, int nsum, double* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = warp; k < nsum; k += WVS_THREADS / 32) {
    const double* r = red + (size_t)k * WVS_THREADS + lane;
    double s = r[0];
#pragma unroll
    for (int j = 1; j < WVS_THREADS / 32; ++j) s += r[32 * j];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[k] = s;
  }
}
