// waveome_b200 — element-wise passes of one evaluation: the Gram builder and the fused gradient reduction.
//
//   gram   K = sum_c prod_f k_f(X[:,d_f]; theta) + sigma^2 I  (lower 64x64 tiles), RHS row d^T = (y - c)^T at row n,
//          identity padding                                          (reference: the kernel-tree ops of SURVEY §3.4)
//   grad   partial[b][tile][slot] = sum over the tile of wgt_ij W_ij dK_ij/dtheta_slot,  W = alpha alpha^T - K^{-1};
//          dK/dtheta is regenerated from the kernel program, nothing of size n^2 is materialised.
//
// Both interpret the flat kernel program (sum of products of leaves).  A CTA stages the program, the constrained
// parameters, per-leaf constants and the 2^(j/64) table ONCE and then walks `tpc` tiles of the same model; a
// thread owns an MR x 4 micro-tile, a warp a compact (8 MR) x 32 region, so that after the host has sorted the rows by
// their categorical columns a zero categorical mask usually covers whole warps and the transcendental factors of
// categorical x numeric products are skipped.  The passes are bound by FP64 issue (one 2^u per squared-exponential
// leaf and element), not by the 8 n^2 bytes they write / read; see DESIGN.md §4.
#pragma once
#include "wv_kernels.cuh"

// Micro-tile rows per thread and the resident CTAs per SM the register allocation aims for, per kernel.  Measured on
// config 3 (B = 2000, n = 600), gram / grad ms per evaluation: MR 4 x (2 CTAs: 4.67 / 6.08, 3 CTAs: 5.72 / 5.56,
// 4 CTAs: 7.90 / 5.96), MR 2 x (1 CTA: 5.89 / 9.22, 2 CTAs: 5.42 / 7.00, 3 CTAs: 10.7 / 8.86).
#ifndef WV_ELEM_MR
#define WV_ELEM_MR 4
#endif
#ifndef WV_GRAM_MINB
#define WV_GRAM_MINB 2
#endif
#ifndef WV_GRAD_MINB
#define WV_GRAD_MINB 3
#endif
#define WV_ELEM_NE (WV_ELEM_MR * 4)                         // elements per thread
#undef WV_ELEM_THREADS
#define WV_ELEM_THREADS (WV_NB * WV_NB / WV_ELEM_NE)        // 256
#define WV_ELEM_WARPS (WV_ELEM_THREADS / 32)
// Tiles walked by one CTA (it stages the model once): chosen per launch, 8 when the launch has work for many waves
// (staging is ~4 % of the Gram pass at 4 tiles), down to 1 when few models are active and parallelism matters more.
#define WV_ELEM_TPC_MAX 8
__host__ inline int wv_elem_tpc(long total_tiles) {
  long t = total_tiles / 2400;      // aim at >= ~8 CTAs per resident slot (148 SMs x 2)
  return (int)(t < 1 ? 1 : (t > WV_ELEM_TPC_MAX ? WV_ELEM_TPC_MAX : t));
}

struct WvElemSmem {
  int n_comp, n_leaves, n_slots, noise_slot, mean_slot, n_dims;
  unsigned comp_mask, pad1;
  int comp_start[WV_MAX_COMP + 1];
  int dims[WV_MAX_DIMS];
  int slot_x[WV_MAX_SLOTS];                // packed index of the slot, -1 if frozen
  WvLeaf leaves[WV_MAX_LEAVES];
  double theta[WV_MAX_SLOTS];
  double lc[WV_MAX_LEAVES];                // per-leaf constant: SE sqrt(log2(e)/2) / lengthscale
  double lv[WV_MAX_LEAVES];                // SE: log2(variance), folded into the exponent of the Gram value
  double tab[WV_EXP2_TAB];
  double xr[WV_MAX_DIMS][WV_NB];
  double xc[WV_MAX_DIMS][WV_NB];
  double red[WV_MAX_SLOTS][WV_ELEM_WARPS]; // per-warp partial sums (grad only)
};

// once per CTA: program, theta, per-leaf constants, exp2 table
__device__ __forceinline__ void wv_elem_stage_model(const WvBatchDev& bd, int b, const double* __restrict__ xall,
                                                    WvElemSmem& sm) {
  const WvProgram* gp = bd.programs + bd.prog_id[b];
  const int nl = gp->n_leaves, ns = gp->n_slots, nc = gp->n_comp;
  if (threadIdx.x == 0) {
    sm.n_comp = nc; sm.n_leaves = nl; sm.n_slots = ns; sm.noise_slot = gp->noise_slot; sm.mean_slot = gp->mean_slot;
    sm.n_dims = gp->n_dims;
    sm.comp_mask = bd.comp_mask[b];
  }
  for (int i = threadIdx.x; i <= nc; i += blockDim.x) sm.comp_start[i] = gp->comp_start[i];
  for (int i = threadIdx.x; i < WV_MAX_DIMS; i += blockDim.x) sm.dims[i] = gp->dims[i];
  {
    const int nw = nl * (int)(sizeof(WvLeaf) / 4);
    const int32_t* src = reinterpret_cast<const int32_t*>(gp->leaves);
    int32_t* dst = reinterpret_cast<int32_t*>(sm.leaves);
    for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
  }
  const double* x = xall + (size_t)b * bd.P;
  for (int s = threadIdx.x; s < ns; s += blockDim.x) {
    const WvSlot& sl = gp->slots[s];
    sm.slot_x[s] = sl.xindex;
    sm.theta[s] = sl.xindex >= 0 ? wv_transform(sl.transform, x[sl.xindex], sl.shift) : sl.fixed;
  }
  for (int j = threadIdx.x; j < WV_EXP2_TAB; j += blockDim.x) wv_exp2_table_entry(j, &sm.tab[j]);
  __syncthreads();
  for (int l = threadIdx.x; l < nl; l += blockDim.x) {
    const WvLeaf lf = sm.leaves[l];
    // exp(-r2/2) = 2^(-(s (x_i - x_j))^2),  s = sqrt(log2(e) / 2) / lengthscale
    sm.lc[l] = lf.type == WV_LEAF_SE ? 0.84932180028801907 / sm.theta[lf.s_ls] : 0.0;
    sm.lv[l] = lf.type == WV_LEAF_SE ? log2(sm.theta[lf.s_var]) : 0.0;
  }
}

// per tile: covariate columns of the tile's rows and columns
__device__ __forceinline__ void wv_elem_stage_tile(const WvBatchDev& bd, int ti, int tj, WvElemSmem& sm) {
  __syncthreads();      // previous tile done with xr / xc (and red), model staging visible
  for (int i = threadIdx.x; i < sm.n_dims * WV_NB; i += blockDim.x) {
    const int d = i / WV_NB, r = i % WV_NB;
    const double* col = bd.Xt + (size_t)sm.dims[d] * bd.npad;
    sm.xr[d][r] = col[ti * WV_NB + r];
    sm.xc[d][r] = col[tj * WV_NB + r];
  }
  __syncthreads();
}

// thread -> micro-tile: warp w owns rows (w>>1) * 4 MR .., cols (w&1) * 32 ..; lane l the MR x 4 micro-tile at
// (+ (l>>3) MR, + (l&7) 4).
__device__ __forceinline__ void wv_elem_coords(int& r_off, int& c_off, bool& above_diag, bool diag_tile) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = (warp >> 1) * (4 * WV_ELEM_MR), wc = (warp & 1) * 32;
  r_off = wr + (lane >> 3) * WV_ELEM_MR;
  c_off = wc + (lane & 7) * 4;
  above_diag = diag_tile && wc > wr + 4 * WV_ELEM_MR - 1;     // every element of the warp's region has col > row
}

__device__ __forceinline__ bool wv_warp_all_zero_ne(const double (&v)[WV_ELEM_NE]) {
  bool nz = false;
#pragma unroll
  for (int e = 0; e < WV_ELEM_NE; ++e) nz |= (v[e] != 0.0);
  return !__any_sync(0xffffffffu, nz);
}

// Leaves with library transcendentals (Matern, periodic) or rare ones (polynomial, empty): kept out of line so that
// their register needs do not set the allocation of the whole kernel (the common squared-exponential / categorical /
// linear / constant leaves are inlined below).  prod[e] = (first ? 1 : prod[e]) * k(x_i[a], x_j[b]),  e = a*4+b.
__device__ __noinline__ void wv_leaf_mul_slow(const WvLeaf lf, const double* __restrict__ theta, bool first,
                                              const double* __restrict__ xi, const double* __restrict__ xj,
                                              double* __restrict__ prod) {
  const double var = lf.s_var >= 0 ? theta[lf.s_var] : 1.0;
  for (int a = 0; a < WV_ELEM_MR; ++a)
    for (int b = 0; b < 4; ++b) {
      double v;
      switch (lf.type) {
        case WV_LEAF_M12:
        case WV_LEAF_M32:
        case WV_LEAF_M52: {
          // gpflow: r = sqrt(max(r2, 1e-36)), r2 = |a|^2 + |b|^2 - 2ab on a = x / ell (exactly zero on the diagonal)
          const double inv = 1.0 / theta[lf.s_ls];
          const double pa = xi[a] * inv, pb = xj[b] * inv;
          const double r2 = __dadd_rn(__dmul_rn(-2.0 * pa, pb), __dadd_rn(__dmul_rn(pa, pa), __dmul_rn(pb, pb)));
          const double r = sqrt(fmax(r2, 1e-36));
          if (lf.type == WV_LEAF_M12) v = exp(-r);
          else if (lf.type == WV_LEAF_M32) { const double q = 1.7320508075688772 * r; v = (1.0 + q) * exp(-q); }
          else { const double q = 2.23606797749979 * r; v = (1.0 + q + 5.0 / 3.0 * r * r) * exp(-q); }
          v *= var;
        } break;
        case WV_LEAF_PERIODIC: {
          const double ell = theta[lf.s_ls], per = theta[lf.s_aux];
          const double arg = 3.141592653589793 * (xi[a] - xj[b]) / per;
          const double ss = sin(arg) / ell;
          v = var * exp(-0.5 * (ss * ss));
        } break;
        case WV_LEAF_POLY: v = wv_powi(var * (xi[a] * xj[b]) + theta[lf.s_ls], lf.degree); break;
        default: v = 0.0; break;
      }
      prod[a * 4 + b] = first ? v : prod[a * 4 + b] * v;
    }
}

template <bool FIRST>
__device__ __forceinline__ void wv_leaf_mul_ne(const WvElemSmem& sm, int l, const double (&xi)[WV_ELEM_MR],
                                               const double (&xj)[4], double (&prod)[WV_ELEM_NE]) {
  const WvLeaf lf = sm.leaves[l];
  const double var = lf.s_var >= 0 ? sm.theta[lf.s_var] : 1.0;
#define WV_PUT(e, v) prod[e] = FIRST ? (v) : prod[e] * (v)
  switch (lf.type) {
    case WV_LEAF_SE: {
      // variance * exp(-r2 / 2) = 2^(log2(variance) - (s d)^2): the variance rides in the exponent
      const double s = sm.lc[l], lv = sm.lv[l];
      double ai[WV_ELEM_MR], aj[4];
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a) ai[a] = xi[a] * s;
#pragma unroll
      for (int b = 0; b < 4; ++b) aj[b] = xj[b] * s;
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const double d = ai[a] - aj[b];
          WV_PUT(a * 4 + b, wv_exp2_lo(fma(-d, d, lv), sm.tab));
        }
    } break;
    case WV_LEAF_LINEAR:
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) WV_PUT(a * 4 + b, var * (xi[a] * xj[b]));
      break;
    case WV_LEAF_CONST:
#pragma unroll
      for (int e = 0; e < WV_ELEM_NE; ++e) WV_PUT(e, var);
      break;
    case WV_LEAF_CAT: {
      double ci[WV_ELEM_MR], cj[4];
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a) ci[a] = rint(xi[a]);
#pragma unroll
      for (int b = 0; b < 4; ++b) cj[b] = rint(xj[b]);
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) WV_PUT(a * 4 + b, ci[a] == cj[b] ? var : 0.0);
    } break;
    default: {   // address-taken copies, so that prod / xi / xj themselves stay in registers
      double tp[WV_ELEM_NE], txi[WV_ELEM_MR], txj[4];
#pragma unroll
      for (int e = 0; e < WV_ELEM_NE; ++e) tp[e] = FIRST ? 1.0 : prod[e];
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a) txi[a] = xi[a];
#pragma unroll
      for (int b = 0; b < 4; ++b) txj[b] = xj[b];
      wv_leaf_mul_slow(lf, sm.theta, FIRST, txi, txj, tp);
#pragma unroll
      for (int e = 0; e < WV_ELEM_NE; ++e) prod[e] = tp[e];
    } break;
  }
#undef WV_PUT
}

// gradient sums of one leaf over the micro-tile: s[0..2] = sum_e wo[e] * d k_e / d (variance, lengthscale|offset, period)
// (wo = W weight times the product of the other leaves of the component); out-of-line part, see wv_leaf_mul_slow
__device__ __noinline__ void wv_leaf_grad_sums_slow(const WvLeaf lf, const double* __restrict__ theta,
                                                    const double* __restrict__ xi, const double* __restrict__ xj,
                                                    const double* __restrict__ wo, double* __restrict__ s) {
  double s_var = 0.0, s_ls = 0.0, s_aux = 0.0;
  const double var = lf.s_var >= 0 ? theta[lf.s_var] : 1.0;
  for (int a = 0; a < WV_ELEM_MR; ++a)
    for (int b = 0; b < 4; ++b) {
      const double w = wo[a * 4 + b];
      switch (lf.type) {
        case WV_LEAF_M12:
        case WV_LEAF_M32:
        case WV_LEAF_M52: {
          const double inv = 1.0 / theta[lf.s_ls];
          const double pa = xi[a] * inv, pb = xj[b] * inv;
          const double r2 = __dadd_rn(__dmul_rn(-2.0 * pa, pb), __dadd_rn(__dmul_rn(pa, pa), __dmul_rn(pb, pb)));
          const double r = sqrt(fmax(r2, 1e-36));
          double e, de;  // de = dE/dr
          if (lf.type == WV_LEAF_M12) { e = exp(-r); de = -e; }
          else if (lf.type == WV_LEAF_M32) {
            const double q = 1.7320508075688772 * r, ex = exp(-q);
            e = (1.0 + q) * ex; de = -3.0 * r * ex;
          } else {
            const double q = 2.23606797749979 * r, ex = exp(-q);
            e = (1.0 + q + 5.0 / 3.0 * r * r) * ex; de = -(5.0 / 3.0) * r * (1.0 + q) * ex;
          }
          s_var += w * e;
          if (r2 > 1e-36) s_ls += w * de * (-r);
        } break;
        case WV_LEAF_PERIODIC: {
          const double ell = theta[lf.s_ls], per = theta[lf.s_aux];
          const double arg = 3.141592653589793 * (xi[a] - xj[b]) / per;
          double sn, cs;
          sincos(arg, &sn, &cs);
          const double ss = sn / ell;
          const double r2 = ss * ss;
          const double t = w * exp(-0.5 * r2);
          s_var += t;
          s_ls = fma(t, r2, s_ls);
          s_aux += t * (ss * cs) * arg;
        } break;
        case WV_LEAF_POLY: {
          const double xx = xi[a] * xj[b];
          const double db = lf.degree * wv_powi(var * xx + theta[lf.s_ls], lf.degree - 1);
          s_var += w * db * xx;
          s_ls += w * db;
        } break;
        default: break;
      }
    }
  if (lf.type == WV_LEAF_M12 || lf.type == WV_LEAF_M32 || lf.type == WV_LEAF_M52) s_ls *= var / theta[lf.s_ls];
  if (lf.type == WV_LEAF_PERIODIC) {
    const double ell = theta[lf.s_ls], per = theta[lf.s_aux];
    s_ls *= var / ell;
    s_aux *= var / (ell * per);
  }
  s[0] = s_var; s[1] = s_ls; s[2] = s_aux;
}

__device__ __forceinline__ void wv_leaf_grad_sums_ne(const WvElemSmem& sm, int l, const double (&xi)[WV_ELEM_MR],
                                                     const double (&xj)[4], const double (&wo)[WV_ELEM_NE],
                                                     double& s_var, double& s_ls, double& s_aux) {
  const WvLeaf lf = sm.leaves[l];
  s_var = 0.0; s_ls = 0.0; s_aux = 0.0;
  switch (lf.type) {
    case WV_LEAF_SE: {
      const double var = sm.theta[lf.s_var], ell = sm.theta[lf.s_ls], s = sm.lc[l];
      double ai[WV_ELEM_MR], aj[4];
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a) ai[a] = xi[a] * s;
#pragma unroll
      for (int b = 0; b < 4; ++b) aj[b] = xj[b] * s;
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const double d = ai[a] - aj[b];
          const double u = d * d;                                   // = r2 log2(e) / 2
          const double t = wo[a * 4 + b] * wv_exp2_lo(-u, sm.tab);
          s_var += t;
          s_ls = fma(t, u, s_ls);
        }
      s_ls *= 1.3862943611198906 * var / ell;                       // r2 = 2 ln2 u;  dk/dell = var e r2 / ell
    } break;
    case WV_LEAF_LINEAR:
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) s_var += wo[a * 4 + b] * (xi[a] * xj[b]);
      break;
    case WV_LEAF_CONST:
#pragma unroll
      for (int e = 0; e < WV_ELEM_NE; ++e) s_var += wo[e];
      break;
    case WV_LEAF_CAT: {
      double ci[WV_ELEM_MR], cj[4];
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a) ci[a] = rint(xi[a]);
#pragma unroll
      for (int b = 0; b < 4; ++b) cj[b] = rint(xj[b]);
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) s_var += ci[a] == cj[b] ? wo[a * 4 + b] : 0.0;
    } break;
    default: {
      double s3[3], tw[WV_ELEM_NE], txi[WV_ELEM_MR], txj[4];
#pragma unroll
      for (int e = 0; e < WV_ELEM_NE; ++e) tw[e] = wo[e];
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a) txi[a] = xi[a];
#pragma unroll
      for (int b = 0; b < 4; ++b) txj[b] = xj[b];
      wv_leaf_grad_sums_slow(lf, sm.theta, txi, txj, tw, s3);
      s_var = s3[0]; s_ls = s3[1]; s_aux = s3[2];
    } break;
  }
}

__device__ __forceinline__ void wv_elem_load_x(const WvElemSmem& sm, int dim, int r_off, int c_off,
                                               double (&xi)[WV_ELEM_MR], double (&xj)[4]) {
#pragma unroll
  for (int a = 0; a < WV_ELEM_MR; ++a) xi[a] = sm.xr[dim][r_off + a];
#pragma unroll
  for (int b = 0; b < 4; ++b) xj[b] = sm.xc[dim][c_off + b];
}

// acc[e] = sum over the enabled components of the product of their leaves on this thread's micro-tile, covariates
// taken from sm.xr (rows) / sm.xc (columns).  Whole warps skip the transcendental factors of a product whose partial
// product (a categorical mask) is zero for the warp.
__device__ __forceinline__ void wv_elem_eval_kernel_tree(const WvElemSmem& sm, int r_off, int c_off,
                                                         double (&acc)[WV_ELEM_NE]) {
#pragma unroll
  for (int e = 0; e < WV_ELEM_NE; ++e) acc[e] = 0.0;
  for (int c = 0; c < sm.n_comp; ++c) {
    if (!((sm.comp_mask >> c) & 1u)) continue;
    double prod[WV_ELEM_NE];
    const int l0 = sm.comp_start[c], l1 = sm.comp_start[c + 1];
    bool skip = l1 <= l0;
    for (int l = l0; l < l1; ++l) {
      const int type = sm.leaves[l].type;
      if (l > l0 && !wv_leaf_is_cheap(type) && wv_warp_all_zero_ne(prod)) { skip = true; break; }
      double xi[WV_ELEM_MR], xj[4];
      wv_elem_load_x(sm, sm.leaves[l].dim, r_off, c_off, xi, xj);
      if (l == l0) wv_leaf_mul_ne<true>(sm, l, xi, xj, prod);
      else wv_leaf_mul_ne<false>(sm, l, xi, xj, prod);
    }
    if (!skip) {
#pragma unroll
      for (int e = 0; e < WV_ELEM_NE; ++e) acc[e] += prod[e];
    }
  }
}

// =============================================================================================
// gram.  grid (ceil(n_lower_tiles / tpc), n_active), WV_ELEM_THREADS threads.
// Algorithmic traffic: 8 n^2 bytes written (lower tiles actually written: ~4 n^2).
// =============================================================================================
__global__ void __launch_bounds__(WV_ELEM_THREADS, WV_GRAM_MINB) wv_gram_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                                     const double* __restrict__ xall, int ntiles, int tpc) {
  WvElemSmem& sm = *reinterpret_cast<WvElemSmem*>(wv_smem_raw);
  const int b = active[blockIdx.y];
  wv_elem_stage_model(bd, b, xall, sm);
  const int n = bd.n;
  double* Ab = bd.A + (size_t)b * bd.npad * bd.npad;
  const double* yb = bd.Y + (size_t)b * bd.npad;
  // variational path: the "observations" are the Gaussian sites (mean eta/lam, noise variance jitter + 1/lam)
  const double* lam = bd.site_lam ? bd.site_lam + (size_t)b * bd.npad : nullptr;
  const double* eta = bd.site_eta ? bd.site_eta + (size_t)b * bd.npad : nullptr;
  const int t1 = min(ntiles, (int)(blockIdx.x + 1) * tpc);
  for (int t = blockIdx.x * tpc; t < t1; ++t) {
    int ti, tj;
    wv_tile_from_linear(t, ti, tj);
    wv_elem_stage_tile(bd, ti, tj, sm);
    int r_off, c_off;
    bool above;
    wv_elem_coords(r_off, c_off, above, ti == tj);
    if (above) continue;                          // the strict upper part of a diagonal tile is never read
    double acc[WV_ELEM_NE];
    wv_elem_eval_kernel_tree(sm, r_off, c_off, acc);
    const double s2 = sm.theta[sm.noise_slot];
    const double cmean = sm.mean_slot >= 0 ? sm.theta[sm.mean_slot] : 0.0;
#pragma unroll
    for (int a = 0; a < WV_ELEM_MR; ++a) {
      const int gi = ti * WV_NB + r_off + a;
      double out[4];
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int gj = tj * WV_NB + c_off + bb;
        double v;
        if (gi < n && gj < n) v = acc[a * 4 + bb] + (gi == gj ? (lam ? bd.jitter + 1.0 / lam[gi] : s2) : 0.0);
        else if (gi == n && gj < n) v = (lam ? eta[gj] / lam[gj] : yb[gj]) - cmean;     // RHS row d^T
        else v = (gi == gj) ? 1.0 : 0.0;                     // identity padding (incl. A[n][n] = 1)
        out[bb] = v;
      }
      double2* dst = reinterpret_cast<double2*>(Ab + (size_t)gi * bd.npad + tj * WV_NB + c_off);
      dst[0] = make_double2(out[0], out[1]);
      dst[1] = make_double2(out[2], out[3]);
    }
  }
}

// =============================================================================================
// grad.  grid (ceil(n_lower_tiles / tpc), n_active), WV_ELEM_THREADS threads.
//   wgt = 2 below the diagonal, 1 on it, 0 above / outside [0,n).  Algorithmic traffic: 8 n^2 bytes read.
// =============================================================================================
__global__ void __launch_bounds__(WV_ELEM_THREADS, WV_GRAD_MINB) wv_grad_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                                     const double* __restrict__ xall, int ntiles, int tpc) {
  WvElemSmem& sm = *reinterpret_cast<WvElemSmem*>(wv_smem_raw);
  const int b = active[blockIdx.y];
  wv_elem_stage_model(bd, b, xall, sm);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = bd.n, ld = bd.npad;
  const double* Kb = bd.A + (size_t)b * ld * ld;
  const double* al = bd.alpha + (size_t)b * ld;
  const int t1 = min(ntiles, (int)(blockIdx.x + 1) * tpc);
  for (int t = blockIdx.x * tpc; t < t1; ++t) {
    int ti, tj;
    wv_tile_from_linear(t, ti, tj);
    wv_elem_stage_tile(bd, ti, tj, sm);
    for (int i = threadIdx.x; i < WV_MAX_SLOTS * WV_ELEM_WARPS; i += blockDim.x) (&sm.red[0][0])[i] = 0.0;
    __syncthreads();
    int r_off, c_off;
    bool above;
    wv_elem_coords(r_off, c_off, above, ti == tj);
    if (!above) {
      double w[WV_ELEM_NE];
      double trw = 0.0;
      double aj[4];
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) aj[bb] = al[tj * WV_NB + c_off + bb];
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a) {
        const int gi = ti * WV_NB + r_off + a;
        const double2* src = reinterpret_cast<const double2*>(Kb + (size_t)gi * ld + tj * WV_NB + c_off);
        const double2 k01 = src[0], k23 = src[1];
        const double kin[4] = {k01.x, k01.y, k23.x, k23.y};
        const double ai = al[gi];
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          const int gj = tj * WV_NB + c_off + bb;
          const double wv = ai * aj[bb] - kin[bb];
          const bool in = gi < n && gj < n;
          w[a * 4 + bb] = in ? (gi > gj ? 2.0 * wv : (gi == gj ? wv : 0.0)) : 0.0;
          if (gi == gj && gi < n) trw += wv;
        }
      }
      for (int c = 0; c < sm.n_comp; ++c) {
        if (!((sm.comp_mask >> c) & 1u)) continue;
        const int l0 = sm.comp_start[c], l1 = sm.comp_start[c + 1];
        for (int l = l0; l < l1; ++l) {
          const WvLeaf lf = sm.leaves[l];
          const bool tv = lf.s_var >= 0 && sm.slot_x[lf.s_var] >= 0;
          const bool tl = lf.s_ls >= 0 && sm.slot_x[lf.s_ls] >= 0;
          const bool ta = lf.s_aux >= 0 && sm.slot_x[lf.s_aux] >= 0;
          if (!(tv || tl || ta)) continue;
          double wo[WV_ELEM_NE];
#pragma unroll
          for (int e = 0; e < WV_ELEM_NE; ++e) wo[e] = w[e];
          bool skip = false;
          for (int l2 = l0; l2 < l1; ++l2) {
            if (l2 == l) continue;
            const int type2 = sm.leaves[l2].type;
            if (!wv_leaf_is_cheap(type2) && wv_warp_all_zero_ne(wo)) { skip = true; break; }
            double xi[WV_ELEM_MR], xj[4];
            wv_elem_load_x(sm, sm.leaves[l2].dim, r_off, c_off, xi, xj);
            wv_leaf_mul_ne<false>(sm, l2, xi, xj, wo);
          }
          if (skip || (!wv_leaf_is_cheap(lf.type) && wv_warp_all_zero_ne(wo))) continue;   // every contribution is zero
          double xi[WV_ELEM_MR], xj[4];
          wv_elem_load_x(sm, lf.dim, r_off, c_off, xi, xj);
          double sv, sl, sa;
          wv_leaf_grad_sums_ne(sm, l, xi, xj, wo, sv, sl, sa);
          for (int o = 16; o > 0; o >>= 1) {
            sv += __shfl_xor_sync(0xffffffffu, sv, o);
            sl += __shfl_xor_sync(0xffffffffu, sl, o);
            if (ta) sa += __shfl_xor_sync(0xffffffffu, sa, o);
          }
          if (lane == 0) {   // several leaves may share a slot: accumulate (warp-private column, no race)
            if (tv) sm.red[lf.s_var][warp] += sv;
            if (tl) sm.red[lf.s_ls][warp] += sl;
            if (ta) sm.red[lf.s_aux][warp] += sa;
          }
        }
      }
      for (int o = 16; o > 0; o >>= 1) trw += __shfl_xor_sync(0xffffffffu, trw, o);
      if (lane == 0) sm.red[sm.noise_slot][warp] += trw;
    }
    __syncthreads();
    double* dst = bd.partial + ((size_t)b * ntiles + t) * bd.n_slots_max;
    for (int s = threadIdx.x; s < sm.n_slots; s += blockDim.x) {
      double tt = 0.0;
#pragma unroll
      for (int wq = 0; wq < WV_ELEM_WARPS; ++wq) tt += sm.red[s][wq];
      dst[s] = tt;
    }
  }
}


// =============================================================================================
// cross mean: mean[b][i] = c_b + sum_j k_b(xnew_i, x_j) alpha_b[j]   (posterior mean of gpflow GPR.predict_f at new
// inputs; alpha = (K + sigma^2 I)^{-1} (y - c) is left in bd.alpha by the last evaluation).
// grid (ceil(m / 64), n_models), WV_ELEM_THREADS threads; one CTA owns 64 new points and walks the training tiles.
// Xnew_t: [D][mpad] column-major, zero padded.  Fixed-order reductions (bit-reproducible).
// =============================================================================================
__global__ void __launch_bounds__(WV_ELEM_THREADS, WV_GRAM_MINB) wv_cross_mean_kernel(
    WvBatchDev bd, const double* __restrict__ xall, const double* __restrict__ Xnew_t, int m, int mpad,
    double* __restrict__ mean) {
  WvElemSmem& sm = *reinterpret_cast<WvElemSmem*>(wv_smem_raw);
  const int b = blockIdx.y, ti = blockIdx.x;
  wv_elem_stage_model(bd, b, xall, sm);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* al = bd.alpha + (size_t)b * bd.npad;
  int r_off, c_off;
  bool above;
  wv_elem_coords(r_off, c_off, above, false);
  double rs[WV_ELEM_MR];
#pragma unroll
  for (int a = 0; a < WV_ELEM_MR; ++a) rs[a] = 0.0;
  for (int tj = 0; tj < bd.nt; ++tj) {
    __syncthreads();
    for (int i = threadIdx.x; i < sm.n_dims * WV_NB; i += blockDim.x) {
      const int d = i / WV_NB, r = i % WV_NB;
      sm.xr[d][r] = Xnew_t[(size_t)sm.dims[d] * mpad + ti * WV_NB + r];
      sm.xc[d][r] = bd.Xt[(size_t)sm.dims[d] * bd.npad + tj * WV_NB + r];
    }
    __syncthreads();
    double acc[WV_ELEM_NE];
    wv_elem_eval_kernel_tree(sm, r_off, c_off, acc);
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int gj = tj * WV_NB + c_off + bb;
      const double aj = gj < bd.n ? al[gj] : 0.0;
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a) rs[a] = fma(acc[a * 4 + bb], aj, rs[a]);
    }
  }
  // the 8 lanes (lane & 7) of a row group, then the two warps that share the rows
  __syncthreads();
#pragma unroll
  for (int a = 0; a < WV_ELEM_MR; ++a) {
    double v = rs[a];
    for (int o = 1; o < 8; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((lane & 7) == 0) sm.red[r_off + a][warp & 1] = v;
  }
  __syncthreads();
  const double cmean = sm.mean_slot >= 0 ? sm.theta[sm.mean_slot] : 0.0;
  for (int r = threadIdx.x; r < WV_NB; r += blockDim.x) {
    const int gi = ti * WV_NB + r;
    if (gi < m) mean[(size_t)b * m + gi] = cmean + (sm.red[r][0] + sm.red[r][1]);
  }
}


// =============================================================================================
// cross variance: var f(xnew_i) = k(xnew_i, xnew_i) - k*_i^T W k*_i,  W = (K + sigma^2 I)^-1 (lower tiles of A after an
// evaluation), k*_i = k(xnew_i, X)  (gpflow GPR.predict_f variance, full_cov = False).
// grid (n_lower_tiles, ceil(m / 64), n_models), WV_ELEM_THREADS threads: CTA (t, ti, b) forms the two cross-Gram tiles
// K*_j = k(xnew[ti], X[tj]) and K*_k = k(xnew[ti], X[tk]) of the W tile (tj >= tk) in shared memory and reduces
//     part[b][t][i] = w sum_{j,k} K*_j[i][j] W[j][k] K*_k[i][k],   w = 2 off the diagonal, 1 on it;
// the tile t = 0 CTA also writes the prior variance k(xnew_i, xnew_i).  wv_cross_var_reduce_kernel sums the parts
// in fixed order.  A post-fit pass: recomputing the cross-Gram tiles per W tile keeps it a single simple kernel.
// =============================================================================================
struct WvCrossVarSmem {
  WvElemSmem el;
  double Kj[WV_NB * (WV_NB + 1)];
  double Kk[WV_NB * (WV_NB + 1)];
  double W[WV_NB * (WV_NB + 1)];
};

__global__ void __launch_bounds__(WV_ELEM_THREADS, 1) wv_cross_var_kernel(
    WvBatchDev bd, const double* __restrict__ xall, const double* __restrict__ Xnew_t, int m, int mpad,
    double* __restrict__ part, double* __restrict__ prior) {
  WvCrossVarSmem& cs = *reinterpret_cast<WvCrossVarSmem*>(wv_smem_raw);
  WvElemSmem& sm = cs.el;
  const int b = blockIdx.z, ti = blockIdx.y, t = blockIdx.x;
  const int ntiles = gridDim.x, LDK = WV_NB + 1;
  int tj, tk;
  wv_tile_from_linear(t, tj, tk);
  wv_elem_stage_model(bd, b, xall, sm);
  int r_off, c_off;
  bool above;
  wv_elem_coords(r_off, c_off, above, false);
  // ---- the two cross-Gram tiles (columns beyond n are zero) and, for t = 0, the prior variances
  for (int which = 0; which < (t == 0 ? 3 : 2); ++which) {
    const int tc = which == 0 ? tj : tk;
    __syncthreads();
    for (int i = threadIdx.x; i < sm.n_dims * WV_NB; i += blockDim.x) {
      const int d = i / WV_NB, r = i % WV_NB;
      const double xr = Xnew_t[(size_t)sm.dims[d] * mpad + ti * WV_NB + r];
      sm.xr[d][r] = xr;
      sm.xc[d][r] = which == 2 ? xr : bd.Xt[(size_t)sm.dims[d] * bd.npad + tc * WV_NB + r];
    }
    __syncthreads();
    double acc[WV_ELEM_NE];
    wv_elem_eval_kernel_tree(sm, r_off, c_off, acc);
    if (which == 2) {
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a)
#pragma unroll
        for (int bb = 0; bb < 4; ++bb)
          if (r_off + a == c_off + bb && ti * WV_NB + r_off + a < m)
            prior[(size_t)b * m + ti * WV_NB + r_off + a] = acc[a * 4 + bb];
    } else {
      double* dst = which == 0 ? cs.Kj : cs.Kk;
#pragma unroll
      for (int a = 0; a < WV_ELEM_MR; ++a)
#pragma unroll
        for (int bb = 0; bb < 4; ++bb)
          dst[(r_off + a) * LDK + c_off + bb] = (tc * WV_NB + c_off + bb < bd.n) ? acc[a * 4 + bb] : 0.0;
    }
  }
  // ---- the W tile
  const double* Wg = bd.A + (size_t)b * bd.npad * bd.npad + (size_t)tj * WV_NB * bd.npad + tk * WV_NB;
  for (int i = threadIdx.x; i < WV_NB * WV_NB; i += blockDim.x) {
    const int r = i >> 6, c = i & 63;
    // diagonal tiles of K^-1 hold their lower 8x8 blocks only (wv_kinv_kernel): mirror
    cs.W[r * LDK + c] = (tj == tk && (c >> 3) > (r >> 3)) ? Wg[(size_t)c * bd.npad + r] : Wg[(size_t)r * bd.npad + c];
  }
  __syncthreads();
  // ---- thread (i = tid / 4, q = tid % 4): T[i][k] = sum_j K*_j[i][j] W[j][k] for k in [16 q, 16 q + 16), then the dot
  const int i = threadIdx.x >> 2, q = threadIdx.x & 3;
  double tacc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) tacc[k] = 0.0;
  for (int j = 0; j < WV_NB; ++j) {
    const double a = cs.Kj[i * LDK + j];
#pragma unroll
    for (int k = 0; k < 16; ++k) tacc[k] = fma(a, cs.W[j * LDK + q * 16 + k], tacc[k]);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s = fma(tacc[k], cs.Kk[i * LDK + q * 16 + k], s);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if (q == 0 && ti * WV_NB + i < m) part[((size_t)b * ntiles + t) * m + ti * WV_NB + i] = (tj == tk ? 1.0 : 2.0) * s;
}

__global__ void wv_cross_var_reduce_kernel(int B, int ntiles, int m, const double* __restrict__ part,
                                           const double* __restrict__ prior, double* __restrict__ var) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * m) return;
  const size_t b = idx / m, i = idx % m;
  double s = 0.0;
  for (int t = 0; t < ntiles; ++t) s += part[(b * ntiles + t) * m + i];
  var[idx] = prior[idx] - s;
}
