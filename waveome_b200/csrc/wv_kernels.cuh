// waveome_b200 — sm_100a kernels for one batched LML+gradient evaluation (objective A, exact GPR).
//
// Per evaluation of a batch of models that share X [n, D]:
//   gram      K = sum_c prod_f k_f(X[:,d_f]; theta) + sigma^2 I       (+ RHS row d^T = (y - c)^T at row n)
//   chol      blocked left-looking Cholesky, 64x64 tiles, FP64 DMMA (mma.sync.m8n8k4.f64) trailing updates
//   trtri     Mt = L^{-T} (upper) by block rows; the augmented RHS row yields z = L^{-1} d and -alpha for free
//   kinv      K^{-1} = Mt Mt^T  (lower tiles, DMMA)
//   grad      0.5 * sum_ij (alpha_i alpha_j - K^{-1}_ij) dK_ij/dtheta_p, dK regenerated on the fly
//   finalize  LML, priors, chain rule through the bijectors -> f = -(LML + log prior), df/dx
//
// Layout in HBM, per model b (npad = 64 * ceil((n + 1) / 64), ld = npad):
//   A [npad x npad]  lower tiles: K + sigma^2 I  ->  L  ->  K^{-1};  row n holds d^T -> z^T
//   Mt[npad x npad]  upper tiles: L^{-T};  column n holds -alpha before `extract` zeroes it
//   Dinv[nt][64x64]  row-major inverses of the diagonal Cholesky blocks
// Rows/cols in (n, npad) carry an identity so that no kernel needs ragged-edge special cases.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif
#include "wv_common.cuh"

#define WV_NB 64
#define WV_BK 16
#define WV_LDS (WV_BK + 4)   // smem row stride (doubles) of a pipeline stage: (row*20 + k) mod 16 distinct -> conflict-free DMMA fragment loads
#define WV_LDT 68            // smem row stride of a 64x64 DMMA operand tile (68 mod 16 == 4)
#define WV_LDP 65            // smem row stride of a 64x64 scalar-access tile (potrf / trtri of a diagonal block)
// cp.async operand ring of the 64x64 tile GEMM.  Two stages (41 KB): five CTAs per SM for the kernels that only need the
// ring (kinv, syrk, trtri levels) -- measured against three stages (61 KB, three CTAs): kinv 5.79 -> 5.37 ms per
// 2000-model evaluation, n = 8192 Cholesky 8.97 -> 8.71 ms; occupancy beats pipeline depth here.
#ifndef WV_STAGES
#define WV_STAGES 2
#endif
#define WV_GEMM_THREADS 128
#define WV_ELEM_THREADS 256

struct WvBatchDev {
  int n, D, B, npad, nt, n8, P;     // P = stride of x / grad (max packed params)
  int n_slots_max;                  // stride of theta / partial sums
  const double* Xt;                 // [D][npad] covariates, column-major, zero padded
  const double* Y;                  // [B][npad] outcomes, zero padded
  const WvProgram* programs;        // [n_programs]
  const int* prog_id;               // [B]
  double* A;                        // [B][npad][npad]
  double* Mt;                       // [B][npad][npad]
  double* Dinv;                     // [B][nt][64][64]
  double* alpha;                    // [B][npad]
  double* logdet_part;              // [B][nt]
  double* quad;                     // [B]
  double* partial;                  // [B][n_tiles][n_slots_max]
  int* chol_fail;                   // [B]
  const unsigned* comp_mask;        // [B] bit c = additive component c of the model's program takes part (default all)
  int* step_flag;                   // [2][B][nt] + 1: epoch of the last finished diagonal block (fused Cholesky step), then the
                                    // per-row column counters of the fused large-n panel (wv_chol_panel_fused_kernel), then its work counter
  // variational path for count likelihoods (nullptr / 0 on the Gaussian path), see wv_site_update_kernel:
  int lik;                          // 0 gaussian, 1 poisson (exp link), 2 negative binomial (log link, fixed alpha)
  double lik_param;                 // negative binomial: alpha when the programs' noise slot is frozen; a trainable
                                    // noise slot (Exp bijector) IS alpha on this path (waveome/likelihoods.py:24-28)
  double jitter;                    // gpflow default_jitter() added to K on the variational path
  double* site_lam;                 // [B][npad] precision of the Gaussian pseudo-observation of every row
  double* site_eta;                 // [B][npad] precision x mean of the pseudo-observation
  double* vgp_extra;                // [B] sum_i E_i + 1/2 log(2 pi / lam_i) + lam_i/2 ((ytilde_i - m_i)^2 + v_i)
  double lik_param2;                // ZINB: km when the programs carry no (trainable) second likelihood slot
  double* vgp_dlik2;                // [B] sum_i dE_i/d(km)
  double* vgp_dlik;                 // [B] sum_i dE_i/d(alpha): gradient of the bound wrt the negative-binomial dispersion
};

// per-model state of the site iteration (device arrays owned by the batch)
struct WvVgpState {
  double *lam_p, *eta_p, *lam_t, *eta_t;   // [B][npad] previous accepted sites, their targets
  double *F_prev, *rho;                    // [B]
  double *fmean, *fvar;                    // [B][npad] posterior mean / variance of f at the training inputs
  const double* lgam;                      // [B][npad] lgamma(y + 1)
  int *first, *inner_task, *sweeps, *good; // [B]
  int* at_bound;                           // [B] sites at the lower precision bound in the last sweep
  double tol, soft_tol;                    // converged below tol; at the sweep cap accepted without a flag below soft_tol
  int max_sweeps;
};

// ---------------------------------------------------------------------------------------------
// optional per-kernel-class timing with CUDA events on the launching stream (bench.py's live roofline)
// ---------------------------------------------------------------------------------------------
enum WvKernelClass { WV_K_GRAM = 0, WV_K_CHOL_DIAG, WV_K_CHOL_PANEL, WV_K_TRTRI, WV_K_EXTRACT, WV_K_KINV, WV_K_GRAD,
                     WV_K_FINALIZE, WV_K_LBFGS, WV_K_CHOL_SYRK, WV_K_SITES, WV_K_NCLASS };

#ifndef __CUDACC_RTC__     // host-side launch bookkeeping: not part of the NVRTC prelude
// second stream + events of the large-n path (look-ahead: the next panel is factorised while the bulk of the trailing
// update of the current one still runs), and the tile count from which that path is taken
struct WvAux {
  cudaStream_t side = nullptr;
  cudaEvent_t ev_panel = nullptr, ev_bulk = nullptr;
  int big_nt = 16;
  int panel_tiles = 4;       // tile columns per panel of the large-n right-looking Cholesky
  int resident_ctas = 444;   // 3 CTAs x 148 SMs: a Cholesky step is fused into one launch only if it fits
  const void* tmap_mt = nullptr;   // CUtensorMap of the current batch's Mt (kinv through a TMA operand ring, WV_KINV_TMA=1)
  int panel_ctas = 148;      // persistent CTAs of that launch (WV_PANEL_CTAS; default one per SM)
  int panel_fused = 1;       // large-n path: all column steps of a panel in one launch (flags instead of launch boundaries)
  int trtri_rows = 0;   // 1: the batched schedule's triangular inverse as one row-wise launch (wv_trtri_rows_kernel)
  int chol_all = 2;     // nt < big_nt: the whole Cholesky in one persistent launch (wv_chol_all_kernel).  0: never (beyond the
                        // few-models case), 1: always, 2: for a batch that has the device to itself (wv_batch_set_solo)
                        // while n_active * nt <= chol_all_max (WV_CHOL_ALL, WV_CHOL_ALL_MAX)
  long chol_all_max = 10000;
  long chol_all_min = 300;   // (WV_CHOL_ALL_MIN) below: the fused launch per column
  int solo = 0;         // the batch being evaluated declared itself alone on the device
  int trtri_all = 1;    // solo batches: the triangular inverse as one persistent launch (wv_trtri_all_kernel; WV_TRTRI_ALL)
  long trtri_all_max = 5000;       // ... while n_active * nt <= this (WV_TRTRI_ALL_MAX; measured crossover ~600 models at nt = 10)
  int trtri_ctas = 592; // its persistent CTAs (4 per SM; WV_TRTRI_CTAS)
  int few_models = 1;   // few models in flight: one launch each for the Cholesky and the triangular inverse (WV_FEW_MODELS=0: off)
  int chol_lag = 640;   // work items between a panel tile (j + 1, j) and the diagonal block j + 1 that needs it (WV_CHOL_LAG)
  int epoch = 0;   // evaluation counter of the engine: the value the diagonal CTAs publish in step_flag
};
struct WvProfiler {
  bool enabled = false;
  cudaEvent_t ev[256];
  int cls[256];
  int n_ev = 0, n_alloc = 0;
  double ms[WV_K_NCLASS] = {0};
  long long launches[WV_K_NCLASS] = {0};
  // mark the end of a kernel of class c (c < 0: start marker)
  void mark(int c, cudaStream_t st) {
    if (!enabled) return;
    if (n_ev == 256) resolve();
    if (n_ev >= n_alloc) { cudaEventCreate(&ev[n_alloc]); ++n_alloc; }
    cudaEventRecord(ev[n_ev], st);
    cls[n_ev++] = c;
  }
  void resolve() {   // requires the stream to be idle or blocks until the last event completes
    if (n_ev == 0) return;
    cudaEventSynchronize(ev[n_ev - 1]);
    for (int i = 1; i < n_ev; ++i) {
      if (cls[i] < 0) continue;
      float t = 0.f;
      cudaEventElapsedTime(&t, ev[i - 1], ev[i]);
      ms[cls[i]] += t;
      launches[cls[i]] += 1;
    }
    n_ev = 0;
  }
  void destroy() { for (int i = 0; i < n_alloc; ++i) cudaEventDestroy(ev[i]); n_alloc = 0; n_ev = 0; }
};
#endif  // __CUDACC_RTC__

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void wv_cp_async16(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void wv_cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void wv_cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void wv_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// lower-triangular tile index -> (ti, tj), tj <= ti
__device__ __forceinline__ void wv_tile_from_linear(int t, int& ti, int& tj) {
  int i = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((i + 1) * (i + 2) / 2 <= t) ++i;
  while (i * (i + 1) / 2 > t) --i;
  ti = i;
  tj = t - i * (i + 1) / 2;
}

// ---------------------------------------------------------------------------------------------
// helpers of the kernel-tree leaves (the leaves themselves live in wv_elem.cuh)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double wv_powi(double b, int d) {
  double r = 1.0;
  for (int i = 0; i < d; ++i) r *= b;
  return r;
}

// "cheap" leaves have no transcendental: they are evaluated first inside a product so that an all-zero
// categorical mask can skip the expensive factors (the encoder orders the leaves of a component this way).
__device__ __forceinline__ bool wv_leaf_is_cheap(int type) {
  return type == WV_LEAF_CAT || type == WV_LEAF_CONST || type == WV_LEAF_LINEAR || type == WV_LEAF_EMPTY;
}

// Role of a warp inside the 2 x 2 warp grid of a 64x64 tile product.  Hardware places warp w of every CTA on SM
// sub-partition w mod 4, and the roles are not equally heavy: the padding ("dead") half of a model's last tile row and
// the zero half of a triangular operand always fall on the same roles.  Flipping the row role with the CTA's parity
// spreads the light roles over all four tensor pipes of the SM.
__device__ __forceinline__ int wv_warp_role() {
  return (int)(threadIdx.x >> 5) ^ ((int)((blockIdx.x + blockIdx.y) & 1u) << 1);
}

// ---------------------------------------------------------------------------------------------
// 64x64 NT tile GEMM on the FP64 tensor pipe:
//   acc[m][n] = sum_{k in [k0,k1)} Ag[m*ld + k] * Bg[n*ld + k],  k0 % 16 == 0, k1 % 8 == 0
// 128 threads = 4 warps (2x2), each warp owns a 32x32 sub-tile = 4x4 DMMA.8x8x4 accumulators.
// Operands are staged with a 3-deep cp.async ring (16B chunks, zero-fill beyond k1).
// ---------------------------------------------------------------------------------------------
struct WvGemmSmem {
  double a[WV_STAGES][WV_NB * WV_LDS];
  double b[WV_STAGES][WV_NB * WV_LDS];
};

#define WV_LDN 68           // smem row stride of an NN B stage: [k][n], 16 k-rows x 64 columns (68 mod 16 == 4)

// BNN = false: B operand rows are n, k contiguous (Bg[n * ld + k]);  BNN = true: B operand rows are k, n contiguous
// (Bg[k * ld + n]) -- the stage is then kept [k][n] with stride WV_LDN and the DMMA B fragments read it transposed.
template <bool BNN>
__device__ __forceinline__ void wv_gemm_issue(WvGemmSmem& sm, int stage, const double* __restrict__ Ag,
                                              const double* __restrict__ Bg, int ld, int kc, int k1) {
  const int t = threadIdx.x;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int q = t + WV_GEMM_THREADS * r;
    int row = q >> 3, ch = q & 7;
    int k = kc + ch * 2;
    bool ok = k < k1;
    int ks = ok ? k : kc;
    wv_cp_async16(&sm.a[stage][row * WV_LDS + ch * 2], Ag + (size_t)row * ld + ks, ok);
    if (!BNN) {
      wv_cp_async16(&sm.b[stage][row * WV_LDS + ch * 2], Bg + (size_t)row * ld + ks, ok);
    } else {
      int kr = q >> 5, cn = (q & 31) * 2;
      bool okb = kc + kr < k1;
      wv_cp_async16(&sm.b[stage][kr * WV_LDN + cn], Bg + (size_t)(okb ? kc + kr : kc) * ld + cn, okb);
    }
  }
}

// `dead`: this warp's 32x32 part of the tile is padding (rows / columns beyond the model's n + 1): it keeps staging
// operands and meeting the barriers but issues no DMMA, which frees the tensor pipe for the other warps of the SM
// (the last tile row / column of a model holds (n + 1) mod 64 real rows -- 25 of 64 at n = 600).
// One 16-deep chunk of a product whose A operand is the UPPER TRIANGULAR diagonal tile of Mt = L^{-T} (row m is zero for
// k < m): the 8-row group mi of row role WM needs the 4-deep step at k = 16 C + kk only if 16 C + kk + 4 > 32 WM + 8 mi.
// C and WM are template constants, so the skipped DMMAs cost nothing -- 44 % of the tile product.
template <int C, int WM>
__device__ __forceinline__ void wv_chunk_tri_a(const double* __restrict__ as, const double* __restrict__ bs,
                                               double (&acc)[4][4][2]) {
#pragma unroll
  for (int kk = 0; kk < WV_BK; kk += 4) {
    if (16 * C + kk + 4 <= 32 * WM) continue;
    double af[4], bf[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      af[i] = (16 * C + kk + 4 > 32 * WM + 8 * i) ? as[i * 8 * WV_LDS + kk] : 0.0;
      bf[i] = bs[i * 8 * WV_LDS + kk];
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
      if (16 * C + kk + 4 > 32 * WM + 8 * mi) {
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) wv_dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
      }
  }
}

// TRIA: the first 64 k-columns of the A operand (k0 is a tile boundary) are an upper triangular tile
template <bool BNN, bool TRIA = false>
__device__ __forceinline__ void wv_gemm_64(WvGemmSmem& sm, const double* __restrict__ Ag,
                                           const double* __restrict__ Bg, int ld, int k0, int k1,
                                           double (&acc)[4][4][2], bool dead = false) {
  const int lane = threadIdx.x & 31, warp = wv_warp_role();
  const int wm = warp >> 1, wn = warp & 1;
  const int fr = lane >> 2, fk = lane & 3;
  const int nchunks = (k1 - k0 + WV_BK - 1) / WV_BK;
#pragma unroll
  for (int s = 0; s < WV_STAGES - 1; ++s) {
    if (s < nchunks) wv_gemm_issue<BNN>(sm, s, Ag, Bg, ld, k0 + s * WV_BK, k1);
    wv_cp_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    wv_cp_wait<WV_STAGES - 2>();
    __syncthreads();
    {
      int cn = c + WV_STAGES - 1;
      if (cn < nchunks) wv_gemm_issue<BNN>(sm, cn % WV_STAGES, Ag, Bg, ld, k0 + cn * WV_BK, k1);
      wv_cp_commit();
    }
    if (dead) continue;
    const double* as = sm.a[c % WV_STAGES] + (wm * 32 + fr) * WV_LDS + fk;
    const double* bs = BNN ? sm.b[c % WV_STAGES] + fk * WV_LDN + wn * 32 + fr
                           : sm.b[c % WV_STAGES] + (wn * 32 + fr) * WV_LDS + fk;
    if (TRIA && !BNN && c < 4) {
      if (wm == 0) {
        if (c == 0) wv_chunk_tri_a<0, 0>(as, bs, acc);
        else if (c == 1) wv_chunk_tri_a<1, 0>(as, bs, acc);
        else if (c == 2) wv_chunk_tri_a<2, 0>(as, bs, acc);
        else wv_chunk_tri_a<3, 0>(as, bs, acc);
      } else {
        if (c == 2) wv_chunk_tri_a<2, 1>(as, bs, acc);
        else if (c == 3) wv_chunk_tri_a<3, 1>(as, bs, acc);       // chunks 0, 1: rows 32.. are zero for k < 32
      }
      continue;
    }
#pragma unroll
    for (int kk = 0; kk < WV_BK; kk += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        af[i] = as[i * 8 * WV_LDS + kk];
        bf[i] = BNN ? bs[kk * WV_LDN + i * 8] : bs[i * 8 * WV_LDS + kk];
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) wv_dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
    }
  }
  wv_cp_wait<0>();
  __syncthreads();
}
__device__ __forceinline__ void wv_gemm_nt_64(WvGemmSmem& sm, const double* __restrict__ Ag,
                                              const double* __restrict__ Bg, int ld, int k0, int k1,
                                              double (&acc)[4][4][2], bool dead = false) {
  wv_gemm_64<false>(sm, Ag, Bg, ld, k0, k1, acc, dead);
}
// A's first k-tile is the upper triangular diagonal tile of Mt (trtri, kinv)
__device__ __forceinline__ void wv_gemm_nt_64_tria(WvGemmSmem& sm, const double* __restrict__ Ag,
                                                   const double* __restrict__ Bg, int ld, int k0, int k1,
                                                   double (&acc)[4][4][2], bool dead = false) {
  wv_gemm_64<false, true>(sm, Ag, Bg, ld, k0, k1, acc, dead);
}

// Packed lower-triangular 64x64 block (the inverse of a Cholesky diagonal block) in shared memory: the 8 rows of row
// block nb keep their first 8 (nb + 1) columns with row stride 8 (nb + 1) + 4 (== 4 or 12 mod 16: the DMMA fragment
// loads stay conflict free), 2560 doubles instead of 64 x 68 -- with it the panel / trtri CTAs need 55 KB and four fit
// on an SM.
#define WV_DP_DOUBLES 2560
__device__ __forceinline__ int wv_dp_offset(int nb) { return 32 * nb * (nb + 2); }
__device__ __forceinline__ int wv_dp_stride(int nb) { return 8 * (nb + 1) + 4; }

// gmem row-major 64x64 (zeros above the diagonal) -> packed smem; L2 loads (in the fused Cholesky step the block was
// written by another SM of the same launch)
__device__ __forceinline__ void wv_dp_load(double* __restrict__ Dp, const double* __restrict__ Dg_) {
  const double2* Dg = reinterpret_cast<const double2*>(Dg_);
  for (int i = threadIdx.x; i < WV_NB * WV_NB / 2; i += WV_GEMM_THREADS) {
    const int rr = i >> 5, c2 = (i & 31) * 2, nb = rr >> 3;
    if (c2 < 8 * (nb + 1))
      *reinterpret_cast<double2*>(&Dp[wv_dp_offset(nb) + (rr & 7) * wv_dp_stride(nb) + c2]) = __ldcg(Dg + i);
  }
}

// second-stage product from shared memory operands (64x64x64): acc[m][n] = sum_k Ts[m][k] * B[n][k], B LOWER TRIANGULAR
// and packed (wv_dp_load)
__device__ __forceinline__ void wv_gemm_nt_smem64(const double* __restrict__ Ts, const double* __restrict__ Dp,
                                                  double (&acc)[4][4][2], bool dead = false) {
  if (dead) return;
  const int lane = threadIdx.x & 31, warp = wv_warp_role();
  const int wm = warp >> 1, wn = warp & 1;
  const int fr = lane >> 2, fk = lane & 3;
  const double* as = Ts + (wm * 32 + fr) * WV_LDT + fk;
  const double* bs[4];
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) bs[ni] = Dp + wv_dp_offset(wn * 4 + ni) + fr * wv_dp_stride(wn * 4 + ni) + fk;
  // column block nb = wn * 4 + ni of the result only needs k < 8 (nb + 1) -- 56 % of the tile products, and the warps
  // of the left half finish after k = 32
  const int kend = (wn * 4 + 4) * 8;
#pragma unroll 4
  for (int kk = 0; kk < kend; kk += 4) {
    double af[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) af[i] = as[i * 8 * WV_LDT + kk];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
      if (kk < (wn * 4 + ni + 1) * 8) {
        const double bf = bs[ni][kk];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) wv_dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf);
      }
  }
}

// accumulator fragment coordinates of this thread: row = r0 + mi*8, col = c0 + ni*8 (+0,+1)
__device__ __forceinline__ void wv_frag_origin(int& r0, int& c0) {
  const int lane = threadIdx.x & 31, warp = wv_warp_role();
  r0 = (warp >> 1) * 32 + (lane >> 2);
  c0 = (warp & 1) * 32 + (lane & 3) * 2;
}

__device__ __forceinline__ void wv_zero_acc(double (&acc)[4][4][2]) {
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
}
