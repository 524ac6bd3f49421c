// waveome_b200 — kernels of one batched LML+gradient evaluation.  See wv_kernels.cuh for the plan.
#include "wv_kernels.cuh"

// dynamic shared memory is carved by hand; the GEMM pipeline and the epilogue tiles alias each other.
extern __shared__ __align__(16) unsigned char wv_smem_raw[];

// =============================================================================================
// gram: lower tiles of K + sigma^2 I, RHS row, identity padding.
// grid (n_lower_tiles, n_active), 256 threads, each thread a 4x4 micro-tile.
// Algorithmic traffic: 8 n^2 bytes written (lower half + diagonal tiles actually written: ~4 n^2).
// =============================================================================================
struct WvElemSmem {
  WvProgram pg;
  double theta[WV_MAX_SLOTS];
  double xr[WV_MAX_DIMS][WV_NB];
  double xc[WV_MAX_DIMS][WV_NB];
  double red[WV_MAX_SLOTS][8];   // per-warp partial sums (grad only)
};

__device__ __forceinline__ void wv_elem_prologue(const WvBatchDev& bd, int b, int ti, int tj, const double* xall,
                                                 WvElemSmem& sm) {
  const WvProgram* gp = bd.programs + bd.prog_id[b];
  const int nwords = sizeof(WvProgram) / 4;
  const int32_t* src = reinterpret_cast<const int32_t*>(gp);
  int32_t* dst = reinterpret_cast<int32_t*>(&sm.pg);
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  wv_load_theta(&sm.pg, xall + (size_t)b * bd.P, sm.theta);
  for (int i = threadIdx.x; i < sm.pg.n_dims * WV_NB; i += blockDim.x) {
    int d = i / WV_NB, r = i % WV_NB;
    const double* col = bd.Xt + (size_t)sm.pg.dims[d] * bd.npad;
    sm.xr[d][r] = col[ti * WV_NB + r];
    sm.xc[d][r] = col[tj * WV_NB + r];
  }
  __syncthreads();
}

// thread -> micro-tile mapping: warp w owns the compact 16x32 region rows (w>>1)*16.., cols (w&1)*32..; lane l the
// 4x4 micro-tile at (+ (l>>3)*4, + (l&7)*4).  Compact warp regions make "the categorical mask is zero for the whole
// warp" common once the rows are sorted by their categorical columns (wv_batch_create does that).
__device__ __forceinline__ void wv_elem_coords(int& r_off, int& c_off, bool& above_diag, bool diag_tile) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = (warp >> 1) * 16, wc = (warp & 1) * 32;
  r_off = wr + (lane >> 3) * 4;
  c_off = wc + (lane & 7) * 4;
  above_diag = diag_tile && wc > wr + 15;     // every element of the warp's region has col > row
}

__global__ void __launch_bounds__(WV_ELEM_THREADS) wv_gram_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                                  const double* __restrict__ xall) {
  WvElemSmem& sm = *reinterpret_cast<WvElemSmem*>(wv_smem_raw);
  const int b = active[blockIdx.y];
  int ti, tj;
  wv_tile_from_linear(blockIdx.x, ti, tj);
  wv_elem_prologue(bd, b, ti, tj, xall, sm);
  int r_off, c_off;
  bool above;
  wv_elem_coords(r_off, c_off, above, ti == tj);
  if (above) return;                            // the strict upper part of a diagonal tile is never read
  const int n = bd.n;
  double acc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.0;
  for (int c = 0; c < sm.pg.n_comp; ++c) {
    double prod[16];
    const int l0 = sm.pg.comp_start[c], l1 = sm.pg.comp_start[c + 1];
    bool skip = false;
    for (int l = l0; l < l1; ++l) {
      const WvLeaf lf = sm.pg.leaves[l];
      if (l > l0 && !wv_leaf_is_cheap(lf.type) && wv_warp_all_zero(prod)) { skip = true; break; }
      double xi[4], xj[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { xi[a] = sm.xr[lf.dim][r_off + a]; xj[a] = sm.xc[lf.dim][c_off + a]; }
      if (l == l0) wv_leaf_mul<true>(lf, sm.theta, xi, xj, prod);
      else wv_leaf_mul<false>(lf, sm.theta, xi, xj, prod);
    }
    if (!skip) {
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] += prod[e];
    }
  }
  const double s2 = sm.theta[sm.pg.noise_slot];
  const double cmean = sm.pg.mean_slot >= 0 ? sm.theta[sm.pg.mean_slot] : 0.0;
  double* Ab = bd.A + (size_t)b * bd.npad * bd.npad;
  const double* yb = bd.Y + (size_t)b * bd.npad;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int gi = ti * WV_NB + r_off + a;
    double out[4];
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int gj = tj * WV_NB + c_off + bb;
      double v;
      if (gi < n && gj < n) v = acc[a * 4 + bb] + (gi == gj ? s2 : 0.0);
      else if (gi == n && gj < n) v = yb[gj] - cmean;     // RHS row d^T
      else v = (gi == gj) ? 1.0 : 0.0;                     // identity padding (incl. A[n][n] = 1)
      out[bb] = v;
    }
    double2* dst = reinterpret_cast<double2*>(Ab + (size_t)gi * bd.npad + tj * WV_NB + c_off);
    dst[0] = make_double2(out[0], out[1]);
    dst[1] = make_double2(out[2], out[3]);
  }
}

// =============================================================================================
// chol_diag(j): T = A[j,j] - sum_{k<j} L[j,k] L[j,k]^T ; L_jj = chol(T) ; Linv_jj = L_jj^{-1}
// grid (n_active), 128 threads.  Row n (the RHS row) takes part as an ordinary row but is never a pivot.
//
// The 64x64 block is handled as 2x2 blocks of 32: each 32x32 Cholesky and triangular inverse runs in the registers
// of ONE warp (lane r owns row r of L, lane j owns column j of L^{-1}; pivots and multipliers travel by shuffles,
// everything fully unrolled so all register indices are static), the coupling products are 32^3 DMMA GEMMs from
// shared memory:   L21 = T21 X11^T,  T22 -= L21 L21^T,  X21 = -X22 (L21 X11).
// =============================================================================================
struct WvDiagSmem {
  union {
    WvGemmSmem g;
    struct {
      double T[WV_NB * WV_LDT];   // T -> L (lower);  [0:32, 32:64] is scratch for X11^T
      double X[WV_NB * WV_LDT];   // L^{-1} (lower);  [0:32, 32:64] is scratch for (L21 X11)^T
    } e;
  };
  double invd[WV_NB];
  double logsum[2];
  int fail;
};

// 32x32 Cholesky in registers: lane r holds row r in a[0..31] (lower part meaningful).  `rhs` (warp-uniform, -1 if
// none) is the local index of the augmented RHS row: unit diagonal, never a pivot.  The dependent chain per column is
// shuffle -> rsqrt -> multiply -> shuffle -> fma; logs are taken afterwards, one pivot per lane, in parallel.
// Returns sum of log(diag) over columns c < nreal; sets fail if a pivot is <= 0 (NaN pivots flow through, as in
// Eigen's LLT).  myinv = 1 / L[lane][lane].
__device__ __forceinline__ double wv_potrf32(double (&a)[32], double& myinv, int rhs, int nreal, bool& fail) {
  const int lane = threadIdx.x & 31;
  double mypiv = 1.0;
  myinv = 1.0;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    double d = __shfl_sync(0xffffffffu, a[c], c);
    if (c == rhs) d = 1.0;
    if (d <= 0.0) fail = true;
    const double inv = rsqrt(d);
    double l = a[c] * inv;                 // lane c: d * rsqrt(d) = sqrt(d)
    if (lane == c) { mypiv = d; myinv = inv; }
    if (lane < c) l = 0.0;
    a[c] = l;
#pragma unroll
    for (int c2 = c + 1; c2 < 32; ++c2) {
      const double v = __shfl_sync(0xffffffffu, l, c2);
      a[c2] = fma(-l, v, a[c2]);
    }
  }
  double lg = lane < nreal ? 0.5 * log(mypiv) : 0.0;
  for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
  return lg;
}

// 32x32 lower-triangular inverse: L (row stride ldl) and 1/diag(L) are read from shared memory with warp-uniform
// (broadcast) loads; lane j produces column j of X = L^{-1} in x[0..31] (x[r] = X[r][j], zero for r < j).
// Four partial accumulators keep the dependent FMA chain at r/4.
__device__ __forceinline__ void wv_trtri32(const double* __restrict__ Ls, int ldl, const double* __restrict__ invd,
                                           double (&x)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    double s0 = (lane == r) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int k = 0; k < r; ++k) {
      const double lv = Ls[r * ldl + k];
      if ((k & 3) == 0) s0 = fma(-lv, x[k], s0);
      else if ((k & 3) == 1) s1 = fma(-lv, x[k], s1);
      else if ((k & 3) == 2) s2 = fma(-lv, x[k], s2);
      else s3 = fma(-lv, x[k], s3);
    }
    x[r] = ((s0 + s1) + (s2 + s3)) * invd[r];
  }
}

// C[32x32] = sum_k A[m][k] B[n][k], k < 32, operands in shared memory (row strides lda, ldb); 4 warps, 16x16 each.
__device__ __forceinline__ void wv_gemm32_nt(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                             double (&acc)[2][2][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const double* as = A + ((warp >> 1) * 16 + fr) * lda + fk;
  const double* bs = B + ((warp & 1) * 16 + fr) * ldb + fk;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll
  for (int kk = 0; kk < 32; kk += 4) {
    double af[2], bf[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) { af[i] = as[i * 8 * lda + kk]; bf[i] = bs[i * 8 * ldb + kk]; }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) wv_dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
  }
}
// row/col of this thread's first accumulator element inside the 32x32 result (+ mi*8 rows, + ni*8 cols, +0/+1 col)
__device__ __forceinline__ void wv_frag32_origin(int& r0, int& c0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  r0 = (warp >> 1) * 16 + (lane >> 2);
  c0 = (warp & 1) * 16 + (lane & 3) * 2;
}

__global__ void __launch_bounds__(WV_GEMM_THREADS, 3) wv_chol_diag_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                                       int j) {
  WvDiagSmem& sm = *reinterpret_cast<WvDiagSmem*>(wv_smem_raw);
  const int b = active[blockIdx.x];
  const int ld = bd.npad;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* Ab = bd.A + (size_t)b * ld * ld;
  const double* Lrow = Ab + (size_t)j * WV_NB * ld;
  {
    double acc[4][4][2];
    wv_zero_acc(acc);
    if (threadIdx.x == 0) { sm.fail = 0; sm.logsum[0] = sm.logsum[1] = 0.0; }
    if (j > 0) wv_gemm_nt_64(sm.g, Lrow, Lrow, ld, 0, j * WV_NB, acc);
    else __syncthreads();
    int r0, c0;
    wv_frag_origin(r0, c0);
    const double* Tg = Ab + (size_t)j * WV_NB * ld + j * WV_NB;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        int r = r0 + mi * 8, c = c0 + ni * 8;
        double2 a = *reinterpret_cast<const double2*>(Tg + (size_t)r * ld + c);
        *reinterpret_cast<double2*>(&sm.e.T[r * WV_LDT + c]) = make_double2(a.x - acc[mi][ni][0], a.y - acc[mi][ni][1]);
      }
  }
  for (int i = threadIdx.x; i < WV_NB * WV_LDT; i += WV_GEMM_THREADS) sm.e.X[i] = 0.0;
  __syncthreads();
  double* T = sm.e.T;
  double* X = sm.e.X;
  const int rhs = bd.n - j * WV_NB;                       // local index of the RHS row (may be outside [0,64))
  const int nreal = min(WV_NB, bd.n - j * WV_NB);         // pivots that belong to K (log-det terms)

  // ---- block (1,1): warp 0
  if (warp == 0) {
    bool fail = false;
    {
      double a[32], myinv;
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = T[lane * WV_LDT + c];
      const double ls = wv_potrf32(a, myinv, rhs, nreal, fail);
#pragma unroll
      for (int c = 0; c < 32; ++c) T[lane * WV_LDT + c] = a[c];      // L11 (zeros above the diagonal)
      sm.invd[lane] = myinv;
      if (lane == 0) { sm.logsum[0] = ls; if (fail) sm.fail = 1; }
    }
    __syncwarp();
    double x[32];
    wv_trtri32(T, WV_LDT, sm.invd, x);
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      X[c * WV_LDT + lane] = x[c];                  // X11[r=c][j=lane]
      T[lane * WV_LDT + 32 + c] = x[c];             // scratch: X11^T[j=lane][r=c]
    }
  }
  __syncthreads();
  int r0, c0;
  wv_frag32_origin(r0, c0);
  double acc[2][2][2];
  // ---- L21 = T21 X11^T   (in place over T21)
  wv_gemm32_nt(T + 32 * WV_LDT, WV_LDT, X, WV_LDT, acc);
  __syncthreads();
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni)
      *reinterpret_cast<double2*>(&T[(32 + r0 + mi * 8) * WV_LDT + c0 + ni * 8]) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
  __syncthreads();
  // ---- T22 -= L21 L21^T
  wv_gemm32_nt(T + 32 * WV_LDT, WV_LDT, T + 32 * WV_LDT, WV_LDT, acc);
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) {
      double2* p = reinterpret_cast<double2*>(&T[(32 + r0 + mi * 8) * WV_LDT + 32 + c0 + ni * 8]);
      double2 v = *p;
      *p = make_double2(v.x - acc[mi][ni][0], v.y - acc[mi][ni][1]);
    }
  __syncthreads();
  // ---- block (2,2): warp 0
  if (warp == 0) {
    bool fail = false;
    {
      double a[32], myinv;
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = T[(32 + lane) * WV_LDT + 32 + c];
      const double ls = wv_potrf32(a, myinv, rhs - 32, nreal - 32, fail);
#pragma unroll
      for (int c = 0; c < 32; ++c) T[(32 + lane) * WV_LDT + 32 + c] = a[c];      // L22
      sm.invd[32 + lane] = myinv;
      if (lane == 0) { sm.logsum[1] = ls; if (fail) sm.fail = 1; }
    }
    __syncwarp();
    double x[32];
    wv_trtri32(T + 32 * WV_LDT + 32, WV_LDT, sm.invd + 32, x);
#pragma unroll
    for (int c = 0; c < 32; ++c) X[(32 + c) * WV_LDT + 32 + lane] = x[c];      // X22[r=c][j=lane]
  }
  __syncthreads();
  // ---- P^T = X11^T-rows x L21-rows:  Pt[n][m] = sum_k X11[k][n] L21[m][k]   -> scratch X[0:32, 32:64]
  wv_gemm32_nt(T + 32, WV_LDT, T + 32 * WV_LDT, WV_LDT, acc);
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni)
      *reinterpret_cast<double2*>(&X[(r0 + mi * 8) * WV_LDT + 32 + c0 + ni * 8]) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
  __syncthreads();
  // ---- X21 = -X22 P :  X21[m][n] = -sum_k X22[m][k] Pt[n][k]
  wv_gemm32_nt(X + 32 * WV_LDT + 32, WV_LDT, X + 32, WV_LDT, acc);
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni)
      *reinterpret_cast<double2*>(&X[(32 + r0 + mi * 8) * WV_LDT + c0 + ni * 8]) = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
  __syncthreads();

  // ---- write L_jj (lower, zeros above), Linv_jj (row-major) and Linv_jj^T into Mt[j,j]
  double* Tg = Ab + (size_t)j * WV_NB * ld + j * WV_NB;
  double* Mg = bd.Mt + (size_t)b * ld * ld + (size_t)j * WV_NB * ld + j * WV_NB;
  double* Dg = bd.Dinv + ((size_t)b * bd.nt + j) * WV_NB * WV_NB;
  for (int i = threadIdx.x; i < WV_NB * WV_NB; i += WV_GEMM_THREADS) {
    int rr = i >> 6, cc = i & 63;
    Tg[(size_t)rr * ld + cc] = cc <= rr ? T[rr * WV_LDT + cc] : 0.0;
    Dg[i] = cc <= rr ? X[rr * WV_LDT + cc] : 0.0;
    Mg[(size_t)rr * ld + cc] = rr <= cc ? X[cc * WV_LDT + rr] : 0.0;
  }
  if (threadIdx.x == 0) {
    bd.logdet_part[(size_t)b * bd.nt + j] = sm.logsum[0] + sm.logsum[1];
    if (sm.fail) bd.chol_fail[b] = 1;
  }
}

// =============================================================================================
// generic tile step:  out = sign * (C_in - sum_{k in [k0,k1)} Arow[.,k] Brow[.,k]^T) * Dinv^T
//   chol_panel(j): tile (i,j), i>j   : L[i,j]  = (A[i,j] - sum_{k<j} L[i,k] L[j,k]^T) Linv_jj^T
//   trtri(i)     : tile (j,i), j<i   : Mt[j,i] = -( sum_{k=j..i-1} Mt[j,k] L[i,k]^T ) Linv_ii^T
// grid (n_tiles_in_step, n_active), 128 threads.
// =============================================================================================
struct WvPanelSmem {
  union {
    WvGemmSmem g;
    struct {
      double T[WV_NB * WV_LDT];
      double D[WV_NB * WV_LDT];
    } e;
  };
};

template <int MODE>   // 0 = chol_panel, 1 = trtri
__global__ void __launch_bounds__(WV_GEMM_THREADS) wv_panel_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                                   int step) {
  WvPanelSmem& sm = *reinterpret_cast<WvPanelSmem*>(wv_smem_raw);
  const int b = active[blockIdx.y];
  const int ld = bd.npad;
  double* Ab = bd.A + (size_t)b * ld * ld;
  double* Mb = bd.Mt + (size_t)b * ld * ld;
  const double *Ag, *Bg;
  double* Out;
  const double* Cin;
  int k0, k1;
  if (MODE == 0) {
    const int i = step + 1 + blockIdx.x, j = step;
    Ag = Ab + (size_t)i * WV_NB * ld;
    Bg = Ab + (size_t)j * WV_NB * ld;
    k0 = 0; k1 = j * WV_NB;
    Out = Ab + (size_t)i * WV_NB * ld + j * WV_NB;
    Cin = Out;
  } else {
    const int i = step, j = blockIdx.x;
    Ag = Mb + (size_t)j * WV_NB * ld;
    Bg = Ab + (size_t)i * WV_NB * ld;
    k0 = j * WV_NB; k1 = i * WV_NB;
    Out = Mb + (size_t)j * WV_NB * ld + i * WV_NB;
    Cin = nullptr;
  }
  double acc[4][4][2];
  wv_zero_acc(acc);
  if (k1 > k0) wv_gemm_nt_64(sm.g, Ag, Bg, ld, k0, k1, acc);
  int r0, c0;
  wv_frag_origin(r0, c0);
  // T = C_in - acc  (or -acc ... the sign is applied at the end for trtri) -> smem as the A operand of the 2nd product
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      int r = r0 + mi * 8, c = c0 + ni * 8;
      double2 v;
      if (MODE == 0) {
        double2 a = *reinterpret_cast<const double2*>(Cin + (size_t)r * ld + c);
        v = make_double2(a.x - acc[mi][ni][0], a.y - acc[mi][ni][1]);
      } else {
        v = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
      }
      *reinterpret_cast<double2*>(&sm.e.T[r * WV_LDT + c]) = v;
    }
  // D = Linv of the step's diagonal block (row-major)
  const double2* Dg = reinterpret_cast<const double2*>(bd.Dinv + ((size_t)b * bd.nt + step) * WV_NB * WV_NB);
  for (int i = threadIdx.x; i < WV_NB * WV_NB / 2; i += WV_GEMM_THREADS) {
    int rr = i >> 5, c2 = (i & 31) * 2;
    *reinterpret_cast<double2*>(&sm.e.D[rr * WV_LDT + c2]) = Dg[i];
  }
  __syncthreads();
  wv_zero_acc(acc);
  wv_gemm_nt_smem64(sm.e.T, sm.e.D, acc);
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      int r = r0 + mi * 8, c = c0 + ni * 8;
      *reinterpret_cast<double2*>(Out + (size_t)r * ld + c) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
    }
}

// =============================================================================================
// extract: alpha_j = -Mt[j][n], quad = |L[n][0:n]|^2 = |L^{-1} d|^2 ; then clear column n of Mt so that
// kinv = Mt Mt^T excludes the augmented row.  grid (n_active), 256 threads.
// =============================================================================================
__global__ void __launch_bounds__(256) wv_extract_kernel(WvBatchDev bd, const int* __restrict__ active) {
  __shared__ double red[8];
  const int b = active[blockIdx.x];
  const int ld = bd.npad, n = bd.n;
  double* Mb = bd.Mt + (size_t)b * ld * ld;
  const double* zrow = bd.A + (size_t)b * ld * ld + (size_t)n * ld;
  double q = 0.0;
  for (int jx = threadIdx.x; jx < ld; jx += blockDim.x) {
    double a = 0.0;
    if (jx < n) {
      a = -Mb[(size_t)jx * ld + n];
      double z = zrow[jx];
      q += z * z;
    }
    bd.alpha[(size_t)b * ld + jx] = a;
    if (jx <= n) Mb[(size_t)jx * ld + n] = 0.0;
  }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    bd.quad[b] = s;
  }
}

// =============================================================================================
// kinv: A[i,j] (j<=i) = sum_{k >= i*64}^{n8} Mt[i][k] Mt[j][k]^T.   grid (n_lower_tiles, n_active), 128 threads.
// =============================================================================================
__global__ void __launch_bounds__(WV_GEMM_THREADS) wv_kinv_kernel(WvBatchDev bd, const int* __restrict__ active) {
  WvGemmSmem& sm = *reinterpret_cast<WvGemmSmem*>(wv_smem_raw);
  const int b = active[blockIdx.y];
  const int ld = bd.npad;
  int ti, tj;
  wv_tile_from_linear(blockIdx.x, ti, tj);
  const double* Mb = bd.Mt + (size_t)b * ld * ld;
  double acc[4][4][2];
  wv_zero_acc(acc);
  wv_gemm_nt_64(sm, Mb + (size_t)ti * WV_NB * ld, Mb + (size_t)tj * WV_NB * ld, ld, ti * WV_NB, bd.n8, acc);
  int r0, c0;
  wv_frag_origin(r0, c0);
  double* Out = bd.A + (size_t)b * ld * ld + (size_t)ti * WV_NB * ld + tj * WV_NB;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      int r = r0 + mi * 8, c = c0 + ni * 8;
      *reinterpret_cast<double2*>(Out + (size_t)r * ld + c) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
    }
}

// =============================================================================================
// grad: partial[b][tile][slot] = sum over the tile of wgt_ij * W_ij * dK_ij/dtheta_slot,
//   W = alpha alpha^T - K^{-1};  wgt = 2 below the diagonal, 1 on it, 0 above / outside [0,n).
// dK/dtheta is regenerated from the kernel program; nothing of size n^2 is materialised.
// grid (n_lower_tiles, n_active), 256 threads.  Algorithmic traffic: 8 n^2 bytes read.
// =============================================================================================
__global__ void __launch_bounds__(WV_ELEM_THREADS) wv_grad_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                                  const double* __restrict__ xall) {
  WvElemSmem& sm = *reinterpret_cast<WvElemSmem*>(wv_smem_raw);
  const int b = active[blockIdx.y];
  int ti, tj;
  wv_tile_from_linear(blockIdx.x, ti, tj);
  wv_elem_prologue(bd, b, ti, tj, xall, sm);
  int r_off, c_off;
  bool above;
  wv_elem_coords(r_off, c_off, above, ti == tj);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = bd.n, ld = bd.npad;
  for (int i = threadIdx.x; i < WV_MAX_SLOTS * 8; i += blockDim.x) (&sm.red[0][0])[i] = 0.0;
  __syncthreads();
  if (!above) {
    const double* Kb = bd.A + (size_t)b * ld * ld;
    const double* al = bd.alpha + (size_t)b * ld;
    double w[16];
    double trw = 0.0;
    double aj[4];
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) aj[bb] = al[tj * WV_NB + c_off + bb];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int gi = ti * WV_NB + r_off + a;
      const double2* src = reinterpret_cast<const double2*>(Kb + (size_t)gi * ld + tj * WV_NB + c_off);
      double2 k01 = src[0], k23 = src[1];
      double kin[4] = {k01.x, k01.y, k23.x, k23.y};
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
    for (int c = 0; c < sm.pg.n_comp; ++c) {
      const int l0 = sm.pg.comp_start[c], l1 = sm.pg.comp_start[c + 1];
      for (int l = l0; l < l1; ++l) {
        const WvLeaf lf = sm.pg.leaves[l];
        const bool tv = lf.s_var >= 0 && sm.pg.slots[lf.s_var].xindex >= 0;
        const bool tl = lf.s_ls >= 0 && sm.pg.slots[lf.s_ls].xindex >= 0;
        const bool ta = lf.s_aux >= 0 && sm.pg.slots[lf.s_aux].xindex >= 0;
        if (!(tv || tl || ta)) continue;
        double wo[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) wo[e] = w[e];
        bool skip = false;
        for (int l2 = l0; l2 < l1; ++l2) {
          if (l2 == l) continue;
          const WvLeaf lo = sm.pg.leaves[l2];
          if (!wv_leaf_is_cheap(lo.type) && wv_warp_all_zero(wo)) { skip = true; break; }
          double xi[4], xj[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) { xi[a] = sm.xr[lo.dim][r_off + a]; xj[a] = sm.xc[lo.dim][c_off + a]; }
          wv_leaf_mul<false>(lo, sm.theta, xi, xj, wo);
        }
        if (skip || (!wv_leaf_is_cheap(lf.type) && wv_warp_all_zero(wo))) continue;   // every contribution is zero
        double xi[4], xj[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) { xi[a] = sm.xr[lf.dim][r_off + a]; xj[a] = sm.xc[lf.dim][c_off + a]; }
        double sv, sl, sa;
        wv_leaf_grad_sums(lf, sm.theta, xi, xj, wo, sv, sl, sa);
        for (int o = 16; o > 0; o >>= 1) {
          sv += __shfl_xor_sync(0xffffffffu, sv, o);
          sl += __shfl_xor_sync(0xffffffffu, sl, o);
          sa += __shfl_xor_sync(0xffffffffu, sa, o);
        }
        if (lane == 0) {   // several leaves may share a slot: accumulate (warp-private column, no race)
          if (tv) sm.red[lf.s_var][warp] += sv;
          if (tl) sm.red[lf.s_ls][warp] += sl;
          if (ta) sm.red[lf.s_aux][warp] += sa;
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) trw += __shfl_xor_sync(0xffffffffu, trw, o);
    if (lane == 0) sm.red[sm.pg.noise_slot][warp] += trw;
  }
  __syncthreads();
  const int ntiles = gridDim.x;
  double* dst = bd.partial + ((size_t)b * ntiles + blockIdx.x) * bd.n_slots_max;
  for (int s = threadIdx.x; s < sm.pg.n_slots; s += blockDim.x) {
    double t = 0.0;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) t += sm.red[s][wq];
    dst[s] = t;
  }
}

// =============================================================================================
// finalize: f = -(LML + log prior), df/dx; status bits.  grid (n_active), 64 threads (one per slot).
// =============================================================================================
__global__ void __launch_bounds__(64) wv_finalize_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                         const double* __restrict__ xall, int ntiles,
                                                         double* __restrict__ f_out, double* __restrict__ g_out,
                                                         double* __restrict__ lml_out, int* __restrict__ status_out) {
  __shared__ double s_lp[64];
  __shared__ double s_sum_alpha;
  __shared__ int s_bad;
  const int b = active[blockIdx.x];
  const WvProgram* pg = bd.programs + bd.prog_id[b];
  const double* x = xall + (size_t)b * bd.P;
  const int s = threadIdx.x;
  if (s == 0) s_bad = 0;
  // sum(alpha) for the mean gradient (fixed order: lane-strided then tree)
  {
    double t = 0.0;
    const double* al = bd.alpha + (size_t)b * bd.npad;
    if (s < 32) {
      for (int i = s; i < bd.n; i += 32) t += al[i];
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (s == 0) s_sum_alpha = t;
    }
  }
  __syncthreads();
  double lp = 0.0;
  if (s < pg->n_slots) {
    const WvSlot sl = pg->slots[s];
    if (sl.xindex >= 0) {
      const double u = x[sl.xindex];
      const double v = wv_transform(sl.transform, u, sl.shift);
      double dlp;
      wv_prior(sl, v, &lp, &dlp);
      double dl = 0.0;
      const double* part = bd.partial + (size_t)b * ntiles * bd.n_slots_max + s;
      for (int t = 0; t < ntiles; ++t) dl += part[(size_t)t * bd.n_slots_max];
      dl *= 0.5;
      if (s == pg->mean_slot) dl = s_sum_alpha;
      const double g = -(dl + dlp) * wv_transform_grad(sl.transform, u);
      g_out[(size_t)b * bd.P + sl.xindex] = g;
      if (!isfinite(g)) atomicOr(&s_bad, 1);
    }
  }
  s_lp[s] = lp;
  __syncthreads();
  if (s == 0) {
    double lps = 0.0;
    for (int i = 0; i < pg->n_slots; ++i) lps += s_lp[i];
    double logdet = 0.0;
    for (int jb = 0; jb < bd.nt; ++jb) logdet += bd.logdet_part[(size_t)b * bd.nt + jb];
    const double lml = -0.5 * bd.quad[b] - 0.5 * bd.n * 1.8378770664093453 - logdet;
    const double f = -(lml + lps);
    f_out[b] = f;
    lml_out[b] = lml;
    int st = 0;
    if (bd.chol_fail[b]) st |= WV_STATUS_CHOL_FAIL;
    if (!isfinite(f) || s_bad) st |= WV_STATUS_NONFINITE;
    status_out[b] = st;
    for (int i = pg->n_x; i < bd.P; ++i) g_out[(size_t)b * bd.P + i] = 0.0;
  }
}

// =============================================================================================
// host-side launch sequence of one evaluation (enqueued on `stream`, no host sync)
// =============================================================================================
size_t wv_smem_gemm_bytes() { return sizeof(WvPanelSmem) > sizeof(WvDiagSmem) ? sizeof(WvPanelSmem) : sizeof(WvDiagSmem); }

static bool g_attr_done = false;
static cudaError_t wv_set_attrs() {
  if (g_attr_done) return cudaSuccess;
  cudaError_t e;
#define WV_ATTR(k, bytes) \
  e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
  if (e != cudaSuccess) return e;
  WV_ATTR(wv_gram_kernel, sizeof(WvElemSmem));
  WV_ATTR(wv_grad_kernel, sizeof(WvElemSmem));
  WV_ATTR(wv_chol_diag_kernel, sizeof(WvDiagSmem));
  WV_ATTR(wv_panel_kernel<0>, sizeof(WvPanelSmem));
  WV_ATTR(wv_panel_kernel<1>, sizeof(WvPanelSmem));
  WV_ATTR(wv_kinv_kernel, sizeof(WvGemmSmem));
#undef WV_ATTR
  g_attr_done = true;
  return cudaSuccess;
}

// Returns the number of kernel launches enqueued (for bench.py's gpu_launches), or -1 on error.
int wv_enqueue_eval(const WvBatchDev& bd, const int* d_active, int n_active, const double* d_x, double* d_f,
                    double* d_g, double* d_lml, int* d_status, cudaStream_t st, WvProfiler* pf) {
  if (n_active <= 0) return 0;
  if (wv_set_attrs() != cudaSuccess) return -1;
  int launches = 0;
  const int nt = bd.nt;
  const int ntiles = nt * (nt + 1) / 2;
  WvProfiler none;
  if (!pf) pf = &none;
  cudaMemsetAsync(bd.chol_fail, 0, sizeof(int) * bd.B, st);
  pf->mark(-1, st);
  wv_gram_kernel<<<dim3(ntiles, n_active), WV_ELEM_THREADS, sizeof(WvElemSmem), st>>>(bd, d_active, d_x);
  pf->mark(WV_K_GRAM, st);
  ++launches;
  for (int j = 0; j < nt; ++j) {
    wv_chol_diag_kernel<<<dim3(n_active), WV_GEMM_THREADS, sizeof(WvDiagSmem), st>>>(bd, d_active, j);
    pf->mark(WV_K_CHOL_DIAG, st);
    ++launches;
    if (j + 1 < nt) {
      wv_panel_kernel<0><<<dim3(nt - j - 1, n_active), WV_GEMM_THREADS, sizeof(WvPanelSmem), st>>>(bd, d_active, j);
      pf->mark(WV_K_CHOL_PANEL, st);
      ++launches;
    }
  }
  for (int i = 1; i < nt; ++i) {
    wv_panel_kernel<1><<<dim3(i, n_active), WV_GEMM_THREADS, sizeof(WvPanelSmem), st>>>(bd, d_active, i);
    pf->mark(WV_K_TRTRI, st);
    ++launches;
  }
  wv_extract_kernel<<<dim3(n_active), 256, 0, st>>>(bd, d_active);
  pf->mark(WV_K_EXTRACT, st);
  wv_kinv_kernel<<<dim3(ntiles, n_active), WV_GEMM_THREADS, sizeof(WvGemmSmem), st>>>(bd, d_active);
  pf->mark(WV_K_KINV, st);
  wv_grad_kernel<<<dim3(ntiles, n_active), WV_ELEM_THREADS, sizeof(WvElemSmem), st>>>(bd, d_active, d_x);
  pf->mark(WV_K_GRAD, st);
  wv_finalize_kernel<<<dim3(n_active), 64, 0, st>>>(bd, d_active, d_x, ntiles, d_f, d_g, d_lml, d_status);
  pf->mark(WV_K_FINALIZE, st);
  launches += 4;
  if (cudaGetLastError() != cudaSuccess) return -1;
  return launches;
}
