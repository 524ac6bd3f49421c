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
// chol_diag(j): T = A[j,j] - sum_{k0<=k<j} L[j,k] L[j,k]^T ; L_jj = chol(T) ; Linv_jj = L_jj^{-1}
// (k0 = 0: left-looking over the whole row; k0 = first column of the current panel on the large-n path, where the
// columns before it have already been applied by the right-looking trailing updates)
// grid (n_active), 128 threads.  Row n (the RHS row) takes part as an ordinary row but is never a pivot.
//
// This kernel is the serial link of the factorisation (one CTA per model and step), so it is built for latency:
// the 64x64 block is processed in four 16-column steps.  Only the 16x16 diagonal Cholesky + triangular inverse is
// scalar work (one warp, rows in registers, pivots by shuffle); the panel scaling L_ik = T_ik X_kk^T, the trailing
// update T_ij -= L_ik L_jk^T and the assembly of the 64x64 inverse by recursive doubling (X_SF = -X_SS (L_SF X_FF))
// are 8x8x4 DMMA tile products on shared-memory operands spread over the four warps.
// =============================================================================================
struct WvDiagSmem {
  union {
    WvGemmSmem g;
    struct {
      double T[WV_NB * WV_LDT];   // T -> L (lower)
      double X[WV_NB * WV_LDT];   // L^{-1} (lower); the upper off-diagonal blocks are scratch for (L_SF X_FF)^T
    } e;
  };
  double invd[16];
  double logsum;
  int fail;
};

// 16x16 Cholesky in registers: lane r < 16 holds row r in a[0..15] (lower part meaningful).  `rhs` (warp-uniform,
// outside [0,16) if none) is the local index of the augmented RHS row: unit diagonal, never a pivot.  The dependent
// chain per column is shuffle -> rsqrt -> multiply -> shuffle -> fma; logs are taken afterwards, one pivot per lane.
// Returns sum of log(diag) over columns c < nreal; sets fail if a pivot is <= 0 (NaN pivots flow through, as in
// Eigen's LLT).  myinv = 1 / L[lane][lane].
__device__ __forceinline__ double wv_potrf16(double (&a)[16], double& myinv, int rhs, int nreal, bool& fail) {
  const int lane = threadIdx.x & 31;
  double mypiv = 1.0;
  myinv = 1.0;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    double d = __shfl_sync(0xffffffffu, a[c], c);
    if (c == rhs) d = 1.0;
    if (d <= 0.0) fail = true;
    const double inv = rsqrt(d);
    double l = a[c] * inv;                 // lane c: d * rsqrt(d) = sqrt(d)
    if (lane == c) { mypiv = d; myinv = inv; }
    if (lane < c) l = 0.0;
    a[c] = l;
#pragma unroll
    for (int c2 = c + 1; c2 < 16; ++c2) {
      const double v = __shfl_sync(0xffffffffu, l, c2);
      a[c2] = fma(-l, v, a[c2]);
    }
  }
  double lg = lane < nreal ? 0.5 * log(mypiv) : 0.0;
  for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
  return lg;
}

// 16x16 lower-triangular inverse: L (row stride ldl) and 1/diag(L) are read from shared memory with warp-uniform
// (broadcast) loads; lane j < 16 produces column j of X = L^{-1} in x[0..15] (x[r] = X[r][j], zero for r < j).
__device__ __forceinline__ void wv_trtri16(const double* __restrict__ Ls, int ldl, const double* __restrict__ invd,
                                           double (&x)[16]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
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

// One 8x8 DMMA output tile from shared-memory operands: (c0, c1) += sum_{k in [ka, kb)} A[fr][k] * B(k, fr), with
// B(k, n) = Bp[n * ldb + k] (BT = false, "NT") or Bp[k * ldb + n] (BT = true).  ka, kb multiples of 4.  The thread
// holds the elements (row fr, cols 2 fk, 2 fk + 1) of the tile.  Strides congruent to 4 mod 16 make all three access
// patterns bank-conflict free.
template <bool BT, bool NEGA>
__device__ __forceinline__ void wv_tile8(const double* __restrict__ Ap, int lda, const double* __restrict__ Bp, int ldb,
                                         int ka, int kb, double& c0, double& c1) {
  const int lane = threadIdx.x & 31, fr = lane >> 2, fk = lane & 3;
  for (int k = ka; k < kb; k += 4) {
    double a = Ap[fr * lda + k + fk];
    const double b = BT ? Bp[(k + fk) * ldb + fr] : Bp[fr * ldb + k + fk];
    if (NEGA) a = -a;
    wv_dmma(c0, c1, a, b);
  }
}

// X_SF = -X_SS (L_SF X_FF) for the diagonal sub-blocks F = [f0, f0+h), S = [f0+h, f0+2h) of the 64x64 block
// (recursive doubling of the triangular inverse).  Two phases separated by the caller's barrier:
//   phase 0: P = L_SF X_FF, stored transposed in the (otherwise unused) upper block X[F rows][S cols]
//   phase 1: X_SF = -X_SS P
// `w`/`nw`: index of this warp among the nw warps that share the block's h/8 x h/8 output tiles.
__device__ __forceinline__ void wv_inv_couple(double* __restrict__ T, double* __restrict__ X, int f0, int h, int phase,
                                              int w, int nw) {
  const int lane = threadIdx.x & 31, fr = lane >> 2, fk = lane & 3;
  const int s0 = f0 + h, nb = h >> 3;
  for (int t = w; t < nb * nb; t += nw) {
    const int mi = t / nb, ni = t % nb;
    double c0 = 0.0, c1 = 0.0;
    if (phase == 0) {
      // P[m][n] = sum_{k >= n} L[s0+m][f0+k] X[f0+k][f0+n]      (X_FF lower triangular: k from the tile's first column)
      wv_tile8<true, false>(T + (s0 + mi * 8) * WV_LDT + f0, WV_LDT, X + f0 * WV_LDT + f0 + ni * 8, WV_LDT, ni * 8, h, c0, c1);
      X[(f0 + ni * 8 + 2 * fk) * WV_LDT + s0 + mi * 8 + fr] = c0;          // P^T
      X[(f0 + ni * 8 + 2 * fk + 1) * WV_LDT + s0 + mi * 8 + fr] = c1;
    } else {
      // X_SF[m][n] = -sum_{k <= m} X[s0+m][s0+k] P[k][n],  P[k][n] = X[f0+n][s0+k]
      wv_tile8<false, true>(X + (s0 + mi * 8) * WV_LDT + s0, WV_LDT, X + (f0 + ni * 8) * WV_LDT + s0, WV_LDT, 0, mi * 8 + 8, c0, c1);
      *reinterpret_cast<double2*>(&X[(s0 + mi * 8 + fr) * WV_LDT + f0 + ni * 8 + 2 * fk]) = make_double2(c0, c1);
    }
  }
}

__global__ void __launch_bounds__(WV_GEMM_THREADS, 3) wv_chol_diag_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                                       int j, int k0) {
  WvDiagSmem& sm = *reinterpret_cast<WvDiagSmem*>(wv_smem_raw);
  const int b = active[blockIdx.x];
  const int ld = bd.npad;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  double* Ab = bd.A + (size_t)b * ld * ld;
  const double* Lrow = Ab + (size_t)j * WV_NB * ld;
  {
    double acc[4][4][2];
    wv_zero_acc(acc);
    if (threadIdx.x == 0) { sm.fail = 0; sm.logsum = 0.0; }
    if (j * WV_NB > k0) wv_gemm_nt_64(sm.g, Lrow, Lrow, ld, k0, j * WV_NB, acc);
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

  for (int kb = 0; kb < 4; ++kb) {
    const int o = kb * 16;
    // ---- (a) 16x16 diagonal block: Cholesky + inverse, warp 0
    if (warp == 0) {
      bool fail = false;
      double a[16], myinv;
      const int row = lane & 15;
#pragma unroll
      for (int c = 0; c < 16; ++c) a[c] = T[(o + row) * WV_LDT + o + c];
      const double ls = wv_potrf16(a, myinv, rhs - o, nreal - o, fail);
      if (lane < 16) {
#pragma unroll
        for (int c = 0; c < 16; ++c) T[(o + lane) * WV_LDT + o + c] = a[c];      // L_kk (zeros above the diagonal)
        sm.invd[lane] = myinv;
      }
      if (lane == 0) { sm.logsum += ls; if (fail) sm.fail = 1; }
      __syncwarp();
      double x[16];
      wv_trtri16(T + o * WV_LDT + o, WV_LDT, sm.invd, x);
      if (lane < 16) {
#pragma unroll
        for (int r = 0; r < 16; ++r) X[(o + r) * WV_LDT + o + lane] = x[r];      // X_kk[r][j = lane]
      }
    }
    __syncthreads();
    if (kb == 3) break;
    // ---- (b) L_ik = T_ik X_kk^T for the row blocks below: one warp owns whole 8-row blocks, so it may overwrite T_ik
    const int nrb = (WV_NB - o - 16) >> 3;
    for (int rb = warp; rb < nrb; rb += 4) {
      const int m0 = o + 16 + rb * 8;
      double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
      wv_tile8<false, false>(T + m0 * WV_LDT + o, WV_LDT, X + o * WV_LDT + o, WV_LDT, 0, 8, c[0][0], c[0][1]);
      wv_tile8<false, false>(T + m0 * WV_LDT + o, WV_LDT, X + (o + 8) * WV_LDT + o, WV_LDT, 0, 16, c[1][0], c[1][1]);
      __syncwarp();
      *reinterpret_cast<double2*>(&T[(m0 + fr) * WV_LDT + o + 2 * fk]) = make_double2(c[0][0], c[0][1]);
      *reinterpret_cast<double2*>(&T[(m0 + fr) * WV_LDT + o + 8 + 2 * fk]) = make_double2(c[1][0], c[1][1]);
    }
    __syncthreads();
    // ---- (c) trailing update of the lower 8x8 tiles: T_ij -= L_ik L_jk^T
    const int ntl = nrb * (nrb + 1) / 2;
    for (int t = warp; t < ntl; t += 4) {
      int ti = 0;
      while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
      const int tj = t - ti * (ti + 1) / 2;
      const int m0 = o + 16 + ti * 8, n0 = o + 16 + tj * 8;
      double2* cp = reinterpret_cast<double2*>(&T[(m0 + fr) * WV_LDT + n0 + 2 * fk]);
      double2 cv = *cp;
      wv_tile8<false, true>(T + m0 * WV_LDT + o, WV_LDT, T + n0 * WV_LDT + o, WV_LDT, 0, 16, cv.x, cv.y);
      *cp = cv;
    }
    __syncthreads();
  }
  // ---- assemble the 64x64 inverse: level 1 (16-blocks: pairs (0,1) and (2,3)), level 2 (32-blocks)
  for (int phase = 0; phase < 2; ++phase) {
    wv_inv_couple(T, X, (warp >> 1) * 32, 16, phase, warp & 1, 2);
    __syncthreads();
  }
  for (int phase = 0; phase < 2; ++phase) {
    wv_inv_couple(T, X, 0, 32, phase, warp, 4);
    __syncthreads();
  }

  // ---- write L_jj (lower, zeros above), Linv_jj (row-major) and Linv_jj^T into Mt[j,j]
  double* Tg = Ab + (size_t)j * WV_NB * ld + j * WV_NB;
  double* Mg = bd.Mt + (size_t)b * ld * ld + (size_t)j * WV_NB * ld + j * WV_NB;
  double* Dg = bd.Dinv + ((size_t)b * bd.nt + j) * WV_NB * WV_NB;
  for (int i = threadIdx.x; i < WV_NB * WV_NB / 2; i += WV_GEMM_THREADS) {
    const int rr = i >> 5, cc = (i & 31) * 2;
    const double2 tv = *reinterpret_cast<const double2*>(&T[rr * WV_LDT + cc]);
    const double2 xv = *reinterpret_cast<const double2*>(&X[rr * WV_LDT + cc]);
    *reinterpret_cast<double2*>(Tg + (size_t)rr * ld + cc) = make_double2(cc <= rr ? tv.x : 0.0, cc + 1 <= rr ? tv.y : 0.0);
    *reinterpret_cast<double2*>(Dg + rr * WV_NB + cc) = make_double2(cc <= rr ? xv.x : 0.0, cc + 1 <= rr ? xv.y : 0.0);
    const double m0 = rr <= cc ? X[cc * WV_LDT + rr] : 0.0, m1 = rr <= cc + 1 ? X[(cc + 1) * WV_LDT + rr] : 0.0;
    *reinterpret_cast<double2*>(Mg + (size_t)rr * ld + cc) = make_double2(m0, m1);
  }
  if (threadIdx.x == 0) {
    bd.logdet_part[(size_t)b * bd.nt + j] = sm.logsum;
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
                                                                   int step, int kstart) {
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
    k0 = kstart; k1 = j * WV_NB;
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
// Large-n path (nt >= WvAux::big_nt, e.g. config 4: one n = 8192 model).  The left-looking batched scheme above has
// one CTA per model on its critical path; for few, large models the factorisation is reorganised so that almost all
// flops sit in wide launches of the same 64x64 DMMA tile GEMM:
//   Cholesky   right-looking over panels of WV_PANEL_TILES tile columns: the panel is factorised with the kernels
//              above (k0 = first panel column), then `wv_syrk_kernel` applies it to the trailing matrix.  The update
//              is split into the columns of the NEXT panel (main stream) and the rest (side stream), so that the next
//              panel factorisation overlaps the bulk of the update (look-ahead of one panel).
//   L^{-T}     recursive doubling: at level m the inverse of every aligned 2m-tile diagonal block is assembled from
//              its two m-tile halves, Mt_FS = -Mt_FF (L_SF^T Mt_SS), as two launches over all blocks of the level
//              (`wv_trtri_level_kernel<1>`: U = Mt_FF L_SF^T into the unused upper tiles of A;  <2>: Mt_FS = -U Mt_SS,
//              the only product on the path whose B operand is not k-contiguous -> wv_gemm_64<true>).
// =============================================================================================
#define WV_PANEL_TILES 4

// A[ti,tj] -= sum_{k in [k0,k1)} L[ti,k] L[tj,k]^T   for tj in [c_lo, c_hi), ti in [tj, nt).
// c_hi == nt: triangular enumeration of the whole trailing block; otherwise a (nt - c_lo) x (c_hi - c_lo) rectangle
// whose above-diagonal CTAs exit.  grid (tiles, n_active), 128 threads.
__global__ void __launch_bounds__(WV_GEMM_THREADS) wv_syrk_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                                  int c_lo, int c_hi, int k0, int k1) {
  WvGemmSmem& sm = *reinterpret_cast<WvGemmSmem*>(wv_smem_raw);
  const int b = active[blockIdx.y];
  const int ld = bd.npad;
  int ti, tj;
  if (c_hi == bd.nt) {
    wv_tile_from_linear(blockIdx.x, ti, tj);
  } else {
    const int w = c_hi - c_lo;
    ti = blockIdx.x / w; tj = blockIdx.x % w;
    if (ti < tj) return;
  }
  ti += c_lo; tj += c_lo;
  double* Ab = bd.A + (size_t)b * ld * ld;
  double acc[4][4][2];
  wv_zero_acc(acc);
  wv_gemm_nt_64(sm, Ab + (size_t)ti * WV_NB * ld, Ab + (size_t)tj * WV_NB * ld, ld, k0, k1, acc);
  int r0, c0;
  wv_frag_origin(r0, c0);
  double* Out = Ab + (size_t)ti * WV_NB * ld + tj * WV_NB;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      double2* p = reinterpret_cast<double2*>(Out + (size_t)(r0 + mi * 8) * ld + c0 + ni * 8);
      double2 v = *p;
      *p = make_double2(v.x - acc[mi][ni][0], v.y - acc[mi][ni][1]);
    }
}

// level m (in tiles): block q has F = [2qm, (2q+1)m), S = [(2q+1)m, min((2q+2)m, nt)).
// grid (n_blocks * m * m, n_active): tile (tF, tS), tF in F, tS in S.
//   STEP 1: U[tF,tS]  =  sum_{k in [tF, end F)} Mt[tF,k] L[tS,k]^T          -> A[tF,tS]  (upper tile, scratch)
//   STEP 2: Mt[tF,tS] = -sum_{k in [start S, tS]} U[tF,k] Mt[k,tS]          (NN product)
template <int STEP>
__global__ void __launch_bounds__(WV_GEMM_THREADS) wv_trtri_level_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                                         int m) {
  WvGemmSmem& sm = *reinterpret_cast<WvGemmSmem*>(wv_smem_raw);
  const int b = active[blockIdx.y];
  const int ld = bd.npad;
  const int q = blockIdx.x / (m * m), r = blockIdx.x % (m * m);
  const int tF = 2 * q * m + r / m, tS = (2 * q + 1) * m + r % m;
  if (tS >= bd.nt) return;
  const int mid = (2 * q + 1) * m;
  double* Ab = bd.A + (size_t)b * ld * ld;
  double* Mb = bd.Mt + (size_t)b * ld * ld;
  double acc[4][4][2];
  wv_zero_acc(acc);
  double* Out;
  if (STEP == 1) {
    wv_gemm_64<false>(sm, Mb + (size_t)tF * WV_NB * ld, Ab + (size_t)tS * WV_NB * ld, ld, tF * WV_NB, mid * WV_NB, acc);
    Out = Ab + (size_t)tF * WV_NB * ld + tS * WV_NB;
  } else {
    wv_gemm_64<true>(sm, Ab + (size_t)tF * WV_NB * ld, Mb + tS * WV_NB, ld, mid * WV_NB, (tS + 1) * WV_NB, acc);
    Out = Mb + (size_t)tF * WV_NB * ld + tS * WV_NB;
  }
  int r0, c0;
  wv_frag_origin(r0, c0);
  const double sgn = STEP == 1 ? 1.0 : -1.0;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
      *reinterpret_cast<double2*>(Out + (size_t)(r0 + mi * 8) * ld + c0 + ni * 8) =
          make_double2(sgn * acc[mi][ni][0], sgn * acc[mi][ni][1]);
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
// finalize: f = -(LML + log prior), df/dx; status bits.  grid (n_active), 64 * G threads: thread (g, s) sums the
// per-tile partials of slot s over tiles t = g mod G (fixed order), the G group sums are added in fixed order.
// G = 1 up to 128 tiles; the large-n path (thousands of tiles per model) uses G = 16.
// =============================================================================================
#define WV_FIN_MAXG 16
__global__ void __launch_bounds__(64 * WV_FIN_MAXG) wv_finalize_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                         const double* __restrict__ xall, int ntiles,
                                                         double* __restrict__ f_out, double* __restrict__ g_out,
                                                         double* __restrict__ lml_out, int* __restrict__ status_out) {
  __shared__ double s_lp[64];
  __shared__ double s_part[WV_FIN_MAXG][64];
  __shared__ double s_sum_alpha;
  __shared__ int s_bad;
  const int b = active[blockIdx.x];
  const WvProgram* pg = bd.programs + bd.prog_id[b];
  const double* x = xall + (size_t)b * bd.P;
  const int s = threadIdx.x & 63, grp = threadIdx.x >> 6, G = blockDim.x >> 6;
  if (threadIdx.x == 0) s_bad = 0;
  // sum(alpha) for the mean gradient (fixed order: lane-strided then tree)
  if (threadIdx.x < 32) {
    double t = 0.0;
    const double* al = bd.alpha + (size_t)b * bd.npad;
    for (int i = threadIdx.x; i < bd.n; i += 32) t += al[i];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) s_sum_alpha = t;
  }
  {
    double dl = 0.0;
    if (s < pg->n_slots) {
      const double* part = bd.partial + (size_t)b * ntiles * bd.n_slots_max + s;
      for (int t = grp; t < ntiles; t += G) dl += part[(size_t)t * bd.n_slots_max];
    }
    s_part[grp][s] = dl;
  }
  __syncthreads();
  double lp = 0.0;
  if (grp == 0 && s < pg->n_slots) {
    const WvSlot sl = pg->slots[s];
    if (sl.xindex >= 0) {
      const double u = x[sl.xindex];
      const double v = wv_transform(sl.transform, u, sl.shift);
      double dlp;
      wv_prior(sl, v, &lp, &dlp);
      double dl = 0.0;
      for (int g = 0; g < G; ++g) dl += s_part[g][s];
      dl *= 0.5;
      if (s == pg->mean_slot) dl = s_sum_alpha;
      const double g = -(dl + dlp) * wv_transform_grad(sl.transform, u);
      g_out[(size_t)b * bd.P + sl.xindex] = g;
      if (!isfinite(g)) atomicOr(&s_bad, 1);
    }
  }
  if (grp == 0) s_lp[s] = lp;
  __syncthreads();
  if (threadIdx.x == 0) {
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
  WV_ATTR(wv_syrk_kernel, sizeof(WvGemmSmem));
  WV_ATTR(wv_trtri_level_kernel<1>, sizeof(WvGemmSmem));
  WV_ATTR(wv_trtri_level_kernel<2>, sizeof(WvGemmSmem));
#undef WV_ATTR
  g_attr_done = true;
  return cudaSuccess;
}

// Cholesky + L^{-T} of the large-n path (see the comment above wv_syrk_kernel).  Returns launches or -1.
static int wv_enqueue_factor_big(const WvBatchDev& bd, const int* d_active, int n_active, cudaStream_t st,
                                 WvProfiler* pf, const WvAux& aux) {
  const int nt = bd.nt;
  int launches = 0;
  bool bulk_pending = false;
  for (int p0 = 0; p0 < nt; p0 += WV_PANEL_TILES) {
    const int p1 = p0 + WV_PANEL_TILES < nt ? p0 + WV_PANEL_TILES : nt;
    for (int j = p0; j < p1; ++j) {
      wv_chol_diag_kernel<<<dim3(n_active), WV_GEMM_THREADS, sizeof(WvDiagSmem), st>>>(bd, d_active, j, p0 * WV_NB);
      pf->mark(WV_K_CHOL_DIAG, st);
      ++launches;
      if (j + 1 < nt) {
        wv_panel_kernel<0><<<dim3(nt - j - 1, n_active), WV_GEMM_THREADS, sizeof(WvPanelSmem), st>>>(bd, d_active, j,
                                                                                                   p0 * WV_NB);
        pf->mark(WV_K_CHOL_PANEL, st);
        ++launches;
      }
    }
    if (p1 >= nt) break;
    const int a_hi = p1 + WV_PANEL_TILES < nt ? p1 + WV_PANEL_TILES : nt;   // columns of the next panel
    cudaEventRecord(aux.ev_panel, st);
    // next panel's columns on the main stream (they also received the previous bulk update: wait for it)
    if (bulk_pending) cudaStreamWaitEvent(st, aux.ev_bulk, 0);
    if (a_hi == nt) {
      const int w = nt - p1;
      wv_syrk_kernel<<<dim3(w * (w + 1) / 2, n_active), WV_GEMM_THREADS, sizeof(WvGemmSmem), st>>>(
          bd, d_active, p1, nt, p0 * WV_NB, p1 * WV_NB);
      bulk_pending = false;
    } else {
      wv_syrk_kernel<<<dim3((nt - p1) * (a_hi - p1), n_active), WV_GEMM_THREADS, sizeof(WvGemmSmem), st>>>(
          bd, d_active, p1, a_hi, p0 * WV_NB, p1 * WV_NB);
      // the rest of the trailing matrix on the side stream, overlapping the next panel factorisation
      cudaStreamWaitEvent(aux.side, aux.ev_panel, 0);
      const int w = nt - a_hi;
      wv_syrk_kernel<<<dim3(w * (w + 1) / 2, n_active), WV_GEMM_THREADS, sizeof(WvGemmSmem), aux.side>>>(
          bd, d_active, a_hi, nt, p0 * WV_NB, p1 * WV_NB);
      cudaEventRecord(aux.ev_bulk, aux.side);
      bulk_pending = true;
      ++launches;
    }
    pf->mark(WV_K_CHOL_SYRK, st);
    ++launches;
  }
  if (bulk_pending) cudaStreamWaitEvent(st, aux.ev_bulk, 0);
  for (int m = 1; m < nt; m *= 2) {
    const int nblk = (nt + 2 * m - 1) / (2 * m);
    wv_trtri_level_kernel<1><<<dim3(nblk * m * m, n_active), WV_GEMM_THREADS, sizeof(WvGemmSmem), st>>>(bd, d_active, m);
    wv_trtri_level_kernel<2><<<dim3(nblk * m * m, n_active), WV_GEMM_THREADS, sizeof(WvGemmSmem), st>>>(bd, d_active, m);
    pf->mark(WV_K_TRTRI, st);
    launches += 2;
  }
  return launches;
}

// Returns the number of kernel launches enqueued (for bench.py's gpu_launches), or -1 on error.
int wv_enqueue_eval(const WvBatchDev& bd, const int* d_active, int n_active, const double* d_x, double* d_f,
                    double* d_g, double* d_lml, int* d_status, cudaStream_t st, WvProfiler* pf, const WvAux* aux) {
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
  if (aux && aux->side && nt >= aux->big_nt) {
    int l = wv_enqueue_factor_big(bd, d_active, n_active, st, pf, *aux);
    if (l < 0) return -1;
    launches += l;
  } else {
    for (int j = 0; j < nt; ++j) {
      wv_chol_diag_kernel<<<dim3(n_active), WV_GEMM_THREADS, sizeof(WvDiagSmem), st>>>(bd, d_active, j, 0);
      pf->mark(WV_K_CHOL_DIAG, st);
      ++launches;
      if (j + 1 < nt) {
        wv_panel_kernel<0><<<dim3(nt - j - 1, n_active), WV_GEMM_THREADS, sizeof(WvPanelSmem), st>>>(bd, d_active, j, 0);
        pf->mark(WV_K_CHOL_PANEL, st);
        ++launches;
      }
    }
    for (int i = 1; i < nt; ++i) {
      wv_panel_kernel<1><<<dim3(i, n_active), WV_GEMM_THREADS, sizeof(WvPanelSmem), st>>>(bd, d_active, i, 0);
      pf->mark(WV_K_TRTRI, st);
      ++launches;
    }
  }
  wv_extract_kernel<<<dim3(n_active), 256, 0, st>>>(bd, d_active);
  pf->mark(WV_K_EXTRACT, st);
  wv_kinv_kernel<<<dim3(ntiles, n_active), WV_GEMM_THREADS, sizeof(WvGemmSmem), st>>>(bd, d_active);
  pf->mark(WV_K_KINV, st);
  wv_grad_kernel<<<dim3(ntiles, n_active), WV_ELEM_THREADS, sizeof(WvElemSmem), st>>>(bd, d_active, d_x);
  pf->mark(WV_K_GRAD, st);
  wv_finalize_kernel<<<dim3(n_active), ntiles > 128 ? 64 * WV_FIN_MAXG : 64, 0, st>>>(bd, d_active, d_x, ntiles, d_f, d_g,
                                                                                  d_lml, d_status);
  pf->mark(WV_K_FINALIZE, st);
  launches += 4;
  if (cudaGetLastError() != cudaSuccess) return -1;
  return launches;
}
