// waveome_b200 — kernels of one batched LML+gradient evaluation.  See wv_kernels.cuh for the plan.
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include "wv_kernels.cuh"
#include "wv_rtc.h"

// dynamic shared memory is carved by hand; the GEMM pipeline and the epilogue tiles alias each other.
extern __shared__ __align__(16) unsigned char wv_smem_raw[];

// phase clocks of the serial kernels for scratch/diag_bench.cu (compiled out of the library)
#ifdef WV_DIAG_CLOCK
__device__ long long wv_dbg_clk[64];
#define WV_CLK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) wv_dbg_clk[i] = clock64(); } while (0)
#else
#define WV_CLK(i) do { } while (0)
#endif

#include "wv_elem.cuh"   // wv_gram_kernel, wv_grad_kernel

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
    WvGemmSmem g;                 // used as 2 * WV_STAGES single-operand stages by wv_syrk_self_64 (A and B are the same rows)
    struct {
      double T[WV_NB * WV_LDT];   // T -> L (lower, incl. diagonal); strict upper part: L^{-T} as it is assembled
      double X[WV_NB * WV_LDT];   // L^{-1} (lower); the upper off-diagonal blocks are scratch for (L_SF X_FF)^T
    } e;
  };
  double piv[WV_NB];
  double invd[16];
  double lcol[2][16];
  double red[2];
  int fail;
};

// Lower 8x8 tiles (i >= j) of the 64x64 product, spread over the four warps (10 / 9 / 9 / 8 tiles): the diagonal
// update is symmetric, so the strictly upper tiles are never formed.  X(q, i, j): accumulator q of this warp is tile
// (i, j).
#define WV_SYM_W0(X) X(0,0,0) X(1,1,0) X(2,1,1) X(3,2,0) X(4,2,1) X(5,2,2) X(6,3,0) X(7,3,1) X(8,3,2) X(9,3,3)
#define WV_SYM_W1(X) X(0,4,0) X(1,4,1) X(2,4,2) X(3,4,3) X(4,4,4) X(5,5,0) X(6,5,1) X(7,5,2) X(8,5,3)
#define WV_SYM_W2(X) X(0,5,4) X(1,5,5) X(2,6,0) X(3,6,1) X(4,6,2) X(5,6,3) X(6,6,4) X(7,6,5) X(8,6,6)
#define WV_SYM_W3(X) X(0,7,0) X(1,7,1) X(2,7,2) X(3,7,3) X(4,7,4) X(5,7,5) X(6,7,6) X(7,7,7)

// acc = lower tiles of sum_{k in [k0,k1)} R[m][k] R[n][k] for one 64-row operand R (row stride ld): the diagonal-tile
// update of the Cholesky.  Both DMMA operands come from ONE staged copy (2 * WV_STAGES stages in the memory of the A/B
// pairs): this kernel runs one CTA per model on the critical path and is bound by the FP64 tensor rate of a single
// SM and by bytes in flight, not by bandwidth.
#define WV_SELF_STAGES (2 * WV_STAGES)
__device__ __forceinline__ void wv_self_issue(double* __restrict__ stage, const double* __restrict__ Rg, int ld, int kc,
                                              int k1) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int q = threadIdx.x + WV_GEMM_THREADS * r;
    const int row = q >> 3, ch = q & 7;
    const int k = kc + ch * 2;
    const bool ok = k < k1;
    wv_cp_async16(&stage[row * WV_LDS + ch * 2], Rg + (size_t)row * ld + (ok ? k : kc), ok);
  }
}
__device__ __forceinline__ void wv_syrk_self_64(WvGemmSmem& sm, const double* __restrict__ Rg, int ld, int k0, int k1,
                                                double (&acc)[10][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  double* base = &sm.a[0][0];     // a[..] and b[..] are contiguous: 2 * WV_STAGES stages of 64 x WV_LDS doubles
  const int nchunks = (k1 - k0 + WV_BK - 1) / WV_BK;
#pragma unroll
  for (int s = 0; s < WV_SELF_STAGES - 1; ++s) {
    if (s < nchunks) wv_self_issue(base + s * (WV_NB * WV_LDS), Rg, ld, k0 + s * WV_BK, k1);
    wv_cp_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    wv_cp_wait<WV_SELF_STAGES - 2>();
    __syncthreads();
    {
      const int cn = c + WV_SELF_STAGES - 1;
      if (cn < nchunks) wv_self_issue(base + (cn % WV_SELF_STAGES) * (WV_NB * WV_LDS), Rg, ld, k0 + cn * WV_BK, k1);
      wv_cp_commit();
    }
    const double* st = base + (c % WV_SELF_STAGES) * (WV_NB * WV_LDS) + fr * WV_LDS + fk;
#pragma unroll
    for (int kk = 0; kk < WV_BK; kk += 4) {
      double f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = st[i * 8 * WV_LDS + kk];
#define WV_X(q, i, j) wv_dmma(acc[q][0], acc[q][1], f[i], f[j]);
      if (warp == 0) { WV_SYM_W0(WV_X) }
      else if (warp == 1) { WV_SYM_W1(WV_X) }
      else if (warp == 2) { WV_SYM_W2(WV_X) }
      else { WV_SYM_W3(WV_X) }
#undef WV_X
    }
  }
  wv_cp_wait<0>();
  __syncthreads();
}

// 16x16 Cholesky in the registers of one warp: lane r < 16 holds row r of the block in a[0..15] (lower part
// meaningful).  `rhs` (warp-uniform, outside [0,16) if none) is the local index of the augmented RHS row: unit
// diagonal, never a pivot.  Per column the dependent chain is  pivot shuffle -> rsqrt -> multiply -> fma (next pivot,
// formed by its own lane before anything else); the rank-1 update of the other columns reads the column through
// shared memory (broadcast loads, double buffered) and stays off that chain.  Sets fail if a pivot is <= 0 (NaN pivots
// flow through, as in Eigen's LLT).  mypiv / myinv: pivot of row `lane` and 1 / L[lane][lane].
__device__ __forceinline__ void wv_potrf16(double (&a)[16], double (*lcol)[16], double& mypiv, double& myinv, int rhs,
                                           bool& fail) {
  const int lane = threadIdx.x & 31;
  mypiv = 1.0; myinv = 1.0;
  double d = __shfl_sync(0xffffffffu, a[0], 0);
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    if (c == rhs) d = 1.0;
    if (d <= 0.0) fail = true;
    const double inv = rsqrt(d);
    double l = a[c] * inv;                 // lane c: d * rsqrt(d) = sqrt(d)
    if (lane == c) { mypiv = d; myinv = inv; }
    if (lane < c) l = 0.0;
    a[c] = l;
    if (c < 15) {
      const double dn = fma(-l, l, a[c + 1]);          // lane c+1: its own diagonal entry = the next pivot
      d = __shfl_sync(0xffffffffu, dn, c + 1);
      if (lane < 16) lcol[c & 1][lane] = l;
      __syncwarp();
#pragma unroll
      for (int c2 = c + 1; c2 < 16; ++c2) a[c2] = fma(-l, lcol[c & 1][c2], a[c2]);
    }
    WV_CLK(20 + c);
  }
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

// X_SF = -X_SS (L_SF X_FF) for the diagonal sub-blocks F = [f0, f0+h), S = [f0+h, f0+2h) of the 64x64 block
// (recursive doubling of the triangular inverse), h = 8 NBT.  Two phases (the caller synchronises between them):
//   phase 0: P = L_SF X_FF, stored transposed in the (otherwise unused) upper block X[F rows][S cols];
//            w = tile COLUMN (all its NBT tiles share the k range [8w, h): X_FF is lower triangular)
//   phase 1: X_SF = -X_SS P, also stored transposed into the strict upper part of T (-> Mt);
//            w = tile ROW (k range [0, 8w + 8): X_SS is lower triangular)
// The NBT tiles of a call are independent accumulators, interleaved for latency.
template <int NBT>
__device__ __forceinline__ void wv_inv_couple(double* __restrict__ T, double* __restrict__ X, int f0, int phase, int w) {
  const int lane = threadIdx.x & 31, fr = lane >> 2, fk = lane & 3;
  const int h = NBT * 8, s0 = f0 + h;
  double c[NBT][2];
#pragma unroll
  for (int q = 0; q < NBT; ++q) c[q][0] = c[q][1] = 0.0;
  if (phase == 0) {
    const int ni = w;
    for (int k = ni * 8; k < h; k += 4) {
      const double bv = X[(f0 + k + fk) * WV_LDT + f0 + ni * 8 + fr];          // X_FF[k][n]
#pragma unroll
      for (int mi = 0; mi < NBT; ++mi) wv_dmma(c[mi][0], c[mi][1], T[(s0 + mi * 8 + fr) * WV_LDT + f0 + k + fk], bv);
    }
#pragma unroll
    for (int mi = 0; mi < NBT; ++mi) {                                          // P^T
      X[(f0 + ni * 8 + 2 * fk) * WV_LDT + s0 + mi * 8 + fr] = c[mi][0];
      X[(f0 + ni * 8 + 2 * fk + 1) * WV_LDT + s0 + mi * 8 + fr] = c[mi][1];
    }
  } else {
    const int mi = w;
    for (int k = 0; k < mi * 8 + 8; k += 4) {
      const double av = -X[(s0 + mi * 8 + fr) * WV_LDT + s0 + k + fk];          // -X_SS[m][k]
#pragma unroll
      for (int ni = 0; ni < NBT; ++ni) wv_dmma(c[ni][0], c[ni][1], av, X[(f0 + ni * 8 + fr) * WV_LDT + s0 + k + fk]);
    }
#pragma unroll
    for (int ni = 0; ni < NBT; ++ni) {
      *reinterpret_cast<double2*>(&X[(s0 + mi * 8 + fr) * WV_LDT + f0 + ni * 8 + 2 * fk]) = make_double2(c[ni][0], c[ni][1]);
      T[(f0 + ni * 8 + 2 * fk) * WV_LDT + s0 + mi * 8 + fr] = c[ni][0];
      T[(f0 + ni * 8 + 2 * fk + 1) * WV_LDT + s0 + mi * 8 + fr] = c[ni][1];
    }
  }
}

// inverse of the 16x16 diagonal block at offset o by ONE warp: X_kk into X (lower), X_kk^T into the strict upper
// part of T's diagonal block (-> Mt)
__device__ __forceinline__ void wv_diag_block_inverse(double* __restrict__ T, double* __restrict__ X,
                                                      const double* __restrict__ invd, int o) {
  const int lane = threadIdx.x & 31;
  double x[16];
  wv_trtri16(T + o * WV_LDT + o, WV_LDT, invd, x);
  if (lane < 16) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      X[(o + r) * WV_LDT + o + lane] = x[r];                          // X_kk[r][j = lane]
      if (r > lane) T[(o + lane) * WV_LDT + o + r] = x[r];            // T[lane][r] = X[r][lane]
    }
  }
}

__device__ __forceinline__ void wv_diag_body(const WvBatchDev& bd, int b, int j, int k0, int epoch) {
  WvDiagSmem& sm = *reinterpret_cast<WvDiagSmem*>(wv_smem_raw);
  const int ld = bd.npad;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  double* Ab = bd.A + (size_t)b * ld * ld;
  const double* Lrow = Ab + (size_t)j * WV_NB * ld;
  double* Tg = Ab + (size_t)j * WV_NB * ld + j * WV_NB;
  WV_CLK(0);
  {
    double acc[10][2];
#pragma unroll
    for (int q = 0; q < 10; ++q) acc[q][0] = acc[q][1] = 0.0;
    if (threadIdx.x == 0) sm.fail = 0;
    if (j * WV_NB > k0) wv_syrk_self_64(sm.g, Lrow, ld, k0, j * WV_NB, acc);
    else __syncthreads();
#define WV_X(q, i, j)                                                                                         \
    {                                                                                                         \
      const double2 av = *reinterpret_cast<const double2*>(Tg + (size_t)((i) * 8 + fr) * ld + (j) * 8 + 2 * fk); \
      *reinterpret_cast<double2*>(&sm.e.T[((i) * 8 + fr) * WV_LDT + (j) * 8 + 2 * fk]) =                      \
          make_double2(av.x - acc[q][0], av.y - acc[q][1]);                                                   \
    }
    if (warp == 0) { WV_SYM_W0(WV_X) }
    else if (warp == 1) { WV_SYM_W1(WV_X) }
    else if (warp == 2) { WV_SYM_W2(WV_X) }
    else { WV_SYM_W3(WV_X) }
#undef WV_X
  }
  __syncthreads();
  WV_CLK(1);
  double* T = sm.e.T;
  double* X = sm.e.X;
  const int rhs = bd.n - j * WV_NB;                       // local index of the RHS row (may be outside [0,64))
  const int nreal = min(WV_NB, bd.n - j * WV_NB);         // pivots that belong to K (log-det terms)

  for (int kb = 0; kb < 4; ++kb) {
    const int o = kb * 16;
    // ---- (a) 16x16 diagonal block: Cholesky, warp 0 (the serial chain of the whole kernel)
    if (warp == 0) {
      bool fail = false;
      double a[16], mypiv, myinv;
      const int row = lane & 15;
      __syncwarp();                       // the block's last trailing update was written by other lanes of this warp
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        const double2 v = *reinterpret_cast<const double2*>(&T[(o + row) * WV_LDT + o + c]);
        a[c] = v.x; a[c + 1] = v.y;
      }
      WV_CLK(18);
      wv_potrf16(a, sm.lcol, mypiv, myinv, rhs - o, fail);
      WV_CLK(19);
      if (lane < 16) {
#pragma unroll
        for (int c = 0; c < 16; c += 2)
          if (c <= lane)      // on/below the diagonal only: the strict upper part will receive X_kk^T
            *reinterpret_cast<double2*>(&T[(o + lane) * WV_LDT + o + c]) =
                make_double2(a[c], c + 1 <= lane ? a[c + 1] : T[(o + lane) * WV_LDT + o + c + 1]);
        sm.piv[o + lane] = mypiv;
        sm.invd[lane] = myinv;
      }
      if (fail && lane == 0) sm.fail = 1;          // the pivots are warp-uniform, so is `fail`
      WV_CLK(36);
    }
    __syncthreads();
    WV_CLK(2 + kb * 3);
    const int nrows = WV_NB - o - 16;
    // ---- (b) rows below: solve x L_kk^T = t, one row per thread (warps 0..1), while warp 3 inverts the diagonal block
    if ((int)threadIdx.x < nrows) {
      double* trow = T + (o + 16 + threadIdx.x) * WV_LDT + o;
      double x[16];
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        const double2 v = *reinterpret_cast<const double2*>(trow + c);
        x[c] = v.x; x[c + 1] = v.y;
      }
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        double s0 = x[c], s1 = 0.0;
#pragma unroll
        for (int k = 0; k < c; ++k) {
          const double lv = T[(o + c) * WV_LDT + o + k];
          if (k & 1) s1 = fma(-lv, x[k], s1); else s0 = fma(-lv, x[k], s0);
        }
        x[c] = (s0 + s1) * sm.invd[c];
      }
#pragma unroll
      for (int c = 0; c < 16; c += 2) *reinterpret_cast<double2*>(trow + c) = make_double2(x[c], x[c + 1]);
    } else if (warp == 3) {
      wv_diag_block_inverse(T, X, sm.invd, o);
    }
    __syncthreads();
    WV_CLK(3 + kb * 3);
    if (kb == 3) break;
    // ---- (c) trailing update of the lower 8x8 tiles: T_ij -= L_ik L_jk^T.  Warp 0 takes the three tiles of the next
    //      diagonal block and goes straight on to factorise it; warps 1..3 share the rest (<= 6 independent tiles
    //      each); the barrier after the next (a) orders their writes before anybody reads them.
    {
      const int nrb = nrows >> 3;
      const int ntl = nrb * (nrb + 1) / 2;
      double2 cv[6];
      int mo[6], no[6];
      bool on[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const int t = warp == 0 ? q : 3 + (warp - 1) + 3 * q;
        on[q] = warp == 0 ? q < 3 : t < ntl;
        int ti = 0;
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        mo[q] = o + 16 + ti * 8; no[q] = o + 16 + (t - ti * (ti + 1) / 2) * 8;
        if (on[q]) cv[q] = *reinterpret_cast<const double2*>(&T[(mo[q] + fr) * WV_LDT + no[q] + 2 * fk]);
      }
#pragma unroll
      for (int k = 0; k < 16; k += 4) {
#pragma unroll
        for (int q = 0; q < 6; ++q)
          if (on[q]) wv_dmma(cv[q].x, cv[q].y, -T[(mo[q] + fr) * WV_LDT + o + k + fk], T[(no[q] + fr) * WV_LDT + o + k + fk]);
      }
#pragma unroll
      for (int q = 0; q < 6; ++q)
        if (on[q]) *reinterpret_cast<double2*>(&T[(mo[q] + fr) * WV_LDT + no[q] + 2 * fk]) = cv[q];
    }
    WV_CLK(4 + kb * 3);
  }
  // ---- assemble the 64x64 inverse.  Level 1 (16-blocks): warp 2 couples the pair (0,1), warp 3 the pair (2,3), each
  //      alone (both phases, warp-level sync only); meanwhile warps 0..1 write L_jj, which is final, to global memory.
  if (warp >= 2) {
    const int f0 = (warp - 2) * 32;
    wv_inv_couple<2>(T, X, f0, 0, 0);
    wv_inv_couple<2>(T, X, f0, 0, 1);
    __syncwarp();
    wv_inv_couple<2>(T, X, f0, 1, 0);
    wv_inv_couple<2>(T, X, f0, 1, 1);
  } else {
    for (int i = threadIdx.x; i < WV_NB * WV_NB / 2; i += 64) {
      const int rr = i >> 5, cc = (i & 31) * 2;
      const double2 tv = *reinterpret_cast<const double2*>(&T[rr * WV_LDT + cc]);
      *reinterpret_cast<double2*>(Tg + (size_t)rr * ld + cc) = make_double2(cc <= rr ? tv.x : 0.0, cc + 1 <= rr ? tv.y : 0.0);
    }
  }
  __syncthreads();
  //      Level 2 (32-blocks), all four warps.
  for (int phase = 0; phase < 2; ++phase) {
    wv_inv_couple<4>(T, X, 0, phase, warp);
    __syncthreads();
  }
  WV_CLK(12);

  // ---- write Linv_jj (row-major) and Linv_jj^T (-> Mt[j,j])
  double* Mg = bd.Mt + (size_t)b * ld * ld + (size_t)j * WV_NB * ld + j * WV_NB;
  double* Dg = bd.Dinv + ((size_t)b * bd.nt + j) * WV_NB * WV_NB;
  for (int i = threadIdx.x; i < WV_NB * WV_NB / 2; i += WV_GEMM_THREADS) {
    const int rr = i >> 5, cc = (i & 31) * 2;
    const double2 tv = *reinterpret_cast<const double2*>(&T[rr * WV_LDT + cc]);
    const double2 xv = *reinterpret_cast<const double2*>(&X[rr * WV_LDT + cc]);
    const double xd = X[rr * WV_LDT + rr];
    *reinterpret_cast<double2*>(Dg + rr * WV_NB + cc) = make_double2(cc <= rr ? xv.x : 0.0, cc + 1 <= rr ? xv.y : 0.0);
    *reinterpret_cast<double2*>(Mg + (size_t)rr * ld + cc) =
        make_double2(cc > rr ? tv.x : (cc == rr ? xd : 0.0), cc + 1 > rr ? tv.y : (cc + 1 == rr ? xd : 0.0));
  }
  // log-determinant terms, off the serial chain: one pivot per thread, fixed-order reduction
  if (threadIdx.x < WV_NB) {
    double lg = (int)threadIdx.x < nreal ? 0.5 * log(sm.piv[threadIdx.x]) : 0.0;
    for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
    if (lane == 0) sm.red[warp] = lg;
  }
  __threadfence();                       // L_jj / Linv_jj / Mt_jj visible device-wide before the flag
  __syncthreads();
  if (threadIdx.x == 0) {
    bd.logdet_part[(size_t)b * bd.nt + j] = sm.red[0] + sm.red[1];
    if (sm.fail) bd.chol_fail[b] = 1;
    __threadfence();
    *reinterpret_cast<volatile int*>(bd.step_flag + (size_t)b * bd.nt + j) = epoch;   // releases the panel CTAs
  }
  WV_CLK(13);
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
      double D[WV_DP_DOUBLES];      // packed triangular inverse block (wv_dp_load)
    } e;
  };
};

// body shared by the panel tiles of the fused Cholesky step (MODE 0, waits for the diagonal CTA's flag between the
// two products) and the triangular-inverse step (MODE 1)
template <int MODE>
__device__ __forceinline__ void wv_panel_body(const WvBatchDev& bd, int b, int step, int tile, int kstart, int epoch) {
  WvPanelSmem& sm = *reinterpret_cast<WvPanelSmem*>(wv_smem_raw);
  const int ld = bd.npad;
  double* Ab = bd.A + (size_t)b * ld * ld;
  double* Mb = bd.Mt + (size_t)b * ld * ld;
  const double *Ag, *Bg;
  double* Out;
  const double* Cin;
  int k0, k1;
  if (MODE == 0) {
    const int i = step + 1 + tile, j = step;
    Ag = Ab + (size_t)i * WV_NB * ld;
    Bg = Ab + (size_t)j * WV_NB * ld;
    k0 = kstart; k1 = j * WV_NB;
    Out = Ab + (size_t)i * WV_NB * ld + j * WV_NB;
    Cin = Out;
  } else {
    const int i = step, j = tile;
    Ag = Mb + (size_t)j * WV_NB * ld;
    Bg = Ab + (size_t)i * WV_NB * ld;
    k0 = j * WV_NB; k1 = i * WV_NB;
    Out = Mb + (size_t)j * WV_NB * ld + i * WV_NB;
    Cin = nullptr;
  }
  // padding: the last tile row (MODE 0: output rows) / last tile column (MODE 1: output columns) of a model holds
  // (n + 1) - 64 (nt - 1) real rows; a warp whose 32 rows / columns are all padding produces zeros that are already in
  // memory (gram rewrites the identity padding, Mt's padding columns are never non-zero)
  const int real_last = bd.n + 1 - (bd.nt - 1) * WV_NB;
  const int warp_ = wv_warp_role();
  bool dead;
  if (MODE == 0) dead = (step + 1 + tile == bd.nt - 1) && real_last <= 32 && (warp_ >> 1) == 1;
  else dead = (step == bd.nt - 1) && real_last <= 32 && (warp_ & 1) == 1;
  double acc[4][4][2];
  wv_zero_acc(acc);
  if (k1 > k0) {
    if (MODE == 1) wv_gemm_nt_64_tria(sm.g, Ag, Bg, ld, k0, k1, acc, dead);     // Mt[j,j] is upper triangular
    else wv_gemm_nt_64(sm.g, Ag, Bg, ld, k0, k1, acc, dead);
  }
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
  if (MODE == 0) {
    // the inverse of the diagonal block is produced by CTA x = 0 of the same launch (dispatched before this one)
    if (threadIdx.x == 0) {
      const volatile int* f = reinterpret_cast<const volatile int*>(bd.step_flag + (size_t)b * bd.nt + step);
      while (*f != epoch) __nanosleep(40);
      __threadfence();
    }
    __syncthreads();
  }
  // D = Linv of the step's diagonal block
  wv_dp_load(sm.e.D, bd.Dinv + ((size_t)b * bd.nt + step) * WV_NB * WV_NB);
  __syncthreads();
  wv_zero_acc(acc);
  wv_gemm_nt_smem64(sm.e.T, sm.e.D, acc, dead);
  if (dead) return;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      int r = r0 + mi * 8, c = c0 + ni * 8;
      *reinterpret_cast<double2*>(Out + (size_t)r * ld + c) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
    }
}

// One column step of the Cholesky factorisation in ONE launch: CTA x = 0 factorises the diagonal block (wv_diag_body),
// CTAs x >= 1 are the panel tiles below it.  A panel tile first accumulates its update sum_k L[i,k] L[j,k]^T, which does
// not depend on the diagonal block, and only then waits for the diagonal CTA's flag: the serial 64x64 factorisation
// overlaps the tensor-pipe work instead of preceding it.  The wait cannot deadlock: CTAs are dispatched in linear
// block order, so the diagonal CTA of a model (x = 0) is resident or finished whenever one of its panel CTAs runs.
// grid (nt - j, n_active), 128 threads.
// With many models in flight the overlap is not worth having waiting tiles hold SM slots: the host then launches the
// diagonal CTAs (grid.x = 1, x0 = 0) and the panel tiles (x0 = 1) separately and the flag is already set.
__global__ void __launch_bounds__(WV_GEMM_THREADS, 3) wv_chol_step_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                                       int j, int k0, int epoch, int x0) {
  const int b = active[blockIdx.y];
  const int x = blockIdx.x + x0;
  if (x == 0) wv_diag_body(bd, b, j, k0, epoch);
  else wv_panel_body<0>(bd, b, j, x - 1, k0, epoch);
}

// the panel tiles alone (many models in flight: the diagonal blocks were factorised by the launch before): without the
// diagonal body the kernel needs ~100 registers and 55 KB, four CTAs per SM instead of three
__global__ void __launch_bounds__(WV_GEMM_THREADS, 4) wv_chol_panel_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                                        int j, int k0, int epoch) {
  wv_panel_body<0>(bd, active[blockIdx.y], j, blockIdx.x, k0, epoch);
}

// ---------------------------------------------------------------------------------------------
// Large-n path: ALL column steps of one panel [p0, p1) in ONE launch.  With one launch per step the serial chain of the
// factorisation pays, per 64 columns, a diagonal block (~30 us), the tail of that step's panel tiles and a launch
// boundary (58 us per step measured at n = 8192: 7.4 of the 8.8 ms Cholesky).  Here the work items of the panel's steps --
// per column j the diagonal block (x = 0) and the tiles (i, j), i > j -- form one grid in column order, and every
// dependency is a flag in global memory:
//   diagonal block j     waits until row j has received the panel's earlier columns   (row_done[j] >= j - p0)
//   tile (i, j)          waits for rows i and j likewise, accumulates, then waits for the diagonal block's flag
//                        (step_flag, inside wv_panel_body<0>) and publishes row_done[i] = j - p0 + 1
// so the chain per column is  diagonal block -> ONE tile (j + 1, j) -> next diagonal block, while the other tiles of a
// column overlap the next diagonal block.  CTAs are dispatched in linear block order and every flag is set by a block
// with a smaller index, so a spinning CTA never waits for one that cannot become resident.  Tiles of a row complete in
// column order (tile (i, j) reads tile (i, j - 1)), hence one counter per row suffices; its value carries a tag of
// (evaluation epoch, panel) so that stale counts of earlier panels / evaluations never match.
// grid (sum_j (nt - j), n_active), 128 threads.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void wv_wait_row(const int* row_done, int tag, int need) {
  if (need <= 0) return;
  const volatile int* f = reinterpret_cast<const volatile int*>(row_done);
  for (;;) {
    const int v = *f;
    if ((v >> 3) == tag && (v & 7) >= need) break;
    __nanosleep(40);
  }
}

__global__ void __launch_bounds__(WV_GEMM_THREADS, 3) wv_chol_panel_fused_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                                              int n_active, int p0, int p1, int items,
                                                                              int epoch, int tag) {
  // PERSISTENT: a few CTAs per SM pull the work items in index order from a counter.  One CTA per item would fill every
  // SM slot with spinning CTAs and starve the trailing update that runs beside this kernel on the side stream
  // (measured: 8.54 ms per n = 8192 Cholesky with one CTA per item).  Items are taken in order, so the lowest unfinished
  // item is always held by a running CTA and its dependencies -- all of lower index -- are finished: no deadlock.
  __shared__ int s_item;
  int* counter = bd.step_flag + 2 * (size_t)bd.B * bd.nt;
  const int total = items * n_active;
  for (;;) {
    __syncthreads();                      // previous item done with the shared-memory tiles
    if (threadIdx.x == 0) s_item = atomicAdd(counter, 1);
    __syncthreads();
    const int it = s_item;
    if (it >= total) return;
    const int b = active[it / items];
    int j = p0, x = it % items;
    while (x >= bd.nt - j) { x -= bd.nt - j; ++j; }          // work item -> (column j, x = 0 diagonal | tile j + x)
    int* row_done = bd.step_flag + (size_t)bd.B * bd.nt + (size_t)b * bd.nt;
    const int need = j - p0;
    if (threadIdx.x == 0) {
      wv_wait_row(row_done + j, tag, need);
      if (x > 0) wv_wait_row(row_done + j + x, tag, need);
      __threadfence();
    }
    __syncthreads();
    if (x == 0) {
      wv_diag_body(bd, b, j, p0 * WV_NB, epoch);
      continue;
    }
    wv_panel_body<0>(bd, b, j, x - 1, p0 * WV_NB, epoch);
    __threadfence();                       // the tile is visible device-wide before the row counter moves
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(row_done + j + x) = (tag << 3) | (need + 1);
  }
}

// ---------------------------------------------------------------------------------------------
// Many models of moderate n (the batched schedule, nt < big_nt): the WHOLE left-looking factorisation in ONE persistent
// launch.  With one launch pair per column every CTA of the device runs the diagonal blocks at the same time -- a serial
// 64x64 factorisation per CTA that leaves the tensor pipe idle (1.8 of the 7.7 ms Cholesky of 2000 models, n = 600) --
// and then every CTA runs panel tiles.  Here the work items of all columns form one ordered list that a few CTAs per SM
// pull from a counter, and the diagonal blocks of column j + 1 are INTERLEAVED with the panel tiles of column j:
//   items [0, B)                              diagonal block (m, 0) of every model m
//   super-step j = 0 .. nt-2, group q = 0 .. B + lag_j - 1, nt - j slots per group:
//       slots 0 .. nt-j-2                     panel tiles (q, row j + 1 + slot, column j)              [if q < B]
//       slot  nt-j-1                          diagonal block (q - lag_j, j + 1)                        [if q >= lag_j]
// so at any time about one CTA in nt - j runs a diagonal block while the others keep the DMMA pipe busy.  lag_j models
// (> the items in flight / slots per group) separate a diagonal block from the tile (j + 1, j) it depends on, so it
// almost never waits.  Dependencies are the flags of the large-n panel kernel: step_flag[b][j] = epoch when diagonal
// block j is done (waited for inside wv_panel_body<0>), row counter row_done[b][i] = (epoch, columns finished in row i).
// Every dependency of an item has a lower index and items are pulled in order: the lowest unfinished item is always held by
// a running CTA whose dependencies are finished -- no deadlock.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void wv_wait_cols(const int* row_done, int ep, int need) {
  if (need <= 0) return;
  const volatile int* f = reinterpret_cast<const volatile int*>(row_done);
  for (;;) {
    const int w = -*f - 1;               // this kernel's counters are stored negated: never mistaken for (or by) the
    if (w >= 0 && (w >> 8) == ep && (w & 255) >= need) break;      // large-n panel kernel's, which share the array
    __nanosleep(40);
  }
}

__global__ void __launch_bounds__(WV_GEMM_THREADS, 3) wv_chol_all_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                                      int n_active, int lag_items, int epoch) {
  __shared__ int s_item;
  int* counter = bd.step_flag + 2 * (size_t)bd.B * bd.nt;
  const int nt = bd.nt;
  const int ep = epoch & 0x3fffff;
  for (;;) {
    __syncthreads();                      // previous item done with the shared-memory tiles
    if (threadIdx.x == 0) s_item = atomicAdd(counter, 1);
    __syncthreads();
    int it = s_item;
    int m, j = 0, x = 0;
    if (it < n_active) {
      m = it;
    } else {
      it -= n_active;
      int lag;
      for (;; ++j) {
        if (j >= nt - 1) return;
        lag = (lag_items + nt - j - 1) / (nt - j);
        const int sz = (n_active + lag) * (nt - j);
        if (it < sz) break;
        it -= sz;
      }
      const int gsz = nt - j;
      const int q = it / gsz, sl = it - q * gsz;
      if (sl < gsz - 1) {
        if (q >= n_active) continue;
        m = q; x = sl + 1;
      } else {
        if (q < lag) continue;
        m = q - lag; j += 1;
      }
    }
    const int b = active[m];
    int* row_done = bd.step_flag + (size_t)bd.B * bd.nt + (size_t)b * bd.nt;
    if (threadIdx.x == 0) {
      wv_wait_cols(row_done + j, ep, j);
      if (x > 0) wv_wait_cols(row_done + j + x, ep, j);
      __threadfence();
    }
    __syncthreads();
    if (x == 0) {
      wv_diag_body(bd, b, j, 0, epoch);
      continue;
    }
    wv_panel_body<0>(bd, b, j, x - 1, 0, epoch);
    __threadfence();                       // the tile is visible device-wide before the row counter moves
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(row_done + j + x) = -((ep << 8) | (j + 1)) - 1;
  }
}

// enqueue one Cholesky column step; fused into one launch iff all its CTAs can be resident at once
static int wv_launch_chol_step(const WvBatchDev& bd, const int* d_active, int n_active, int j, int k0, cudaStream_t st,
                               WvProfiler* pf, const WvAux& aux) {
  const int nt = bd.nt;
  const size_t smem = sizeof(WvPanelSmem) > sizeof(WvDiagSmem) ? sizeof(WvPanelSmem) : sizeof(WvDiagSmem);
  if ((long)n_active * (nt - j) <= aux.resident_ctas || j + 1 == nt) {
    wv_chol_step_kernel<<<dim3(nt - j, n_active), WV_GEMM_THREADS, smem, st>>>(bd, d_active, j, k0, aux.epoch, 0);
    pf->mark(WV_K_CHOL_PANEL, st);
    return 1;
  }
  wv_chol_step_kernel<<<dim3(1, n_active), WV_GEMM_THREADS, smem, st>>>(bd, d_active, j, k0, aux.epoch, 0);
  pf->mark(WV_K_CHOL_DIAG, st);
  wv_chol_panel_kernel<<<dim3(nt - j - 1, n_active), WV_GEMM_THREADS, sizeof(WvPanelSmem), st>>>(bd, d_active, j, k0, aux.epoch);
  pf->mark(WV_K_CHOL_PANEL, st);
  return 2;
}

// trtri(i): tiles (j, i), j < i, of Mt = L^{-T}.  grid (i, n_active), 128 threads.
__global__ void __launch_bounds__(WV_GEMM_THREADS) wv_trtri_kernel(WvBatchDev bd, const int* __restrict__ active, int step) {
  wv_panel_body<1>(bd, active[blockIdx.y], step, blockIdx.x, 0, 0);
}

// The whole triangular inverse in ONE launch: the tile rows of Mt = L^{-T} are independent (row j needs L and its own
// earlier tiles only), so CTA (j, model) walks its row left to right, i = j + 1 .. nt - 1.  The tiles it has just written
// come back from L2 instead of DRAM and the nt - 1 dependent launches (and their tails) become one.  Heavy rows first
// (j = 0 has nt - 1 tiles).  grid (nt - 1, n_active), 128 threads.
__global__ void __launch_bounds__(WV_GEMM_THREADS) wv_trtri_rows_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                                        int paired) {
  const int b = active[blockIdx.y];
  // paired: CTA p takes row p and then row nt - 2 - p (nt - 1 - p and p + 1 tiles: every CTA walks nt tiles) -- the
  // load-balanced form of the one-launch inverse; unpaired: CTA j takes row j alone (grid nt - 1)
  for (int half = 0; half < (paired ? 2 : 1); ++half) {
    const int j = half == 0 ? (int)blockIdx.x : bd.nt - 2 - (int)blockIdx.x;
    if (half == 1 && j <= (int)blockIdx.x) break;
    for (int i = j + 1; i < bd.nt; ++i) {
      wv_panel_body<1>(bd, b, i, j, 0, 0);
      __syncthreads();          // the tile just stored is an operand of the next one (block-scope ordering)
    }
  }
}

// The triangular inverse as ONE persistent launch for a batch that has the device to itself: the tiles (j, i) of all
// steps i = 1 .. nt-1 form one list in step order, pulled from the work counter of wv_chol_all_kernel by a few CTAs per SM.
// Tile (j, i) needs the earlier tiles of its own row (j, j+1 .. i-1) -- a per-row counter, kept in the row counters of the
// Cholesky kernel with an offset of 128 so that a leftover Cholesky count never satisfies a wait -- and L / Linv, which are
// final.  Dependencies have lower indices and items are pulled in order: no deadlock.  Unlike the row-per-CTA form
// (wv_trtri_rows_kernel) the load is balanced, and the nine launch boundaries with their tails are gone.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WV_GEMM_THREADS) wv_trtri_all_kernel(WvBatchDev bd, const int* __restrict__ active,
                                                                     int n_active, int epoch) {
  __shared__ int s_item;
  int* counter = bd.step_flag + 2 * (size_t)bd.B * bd.nt;
  const int nt = bd.nt;
  const int ep = epoch & 0x3fffff;
  for (;;) {
    __syncthreads();                      // previous item done with the shared-memory tiles
    if (threadIdx.x == 0) s_item = atomicAdd(counter, 1);
    __syncthreads();
    int it = s_item, i = 1;
    for (;; ++i) {
      if (i >= nt) return;
      const int sz = n_active * i;
      if (it < sz) break;
      it -= sz;
    }
    const int m = it / i, j = it - m * i;
    const int b = active[m];
    int* row_done = bd.step_flag + (size_t)bd.B * bd.nt + (size_t)b * bd.nt;
    if (threadIdx.x == 0) {
      if (i - 1 > j) wv_wait_cols(row_done + j, ep, 128 + (i - 1 - j));
      __threadfence();
    }
    __syncthreads();
    wv_panel_body<1>(bd, b, i, j, 0, 0);
    __threadfence();                       // the tile is visible device-wide before the row counter moves
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(row_done + j) = -((ep << 8) | (128 + i - j)) - 1;
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
  if (ti == tj) {
    // diagonal tile: Mt_i Mt_i^T is symmetric -- one staged operand, only the lower 8x8 blocks are formed (36 of 64);
    // the strict upper blocks of the tile keep stale values, every reader mirrors or masks them
    double sacc[10][2];
#pragma unroll
    for (int q = 0; q < 10; ++q) sacc[q][0] = sacc[q][1] = 0.0;
    wv_syrk_self_64(sm, Mb + (size_t)ti * WV_NB * ld, ld, ti * WV_NB, bd.n8, sacc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    double* Out = bd.A + (size_t)b * ld * ld + (size_t)ti * WV_NB * ld + ti * WV_NB;
#define WV_X(q, i, j) \
    *reinterpret_cast<double2*>(Out + (size_t)((i) * 8 + fr) * ld + (j) * 8 + 2 * fk) = make_double2(sacc[q][0], sacc[q][1]);
    if (warp == 0) { WV_SYM_W0(WV_X) }
    else if (warp == 1) { WV_SYM_W1(WV_X) }
    else if (warp == 2) { WV_SYM_W2(WV_X) }
    else { WV_SYM_W3(WV_X) }
#undef WV_X
    return;
  }
  double acc[4][4][2];
  wv_zero_acc(acc);
  // rows / columns beyond n of K^-1 are never read (the gradient pass masks them): padding warps of the last tile
  // row / column skip their products
  const int real_last = bd.n + 1 - (bd.nt - 1) * WV_NB;
  const int warp_ = wv_warp_role();
  const bool dead = real_last <= 32 && ((ti == bd.nt - 1 && (warp_ >> 1) == 1) || (tj == bd.nt - 1 && (warp_ & 1) == 1));
  // the first k-tile of the A operand is Mt[ti,ti], upper triangular
  wv_gemm_nt_64_tria(sm, Mb + (size_t)ti * WV_NB * ld, Mb + (size_t)tj * WV_NB * ld, ld, ti * WV_NB, bd.n8, acc, dead);
  if (dead) return;
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
// kinv with a TMA operand ring (experiment, WV_KINV_TMA=1; result in DESIGN.md section 4): the two 64 x 16 fp64 operand
// boxes of a k-chunk are fetched by ONE thread with cp.async.bulk.tensor.2d (tensor map over Mt as a [B npad, npad]
// matrix, 128-byte swizzle) and land on an mbarrier; the other 127 threads issue no load instructions at all.
// With 128-byte rows the DMMA fragment loads (8 rows x 4 k per operand) are bank-conflict free only if the 8 rows of a
// fragment differ in (row & 7) in a way the XOR swizzle separates: fragment row fr of block mi is matrix row
//     R(mi, fr) = 32 wm + 16 (mi >> 1) + 2 fr + (mi & 1)
// (rows of stride 2: (R & 7) runs over {0,2,4,6} or {1,3,5,7} within a half-warp, so the two 16-byte chunks of the 4
// k-values land in 8 distinct chunk slots).  Columns likewise; a thread then owns 4 adjacent output columns per (mi, ni
// pair) and stores them as two double2.
// =============================================================================================
#ifndef WV_TMA_STAGES
#define WV_TMA_STAGES 3
#endif
struct WvTmaSmem {
  double a[WV_TMA_STAGES][WV_NB * WV_BK];      // 8 KB per box, 1024-byte aligned (swizzle atom)
  double b[WV_TMA_STAGES][WV_NB * WV_BK];
  unsigned long long bar[WV_TMA_STAGES];
};

__device__ __forceinline__ void wv_mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void wv_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void wv_mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void wv_tma_load_2d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// byte offset of element (row r, k) of a 64 x 16 fp64 box written with CU_TENSOR_MAP_SWIZZLE_128B
__device__ __forceinline__ int wv_swz(int r, int k) { return r * 128 + ((((k >> 1) ^ (r & 7)) << 4) | ((k & 1) << 3)); }

__global__ void __launch_bounds__(WV_GEMM_THREADS) wv_kinv_tma_kernel(const __grid_constant__ CUtensorMap tmap, WvBatchDev bd,
                                                                      const int* __restrict__ active) {
  const int b = active[blockIdx.y];
  const int ld = bd.npad;
  int ti, tj;
  wv_tile_from_linear(blockIdx.x, ti, tj);
  const double* Mb = bd.Mt + (size_t)b * ld * ld;
  if (ti == tj) {      // symmetric diagonal tile: the single-operand cp.async path of wv_kinv_kernel
    WvGemmSmem& sm = *reinterpret_cast<WvGemmSmem*>(wv_smem_raw);
    double sacc[10][2];
#pragma unroll
    for (int q = 0; q < 10; ++q) sacc[q][0] = sacc[q][1] = 0.0;
    wv_syrk_self_64(sm, Mb + (size_t)ti * WV_NB * ld, ld, ti * WV_NB, bd.n8, sacc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    double* Out = bd.A + (size_t)b * ld * ld + (size_t)ti * WV_NB * ld + ti * WV_NB;
#define WV_X(q, i, j) \
    *reinterpret_cast<double2*>(Out + (size_t)((i) * 8 + fr) * ld + (j) * 8 + 2 * fk) = make_double2(sacc[q][0], sacc[q][1]);
    if (warp == 0) { WV_SYM_W0(WV_X) }
    else if (warp == 1) { WV_SYM_W1(WV_X) }
    else if (warp == 2) { WV_SYM_W2(WV_X) }
    else { WV_SYM_W3(WV_X) }
#undef WV_X
    return;
  }
  WvTmaSmem& sm = *reinterpret_cast<WvTmaSmem*>((reinterpret_cast<uintptr_t>(wv_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int lane = threadIdx.x & 31, warp = wv_warp_role();
  const int wm = warp >> 1, wn = warp & 1, fr = lane >> 2, fk = lane & 3;
  const int real_last = bd.n + 1 - (bd.nt - 1) * WV_NB;
  const bool dead = real_last <= 32 && ((ti == bd.nt - 1 && wm == 1) || (tj == bd.nt - 1 && wn == 1));
  const int k0 = ti * WV_NB, nch = (bd.n8 - k0 + WV_BK - 1) / WV_BK;
  const int row_a = b * ld + ti * WV_NB, row_b = b * ld + tj * WV_NB;
  if (threadIdx.x == 0) {
    for (int s_ = 0; s_ < WV_TMA_STAGES; ++s_) wv_mbar_init(&sm.bar[s_], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int s_ = 0; s_ < WV_TMA_STAGES && s_ < nch; ++s_) {
      wv_mbar_expect_tx(&sm.bar[s_], 2 * WV_NB * WV_BK * 8);
      wv_tma_load_2d(sm.a[s_], &tmap, &sm.bar[s_], k0 + s_ * WV_BK, row_a);
      wv_tma_load_2d(sm.b[s_], &tmap, &sm.bar[s_], k0 + s_ * WV_BK, row_b);
    }
  double acc[4][4][2];
  wv_zero_acc(acc);
  for (int c = 0; c < nch; ++c) {
    const int st = c % WV_TMA_STAGES;
    wv_mbar_wait(&sm.bar[st], (c / WV_TMA_STAGES) & 1);
    if (!dead) {
      const char* as = reinterpret_cast<const char*>(sm.a[st]);
      const char* bs = reinterpret_cast<const char*>(sm.b[st]);
#pragma unroll
      for (int kk = 0; kk < WV_BK; kk += 4) {
        // Mt[ti,ti] (chunks 0..3) is upper triangular: rows >= r are zero for k < r; 16-row granularity here
        const int kabs = c * WV_BK + kk;
        const bool lo = !(c < 4 && kabs + 4 <= wm * 32), hi = !(c < 4 && kabs + 4 <= wm * 32 + 16);
        if (!hi && !lo) continue;
        double bf[4];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
          bf[ni] = *reinterpret_cast<const double*>(bs + wv_swz(wn * 32 + (ni >> 1) * 16 + 2 * fr + (ni & 1), kk + fk));
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          if (mi < 2 ? !lo : !hi) continue;
          const double af = *reinterpret_cast<const double*>(as + wv_swz(wm * 32 + (mi >> 1) * 16 + 2 * fr + (mi & 1), kk + fk));
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) wv_dmma(acc[mi][ni][0], acc[mi][ni][1], af, bf[ni]);
        }
      }
    }
    __syncthreads();                                        // every warp is done with stage st
    if (threadIdx.x == 0 && c + WV_TMA_STAGES < nch) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy reads before the async-proxy refill
      wv_mbar_expect_tx(&sm.bar[st], 2 * WV_NB * WV_BK * 8);
      wv_tma_load_2d(sm.a[st], &tmap, &sm.bar[st], k0 + (c + WV_TMA_STAGES) * WV_BK, row_a);
      wv_tma_load_2d(sm.b[st], &tmap, &sm.bar[st], k0 + (c + WV_TMA_STAGES) * WV_BK, row_b);
    }
  }
  if (dead) return;
  double* Out = bd.A + (size_t)b * ld * ld + (size_t)ti * WV_NB * ld + tj * WV_NB;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
    const int r = wm * 32 + (mi >> 1) * 16 + 2 * fr + (mi & 1);
#pragma unroll
    for (int np = 0; np < 2; ++np) {       // ni pair (2 np, 2 np + 1): columns 16 np + 4 fk .. + 3
      double* o = Out + (size_t)r * ld + wn * 32 + np * 16 + 4 * fk;
      *reinterpret_cast<double2*>(o) = make_double2(acc[mi][2 * np][0], acc[mi][2 * np + 1][0]);
      *reinterpret_cast<double2*>(o + 2) = make_double2(acc[mi][2 * np][1], acc[mi][2 * np + 1][1]);
    }
  }
}

// tensor map of Mt as a row-major [B npad, npad] fp64 matrix, box 16 (k) x 64 (rows), 128-byte swizzle
int wv_make_tmap_mt(const WvBatchDev& bd, void* out128) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)bd.npad, (cuuint64_t)bd.B * bd.npad};
  const cuuint64_t strides[1] = {(cuuint64_t)bd.npad * sizeof(double)};
  const cuuint32_t box[2] = {WV_BK, WV_NB};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(reinterpret_cast<CUtensorMap*>(out128), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)bd.Mt, dims, strides,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1;
}

// =============================================================================================
// Large-n path (nt >= WvAux::big_nt, e.g. config 4: one n = 8192 model).  The left-looking batched scheme above has
// one CTA per model on its critical path; for few, large models the factorisation is reorganised so that almost all
// flops sit in wide launches of the same 64x64 DMMA tile GEMM:
//   Cholesky   right-looking over panels of WvAux::panel_tiles tile columns: the panel is factorised with the kernels
//              above (k0 = first panel column), then `wv_syrk_kernel` applies it to the trailing matrix.  The update
//              is split into the columns of the NEXT panel (main stream) and the rest (side stream), so that the next
//              panel factorisation overlaps the bulk of the update (look-ahead of one panel).
//   L^{-T}     recursive doubling: at level m the inverse of every aligned 2m-tile diagonal block is assembled from
//              its two m-tile halves, Mt_FS = -Mt_FF (L_SF^T Mt_SS), as two launches over all blocks of the level
//              (`wv_trtri_level_kernel<1>`: U = Mt_FF L_SF^T into the unused upper tiles of A;  <2>: Mt_FS = -U Mt_SS,
//              the only product on the path whose B operand is not k-contiguous -> wv_gemm_64<true>).
// =============================================================================================

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
      if ((bd.lik == 2 || bd.lik == 4 || bd.lik == 5) && s == pg->noise_slot) dl = bd.vgp_dlik[b];   // the slot is the likelihood parameter
      if (bd.lik == 5 && s == pg->lik_slot2) dl = bd.vgp_dlik2[b];
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
    double lml = -0.5 * bd.quad[b] - 0.5 * bd.n * 1.8378770664093453 - logdet;
    if (bd.vgp_extra) lml += bd.vgp_extra[b];     // variational path: ELBO at the converged sites
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
static size_t wv_smem_gemm_bytes() { return sizeof(WvPanelSmem) > sizeof(WvDiagSmem) ? sizeof(WvPanelSmem) : sizeof(WvDiagSmem); }

// dynamic shared-memory limits are a per-device function attribute; setting them twice (two host threads racing on the
// first evaluation) is harmless
static bool g_attr_done[64] = {};
static cudaError_t wv_set_attrs() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && g_attr_done[dev]) return cudaSuccess;
#define WV_ATTR(k, bytes) \
  e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
  if (e != cudaSuccess) return e;
  WV_ATTR(wv_gram_kernel, sizeof(WvElemSmem));
  WV_ATTR(wv_grad_kernel, sizeof(WvElemSmem));
  WV_ATTR(wv_cross_mean_kernel, sizeof(WvElemSmem));
  WV_ATTR(wv_cross_var_kernel, sizeof(WvCrossVarSmem));
  WV_ATTR(wv_chol_step_kernel, wv_smem_gemm_bytes());
  WV_ATTR(wv_chol_panel_kernel, sizeof(WvPanelSmem));
  WV_ATTR(wv_chol_panel_fused_kernel, wv_smem_gemm_bytes());
  WV_ATTR(wv_chol_all_kernel, wv_smem_gemm_bytes());
  WV_ATTR(wv_trtri_kernel, sizeof(WvPanelSmem));
  WV_ATTR(wv_trtri_rows_kernel, sizeof(WvPanelSmem));
  WV_ATTR(wv_trtri_all_kernel, sizeof(WvPanelSmem));
  WV_ATTR(wv_kinv_kernel, sizeof(WvGemmSmem));
  WV_ATTR(wv_kinv_tma_kernel, sizeof(WvTmaSmem) + 1024);
  WV_ATTR(wv_syrk_kernel, sizeof(WvGemmSmem));
  WV_ATTR(wv_trtri_level_kernel<1>, sizeof(WvGemmSmem));
  WV_ATTR(wv_trtri_level_kernel<2>, sizeof(WvGemmSmem));
#undef WV_ATTR
  if (dev >= 0 && dev < 64) g_attr_done[dev] = true;
  return cudaSuccess;
}

// Cholesky + L^{-T} of the large-n path (see the comment above wv_syrk_kernel).  Returns launches or -1.
static int wv_enqueue_factor_big(const WvBatchDev& bd, const int* d_active, int n_active, cudaStream_t st,
                                 WvProfiler* pf, const WvAux& aux, bool chol_only = false) {
  const int nt = bd.nt;
  int launches = 0;
  bool bulk_pending = false;
  const int PT = aux.panel_tiles;
  for (int p0 = 0; p0 < nt; p0 += PT) {
    const int p1 = p0 + PT < nt ? p0 + PT : nt;
    if (aux.panel_fused && p1 - p0 <= 6) {
      int items = 0;
      for (int j = p0; j < p1; ++j) items += nt - j;
      const int tag = (aux.epoch * 256 + p0 / PT) & 0x0fffffff;
      const long total = (long)items * n_active;
      const int ctas = (int)(total < aux.panel_ctas ? total : aux.panel_ctas);
      cudaMemsetAsync(bd.step_flag + 2 * (size_t)bd.B * bd.nt, 0, sizeof(int), st);      // the work counter
      wv_chol_panel_fused_kernel<<<ctas, WV_GEMM_THREADS, wv_smem_gemm_bytes(), st>>>(bd, d_active, n_active, p0, p1, items,
                                                                                   aux.epoch, tag);
      pf->mark(WV_K_CHOL_PANEL, st);
      ++launches;
    } else {
      for (int j = p0; j < p1; ++j) launches += wv_launch_chol_step(bd, d_active, n_active, j, p0 * WV_NB, st, pf, aux);
    }
    if (p1 >= nt) break;
    const int a_hi = p1 + PT < nt ? p1 + PT : nt;   // columns of the next panel
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
  if (chol_only) return launches;
  for (int m = 1; m < nt; m *= 2) {
    const int nblk = (nt + 2 * m - 1) / (2 * m);
    wv_trtri_level_kernel<1><<<dim3(nblk * m * m, n_active), WV_GEMM_THREADS, sizeof(WvGemmSmem), st>>>(bd, d_active, m);
    wv_trtri_level_kernel<2><<<dim3(nblk * m * m, n_active), WV_GEMM_THREADS, sizeof(WvGemmSmem), st>>>(bd, d_active, m);
    pf->mark(WV_K_TRTRI, st);
    launches += 2;
  }
  return launches;
}

// =============================================================================================
// Variational path for count likelihoods (BASELINE configs[4]; reference: gpflow VGP / PSVGP with Z = X,
// waveome/model_fitting.py:158-185, waveome/likelihoods.py:16-79).  For a factorising likelihood the optimal Gaussian q
// of the VGP bound is the posterior of a GP with Gaussian pseudo-observations ("sites": precision lam_i, precision x
// mean eta_i), so maximising the bound over q for fixed hyper-parameters is a fixed point of heteroscedastic GPR
// solves -- each sweep is ONE run of the factorisation above with noise jitter + 1/lam_i and data eta_i/lam_i:
//     m = ytilde - D alpha,   v = D - D^2 diag((K + D)^-1),   D = 1/lam
//     targets:  lam_i <- -2 dE_i/dv_i,   eta_i <- dE_i/dm_i + lam_i m_i          (E_i = E_q log p(y_i | f_i))
//     F = log N(ytilde; c, K + D) + sum_i [E_i + 1/2 log(2 pi / lam_i) + lam_i/2 ((ytilde_i - m_i)^2 + v_i)]
// F is the exact ELBO of the q defined by the current sites (a lower bound at every sweep), which makes the step
// safe: a damped move (1 - rho) sites + rho targets is accepted only if F did not decrease, otherwise rho is halved
// and the move is retried from the last accepted sites.  At the fixed point dF/dtheta = 1/2 tr((alpha alpha^T -
// (K + D)^-1) dK/dtheta) and dF/dc = sum(alpha) (envelope theorem): the Gaussian-path gradient kernel, unchanged.
// grid (n_list), 256 threads.
// =============================================================================================
// digamma: recurrence up to x >= 10, then the asymptotic series (|error| < 1e-15 there)
__device__ __forceinline__ double wv_digamma(double x) {
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  return r + log(x) - 0.5 / x -
         f * (1.0 / 12.0 - f * (1.0 / 120.0 - f * (1.0 / 252.0 - f * (1.0 / 240.0 - f * (1.0 / 132.0)))));
}

// smallest site precision: 1e-300 ("never") for the log-concave likelihoods; the ZINB zero branch is not log-concave, its
// bound is maximised over the Gaussian family with site precisions >= 1e-6 (oracle/vgp_oracle.py LAM_MIN)
__device__ __forceinline__ double wv_lam_min(int lik) { return lik == 5 ? 1e-6 : 1e-300; }

__device__ __forceinline__ void wv_var_exp(int lik, double alpha_nb, double km, double y, double lgam, double m, double v,
                                           double& E, double& g, double& h, double& da, double& da2) {
  da = 0.0; da2 = 0.0;
  if (lik == 1) {                       // Poisson, exp link: closed form (gpflow.likelihoods.Poisson)
    const double r = exp(m + 0.5 * v);
    E = y * m - r - lgam;
    g = y - r;
    h = -0.5 * r;
    return;
  }
  if (lik == 4) {                       // Gamma, exp link, shape a = alpha_nb: closed form (gpflow.likelihoods.Gamma)
    const double a = alpha_nb;
    const double r = y * exp(-m + 0.5 * v);
    E = -a * m - lgamma(a) + (a - 1.0) * log(y) - r;
    g = -a + r;
    h = -0.5 * r;
    da = -m - wv_digamma(a) + log(y);
    return;
  }
  // negative binomial, log link (waveome/likelihoods.py:68-79), 20-point Gauss-Hermite; derivatives by Bonnet / Price
  const double gx[10] = {0.2453407083009012, 0.7374737285453944, 1.2340762153953231, 1.7385377121165861,
                         2.2549740020892757, 2.7888060584281305, 3.3478545673832163, 3.9447640401156252,
                         4.6036824495507442, 5.3874808900112328};
  const double gw[10] = {4.622436696006101e-01, 2.866755053628341e-01, 1.090172060200233e-01, 2.481052088746361e-02,
                         3.243773342237862e-03, 2.283386360163540e-04, 7.802556478532064e-06, 1.086069370769282e-07,
                         4.399340992273181e-10, 2.229393645534151e-13};
  const double sd = sqrt(2.0 * v);
  if (lik == 3) {                       // Bernoulli, gpflow inv_probit link p = 1e-3 + (1 - 2e-3) Phi(f), y in {0, 1}
    double se = 0.0, s1 = 0.0, s2 = 0.0;
    for (int q = 0; q < 20; ++q) {
      const double x = q < 10 ? -gx[9 - q] : gx[q - 10];
      const double w = (q < 10 ? gw[9 - q] : gw[q - 10]) * 0.5641895835477563;
      const double f = m + sd * x;
      const double ph = 0.3989422804014327 * exp(-0.5 * f * f) * (1.0 - 2e-3);      // dp/df
      const double p = 1e-3 + (1.0 - 2e-3) * 0.5 * erfc(-f * 0.7071067811865476);
      const double pp = y > 0.5 ? p : 1.0 - p;           // probability of the observed class
      const double d1 = (y > 0.5 ? ph : -ph) / pp;       // d log pp / df
      const double d2 = -f * d1 - d1 * d1;               // d2 log pp / df2  (dph/df = -f ph)
      se += w * log(pp);
      s1 += w * d1;
      s2 += w * d2;
    }
    E = se; g = s1; h = 0.5 * s2;
    return;
  }
  const double k = 1.0 / alpha_nb;
  if (lik == 5 && y == 0.0) {
    // ZINB (waveome/likelihoods.py:96-139), y = 0: log(psi + (1 - psi) NB(0)) = log(km + u) - log(km + e^f),
    // u = e^f (1 + alpha e^f)^(-1/alpha), psi = km / (km + e^f)
    double se = 0.0, s1 = 0.0, s2 = 0.0, sa = 0.0, sm = 0.0;
    for (int q = 0; q < 20; ++q) {
      const double x = q < 10 ? -gx[9 - q] : gx[q - 10];
      const double w = (q < 10 ? gw[9 - q] : gw[q - 10]) * 0.5641895835477563;
      const double f = m + sd * x;
      const double ef = exp(f);
      const double am = 1.0 + alpha_nb * ef, l1 = log1p(alpha_nb * ef);
      const double u = ef * exp(-l1 * k);
      const double P = km + u, Q = km + ef;
      const double r = (1.0 + (alpha_nb - 1.0) * ef) / am;        // d log u / df
      const double u1 = u * r, u2 = u * (r * r - ef / (am * am));
      se += w * (log(P) - log(Q));
      s1 += w * (u1 / P - ef / Q);
      s2 += w * (u2 / P - (u1 / P) * (u1 / P) - km * ef / (Q * Q));
      sa += w * (u * (l1 * k * k - ef * k / am) / P);
      sm += w * (1.0 / P - 1.0 / Q);
    }
    E = se; g = s1; h = 0.5 * s2; da = sa; da2 = sm;
    return;
  }
  const double cst = lgamma(k + y) - lgam - lgamma(k);
  double se = 0.0, s1 = 0.0, s2 = 0.0, sk = 0.0, sm = 0.0;
  const double dcst = wv_digamma(k + y) - wv_digamma(k);          // d cst / dk
  for (int q = 0; q < 20; ++q) {
    const double x = q < 10 ? -gx[9 - q] : gx[q - 10];
    const double w = (q < 10 ? gw[9 - q] : gw[q - 10]) * 0.5641895835477563;      // / sqrt(pi)
    const double f = m + sd * x;
    const double ef = exp(f);
    if (lik == 5) {        // ZINB, y > 0: log(1 - psi) + NB = f - log(km + e^f) + NB
      const double Q = km + ef;
      se += w * (f - log(Q));
      s1 += w * (km / Q);
      s2 += w * (-km * ef / (Q * Q));
      sm += w * (-1.0 / Q);
    }
    const double lp = cst + y * (f - log(ef + k)) - k * log1p(ef * alpha_nb);
    const double t = alpha_nb * ef / (1.0 + alpha_nb * ef);
    se += w * lp;
    s1 += w * (y - (y + k) * t);
    s2 += w * (-(y + k) * t / (1.0 + alpha_nb * ef));
    // d log p / dk = psi(k+y) - psi(k) - y / (e^f + k) - log(1 + e^f / k) + e^f / (k + e^f)
    sk += w * (dcst - y / (ef + k) - log1p(ef * alpha_nb) + ef / (k + ef));
  }
  E = se; g = s1; h = 0.5 * s2;
  da = -k * k * sk;                     // dk / dalpha = -1 / alpha^2
  da2 = sm;
}

__global__ void __launch_bounds__(256) wv_site_update_kernel(WvBatchDev bd, WvVgpState vs, const int* __restrict__ list,
                                                             const double* __restrict__ xall) {
  __shared__ double red[8];
  __shared__ double s_bcast[2];
  __shared__ double s_dal, s_dal2;
  __shared__ int s_dec, s_nbound;
  const int b = list[blockIdx.x];
  const int n = bd.n, ld = bd.npad;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const WvProgram* pg = bd.programs + bd.prog_id[b];
  double cmean = 0.0;
  if (pg->mean_slot >= 0) {
    const WvSlot& sl = pg->slots[pg->mean_slot];
    cmean = sl.xindex >= 0 ? wv_transform(sl.transform, xall[(size_t)b * bd.P + sl.xindex], sl.shift) : sl.fixed;
  }
  double lik_param = bd.lik_param, lik_param2 = bd.lik_param2;
  if (bd.lik == 2 || bd.lik == 4 || bd.lik == 5) {   // a trainable noise slot is the likelihood parameter (NB / ZINB alpha, Gamma shape)
    const WvSlot& sl = pg->slots[pg->noise_slot];
    if (sl.xindex >= 0) lik_param = wv_transform(sl.transform, xall[(size_t)b * bd.P + sl.xindex], sl.shift);
  }
  if (bd.lik == 5 && pg->lik_slot2 >= 0) {           // ZINB km
    const WvSlot& sl = pg->slots[pg->lik_slot2];
    lik_param2 = sl.xindex >= 0 ? wv_transform(sl.transform, xall[(size_t)b * bd.P + sl.xindex], sl.shift) : sl.fixed;
  }
  const double lam_min = wv_lam_min(bd.lik);
  double* lam = bd.site_lam + (size_t)b * ld;
  double* eta = bd.site_eta + (size_t)b * ld;
  double* lam_p = vs.lam_p + (size_t)b * ld;
  double* eta_p = vs.eta_p + (size_t)b * ld;
  double* lam_t = vs.lam_t + (size_t)b * ld;
  double* eta_t = vs.eta_t + (size_t)b * ld;
  const double* al = bd.alpha + (size_t)b * ld;
  const double* Ab = bd.A + (size_t)b * ld * ld;
  const double* yb = bd.Y + (size_t)b * ld;
  const double* lg = vs.lgam + (size_t)b * ld;
  // ---- pass 1: posterior marginals, variational expectations, the bound, the targets of the next move
  double part = 0.0, dmax = 0.0, dalpha = 0.0, dalpha2 = 0.0;
  int nbound = 0;
  if (threadIdx.x == 0) s_nbound = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double l = lam[i], e = eta[i];
    const double D = 1.0 / l, yt = e * D;
    const double m = yt - D * al[i];
    const double v = D - D * D * Ab[(size_t)i * ld + i];
    double E, g, h, da, da2;
    wv_var_exp(bd.lik, lik_param, lik_param2, yb[i], lg[i], m, v, E, g, h, da, da2);
    dalpha += da;
    dalpha2 += da2;
    part += E + 0.5 * log(6.283185307179586 / l) + 0.5 * l * ((yt - m) * (yt - m) + v);
    const double lt = fmax(-2.0 * h, lam_min);
    if (bd.lik == 5 && -2.0 * h < lam_min) nbound = 1;
    const double et = g + lt * m;
    vs.fmean[(size_t)b * ld + i] = m;
    vs.fvar[(size_t)b * ld + i] = v;
    // kept in registers would be cheaper, but a rejected sweep must not overwrite the accepted targets: stage in fmean
    // / fvar's neighbours below only after the decision
    const double d1 = fabs(lt - l) / (fabs(l) + 1e-300), d2 = fabs(et - e) / (1.0 + fabs(e));
    dmax = fmax(dmax, fmax(d1, d2));
    if (!(d1 == d1) || !(d2 == d2)) dmax = INFINITY;
  }
  if (nbound) atomicOr(&s_nbound, 1);
  // deterministic block reductions (sum, max)
  for (int o = 16; o > 0; o >>= 1) {
    part += __shfl_xor_sync(0xffffffffu, part, o);
    dalpha += __shfl_xor_sync(0xffffffffu, dalpha, o);
    dalpha2 += __shfl_xor_sync(0xffffffffu, dalpha2, o);
    dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  }
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double ssum = 0.0;
    for (int w = 0; w < 8; ++w) ssum += red[w];
    s_bcast[0] = ssum;
  }
  __syncthreads();
  if (lane == 0) red[warp] = dalpha;
  __syncthreads();
  if (threadIdx.x == 0) {
    double ssum = 0.0;
    for (int w = 0; w < 8; ++w) ssum += red[w];
    s_dal = ssum;
  }
  __syncthreads();
  if (lane == 0) red[warp] = dalpha2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double ssum = 0.0;
    for (int w = 0; w < 8; ++w) ssum += red[w];
    s_dal2 = ssum;
  }
  __syncthreads();
  if (lane == 0) red[warp] = dmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    double mx = 0.0;
    for (int w = 0; w < 8; ++w) mx = fmax(mx, red[w]);
    double logdet = 0.0;
    for (int jb = 0; jb < bd.nt; ++jb) logdet += bd.logdet_part[(size_t)b * bd.nt + jb];
    const double extra = s_bcast[0];
    const double F = -0.5 * bd.quad[b] - 0.5 * n * 1.8378770664093453 - logdet + extra;
    const bool finite = (F - F == 0.0) && !bd.chol_fail[b];
    const int first = vs.first[b];
    // 0 converged, 1 accepted (move on), 2 rejected (retry with half the step), 3 give up, 4 sweep cap on an accepted state
    int dec;
    double rho = vs.rho[b];
    int good = vs.good[b];
    const int sweeps = vs.sweeps[b] + 1;
    if (!finite || (!first && !(F >= vs.F_prev[b] - 1e-13 * (1.0 + fabs(vs.F_prev[b]))))) {
      if (first || rho < 1e-6) dec = 3;
      else { dec = 2; rho *= 0.5; good = 0; }
    } else if (mx < vs.tol) {
      dec = 0;
    } else {
      dec = 1;
      vs.F_prev[b] = F;
      if (!first && ++good >= 3) { rho = fmin(1.0, 1.5 * rho); good = 0; }     // grow the step only after a calm stretch
    }
    int flag = (dec == 1 || dec == 2) ? 1 : (dec == 0 ? 0 : -1);
    if (sweeps >= vs.max_sweeps && dec == 1) { dec = 4; flag = mx < vs.soft_tol ? 0 : -1; }
    if (sweeps >= vs.max_sweeps && dec == 2) { dec = 3; flag = -1; }
    if (dec == 0 || dec == 4) { bd.vgp_extra[b] = extra; bd.vgp_dlik[b] = s_dal; bd.vgp_dlik2[b] = s_dal2; vs.at_bound[b] = s_nbound; }
    vs.rho[b] = rho;
    vs.good[b] = good;
    vs.first[b] = 0;
    vs.sweeps[b] = sweeps;
    vs.inner_task[b] = flag;
    s_dec = dec;
    s_bcast[1] = rho;
  }
  __syncthreads();
  const int dec = s_dec;
  const double rho = s_bcast[1];
  // ---- pass 2: move the sites
  if (dec == 1) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double l = lam[i], e = eta[i];
      const double D = 1.0 / l, yt = e * D;
      const double m = vs.fmean[(size_t)b * ld + i], v = vs.fvar[(size_t)b * ld + i];
      double E, g, h, da, da2;
      wv_var_exp(bd.lik, lik_param, lik_param2, yb[i], lg[i], m, v, E, g, h, da, da2);
      const double lt = fmax(-2.0 * h, lam_min), et = g + lt * m;
      lam_p[i] = l; eta_p[i] = e; lam_t[i] = lt; eta_t[i] = et;
      lam[i] = (1.0 - rho) * l + rho * lt;
      eta[i] = (1.0 - rho) * e + rho * et;
      (void)yt;
    }
  } else if (dec == 2) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      lam[i] = (1.0 - rho) * lam_p[i] + rho * lam_t[i];
      eta[i] = (1.0 - rho) * eta_p[i] + rho * eta_t[i];
    }
  } else if (dec == 3) {
    // no acceptable move: fall back to the last accepted sites if there are any (the factorisation in memory then
    // belongs to other sites; the caller flags the evaluation as non-finite)
    if (!vs.first[b] && vs.sweeps[b] > 1)
      for (int i = threadIdx.x; i < n; i += blockDim.x) { lam[i] = lam_p[i]; eta[i] = eta_p[i]; }
  }
}

// ordered compaction of list entries whose flag is 1 (single CTA)
__global__ void wv_compact_list_kernel(const int* __restrict__ list, int n_list, const int* __restrict__ flag, int* out,
                                       int* count) {
  __shared__ int warp_tot[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int start = 0; start < n_list; start += blockDim.x) {
    const int i = start + threadIdx.x;
    const int b = i < n_list ? list[i] : -1;
    const bool act = b >= 0 && flag[b] == 1;
    const unsigned bal = __ballot_sync(0xffffffffu, act);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (act) out[off + __popc(bal & ((1u << lane) - 1))] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < nw; ++w) t += warp_tot[w];
      base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = base;
}

__global__ void wv_vgp_begin_kernel(WvVgpState vs, const int* __restrict__ list, int n_list) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_list) return;
  const int b = list[i];
  vs.first[b] = 1; vs.rho[b] = 1.0; vs.sweeps[b] = 0; vs.good[b] = 0; vs.inner_task[b] = 1;
}

// status bits of the site iteration, OR-ed into the evaluation status after finalize
__global__ void wv_vgp_status_kernel(WvVgpState vs, const int* __restrict__ list, int n_list, int* __restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_list) return;
  const int b = list[i];
  if (vs.inner_task[b] == -1) status[b] |= (vs.sweeps[b] >= vs.max_sweeps ? WV_STATUS_INNER_CAP : WV_STATUS_NONFINITE);
  if (vs.at_bound[b]) status[b] |= WV_STATUS_SITE_BOUND;
}

// =============================================================================================
// Objective (B) at GIVEN variational parameters: the whitened VGP / SVGP-with-Z = X bound the live reference API hands to
// its optimisers (waveome/model_classes.py:1082-1126 PSVGP -> gpflow.models.SVGP.elbo; mirror
// waveome/model_types_DEPR.py:126-158; VGP: waveome/model_fitting.py:158-185):
//     L L^T = K + jitter I,  f_mean = c + L q_mu,  f_var_i = |(L tril(q_sqrt))_i|^2
//     ELBO = sum_i E_{N(f_mean_i, f_var_i)}[log p(y_i | f_i)] - 1/2 (|q_mu|^2 + |tril(q_sqrt)|_F^2 - n - 2 sum log|q_sqrt_ii|)
// L is read from the lower tiles of A after the Cholesky steps (rows in the CALLER's order: the batch must have been
// created with keep_row_order, the whitening is not permutation invariant).  A validation entry point, not a hot path:
// plain FP64 FMAs, grid (row blocks of 64, models), 256 threads; thread t owns the columns k = t, t + 256, ... of L S.
// part[b][blk] = sum of E_i over the block's rows (block 0 adds -KL): the host adds the blocks in order.
// =============================================================================================
__global__ void __launch_bounds__(256) wv_elbo_rows_kernel(WvBatchDev bd, const double* __restrict__ xall,
                                                           const double* __restrict__ qmu, const double* __restrict__ qs,
                                                           double* __restrict__ part, int nblk) {
  __shared__ double red[8];
  __shared__ double s_row[2];
  const int b = blockIdx.y, blk = blockIdx.x;
  const int n = bd.n, ld = bd.npad;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const WvProgram* pg = bd.programs + bd.prog_id[b];
  const double* x = xall + (size_t)b * bd.P;
  auto slot_value = [&](int s_) {
    const WvSlot& sl = pg->slots[s_];
    return sl.xindex >= 0 ? wv_transform(sl.transform, x[sl.xindex], sl.shift) : sl.fixed;
  };
  const double cmean = pg->mean_slot >= 0 ? slot_value(pg->mean_slot) : 0.0;
  const double p1 = slot_value(pg->noise_slot);                 // Gaussian variance | NB / ZINB alpha | Gamma shape
  double lik_param = bd.lik_param, lik_param2 = bd.lik_param2;
  if (bd.lik == 2 || bd.lik == 4 || bd.lik == 5) { if (pg->slots[pg->noise_slot].xindex >= 0) lik_param = p1; }
  if (bd.lik == 5 && pg->lik_slot2 >= 0) lik_param2 = slot_value(pg->lik_slot2);
  const double* Lb = bd.A + (size_t)b * ld * ld;
  const double* mu = qmu + (size_t)b * n;
  const double* S = qs + (size_t)b * n * n;
  const double* yb = bd.Y + (size_t)b * ld;
  double esum = 0.0;                       // thread 0 accumulates the block's E_i in row order
  const int i1 = min(n, (blk + 1) * 64);
  for (int i = blk * 64; i < i1; ++i) {
    const double* Li = Lb + (size_t)i * ld;
    double fm = 0.0, fv = 0.0;
    for (int k = threadIdx.x; k <= i; k += blockDim.x) {
      double s_ = 0.0;
      for (int j = k; j <= i; ++j) s_ = fma(Li[j], S[(size_t)j * n + k], s_);
      fv = fma(s_, s_, fv);
      fm = fma(Li[k], mu[k], fm);
    }
    for (int o = 16; o > 0; o >>= 1) { fm += __shfl_xor_sync(0xffffffffu, fm, o); fv += __shfl_xor_sync(0xffffffffu, fv, o); }
    __syncthreads();
    if (lane == 0) red[warp] = fm;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < 8; ++w) t += red[w]; s_row[0] = t; }
    __syncthreads();
    if (lane == 0) red[warp] = fv;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      const double m_ = cmean + s_row[0], v_ = t, y_ = yb[i];
      double E;
      if (bd.lik == 0) {
        E = -0.5 * log(6.283185307179586 * p1) - ((y_ - m_) * (y_ - m_) + v_) / (2.0 * p1);
      } else {
        double g_, h_, da_, da2_;
        wv_var_exp(bd.lik, lik_param, lik_param2, y_, lgamma(y_ + 1.0), m_, v_, E, g_, h_, da_, da2_);
      }
      esum += E;
    }
  }
  if (blk == 0) {          // - KL[q(v) || N(0, I)]
    double kl = 0.0;
    for (size_t e = threadIdx.x; e < (size_t)n * n; e += blockDim.x) {
      const size_t r = e / n, c = e % n;
      if (c <= r) { const double sv = S[e]; kl += sv * sv; if (c == r) kl -= 2.0 * log(fabs(sv)); }
    }
    for (int k = threadIdx.x; k < n; k += blockDim.x) kl += mu[k] * mu[k];
    for (int o = 16; o > 0; o >>= 1) kl += __shfl_xor_sync(0xffffffffu, kl, o);
    __syncthreads();
    if (lane == 0) red[warp] = kl;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < 8; ++w) t += red[w]; esum -= 0.5 * (t - n); }
  }
  if (threadIdx.x == 0) part[(size_t)b * nblk + blk] = esum;
}

// log prior of the trainable parameters of every model (the same wv_prior the finalize kernel uses)
__global__ void wv_logprior_kernel(WvBatchDev bd, const double* __restrict__ xall, double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bd.B) return;
  const WvProgram* pg = bd.programs + bd.prog_id[b];
  double lp = 0.0;
  for (int s_ = 0; s_ < pg->n_slots; ++s_) {
    const WvSlot sl = pg->slots[s_];
    if (sl.xindex < 0) continue;
    double l, d;
    wv_prior(sl, wv_transform(sl.transform, xall[(size_t)b * bd.P + sl.xindex], sl.shift), &l, &d);
    lp += l;
  }
  out[b] = lp;
}

int wv_enqueue_elbo(const WvBatchDev& bd, const double* d_x, const double* d_qmu, const double* d_qs, double* d_part,
                    int nblk, double* d_logprior, cudaStream_t st) {
  wv_elbo_rows_kernel<<<dim3(nblk, bd.B), 256, 0, st>>>(bd, d_x, d_qmu, d_qs, d_part, nblk);
  wv_logprior_kernel<<<(bd.B + 63) / 64, 64, 0, st>>>(bd, d_x, d_logprior);
  return cudaGetLastError() == cudaSuccess ? 2 : -1;
}

// Gram + Cholesky + L^{-T} + alpha + K^{-1} of the listed models.  Returns launches or -1.
// launch of a run-time specialised element-wise kernel (same grid as the interpreter kernel it replaces)
static cudaError_t wv_launch_spec(const void* kern, int smem, const double* tab12, const WvBatchDev& bd, const int* d_active,
                                  int n_active, const double* d_x, int ntiles, int tpc, cudaStream_t st) {
  void* args[] = {(void*)&bd, (void*)&d_active, (void*)&d_x, (void*)&tab12, (void*)&ntiles, (void*)&tpc};
  return cudaLaunchKernel(kern, dim3((ntiles + tpc - 1) / tpc, n_active), dim3(WV_ELEM_THREADS), args, (size_t)smem, st);
}

int wv_enqueue_factor(const WvBatchDev& bd, const int* d_active, int n_active, const double* d_x, cudaStream_t st,
                      WvProfiler* pf, const WvAux* aux, const WvSpecLaunch* spec, bool chol_only) {
  if (n_active <= 0) return 0;
  if (wv_set_attrs() != cudaSuccess || !aux) return -1;
  int launches = 0;
  const int nt = bd.nt;
  const int ntiles = nt * (nt + 1) / 2;
  pf->mark(-1, st);
  const int tpc = wv_elem_tpc((long)ntiles * n_active);
  if (spec && spec->gram) {
    if (wv_launch_spec(spec->gram, spec->gram_smem, spec->tab12, bd, d_active, n_active, d_x, ntiles, tpc, st) != cudaSuccess)
      return -1;
  } else {
    wv_gram_kernel<<<dim3((ntiles + tpc - 1) / tpc, n_active), WV_ELEM_THREADS, sizeof(WvElemSmem), st>>>(
        bd, d_active, d_x, ntiles, tpc);
  }
  pf->mark(WV_K_GRAM, st);
  ++launches;
  if (aux->side && nt >= aux->big_nt) {
    int l = wv_enqueue_factor_big(bd, d_active, n_active, st, pf, *aux, chol_only);
    if (l < 0) return -1;
    launches += l;
    if (chol_only) return cudaGetLastError() == cudaSuccess ? launches : -1;
  } else {
    // few models (the tail rounds of a fit, small searches): the device is empty and the evaluation is a chain of
    // dependent launches -- one persistent launch for the whole factorisation (no lag: a diagonal block follows its
    // model's tiles directly and spins for the one it needs) and one for the triangular inverse (a CTA per tile row)
    const long tiles_all = (long)n_active * nt * (nt + 1) / 2;
    const bool few = aux->few_models && nt >= 2 && nt < 200 && tiles_all <= aux->resident_ctas;
    // ... and the same persistent launch (with the lag) for batches of up to about a thousand models that have the device
    // to themselves: it removes the launch pair per column and hides the diagonal chains behind panel tiles, which
    // outweighs its lower panel occupancy (3 instead of 4 CTAs per SM) until the batch fills every launch several times
    // over.  Beside other streams' work the persistent CTAs hold every slot and the streams stop filling each other's
    // gaps (measured: DESIGN.md, the section on the one-launch Cholesky)
    const long cols_all = (long)n_active * nt;
    const bool mid = nt >= 2 && nt < 200 &&
                     ((aux->chol_all == 1 && cols_all > 2L * aux->resident_ctas) ||
                      (aux->chol_all == 2 && aux->solo && cols_all <= aux->chol_all_max && cols_all > aux->chol_all_min));
    if (few || mid) {
      cudaMemsetAsync(bd.step_flag + 2 * (size_t)bd.B * bd.nt, 0, sizeof(int), st);      // the work counter
      const int ctas = few ? (int)tiles_all : aux->resident_ctas;
      wv_chol_all_kernel<<<ctas, WV_GEMM_THREADS, wv_smem_gemm_bytes(), st>>>(bd, d_active, n_active,
                                                                              few ? 0 : (int)(cols_all / 2 < aux->chol_lag ? cols_all / 2 : aux->chol_lag), aux->epoch);
      pf->mark(WV_K_CHOL_PANEL, st);
      ++launches;
    } else {
      for (int j = 0; j < nt; ++j) launches += wv_launch_chol_step(bd, d_active, n_active, j, 0, st, pf, *aux);
    }
    if (chol_only) return cudaGetLastError() == cudaSuccess ? launches : -1;
    const long tri_tiles = (long)n_active * nt * (nt - 1) / 2;
    if (aux->trtri_all && aux->solo && !few && nt > 1 && nt < 120 && tri_tiles > aux->resident_ctas &&
        cols_all <= aux->trtri_all_max) {
      cudaMemsetAsync(bd.step_flag + 2 * (size_t)bd.B * bd.nt, 0, sizeof(int), st);      // the work counter
      const int ctas = aux->trtri_ctas;
      wv_trtri_all_kernel<<<ctas, WV_GEMM_THREADS, sizeof(WvPanelSmem), st>>>(bd, d_active, n_active, aux->epoch);
      pf->mark(WV_K_TRTRI, st);
      ++launches;
    } else if ((aux->trtri_rows || few) && nt > 1) {
      const int paired = aux->trtri_rows == 2;
      wv_trtri_rows_kernel<<<dim3(paired ? nt / 2 : nt - 1, n_active), WV_GEMM_THREADS, sizeof(WvPanelSmem), st>>>(bd, d_active,
                                                                                                              paired);
      pf->mark(WV_K_TRTRI, st);
      ++launches;
    } else {
      for (int i = 1; i < nt; ++i) {
        wv_trtri_kernel<<<dim3(i, n_active), WV_GEMM_THREADS, sizeof(WvPanelSmem), st>>>(bd, d_active, i);
        pf->mark(WV_K_TRTRI, st);
        ++launches;
      }
    }
  }
  wv_extract_kernel<<<dim3(n_active), 256, 0, st>>>(bd, d_active);
  pf->mark(WV_K_EXTRACT, st);
  if (aux->tmap_mt) {
    wv_kinv_tma_kernel<<<dim3(ntiles, n_active), WV_GEMM_THREADS, sizeof(WvTmaSmem) + 1024, st>>>(
        *reinterpret_cast<const CUtensorMap*>(aux->tmap_mt), bd, d_active);
  } else {
    wv_kinv_kernel<<<dim3(ntiles, n_active), WV_GEMM_THREADS, sizeof(WvGemmSmem), st>>>(bd, d_active);
  }
  pf->mark(WV_K_KINV, st);
  launches += 2;
  return cudaGetLastError() == cudaSuccess ? launches : -1;
}

// gradient reduction + objective of the listed models (after wv_enqueue_factor)
int wv_enqueue_grad_finalize(const WvBatchDev& bd, const int* d_active, int n_active, const double* d_x, double* d_f,
                             double* d_g, double* d_lml, int* d_status, cudaStream_t st, WvProfiler* pf,
                             const WvSpecLaunch* spec) {
  if (n_active <= 0) return 0;
  const int nt = bd.nt;
  const int ntiles = nt * (nt + 1) / 2;
  const int tpc = wv_elem_tpc((long)ntiles * n_active);
  if (spec && spec->grad) {
    if (wv_launch_spec(spec->grad, spec->grad_smem, spec->tab12, bd, d_active, n_active, d_x, ntiles, tpc, st) != cudaSuccess)
      return -1;
  } else {
    wv_grad_kernel<<<dim3((ntiles + tpc - 1) / tpc, n_active), WV_ELEM_THREADS, sizeof(WvElemSmem), st>>>(
        bd, d_active, d_x, ntiles, tpc);
  }
  pf->mark(WV_K_GRAD, st);
  wv_finalize_kernel<<<dim3(n_active), ntiles > 128 ? 64 * WV_FIN_MAXG : 64, 0, st>>>(bd, d_active, d_x, ntiles, d_f, d_g,
                                                                                  d_lml, d_status);
  pf->mark(WV_K_FINALIZE, st);
  return cudaGetLastError() == cudaSuccess ? 2 : -1;
}

// one sweep of the site iteration after wv_enqueue_factor, and the compaction of the models that go on
int wv_enqueue_site_sweep(const WvBatchDev& bd, const WvVgpState& vs, const int* d_list, int n_list, const double* d_x,
                          int* d_next, int* d_count, cudaStream_t st, WvProfiler* pf) {
  wv_site_update_kernel<<<dim3(n_list), 256, 0, st>>>(bd, vs, d_list, d_x);
  wv_compact_list_kernel<<<1, 1024, 0, st>>>(d_list, n_list, vs.inner_task, d_next, d_count);
  pf->mark(WV_K_SITES, st);
  return cudaGetLastError() == cudaSuccess ? 2 : -1;
}

int wv_enqueue_vgp_begin(const WvVgpState& vs, const int* d_list, int n_list, cudaStream_t st) {
  wv_vgp_begin_kernel<<<(n_list + 255) / 256, 256, 0, st>>>(vs, d_list, n_list);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int wv_enqueue_vgp_status(const WvVgpState& vs, const int* d_list, int n_list, int* d_status, cudaStream_t st) {
  wv_vgp_status_kernel<<<(n_list + 255) / 256, 256, 0, st>>>(vs, d_list, n_list, d_status);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// Returns the number of kernel launches enqueued (for bench.py's gpu_launches), or -1 on error.
int wv_enqueue_eval(const WvBatchDev& bd, const int* d_active, int n_active, const double* d_x, double* d_f,
                    double* d_g, double* d_lml, int* d_status, cudaStream_t st, WvProfiler* pf, const WvAux* aux,
                    const WvSpecLaunch* spec) {
  if (n_active <= 0) return 0;
  WvProfiler none;
  if (!pf) pf = &none;
  cudaMemsetAsync(bd.chol_fail, 0, sizeof(int) * bd.B, st);
  const int l1 = wv_enqueue_factor(bd, d_active, n_active, d_x, st, pf, aux, spec, false);
  if (l1 < 0) return -1;
  const int l2 = wv_enqueue_grad_finalize(bd, d_active, n_active, d_x, d_f, d_g, d_lml, d_status, st, pf, spec);
  if (l2 < 0) return -1;
  return l1 + l2;
}

// posterior mean at new inputs for every model of the batch (uses bd.alpha of the last evaluation at d_x)
int wv_enqueue_cross_mean(const WvBatchDev& bd, const double* d_x, const double* d_xnew_t, int m, int mpad, double* d_mean,
                          cudaStream_t st) {
  if (wv_set_attrs() != cudaSuccess) return -1;
  wv_cross_mean_kernel<<<dim3(mpad / WV_NB, bd.B), WV_ELEM_THREADS, sizeof(WvElemSmem), st>>>(bd, d_x, d_xnew_t, m, mpad,
                                                                                             d_mean);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// predictive variance of f at new inputs for every model (uses K^-1 left in A by the last evaluation at d_x)
int wv_enqueue_cross_var(const WvBatchDev& bd, const double* d_x, const double* d_xnew_t, int m, int mpad, double* d_part,
                         double* d_prior, double* d_var, cudaStream_t st) {
  if (wv_set_attrs() != cudaSuccess) return -1;
  const int ntiles = bd.nt * (bd.nt + 1) / 2;
  wv_cross_var_kernel<<<dim3(ntiles, mpad / WV_NB, bd.B), WV_ELEM_THREADS, sizeof(WvCrossVarSmem), st>>>(
      bd, d_x, d_xnew_t, m, mpad, d_part, d_prior);
  const size_t total = (size_t)bd.B * m;
  wv_cross_var_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(bd.B, ntiles, m, d_part, d_prior, d_var);
  return cudaGetLastError() == cudaSuccess ? 2 : -1;
}
