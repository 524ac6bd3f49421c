// single-warp issue costs: independent double shuffles, smem broadcast round trip
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* clk, double x0) {
  __shared__ double sh[64];
  const int N = 64, lane = threadIdx.x;
  double v[16];
  for (int q = 0; q < 16; ++q) v[q] = x0 + lane * 1e-3 + q;
  long long t0, t1;
  // (0) 16 independent double shuffles + 16 fma per iteration
  __syncwarp(); t0 = clock64();
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = fma(__shfl_sync(0xffffffffu, v[q], q), 1.0000001, v[(q + 1) & 15]);
  }
  t1 = clock64(); if (lane == 0) clk[0] = (t1 - t0) / N;
  // (1) dependent double shuffle chain, runtime lane
  double x = v[3];
  int src = (lane * 7 + 3) & 31;
  __syncwarp(); t0 = clock64();
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int q = 0; q < 16; ++q) x = __shfl_sync(0xffffffffu, x, src) + 1.0;
  }
  t1 = clock64(); if (lane == 0) clk[1] = (t1 - t0) / N;
  // (2) STS -> syncwarp -> 16 broadcast LDS.64 + 16 fma per iteration
  __syncwarp(); t0 = clock64();
  for (int i = 0; i < N; ++i) {
    sh[lane] = x;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = fma(sh[q], 1.0000001, v[q]);
    x = v[i & 15];
    __syncwarp();
  }
  t1 = clock64(); if (lane == 0) clk[2] = (t1 - t0) / N;
  // (3) STS -> syncwarp -> one LDS dependent round trip
  __syncwarp(); t0 = clock64();
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { sh[lane] = x; __syncwarp(); x = sh[(lane + 1) & 31] + 1.0; __syncwarp(); }
  }
  t1 = clock64(); if (lane == 0) clk[3] = (t1 - t0) / N;
  double s = x;
  for (int q = 0; q < 16; ++q) s += v[q];
  out[lane] = s;
}
int main() {
  double* out; long long* clk;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&clk, 16 * 8);
  for (int r = 0; r < 2; ++r) k<<<1, 32>>>(out, clk, 1.7);
  long long h[16]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  printf("16 indep double SHFL + 16 DFMA        : %lld cycles / iteration\n", h[0]);
  printf("16 dependent (double SHFL + DADD)     : %lld cycles / iteration (%.1f per pair)\n", h[1], h[1] / 16.0);
  printf("STS, sync, 16 bcast LDS.64 + 16 DFMA  : %lld cycles / iteration\n", h[2]);
  printf("4 x (STS, sync, LDS, DADD, sync) chain: %lld cycles / iteration (%.1f per round trip)\n", h[3], h[3] / 4.0);
  return 0;
}
