// dependent-chain latencies of the ops on the diagonal-block critical path (single warp): cycles per op
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double* out, long long* clk, double x0) {
  const int N = 256;
  double x = x0 + threadIdx.x * 1e-3, y = 1.0 + 1e-9 * threadIdx.x, c1 = 0;
  long long t0, t1;
  int i = 0;
#define RUN(idx, BODY) \
  __syncwarp(); t0 = clock64(); \
  _Pragma("unroll 16") for (i = 0; i < N; ++i) { BODY; } \
  t1 = clock64(); if (threadIdx.x == 0) clk[idx] = t1 - t0;
  RUN(0, x = fma(x, y, 1e-9));
  RUN(1, x = x * y);
  RUN(2, x = x + y);
  RUN(3, x = rsqrt(x) + 1.5);
  RUN(4, x = __shfl_sync(0xffffffffu, x, (i + 1) & 31));
  RUN(5, x = sqrt(x) + 1.5);
  RUN(6, x = 1.0 / x + 1.5);
  RUN(7, dmma(x, c1, y, y));
  RUN(8, x = log(x) + 3.0);
  RUN(9, x = exp(-x * 1e-3) + 1.0);
  { float f = (float)x; RUN(10, f = rsqrtf(f) + 1.5f); x += f; }
  { float f = (float)x; RUN(11, f = fmaf(f, 1.0001f, 1e-9f)); x += f; }
  { int q = (int)x; RUN(12, q = __shfl_sync(0xffffffffu, q, (i + 1) & 31)); x += q; }
  out[threadIdx.x] = x + c1;
}
int main() {
  double* out; long long* clk;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&clk, 16 * 8);
  for (int r = 0; r < 2; ++r) k<<<1, 32>>>(out, clk, 1.7);
  long long h[16]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  const char* nm[] = {"DFMA", "DMUL", "DADD", "rsqrt(double)+DADD", "SHFL double", "sqrt(double)+DADD", "1/x double + DADD", "DMMA m8n8k4 (dep. accumulator)", "log(double)+DADD", "exp(double)+..", "rsqrtf+FADD", "FFMA", "SHFL int"};
  for (int i = 0; i < 13; ++i) printf("%-34s %7.1f cycles/op\n", nm[i], h[i] / 256.0);
  return 0;
}
