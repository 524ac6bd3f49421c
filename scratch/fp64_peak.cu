// Microbenchmark: FP64 DMMA (mma.sync f64) shapes vs DFMA peak on sm_100a. Scratch (not product).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__global__ void k_dfma(double* out, int iters) {
  double a[8], b = 1.0000001, c = 0.9999999;
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);
  }
  double s = 0; for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma884(double* out, int iters) {
  double c[8][2]; double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0; for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma1684(double* out, int iters) {
  double c[4][4]; double a0 = threadIdx.x * 1e-3, a1 = 0.5, b = 1.0 + threadIdx.x * 1e-6;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b));
  }
  double s = 0; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma1688(double* out, int iters) {
  double c[4][4]; double a[4], b[2];
  for (int i = 0; i < 4; i++) a[i] = threadIdx.x * 1e-3 + i;
  b[0] = 1.0; b[1] = 0.5;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  }
  double s = 0; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double* out, int iters) {
  double c[4][4]; double a[8], b[4];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; i++) b[i] = 1.0 / (i + 1);
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024));
  int iters = 20000;
  for (int wps = 4; wps <= 32; wps *= 2) {   // warps per SM (1 CTA/SM... use 2 CTAs per SM of wps/2 warps)
    int threads = wps * 32 / 2; if (threads > 1024) threads = 1024; int blocks = 148 * 2;
    double tot_warps = (double)blocks * threads / 32;
    float ms;
    ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); });
    printf("warps/SM %2d  DFMA      : %8.2f TFLOP/s\n", wps, tot_warps * 32 * 8.0 * iters * 2 / ms * 1e-9);
    ms = timeit([&] { k_dmma884<<<blocks, threads>>>(out, iters); });
    printf("warps/SM %2d  DMMA 8x8x4 : %8.2f TFLOP/s\n", wps, tot_warps * 8.0 * iters * 2 * 8 * 8 * 4 / ms * 1e-9);
    ms = timeit([&] { k_dmma1684<<<blocks, threads>>>(out, iters); });
    printf("warps/SM %2d  DMMA 16x8x4: %8.2f TFLOP/s\n", wps, tot_warps * 4.0 * iters * 2 * 16 * 8 * 4 / ms * 1e-9);
    ms = timeit([&] { k_dmma1688<<<blocks, threads>>>(out, iters); });
    printf("warps/SM %2d  DMMA 16x8x8: %8.2f TFLOP/s\n", wps, tot_warps * 4.0 * iters * 2 * 16 * 8 * 8 / ms * 1e-9);
    ms = timeit([&] { k_dmma16816<<<blocks, threads>>>(out, iters); });
    printf("warps/SM %2d  DMMA16x8x16: %8.2f TFLOP/s\n", wps, tot_warps * 4.0 * iters * 2 * 16 * 8 * 16 / ms * 1e-9);
  }
  return 0;
}
