// Latency harness for wv_chol_diag_kernel / wv_panel_kernel<0>: solo launch time (CUDA events) and in-kernel phase clocks.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DWV_DIAG_CLOCK -o scratch/diag_bench scratch/diag_bench.cu
#include <cstdio>
#include <vector>
#include "../waveome_b200/csrc/wv_eval.cu"

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 1;
  const int nt = 4, npad = nt * 64, n = npad - 9;
  WvBatchDev bd{};
  bd.n = n; bd.D = 1; bd.B = B; bd.npad = npad; bd.nt = nt; bd.n8 = (n + 8) / 8 * 8; bd.P = 4; bd.n_slots_max = 4;
  std::vector<double> h((size_t)npad * npad, 0.0);
  for (int i = 0; i < npad; ++i)
    for (int j = 0; j <= i; ++j) h[(size_t)i * npad + j] = i == j ? 4.0 : 0.5 / (1.0 + (i - j));
  double *A, *Mt, *Dinv, *ld_;
  int *fail, *act;
  cudaMalloc(&A, sizeof(double) * B * npad * npad); cudaMalloc(&Mt, sizeof(double) * B * npad * npad);
  cudaMalloc(&Dinv, sizeof(double) * B * nt * 4096); cudaMalloc(&ld_, sizeof(double) * B * nt);
  cudaMalloc(&fail, sizeof(int) * B); cudaMalloc(&act, sizeof(int) * B);
  std::vector<int> ia(B); for (int i = 0; i < B; ++i) ia[i] = i;
  cudaMemcpy(act, ia.data(), sizeof(int) * B, cudaMemcpyHostToDevice);
  for (int b = 0; b < B; ++b) cudaMemcpy(A + (size_t)b * npad * npad, h.data(), sizeof(double) * npad * npad, cudaMemcpyHostToDevice);
  bd.A = A; bd.Mt = Mt; bd.Dinv = Dinv; bd.logdet_part = ld_; bd.chol_fail = fail;
  wv_set_attrs();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  // factorise columns 0..2 once so that j = 3 has real data
  int* flags; cudaMalloc(&flags, sizeof(int) * B * nt); cudaMemset(flags, 0, sizeof(int) * B * nt);
  bd.step_flag = flags;
  int epoch = 0;
  const size_t smem = wv_smem_gemm_bytes();
  for (int j = 0; j < 3; ++j) wv_chol_step_kernel<<<dim3(nt - j, B), 128, smem>>>(bd, act, j, 0, ++epoch, 0);
  cudaDeviceSynchronize();
  for (int k0 : {0, 192}) {
    const int reps = 50;
    for (int w = 0; w < 5; ++w) wv_chol_step_kernel<<<dim3(1, B), 128, smem>>>(bd, act, 3, k0, ++epoch, 0);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) wv_chol_step_kernel<<<dim3(1, B), 128, smem>>>(bd, act, 3, k0, ++epoch, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long clk[64];
    cudaMemcpyFromSymbol(clk, wv_dbg_clk, sizeof(clk));
    printf("diag B=%d K=%d: %.2f us/launch (back-to-back); phases (cycles from start):", B, 192 - k0, ms * 1e3 / reps);
    for (int i = 1; i <= 13; ++i) printf(" %lld", clk[i] - clk[0]);
    printf("\n   last potrf16: load->start %lld, columns:", clk[18] - clk[11]);
    for (int i = 20; i < 36; ++i) printf(" %lld", clk[i] - clk[i == 20 ? 18 : i - 1]);
    printf("  | end %lld, stores %lld, to barrier %lld\n", clk[19] - clk[35], clk[36] - clk[19], clk[11] + 0 - clk[36]);
  }
  {
    const int reps = 50;
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) wv_chol_step_kernel<<<dim3(2, B), 128, smem>>>(bd, act, 2, 0, ++epoch, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("fused step (diag + 1 panel tile) B=%d K=128: %.2f us/launch\n", B, ms * 1e3 / reps);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
