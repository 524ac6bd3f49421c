// waveome_b200 — run-time specialised element-wise kernels: host-side handles (see wv_rtc.cu, wv_spec.cuh)
#pragma once
#include <string>
#include <vector>

// what the launch sequence needs to run a batch's Gram / gradient pass on its specialised kernels
struct WvSpecLaunch {
  const void* gram = nullptr;      // cudaKernel_t of the generated kernels (cudaLaunchKernel takes them as is)
  const void* grad = nullptr;
  int gram_smem = 0, grad_smem = 0;
  const double* tab12 = nullptr;   // device table 2^(j/4096)
};

int wv_rtc_compile(const char* src, std::vector<char>* cubin, std::string* log);
int wv_rtc_get_kernels(const char* key, const char* src, const char* gram_name, const char* grad_name, WvSpecLaunch* out,
                       std::string* err);
int wv_rtc_exp2_table(int device, const double** out, std::string* err);
void wv_rtc_set_cache_dir(const char* dir);
int wv_rtc_precompile(const char* key, const char* src, std::string* err);
