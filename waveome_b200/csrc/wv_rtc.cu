// waveome_b200 — run-time compilation of the specialised element-wise kernels (waveome_b200/specialize.py writes the
// CUDA text, csrc/wv_spec.cuh is its prelude).  NVRTC is loaded with dlopen so that the library itself has no link-time
// dependency on it; the cubin is loaded through the runtime's library API (cudaLibraryLoadData / cudaLibraryGetKernel),
// and the kernels are launched with cudaLaunchKernel like any other.  Compiled libraries are cached per process by the
// hash of their source text.
#include <dlfcn.h>
#include <unistd.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "wv_rtc.h"
#include "wv_common.cuh"

namespace {

typedef struct _nvrtcProgram* nvrtcProgram;
typedef int nvrtcResult;

struct NvrtcApi {
  void* handle = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  const char* (*GetErrorString)(nvrtcResult) = nullptr;
  std::string error;
};

std::mutex g_mu;
NvrtcApi g_nvrtc;
bool g_nvrtc_tried = false;

bool load_nvrtc_locked() {
  if (g_nvrtc_tried) return g_nvrtc.handle != nullptr;
  g_nvrtc_tried = true;
  std::vector<std::string> names;
  if (const char* p = getenv("WV_NVRTC_LIB")) names.push_back(p);
  names.push_back("libnvrtc.so.12");
  names.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
  names.push_back("libnvrtc.so");
  for (const auto& n : names) {
    g_nvrtc.handle = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (g_nvrtc.handle) break;
  }
  if (!g_nvrtc.handle) {
    g_nvrtc.error = std::string("libnvrtc not found (") + (dlerror() ? dlerror() : "dlopen failed") + ")";
    return false;
  }
#define WV_SYM(field, name)                                                        \
  *reinterpret_cast<void**>(&g_nvrtc.field) = dlsym(g_nvrtc.handle, name);        \
  if (!g_nvrtc.field) { g_nvrtc.error = std::string("libnvrtc lacks ") + name; g_nvrtc.handle = nullptr; return false; }
  WV_SYM(CreateProgram, "nvrtcCreateProgram");
  WV_SYM(CompileProgram, "nvrtcCompileProgram");
  WV_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
  WV_SYM(GetCUBIN, "nvrtcGetCUBIN");
  WV_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
  WV_SYM(GetProgramLog, "nvrtcGetProgramLog");
  WV_SYM(DestroyProgram, "nvrtcDestroyProgram");
  WV_SYM(GetErrorString, "nvrtcGetErrorString");
#undef WV_SYM
  return true;
}

struct LoadedLib {
  cudaLibrary_t lib = nullptr;
  std::vector<char> cubin;     // kept alive: the runtime may reference the image lazily
};
std::map<std::string, LoadedLib*> g_libs;      // source text hash (hex, from the caller) -> library

double* g_tab12[64] = {};
std::string g_cache_dir;      // cubins by source hash ("" = no disk cache)

bool read_file(const std::string& path, std::vector<char>* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  bool ok = n > 0;
  if (ok) { out->resize((size_t)n); ok = fread(out->data(), 1, (size_t)n, f) == (size_t)n; }
  fclose(f);
  return ok;
}
void write_file_atomic(const std::string& path, const std::vector<char>& data) {
  const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;                                   // a read-only cache directory is not an error
  const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
  fclose(f);
  if (ok) rename(tmp.c_str(), path.c_str()); else remove(tmp.c_str());
}

}  // namespace

// Compile `src` for sm_100a.  Needs no GPU (the CPU test suite checks that every generated text compiles).
int wv_rtc_compile(const char* src, std::vector<char>* cubin, std::string* log) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (!load_nvrtc_locked()) { *log = g_nvrtc.error; return -1; }
  nvrtcProgram prog = nullptr;
  nvrtcResult r = g_nvrtc.CreateProgram(&prog, src, "wv_specialized.cu", 0, nullptr, nullptr);
  if (r != 0) { *log = std::string("nvrtcCreateProgram: ") + g_nvrtc.GetErrorString(r); return -1; }
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
  r = g_nvrtc.CompileProgram(prog, 3, opts);
  size_t ls = 0;
  g_nvrtc.GetProgramLogSize(prog, &ls);
  if (ls > 1) { log->resize(ls); g_nvrtc.GetProgramLog(prog, &(*log)[0]); }
  if (r != 0) {
    *log = std::string("nvrtcCompileProgram: ") + g_nvrtc.GetErrorString(r) + "\n" + *log;
    g_nvrtc.DestroyProgram(&prog);
    return -1;
  }
  size_t cs = 0;
  r = g_nvrtc.GetCUBINSize(prog, &cs);
  if (r != 0 || cs == 0) { *log = "nvrtcGetCUBINSize failed"; g_nvrtc.DestroyProgram(&prog); return -1; }
  cubin->resize(cs);
  r = g_nvrtc.GetCUBIN(prog, cubin->data());
  g_nvrtc.DestroyProgram(&prog);
  if (r != 0) { *log = "nvrtcGetCUBIN failed"; return -1; }
  return 0;
}

// Compile (or fetch from the per-process cache) and resolve the two kernels.
int wv_rtc_get_kernels(const char* key, const char* src, const char* gram_name, const char* grad_name, WvSpecLaunch* out,
                       std::string* err) {
  LoadedLib* L = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_libs.find(key);
    if (it != g_libs.end()) L = it->second;
  }
  if (!L) {
    LoadedLib* fresh = new LoadedLib();
    std::string log, path;
    {
      std::lock_guard<std::mutex> lock(g_mu);
      if (!g_cache_dir.empty()) path = g_cache_dir + "/" + key + ".cubin";
    }
    cudaError_t e = cudaErrorUnknown;
    if (!path.empty() && read_file(path, &fresh->cubin)) {      // compiled by an earlier process (or by build())
      e = cudaLibraryLoadData(&fresh->lib, fresh->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
      if (e != cudaSuccess) { cudaGetLastError(); fresh->cubin.clear(); }
    }
    if (e != cudaSuccess) {
      if (wv_rtc_compile(src, &fresh->cubin, &log) != 0) { *err = log; delete fresh; return -1; }
      if (!path.empty()) write_file_atomic(path, fresh->cubin);
      e = cudaLibraryLoadData(&fresh->lib, fresh->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    }
    if (e != cudaSuccess) { *err = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e); delete fresh; return -1; }
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_libs.find(key);
    if (it != g_libs.end()) { cudaLibraryUnload(fresh->lib); delete fresh; L = it->second; }
    else { g_libs[key] = fresh; L = fresh; }
  }
  cudaKernel_t kg = nullptr, kd = nullptr;
  cudaError_t e = cudaLibraryGetKernel(&kg, L->lib, gram_name);
  if (e == cudaSuccess) e = cudaLibraryGetKernel(&kd, L->lib, grad_name);
  if (e != cudaSuccess) { *err = std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(e); return -1; }
  out->gram = (const void*)kg;
  out->grad = (const void*)kd;
  return 0;
}

void wv_rtc_set_cache_dir(const char* dir) {
  std::lock_guard<std::mutex> lock(g_mu);
  g_cache_dir = dir ? dir : "";
}

// compile into the disk cache without loading (build step: needs no GPU).  Returns 0, or 1 if the entry existed.
int wv_rtc_precompile(const char* key, const char* src, std::string* err) {
  std::string path;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_cache_dir.empty()) { *err = "no cache directory set"; return -1; }
    path = g_cache_dir + "/" + key + ".cubin";
  }
  std::vector<char> cubin;
  if (read_file(path, &cubin)) return 1;
  if (wv_rtc_compile(src, &cubin, err) != 0) return -1;
  write_file_atomic(path, cubin);
  return 0;
}

// 2^(j/2048), j = 0..2047, correctly rounded from long double; one copy per device, never freed
int wv_rtc_exp2_table(int device, const double** out, std::string* err) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (device < 0 || device >= 64) { *err = "device index beyond the table cache"; return -1; }
  if (!g_tab12[device]) {
    std::vector<double> h(WV_EXP2_BIG_TAB);
    for (int j = 0; j < WV_EXP2_BIG_TAB; ++j) h[j] = (double)exp2l((long double)j / WV_EXP2_BIG_TAB);
    double* d = nullptr;
    cudaError_t e = cudaMalloc(&d, h.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { *err = std::string("exp2 table: ") + cudaGetErrorString(e); return -1; }
    g_tab12[device] = d;
  }
  *out = g_tab12[device];
  return 0;
}
