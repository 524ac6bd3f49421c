// waveome_b200 — C ABI (include/waveome_b200.h): engine / batch lifecycle, batched evaluation and the
// device-resident L-BFGS-B fit loop.  No CPU fallback: every entry point needs a CUDA device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <set>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/waveome_b200.h"
#include "wv_kernels.cuh"
#include "wv_lbfgsb.h"
#include "wv_rtc.h"

int wv_enqueue_eval(const WvBatchDev& bd, const int* d_active, int n_active, const double* d_x, double* d_f,
                    double* d_g, double* d_lml, int* d_status, cudaStream_t st, WvProfiler* pf, const WvAux* aux,
                    const WvSpecLaunch* spec);

int wv_enqueue_factor(const WvBatchDev& bd, const int* d_active, int n_active, const double* d_x, cudaStream_t st,
                      WvProfiler* pf, const WvAux* aux, const WvSpecLaunch* spec, bool chol_only);
int wv_make_tmap_mt(const WvBatchDev& bd, void* out128);
int wv_enqueue_elbo(const WvBatchDev& bd, const double* d_x, const double* d_qmu, const double* d_qs, double* d_part,
                    int nblk, double* d_logprior, cudaStream_t st);
int wv_enqueue_grad_finalize(const WvBatchDev& bd, const int* d_active, int n_active, const double* d_x, double* d_f,
                             double* d_g, double* d_lml, int* d_status, cudaStream_t st, WvProfiler* pf,
                             const WvSpecLaunch* spec);
int wv_enqueue_site_sweep(const WvBatchDev& bd, const WvVgpState& vs, const int* d_list, int n_list, const double* d_x,
                          int* d_next, int* d_count, cudaStream_t st, WvProfiler* pf);
int wv_enqueue_vgp_begin(const WvVgpState& vs, const int* d_list, int n_list, cudaStream_t st);
int wv_enqueue_vgp_status(const WvVgpState& vs, const int* d_list, int n_list, int* d_status, cudaStream_t st);
int wv_enqueue_cross_mean(const WvBatchDev& bd, const double* d_x, const double* d_xnew_t, int m, int mpad, double* d_mean,
                          cudaStream_t st);

int wv_enqueue_cross_var(const WvBatchDev& bd, const double* d_x, const double* d_xnew_t, int m, int mpad, double* d_part,
                         double* d_prior, double* d_var, cudaStream_t st);

static thread_local std::string g_err;
static int wv_fail(const std::string& m) { g_err = m; return -1; }
#define WV_CUDA(x)                                                                                   \
  do {                                                                                               \
    cudaError_t e_ = (x);                                                                            \
    if (e_ != cudaSuccess) return wv_fail(std::string(#x) + ": " + cudaGetErrorString(e_));          \
  } while (0)

// Device buffers of destroyed batches, kept for the next batch on the same GPU (oldest first): cudaMalloc / cudaFree of
// the multi-GB workspaces costs 0.1 - 2 s each and every cudaFree of a small one synchronises the device, which is
// visible next to a 4 s fit (fit -> post-fit batches, search levels).  Small buffers are binned to powers of two so that
// batches of different sizes share them.  ONE cache per device, shared by all engines of the process (several engines
// = several streams fitting sub-batches concurrently, model_fitting.fit_replicated): a buffer enters it only after its
// batch's stream has been synchronised (wv_batch_destroy), so any engine may take it; a failed cudaMalloc flushes what
// every engine of the device has parked.
struct WvDeviceCache {
  std::mutex mu;
  std::vector<std::pair<size_t, void*>> cache;
  size_t cache_bytes = 0;
  int engines = 0;
};
static const int WV_MAX_DEVICES = 64;
static WvDeviceCache g_dev_cache[WV_MAX_DEVICES];

struct wv_engine {
  int device;
  cudaStream_t stream;
  WvAux aux;   // side stream / events of the large-n look-ahead schedule
  WvDeviceCache* dc;
  // pinned scratch for the per-round counters the host reads back (active models, unconverged sites).  Owned by the
  // engine, not the batch: cudaFreeHost waits for the whole device, which would park a finished sub-batch's host thread
  // until every other stream has drained.  The batches of an engine are driven by one host thread, one call at a time.
  int* h_count;
};
static const size_t WV_CACHE_SMALL_BYTES = (size_t)1 << 20;      // below: power-of-two bins
static const size_t WV_CACHE_MAX_ENTRIES = 4096;    // small buffers are cheap to keep; evicting costs a device-wide sync
static size_t WV_CACHE_MAX_BYTES = (size_t)110 << 30;      // of 180 GB (env WV_CACHE_MAX_GB); a failed cudaMalloc flushes the cache anyway

static void wv_cache_flush_locked(WvDeviceCache* dc) {
  for (auto& c : dc->cache) cudaFree(c.second);
  dc->cache.clear();
  dc->cache_bytes = 0;
}

static void wv_cache_put(wv_engine* e, void* p, size_t bytes) {
  WvDeviceCache* dc = e->dc;
  std::lock_guard<std::mutex> lock(dc->mu);
  dc->cache.push_back({bytes, p});
  dc->cache_bytes += bytes;
  while (!dc->cache.empty() && (dc->cache.size() > WV_CACHE_MAX_ENTRIES || dc->cache_bytes > WV_CACHE_MAX_BYTES)) {
    cudaFree(dc->cache.front().second);
    dc->cache_bytes -= dc->cache.front().first;
    dc->cache.erase(dc->cache.begin());
  }
}

struct wv_batch {
  wv_engine* eng;
  WvBatchDev bd;
  std::vector<std::pair<void*, size_t>> allocs;
  // device buffers
  double *d_x, *d_f, *d_g, *d_lml;
  int *d_status, *d_active, *d_active2, *d_count, *d_task, *d_nx, *d_iter, *d_neval, *d_st2;
  WvLbScalars* d_lbs;
  double* d_lbw;
  int lb_m_alloc;
  WvLbOpts lb_opts;              // the fit in progress (wv_batch_fit_lbfgs_begin / _run / _report)
  int lb_n_active = 0;
  int *lb_cur = nullptr, *lb_nxt = nullptr;
  long lb_guard = 0, lb_guard_max = 0;
  int* h_count;   // the engine's pinned counters
  int64_t bytes, launches, rounds, model_evals;
  WvVgpState vgp;         // site-iteration state (variational path), arrays allocated by wv_batch_set_likelihood
  int *d_inner1, *d_inner2;
  int64_t site_sweeps;
  const double* last_x;   // device pointer of the parameters of the last full evaluation (alpha belongs to them)
  WvProfiler prof;
  std::vector<int> perm;   // device row i holds caller row perm[i] (rows sorted by their categorical columns)
  WvSpecLaunch spec;       // run-time specialised Gram / gradient kernels of the batch's program (wv_batch_specialize)
  bool has_spec = false;
  int n_programs = 0;
  alignas(64) unsigned char tmap_mt[128];   // CUtensorMap over Mt (WV_KINV_TMA=1)
  bool has_tmap = false;
  int solo = 0;                  // wv_batch_set_solo: nothing else runs on the device beside this batch's calls
  bool keep_row_order = false;   // WV_BATCH_KEEP_ROW_ORDER: device rows = caller rows (needed by wv_batch_eval_elbo)
};

template <typename T> static int wv_alloc(wv_batch* b, T** p, size_t count) {
  void* q = nullptr;
  size_t bytes = count * sizeof(T);
  if (bytes == 0) bytes = sizeof(T);
  if (bytes < WV_CACHE_SMALL_BYTES) {
    size_t bin = 512;
    while (bin < bytes) bin <<= 1;
    bytes = bin;
  }
  {        // best fit from the device's cache (at most 25 % larger than asked)
    WvDeviceCache* dc = b->eng->dc;
    std::lock_guard<std::mutex> lock(dc->mu);
    auto& cache = dc->cache;
    int best = -1;
    for (int i = 0; i < (int)cache.size(); ++i)
      if (cache[i].first >= bytes && cache[i].first <= bytes + bytes / 4 && (best < 0 || cache[i].first < cache[best].first))
        best = i;
    if (best >= 0) {
      q = cache[best].second;
      bytes = cache[best].first;
      dc->cache_bytes -= bytes;
      cache.erase(cache.begin() + best);
    }
  }
  if (!q) {
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) {     // give the cached buffers back and retry
      cudaGetLastError();
      {
        std::lock_guard<std::mutex> lock(b->eng->dc->mu);
        wv_cache_flush_locked(b->eng->dc);
      }
      e = cudaMalloc(&q, bytes);
    }
    if (e != cudaSuccess) return wv_fail(std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
  }
  b->allocs.push_back({q, bytes});
  b->bytes += (int64_t)bytes;
  *p = reinterpret_cast<T*>(q);
  return 0;
}

extern "C" const char* wv_last_error(void) { return g_err.c_str(); }
extern "C" const char* wv_version(void) { return "waveome_b200 0.1 (sm_100a, fp64 DMMA)"; }

extern "C" int wv_engine_create2(int device, int flags, wv_engine** out);
extern "C" int wv_engine_create(int device, wv_engine** out) { return wv_engine_create2(device, 0, out); }

extern "C" int wv_engine_create2(int device, int flags, wv_engine** out) {
  if (!out) return wv_fail("wv_engine_create: out is null");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return wv_fail(std::string("wv_engine_create: no CUDA device (") + cudaGetErrorString(e) +
                   "); waveome_b200 has no CPU fallback");
  if (device < 0 || device >= count) return wv_fail("wv_engine_create: bad device index");
  WV_CUDA(cudaSetDevice(device));
  if (device >= WV_MAX_DEVICES) return wv_fail("wv_engine_create: device index beyond the cache table");
  wv_engine* eng = new wv_engine();
  eng->device = device;
  eng->dc = &g_dev_cache[device];
  {
    std::lock_guard<std::mutex> lock(eng->dc->mu);
    ++eng->dc->engines;
  }
  // the main stream carries the serial chain of the factorisation (diagonal blocks, panels): it gets the highest
  // priority so that its few CTAs are placed ahead of the bulk trailing updates queued on the side stream
  // ... except behind the streams of WV_ENGINE_HIGH_PRIORITY engines: those carry the few straggler models of a batch
  // whose other outcomes have moved on (kernel search), tiny dependent launches that must not queue behind the thousands
  // of CTAs of the next level's batch
  int prio_lo = 0, prio_hi = 0;
  WV_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  const int prio_main = (prio_lo - prio_hi >= 2 && !(flags & 1)) ? prio_hi + 1 : prio_hi;
  WV_CUDA(cudaStreamCreateWithPriority(&eng->stream, cudaStreamNonBlocking, prio_main));
  WV_CUDA(cudaStreamCreateWithPriority(&eng->aux.side, cudaStreamNonBlocking, prio_lo));
  WV_CUDA(cudaEventCreateWithFlags(&eng->aux.ev_panel, cudaEventDisableTiming));
  WV_CUDA(cudaEventCreateWithFlags(&eng->aux.ev_bulk, cudaEventDisableTiming));
  WV_CUDA(cudaMallocHost((void**)&eng->h_count, 4 * sizeof(int)));
  if (const char* v = getenv("WV_BIG_NT")) eng->aux.big_nt = atoi(v) > 1 ? atoi(v) : 2;
  if (const char* v = getenv("WV_CACHE_MAX_GB")) WV_CACHE_MAX_BYTES = (size_t)(atof(v) > 0 ? atof(v) : 0) << 30;
  if (const char* v = getenv("WV_TRTRI_ROWS")) eng->aux.trtri_rows = atoi(v);      // 1: one CTA per row, 2: balanced row pairs
  if (const char* v = getenv("WV_PANEL_TILES")) eng->aux.panel_tiles = atoi(v) > 0 ? atoi(v) : 4;
  if (const char* v = getenv("WV_FEW_MODELS")) eng->aux.few_models = atoi(v) != 0;
  if (const char* v = getenv("WV_CHOL_ALL")) eng->aux.chol_all = atoi(v);
  if (const char* v = getenv("WV_CHOL_ALL_MAX")) eng->aux.chol_all_max = atol(v);
  if (const char* v = getenv("WV_CHOL_ALL_MIN")) eng->aux.chol_all_min = atol(v);
  if (const char* v = getenv("WV_TRTRI_ALL")) eng->aux.trtri_all = atoi(v);
  if (const char* v = getenv("WV_TRTRI_ALL_MAX")) eng->aux.trtri_all_max = atol(v);
  if (const char* v = getenv("WV_CHOL_LAG")) eng->aux.chol_lag = atoi(v) > 0 ? atoi(v) : 640;
  if (const char* v = getenv("WV_PANEL_FUSED")) eng->aux.panel_fused = atoi(v) != 0;
  if (const char* v = getenv("WV_PANEL_CTAS")) eng->aux.panel_ctas = atoi(v) > 0 ? atoi(v) : 148;
  {
    cudaDeviceProp prop;
    WV_CUDA(cudaGetDeviceProperties(&prop, device));
    eng->aux.resident_ctas = 3 * prop.multiProcessorCount;
    if (!getenv("WV_PANEL_CTAS")) eng->aux.panel_ctas = prop.multiProcessorCount;
    eng->aux.trtri_ctas = 4 * prop.multiProcessorCount;
    if (const char* v = getenv("WV_TRTRI_CTAS")) eng->aux.trtri_ctas = atoi(v) > 0 ? atoi(v) : eng->aux.trtri_ctas;
  }
  *out = eng;
  return 0;
}

// Re-home a batch: its later calls run on engine e's stream.  No call on the batch may be in progress; both engines must
// be on the batch's device.  (Every entry point synchronises its stream before it returns, so nothing is in flight.)
extern "C" int wv_batch_set_engine(wv_batch* b, wv_engine* e);

extern "C" void wv_engine_destroy(wv_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  {        // the last engine of the device gives the parked buffers back
    std::lock_guard<std::mutex> lock(e->dc->mu);
    if (--e->dc->engines == 0) wv_cache_flush_locked(e->dc);
  }
  cudaStreamDestroy(e->stream);
  if (e->aux.side) cudaStreamDestroy(e->aux.side);
  if (e->aux.ev_panel) cudaEventDestroy(e->aux.ev_panel);
  if (e->aux.ev_bulk) cudaEventDestroy(e->aux.ev_bulk);
  if (e->h_count) cudaFreeHost(e->h_count);
  delete e;
}

extern "C" void* wv_engine_stream(wv_engine* e) { return e ? (void*)e->stream : nullptr; }

extern "C" int wv_engine_set_large_n_tiles(wv_engine* e, int nt) {
  if (!e) return wv_fail("wv_engine_set_large_n_tiles: null engine");
  if (nt < 2) return wv_fail("wv_engine_set_large_n_tiles: threshold must be >= 2 tiles");
  e->aux.big_nt = nt;
  return 0;
}

static int wv_build_program(const wv_program_desc& d, int D, WvProgram* p) {
  memset(p, 0, sizeof(WvProgram));
  if (d.n_comp < 0 || d.n_comp > WV_MAX_COMP) return wv_fail("program: too many components (max 32)");
  if (d.n_leaves < 0 || d.n_leaves > WV_MAX_LEAVES) return wv_fail("program: too many leaves (max 64)");
  if (d.n_slots <= 0 || d.n_slots > WV_MAX_SLOTS) return wv_fail("program: too many parameter slots (max 64)");
  if (d.noise_slot < 0 || d.noise_slot >= d.n_slots) return wv_fail("program: bad noise_slot");
  if (d.mean_slot >= d.n_slots) return wv_fail("program: bad mean_slot");
  if (d.lik_slot2 < -1 || d.lik_slot2 >= d.n_slots) return wv_fail("program: bad lik_slot2");
  p->n_comp = d.n_comp; p->n_leaves = d.n_leaves; p->n_slots = d.n_slots;
  p->noise_slot = d.noise_slot; p->mean_slot = d.mean_slot; p->lik_slot2 = d.lik_slot2;
  for (int c = 0; c <= d.n_comp; ++c) {
    p->comp_start[c] = d.comp_start[c];
    if (c > 0 && (d.comp_start[c] < d.comp_start[c - 1] || d.comp_start[c] > d.n_leaves))
      return wv_fail("program: comp_start not monotone");
  }
  int nd = 0;
  auto chk = [&](int s) { return s >= -1 && s < d.n_slots; };
  for (int l = 0; l < d.n_leaves; ++l) {
    int dim = d.leaf_dim[l];
    if (dim < 0 || dim >= D) return wv_fail("program: leaf dim out of range");
    int k = 0;
    while (k < nd && p->dims[k] != dim) ++k;
    if (k == nd) {
      if (nd == WV_MAX_DIMS) return wv_fail("program: more than 16 distinct covariate columns");
      p->dims[nd++] = dim;
    }
    WvLeaf& lf = p->leaves[l];
    lf.type = d.leaf_type[l]; lf.dim = k;
    lf.s_var = d.leaf_s_var[l]; lf.s_ls = d.leaf_s_ls[l]; lf.s_aux = d.leaf_s_aux[l];
    lf.degree = d.leaf_degree ? d.leaf_degree[l] : 0;
    if (lf.type < 0 || lf.type > WV_LEAF_EMPTY) return wv_fail("program: unknown leaf type");
    if (!chk(lf.s_var) || !chk(lf.s_ls) || !chk(lf.s_aux)) return wv_fail("program: leaf slot out of range");
    bool need_ls = lf.type <= WV_LEAF_PERIODIC || lf.type == WV_LEAF_POLY;
    if (need_ls && lf.s_ls < 0) return wv_fail("program: leaf needs a lengthscale/offset slot");
    if (lf.type == WV_LEAF_PERIODIC && lf.s_aux < 0) return wv_fail("program: periodic leaf needs a period slot");
    if (lf.type != WV_LEAF_EMPTY && lf.s_var < 0) return wv_fail("program: leaf needs a variance slot");
  }
  p->n_dims = nd;
  int nx = 0;
  for (int s = 0; s < d.n_slots; ++s) {
    WvSlot& sl = p->slots[s];
    sl.transform = d.slot_transform[s]; sl.xindex = d.slot_xindex[s]; sl.prior = d.slot_prior[s];
    sl.fixed = d.slot_fixed[s]; sl.shift = d.slot_shift[s]; sl.pa = d.slot_pa[s]; sl.pb = d.slot_pb[s];
    if (sl.xindex >= 0) nx = sl.xindex + 1 > nx ? sl.xindex + 1 : nx;
  }
  p->n_x = nx;
  return 0;
}

extern "C" int wv_batch_create2(wv_engine* e, const wv_batch_desc* d, int32_t flags, wv_batch** out);
extern "C" int wv_batch_create(wv_engine* e, const wv_batch_desc* d, wv_batch** out) {
  return wv_batch_create2(e, d, 0, out);
}

extern "C" int wv_batch_create2(wv_engine* e, const wv_batch_desc* d, int32_t flags, wv_batch** out) {
  if (!e || !d || !out) return wv_fail("wv_batch_create: null argument");
  if (d->n <= 0 || d->B <= 0 || d->D <= 0 || d->P <= 0 || d->n_programs <= 0)
    return wv_fail("wv_batch_create: n, D, B, P, n_programs must be positive");
  WV_CUDA(cudaSetDevice(e->device));
  wv_batch* b = new wv_batch();
  b->eng = e; b->bytes = 0; b->launches = b->rounds = b->model_evals = 0;
  b->last_x = nullptr; b->d_inner1 = b->d_inner2 = nullptr; b->site_sweeps = 0;
  memset(&b->vgp, 0, sizeof(b->vgp));
  b->d_lbs = nullptr; b->d_lbw = nullptr; b->lb_m_alloc = 0; b->h_count = nullptr;
  WvBatchDev& bd = b->bd;
  bd.n = d->n; bd.D = d->D; bd.B = d->B; bd.P = d->P;
  bd.n8 = (d->n + 1 + 7) / 8 * 8;
  bd.nt = (bd.n8 + WV_NB - 1) / WV_NB;
  bd.npad = bd.nt * WV_NB;
  std::vector<WvProgram> progs(d->n_programs);
  int smax = 1;
  for (int i = 0; i < d->n_programs; ++i) {
    if (wv_build_program(d->programs[i], d->D, &progs[i]) != 0) { delete b; return -1; }
    if (progs[i].n_x > d->P) { delete b; return wv_fail("wv_batch_create: program has more trainable params than P"); }
    smax = progs[i].n_slots > smax ? progs[i].n_slots : smax;
  }
  bd.n_slots_max = smax;
  b->n_programs = d->n_programs;
  for (int i = 0; i < d->B; ++i)
    if (d->prog_id[i] < 0 || d->prog_id[i] >= d->n_programs) { delete b; return wv_fail("wv_batch_create: bad prog_id"); }
  const size_t np = bd.npad, B = d->B;
  const int ntiles = bd.nt * (bd.nt + 1) / 2;
  double *dXt, *dY;
  WvProgram* dprog;
  int* dpid;
#define WV_TRY(x) if ((x) != 0) { wv_batch_destroy(b); return -1; }
  WV_TRY(wv_alloc(b, &dXt, (size_t)d->D * np));
  WV_TRY(wv_alloc(b, &dY, B * np));
  WV_TRY(wv_alloc(b, &dprog, (size_t)d->n_programs));
  WV_TRY(wv_alloc(b, &dpid, B));
  WV_TRY(wv_alloc(b, &bd.A, B * np * np));
  WV_TRY(wv_alloc(b, &bd.Mt, B * np * np));
  WV_TRY(wv_alloc(b, &bd.Dinv, B * bd.nt * WV_NB * WV_NB));
  WV_TRY(wv_alloc(b, &bd.alpha, B * np));
  WV_TRY(wv_alloc(b, &bd.logdet_part, B * bd.nt));
  WV_TRY(wv_alloc(b, &bd.quad, B));
  WV_TRY(wv_alloc(b, &bd.partial, B * ntiles * smax));
  WV_TRY(wv_alloc(b, &bd.chol_fail, B));
  WV_TRY(wv_alloc(b, &bd.step_flag, 2 * B * bd.nt + 1));
  unsigned* dmask;
  WV_TRY(wv_alloc(b, &dmask, B));
  WV_TRY(wv_alloc(b, &b->d_x, B * d->P));
  WV_TRY(wv_alloc(b, &b->d_g, B * d->P));
  WV_TRY(wv_alloc(b, &b->d_f, B));
  WV_TRY(wv_alloc(b, &b->d_lml, B));
  WV_TRY(wv_alloc(b, &b->d_status, B));
  WV_TRY(wv_alloc(b, &b->d_active, B));
  WV_TRY(wv_alloc(b, &b->d_active2, B));
  WV_TRY(wv_alloc(b, &b->d_count, 4));
  WV_TRY(wv_alloc(b, &b->d_task, B));
  WV_TRY(wv_alloc(b, &b->d_nx, B));
  WV_TRY(wv_alloc(b, &b->d_iter, B));
  WV_TRY(wv_alloc(b, &b->d_neval, B));
  WV_TRY(wv_alloc(b, &b->d_st2, B));
#undef WV_TRY
  bd.Xt = dXt; bd.Y = dY; bd.programs = dprog; bd.prog_id = dpid; bd.comp_mask = dmask;
  bd.lik = 0; bd.lik_param = 0.0; bd.jitter = 0.0; bd.site_lam = nullptr; bd.site_eta = nullptr; bd.vgp_extra = nullptr; bd.vgp_dlik = nullptr; bd.vgp_dlik2 = nullptr; bd.lik_param2 = 1.0;
  cudaStream_t st = e->stream;
  // Row order on the device: sorted lexicographically by the categorical columns the programs use (fewest levels
  // first).  The marginal likelihood is invariant under a simultaneous permutation of X rows and y entries; the sort
  // turns the categorical masks into block patterns, so whole warps of the Gram / gradient kernels can skip the
  // transcendental factors of categorical x numeric products where the mask is zero.
  b->perm.resize(d->n);
  for (int i = 0; i < d->n; ++i) b->perm[i] = i;
  b->keep_row_order = (flags & 1) != 0;
  if (!b->keep_row_order) {
    std::set<int> cat;
    for (int p = 0; p < d->n_programs; ++p)
      for (int l = 0; l < d->programs[p].n_leaves; ++l)
        if (d->programs[p].leaf_type[l] == WV_LEAF_CAT) cat.insert(d->programs[p].leaf_dim[l]);
    std::vector<std::pair<int, int>> order;   // (levels, dim)
    for (int k : cat) {
      std::set<long long> lv;
      for (int i = 0; i < d->n; ++i) lv.insert((long long)rint(d->X[(size_t)i * d->D + k]));
      order.push_back({(int)lv.size(), k});
    }
    std::sort(order.begin(), order.end());
    const double* Xh = d->X;
    const int Dh = d->D;
    std::stable_sort(b->perm.begin(), b->perm.end(), [&](int a, int c) {
      for (auto& o : order) {
        double va = rint(Xh[(size_t)a * Dh + o.second]), vc = rint(Xh[(size_t)c * Dh + o.second]);
        if (va != vc) return va < vc;
      }
      return false;
    });
  }
  // X -> column-major, zero padded
  std::vector<double> xt((size_t)d->D * np, 0.0);
  for (int i = 0; i < d->n; ++i)
    for (int k = 0; k < d->D; ++k) xt[(size_t)k * np + i] = d->X[(size_t)b->perm[i] * d->D + k];
  std::vector<double> yp((size_t)B * d->n);
  for (size_t m = 0; m < B; ++m)
    for (int i = 0; i < d->n; ++i) yp[m * d->n + i] = d->Y[m * d->n + b->perm[i]];
  std::vector<int> ident(B);
  for (size_t i = 0; i < B; ++i) ident[i] = (int)i;
  cudaError_t ce = cudaSuccess;
  auto step = [&](cudaError_t r) { if (ce == cudaSuccess) ce = r; };
  step(cudaMemcpyAsync(dXt, xt.data(), xt.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  step(cudaMemcpyAsync(dprog, progs.data(), progs.size() * sizeof(WvProgram), cudaMemcpyHostToDevice, st));
  step(cudaMemcpyAsync(dpid, d->prog_id, B * sizeof(int), cudaMemcpyHostToDevice, st));
  step(cudaMemcpyAsync(b->d_active, ident.data(), B * sizeof(int), cudaMemcpyHostToDevice, st));
  step(cudaMemsetAsync(bd.A, 0, B * np * np * sizeof(double), st));
  step(cudaMemsetAsync(bd.Mt, 0, B * np * np * sizeof(double), st));
  step(cudaMemsetAsync(bd.step_flag, 0, (2 * B * bd.nt + 1) * sizeof(int), st));
  step(cudaMemsetAsync(dmask, 0xff, B * sizeof(unsigned), st));
  step(cudaMemsetAsync(dY, 0, B * np * sizeof(double), st));
  step(cudaMemcpy2DAsync(dY, np * sizeof(double), yp.data(), (size_t)d->n * sizeof(double), (size_t)d->n * sizeof(double), B,
                         cudaMemcpyHostToDevice, st));
  b->h_count = e->h_count;
  if (const char* v = getenv("WV_KINV_TMA"))
    if (atoi(v) != 0 && bd.nt < e->aux.big_nt) b->has_tmap = wv_make_tmap_mt(bd, b->tmap_mt) == 0;
  step(cudaStreamSynchronize(st));
  if (ce != cudaSuccess) {
    wv_batch_destroy(b);
    return wv_fail(std::string("wv_batch_create: ") + cudaGetErrorString(ce));
  }
  *out = b;
  return 0;
}

extern "C" void wv_batch_destroy(wv_batch* b) {
  if (!b) return;
  cudaSetDevice(b->eng->device);
  cudaStreamSynchronize(b->eng->stream);
  b->prof.destroy();
  for (auto& a : b->allocs) wv_cache_put(b->eng, a.first, a.second);
  delete b;
}

// ---------------------------------------------------------------------------------------------
// run-time specialised element-wise kernels (waveome_b200/specialize.py generates the text)
// ---------------------------------------------------------------------------------------------
extern "C" int wv_rtc_check(const char* src, char* log, int log_len) {
  if (!src) return wv_fail("wv_rtc_check: null source");
  std::vector<char> cubin;
  std::string lg;
  const int rc = wv_rtc_compile(src, &cubin, &lg);
  if (log && log_len > 0) { strncpy(log, lg.c_str(), (size_t)log_len - 1); log[log_len - 1] = 0; }
  if (rc != 0) return wv_fail("wv_rtc_check: " + lg);
  return (int)cubin.size();
}

extern "C" void wv_rtc_set_cache(const char* dir) { wv_rtc_set_cache_dir(dir); }

extern "C" int wv_rtc_precompile_text(const char* key, const char* src) {
  if (!key || !src) return wv_fail("wv_rtc_precompile_text: null argument");
  std::string err;
  const int rc = wv_rtc_precompile(key, src, &err);
  if (rc < 0) return wv_fail("wv_rtc_precompile_text: " + err);
  return rc;
}

extern "C" int wv_batch_specialize(wv_batch* b, const char* key, const char* src, const char* gram_name,
                                   const char* grad_name, int32_t gram_smem, int32_t grad_smem) {
  if (!b) return wv_fail("wv_batch_specialize: null batch");
  if (!src) { b->has_spec = false; return 0; }        // back to the interpreter kernels
  if (!key || !gram_name || !grad_name) return wv_fail("wv_batch_specialize: null argument");
  WV_CUDA(cudaSetDevice(b->eng->device));
  WvSpecLaunch sp;
  std::string err;
  if (wv_rtc_get_kernels(key, src, gram_name, grad_name, &sp, &err) != 0) return wv_fail("wv_batch_specialize: " + err);
  if (wv_rtc_exp2_table(b->eng->device, &sp.tab12, &err) != 0) return wv_fail("wv_batch_specialize: " + err);
  sp.gram_smem = gram_smem; sp.grad_smem = grad_smem;
  WV_CUDA(cudaFuncSetAttribute(sp.gram, cudaFuncAttributeMaxDynamicSharedMemorySize, gram_smem));
  WV_CUDA(cudaFuncSetAttribute(sp.grad, cudaFuncAttributeMaxDynamicSharedMemorySize, grad_smem));
  b->spec = sp;
  b->has_spec = true;
  return 0;
}

extern "C" void wv_batch_profile_enable(wv_batch* b, int on) {
  if (!b) return;
  b->prof.enabled = on != 0;
  if (on) { for (int i = 0; i < WV_K_NCLASS; ++i) { b->prof.ms[i] = 0; b->prof.launches[i] = 0; } b->prof.n_ev = 0; }
}

extern "C" int wv_batch_profile_read(wv_batch* b, double* ms, int64_t* launches, int n) {
  if (!b || !ms || !launches) return wv_fail("wv_batch_profile_read: null argument");
  cudaSetDevice(b->eng->device);
  cudaStreamSynchronize(b->eng->stream);
  b->prof.resolve();
  for (int i = 0; i < n && i < WV_K_NCLASS; ++i) { ms[i] = b->prof.ms[i]; launches[i] = b->prof.launches[i]; }
  return WV_K_NCLASS;
}

extern "C" int64_t wv_batch_workspace_bytes(const wv_batch* b) { return b ? b->bytes : 0; }
extern "C" int wv_batch_set_likelihood(wv_batch* b, int32_t kind, double param);

extern "C" int wv_batch_set_y(wv_batch* b, const double* Y) {
  if (!b || !Y) return wv_fail("wv_batch_set_y: null argument");
  WV_CUDA(cudaSetDevice(b->eng->device));
  const WvBatchDev& bd = b->bd;
  std::vector<double> yp((size_t)bd.B * bd.n);
  for (size_t m = 0; m < (size_t)bd.B; ++m)
    for (int i = 0; i < bd.n; ++i) yp[m * bd.n + i] = Y[m * bd.n + b->perm[i]];
  WV_CUDA(cudaMemcpy2DAsync((void*)bd.Y, (size_t)bd.npad * sizeof(double), yp.data(), (size_t)bd.n * sizeof(double),
                            (size_t)bd.n * sizeof(double), bd.B, cudaMemcpyHostToDevice, b->eng->stream));
  WV_CUDA(cudaStreamSynchronize(b->eng->stream));
  if (bd.lik != 0) return wv_batch_set_likelihood2(b, bd.lik, bd.lik_param, bd.lik_param2);     // new counts: restart the sites
  return 0;
}

// ---------------------------------------------------------------------------------------------
// likelihood of the batch: 0 gaussian (default), 1 poisson, 2 negative binomial (param = alpha)
// ---------------------------------------------------------------------------------------------
__global__ void wv_site_init_kernel(int B, int n, int npad, int kind, const double* __restrict__ Y, double* lam,
                                    double* eta, double* lgam) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * npad) return;
  const int r = (int)(i % npad);
  const double y = r < n ? Y[i] : 0.0;
  lam[i] = 1.0;                   // unit-precision pseudo-observations at a link-scale guess of f
  eta[i] = kind == 3 ? (y > 0.5 ? 1.0 : -1.0) : (kind == 4 ? log(fmax(y, 1e-12)) : log(y + 1.0));
  lgam[i] = lgamma(y + 1.0);
}

extern "C" int wv_batch_set_likelihood2(wv_batch* b, int32_t kind, double param, double param2);
extern "C" int wv_batch_set_likelihood(wv_batch* b, int32_t kind, double param) {
  return wv_batch_set_likelihood2(b, kind, param, 1.0);
}

extern "C" int wv_batch_set_likelihood2(wv_batch* b, int32_t kind, double param, double param2) {
  if (!b) return wv_fail("wv_batch_set_likelihood: null batch");
  if (kind < 0 || kind > 5)
    return wv_fail("wv_batch_set_likelihood: kind must be 0 (gaussian), 1 (poisson), 2 (negative binomial), 3 (bernoulli), "
                   "4 (gamma) or 5 (zero-inflated negative binomial)");
  if ((kind == 2 || kind == 4 || kind == 5) && !(param > 0.0))
    return wv_fail("wv_batch_set_likelihood: the likelihood parameter must be positive");
  if (kind == 5 && !(param2 > 0.0)) return wv_fail("wv_batch_set_likelihood: km must be positive");
  WV_CUDA(cudaSetDevice(b->eng->device));
  WvBatchDev& bd = b->bd;
  bd.lik = kind; bd.lik_param = param; bd.lik_param2 = param2;
  if (kind == 0) { bd.site_lam = bd.site_eta = bd.vgp_extra = bd.vgp_dlik = bd.vgp_dlik2 = nullptr; bd.jitter = 0.0; return 0; }
  const size_t B = bd.B, np = bd.npad;
  if (!b->vgp.lam_p) {
    double* lg = nullptr;
    if (wv_alloc(b, &bd.site_lam, B * np) || wv_alloc(b, &bd.site_eta, B * np) || wv_alloc(b, &bd.vgp_extra, B) || wv_alloc(b, &bd.vgp_dlik, B) || wv_alloc(b, &bd.vgp_dlik2, B) ||
        wv_alloc(b, &b->vgp.lam_p, B * np) || wv_alloc(b, &b->vgp.eta_p, B * np) || wv_alloc(b, &b->vgp.lam_t, B * np) ||
        wv_alloc(b, &b->vgp.eta_t, B * np) || wv_alloc(b, &b->vgp.fmean, B * np) || wv_alloc(b, &b->vgp.fvar, B * np) ||
        wv_alloc(b, &lg, B * np) || wv_alloc(b, &b->vgp.F_prev, B) || wv_alloc(b, &b->vgp.rho, B) ||
        wv_alloc(b, &b->vgp.first, B) || wv_alloc(b, &b->vgp.inner_task, B) || wv_alloc(b, &b->vgp.sweeps, B) ||
        wv_alloc(b, &b->vgp.good, B) || wv_alloc(b, &b->vgp.at_bound, B) ||
        wv_alloc(b, &b->d_inner1, B) || wv_alloc(b, &b->d_inner2, B))
      return -1;
    b->vgp.lgam = lg;
  }
  bd.jitter = 1e-6;                 // gpflow.config.default_jitter()
  b->vgp.tol = 1e-8;        // relative move of the sites; the bound is stationary there, so its error is ~tol^2
  b->vgp.soft_tol = 1e-3;
  b->vgp.max_sweeps = 80;
  cudaStream_t st = b->eng->stream;
  WV_CUDA(cudaMemsetAsync(bd.vgp_extra, 0, B * sizeof(double), st));
  WV_CUDA(cudaMemsetAsync(bd.vgp_dlik, 0, B * sizeof(double), st));
  WV_CUDA(cudaMemsetAsync(bd.vgp_dlik2, 0, B * sizeof(double), st));
  WV_CUDA(cudaMemsetAsync(b->vgp.at_bound, 0, B * sizeof(int), st));
  wv_site_init_kernel<<<(unsigned)((B * np + 255) / 256), 256, 0, st>>>((int)B, bd.n, (int)np, kind, bd.Y, bd.site_lam, bd.site_eta,
                                                                       (double*)b->vgp.lgam);
  WV_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// posterior mean and variance of f at the training inputs after the last evaluation (variational path), HOST [B, n]
extern "C" int wv_batch_get_latent(wv_batch* b, double* fmean, double* fvar) {
  if (!b || !fmean || !fvar) return wv_fail("wv_batch_get_latent: null argument");
  if (b->bd.lik == 0) return wv_fail("wv_batch_get_latent: the batch has a Gaussian likelihood (use wv_batch_get_alpha)");
  WV_CUDA(cudaSetDevice(b->eng->device));
  const WvBatchDev& bd = b->bd;
  std::vector<double> t1((size_t)bd.B * bd.npad), t2((size_t)bd.B * bd.npad);
  WV_CUDA(cudaMemcpyAsync(t1.data(), b->vgp.fmean, t1.size() * sizeof(double), cudaMemcpyDeviceToHost, b->eng->stream));
  WV_CUDA(cudaMemcpyAsync(t2.data(), b->vgp.fvar, t2.size() * sizeof(double), cudaMemcpyDeviceToHost, b->eng->stream));
  WV_CUDA(cudaStreamSynchronize(b->eng->stream));
  for (size_t m = 0; m < (size_t)bd.B; ++m)
    for (int i = 0; i < bd.n; ++i) {
      fmean[m * bd.n + b->perm[i]] = t1[m * bd.npad + i];
      fvar[m * bd.n + b->perm[i]] = t2[m * bd.npad + i];
    }
  return 0;
}

extern "C" int wv_batch_set_solo(wv_batch* b, int solo) {
  if (!b) return wv_fail("wv_batch_set_solo: null batch");
  b->solo = solo != 0;
  return 0;
}

extern "C" int wv_batch_set_component_mask(wv_batch* b, const uint32_t* mask) {
  if (!b || !mask) return wv_fail("wv_batch_set_component_mask: null argument");
  WV_CUDA(cudaSetDevice(b->eng->device));
  WV_CUDA(cudaMemcpyAsync((void*)b->bd.comp_mask, mask, (size_t)b->bd.B * sizeof(uint32_t), cudaMemcpyHostToDevice,
                          b->eng->stream));
  WV_CUDA(cudaStreamSynchronize(b->eng->stream));
  return 0;
}

extern "C" void wv_batch_counters(const wv_batch* b, int64_t* launches, int64_t* rounds, int64_t* model_evals) {
  if (!b) return;
  if (launches) *launches = b->launches;
  if (rounds) *rounds = b->rounds;
  if (model_evals) *model_evals = b->model_evals;
}

static int wv_eval_all(wv_batch* b, const double* d_x, double* d_f, double* d_g, double* d_lml, int* d_status,
                       const int* d_active, int n_active) {
  b->eng->aux.epoch += 1;
  b->eng->aux.tmap_mt = b->has_tmap ? b->tmap_mt : nullptr;
  b->eng->aux.solo = b->solo;
  cudaStream_t st = b->eng->stream;
  if (b->bd.lik == 0) {
    int l = wv_enqueue_eval(b->bd, d_active, n_active, d_x, d_f, d_g, d_lml, d_status, st, &b->prof, &b->eng->aux,
                            b->has_spec ? &b->spec : nullptr);
    if (l < 0) return wv_fail(std::string("kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    b->launches += l; b->rounds += 1; b->model_evals += n_active;
    return 0;
  }
  // variational path: site sweeps (one factorisation each) until every listed model has converged, then the gradient
  if (n_active <= 0) return 0;
  int l = wv_enqueue_vgp_begin(b->vgp, d_active, n_active, st);
  const int* cur = d_active;
  int n_in = n_active;
  int* bufs[2] = {b->d_inner1, b->d_inner2};
  int which = 0;
  for (int sweep = 0; sweep < b->vgp.max_sweeps + 2 && n_in > 0; ++sweep) {
    b->eng->aux.epoch += 1;
    WV_CUDA(cudaMemsetAsync(b->bd.chol_fail, 0, sizeof(int) * b->bd.B, st));
    int l1 = wv_enqueue_factor(b->bd, cur, n_in, d_x, st, &b->prof, &b->eng->aux, b->has_spec ? &b->spec : nullptr, false);
    int l2 = l1 < 0 ? -1 : wv_enqueue_site_sweep(b->bd, b->vgp, cur, n_in, d_x, bufs[which], b->d_count + 1, st, &b->prof);
    if (l1 < 0 || l2 < 0) return wv_fail(std::string("kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    l += l1 + l2;
    WV_CUDA(cudaMemcpyAsync(b->h_count + 1, b->d_count + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    WV_CUDA(cudaStreamSynchronize(st));
    b->prof.resolve();
    b->site_sweeps += n_in;
    n_in = b->h_count[1];
    cur = bufs[which];
    which ^= 1;
  }
  WvProfiler none;
  int l3 = wv_enqueue_grad_finalize(b->bd, d_active, n_active, d_x, d_f, d_g, d_lml, d_status, st, &b->prof,
                                    b->has_spec ? &b->spec : nullptr);
  int l4 = l3 < 0 ? -1 : wv_enqueue_vgp_status(b->vgp, d_active, n_active, d_status, st);
  if (l3 < 0 || l4 < 0) return wv_fail(std::string("kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
  b->launches += l + l3 + l4; b->rounds += 1; b->model_evals += n_active;
  return 0;
}

__global__ void wv_iota_kernel(int* a, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

extern "C" int wv_batch_eval_device(wv_batch* b, const double* d_x, double* d_f, double* d_grad, double* d_lml,
                                    int32_t* d_status) {
  if (!b || !d_x || !d_f || !d_grad || !d_lml || !d_status) return wv_fail("wv_batch_eval_device: null argument");
  WV_CUDA(cudaSetDevice(b->eng->device));
  const int B = b->bd.B;
  wv_iota_kernel<<<(B + 255) / 256, 256, 0, b->eng->stream>>>(b->d_active, B);
  b->launches += 1;
  b->last_x = d_x;
  return wv_eval_all(b, d_x, d_f, d_grad, d_lml, d_status, b->d_active, B);
}

extern "C" int wv_batch_eval(wv_batch* b, const double* x, double* f, double* grad, double* lml, int32_t* status) {
  if (!b || !x || !f || !grad || !lml || !status) return wv_fail("wv_batch_eval: null argument");
  WV_CUDA(cudaSetDevice(b->eng->device));
  const size_t B = b->bd.B, P = b->bd.P;
  cudaStream_t st = b->eng->stream;
  WV_CUDA(cudaMemcpyAsync(b->d_x, x, B * P * sizeof(double), cudaMemcpyHostToDevice, st));
  if (wv_batch_eval_device(b, b->d_x, b->d_f, b->d_g, b->d_lml, b->d_status) != 0) return -1;
  WV_CUDA(cudaMemcpyAsync(f, b->d_f, B * sizeof(double), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaMemcpyAsync(grad, b->d_g, B * P * sizeof(double), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaMemcpyAsync(lml, b->d_lml, B * sizeof(double), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaMemcpyAsync(status, b->d_status, B * sizeof(int), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// device-resident L-BFGS-B: one thread per model
// ---------------------------------------------------------------------------------------------
#define WV_LB_CHOLFAIL 7

__global__ void wv_lb_init_kernel(int B, int P, int m, WvLbScalars* sc, double* work, size_t wstride, double* x,
                                  double* g, int* task) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  WvLbState L;
  L.bind(sc + b, x + (size_t)b * P, g + (size_t)b * P, work + (size_t)b * wstride, P, m);
  wv_lb_start(L);
  task[b] = WV_LB_FG;
}

__global__ void wv_lb_step_kernel(const int* __restrict__ active, int n_active, const int* __restrict__ nx_of_model,
                                  int Pstride, int m, WvLbOpts opts, WvLbScalars* sc, double* work, size_t wstride,
                                  double* x, double* g, const double* __restrict__ f, const int* __restrict__ status,
                                  int* task) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_active) return;
  const int b = active[i];
  WvLbState L;
  // the optimiser sees only the model's own trainable parameters (n_x <= Pstride)
  L.bind(sc + b, x + (size_t)b * Pstride, g + (size_t)b * Pstride, work + (size_t)b * wstride, nx_of_model[b], m);
  double fb = f[b];
  if (status[b] & WV_STATUS_CHOL_FAIL) {
    // TensorFlow raises InvalidArgumentError here.  At the start point, or with policy 1, the fit is abandoned;
    // otherwise the trial counts as a non-finite evaluation and the line search backs out of it.
    if (opts.chol_fail_policy == 1 || sc[b].first) {
      task[b] = WV_LB_CHOLFAIL;
      return;
    }
    fb = nan("");
    for (int k = 0; k < L.P; ++k) L.g[k] = fb;
  }
  task[b] = wv_lb_step(L, opts, fb);
}

// The same step with the model's optimiser state staged in shared memory: one WARP per model copies workspace, iterate,
// gradient and scalars in (coalesced), runs the state machine on them with the warp policy of wv_lbfgsb.h (lane 0 owns the
// scalars, independent dot products / columns / right-hand sides go to different lanes), and copies them back.
// The thread-per-model kernel above walks ~9 KB of private state per model through dependent, uncoalesced global loads
// (0.46 ms per round at 2000 models); lane 0 alone on the staged state took 156-170 us per round (round 1).  The
// arithmetic of every value is identical in all three forms, so are the results (tests/test_lbfgs_warp_gpu.py).
__global__ void wv_lb_step_warp_kernel(const int* __restrict__ active, int n_active, const int* __restrict__ nx_of_model,
                                       int Pstride, int m, WvLbOpts opts, WvLbScalars* sc, double* work, size_t wstride,
                                       double* x, double* g, const double* __restrict__ f,
                                       const int* __restrict__ status, int* task) {
  extern __shared__ double wv_lb_sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int i = blockIdx.x * wpc + warp;
  if (i >= n_active) return;
  const int b = active[i];
  const int nsc = (int)(sizeof(WvLbScalars) / sizeof(double));
  double* sw = wv_lb_sm + (size_t)warp * (wstride + 2 * Pstride + nsc);
  double* sx = sw + wstride;
  double* sg = sx + Pstride;
  double* ssc = sg + Pstride;
  double* gw = work + (size_t)b * wstride;
  double* gx = x + (size_t)b * Pstride;
  double* gg = g + (size_t)b * Pstride;
  double* gsc = reinterpret_cast<double*>(sc + b);
  for (size_t k = lane; k < wstride; k += 32) sw[k] = gw[k];
  for (int k = lane; k < Pstride; k += 32) { sx[k] = gx[k]; sg[k] = gg[k]; }
  for (int k = lane; k < nsc; k += 32) ssc[k] = gsc[k];
  __syncwarp();
  {
    WvLbScalars* ls = reinterpret_cast<WvLbScalars*>(ssc);
    WvLbState L;
    L.bind(ls, sx, sg, sw, nx_of_model[b], m);
    double fb = f[b];
    bool run = true;
    if (status[b] & WV_STATUS_CHOL_FAIL) {
      if (opts.chol_fail_policy == 1 || ls->first) {
        if (lane == 0) task[b] = WV_LB_CHOLFAIL;
        run = false;
      } else {
        fb = nan("");
        for (int k = lane; k < L.P; k += 32) L.g[k] = fb;
      }
    }
    __syncwarp();
    if (run) {
      WvExWarp ex;
      ex.ln = lane;
      const int t = wv_lb_step(ex, L, opts, fb);
      if (lane == 0) task[b] = t;
    }
  }
  __syncwarp();
  for (size_t k = lane; k < wstride; k += 32) gw[k] = sw[k];
  for (int k = lane; k < Pstride; k += 32) { gx[k] = sx[k]; gg[k] = sg[k]; }
  for (int k = lane; k < nsc; k += 32) gsc[k] = ssc[k];
}

// ordered compaction of the models that still need an evaluation (single CTA, warp ballots)
__global__ void wv_compact_kernel(const int* __restrict__ task, int B, int* out, int* count, int want_active = 1) {
  __shared__ int warp_tot[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int start = 0; start < B; start += blockDim.x) {
    int b = start + threadIdx.x;
    bool act = b < B && (task[b] == WV_LB_FG) == (want_active != 0);
    unsigned bal = __ballot_sync(0xffffffffu, act);
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

__global__ void wv_nx_kernel(const WvProgram* progs, const int* prog_id, int B, int* nx) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) nx[b] = progs[prog_id[b]].n_x;
}

__global__ void wv_lb_report_kernel(int B, const WvLbScalars* sc, const int* task, const int* eval_status,
                                    int* n_iter, int* n_eval, int* status) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  n_iter[b] = sc[b].nit;
  n_eval[b] = sc[b].neval;
  int st = eval_status[b];
  int t = task[b];
  if (t == WV_LB_MAXITER || t == WV_LB_MAXFUN) st |= WV_STATUS_MAXITER;
  if (t == WV_LB_ABNORMAL) st |= WV_STATUS_LINESEARCH;
  if (t == WV_LB_CHOLFAIL) st |= WV_STATUS_CHOL_FAIL;
  status[b] = st;
}

extern "C" int wv_batch_set_engine(wv_batch* b, wv_engine* e) {
  if (!b || !e) return wv_fail("wv_batch_set_engine: null argument");
  if (e->device != b->eng->device) return wv_fail("wv_batch_set_engine: the engine is on another device");
  WV_CUDA(cudaSetDevice(e->device));
  WV_CUDA(cudaStreamSynchronize(b->eng->stream));
  // the dependency flags in the batch's buffers carry evaluation epochs of the old engine: the new one must stay ahead
  if (e->aux.epoch < b->eng->aux.epoch) e->aux.epoch = b->eng->aux.epoch;
  b->eng = e;
  b->h_count = e->h_count;
  return 0;
}

// The fit in three calls, so that a caller can take the results of the models that are finished while a few stragglers
// keep iterating (the kernel search: a level batch lasts as long as its slowest model, waveome_b200/kernel_search.py):
//   begin   upload the starts, initialise every model's state machine
//   run     rounds of {evaluate the active models, advance their state machines, compact} while more than `min_active`
//           models are active
//   report  objective / LML / status at the current iterates, counters, and which models are finished
extern "C" int wv_batch_fit_lbfgs_begin(wv_batch* b, const double* x, const wv_lbfgs_opts* o) {
  if (!b || !x || !o) return wv_fail("wv_batch_fit_lbfgs_begin: null argument");
  if (o->maxcor < 1 || o->maxcor > WV_LB_MAXCOR) return wv_fail("wv_batch_fit_lbfgs: maxcor must be in [1, 20]");
  if (o->maxls < 1) return wv_fail("wv_batch_fit_lbfgs: maxls must be positive");
  WV_CUDA(cudaSetDevice(b->eng->device));
  const int B = b->bd.B, P = b->bd.P, m = o->maxcor;
  cudaStream_t st = b->eng->stream;
  const size_t wstride = wv_lb_work_doubles(P, m);
  if (b->lb_m_alloc < m) {
    if (wv_alloc(b, &b->d_lbs, (size_t)B) != 0) return -1;
    if (wv_alloc(b, &b->d_lbw, (size_t)B * wstride) != 0) return -1;
    b->lb_m_alloc = m;
  }
  WvLbOpts& opts = b->lb_opts;
  opts.m = m; opts.maxiter = o->maxiter; opts.maxfun = o->maxfun; opts.maxls = o->maxls;
  opts.ftol = o->ftol; opts.pgtol = o->gtol; opts.chol_fail_policy = o->chol_fail_policy; opts.reserved = 0;
  const int tb = 64, gb = (B + tb - 1) / tb;
  WV_CUDA(cudaMemcpyAsync(b->d_x, x, (size_t)B * P * sizeof(double), cudaMemcpyHostToDevice, st));
  wv_nx_kernel<<<gb, tb, 0, st>>>(b->bd.programs, b->bd.prog_id, B, b->d_nx);
  wv_lb_init_kernel<<<gb, tb, 0, st>>>(B, P, m, b->d_lbs, b->d_lbw, wstride, b->d_x, b->d_g, b->d_task);
  wv_iota_kernel<<<(B + 255) / 256, 256, 0, st>>>(b->d_active, B);
  b->launches += 3;
  b->lb_n_active = B;
  b->lb_cur = b->d_active;
  b->lb_nxt = b->d_active2;
  b->lb_guard = 0;
  b->lb_guard_max = (long)o->maxfun + (long)o->maxiter + 1000;
  // the pageable source has been consumed by the time cudaMemcpyAsync returns, but not a pinned one
  WV_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int wv_batch_fit_lbfgs_run(wv_batch* b, int32_t min_active, int32_t* n_active_out) {
  if (!b || b->lb_m_alloc < 1 || !b->lb_cur) return wv_fail("wv_batch_fit_lbfgs_run: no fit in progress");
  WV_CUDA(cudaSetDevice(b->eng->device));
  const int B = b->bd.B, P = b->bd.P, m = b->lb_opts.m;
  cudaStream_t st = b->eng->stream;
  const size_t wstride = wv_lb_work_doubles(P, m);
  const WvLbOpts opts = b->lb_opts;
  const int tb = 64;
  int n_active = b->lb_n_active;
  int* cur = b->lb_cur;
  int* nxt = b->lb_nxt;
  // WV_LB_SERIAL=1: the thread-per-model form of the optimiser step (the tests compare the warp form against it)
  const char* lb_env = getenv("WV_LB_SERIAL");
  const bool lb_serial = lb_env && atoi(lb_env) != 0;
  while (n_active > (min_active > 0 ? min_active : 0)) {
    if (wv_eval_all(b, b->d_x, b->d_f, b->d_g, b->d_lml, b->d_status, cur, n_active) != 0) return -1;
    {
      static_assert(sizeof(WvLbScalars) % sizeof(double) == 0, "WvLbScalars is copied as doubles");
      const size_t per = (wstride + 2 * (size_t)P + sizeof(WvLbScalars) / sizeof(double)) * sizeof(double);
      const int wpc = (int)std::min<size_t>(4, (48 * 1024) / per);     // warps (= models) per CTA within 48 KB
      if (wpc >= 1 && !lb_serial)
        wv_lb_step_warp_kernel<<<(n_active + wpc - 1) / wpc, wpc * 32, wpc * per, st>>>(
            cur, n_active, b->d_nx, P, m, opts, b->d_lbs, b->d_lbw, wstride, b->d_x, b->d_g, b->d_f, b->d_status, b->d_task);
      else      // very long memories: the state does not fit, one thread per model on global memory
        wv_lb_step_kernel<<<(n_active + tb - 1) / tb, tb, 0, st>>>(cur, n_active, b->d_nx, P, m, opts, b->d_lbs, b->d_lbw,
                                                                  wstride, b->d_x, b->d_g, b->d_f, b->d_status, b->d_task);
    }
    wv_compact_kernel<<<1, 1024, 0, st>>>(b->d_task, B, nxt, b->d_count);
    b->prof.mark(WV_K_LBFGS, st);
    b->launches += 2;
    WV_CUDA(cudaMemcpyAsync(b->h_count, b->d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    WV_CUDA(cudaStreamSynchronize(st));
    b->prof.resolve();
    n_active = b->h_count[0];
    int* tmp = cur; cur = nxt; nxt = tmp;
    b->lb_n_active = n_active; b->lb_cur = cur; b->lb_nxt = nxt;
    if (++b->lb_guard > b->lb_guard_max) return wv_fail("wv_batch_fit_lbfgs: iteration guard tripped");
  }
  if (n_active_out) *n_active_out = n_active;
  return 0;
}

extern "C" int wv_batch_fit_lbfgs_report(wv_batch* b, double* x, double* f, double* lml, int32_t* n_iter, int32_t* n_eval,
                                         int32_t* status, int32_t* finished) {
  if (!b || !x || !f || !lml || !n_iter || !n_eval || !status) return wv_fail("wv_batch_fit_lbfgs_report: null argument");
  if (b->lb_m_alloc < 1 || !b->lb_cur) return wv_fail("wv_batch_fit_lbfgs_report: no fit in progress");
  WV_CUDA(cudaSetDevice(b->eng->device));
  const int B = b->bd.B, P = b->bd.P;
  cudaStream_t st = b->eng->stream;
  const int tb = 64, gb = (B + tb - 1) / tb;
  // objective, LML and status at the returned optimum (waveome/model_fitting.py:316 log_posterior_density); the active
  // list is not touched (d_iter doubles as the list of models to evaluate).  f / lml / status of a model that is still
  // iterating are not meaningful.
  int n_list = B;
  if (b->lb_n_active > 0) {     // mid-fit: the finished models only (an evaluation of the others would advance their sites)
    wv_compact_kernel<<<1, 1024, 0, st>>>(b->d_task, B, b->d_iter, b->d_count, 0);
    WV_CUDA(cudaMemcpyAsync(b->h_count, b->d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    WV_CUDA(cudaStreamSynchronize(st));
    n_list = b->h_count[0];
  } else {
    wv_iota_kernel<<<(B + 255) / 256, 256, 0, st>>>(b->d_iter, B);
  }
  b->launches += 1;
  if (n_list > 0 && wv_eval_all(b, b->d_x, b->d_f, b->d_g, b->d_lml, b->d_status, b->d_iter, n_list) != 0) return -1;
  b->last_x = b->d_x;
  wv_lb_report_kernel<<<gb, tb, 0, st>>>(B, b->d_lbs, b->d_task, b->d_status, b->d_iter, b->d_neval, b->d_st2);
  b->launches += 1;
  WV_CUDA(cudaMemcpyAsync(x, b->d_x, (size_t)B * P * sizeof(double), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaMemcpyAsync(f, b->d_f, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaMemcpyAsync(lml, b->d_lml, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaMemcpyAsync(n_iter, b->d_iter, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaMemcpyAsync(n_eval, b->d_neval, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaMemcpyAsync(status, b->d_st2, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (finished) WV_CUDA(cudaMemcpyAsync(finished, b->d_task, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaStreamSynchronize(st));
  if (finished)
    for (int i = 0; i < B; ++i) finished[i] = finished[i] != WV_LB_FG;
  WV_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int wv_batch_fit_lbfgs(wv_batch* b, double* x, const wv_lbfgs_opts* o, double* f, double* lml,
                                  int32_t* n_iter, int32_t* n_eval, int32_t* status) {
  if (!b || !x || !o || !f || !lml || !n_iter || !n_eval || !status) return wv_fail("wv_batch_fit_lbfgs: null argument");
  if (wv_batch_fit_lbfgs_begin(b, x, o) != 0) return -1;
  if (wv_batch_fit_lbfgs_run(b, 0, nullptr) != 0) return -1;
  return wv_batch_fit_lbfgs_report(b, x, f, lml, n_iter, n_eval, status, nullptr);
}

// ---------------------------------------------------------------------------------------------
// device-resident Adam with the reference's schedule (waveome/model_classes.py:344-462, the optimiser kernel_test uses by
// default, waveome/model_search.py:2284-2297): Adam(learning_rate) steps on the unconstrained hyper-parameters, a
// checkpoint every `check_every` steps (loss after the step; parameter snapshot; learning-rate decay
// lr0 * decay^(i / decay_every) every `decay_every` steps), stop when the loss decreased by less than
// `convergence_threshold` between two checkpoints, or -- restoring the last snapshot -- on a NaN checkpoint loss (the
// reference stops there with the NaN values in place; the engine returns the last finite checkpoint) or when a step
// hits a failed factorisation (TensorFlow's InvalidArgumentError there).  The reference alternates each Adam step
// with a natural-gradient step of size gamma on (q_mu, q_sqrt); the engine's objective is the bound already maximised
// over q, i.e. the gamma = 1 limit for a Gaussian likelihood, and the exact inner maximisation for the others.
// One thread per model; the loss after step i is the value of the evaluation that opens step i + 1.
// ---------------------------------------------------------------------------------------------
struct WvAdamState {
  double lr, prev_loss;
  int it, n_loss, done, why;      // why: 0 converged, 1 maxiter, 2 NaN loss, 3 restored after a failed factorisation
};

__global__ void wv_adam_init_kernel(int B, int P, double lr0, WvAdamState* st, double* m, double* v, double* xprev,
                                    const double* __restrict__ x, int* task) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  st[b].lr = lr0; st[b].prev_loss = 0.0; st[b].it = 0; st[b].n_loss = 0; st[b].done = 0; st[b].why = 0;
  for (int k = 0; k < P; ++k) { m[(size_t)b * P + k] = 0.0; v[(size_t)b * P + k] = 0.0; xprev[(size_t)b * P + k] = x[(size_t)b * P + k]; }
  task[b] = WV_LB_FG;
}

__global__ void wv_adam_step_kernel(const int* __restrict__ active, int n_active, const int* __restrict__ nx_of_model,
                                    int P, wv_adam_opts o, WvAdamState* st, double* m, double* v, double* xprev, double* x,
                                    const double* __restrict__ g, const double* __restrict__ f,
                                    const int* __restrict__ status, int* task) {
  const int i_ = blockIdx.x * blockDim.x + threadIdx.x;
  if (i_ >= n_active) return;
  const int b = active[i_];
  WvAdamState s = st[b];
  if (s.done) return;                             // stopped since the last compaction of the active list
  const int nx = nx_of_model[b];
  double* xb = x + (size_t)b * P;
  double* xp = xprev + (size_t)b * P;
  const double fb = f[b];
  bool stop = false;
  if (status[b] & WV_STATUS_CHOL_FAIL) {          // the step that would start here raises in the reference
    for (int k = 0; k < nx; ++k) xb[k] = xp[k];
    s.why = 3; stop = true;
  } else {
    const int i = s.it;                           // steps done so far; fb is the loss after step i - 1
    if (i >= 1 && (i - 1) % o.check_every == 0) {
      if (fb != fb) {                             // NaN loss: back to the last finite checkpoint
        for (int k = 0; k < nx; ++k) xb[k] = xp[k];
        s.why = 2; stop = true;
      } else {
        for (int k = 0; k < nx; ++k) xp[k] = xb[k];
        if ((i - 1) % o.decay_every == 0) s.lr = o.learning_rate * pow(o.decay_rate, (double)(i - 1) / o.decay_every);
        if (s.n_loss >= 1 && s.prev_loss - fb < o.convergence_threshold) { s.why = 0; stop = true; }
        s.prev_loss = fb;
        s.n_loss += 1;
      }
    }
    if (!stop && i >= o.max_iter) { s.why = 1; stop = true; }
    if (!stop) {
      const double t = (double)(i + 1);
      const double alpha = s.lr * sqrt(1.0 - pow(o.beta2, t)) / (1.0 - pow(o.beta1, t));
      const double* gb = g + (size_t)b * P;
      double* mb = m + (size_t)b * P;
      double* vb = v + (size_t)b * P;
      for (int k = 0; k < nx; ++k) {
        const double gk = gb[k];
        mb[k] = o.beta1 * mb[k] + (1.0 - o.beta1) * gk;
        vb[k] = o.beta2 * vb[k] + (1.0 - o.beta2) * gk * gk;
        xb[k] -= alpha * mb[k] / (sqrt(vb[k]) + o.epsilon);
      }
      s.it = i + 1;
    }
  }
  s.done = stop ? 1 : 0;
  st[b] = s;
  task[b] = stop ? WV_LB_CONV_F : WV_LB_FG;
}

__global__ void wv_adam_report_kernel(int B, const WvAdamState* st, const int* eval_status, int* n_iter, int* status) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  n_iter[b] = st[b].it;
  int s = eval_status[b];
  if (st[b].why == 1) s |= WV_STATUS_MAXITER;
  if (st[b].why == 2) s |= WV_STATUS_NONFINITE;
  if (st[b].why == 3) s |= WV_STATUS_RESTORED;
  status[b] = s;
}

extern "C" int wv_batch_fit_adam(wv_batch* b, double* x, const wv_adam_opts* o, double* f, double* lml, int32_t* n_iter,
                                 int32_t* status) {
  if (!b || !x || !o || !f || !lml || !n_iter || !status) return wv_fail("wv_batch_fit_adam: null argument");
  if (o->max_iter < 1 || o->check_every < 1 || o->decay_every < 1) return wv_fail("wv_batch_fit_adam: bad schedule");
  WV_CUDA(cudaSetDevice(b->eng->device));
  const int B = b->bd.B, P = b->bd.P;
  cudaStream_t st = b->eng->stream;
  WvAdamState* d_st = nullptr;
  double *d_m = nullptr, *d_v = nullptr, *d_xp = nullptr;
  if (wv_alloc(b, &d_st, (size_t)B) || wv_alloc(b, &d_m, (size_t)B * P) || wv_alloc(b, &d_v, (size_t)B * P) ||
      wv_alloc(b, &d_xp, (size_t)B * P))
    return -1;
  const int tb = 64, gb = (B + tb - 1) / tb;
  WV_CUDA(cudaMemcpyAsync(b->d_x, x, (size_t)B * P * sizeof(double), cudaMemcpyHostToDevice, st));
  wv_nx_kernel<<<gb, tb, 0, st>>>(b->bd.programs, b->bd.prog_id, B, b->d_nx);
  wv_adam_init_kernel<<<gb, tb, 0, st>>>(B, P, o->learning_rate, d_st, d_m, d_v, d_xp, b->d_x, b->d_task);
  wv_iota_kernel<<<(B + 255) / 256, 256, 0, st>>>(b->d_active, B);
  b->launches += 3;
  int n_active = B;
  int* cur = b->d_active;
  int* nxt = b->d_active2;
  // the active list is compacted (one host sync) at the checkpoints only: between them every model takes the same steps
  for (long round = 0; n_active > 0 && round <= (long)o->max_iter + 1; ++round) {
    if (wv_eval_all(b, b->d_x, b->d_f, b->d_g, b->d_lml, b->d_status, cur, n_active) != 0) return -1;
    wv_adam_step_kernel<<<(n_active + tb - 1) / tb, tb, 0, st>>>(cur, n_active, b->d_nx, P, *o, d_st, d_m, d_v, d_xp, b->d_x,
                                                               b->d_g, b->d_f, b->d_status, b->d_task);
    b->prof.mark(WV_K_LBFGS, st);
    b->launches += 1;
    const bool sync_now = round == 0 || (round - 1) % o->check_every == 0 || b->bd.lik != 0 || round >= o->max_iter;
    if (sync_now) {
      wv_compact_kernel<<<1, 1024, 0, st>>>(b->d_task, B, nxt, b->d_count);
      b->launches += 1;
      WV_CUDA(cudaMemcpyAsync(b->h_count, b->d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
      WV_CUDA(cudaStreamSynchronize(st));
      b->prof.resolve();
      n_active = b->h_count[0];
      int* tmp = cur; cur = nxt; nxt = tmp;
    }
  }
  wv_iota_kernel<<<(B + 255) / 256, 256, 0, st>>>(b->d_active, B);
  b->launches += 1;
  if (wv_eval_all(b, b->d_x, b->d_f, b->d_g, b->d_lml, b->d_status, b->d_active, B) != 0) return -1;
  b->last_x = b->d_x;
  wv_adam_report_kernel<<<gb, tb, 0, st>>>(B, d_st, b->d_status, b->d_iter, b->d_st2);
  b->launches += 1;
  WV_CUDA(cudaMemcpyAsync(x, b->d_x, (size_t)B * P * sizeof(double), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaMemcpyAsync(f, b->d_f, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaMemcpyAsync(lml, b->d_lml, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaMemcpyAsync(n_iter, b->d_iter, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaMemcpyAsync(status, b->d_st2, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaStreamSynchronize(st));
  WV_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// objective (B) at given variational parameters (validation entry point, see wv_elbo_rows_kernel)
// ---------------------------------------------------------------------------------------------
__global__ void wv_fill_kernel(double* p, size_t n, double v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

extern "C" int wv_batch_eval_elbo(wv_batch* b, const double* x, const double* q_mu, const double* q_sqrt, double jitter,
                                  double* elbo, double* f, int32_t* status) {
  if (!b || !x || !q_mu || !q_sqrt || !elbo || !f || !status) return wv_fail("wv_batch_eval_elbo: null argument");
  if (!b->keep_row_order)
    return wv_fail("wv_batch_eval_elbo: the batch must be created with WV_BATCH_KEEP_ROW_ORDER (wv_batch_create2): the "
                   "whitened variational parameters refer to the Cholesky factor in the caller's row order");
  WV_CUDA(cudaSetDevice(b->eng->device));
  WvBatchDev bd = b->bd;                   // local copy: Gram of K + jitter I through the per-row-noise path, 1 / lam = 0
  const size_t B = bd.B, n = bd.n, np = bd.npad;
  const int nblk = (int)((n + 63) / 64);
  cudaStream_t st = b->eng->stream;
  double *d_qmu, *d_qs, *d_part, *d_lp, *d_inf, *d_zero;
  if (wv_alloc(b, &d_qmu, B * n) || wv_alloc(b, &d_qs, B * n * n) || wv_alloc(b, &d_part, B * nblk) || wv_alloc(b, &d_lp, B) ||
      wv_alloc(b, &d_inf, B * np) || wv_alloc(b, &d_zero, B * np))
    return -1;
  WV_CUDA(cudaMemcpyAsync(b->d_x, x, B * bd.P * sizeof(double), cudaMemcpyHostToDevice, st));
  WV_CUDA(cudaMemcpyAsync(d_qmu, q_mu, B * n * sizeof(double), cudaMemcpyHostToDevice, st));
  WV_CUDA(cudaMemcpyAsync(d_qs, q_sqrt, B * n * n * sizeof(double), cudaMemcpyHostToDevice, st));
  wv_fill_kernel<<<(unsigned)((B * np + 255) / 256), 256, 0, st>>>(d_inf, B * np, INFINITY);
  WV_CUDA(cudaMemsetAsync(d_zero, 0, B * np * sizeof(double), st));
  bd.site_lam = d_inf; bd.site_eta = d_zero; bd.jitter = jitter;
  wv_iota_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(b->d_active, (int)B);
  WV_CUDA(cudaMemsetAsync(bd.chol_fail, 0, sizeof(int) * B, st));
  b->eng->aux.epoch += 1;
  const int l1 = wv_enqueue_factor(bd, b->d_active, (int)B, b->d_x, st, &b->prof, &b->eng->aux,
                                   b->has_spec ? &b->spec : nullptr, true);
  const int l2 = l1 < 0 ? -1 : wv_enqueue_elbo(bd, b->d_x, d_qmu, d_qs, d_part, nblk, d_lp, st);
  if (l1 < 0 || l2 < 0) return wv_fail(std::string("kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
  b->launches += l1 + l2 + 2;
  std::vector<double> part(B * nblk), lp(B);
  std::vector<int> fail(B);
  WV_CUDA(cudaMemcpyAsync(part.data(), d_part, part.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaMemcpyAsync(lp.data(), d_lp, B * sizeof(double), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaMemcpyAsync(fail.data(), bd.chol_fail, B * sizeof(int), cudaMemcpyDeviceToHost, st));
  WV_CUDA(cudaStreamSynchronize(st));
  for (size_t m = 0; m < B; ++m) {
    double s = 0.0;
    for (int k = 0; k < nblk; ++k) s += part[m * nblk + k];
    elbo[m] = s;
    f[m] = -(s + lp[m]);
    status[m] = (fail[m] ? WV_STATUS_CHOL_FAIL : 0) | (std::isfinite(f[m]) ? 0 : WV_STATUS_NONFINITE);
  }
  b->last_x = nullptr;            // A holds L, not K^-1: the post-fit getters need a full evaluation first
  return 0;
}

// ---------------------------------------------------------------------------------------------
// post-fit: alpha and posterior means (gpflow GPR.predict_f mean; waveome/utilities.py:614-707 consumes predict_y means)
// ---------------------------------------------------------------------------------------------
extern "C" int wv_batch_get_alpha(wv_batch* b, double* alpha) {
  if (!b || !alpha) return wv_fail("wv_batch_get_alpha: null argument");
  if (!b->last_x) return wv_fail("wv_batch_get_alpha: no evaluation has been run on this batch");
  WV_CUDA(cudaSetDevice(b->eng->device));
  const WvBatchDev& bd = b->bd;
  std::vector<double> tmp((size_t)bd.B * bd.npad);
  WV_CUDA(cudaMemcpyAsync(tmp.data(), bd.alpha, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, b->eng->stream));
  WV_CUDA(cudaStreamSynchronize(b->eng->stream));
  for (size_t m = 0; m < (size_t)bd.B; ++m)
    for (int i = 0; i < bd.n; ++i) alpha[m * bd.n + b->perm[i]] = tmp[m * bd.npad + i];
  return 0;
}

__global__ void wv_diag_gather_kernel(const double* __restrict__ A, int npad, size_t total, double* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t m = i / npad, r = i % npad;
  out[i] = A[m * npad * npad + r * npad + r];
}

extern "C" int wv_batch_get_kinv_diag(wv_batch* b, double* diag) {
  if (!b || !diag) return wv_fail("wv_batch_get_kinv_diag: null argument");
  if (!b->last_x) return wv_fail("wv_batch_get_kinv_diag: no evaluation has been run on this batch");
  WV_CUDA(cudaSetDevice(b->eng->device));
  const WvBatchDev& bd = b->bd;
  const size_t total = (size_t)bd.B * bd.npad;
  std::vector<double> tmp(total);
  // the diagonal of A (= (K + noise)^-1 after the evaluation), gathered into the (free after extract) alpha-sized
  // scratch of the gradient partial sums
  double* d_tmp = bd.partial;       // [B][tiles][slots] >= B * npad doubles is not guaranteed: allocate when short
  bool own = false;
  const size_t have = (size_t)bd.B * (bd.nt * (bd.nt + 1) / 2) * bd.n_slots_max;
  if (have < total) { WV_CUDA(cudaMalloc(&d_tmp, total * sizeof(double))); own = true; }
  wv_diag_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, b->eng->stream>>>(bd.A, bd.npad, total, d_tmp);
  cudaError_t e1 = cudaMemcpyAsync(tmp.data(), d_tmp, total * sizeof(double), cudaMemcpyDeviceToHost, b->eng->stream);
  cudaError_t e2 = cudaStreamSynchronize(b->eng->stream);
  if (own) cudaFree(d_tmp);
  if (e1 != cudaSuccess || e2 != cudaSuccess) return wv_fail("wv_batch_get_kinv_diag: copy failed");
  for (size_t m = 0; m < (size_t)bd.B; ++m)
    for (int i = 0; i < bd.n; ++i) diag[m * bd.n + b->perm[i]] = tmp[m * bd.npad + i];
  return 0;
}

extern "C" int wv_batch_predict_f(wv_batch* b, const double* Xnew, int32_t m, double* mean, double* var);

extern "C" int wv_batch_predict_mean(wv_batch* b, const double* Xnew, int32_t m, double* mean) {
  return wv_batch_predict_f(b, Xnew, m, mean, nullptr);
}

extern "C" int wv_batch_predict_f(wv_batch* b, const double* Xnew, int32_t m, double* mean, double* var) {
  if (!b || !Xnew || !mean) return wv_fail("wv_batch_predict_f: null argument");
  if (m <= 0) return wv_fail("wv_batch_predict_f: m must be positive");
  if (!b->last_x) return wv_fail("wv_batch_predict_f: no evaluation has been run on this batch");
  WV_CUDA(cudaSetDevice(b->eng->device));
  const WvBatchDev& bd = b->bd;
  const int mpad = (m + WV_NB - 1) / WV_NB * WV_NB;
  std::vector<double> xt((size_t)bd.D * mpad, 0.0);
  for (int i = 0; i < m; ++i)
    for (int k = 0; k < bd.D; ++k) xt[(size_t)k * mpad + i] = Xnew[(size_t)i * bd.D + k];
  double *d_xt = nullptr, *d_mean = nullptr, *d_part = nullptr;
  cudaStream_t st = b->eng->stream;
  const size_t Bm = (size_t)bd.B * m;
  const int ntiles = bd.nt * (bd.nt + 1) / 2;
  // workspaces come from (and return to, at wv_batch_destroy) the device's buffer cache like every other buffer of
  // the batch: a cudaMalloc / cudaFree pair per call synchronises the device under the other streams' feet
  if (wv_alloc(b, &d_xt, xt.size()) != 0 || wv_alloc(b, &d_mean, Bm * (var ? 3 : 1)) != 0) return -1;   // mean | prior | var
  if (var && wv_alloc(b, &d_part, Bm * ntiles) != 0) return -1;
  int rc = 0;
  if (cudaMemcpyAsync(d_xt, xt.data(), xt.size() * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = -1;
  if (rc == 0 && wv_enqueue_cross_mean(bd, b->last_x, d_xt, m, mpad, d_mean, st) < 0) rc = -1;
  if (rc == 0 && cudaMemcpyAsync(mean, d_mean, Bm * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = -1;
  if (rc == 0 && var) {
    if (wv_enqueue_cross_var(bd, b->last_x, d_xt, m, mpad, d_part, d_mean + Bm, d_mean + 2 * Bm, st) < 0) rc = -1;
    if (rc == 0 && cudaMemcpyAsync(var, d_mean + 2 * Bm, Bm * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = -1;
    b->launches += 2;
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) rc = -1;
  b->launches += 1;
  if (rc != 0) return wv_fail(std::string("wv_batch_predict_f: ") + cudaGetErrorString(cudaGetLastError()));
  return 0;
}
