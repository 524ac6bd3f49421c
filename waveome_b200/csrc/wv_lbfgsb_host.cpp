// Host build of the L-BFGS-B state machine in wv_lbfgsb.h.  TEST INFRASTRUCTURE: lets the CPU test suite
// drive exactly the code the device runs (one thread per model) against SciPy's L-BFGS-B.  It is not used
// by the product path — wv_batch_fit_lbfgs runs the same header on the GPU.
#include <vector>
#include <cstring>
#include "wv_lbfgsb.h"

struct HostLb {
  int P, m;
  WvLbOpts opts;
  WvLbScalars sc;
  std::vector<double> x, g, work;
  WvLbState L;
};

extern "C" {
void* wvh_lb_create(int P, int m, int maxiter, int maxfun, int maxls, double ftol, double pgtol) {
  HostLb* h = new HostLb();
  h->P = P; h->m = m;
  h->opts.m = m; h->opts.maxiter = maxiter; h->opts.maxfun = maxfun; h->opts.maxls = maxls;
  h->opts.ftol = ftol; h->opts.pgtol = pgtol; h->opts.chol_fail_policy = 0; h->opts.reserved = 0;
  h->x.assign(P, 0.0); h->g.assign(P, 0.0); h->work.assign(wv_lb_work_doubles(P, m), 0.0);
  memset(&h->sc, 0, sizeof(h->sc));
  h->L.bind(&h->sc, h->x.data(), h->g.data(), h->work.data(), P, m);
  return h;
}
void wvh_lb_destroy(void* p) { delete static_cast<HostLb*>(p); }
void wvh_lb_start(void* p, const double* x0) {
  HostLb* h = static_cast<HostLb*>(p);
  for (int i = 0; i < h->P; ++i) h->x[i] = x0[i];
  wv_lb_start(h->L);
}
// feed f, g evaluated at the current x; returns the new task (0 = evaluate again at wvh_lb_x)
int wvh_lb_step(void* p, double f, const double* g) {
  HostLb* h = static_cast<HostLb*>(p);
  for (int i = 0; i < h->P; ++i) h->g[i] = g[i];
  return wv_lb_step(h->L, h->opts, f);
}
const double* wvh_lb_x(void* p) { return static_cast<HostLb*>(p)->x.data(); }
double wvh_lb_f(void* p) { return static_cast<HostLb*>(p)->sc.f; }
int wvh_lb_iter(void* p) { return static_cast<HostLb*>(p)->sc.nit; }
int wvh_lb_nfev(void* p) { return static_cast<HostLb*>(p)->sc.neval; }
}
