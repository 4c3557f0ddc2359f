// Host build of the team-parallel L-BFGS-B core (HostTeam = one thread) for validation against scipy (tests only).
//   g++ -O2 -shared -fPIC -o liblbfgsb_host.so lbfgsb_host.cpp
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../para-ortho-pca-bo_b200/csrc/lbfgsb_core.cuh"

typedef double (*objective_t)(int n, const double *x, double *g, void *user);

struct CallbackFn {
  objective_t fn;
  void *user;
  int n;
  double operator()(lbfgsb::HostTeam &, const double *x, double *g) { return fn(n, x, g, user); }
};

extern "C" int lbfgsb_host_minimize(int n, double *x, const double *lo, const double *hi, objective_t fn, void *user,
                                    int maxiter, double factr, double pgtol, double *f_out, int *nit, int *nfev) {
  std::vector<double> buf(lbfgsb::work_doubles(n) + 16, 0.0);
  lbfgsb::Work w;
  lbfgsb::carve(w, n, buf.data());
  std::memcpy(w.x, x, n * sizeof(double));
  std::memcpy(w.l, lo, n * sizeof(double));
  std::memcpy(w.u, hi, n * sizeof(double));
  lbfgsb::HostTeam tm;
  CallbackFn cb{fn, user, n};
  int flag = 0;
  lbfgsb::Result r = lbfgsb::minimize(tm, w, cb, maxiter, factr, pgtol, &flag);
  std::memcpy(x, w.x, n * sizeof(double));
  *f_out = r.f; *nit = r.nit; *nfev = r.nfev;
  return r.status;
}
