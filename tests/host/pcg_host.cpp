/*
 * TEST HARNESS (not product code): instantiates the host-side control flow of the general preconditioned CG
 * (omegahpetsc_b200/csrc/ohp_pcg.hpp) over plain host arrays, so that the loop that drives the device kernels in
 * ohp_pcg.cu -- Chebyshev recurrence, eigenvalue estimate, null-space projection, convergence test -- can be compared
 * with the oracle on a machine without a GPU.  Built by tests/test_pcg_host.py with g++ into tests/_build/.
 */
#include <vector>

#include "ohp_pcg.hpp"

namespace {
struct HostOps {
  int N;
  const int *rowptr, *colidx;
  const double *vals;
  std::vector<double> v[ohp_pcg::V_COUNT], dinv;
  int64_t global_size() const { return N; }
  int zero(int a) { for (double &t : v[a]) t = 0.0; return 0; }
  int copy(int dst, int src) { v[dst] = v[src]; return 0; }
  int noise(int a) { for (int i = 0; i < N; ++i) v[a][i] = ohp_pcg::hash_noise(i); return 0; }
  int scale(int z, int r, bool jacobi) { for (int i = 0; i < N; ++i) v[z][i] = jacobi ? dinv[i] * v[r][i] : v[r][i]; return 0; }
  int axpy(int y, double a, int x) { for (int i = 0; i < N; ++i) v[y][i] += a * v[x][i]; return 0; }
  int aypx(int y, double a, int x) { for (int i = 0; i < N; ++i) v[y][i] = v[x][i] + a * v[y][i]; return 0; }
  int shift(int z, double s) { for (int i = 0; i < N; ++i) v[z][i] += s; return 0; }
  int dot(int a, int b, double *out) { double s = 0.0; for (int i = 0; i < N; ++i) s += v[a][i] * v[b][i]; *out = s; return 0; }
  int sum(int a, double *out) { double s = 0.0; for (int i = 0; i < N; ++i) s += v[a][i]; *out = s; return 0; }
  int spmv(int x, int y) {
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) s += vals[k] * v[x][colidx[k]];
      v[y][i] = s;
    }
    return 0;
  }
  int cheb_first(int y, int r, double scale) { for (int i = 0; i < N; ++i) v[y][i] = scale * dinv[i] * v[r][i]; return 0; }
  int cheb_iter(int ynew, int ykm1, int yk, int r, int t, double a, double b, double c) {
    spmv(yk, t);
    for (int i = 0; i < N; ++i)
      v[ynew][i] = (ykm1 >= 0 ? a * v[ykm1][i] : 0.0) + b * v[yk][i] + c * (dinv[i] * (v[r][i] - v[t][i]));
    return 0;
  }
};
}  // namespace

extern "C" int pcg_host_solve(int N, const int *rowptr, const int *colidx, const double *vals, const double *b, double *x,
                              double rtol, double abstol, double dtol, int maxit, int pc, int cheb_its, double emin,
                              double emax, int nullspace, int *its, int *reason, double *rnorm, double *eig,
                              double *hist, int hist_cap, int *spmvs) {
  HostOps ops;
  ops.N = N; ops.rowptr = rowptr; ops.colidx = colidx; ops.vals = vals;
  for (auto &vec : ops.v) vec.assign(N, 0.0);
  ops.v[ohp_pcg::V_B].assign(b, b + N);
  ops.dinv.assign(N, 1.0);
  for (int i = 0; i < N; ++i)
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (colidx[k] == i && vals[k] != 0.0) ops.dinv[i] = 1.0 / vals[k];
  ohp_pcg::Options o = {rtol, abstol, dtol, maxit, pc, cheb_its, emin, emax, nullspace != 0, hist, hist_cap};
  ohp_pcg::Result r = {};
  const int rc = ohp_pcg::solve(ops, o, &r);
  for (int i = 0; i < N; ++i) x[i] = ops.v[ohp_pcg::V_X][i];
  *its = r.its; *reason = r.reason; *rnorm = r.rnorm; eig[0] = r.emin; eig[1] = r.emax; *spmvs = r.spmvs;
  return rc;
}

extern "C" double pcg_host_tridiag_max(int n, const double *d, const double *e) { return ohp_pcg::tridiag_lambda_max(n, d, e); }
