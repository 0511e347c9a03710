/*
 * ohp_pcg.hpp -- host-side control flow of the general preconditioned CG: the preconditioners the
 * fused CG kernels of ohp_linalg.cu do not cover (a Chebyshev/Jacobi polynomial) and the constant null space that
 * ex12 attaches to the operator when -bc_type is not dirichlet (ex12.cpp:1534-1538).
 *
 * The loop is written against an `Ops` policy (vectors are small integer handles) so that the SAME control flow that
 * drives the device kernels in ohp_pcg.cu can be instantiated over plain host arrays by the CPU unit test
 * (tests/host/pcg_host.cpp) and compared with the tests' CPU restatement without a GPU.  Nothing here touches memory itself.
 *
 * Restated [PETSc-internal]:
 *   KSPSolve_CG + KSPConvergedDefault (as ohp_ksp_cg: zero initial guess, preconditioned norm, reasons 2/3/-3/-4/-8/-9),
 *   KSP_PCApply -> KSP_RemoveNullSpace (constants projected out of every preconditioned residual),
 *   KSPSolve_Chebyshev as the inner solver of -pc_type ksp -ksp_ksp_type chebyshev -ksp_pc_type jacobi
 *     -ksp_ksp_max_it k (three rotating iterates, zero initial guess, no norms),
 *   KSPChebyshevEstEigSet: lambda_max(D^-1 A) from a 10-step Krylov run on a noise vector (here CG's Lanczos
 *     tridiagonal; PETSc runs GMRES, whose Ritz values are the same for a symmetric operator).
 */
#ifndef OHP_PCG_HPP
#define OHP_PCG_HPP

#include <math.h>
#include <stdint.h>

namespace ohp_pcg {

enum { PC_NONE = 0, PC_JACOBI = 1, PC_CHEBYSHEV = 2 };
enum { kErrInterval = 62 }; /* PETSC_ERR_ARG_WRONG: the Chebyshev interval is empty or not positive */
/* vector handles every Ops implementation provides */
enum { V_X = 0, V_B, V_R, V_Z, V_P, V_W, V_Y0, V_Y1, V_T, V_COUNT };

struct Options {
  double rtol, abstol, dtol;
  int max_it;
  int pc;
  int cheb_its;          /* Chebyshev iterations per application (-ksp_ksp_max_it) */
  double emin, emax;     /* <= 0: derived (see chebyshev_interval) */
  bool nullspace;        /* project the constants out of every preconditioned residual */
  double *hist;          /* host array, may be null */
  int hist_cap;
};
struct Result {
  int its, reason;
  double rnorm, rnorm0, emin, emax;
  int spmvs;             /* operator applications, the eigenvalue estimate included */
};

#define OHP_PCG_TRY(e) do { const int rc__ = (e); if (rc__) return rc__; } while (0)

/* largest eigenvalue of the symmetric tridiagonal (d[0..n), e[0..n-1)): bisection on the Sturm count */
inline double tridiag_lambda_max(int n, const double *d, const double *e) {
  double lo = d[0], hi = d[0];
  for (int i = 0; i < n; ++i) {
    const double r = (i > 0 ? fabs(e[i - 1]) : 0.0) + (i < n - 1 ? fabs(e[i]) : 0.0);
    lo = fmin(lo, d[i] - r);
    hi = fmax(hi, d[i] + r);
  }
  for (int it = 0; it < 200 && hi - lo > 1e-14 * fmax(fabs(hi), fabs(lo)); ++it) {
    const double mid = 0.5 * (lo + hi);
    int below = 0;
    double q = d[0] - mid;
    if (q < 0.0) ++below;
    for (int i = 1; i < n; ++i) {
      if (q == 0.0) q = 1e-300;
      q = d[i] - mid - e[i - 1] * e[i - 1] / q;
      if (q < 0.0) ++below;
    }
    if (below >= n) hi = mid; else lo = mid;
  }
  return 0.5 * (lo + hi);
}

/* lambda_max(D^-1 A): `steps` iterations of Jacobi-preconditioned CG started from the noise vector; uses R, Z, P, W */
template <class Ops> int estimate_emax(Ops &ops, int steps, double *lam, int *spmvs) {
  double td[64], te[64];
  if (steps > 64) steps = 64;
  OHP_PCG_TRY(ops.noise(V_R));
  OHP_PCG_TRY(ops.scale(V_Z, V_R, true));
  OHP_PCG_TRY(ops.copy(V_P, V_Z));
  double beta = 0.0, alpha_old = 1.0, bb_old = 0.0;
  OHP_PCG_TRY(ops.dot(V_Z, V_R, &beta));
  int n = 0;
  for (int j = 0; j < steps && beta > 0.0; ++j) {
    OHP_PCG_TRY(ops.spmv(V_P, V_W));
    ++*spmvs;
    double pw = 0.0;
    OHP_PCG_TRY(ops.dot(V_P, V_W, &pw));
    if (!(pw > 0.0)) break;
    const double alpha = beta / pw;
    td[n] = 1.0 / alpha + (j > 0 ? bb_old / alpha_old : 0.0);
    OHP_PCG_TRY(ops.axpy(V_R, -alpha, V_W));
    OHP_PCG_TRY(ops.scale(V_Z, V_R, true));
    double beta_new = 0.0;
    OHP_PCG_TRY(ops.dot(V_Z, V_R, &beta_new));
    const double bb = beta_new / beta;
    te[n] = sqrt(bb) / alpha;
    ++n;
    OHP_PCG_TRY(ops.aypx(V_P, bb, V_Z));
    beta = beta_new; alpha_old = alpha; bb_old = bb;
  }
  *lam = n ? tridiag_lambda_max(n, td, te) : 1.0;
  return 0;
}

/* the interval the polynomial is built for.  emax <= 0: 1.1 x the estimate (PETSc's default safety factor).
 * emin <= 0: emax / (4 k^2 + 36) -- as a preconditioner for CG (not a multigrid smoother) the polynomial pays most
 * when its interval grows with the degree (measured on the 256^2 box with the CPU harness: total operator applications stay within
 * 5-15 % of plain CG's while the outer iterations, i.e. the global reductions, drop 4-14x); PETSc's smoother default
 * (emin = 0.1 x estimate) is obtained by passing both ends explicitly. */
inline void chebyshev_interval(int k, double est, double *emin, double *emax) {
  if (!(*emax > 0.0)) *emax = 1.1 * est;
  if (!(*emin > 0.0)) *emin = *emax / (4.0 * k * k + 36.0);
}

/* Z = p_k(D^-1 A) D^-1 R, zero initial guess; scratch Y0, Y1, T */
template <class Ops> int chebyshev_apply(Ops &ops, int its, double emin, double emax, int *spmvs) {
  const double scale = 2.0 / (emax + emin), alpha = 1.0 - scale * emin, mu = 1.0 / alpha, omegaprod = 2.0 / alpha;
  double c_km1 = 1.0, c_k = mu;
  int ykm1 = V_Y0, yk = V_Y1, ykp1 = V_Z;
  /* y_0 = 0 is never read or stored: the first step below passes "no y_{k-1}" instead */
  OHP_PCG_TRY(ops.cheb_first(yk, V_R, scale));
  bool km1_is_zero = true;
  for (int k = 1; k < its; ++k) {
    ++*spmvs;
    const double c_kp1 = 2.0 * mu * c_k - c_km1, omega = omegaprod * c_k / c_kp1;
    /* y_{k+1} = (1 - omega) y_{k-1} + omega y_k + scale omega D^-1 (r - A y_k): one operator application and the update
     * (on the device one kernel: the update runs in the SpMV's consumer loop; V_T is scratch for implementations that
     * apply the operator first) */
    OHP_PCG_TRY(ops.cheb_iter(ykp1, km1_is_zero ? -1 : ykm1, yk, V_R, V_T, 1.0 - omega, omega, scale * omega));
    km1_is_zero = false;
    const int tmp = ykm1; ykm1 = yk; yk = ykp1; ykp1 = tmp;
    c_km1 = c_k; c_k = c_kp1;
  }
  if (yk != V_Z) OHP_PCG_TRY(ops.copy(V_Z, yk));
  return 0;
}

template <class Ops> int apply_pc(Ops &ops, const Options &o, double emin, double emax, int *spmvs) {
  if (o.pc == PC_CHEBYSHEV) OHP_PCG_TRY(chebyshev_apply(ops, o.cheb_its < 1 ? 1 : o.cheb_its, emin, emax, spmvs));
  else OHP_PCG_TRY(ops.scale(V_Z, V_R, o.pc == PC_JACOBI));
  if (o.nullspace && ops.global_size() > 0) {
    double s = 0.0;
    OHP_PCG_TRY(ops.sum(V_Z, &s));
    OHP_PCG_TRY(ops.shift(V_Z, -s / (double)ops.global_size()));
  }
  return 0;
}

/* x = 0; solves A x = b.  V_B is read only; the solution is left in V_X. */
template <class Ops> int solve(Ops &ops, const Options &o, Result *res) {
  double emin = o.emin, emax = o.emax;
  int spmvs = 0;
  if (o.pc == PC_CHEBYSHEV) {
    double est = 0.0;
    if (!(emax > 0.0)) OHP_PCG_TRY(estimate_emax(ops, 10, &est, &spmvs));
    chebyshev_interval(o.cheb_its < 1 ? 1 : o.cheb_its, est, &emin, &emax);
    if (!(emin > 0.0) || !(emax > emin)) return kErrInterval; /* 0 < emin < emax or the recurrence has no meaning */
  }
  res->emin = emin; res->emax = emax;
  int reason = 0, i = 0;
  OHP_PCG_TRY(ops.zero(V_X));
  OHP_PCG_TRY(ops.copy(V_R, V_B));
  OHP_PCG_TRY(apply_pc(ops, o, emin, emax, &spmvs));
  double zz = 0.0, beta = 0.0, betaold = 1.0;
  OHP_PCG_TRY(ops.dot(V_Z, V_Z, &zz));
  OHP_PCG_TRY(ops.dot(V_Z, V_R, &beta));
  double dp = sqrt(zz);
  const double rnorm0 = dp, ttol = fmax(o.rtol * rnorm0, o.abstol);
  if (o.hist && o.hist_cap > 0) o.hist[0] = dp;
  if (dp != dp) reason = -9;
  else if (dp <= ttol) reason = (dp < o.abstol) ? 3 : 2;
  while (!reason && i < o.max_it) {
    if (beta == 0.0) { reason = 3; break; }
    if (i == 0) OHP_PCG_TRY(ops.copy(V_P, V_Z));
    else OHP_PCG_TRY(ops.aypx(V_P, beta / betaold, V_Z));
    OHP_PCG_TRY(ops.spmv(V_P, V_W));
    ++spmvs;
    double dpi = 0.0;
    OHP_PCG_TRY(ops.dot(V_P, V_W, &dpi));
    betaold = beta;
    if (!(dpi > 0.0)) { reason = (dpi != dpi) ? -9 : -8; ++i; break; }
    const double a = beta / dpi;
    OHP_PCG_TRY(ops.axpy(V_X, a, V_P));
    OHP_PCG_TRY(ops.axpy(V_R, -a, V_W));
    OHP_PCG_TRY(apply_pc(ops, o, emin, emax, &spmvs));
    OHP_PCG_TRY(ops.dot(V_Z, V_Z, &zz));
    dp = sqrt(zz);
    ++i;
    if (o.hist && i < o.hist_cap) o.hist[i] = dp;
    if (dp != dp) reason = -9;
    else if (dp <= ttol) reason = (dp < o.abstol) ? 3 : 2;
    else if (dp >= o.dtol * rnorm0) reason = -4;
    if (reason) break;
    OHP_PCG_TRY(ops.dot(V_Z, V_R, &beta));
  }
  if (!reason) reason = -3;
  res->its = i; res->reason = reason; res->rnorm = dp; res->rnorm0 = rnorm0; res->spmvs = spmvs;
  return 0;
}

/* the noise vector of the eigenvalue estimate: a hash of the GLOBAL row number, so the estimate (and with it every
 * iterate) does not depend on the partition (PETSc seeds it with PetscRandom, -ksp_chebyshev_esteig_noisy) */
#ifdef __CUDACC__
__host__ __device__
#endif
inline double hash_noise(int64_t i) {
  uint64_t z = (uint64_t)i + 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  z ^= z >> 31;
  return (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

}  // namespace ohp_pcg
#endif
