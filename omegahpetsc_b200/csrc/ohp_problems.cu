/*
 * ohp_problems.cu -- the PetscDS side of the path: ex12's callback sets compiled into the
 * assembly kernels, the problem registry, the FEM evaluators and the Newton driver.
 *
 * Reference: SetupProblem (ex12.cpp:1162-1243) registers (f0,f1) with PetscDSSetResidual and
 * g3 with PetscDSSetJacobian, the exact solution with PetscDSSetExactSolution and the
 * Dirichlet function with DMAddBoundary; DMPlexSetSNESLocalFEM (ex12.cpp:1540) installs the
 * evaluators; SNESSolve (ex12.cpp:1609) runs newtonls + bt over KSP.
 */
#include <math.h>
#include <stdlib.h>
#include <time.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "ohp_fem.cuh"
#include "ohp_internal.h"

namespace {

/* natural boundary data of -bc_type neumann: f0_bd_u (ex12.cpp:179-186), f1_bd_zero (ex12.cpp:188-195), registered for every
 * coefficient variant by SetupProblem (ex12.cpp:1226,1231) */
struct NaturalBd {
  OHP_HD static void f0_bd(OHP_BD_ARGS, double f0[]) {
    f0[0] = 0.0;
    for (int d = 0; d < dim; ++d) f0[0] += -n[d] * 2.0 * x[d];
  }
  OHP_HD static void f1_bd(OHP_BD_ARGS, double f1[]) {
    for (int comp = 0; comp < dim; ++comp) f1[comp] = 0.0;
  }
};

/* exact / Dirichlet data: quadratic_u_2d (ex12.cpp:113-117), quadratic_u_3d (ex12.cpp:408-412) */
struct QuadraticBC : NaturalBd {
  OHP_HD static int bc(int dim, double time, const double x[], int Nc, double *u, void *ctx) {
    if (dim == 3) *u = 2.0 * (x[0] * x[0] + x[1] * x[1] + x[2] * x[2]) / 3.0;
    else *u = x[0] * x[0] + x[1] * x[1];
    return 0;
  }
  OHP_HD static int exact(int dim, double time, const double x[], int Nc, double *u, void *ctx) {
    return bc(dim, time, x, Nc, u, ctx);
  }
};

/* COEFF_NONE: f0_u (ex12.cpp:147-153), f1_u (:198-205), g3_uu (:209-216) */
struct ProblemCoeffNone : QuadraticBC {
  static constexpr bool kJacobianUsesU = false;
  OHP_HD static void f0(OHP_PF_ARGS, double f0[]) { f0[0] = 4.0; }
  OHP_HD static void f1(OHP_PF_ARGS, double f1[]) {
    for (int d = 0; d < dim; ++d) f1[d] = u_x[d];
  }
  OHP_HD static void g3(OHP_PJ_ARGS, double g3[]) {
    for (int d = 0; d < dim; ++d) g3[d * dim + d] = 1.0;
  }
};

/* COEFF_ANALYTIC: f0_analytic_u (ex12.cpp:283-289), f1_analytic_u (:292-299), g3_analytic_uu (:313-320) */
struct ProblemAnalytic : QuadraticBC {
  static constexpr bool kJacobianUsesU = false;
  OHP_HD static void f0(OHP_PF_ARGS, double f0[]) { f0[0] = 6.0 * (x[0] + x[1]); }
  OHP_HD static void f1(OHP_PF_ARGS, double f1[]) {
    for (int d = 0; d < dim; ++d) f1[d] = (x[0] + x[1]) * u_x[d];
  }
  OHP_HD static void g3(OHP_PJ_ARGS, double g3[]) {
    for (int d = 0; d < dim; ++d) g3[d * dim + d] = x[0] + x[1];
  }
};

/* COEFF_NONLINEAR (p-Laplacian, p = 4): f0/f1/g3_analytic_nonlinear_u (ex12.cpp:341-347,350-359,369-383) */
struct ProblemNonlinear : QuadraticBC {
  static constexpr bool kJacobianUsesU = true;
  OHP_HD static void f0(OHP_PF_ARGS, double f0[]) { f0[0] = 16.0 * (x[0] * x[0] + x[1] * x[1]); }
  OHP_HD static void f1(OHP_PF_ARGS, double f1[]) {
    double nu = 0.0;
    for (int d = 0; d < dim; ++d) nu += u_x[d] * u_x[d];
    for (int d = 0; d < dim; ++d) f1[d] = 0.5 * nu * u_x[d];
  }
  OHP_HD static void g3(OHP_PJ_ARGS, double g3[]) {
    double nu = 0.0;
    for (int d = 0; d < dim; ++d) nu += u_x[d] * u_x[d];
    for (int d = 0; d < dim; ++d) {
      g3[d * dim + d] = 0.5 * nu;
      for (int e = 0; e < dim; ++e) g3[d * dim + e] += u_x[d] * u_x[e];
    }
  }
};

/* COEFF_CIRCLE / COEFF_CROSS: f0_circle_u (ex12.cpp:155-166), f0_cross_u (:168-177) with f1_u and g3_uu;
 * Dirichlet and "exact" data circle_u_2d (:127-136), cross_u_2d (:138-145), selected at :1197-1213 */
struct ProblemCircle : NaturalBd {
  static constexpr bool kJacobianUsesU = false;
  OHP_HD static void f0(OHP_PF_ARGS, double f0[]) {
    const double alpha = 500., radius2 = 0.15 * 0.15;
    const double r2 = (x[0] - 0.5) * (x[0] - 0.5) + (x[1] - 0.5) * (x[1] - 0.5);
    const double xi = alpha * (radius2 - r2);
    f0[0] = (-4.0 * alpha - 8.0 * alpha * alpha * r2 * tanh(xi)) * (1.0 / cosh(xi)) * (1.0 / cosh(xi));
  }
  OHP_HD static void f1(OHP_PF_ARGS, double f1[]) { for (int d = 0; d < dim; ++d) f1[d] = u_x[d]; }
  OHP_HD static void g3(OHP_PJ_ARGS, double g3[]) { for (int d = 0; d < dim; ++d) g3[d * dim + d] = 1.0; }
  OHP_HD static int bc(int dim, double time, const double x[], int Nc, double *u, void *ctx) {
    if (dim == 3) return QuadraticBC::bc(dim, time, x, Nc, u, ctx); /* ex12.cpp:1229-1231: 3-D keeps the quadratic */
    const double alpha = 500., radius2 = 0.15 * 0.15;
    const double r2 = (x[0] - 0.5) * (x[0] - 0.5) + (x[1] - 0.5) * (x[1] - 0.5);
    *u = tanh(alpha * (radius2 - r2)) + 1.0;
    return 0;
  }
  OHP_HD static int exact(int dim, double time, const double x[], int Nc, double *u, void *ctx) { return bc(dim, time, x, Nc, u, ctx); }
};
struct ProblemCross : NaturalBd {
  static constexpr bool kJacobianUsesU = false;
  OHP_HD static double cross(const double x[]) {
    const double alpha = 50 * 4;
    const double xy = (x[0] - 0.5) * (x[1] - 0.5);
    return sin(alpha * xy) * (alpha * fabs(xy) < 2 * 3.14159265358979323846 ? 1 : 0.01);
  }
  OHP_HD static void f0(OHP_PF_ARGS, double f0[]) { f0[0] = cross(x); }
  OHP_HD static void f1(OHP_PF_ARGS, double f1[]) { for (int d = 0; d < dim; ++d) f1[d] = u_x[d]; }
  OHP_HD static void g3(OHP_PJ_ARGS, double g3[]) { for (int d = 0; d < dim; ++d) g3[d * dim + d] = 1.0; }
  OHP_HD static int bc(int dim, double time, const double x[], int Nc, double *u, void *ctx) {
    if (dim == 3) return QuadraticBC::bc(dim, time, x, Nc, u, ctx);
    *u = cross(x);
    return 0;
  }
  OHP_HD static int exact(int dim, double time, const double x[], int Nc, double *u, void *ctx) { return bc(dim, time, x, Nc, u, ctx); }
};

struct ProblemEntry {
  ohp_assembly_launcher residual, jacobian, project, l2diff, bd_residual;
};
template <class P> ProblemEntry builtin() {
  return {ohp::launch_residual_any<P>, ohp::launch_jacobian_any<P>, ohp::launch_project_exact_any<P>, ohp::launch_l2_diff_any<P>,
          ohp::launch_bd_residual_any<P>};
}

std::mutex g_reg_mutex;
std::vector<ProblemEntry> &registry() {
  static std::vector<ProblemEntry> r = {builtin<ProblemCoeffNone>(), builtin<ProblemAnalytic>(), builtin<ProblemNonlinear>(),
                                        builtin<ProblemCircle>(), builtin<ProblemCross>()};
  return r;
}

}  // namespace

int ohp_problem_lookup(int id, ohp_assembly_launcher *res, ohp_assembly_launcher *jac, ohp_assembly_launcher *proj,
                       ohp_assembly_launcher *l2) {
  std::lock_guard<std::mutex> lock(g_reg_mutex);
  auto &r = registry();
  if (id < 0 || id >= (int)r.size()) OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "unknown problem id %d", id);
  if (res) *res = r[id].residual;
  if (jac) *jac = r[id].jacobian;
  if (proj) *proj = r[id].project;
  if (l2) *l2 = r[id].l2diff;
  return 0;
}

extern "C" int ohp_problem_register(ohp_assembly_launcher residual, ohp_assembly_launcher jacobian,
                                    ohp_assembly_launcher project_exact, ohp_assembly_launcher l2_diff, int *problem_id) {
  if (!residual || !jacobian || !problem_id) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_problem_register: null launcher");
  std::lock_guard<std::mutex> lock(g_reg_mutex);
  registry().push_back({residual, jacobian, project_exact, l2_diff, nullptr});
  *problem_id = (int)registry().size() - 1;
  return 0;
}

extern "C" int ohp_problem_set_bd_residual(int problem_id, ohp_assembly_launcher bd_residual) {
  std::lock_guard<std::mutex> lock(g_reg_mutex);
  auto &r = registry();
  if (problem_id < 0 || problem_id >= (int)r.size()) OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "ohp_problem_set_bd_residual: unknown problem id %d", problem_id);
  r[problem_id].bd_residual = bd_residual;
  return 0;
}

int ohp_problem_lookup_bd(int id, ohp_assembly_launcher *bd) {
  std::lock_guard<std::mutex> lock(g_reg_mutex);
  auto &r = registry();
  if (id < 0 || id >= (int)r.size()) OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "unknown problem id %d", id);
  *bd = r[id].bd_residual;
  return 0;
}

#define NEED(m, name)                                                                              \
  do {                                                                                             \
    if (!(m) || !(m)->setup_done) OHP_FAIL(OHP_ERR_ORDER, name ": mesh not set up");               \
    OHP_CUDA(cudaSetDevice((m)->ctx->device));                                                     \
  } while (0)

extern "C" int ohp_residual(ohp_mesh m, int problem, ohp_vec u, ohp_vec F) {
  NEED(m, "ohp_residual");
  if (u->n != m->n_owned || F->n != m->n_owned) OHP_FAIL(OHP_ERR_ARG_SIZ, "ohp_residual: vector size mismatch");
  ohp_assembly_launcher fn;
  OHP_TRY(ohp_problem_lookup(problem, &fn, nullptr, nullptr, nullptr));
  if (m->ctx->nranks > 1) OHP_TRY(ohp_halo_exchange(m, u->d)); /* DMGlobalToLocal */
  ohp_assembly_view vw;
  ohp_assembly_view_make(m, u->d, F->d, &vw);
  int rc = fn(&vw);
  m->ctx->launches++;
  if (rc) OHP_FAIL(rc, "ohp_residual: kernel launch failed (%s)", cudaGetErrorString(cudaGetLastError()));
  if (m->bc_type == OHP_BC_NATURAL) { /* DM_BC_NATURAL: the boundary integral of the PetscDSSetBdResidual callbacks (ex12.cpp:1226-1238) */
    ohp_assembly_launcher bd = nullptr;
    OHP_TRY(ohp_problem_lookup_bd(problem, &bd));
    if (!bd) OHP_FAIL(OHP_ERR_SUP, "ohp_residual: the mesh has a natural boundary condition but problem %d registered no boundary residual", problem);
    rc = bd(&vw);
    m->ctx->launches++;
    if (rc) OHP_FAIL(rc, "ohp_residual: boundary kernel launch failed (%s)", cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}

extern "C" int ohp_jacobian(ohp_mesh m, int problem, ohp_vec u, ohp_mat J) {
  NEED(m, "ohp_jacobian");
  if (u->n != m->n_owned || J->mesh != m) OHP_FAIL(OHP_ERR_ARG_SIZ, "ohp_jacobian: size mismatch");
  ohp_assembly_launcher fn;
  OHP_TRY(ohp_problem_lookup(problem, nullptr, &fn, nullptr, nullptr));
  if (m->ctx->nranks > 1) OHP_TRY(ohp_halo_exchange(m, u->d));
  ohp_assembly_view vw;
  ohp_assembly_view_make(m, u->d, J->vals, &vw);
  int rc = fn(&vw);
  m->ctx->launches++;
  if (rc) OHP_FAIL(rc, "ohp_jacobian: kernel launch failed (%s)", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

extern "C" int ohp_project(ohp_mesh m, int problem, int which, ohp_vec u) {
  NEED(m, "ohp_project");
  if (which == 0) return ohp_vec_set(u, 0.0); /* zero(), ex12.cpp:76-80 */
  ohp_assembly_launcher fn;
  OHP_TRY(ohp_problem_lookup(problem, nullptr, nullptr, &fn, nullptr));
  if (!fn) OHP_FAIL(OHP_ERR_SUP, "ohp_project: problem %d has no exact solution registered", problem);
  ohp_assembly_view vw;
  ohp_assembly_view_make(m, nullptr, u->d, &vw);
  int rc = fn(&vw);
  m->ctx->launches++;
  if (rc) OHP_FAIL(rc, "ohp_project: kernel launch failed");
  return 0;
}

extern "C" int ohp_project_function(ohp_mesh m, ohp_assembly_launcher fn, ohp_vec u) {
  NEED(m, "ohp_project_function");
  if (!fn || !u) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_project_function: null argument");
  ohp_assembly_view vw;
  ohp_assembly_view_make(m, nullptr, u->d, &vw);
  int rc = fn(&vw);
  m->ctx->launches++;
  if (rc) OHP_FAIL(rc, "ohp_project_function: kernel launch failed");
  return 0;
}

extern "C" int ohp_l2_error(ohp_mesh m, int problem, ohp_vec u, double *err) {
  NEED(m, "ohp_l2_error");
  ohp_context ctx = m->ctx;
  ohp_assembly_launcher fn = nullptr;
  OHP_TRY(ohp_problem_lookup(problem, nullptr, nullptr, nullptr, &fn)); /* the problem's OWN exact solution */
  if (!fn) OHP_FAIL(OHP_ERR_SUP, "ohp_l2_error: problem %d has no exact solution registered", problem);
  if (ctx->nranks > 1) OHP_TRY(ohp_halo_exchange(m, u->d));
  ohp_assembly_view vw;
  ohp_assembly_view_make(m, u->d, ctx->d_partials, &vw);
  const int grid = std::min(std::max(1, (m->nC + 255) / 256), 1024);
  vw.n_partials = grid;
  int lrc = fn(&vw);
  if (lrc) OHP_FAIL(lrc, "ohp_l2_error: kernel launch failed (%s)", cudaGetErrorString(cudaGetLastError()));
  OHP_CHECK_LAUNCH(ctx);
  std::vector<double> part(grid);
  OHP_CUDA(cudaMemcpyAsync(part.data(), ctx->d_partials, sizeof(double) * grid, cudaMemcpyDeviceToHost, ctx->stream));
  OHP_CUDA(cudaStreamSynchronize(ctx->stream));
  double s = 0.0;
  for (double p : part) s += p;
  if (ctx->nranks > 1) {
    OHP_CUDA(cudaMemcpyAsync(ctx->d_scalars, &s, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    OHP_TRY(ohp_allreduce_sum(ctx, ctx->d_scalars, 1));
    OHP_CUDA(cudaMemcpyAsync(&s, ctx->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    OHP_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  *err = sqrt(s);
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* SNES newtonls + bt                                                                    */
/* ------------------------------------------------------------------------------------ */
extern "C" int ohp_snes_default_options(ohp_snes_options *o) {
  o->rtol = 1e-8; o->abstol = 1e-50; o->stol = 1e-8; o->max_it = 50;
  o->snes_monitor = nullptr; o->ksp_monitor = nullptr; o->monitor_ctx = nullptr;
  return ohp_ksp_default_options(&o->ksp);
}

namespace {
struct PhaseTimer {
  cudaStream_t st;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[3];
  explicit PhaseTimer(cudaStream_t s) : st(s) {}
  int begin(int phase) {
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return 1;
    cudaEventRecord(a, st);
    ev[phase].push_back({a, b});
    return 0;
  }
  void end(int phase) { cudaEventRecord(ev[phase].back().second, st); }
  float total(int phase) {
    float t = 0.f;
    for (auto &p : ev[phase]) {
      float ms = 0.f;
      cudaEventSynchronize(p.second);
      cudaEventElapsedTime(&ms, p.first, p.second);
      t += ms;
    }
    return t;
  }
  ~PhaseTimer() {
    for (auto &v : ev)
      for (auto &p : v) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  }
};
}  // namespace

extern "C" int ohp_snes_solve(ohp_mesh m, int problem, ohp_mat J, ohp_vec u, const ohp_snes_options *o,
                              ohp_snes_result *res) {
  NEED(m, "ohp_snes_solve");
  if (!J || !u || !o || !res) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_snes_solve: null argument");
  ohp_context ctx = m->ctx;
  ohp_vec F = nullptr, y = nullptr, w = nullptr, Fw = nullptr;
  OHP_TRY(ohp_vec_create(m, &F));
  OHP_TRY(ohp_vec_create(m, &y));
  OHP_TRY(ohp_vec_create(m, &w));
  OHP_TRY(ohp_vec_create(m, &Fw));
  PhaseTimer tm(ctx->stream);
  std::vector<double> khist;
  memset(res, 0, sizeof(*res));
  int rc = 0, its = 0, kits = 0, reason = 0;
  double fnorm = 0.0, fnorm0 = 0.0;
  /* OHP_TRACE=1: wall-clock per call (stream drained first), to find host-side gaps */
  static const bool trace = getenv("OHP_TRACE") != nullptr;
  auto now = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
  double t_last = now();
#define SN_TRY(e) do { rc = (e); if (rc) goto done; if (trace) { cudaStreamSynchronize(ctx->stream); const double t__ = now(); \
    fprintf(stderr, "[ohp trace r%d] %8.3f ms  %s\n", ctx->rank, t__ - t_last, #e); t_last = t__; } } while (0)
  tm.begin(0);
  SN_TRY(ohp_residual(m, problem, u, F));
  tm.end(0);
  res->n_residual++;
  SN_TRY(ohp_vec_norm2(F, &fnorm));
  fnorm0 = fnorm;
  if (o->snes_monitor) o->snes_monitor(o->monitor_ctx, 0, fnorm);
  if (fnorm < o->abstol) reason = 2;
  while (!reason && its < o->max_it) {
    tm.begin(1);
    SN_TRY(ohp_jacobian(m, problem, u, J));
    tm.end(1);
    res->n_jacobian++;
    ohp_ksp_result kr;
    ohp_ksp_options ko = o->ksp;
    if (o->ksp_monitor && !ko.history) { /* -ksp_monitor: record the norms on the device, copy them back per solve */
      khist.assign((size_t)std::min<int64_t>((int64_t)ko.max_it + 2, 1 << 24), 0.0);
      ko.history = khist.data(); ko.history_cap = (int32_t)khist.size();
    }
    tm.begin(2);
    SN_TRY(ohp_ksp_cg(J, F, y, &ko, &kr));
    tm.end(2);
    if (o->ksp_monitor) o->ksp_monitor(o->monitor_ctx, its, std::min(kr.its, ko.history_cap - 1), ko.history, kr.reason);
    kits += kr.its;
    res->spmv_ms = kr.spmv_ms; res->spmv_samples = kr.spmv_samples; res->algo = kr.algo;
    if (kr.reason < 0) { reason = -3; break; }
    /* bt line search: accept lambda when 0.5 g^2 <= 0.5 f^2 + lambda * 1e-4 * initslope */
    SN_TRY(ohp_mat_mult(J, y, w));
    double initslope = 0.0;
    SN_TRY(ohp_vec_dot(F, w, &initslope));
    initslope = -initslope;
    if (initslope > 0.0) initslope = -initslope;
    if (initslope == 0.0) initslope = -1.0;
    double lambda = 1.0, gnorm = 0.0;
    for (int bt = 0; bt < 40; ++bt) {
      SN_TRY(ohp_vec_copy(u, w));
      SN_TRY(ohp_vec_axpy(w, -lambda, y));
      tm.begin(0);
      SN_TRY(ohp_residual(m, problem, w, Fw));
      tm.end(0);
      res->n_residual++;
      SN_TRY(ohp_vec_norm2(Fw, &gnorm));
      if (0.5 * gnorm * gnorm <= 0.5 * fnorm * fnorm + lambda * 1e-4 * initslope) break;
      lambda *= 0.5;
    }
    double ynorm = 0.0, xnorm = 0.0;
    SN_TRY(ohp_vec_norm2(y, &ynorm));
    ynorm *= lambda;
    SN_TRY(ohp_vec_copy(w, u));
    SN_TRY(ohp_vec_copy(Fw, F));
    fnorm = gnorm;
    ++its;
    if (o->snes_monitor) o->snes_monitor(o->monitor_ctx, its, fnorm);
    SN_TRY(ohp_vec_norm2(u, &xnorm));
    if (fnorm < o->abstol) reason = 2;
    else if (fnorm <= o->rtol * fnorm0) reason = 3;
    else if (ynorm < o->stol * xnorm) reason = 4;
  }
  if (!reason) reason = -5;
  SN_TRY(ohp_sync(ctx));
  res->its = its; res->ksp_its = kits; res->reason = reason; res->fnorm0 = fnorm0; res->fnorm = fnorm;
  res->ms_residual = tm.total(0); res->ms_jacobian = tm.total(1); res->ms_ksp = tm.total(2);
done:
#undef SN_TRY
  ohp_vec_destroy(F); ohp_vec_destroy(y); ohp_vec_destroy(w); ohp_vec_destroy(Fw);
  return rc;
}
