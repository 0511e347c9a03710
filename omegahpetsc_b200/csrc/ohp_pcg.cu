/*
 * ohp_pcg.cu -- the general preconditioned CG of SURVEY 8(f) rank 4: what the fused CG loops of ohp_linalg.cu do not
 * cover.  ohp_ksp_cg routes here when the preconditioner is the Chebyshev/Jacobi polynomial (-pc_type ksp
 * -ksp_ksp_type chebyshev -ksp_pc_type jacobi -ksp_ksp_max_it k) or when the operator carries the constant null space
 * that ex12 attaches for -bc_type neumann|none (MatNullSpaceCreate(comm, PETSC_TRUE, 0, NULL) + MatSetNullSpace,
 * ex12.cpp:1534-1538).
 *
 * The control flow is ohp_pcg.hpp (host, unit-tested on CPU through tests/host/pcg_host.cpp); this file
 * supplies the device operations it runs on: the library's own SpMV (k_spmv_tma / k_spmv_stream behind
 * ohp_spmv_launch, halo exchange included), its dot product, and the element-wise kernels below.  A polynomial
 * application is k - 1 launches of k_spmv_cheb (ohp_linalg.cu: the SpMV with the Chebyshev update of the row in its
 * consumer loop) without any reduction or host synchronisation; the two dot products of an outer iteration synchronise
 * with the host (they decide the next coefficients), which the degree amortises.
 */
#include <algorithm>

#include "ohp_internal.h"
#include "ohp_pcg.hpp"

namespace {

__global__ void k_pcg_scale(double *__restrict__ z, const double *__restrict__ r, const double *__restrict__ dinv, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) z[i] = dinv ? dinv[i] * r[i] : r[i];
}
__global__ void k_pcg_shift(double *z, int n, double s) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) z[i] += s;
}
__global__ void k_pcg_noise(double *r, int n, long long row_start) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) r[i] = ohp_pcg::hash_noise(row_start + i);
}
/* y_1 = scale D^-1 r */
__global__ void k_pcg_cheb_first(double *__restrict__ y, const double *__restrict__ r, const double *__restrict__ dinv, int n, double scale) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = scale * dinv[i] * r[i];
}
/* y_{k+1} = a y_{k-1} + b y_k + c D^-1 (r - A y_k): 48 bytes per row, one pass */
__global__ void k_pcg_cheb_step(double *__restrict__ ynew, const double *__restrict__ ykm1, const double *__restrict__ yk,
                                const double *__restrict__ r, const double *__restrict__ t, const double *__restrict__ dinv, int n,
                                double a, double b, double c) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    ynew[i] = (ykm1 ? a * ykm1[i] : 0.0) + b * yk[i] + c * (dinv[i] * (r[i] - t[i]));
}

struct DeviceOps {
  ohp_mat A;
  ohp_mesh m;
  ohp_context ctx;
  cudaStream_t st;
  int n, grid;
  ohp_vec_s vec[ohp_pcg::V_COUNT]; /* views: the library's Vec entry points take handles */
  ohp_vec_s ones;
  double *dinv;
  double *owned[ohp_pcg::V_COUNT + 2]; /* what this object allocated */
  int n_alloc;

  DeviceOps() : A(nullptr), m(nullptr), ctx(nullptr), st(nullptr), n(0), grid(1), dinv(nullptr), n_alloc(0) {}
  ~DeviceOps() { for (int i = 0; i < n_alloc; ++i) OHP_FREE(ctx, owned[i]); }

  int alloc(double **p) {
    const size_t len = (size_t)std::max(1, m->n_owned + m->n_ghost);
    OHP_CUDA(OHP_MALLOC(ctx, p, sizeof(double) * len));
    OHP_CUDA(cudaMemsetAsync(*p, 0, sizeof(double) * len, st));
    owned[n_alloc++] = *p;
    return 0;
  }
  int launched() { OHP_CHECK_LAUNCH(ctx); return 0; }

  int64_t global_size() const { return m->n_global; }
  int zero(int a) { return ohp_vec_set(&vec[a], 0.0); }
  int copy(int dst, int src) { return ohp_vec_copy(&vec[src], &vec[dst]); }
  int noise(int a) { k_pcg_noise<<<grid, 256, 0, st>>>(vec[a].d, n, (long long)m->row_start); return launched(); }
  int scale(int z, int r, bool jacobi) { k_pcg_scale<<<grid, 256, 0, st>>>(vec[z].d, vec[r].d, jacobi ? dinv : nullptr, n); return launched(); }
  int axpy(int y, double a, int x) { return ohp_vec_axpy(&vec[y], a, &vec[x]); }
  int aypx(int y, double a, int x) { return ohp_vec_aypx(&vec[y], a, &vec[x]); }
  int shift(int z, double s) { k_pcg_shift<<<grid, 256, 0, st>>>(vec[z].d, n, s); return launched(); }
  int dot(int a, int b, double *out) { return ohp_vec_dot(&vec[a], &vec[b], out); }
  int sum(int a, double *out) { return ohp_vec_dot(&vec[a], &ones, out); }
  int spmv(int x, int y) {
    if (ctx->nranks > 1) OHP_TRY(ohp_halo_exchange(m, vec[x].d));
    return ohp_spmv_launch(A, vec[x].d, vec[y].d);
  }
  int cheb_first(int y, int r, double scale) { k_pcg_cheb_first<<<grid, 256, 0, st>>>(vec[y].d, vec[r].d, dinv, n, scale); return launched(); }
  int cheb_iter(int ynew, int ykm1, int yk, int r, int t, double a, double b, double c) {
    if (ctx->nranks > 1) OHP_TRY(ohp_halo_exchange(m, vec[yk].d));
    const double *km1 = ykm1 >= 0 ? vec[ykm1].d : nullptr;
    int fused = 0;
    OHP_TRY(ohp_spmv_cheb_launch(A, vec[yk].d, km1, vec[r].d, dinv, vec[ynew].d, a, b, c, &fused)); /* update inside the SpMV */
    if (fused) return 0;
    OHP_TRY(ohp_spmv_launch(A, vec[yk].d, vec[t].d));
    k_pcg_cheb_step<<<grid, 256, 0, st>>>(vec[ynew].d, km1, vec[yk].d, vec[r].d, vec[t].d, dinv, n, a, b, c);
    return launched();
  }
};

}  // namespace

int ohp_pcg_solve(ohp_mat A, ohp_vec b, ohp_vec x, const ohp_ksp_options *o, ohp_ksp_result *res) {
  using namespace ohp_pcg;
  ohp_mesh m = A->mesh;
  ohp_context ctx = m->ctx;
  DeviceOps ops;
  ops.A = A; ops.m = m; ops.ctx = ctx; ops.st = ctx->stream; ops.n = m->n_owned;
  ops.grid = (int)std::min<int64_t>(std::max<int64_t>(1, ((int64_t)ops.n + 255) / 256), (int64_t)ctx->num_sms * 8);
  for (int i = 0; i < V_COUNT; ++i) { ops.vec[i].mesh = m; ops.vec[i].n = m->n_owned; ops.vec[i].d = nullptr; }
  ops.vec[V_X].d = x->d;
  ops.vec[V_B].d = b->d;
  for (int i = V_R; i < V_COUNT; ++i) {
    if ((i == V_Y0 || i == V_Y1 || i == V_T) && o->pc != OHP_PC_CHEBYSHEV) continue;
    OHP_TRY(ops.alloc(&ops.vec[i].d));
  }
  ops.ones.mesh = m; ops.ones.n = m->n_owned; ops.ones.d = nullptr;
  if (A->nullspace) {
    OHP_TRY(ops.alloc(&ops.ones.d));
    OHP_TRY(ohp_vec_set(&ops.ones, 1.0));
  }
  if (o->pc != OHP_PC_NONE) {
    const size_t len = (size_t)std::max(1, m->n_owned + m->n_ghost);
    if (!A->dinv) OHP_CUDA(OHP_MALLOC(ctx, &A->dinv, sizeof(double) * len));
    OHP_TRY(ohp_extract_dinv(A));
    ops.dinv = A->dinv;
  }
  Options po;
  po.rtol = o->rtol; po.abstol = o->abstol; po.dtol = o->dtol; po.max_it = o->max_it;
  po.pc = o->pc == OHP_PC_CHEBYSHEV ? PC_CHEBYSHEV : (o->pc == OHP_PC_JACOBI ? PC_JACOBI : PC_NONE);
  po.cheb_its = o->cheb_its > 0 ? o->cheb_its : 8;
  po.emin = o->cheb_emin; po.emax = o->cheb_emax;
  po.nullspace = A->nullspace != 0;
  po.hist = (o->history && o->history_cap > 0) ? o->history : nullptr;
  po.hist_cap = po.hist ? o->history_cap : 0;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  OHP_CUDA(cudaEventCreate(&e0));
  if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0); OHP_FAIL(OHP_ERR_LIB, "ohp_pcg_solve: cudaEventCreate failed"); }
  cudaEventRecord(e0, ctx->stream);
  Result pr = {};
  const int rc = solve(ops, po, &pr);
  cudaEventRecord(e1, ctx->stream);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (rc == kErrInterval)
    OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_ksp_cg: Chebyshev interval [%g, %g] (cheb_emin, cheb_emax; <= 0 means derived) must satisfy 0 < emin < emax",
             po.emin, po.emax);
  if (rc) return rc;
  res->its = pr.its; res->reason = pr.reason; res->rnorm = pr.rnorm; res->rnorm0 = pr.rnorm0;
  res->spmv_ms = 0.f; res->spmv_samples = 0; res->total_ms = ms; res->algo = OHP_ALGO_GENERAL_PCG;
  res->cheb_emin = pr.emin; res->cheb_emax = pr.emax; res->spmvs = pr.spmvs;
  return 0;
}

/* MatSetNullSpace(A, MatNullSpaceCreate(comm, PETSC_TRUE, 0, NULL)) -- ex12.cpp:1534-1538 */
extern "C" int ohp_mat_set_nullspace(ohp_mat A, int has_constant) {
  if (!A) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_mat_set_nullspace: null matrix");
  A->nullspace = has_constant ? 1 : 0;
  return 0;
}

/* MatNullSpaceRemove(nullSpace, u) for the constant null space (ex12.cpp:1667): u -= mean(u), the mean over ALL ranks */
extern "C" int ohp_vec_remove_constant(ohp_vec v) {
  if (!v) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_vec_remove_constant: null vector");
  ohp_mesh m = v->mesh;
  ohp_context ctx = m->ctx;
  OHP_CUDA(cudaSetDevice(ctx->device));
  if (m->n_global <= 0) return 0;
  ohp_vec ones = nullptr;
  OHP_TRY(ohp_vec_create(m, &ones));
  int rc = ohp_vec_set(ones, 1.0);
  double s = 0.0;
  if (!rc) rc = ohp_vec_dot(v, ones, &s);
  ohp_vec_destroy(ones);
  if (rc) return rc;
  const int grid = (int)std::min<int64_t>(std::max<int64_t>(1, ((int64_t)v->n + 255) / 256), (int64_t)ctx->num_sms * 8);
  k_pcg_shift<<<grid, 256, 0, ctx->stream>>>(v->d, v->n, -s / (double)m->n_global);
  OHP_CHECK_LAUNCH(ctx);
  return 0;
}
