/*
 * ohp_patch.cu -- the element -> row scatter map in its cell-parallel form: "patches".
 *
 * The assembly of DMPlexSNESComputeResidualFEM / JacobianFEM (installed by DMPlexSetSNESLocalFEM, ex12.cpp:1540) adds
 * each cell's element vector / matrix into the rows of its vertices.  The first version of this library walked the
 * star of every row and recomputed the cell geometry and quadrature once per (row, cell) incidence: 3x per triangle,
 * 4x per tetrahedron, which made both kernels FP64-issue bound at 0.20 of the HBM roofline (profiles/r01_residual_ncu.txt).
 *
 * A patch is a block of R rows that are close to each other IN SPACE (rows sorted along a Morton curve of their vertex
 * coordinates, so the grouping does not depend on how the mesh happens to be numbered) together with
 *   - the cells of the stars of those rows, each listed ONCE (sorted, so every row still walks its star in ascending
 *     cell order = PETSc's cell-loop order),
 *   - the vertices of those cells, each listed once, with the cells' connectivity rewritten in 16-bit patch-local ids,
 *   - for every (row, cell) incidence the 16-bit patch-local cell number and corner, and the CSR slots of the cell's
 *     corners inside that row.
 * The assembly kernels (include/ohp_fem.cuh: k_residual_patch, k_jacobian_patch) stage a patch's vertices once in shared
 * memory, compute every cell of the patch once into shared memory, and let one thread per row add up its star from
 * there: no atomics, fixed summation order, bitwise reproducible -- and each cell is computed ~1.2x instead of 3-4x.
 * Everything below runs on the device at set-up; the row order of the matrix and of the vectors is untouched (a patch
 * only decides WHICH THREAD BLOCK assembles a row).
 */
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <vector>

#include "ohp_internal.h"

namespace {

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)std::max<int64_t>(1, (n + bs - 1) / bs); }

constexpr int kCandCells = 4096;   /* star entries of one patch (before unique) */
constexpr int kCandVerts = 8192;   /* corners of one patch's cells (before unique) */

template <int DIM> __device__ __forceinline__ unsigned morton_key(const double *x, const double *lo, const double *inv) {
  unsigned q[DIM];
  constexpr int BITS = DIM == 2 ? 16 : 10;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    double t = (x[d] - lo[d]) * inv[d];
    t = fmin(fmax(t, 0.0), 1.0);
    q[d] = (unsigned)(t * (double)((1u << BITS) - 1u));
  }
  unsigned key = 0;
#pragma unroll
  for (int b = 0; b < BITS; ++b)
#pragma unroll
    for (int d = 0; d < DIM; ++d) key |= ((q[d] >> b) & 1u) << (b * DIM + d);
  return key;
}

template <int DIM>
__global__ void k_row_keys(const double *coords, const int32_t *row_vertex, int n_rows, double lo0, double lo1, double lo2, double i0,
                           double i1, double i2, unsigned *key, int32_t *row) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int v = row_vertex[r];
  double x[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) x[d] = coords[(size_t)v * DIM + d];
  const double lo[3] = {lo0, lo1, lo2}, inv[3] = {i0, i1, i2};
  key[r] = morton_key<DIM>(x, lo, inv);
  row[r] = r;
}

__global__ void k_bbox(const double *coords, int64_t n, int dim, double *out /* blocks x 6 */) {
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
    for (int d = 0; d < dim; ++d) { const double x = coords[v * dim + d]; lo[d] = fmin(lo[d], x); hi[d] = fmax(hi[d], x); }
  __shared__ double s[6][8];
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o; o >>= 1) { lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o)); hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o)); }
    if ((threadIdx.x & 31) == 0) { s[d][threadIdx.x >> 5] = lo[d]; s[3 + d][threadIdx.x >> 5] = hi[d]; }
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int d = 0; d < 3; ++d) {
      double a = s[d][0], b = s[3 + d][0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { a = fmin(a, s[d][w]); b = fmax(b, s[3 + d][w]); }
      out[blockIdx.x * 6 + d] = a; out[blockIdx.x * 6 + 3 + d] = b;
    }
}

/* star size of every patch slot (patch order); padding slots have none */
__global__ void k_patch_deg(const int32_t *perm, const int32_t *row_vertex, const int32_t *v2c_ptr, int n_rows, int n_slots, int32_t *deg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  int d = 0;
  if (i < n_rows) { const int v = row_vertex[perm[i]]; d = v2c_ptr[v + 1] - v2c_ptr[v]; }
  deg[i] = d;
}

/* in-place bitonic sort of a[0..P) (P a power of two) by the whole block */
__device__ void block_sort(int32_t *a, int P) {
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const int x = a[i], y = a[ixj];
          const bool up = ((i & k) == 0);
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
}

/* sorted a[0..P) with INT_MAX padding -> distinct values in b[0..count); returns count to every thread */
__device__ int block_unique(const int32_t *a, int P, int32_t *b, int *s_cnt /* blockDim + 1 ints */) {
  const int chunk = (P + blockDim.x - 1) / blockDim.x;
  const int lo = threadIdx.x * chunk, hi = min(lo + chunk, P);
  int c = 0;
  for (int i = lo; i < hi; ++i) c += (a[i] != 0x7fffffff && (i == 0 || a[i] != a[i - 1])) ? 1 : 0;
  s_cnt[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int t = 0; t < (int)blockDim.x; ++t) { const int x = s_cnt[t]; s_cnt[t] = run; run += x; }
    s_cnt[blockDim.x] = run;
  }
  __syncthreads();
  int o = s_cnt[threadIdx.x];
  for (int i = lo; i < hi; ++i)
    if (a[i] != 0x7fffffff && (i == 0 || a[i] != a[i - 1])) b[o++] = a[i];
  __syncthreads();
  return s_cnt[blockDim.x];
}

__device__ __forceinline__ int lower_index(const int32_t *a, int n, int x) {
  int lo = 0, hi = n - 1;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
  return lo;
}

/* One block per patch.  FILL = false: count the patch's distinct cells and vertices.  FILL = true: write its tables. */
template <int NC, bool FILL>
__global__ void k_patch_build(const int32_t *cells, const int32_t *v2c_ptr, const int32_t *v2c, const uint8_t *slot,
                              const int32_t *row_vertex, const int32_t *perm, int n_rows, int R, const int32_t *pstar_ptr,
                              int32_t *ncell, int32_t *nvert, const int32_t *pcell_ptr, const int32_t *pvert_ptr, int32_t *prow,
                              uint16_t *pstar, uint8_t *pslot, uint16_t *pconn, int32_t *pvert, int *err) {
  extern __shared__ int32_t sm[];
  int32_t *ca = sm, *cb = sm + kCandCells, *va = cb + kCandCells, *vb = va + kCandVerts;
  int *s_cnt = reinterpret_cast<int *>(vb + kCandVerts);
  const int p = blockIdx.x, i = p * R + threadIdx.x;
  const bool live = threadIdx.x < R && i < n_rows;
  const int row = live ? perm[i] : -1;
  const int v = live ? row_vertex[row] : 0;
  const int k0 = live ? v2c_ptr[v] : 0, deg = live ? v2c_ptr[v + 1] - k0 : 0;
  const int e_base = pstar_ptr[p * R], e_end = pstar_ptr[min(p * R + R, (int)gridDim.x * R)];
  const int E = e_end - e_base;
  if (E > kCandCells) { if (threadIdx.x == 0) atomicExch(err, 1); return; }
  int PE = 32;
  while (PE < E) PE <<= 1;
  for (int q = threadIdx.x; q < PE; q += blockDim.x) ca[q] = 0x7fffffff;
  __syncthreads();
  const int my = threadIdx.x < R ? pstar_ptr[i] - e_base : 0;
  for (int q = 0; q < deg; ++q) ca[my + q] = v2c[k0 + q] >> 2;
  __syncthreads();
  block_sort(ca, PE);
  const int nu = block_unique(ca, PE, cb, s_cnt);
  if (nu * NC > kCandVerts || nu >= 16384) { if (threadIdx.x == 0) atomicExch(err, 2); return; }
  int PV = 32;
  while (PV < nu * NC) PV <<= 1;
  for (int q = threadIdx.x; q < PV; q += blockDim.x) va[q] = q < nu * NC ? cells[(size_t)cb[q / NC] * NC + (q % NC)] : 0x7fffffff;
  __syncthreads();
  block_sort(va, PV);
  const int nv = block_unique(va, PV, vb, s_cnt);
  if (nv > 65535) { if (threadIdx.x == 0) atomicExch(err, 3); return; }
  if (!FILL) {
    if (threadIdx.x == 0) { ncell[p] = nu; nvert[p] = nv; }
    return;
  }
  if (threadIdx.x < R) prow[i] = row;
  /* (row, cell) incidences in patch-local terms, the row's own CSR slots next to them */
  for (int q = 0; q < deg; ++q) {
    const int e = v2c[k0 + q];
    const int lc = lower_index(cb, nu, e >> 2);
    pstar[e_base + my + q] = (uint16_t)((lc << 2) | (e & 3));
#pragma unroll
    for (int jj = 0; jj < NC; ++jj) pslot[(size_t)(e_base + my + q) * NC + jj] = slot[(size_t)(k0 + q) * NC + jj];
  }
  const int c0 = pcell_ptr[p], w0 = pvert_ptr[p];
  for (int q = threadIdx.x; q < nu * NC; q += blockDim.x)
    pconn[(size_t)c0 * NC + q] = (uint16_t)lower_index(vb, nv, cells[(size_t)cb[q / NC] * NC + (q % NC)]);
  for (int q = threadIdx.x; q < nv; q += blockDim.x) pvert[w0 + q] = vb[q];
}

}  // namespace

void ohp_patch_plan_free(ohp_context ctx, ohp_patch_plan *pl) {
  if (!pl) return;
  OHP_FREE(ctx, pl->prow); OHP_FREE(ctx, pl->pstar_ptr); OHP_FREE(ctx, pl->pcell_ptr); OHP_FREE(ctx, pl->pvert_ptr); OHP_FREE(ctx, pl->pvert);
  OHP_FREE(ctx, pl->pstar); OHP_FREE(ctx, pl->pconn); OHP_FREE(ctx, pl->pslot);
  delete pl;
}

/* called at the end of ohp_mesh_setup; on any limit it leaves m->patch == NULL and the row kernels are used */
int ohp_patch_plan_build(ohp_mesh m) {
  /* OPT-IN (OHP_ASM=patch).  Measured in round 2 on one B200 (profiles/r02_assembly_patch_vs_rows.txt): on the 16.77 M-triangle
   * box the patch residual takes 0.314 ms against 0.346 ms for the row kernel and the patch Jacobian 0.699 ms against 0.628 ms;
   * on the 7.99 M-tet box 0.325 vs 0.527 ms and 2.34 vs 0.72 ms -- computing every cell once removes 2/3 of the FP64 work but the
   * three block-synchronised phases with 128 threads, the 16-bit table reads and the strided shared-memory row accumulators cost
   * as much as it saves, and building the tables (block-wide bitonic sorts per patch) adds 0.25 s (2-D) to 0.5 s (3-D) to
   * ohp_mesh_setup, which the end-to-end metric pays.  So the row kernels stay the default. */
  static const char *on = getenv("OHP_ASM");
  if (!(on && on[0] == 'p')) return 0;
  if (m->n_owned == 0 || m->nC == 0) return 0;
  ohp_context ctx = m->ctx;
  cudaStream_t st = ctx->stream;
  const int nc = m->nc, N = m->n_owned, BS = 256;
  const int R = m->dim == 2 ? 128 : 64;
  const int n_patches = (N + R - 1) / R, n_slots = n_patches * R;
  int rc = 0;
  ohp_patch_plan *pl = new ohp_patch_plan();
  memset(pl, 0, sizeof(*pl));
  pl->R = R; pl->n_patches = n_patches;
  unsigned *key_a = nullptr, *key_b = nullptr;
  int32_t *row_a = nullptr, *perm = nullptr, *deg = nullptr, *ncell = nullptr, *nvert = nullptr;
  double *d_box = nullptr;
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  int64_t total_star = 0, total_cell = 0, total_vert = 0;
  const int smem = (int)sizeof(int32_t) * (2 * kCandCells + 2 * kCandVerts) + (int)sizeof(int) * (BS + 2);
#define PL_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { ohp_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); rc = OHP_ERR_LIB; goto cleanup; } } while (0)
#define PL_TRY(expr) do { rc = (expr); if (rc) goto cleanup; } while (0)
  {
    /* 1. rows along a Morton curve of their vertices */
    const int nb = 256;
    std::vector<double> box((size_t)nb * 6);
    PL_CUDA(OHP_MALLOC(ctx, &d_box, sizeof(double) * nb * 6));
    k_bbox<<<nb, 256, 0, st>>>(m->coords, m->nV, m->dim, d_box);
    ctx->launches++;
    PL_CUDA(cudaMemcpyAsync(box.data(), d_box, sizeof(double) * nb * 6, cudaMemcpyDeviceToHost, st));
    PL_CUDA(cudaStreamSynchronize(st));
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, inv[3] = {0, 0, 0};
    for (int b = 0; b < nb; ++b) for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], box[b * 6 + d]); hi[d] = std::max(hi[d], box[b * 6 + 3 + d]); }
    for (int d = 0; d < m->dim; ++d) inv[d] = hi[d] > lo[d] ? 1.0 / (hi[d] - lo[d]) : 0.0;
    PL_CUDA(OHP_MALLOC(ctx, &key_a, sizeof(unsigned) * N)); PL_CUDA(OHP_MALLOC(ctx, &key_b, sizeof(unsigned) * N));
    PL_CUDA(OHP_MALLOC(ctx, &row_a, sizeof(int32_t) * N)); PL_CUDA(OHP_MALLOC(ctx, &perm, sizeof(int32_t) * N));
    if (m->dim == 2) k_row_keys<2><<<nblk(N, BS), BS, 0, st>>>(m->coords, m->row_vertex, N, lo[0], lo[1], lo[2], inv[0], inv[1], inv[2], key_a, row_a);
    else k_row_keys<3><<<nblk(N, BS), BS, 0, st>>>(m->coords, m->row_vertex, N, lo[0], lo[1], lo[2], inv[0], inv[1], inv[2], key_a, row_a);
    ctx->launches++;
    PL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_a, key_b, row_a, perm, N, 0, 32, st));
    PL_CUDA(OHP_MALLOC(ctx, &tmp, tmp_bytes));
    PL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_a, key_b, row_a, perm, N, 0, 32, st)); /* stable: ties keep row order */
    ctx->launches++;
  }
  /* 2. star sizes in patch order -> offsets of the incidence tables */
  PL_CUDA(OHP_MALLOC(ctx, &deg, sizeof(int32_t) * (n_slots + 1)));
  PL_CUDA(OHP_MALLOC(ctx, &pl->pstar_ptr, sizeof(int32_t) * (n_slots + 1)));
  k_patch_deg<<<nblk(n_slots, BS), BS, 0, st>>>(perm, m->row_vertex, m->v2c_ptr, N, n_slots, deg);
  ctx->launches++;
  PL_TRY(ohp_scan_exclusive_i32(ctx, deg, pl->pstar_ptr, n_slots, &total_star));
  /* 3. count pass, offsets, fill pass */
  PL_CUDA(OHP_MALLOC(ctx, &ncell, sizeof(int32_t) * (n_patches + 1))); PL_CUDA(OHP_MALLOC(ctx, &nvert, sizeof(int32_t) * (n_patches + 1)));
  PL_CUDA(OHP_MALLOC(ctx, &pl->pcell_ptr, sizeof(int32_t) * (n_patches + 1))); PL_CUDA(OHP_MALLOC(ctx, &pl->pvert_ptr, sizeof(int32_t) * (n_patches + 1)));
  if (nc == 3) {
    PL_CUDA(cudaFuncSetAttribute(k_patch_build<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PL_CUDA(cudaFuncSetAttribute(k_patch_build<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_patch_build<3, false><<<n_patches, BS, smem, st>>>(m->cells, m->v2c_ptr, m->v2c, m->slot, m->row_vertex, perm, N, R, pl->pstar_ptr, ncell, nvert,
                                                       nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ctx->d_flag);
  } else {
    PL_CUDA(cudaFuncSetAttribute(k_patch_build<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PL_CUDA(cudaFuncSetAttribute(k_patch_build<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_patch_build<4, false><<<n_patches, BS, smem, st>>>(m->cells, m->v2c_ptr, m->v2c, m->slot, m->row_vertex, perm, N, R, pl->pstar_ptr, ncell, nvert,
                                                       nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ctx->d_flag);
  }
  ctx->launches++;
  PL_CUDA(cudaGetLastError());
  {
    int flag = 0;
    PL_CUDA(cudaMemcpyAsync(&flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    PL_CUDA(cudaStreamSynchronize(st));
    if (flag) { /* a patch exceeds the staging limits (very high vertex degrees): keep the row kernels */
      cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), st);
      rc = -1;
      goto cleanup;
    }
  }
  PL_TRY(ohp_scan_exclusive_i32(ctx, ncell, pl->pcell_ptr, n_patches, &total_cell));
  PL_TRY(ohp_scan_exclusive_i32(ctx, nvert, pl->pvert_ptr, n_patches, &total_vert));
  {
    std::vector<int32_t> hc(n_patches), hv(n_patches), hs(n_slots + 1);
    PL_CUDA(cudaMemcpyAsync(hc.data(), ncell, sizeof(int32_t) * n_patches, cudaMemcpyDeviceToHost, st));
    PL_CUDA(cudaMemcpyAsync(hv.data(), nvert, sizeof(int32_t) * n_patches, cudaMemcpyDeviceToHost, st));
    PL_CUDA(cudaStreamSynchronize(st));
    for (int p = 0; p < n_patches; ++p) { pl->max_cells = std::max(pl->max_cells, hc[p]); pl->max_verts = std::max(pl->max_verts, hv[p]); }
  }
  PL_CUDA(OHP_MALLOC(ctx, &pl->prow, sizeof(int32_t) * n_slots));
  PL_CUDA(OHP_MALLOC(ctx, &pl->pstar, sizeof(uint16_t) * std::max<int64_t>(1, total_star)));
  PL_CUDA(OHP_MALLOC(ctx, &pl->pslot, (size_t)std::max<int64_t>(1, total_star) * nc));
  PL_CUDA(OHP_MALLOC(ctx, &pl->pconn, sizeof(uint16_t) * std::max<int64_t>(1, total_cell) * nc));
  PL_CUDA(OHP_MALLOC(ctx, &pl->pvert, sizeof(int32_t) * std::max<int64_t>(1, total_vert)));
  if (nc == 3) k_patch_build<3, true><<<n_patches, BS, smem, st>>>(m->cells, m->v2c_ptr, m->v2c, m->slot, m->row_vertex, perm, N, R, pl->pstar_ptr, nullptr, nullptr,
                                                                pl->pcell_ptr, pl->pvert_ptr, pl->prow, pl->pstar, pl->pslot, pl->pconn, pl->pvert, ctx->d_flag);
  else k_patch_build<4, true><<<n_patches, BS, smem, st>>>(m->cells, m->v2c_ptr, m->v2c, m->slot, m->row_vertex, perm, N, R, pl->pstar_ptr, nullptr, nullptr,
                                                         pl->pcell_ptr, pl->pvert_ptr, pl->prow, pl->pstar, pl->pslot, pl->pconn, pl->pvert, ctx->d_flag);
  ctx->launches++;
  PL_CUDA(cudaGetLastError());
  PL_CUDA(cudaStreamSynchronize(st));
  pl->total_star = total_star; pl->total_cells = total_cell; pl->total_verts = total_vert;
cleanup:
  OHP_FREE(ctx, key_a); OHP_FREE(ctx, key_b); OHP_FREE(ctx, row_a); OHP_FREE(ctx, perm); OHP_FREE(ctx, deg); OHP_FREE(ctx, ncell);
  OHP_FREE(ctx, nvert); OHP_FREE(ctx, d_box); OHP_FREE(ctx, tmp);
#undef PL_CUDA
#undef PL_TRY
  if (rc) {
    ohp_patch_plan_free(ctx, pl);
    return rc > 0 ? rc : 0; /* rc < 0: limits exceeded, not an error */
  }
  m->patch = pl;
  return 0;
}
