/*
 * ohp_core.cu -- context, DM construction and the device-side section / CSR build.
 *
 * Replaces, on the GPU, what ex12 delegates to PETSc between mesh creation and the solve
 * (file:line are reference call sites; the algorithms behind them are PETSc-internal):
 *   DMPlexInterpolate + support-size-1 boundary marking      ex12.cpp:958-1033
 *   PetscSection (1 dof/vertex, "marker" vertices constrained) ex12.cpp:1309-1321,1237-1239
 *   DMCreateGlobalVector numbering                            ex12.cpp:1505
 *   DMCreateMatrix preallocation (CSR with explicit zeros)    ex12.cpp:1508
 * The Omega_h primitives the reference adapter uses for the same kind of work
 * (each_eq_to / offset_scan / parallel_for, ex12.cpp:587-588,616,629-647) map to the
 * mask + scan + compaction kernels below.
 */
#include <string.h>

#include <algorithm>

#include "ohp_internal.h"

/* ------------------------------------------------------------------------------------ */
/* errors                                                                                */
/* ------------------------------------------------------------------------------------ */
static thread_local char g_err[1024] = "";

void ohp_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char *ohp_last_error(void) { return g_err; }
extern "C" int ohp_version(void) { return OHP_VERSION; }

/* ------------------------------------------------------------------------------------ */
/* context                                                                               */
/* ------------------------------------------------------------------------------------ */
extern "C" int ohp_init(int device, ohp_context *out) {
  if (!out) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_init: null output");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    OHP_FAIL(OHP_ERR_GPU, "ohp_init: no CUDA device (%s); this library has no CPU path",
             e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "ohp_init: device %d of %d", device, ndev);
  OHP_CUDA(cudaSetDevice(device));
  ohp_context c = new ohp_context_s();
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->rank = 0;
  c->nranks = 1;
  cudaDeviceProp prop;
  OHP_CUDA(cudaGetDeviceProperties(&prop, device));
  c->num_sms = prop.multiProcessorCount;
  OHP_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    /* keep freed blocks in the device's pool instead of returning them to the driver (see OHP_MALLOC) */
    cudaMemPool_t pool;
    OHP_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = UINT64_MAX;
    OHP_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  OHP_CUDA(OHP_MALLOC(c, &c->d_partials, sizeof(double) * kMaxPartials * 2));
  OHP_CUDA(OHP_MALLOC(c, &c->d_ticket, sizeof(unsigned) * 8));
  OHP_CUDA(cudaMemset(c->d_ticket, 0, sizeof(unsigned) * 8));
  OHP_CUDA(OHP_MALLOC(c, &c->d_scalars, sizeof(double) * 16));
  OHP_CUDA(cudaMallocHost(&c->h_scalars, sizeof(double) * 16));
  OHP_CUDA(OHP_MALLOC(c, &c->d_cg, sizeof(ohp_cg_state)));
  OHP_CUDA(cudaMallocHost(&c->h_cg, sizeof(ohp_cg_state)));
  OHP_CUDA(OHP_MALLOC(c, &c->d_flag, sizeof(int)));
  OHP_CUDA(cudaMemset(c->d_flag, 0, sizeof(int)));
  *out = c;
  return 0;
}

extern "C" int ohp_finalize(ohp_context c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  ohp_comm_destroy(c);
  OHP_FREE(c, c->trace.buf); OHP_FREE(c, c->trace.count);
  OHP_FREE(c, c->d_partials); OHP_FREE(c, c->d_ticket); OHP_FREE(c, c->d_scalars);
  cudaFreeHost(c->h_scalars); OHP_FREE(c, c->d_cg); cudaFreeHost(c->h_cg); OHP_FREE(c, c->d_flag);
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

extern "C" int ohp_device_name(ohp_context c, char *buf, int buflen) {
  cudaDeviceProp prop;
  OHP_CUDA(cudaGetDeviceProperties(&prop, c->device));
  snprintf(buf, buflen, "%s (sm_%d%d, %d SMs)", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  return 0;
}
extern "C" int ohp_sync(ohp_context c) { OHP_CUDA(cudaStreamSynchronize(c->stream)); return 0; }
extern "C" int ohp_stream(ohp_context c, void **s) { *s = (void *)c->stream; return 0; }
extern "C" int ohp_launch_count(ohp_context c, int64_t *n) { *n = c->launches; return 0; }

/* ------------------------------------------------------------------------------------ */
/* exclusive scan (n inputs -> n+1 outputs, out[n] = total), the shape of Omega_h's      */
/* offset_scan (ex12.cpp:588,647,717).  Three phases, warp-shuffle block scans.          */
/* ------------------------------------------------------------------------------------ */
constexpr int SCAN_T = 512, SCAN_IPT = 8, SCAN_TILE = SCAN_T * SCAN_IPT;

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += t;
  }
  return v;
}

/* exclusive scan of one value per thread across the block; returns the block total via *total */
__device__ __forceinline__ int block_excl_scan(int v, int *total) {
  __shared__ int wsum[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = warp_incl_scan(v);
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int s = lane < nw ? wsum[lane] : 0;
    s = warp_incl_scan(s);
    wsum[lane] = s;
  }
  __syncthreads();
  const int base = wid ? wsum[wid - 1] : 0;
  *total = wsum[nw - 1];
  return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_tile_sums(const int32_t *in, int64_t n, int32_t *sums) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int s = 0;
  for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_T) {
    const int64_t g = base + i;
    if (g < n) s += in[g];
  }
  __shared__ int red[SCAN_T / 32];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < SCAN_T / 32; ++w) t += red[w];
    sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(SCAN_T) k_scan_apply(const int32_t *in, int32_t *out, int64_t n,
                                                       const int32_t *tile_off) {
  __shared__ int tile[SCAN_TILE];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_T) {
    const int64_t g = base + i;
    tile[i] = g < n ? in[g] : 0;
  }
  __syncthreads();
  int loc[SCAN_IPT], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_IPT; ++k) { loc[k] = s; s += tile[threadIdx.x * SCAN_IPT + k]; }
  int total;
  const int off = block_excl_scan(s, &total) + (tile_off ? tile_off[blockIdx.x] : 0);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < SCAN_IPT; ++k) tile[threadIdx.x * SCAN_IPT + k] = loc[k] + off;
  __syncthreads();
  for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_T) {
    const int64_t g = base + i;
    if (g < n) out[g] = tile[i];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0)
    out[n] = total + (tile_off ? tile_off[blockIdx.x] : 0);
}

static int scan_rec(ohp_context ctx, const int32_t *in, int32_t *out, int64_t n) {
  const int64_t nt = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (nt <= 1) {
    k_scan_apply<<<1, SCAN_T, 0, ctx->stream>>>(in, out, n, nullptr);
    OHP_CHECK_LAUNCH(ctx);
    return 0;
  }
  int32_t *sums = nullptr;
  OHP_CUDA(OHP_MALLOC(ctx, &sums, sizeof(int32_t) * (nt + 1)));
  k_scan_tile_sums<<<(unsigned)nt, SCAN_T, 0, ctx->stream>>>(in, n, sums);
  OHP_CHECK_LAUNCH(ctx);
  int rc = scan_rec(ctx, sums, sums, nt);
  if (!rc) {
    k_scan_apply<<<(unsigned)nt, SCAN_T, 0, ctx->stream>>>(in, out, n, sums);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) rc = OHP_ERR_LIB;
  }
  cudaStreamSynchronize(ctx->stream);
  OHP_FREE(ctx, sums);
  return rc;
}

int ohp_scan_exclusive_i32(ohp_context ctx, const int32_t *in, int32_t *out, int64_t n, int64_t *total) {
  if (n == 0) {
    OHP_CUDA(cudaMemsetAsync(out, 0, sizeof(int32_t), ctx->stream));
    if (total) *total = 0;
    return 0;
  }
  OHP_TRY(scan_rec(ctx, in, out, n));
  if (total) {
    int32_t t = 0;
    OHP_CUDA(cudaMemcpyAsync(&t, out + n, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    OHP_CUDA(cudaStreamSynchronize(ctx->stream));
    *total = t;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* DM creation                                                                           */
/* ------------------------------------------------------------------------------------ */
__global__ void k_check_cells(const int32_t *cells, int64_t n, int nV, int *flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (cells[i] < 0 || cells[i] >= nV)) atomicExch(flag, 1);
}

extern "C" int ohp_mesh_create(ohp_context ctx, int dim, int32_t nC, int32_t nV, const int32_t *cells,
                               const double *coords, ohp_mesh *out) {
  if (!ctx || !out) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_mesh_create: null argument");
  if (dim != 2 && dim != 3) OHP_FAIL(OHP_ERR_SUP, "ohp_mesh_create: dim %d (simplices in 2-D/3-D only)", dim);
  if (nC < 0 || nV < 0) OHP_FAIL(OHP_ERR_ARG_SIZ, "ohp_mesh_create: negative size");
  if ((int64_t)nC * (dim + 1) >= (int64_t)1 << 29) OHP_FAIL(OHP_ERR_ARG_SIZ, "ohp_mesh_create: too many cells for 32-bit incidence ids");
  if ((nC && !cells) || (nV && !coords)) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_mesh_create: null arrays");
  const int nc = dim + 1;
  OHP_CUDA(cudaSetDevice(ctx->device));
  ohp_mesh m = new ohp_mesh_s();
  m->ctx = ctx; m->dim = dim; m->nc = nc; m->nC = nC; m->nV = nV;
  m->cells = nullptr; m->coords = nullptr; m->owned = nullptr; m->setup_done = false; m->bc_type = OHP_BC_ESSENTIAL;
  m->v2c_ptr = m->v2c = nullptr; m->bnd = nullptr; m->lidx = m->row_vertex = nullptr; m->ldof_gid = nullptr;
  m->rowptr = m->colidx = nullptr; m->col16 = nullptr; m->slot = nullptr; m->tile_row = nullptr; m->halo = nullptr; m->p2p = nullptr; m->ghost_cta = nullptr;
  m->n_owned = m->n_ghost = m->n_bnd = 0; m->n_global = m->row_start = m->nnz = 0; m->max_row = 0; m->n_tiles = 0; m->max_tile_rows = 0;
  OHP_CUDA(OHP_MALLOC(ctx, &m->cells, sizeof(int32_t) * std::max<int64_t>(1, (int64_t)nC * nc)));
  OHP_CUDA(OHP_MALLOC(ctx, &m->coords, sizeof(double) * std::max<int64_t>(1, (int64_t)nV * dim)));
  OHP_CUDA(cudaMemcpyAsync(m->cells, cells, sizeof(int32_t) * (int64_t)nC * nc, cudaMemcpyHostToDevice, ctx->stream));
  OHP_CUDA(cudaMemcpyAsync(m->coords, coords, sizeof(double) * (int64_t)nV * dim, cudaMemcpyHostToDevice, ctx->stream));
  /* range check of the connectivity on the device (a host pass over 50 M ints costs more than the upload) */
  if (nC) {
    k_check_cells<<<(unsigned)std::max<int64_t>(1, ((int64_t)nC * nc + 255) / 256), 256, 0, ctx->stream>>>(m->cells, (int64_t)nC * nc, nV, ctx->d_flag);
    OHP_CHECK_LAUNCH(ctx);
  }
  int bad = 0;
  OHP_CUDA(cudaMemcpyAsync(&bad, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  OHP_CUDA(cudaStreamSynchronize(ctx->stream));
  if (bad) {
    cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream);
    int64_t i = 0;
    while (i < (int64_t)nC * nc && cells[i] >= 0 && cells[i] < nV) ++i;
    const int64_t bc = std::min<int64_t>(i, (int64_t)nC * nc - 1);
    ohp_mesh_destroy(m);
    OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "ohp_mesh_create: cell %lld references vertex %d outside [0,%d)", (long long)(bc / nc), cells[bc], nV);
  }
  *out = m;
  return 0;
}

extern "C" int ohp_mesh_destroy(ohp_mesh m) {
  if (!m) return 0;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  OHP_FREE(m->ctx, m->cells); OHP_FREE(m->ctx, m->coords); OHP_FREE(m->ctx, m->owned); OHP_FREE(m->ctx, m->v2c_ptr); OHP_FREE(m->ctx, m->v2c);
  OHP_FREE(m->ctx, m->bnd); OHP_FREE(m->ctx, m->lidx); OHP_FREE(m->ctx, m->row_vertex); OHP_FREE(m->ctx, m->ldof_gid); OHP_FREE(m->ctx, m->rowptr);
  OHP_FREE(m->ctx, m->colidx); OHP_FREE(m->ctx, m->col16); OHP_FREE(m->ctx, m->slot); OHP_FREE(m->ctx, m->tile_row);
  OHP_FREE(m->ctx, m->facet_verts); OHP_FREE(m->ctx, m->facet_support); OHP_FREE(m->ctx, m->cell_facets);
  ohp_fused_plan_free(m->ctx, m->fused);
  ohp_pf_plan_free(m->ctx, m->pfplan);
  ohp_patch_plan_free(m->ctx, m->patch);
  ohp_halo_free(m->ctx, m->halo);
  ohp_p2p_free(m->ctx, m->p2p);
  OHP_FREE(m->ctx, m->ghost_cta); OHP_FREE(m->ctx, m->ghost_tile); OHP_FREE(m->ctx, m->ghost_cta_w); OHP_FREE(m->ctx, m->cta_tile0);
  delete m;
  return 0;
}

extern "C" int ohp_mesh_set_ghosts(ohp_mesh m, int32_t n, const int32_t *gv, const int32_t *orank,
                                   const int32_t *overt) {
  if (!m) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_mesh_set_ghosts: null mesh");
  if (m->setup_done) OHP_FAIL(OHP_ERR_ORDER, "ohp_mesh_set_ghosts: called after ohp_mesh_setup");
  if (n > 0 && m->ctx->nranks == 1) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_mesh_set_ghosts: %d leaves but no communicator attached", n);
  for (int i = 0; i < n; ++i) {
    if (gv[i] < 0 || gv[i] >= m->nV) OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "ohp_mesh_set_ghosts: leaf %d vertex %d", i, gv[i]);
    if (orank[i] < 0 || orank[i] >= m->ctx->nranks || orank[i] == m->ctx->rank)
      OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "ohp_mesh_set_ghosts: leaf %d owner rank %d", i, orank[i]);
  }
  m->ghost_vertex.assign(gv, gv + n);
  m->ghost_owner_rank.assign(orank, orank + n);
  m->ghost_owner_vertex.assign(overt, overt + n);
  return 0;
}

extern "C" int ohp_mesh_set_point_sf(ohp_mesh m, int32_t n, const int32_t *leaf, const int32_t *rrank,
                                     const int32_t *rpoint) {
  if (!m) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_mesh_set_point_sf: null mesh");
  std::vector<int64_t> ncells;
  OHP_TRY(ohp_allgather_i64(m->ctx, m->nC, ncells)); /* getNeighborElmCounts, ex12.cpp:528-577 */
  std::vector<int32_t> gv(n), ov(n);
  for (int i = 0; i < n; ++i) {
    if (rrank[i] < 0 || rrank[i] >= m->ctx->nranks) OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "ohp_mesh_set_point_sf: leaf %d root rank %d", i, rrank[i]);
    gv[i] = leaf[i] - m->nC;
    ov[i] = rpoint[i] - (int32_t)ncells[rrank[i]];
    if (ov[i] < 0) OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "ohp_mesh_set_point_sf: root point %d of leaf %d is a cell of rank %d, not a vertex", rpoint[i], i, rrank[i]);
  }
  return ohp_mesh_set_ghosts(m, n, gv.data(), rrank, ov.data());
}

/* ------------------------------------------------------------------------------------ */
/* setup kernels                                                                         */
/* ------------------------------------------------------------------------------------ */
__global__ void k_count_incidence(const int32_t *cells, int64_t n, int32_t *deg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(deg + cells[i], 1);
}

__global__ void k_fill_v2c(const int32_t *cells, int nc, int64_t n, const int32_t *ptr, int32_t *cursor,
                           int32_t *v2c) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = cells[i];
  const int c = (int)(i / nc), b = (int)(i - (int64_t)c * nc);
  const int pos = atomicAdd(cursor + v, 1);
  v2c[ptr[v] + pos] = c * 4 + b;
}

/* the atomic fill order is arbitrary: sort every star ascending so that downstream
 * summation order (and therefore every assembled value) is deterministic */
__global__ void k_sort_v2c(const int32_t *ptr, int32_t *v2c, int nV) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV) return;
  const int a = ptr[v], b = ptr[v + 1];
  for (int i = a + 1; i < b; ++i) {
    const int x = v2c[i];
    int j = i - 1;
    while (j >= a && v2c[j] > x) { v2c[j + 1] = v2c[j]; --j; }
    v2c[j + 1] = x;
  }
}

/* facet (cell minus one vertex) is on the boundary iff exactly one cell contains it
 * (DMPlexGetSupportSize == 1, ex12.cpp:986-997); its vertices get "marker" = 1
 * (ex12.cpp:1001-1010,1023-1026).  The support is counted inside star(facet vertex 0). */
template <int NC>
__global__ void k_mark_boundary(const int32_t *cells, int nC, const int32_t *ptr, const int32_t *v2c, uint8_t *bnd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nC) return;
  int cv[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) cv[j] = cells[(size_t)c * NC + j];
#pragma unroll
  for (int skip = 0; skip < NC; ++skip) {
    int fv[NC - 1], nf = 0;
#pragma unroll
    for (int j = 0; j < NC; ++j) if (j != skip) fv[nf++] = cv[j];
    int support = 0;
    for (int k = ptr[fv[0]]; k < ptr[fv[0] + 1]; ++k) {
      const int oc = v2c[k] >> 2;
      int ov[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) ov[j] = cells[(size_t)oc * NC + j];
      bool all = true;
#pragma unroll
      for (int a = 1; a < NC - 1; ++a) {
        bool found = false;
#pragma unroll
        for (int j = 0; j < NC; ++j) found |= (ov[j] == fv[a]);
        all &= found;
      }
      support += all ? 1 : 0;
    }
    if (support == 1) {
#pragma unroll
      for (int a = 0; a < NC - 1; ++a) bnd[fv[a]] = 1; /* same-value write race, as ex12.cpp:636-644 */
    }
  }
}

/* `constrain` = 0: DM_BC_NATURAL / no boundary condition (ex12.cpp:1237, -bc_type neumann|none): the "marker" vertices
 * keep their dofs */
__global__ void k_free_mask(const uint8_t *bnd, const uint8_t *owned, int nV, int32_t *mask, int32_t *bmask, int constrain) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV) return;
  mask[v] = (owned[v] && !(constrain && bnd[v])) ? 1 : 0;
  bmask[v] = bnd[v] ? 1 : 0;
}

__global__ void k_assign_rows(const int32_t *mask, const int32_t *scan, int nV, int32_t *lidx, int32_t *row_vertex) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV) return;
  if (mask[v]) { lidx[v] = scan[v]; row_vertex[scan[v]] = v; }
  else lidx[v] = -1;
}

/* One warp per row: gather the local dofs of every vertex of every star cell, bitonic-sort them
 * in shared memory, then count (fill == 0) or write (fill == 1) the unique ones. */
constexpr int PAT_CAP = 512, PAT_WARPS = 8;
template <int NC>
__global__ void __launch_bounds__(PAT_WARPS * 32)
k_row_pattern(const int32_t *cells, const int32_t *ptr, const int32_t *v2c, const int32_t *lidx,
              const int32_t *row_vertex, int n_rows, int fill, int32_t *rowlen, const int32_t *rowptr,
              int32_t *colidx, int *err) {
  __shared__ int32_t s_keys[PAT_WARPS][PAT_CAP];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = blockIdx.x * PAT_WARPS + wid;
  if (r >= n_rows) return;
  int32_t *a = s_keys[wid];
  const int v = row_vertex[r];
  const int k0 = ptr[v], deg = ptr[v + 1] - k0;
  const int ncand = deg * NC;
  if (ncand > PAT_CAP) { if (lane == 0) atomicExch(err, 1); return; }
  int P = 32;
  while (P < ncand) P <<= 1;
  for (int i = lane; i < P; i += 32) {
    int key = 0x7fffffff;
    if (i < ncand) {
      const int c = v2c[k0 + i / NC] >> 2;
      const int g = lidx[cells[(size_t)c * NC + (i % NC)]];
      if (g >= 0) key = g;
    }
    a[i] = key;
  }
  __syncwarp();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < P; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const int x = a[i], y = a[ixj];
          const bool up = ((i & k) == 0);
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncwarp();
    }
  int count = 0;
  const int base_out = fill ? rowptr[r] : 0;
  for (int base = 0; base < P; base += 32) {
    const int i = base + lane;
    const int x = a[i];
    const int prev = i > 0 ? a[i - 1] : -1;
    const bool isnew = (x != 0x7fffffff) && (x != prev);
    const unsigned m = __ballot_sync(0xffffffffu, isnew);
    if (fill && isnew) colidx[base_out + count + __popc(m & ((1u << lane) - 1u))] = x;
    count += __popc(m);
  }
  if (!fill && lane == 0) rowlen[r] = count;
}

/* element -> row scatter map, stored from the row's side: for every incidence (v, c) of a row
 * vertex and every vertex j of c, the position of column lidx[c_j] inside row(v), or 255 */
template <int NC>
__global__ void k_slot_map(const int32_t *cells, const int32_t *ptr, const int32_t *v2c, const int32_t *lidx,
                           const int32_t *row_vertex, int n_rows, const int32_t *rowptr,
                           const int32_t *colidx, uint8_t *slot, int *err) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int v = row_vertex[r];
  const int a = rowptr[r], len = rowptr[r + 1] - a;
  if (len > 254) { atomicExch(err, 2); return; }
  for (int k = ptr[v]; k < ptr[v + 1]; ++k) {
    const int c = v2c[k] >> 2;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int g = lidx[cells[(size_t)c * NC + j]];
      int s = 255;
      if (g >= 0) {
        int lo = 0, hi = len - 1;
        while (lo <= hi) {
          const int mid = (lo + hi) >> 1;
          const int x = colidx[a + mid];
          if (x < g) lo = mid + 1; else if (x > g) hi = mid - 1; else { s = mid; break; }
        }
        if (s == 255) atomicExch(err, 3);
      }
      slot[(size_t)k * NC + j] = (uint8_t)s;
    }
  }
}

/* compressed column stream for the SpMV: offset of every column from its row, if all fit 16 bits */
__global__ void k_col16(const int32_t *rowptr, const int32_t *colidx, int n_rows, int16_t *col16, int *overflow) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  bool bad = false;
  for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
    const int d = colidx[k] - r;
    bad |= (d < -32768 || d > 32767);
    col16[k] = (int16_t)d;
  }
  if (bad) atomicExch(overflow, 1);
}

/* which CTAs of the persistent SpMV touch ghost columns.  CTA b owns tiles [n_tiles*b/G, n_tiles*(b+1)/G), or
 * [tile0[b], tile0[b+1]) when a partition table is given */
__global__ void k_ghost_cta(const int32_t *colidx, int64_t nnz, int n_owned, int n_tiles, int tile, uint8_t *flag, const int32_t *tile0) {
  const int64_t ta = tile0 ? tile0[blockIdx.x] : (((int64_t)n_tiles * blockIdx.x) / gridDim.x);
  const int64_t tb = tile0 ? tile0[blockIdx.x + 1] : (((int64_t)n_tiles * (blockIdx.x + 1)) / gridDim.x);
  const int64_t k0 = ta * tile;
  const int64_t k1 = std::min<int64_t>(nnz, tb * tile + 8192);
  int any = 0; /* + 8192: a row started in the last tile may run on into the next CTA's first tile */
  for (int64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) any |= (colidx[k] >= n_owned);
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) flag[blockIdx.x] = any ? 1 : 0;
}

/* per SpMV tile: does it (or the tail of a row it starts) reference a ghost column?  Used by the dynamically scheduled
 * SpMV, where the CTA that gets a tile is not known in advance */
__global__ void k_ghost_tile(const int32_t *colidx, int64_t nnz, int n_owned, int tile, uint8_t *flag) {
  const int64_t k0 = (int64_t)blockIdx.x * tile;
  const int64_t k1 = std::min<int64_t>(nnz, k0 + tile + 8192);
  int any = 0;
  for (int64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) any |= (colidx[k] >= n_owned);
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) flag[blockIdx.x] = any ? 1 : 0;
}

__global__ void k_max_rowlen(const int32_t *rowptr, int n_rows, int32_t *out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  int len = r < n_rows ? rowptr[r + 1] - rowptr[r] : 0;
#pragma unroll
  for (int o = 16; o; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
  if ((threadIdx.x & 31) == 0 && len > 0) atomicMax(out, len);
}

/* first row whose rowptr is >= t * tile: rows [tile_row[t], tile_row[t+1]) form SpMV tile t */
__global__ void k_tile_rows(const int32_t *rowptr, int n_rows, int tile, int n_tiles, int32_t *tile_row) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_tiles) return;
  if (t == n_tiles) { tile_row[t] = n_rows; return; }
  const int64_t target = (int64_t)t * tile;
  int lo = 0, hi = n_rows;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (rowptr[mid] < target) lo = mid + 1; else hi = mid;
  }
  tile_row[t] = lo;
}

__global__ void k_fill_u8(uint8_t *p, int n, uint8_t val) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = val;
}
__global__ void k_clear_owned(uint8_t *owned, const int32_t *gv, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) owned[gv[i]] = 0;
}
__global__ void k_u8_to_i32(const uint8_t *in, int32_t *out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
__global__ void k_i32_to_u8_ghost(const int32_t *in, const uint8_t *owned, uint8_t *out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !owned[i]) out[i] = (uint8_t)(in[i] != 0);
}
__global__ void k_owned_gid(int64_t *gid, int n, int64_t start) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gid[i] = start + i;
}

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)std::max<int64_t>(1, (n + bs - 1) / bs); }

static int check_flag(ohp_context ctx, const char *what) {
  int flag = 0;
  OHP_CUDA(cudaMemcpyAsync(&flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  OHP_CUDA(cudaStreamSynchronize(ctx->stream));
  if (flag) {
    cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream);
    OHP_FAIL(OHP_ERR_SUP, "%s: device-side limit exceeded (code %d: 1 = vertex star too large for the pattern sort, "
             "2 = row longer than 254, 3 = inconsistent pattern)", what, flag);
  }
  return 0;
}

__global__ void k_set_label(uint8_t *bnd, const int32_t *verts, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) bnd[verts[i]] = 1;
}

/* vertex -> cell stars (count, scan, fill, sort); idempotent */
static int mesh_build_stars(ohp_mesh m) {
  if (m->v2c_ptr) return 0;
  ohp_context ctx = m->ctx;
  cudaStream_t st = ctx->stream;
  const int nc = m->nc, nV = m->nV, BS = 256;
  const int64_t ninc = (int64_t)m->nC * nc;
  int32_t *deg;
  OHP_CUDA(OHP_MALLOC(ctx, &deg, sizeof(int32_t) * (nV + 1)));
  OHP_CUDA(cudaMemsetAsync(deg, 0, sizeof(int32_t) * (nV + 1), st));
  OHP_CUDA(OHP_MALLOC(ctx, &m->v2c_ptr, sizeof(int32_t) * (nV + 1)));
  OHP_CUDA(OHP_MALLOC(ctx, &m->v2c, sizeof(int32_t) * std::max<int64_t>(1, ninc)));
  if (ninc) { k_count_incidence<<<nblk(ninc, BS), BS, 0, st>>>(m->cells, ninc, deg); OHP_CHECK_LAUNCH(ctx); }
  OHP_TRY(ohp_scan_exclusive_i32(ctx, deg, m->v2c_ptr, nV, nullptr));
  OHP_CUDA(cudaMemsetAsync(deg, 0, sizeof(int32_t) * (nV + 1), st));
  if (ninc) { k_fill_v2c<<<nblk(ninc, BS), BS, 0, st>>>(m->cells, nc, ninc, m->v2c_ptr, deg, m->v2c); OHP_CHECK_LAUNCH(ctx); }
  k_sort_v2c<<<nblk(nV, BS), BS, 0, st>>>(m->v2c_ptr, m->v2c, nV); OHP_CHECK_LAUNCH(ctx);
  OHP_CUDA(cudaStreamSynchronize(st));
  OHP_FREE(ctx, deg);
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Facets = what DMPlexInterpolate creates between cells and vertices (ex12.cpp:958): edges in 2-D, */
/* triangles in 3-D.  A facet is created by the first cell (lowest index) that contains it, in that  */
/* cell's local facet order -- the order of PETSc's interpolation loop [PETSc-internal, unverified    */
/* here] -- so the numbering is deterministic.  Needed only by callers that walk the Plex themselves  */
/* (DMPlexGetHeightStratum / GetSupportSize / GetCone, ex12.cpp:966-1010).                          */
/* ------------------------------------------------------------------------------------ */
template <int NC> struct FacetTable;
template <> struct FacetTable<3> { /* triangle edges: (v0,v1) (v1,v2) (v2,v0) */
  __host__ __device__ static int v(int f, int k) { return (f + k) % 3; }
};
template <> struct FacetTable<4> { /* tetrahedron faces: (0,1,2) (0,3,1) (0,2,3) (2,1,3) */
  __host__ __device__ static int v(int f, int k) {
    constexpr int F[4][3] = {{0, 1, 2}, {0, 3, 1}, {0, 2, 3}, {2, 1, 3}};
    return F[f][k];
  }
};

template <int NC>
__global__ void k_facet_first(const int32_t *cells, int nC, const int32_t *ptr, const int32_t *v2c, int32_t *is_first,
                              int32_t *mincell, int32_t *support) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)nC * NC) return;
  const int c = (int)(i / NC), f = (int)(i % NC);
  int fv[NC - 1];
#pragma unroll
  for (int k = 0; k < NC - 1; ++k) fv[k] = cells[(size_t)c * NC + FacetTable<NC>::v(f, k)];
  int sup = 0, mc = c;
  for (int k = ptr[fv[0]]; k < ptr[fv[0] + 1]; ++k) {
    const int oc = v2c[k] >> 2;
    bool all = true;
#pragma unroll
    for (int a = 1; a < NC - 1; ++a) {
      bool found = false;
#pragma unroll
      for (int j = 0; j < NC; ++j) found |= (cells[(size_t)oc * NC + j] == fv[a]);
      all &= found;
    }
    if (all) { sup++; mc = min(mc, oc); }
  }
  is_first[i] = (mc == c) ? 1 : 0;
  mincell[i] = mc;
  support[i] = sup;
}

template <int NC>
__global__ void k_facet_fill(const int32_t *cells, int nC, const int32_t *is_first, const int32_t *scan, const int32_t *mincell,
                             const int32_t *support, int32_t *facet_verts, int32_t *facet_support, int32_t *cell_facets) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)nC * NC) return;
  const int c = (int)(i / NC), f = (int)(i % NC);
  int fv[NC - 1];
#pragma unroll
  for (int k = 0; k < NC - 1; ++k) fv[k] = cells[(size_t)c * NC + FacetTable<NC>::v(f, k)];
  int fid;
  if (is_first[i]) {
    fid = scan[i];
#pragma unroll
    for (int k = 0; k < NC - 1; ++k) facet_verts[(size_t)fid * (NC - 1) + k] = fv[k];
    facet_support[fid] = support[i];
  } else {
    /* the same facet seen from the cell that created it: find its local number there */
    const int oc = mincell[i];
    fid = -1;
#pragma unroll
    for (int g = 0; g < NC; ++g) {
      bool same = true;
#pragma unroll
      for (int k = 0; k < NC - 1; ++k) {
        const int w = cells[(size_t)oc * NC + FacetTable<NC>::v(g, k)];
        bool found = false;
#pragma unroll
        for (int a = 0; a < NC - 1; ++a) found |= (w == fv[a]);
        same &= found;
      }
      if (same) fid = scan[(size_t)oc * NC + g];
    }
  }
  cell_facets[i] = fid;
}

static int mesh_build_facets(ohp_mesh m) {
  if (m->facets_built) return 0;
  ohp_context ctx = m->ctx;
  OHP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  OHP_TRY(mesh_build_stars(m));
  const int nc = m->nc, nC = m->nC, BS = 256;
  const int64_t ninc = (int64_t)nC * nc;
  int32_t *is_first, *scan, *mincell, *support;
  OHP_CUDA(OHP_MALLOC(ctx, &is_first, sizeof(int32_t) * (ninc + 1)));
  OHP_CUDA(OHP_MALLOC(ctx, &scan, sizeof(int32_t) * (ninc + 1)));
  OHP_CUDA(OHP_MALLOC(ctx, &mincell, sizeof(int32_t) * (ninc + 1)));
  OHP_CUDA(OHP_MALLOC(ctx, &support, sizeof(int32_t) * (ninc + 1)));
  if (ninc) {
    if (nc == 3) k_facet_first<3><<<nblk(ninc, BS), BS, 0, st>>>(m->cells, nC, m->v2c_ptr, m->v2c, is_first, mincell, support);
    else k_facet_first<4><<<nblk(ninc, BS), BS, 0, st>>>(m->cells, nC, m->v2c_ptr, m->v2c, is_first, mincell, support);
    OHP_CHECK_LAUNCH(ctx);
  }
  int64_t nF = 0;
  OHP_TRY(ohp_scan_exclusive_i32(ctx, is_first, scan, ninc, &nF));
  m->n_facets = (int32_t)nF;
  OHP_CUDA(OHP_MALLOC(ctx, &m->facet_verts, sizeof(int32_t) * std::max<int64_t>(1, nF * (nc - 1))));
  OHP_CUDA(OHP_MALLOC(ctx, &m->facet_support, sizeof(int32_t) * std::max<int64_t>(1, nF)));
  OHP_CUDA(OHP_MALLOC(ctx, &m->cell_facets, sizeof(int32_t) * std::max<int64_t>(1, ninc)));
  if (ninc) {
    if (nc == 3) k_facet_fill<3><<<nblk(ninc, BS), BS, 0, st>>>(m->cells, nC, is_first, scan, mincell, support, m->facet_verts, m->facet_support, m->cell_facets);
    else k_facet_fill<4><<<nblk(ninc, BS), BS, 0, st>>>(m->cells, nC, is_first, scan, mincell, support, m->facet_verts, m->facet_support, m->cell_facets);
    OHP_CHECK_LAUNCH(ctx);
  }
  OHP_CUDA(cudaStreamSynchronize(st));
  OHP_FREE(ctx, is_first); OHP_FREE(ctx, scan); OHP_FREE(ctx, mincell); OHP_FREE(ctx, support);
  m->facets_built = true;
  return 0;
}

extern "C" int ohp_mesh_build_facets(ohp_mesh m, int32_t *n_facets) {
  if (!m) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_mesh_build_facets: null mesh");
  OHP_TRY(mesh_build_facets(m));
  if (n_facets) *n_facets = m->n_facets;
  return 0;
}

extern "C" int ohp_mesh_get_facets(ohp_mesh m, int32_t *facet_verts, int32_t *support, int32_t *cell_facets) {
  if (!m) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_mesh_get_facets: null mesh");
  OHP_TRY(mesh_build_facets(m));
  cudaStream_t st = m->ctx->stream;
  if (facet_verts) OHP_CUDA(cudaMemcpyAsync(facet_verts, m->facet_verts, sizeof(int32_t) * (size_t)m->n_facets * (m->nc - 1), cudaMemcpyDeviceToHost, st));
  if (support) OHP_CUDA(cudaMemcpyAsync(support, m->facet_support, sizeof(int32_t) * (size_t)m->n_facets, cudaMemcpyDeviceToHost, st));
  if (cell_facets) OHP_CUDA(cudaMemcpyAsync(cell_facets, m->cell_facets, sizeof(int32_t) * (size_t)m->nC * m->nc, cudaMemcpyDeviceToHost, st));
  OHP_CUDA(cudaStreamSynchronize(st));
  return 0;
}

/* DMSetLabelValue(dm, "marker", point, 1) for vertex points, collected (ex12.cpp:1023-1026): replaces the
 * library's own boundary detection for the vertices this rank owns; ghost vertices still take their owner's label */
extern "C" int ohp_mesh_set_label(ohp_mesh m, int32_t n, const int32_t *vertices) {
  if (!m || (n > 0 && !vertices)) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_mesh_set_label: null argument");
  if (m->setup_done) OHP_FAIL(OHP_ERR_ORDER, "ohp_mesh_set_label: called after ohp_mesh_setup");
  for (int i = 0; i < n; ++i)
    if (vertices[i] < 0 || vertices[i] >= m->nV) OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "ohp_mesh_set_label: vertex %d outside [0,%d)", vertices[i], m->nV);
  m->user_label.assign(vertices, vertices + n);
  m->has_user_label = true;
  return 0;
}

extern "C" int ohp_mesh_set_bc_type(ohp_mesh m, int bc_type) {
  if (!m) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_mesh_set_bc_type: null mesh");
  if (m->setup_done) OHP_FAIL(OHP_ERR_ORDER, "ohp_mesh_set_bc_type: called after ohp_mesh_setup");
  if (bc_type != OHP_BC_ESSENTIAL && bc_type != OHP_BC_NATURAL && bc_type != OHP_BC_NONE)
    OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "ohp_mesh_set_bc_type: unknown type %d", bc_type);
  m->bc_type = bc_type;
  return 0;
}

extern "C" int ohp_mesh_setup(ohp_mesh m) {
  if (!m) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_mesh_setup: null mesh");
  if (m->setup_done) return 0;
  ohp_context ctx = m->ctx;
  OHP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int nc = m->nc, nV = m->nV, nC = m->nC;
  const int64_t ninc = (int64_t)nC * nc;
  const int BS = 256;

  /* ownership: every vertex that is not an SF leaf */
  OHP_CUDA(OHP_MALLOC(ctx, &m->owned, std::max(1, nV)));
  k_fill_u8<<<nblk(nV, BS), BS, 0, st>>>(m->owned, nV, 1); OHP_CHECK_LAUNCH(ctx);
  const int nleaf = (int)m->ghost_vertex.size();
  if (nleaf) {
    int32_t *d_gv;
    OHP_CUDA(OHP_MALLOC(ctx, &d_gv, sizeof(int32_t) * nleaf));
    OHP_CUDA(cudaMemcpyAsync(d_gv, m->ghost_vertex.data(), sizeof(int32_t) * nleaf, cudaMemcpyHostToDevice, st));
    k_clear_owned<<<nblk(nleaf, BS), BS, 0, st>>>(m->owned, d_gv, nleaf); OHP_CHECK_LAUNCH(ctx);
    OHP_CUDA(cudaStreamSynchronize(st));
    OHP_FREE(ctx, d_gv);
  }

  /* vertex -> cell stars */
  OHP_TRY(mesh_build_stars(m));
  int32_t *deg;
  OHP_CUDA(OHP_MALLOC(ctx, &deg, sizeof(int32_t) * (nV + 1)));

  /* boundary label: facets with one supporting cell (ex12.cpp:986-1010), or the label the caller set point by
   * point with DMSetLabelValue (ex12.cpp:1023-1026) */
  OHP_CUDA(OHP_MALLOC(ctx, &m->bnd, std::max(1, nV)));
  OHP_CUDA(cudaMemsetAsync(m->bnd, 0, std::max(1, nV), st));
  if (m->has_user_label) {
    const int nl = (int)m->user_label.size();
    if (nl) {
      int32_t *d_lab;
      OHP_CUDA(OHP_MALLOC(ctx, &d_lab, sizeof(int32_t) * nl));
      OHP_CUDA(cudaMemcpyAsync(d_lab, m->user_label.data(), sizeof(int32_t) * nl, cudaMemcpyHostToDevice, st));
      k_set_label<<<nblk(nl, BS), BS, 0, st>>>(m->bnd, d_lab, nl); OHP_CHECK_LAUNCH(ctx);
      OHP_CUDA(cudaStreamSynchronize(st));
      OHP_FREE(ctx, d_lab);
    }
  } else if (nC) {
    if (nc == 3) k_mark_boundary<3><<<nblk(nC, BS), BS, 0, st>>>(m->cells, nC, m->v2c_ptr, m->v2c, m->bnd);
    else k_mark_boundary<4><<<nblk(nC, BS), BS, 0, st>>>(m->cells, nC, m->v2c_ptr, m->v2c, m->bnd);
    OHP_CHECK_LAUNCH(ctx);
  }
  int32_t *tmp_a, *tmp_b;
  OHP_CUDA(OHP_MALLOC(ctx, &tmp_a, sizeof(int32_t) * (nV + 1)));
  OHP_CUDA(OHP_MALLOC(ctx, &tmp_b, sizeof(int32_t) * (nV + 1)));
  if (ctx->nranks > 1) {
    /* a ghost vertex's star is incomplete here: take its label from the owner (hazard H2) */
    k_u8_to_i32<<<nblk(nV, BS), BS, 0, st>>>(m->bnd, tmp_a, nV); OHP_CHECK_LAUNCH(ctx);
    OHP_TRY(ohp_vertex_halo_i32(m, tmp_a));
    k_i32_to_u8_ghost<<<nblk(nV, BS), BS, 0, st>>>(tmp_a, m->owned, m->bnd, nV); OHP_CHECK_LAUNCH(ctx);
  }

  /* section: owned unconstrained vertices, ascending local point order */
  k_free_mask<<<nblk(nV, BS), BS, 0, st>>>(m->bnd, m->owned, nV, tmp_a, tmp_b, m->bc_type == OHP_BC_ESSENTIAL ? 1 : 0); OHP_CHECK_LAUNCH(ctx);
  int64_t n_owned = 0, n_bnd = 0;
  OHP_TRY(ohp_scan_exclusive_i32(ctx, tmp_b, deg, nV, &n_bnd));
  OHP_TRY(ohp_scan_exclusive_i32(ctx, tmp_a, deg, nV, &n_owned));
  m->n_owned = (int32_t)n_owned; m->n_bnd = (int32_t)n_bnd;
  OHP_CUDA(OHP_MALLOC(ctx, &m->lidx, sizeof(int32_t) * std::max(1, nV)));
  OHP_CUDA(OHP_MALLOC(ctx, &m->row_vertex, sizeof(int32_t) * std::max(1, m->n_owned)));
  k_assign_rows<<<nblk(nV, BS), BS, 0, st>>>(tmp_a, deg, nV, m->lidx, m->row_vertex); OHP_CHECK_LAUNCH(ctx);
  OHP_CUDA(cudaStreamSynchronize(st));
  OHP_FREE(ctx, tmp_a); OHP_FREE(ctx, tmp_b); OHP_FREE(ctx, deg);

  /* global numbering: ranks concatenated in rank order; ghost dofs + halo plan */
  m->row_start = 0; m->n_global = m->n_owned; m->n_ghost = 0;
  if (ctx->nranks > 1) OHP_TRY(ohp_halo_build(m)); /* sets row_start, n_global, n_ghost, ghost lidx, ldof_gid */
  else {
    OHP_CUDA(OHP_MALLOC(ctx, &m->ldof_gid, sizeof(int64_t) * std::max(1, m->n_owned)));
    k_owned_gid<<<nblk(m->n_owned, BS), BS, 0, st>>>(m->ldof_gid, m->n_owned, 0); OHP_CHECK_LAUNCH(ctx);
  }

  /* CSR pattern: count, scan, fill */
  const int N = m->n_owned;
  int32_t *rowlen;
  OHP_CUDA(OHP_MALLOC(ctx, &rowlen, sizeof(int32_t) * (N + 1)));
  OHP_CUDA(OHP_MALLOC(ctx, &m->rowptr, sizeof(int32_t) * (N + 1)));
  if (N) {
    if (nc == 3) k_row_pattern<3><<<nblk(N, PAT_WARPS), PAT_WARPS * 32, 0, st>>>(m->cells, m->v2c_ptr, m->v2c, m->lidx, m->row_vertex, N, 0, rowlen, nullptr, nullptr, ctx->d_flag);
    else k_row_pattern<4><<<nblk(N, PAT_WARPS), PAT_WARPS * 32, 0, st>>>(m->cells, m->v2c_ptr, m->v2c, m->lidx, m->row_vertex, N, 0, rowlen, nullptr, nullptr, ctx->d_flag);
    OHP_CHECK_LAUNCH(ctx);
  }
  OHP_TRY(check_flag(ctx, "ohp_mesh_setup (pattern count)"));
  OHP_TRY(ohp_scan_exclusive_i32(ctx, rowlen, m->rowptr, N, &m->nnz));
  OHP_FREE(ctx, rowlen);
  OHP_CUDA(OHP_MALLOC(ctx, &m->colidx, sizeof(int32_t) * ohp_padded_nnz(m->nnz)));
  OHP_CUDA(cudaMemsetAsync(m->colidx, 0, sizeof(int32_t) * ohp_padded_nnz(m->nnz), st));
  OHP_CUDA(OHP_MALLOC(ctx, &m->slot, std::max<int64_t>(1, ninc * nc)));
  OHP_CUDA(cudaMemsetAsync(m->slot, 0xff, std::max<int64_t>(1, ninc * nc), st));
  if (N) {
    if (nc == 3) k_row_pattern<3><<<nblk(N, PAT_WARPS), PAT_WARPS * 32, 0, st>>>(m->cells, m->v2c_ptr, m->v2c, m->lidx, m->row_vertex, N, 1, nullptr, m->rowptr, m->colidx, ctx->d_flag);
    else k_row_pattern<4><<<nblk(N, PAT_WARPS), PAT_WARPS * 32, 0, st>>>(m->cells, m->v2c_ptr, m->v2c, m->lidx, m->row_vertex, N, 1, nullptr, m->rowptr, m->colidx, ctx->d_flag);
    OHP_CHECK_LAUNCH(ctx);
    if (nc == 3) k_slot_map<3><<<nblk(N, 128), 128, 0, st>>>(m->cells, m->v2c_ptr, m->v2c, m->lidx, m->row_vertex, N, m->rowptr, m->colidx, m->slot, ctx->d_flag);
    else k_slot_map<4><<<nblk(N, 128), 128, 0, st>>>(m->cells, m->v2c_ptr, m->v2c, m->lidx, m->row_vertex, N, m->rowptr, m->colidx, m->slot, ctx->d_flag);
    OHP_CHECK_LAUNCH(ctx);
  }
  OHP_TRY(check_flag(ctx, "ohp_mesh_setup (pattern fill)"));

  /* 16-bit column offsets (2 B/nnz less SpMV traffic) when the ordering keeps every row's columns within +-32767 */
  if (N) {
    OHP_CUDA(OHP_MALLOC(ctx, &m->col16, sizeof(int16_t) * ohp_padded_nnz(m->nnz)));
    OHP_CUDA(cudaMemsetAsync(m->col16, 0, sizeof(int16_t) * ohp_padded_nnz(m->nnz), st));
    k_col16<<<nblk(N, BS), BS, 0, st>>>(m->rowptr, m->colidx, N, m->col16, ctx->d_flag); OHP_CHECK_LAUNCH(ctx);
    int overflow = 0;
    OHP_CUDA(cudaMemcpyAsync(&overflow, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
    if (overflow) { /* e.g. a space-filling-curve ordering: the 32-bit stream costs 6 % (measured), see ohp_dev.cuh */
      OHP_CUDA(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), st));
      OHP_FREE(ctx, m->col16);
      m->col16 = nullptr;
    }
  }

  /* longest row + SpMV tiles */
  int32_t *d_max;
  OHP_CUDA(OHP_MALLOC(ctx, &d_max, sizeof(int32_t)));
  OHP_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int32_t), st));
  if (N) { k_max_rowlen<<<nblk(N, BS), BS, 0, st>>>(m->rowptr, N, d_max); OHP_CHECK_LAUNCH(ctx); }
  OHP_CUDA(cudaMemcpyAsync(&m->max_row, d_max, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  OHP_CUDA(cudaStreamSynchronize(st));
  OHP_FREE(ctx, d_max);
  m->n_tiles = (int32_t)((m->nnz + kSpmvTile - 1) / kSpmvTile);
  if (m->n_tiles == 0 && N > 0) m->n_tiles = 1;
  OHP_CUDA(OHP_MALLOC(ctx, &m->tile_row, sizeof(int32_t) * (m->n_tiles + 1)));
  k_tile_rows<<<nblk(m->n_tiles + 1, BS), BS, 0, st>>>(m->rowptr, N, kSpmvTile, m->n_tiles, m->tile_row); OHP_CHECK_LAUNCH(ctx);
  if (ctx->nranks > 1 && m->n_tiles > 0) {
    const int G = std::max(1, std::min(m->n_tiles, ctx->num_sms));
    OHP_CUDA(OHP_MALLOC(ctx, &m->ghost_cta, ctx->num_sms));
    OHP_CUDA(cudaMemsetAsync(m->ghost_cta, 1, ctx->num_sms, st));
    k_ghost_cta<<<G, 256, 0, st>>>(m->colidx, m->nnz, m->n_owned, m->n_tiles, kSpmvTile, m->ghost_cta, nullptr); OHP_CHECK_LAUNCH(ctx);
    {
      /* Weighted tile partition for the SpMV of the partitioned CG loops: a CTA that reads ghost columns waits for the
       * neighbours' halo flags and loads x coherently, which made the strips' first and last CTAs the slowest of every
       * launch (device timeline at 8 GPUs: 25-29 us against a median of 23); they get kGhostCtaShare of a full share.
       * The flags are then recomputed for the shifted partition. */
      static const char *sh_env = getenv("OHP_GHOST_CTA_SHARE"); /* percent (A/B runs); 100 = uniform */
      const double share = sh_env ? atof(sh_env) / 100.0 : 0.85;
      std::vector<uint8_t> gc(G);
      OHP_CUDA(cudaMemcpyAsync(gc.data(), m->ghost_cta, G, cudaMemcpyDeviceToHost, st));
      OHP_CUDA(cudaStreamSynchronize(st));
      double total = 0.0;
      for (int b = 0; b < G; ++b) total += gc[b] ? share : 1.0;
      std::vector<int32_t> t0(G + 1, 0);
      double acc = 0.0;
      for (int b = 0; b < G; ++b) {
        acc += gc[b] ? share : 1.0;
        t0[b + 1] = std::max(t0[b], std::min(m->n_tiles, (int)llround((double)m->n_tiles * acc / total)));
      }
      t0[G] = m->n_tiles;
      OHP_CUDA(OHP_MALLOC(ctx, &m->cta_tile0, sizeof(int32_t) * (G + 1)));
      OHP_CUDA(cudaMemcpyAsync(m->cta_tile0, t0.data(), sizeof(int32_t) * (G + 1), cudaMemcpyHostToDevice, st));
      OHP_CUDA(OHP_MALLOC(ctx, &m->ghost_cta_w, ctx->num_sms));
      OHP_CUDA(cudaMemsetAsync(m->ghost_cta_w, 1, ctx->num_sms, st));
      k_ghost_cta<<<G, 256, 0, st>>>(m->colidx, m->nnz, m->n_owned, m->n_tiles, kSpmvTile, m->ghost_cta_w, m->cta_tile0); OHP_CHECK_LAUNCH(ctx);
      OHP_CUDA(cudaStreamSynchronize(st)); /* t0 goes out of scope */
    }
    OHP_CUDA(OHP_MALLOC(ctx, &m->ghost_tile, m->n_tiles));
    k_ghost_tile<<<m->n_tiles, 256, 0, st>>>(m->colidx, m->nnz, m->n_owned, kSpmvTile, m->ghost_tile); OHP_CHECK_LAUNCH(ctx);
  }
  {
    std::vector<int32_t> tr(m->n_tiles + 1);
    OHP_CUDA(cudaMemcpyAsync(tr.data(), m->tile_row, sizeof(int32_t) * (m->n_tiles + 1), cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
    m->max_tile_rows = 0;
    for (int t = 0; t < m->n_tiles; ++t) m->max_tile_rows = std::max(m->max_tile_rows, tr[t + 1] - tr[t]);
  }
  /* can EVERY rank run the TMA-pipelined SpMV?  The CG variants built on it are selected from this global answer, so
   * that all ranks take the same branch (one all-reduce at set-up) */
  m->all_tma = m->max_tile_rows <= kSpmvRowCap - 1;
  if (ctx->nranks > 1 && ctx->comm) {
    const double bad = m->all_tma ? 0.0 : 1.0;
    double sum = 0.0;
    OHP_CUDA(cudaMemcpyAsync(ctx->d_scalars + 12, &bad, sizeof(double), cudaMemcpyHostToDevice, st));
    OHP_TRY(ohp_allreduce_sum(ctx, ctx->d_scalars + 12, 1));
    OHP_CUDA(cudaMemcpyAsync(&sum, ctx->d_scalars + 12, sizeof(double), cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
    m->all_tma = (sum == 0.0);
  }
  OHP_TRY(ohp_patch_plan_build(m)); /* cell-parallel assembly tables; leaves m->patch NULL when a limit is exceeded */
  m->setup_done = true;
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* getters                                                                               */
/* ------------------------------------------------------------------------------------ */
#define NEED_SETUP(m, name)                                                     \
  do {                                                                          \
    if (!(m)) OHP_FAIL(OHP_ERR_ARG_WRONG, name ": null mesh");                  \
    if (!(m)->setup_done) OHP_FAIL(OHP_ERR_ORDER, name ": ohp_mesh_setup has not been called"); \
    OHP_CUDA(cudaSetDevice((m)->ctx->device));                                  \
  } while (0)

extern "C" int ohp_mesh_get_info(ohp_mesh m, ohp_mesh_info *info) {
  NEED_SETUP(m, "ohp_mesh_get_info");
  info->dim = m->dim; info->n_cells = m->nC; info->n_verts = m->nV; info->n_boundary_verts = m->n_bnd;
  info->n_owned = m->n_owned; info->n_ghost = m->n_ghost; info->n_global = m->n_global;
  info->row_start = m->row_start; info->nnz = m->nnz; info->max_row = m->max_row;
  return 0;
}

extern "C" int ohp_mesh_get_cells(ohp_mesh m, int32_t *cells) {
  if (!m || !cells) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_mesh_get_cells: null argument");
  OHP_CUDA(cudaSetDevice(m->ctx->device));
  OHP_CUDA(cudaMemcpyAsync(cells, m->cells, sizeof(int32_t) * (size_t)m->nC * m->nc, cudaMemcpyDeviceToHost, m->ctx->stream));
  OHP_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return 0;
}
extern "C" int ohp_mesh_get_coords(ohp_mesh m, double *coords) {
  if (!m || !coords) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_mesh_get_coords: null argument");
  OHP_CUDA(cudaSetDevice(m->ctx->device));
  OHP_CUDA(cudaMemcpyAsync(coords, m->coords, sizeof(double) * (size_t)m->nV * m->dim, cudaMemcpyDeviceToHost, m->ctx->stream));
  OHP_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return 0;
}

extern "C" int ohp_mesh_get_boundary(ohp_mesh m, uint8_t *bnd) {
  NEED_SETUP(m, "ohp_mesh_get_boundary");
  OHP_CUDA(cudaMemcpyAsync(bnd, m->bnd, m->nV, cudaMemcpyDeviceToHost, m->ctx->stream));
  OHP_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return 0;
}

extern "C" int ohp_mesh_get_numbering(ohp_mesh m, int64_t *gidx) {
  NEED_SETUP(m, "ohp_mesh_get_numbering");
  std::vector<int32_t> lidx(m->nV);
  std::vector<int64_t> gid(m->n_owned + m->n_ghost);
  OHP_CUDA(cudaMemcpyAsync(lidx.data(), m->lidx, sizeof(int32_t) * m->nV, cudaMemcpyDeviceToHost, m->ctx->stream));
  OHP_CUDA(cudaMemcpyAsync(gid.data(), m->ldof_gid, sizeof(int64_t) * gid.size(), cudaMemcpyDeviceToHost, m->ctx->stream));
  OHP_CUDA(cudaStreamSynchronize(m->ctx->stream));
  for (int v = 0; v < m->nV; ++v) gidx[v] = lidx[v] >= 0 ? gid[lidx[v]] : -1;
  return 0;
}

extern "C" int ohp_mesh_get_csr(ohp_mesh m, int32_t *rowptr, int64_t *colidx) {
  NEED_SETUP(m, "ohp_mesh_get_csr");
  std::vector<int32_t> col(m->nnz);
  std::vector<int64_t> gid(m->n_owned + m->n_ghost);
  OHP_CUDA(cudaMemcpyAsync(rowptr, m->rowptr, sizeof(int32_t) * (m->n_owned + 1), cudaMemcpyDeviceToHost, m->ctx->stream));
  OHP_CUDA(cudaMemcpyAsync(col.data(), m->colidx, sizeof(int32_t) * m->nnz, cudaMemcpyDeviceToHost, m->ctx->stream));
  OHP_CUDA(cudaMemcpyAsync(gid.data(), m->ldof_gid, sizeof(int64_t) * gid.size(), cudaMemcpyDeviceToHost, m->ctx->stream));
  OHP_CUDA(cudaStreamSynchronize(m->ctx->stream));
  for (int64_t k = 0; k < m->nnz; ++k) colidx[k] = gid[col[k]];
  return 0;
}

int ohp_assembly_view_make(ohp_mesh m, const double *u, double *out, ohp_assembly_view *vw) {
  vw->dim = m->dim; vw->n_cells = m->nC; vw->n_verts = m->nV; vw->n_rows = m->n_owned;
  vw->cells_dev = m->cells; vw->coords_dev = m->coords; vw->lidx_dev = m->lidx;
  vw->row_vertex_dev = m->row_vertex; vw->v2c_ptr_dev = m->v2c_ptr; vw->v2c_dev = m->v2c;
  vw->slot_dev = m->slot; vw->rowptr_dev = m->rowptr; vw->u_dev = u; vw->out_dev = out;
  vw->stream = (void *)m->ctx->stream;
  vw->owned_dev = m->owned; vw->n_partials = 0; vw->bnd_dev = m->bnd;
  vw->n_patches = 0; vw->max_row = m->max_row;
  if (m->patch) {
    const ohp_patch_plan *pl = m->patch;
    vw->n_patches = pl->n_patches; vw->patch_rows = pl->R; vw->patch_max_cells = pl->max_cells; vw->patch_max_verts = pl->max_verts;
    vw->prow_dev = pl->prow; vw->pstar_ptr_dev = pl->pstar_ptr; vw->pstar_dev = pl->pstar; vw->pslot_dev = pl->pslot;
    vw->pcell_ptr_dev = pl->pcell_ptr; vw->pconn_dev = pl->pconn; vw->pvert_ptr_dev = pl->pvert_ptr; vw->pvert_dev = pl->pvert;
  }
  return 0;
}
