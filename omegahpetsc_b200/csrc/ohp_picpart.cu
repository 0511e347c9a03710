/*
 * ohp_picpart.cu -- device-side extraction of a rank's "core" mesh from a pumipic picpart
 * and the gid -> owner-index join.
 *
 * Replaces the reference's adapter kernels (OMEGA_H_LAMBDAs compiled as CUDA):
 *   K1 getCoreCells2Verts       ex12.cpp:595-616   owned-element compaction + renumbering
 *   K2 markCoreVerts            ex12.cpp:636-644   vertices touched by owned elements
 *   K3 getCoordinatesAndMap     ex12.cpp:657-667   vertex compaction + part->core map
 *   K4 createCoreGhostMask      ex12.cpp:709-714   core && owner != rank
 *   K5 getGhostVtxInfo          ex12.cpp:720-730   compaction of (coreIdx, gid, owner)
 *   K6 findIdxOfGid             ex12.cpp:748-761   O(nVerts x nGhosts) linear search, replaced
 *                                                  here by an open-addressing hash join
 * together with the Omega_h primitives each_eq_to / offset_scan / get_sum they call
 * (ex12.cpp:587-588,629-630,647,717).
 */
#include <algorithm>

#include "ohp_internal.h"

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)std::max<int64_t>(1, (n + bs - 1) / bs); }

template <int NC>
__global__ void k_keep_elems(const int32_t *ev, const int32_t *eown, const int32_t *vown, int nE, int rank,
                             int overlap, int32_t *keep) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nE) return;
  int k = (eown[e] == rank);
  if (overlap) {
#pragma unroll
    for (int j = 0; j < NC; ++j) k |= (vown[ev[(size_t)e * NC + j]] == rank);
  }
  keep[e] = k;
}
template <int NC>
__global__ void k_mark_core(const int32_t *ev, const int32_t *keep, int nE, int32_t *is_core) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nE || !keep[e]) return;
#pragma unroll
  for (int j = 0; j < NC; ++j) is_core[ev[(size_t)e * NC + j]] = 1; /* same-value race, ex12.cpp:640 */
}
template <int DIM>
__global__ void k_compact_verts(const int32_t *is_core, const int32_t *scan, const double *coords,
                                const int64_t *gid, const int32_t *vown, int nV, int rank, double *ccoords,
                                int64_t *cgid, int32_t *part2core, int32_t *ghost_mask, int32_t *owned_mask) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV) return;
  const int core = is_core[v];
  part2core[v] = core ? scan[v] : -1;
  ghost_mask[v] = core && vown[v] != rank;
  owned_mask[v] = core && vown[v] == rank;
  if (core) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) ccoords[(size_t)scan[v] * DIM + d] = coords[(size_t)v * DIM + d];
    cgid[scan[v]] = gid ? gid[v] : (int64_t)v;
  }
}
template <int NC>
__global__ void k_compact_elems(const int32_t *ev, const int32_t *keep, const int32_t *escan,
                                const int32_t *part2core, int nE, int32_t *cev) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nE || !keep[e]) return;
#pragma unroll
  for (int j = 0; j < NC; ++j) cev[(size_t)escan[e] * NC + j] = part2core[ev[(size_t)e * NC + j]];
}
__global__ void k_compact_ghosts(const int32_t *ghost_mask, const int32_t *gscan, const int32_t *part2core,
                                 const int64_t *gid, const int32_t *vown, int nV, int32_t *g_core, int64_t *g_gid,
                                 int32_t *g_owner) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV || !ghost_mask[v]) return;
  const int k = gscan[v];
  g_core[k] = part2core[v];
  g_gid[k] = gid ? gid[v] : (int64_t)v;
  g_owner[k] = vown[v];
}

#define PP_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      ohp_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      rc = OHP_ERR_LIB;                                                                \
      goto cleanup;                                                                    \
    }                                                                                  \
  } while (0)
#define PP_TRY(expr) do { rc = (expr); if (rc) goto cleanup; } while (0)
#define PP_LAUNCH() do { ctx->launches++; PP_CUDA(cudaGetLastError()); } while (0)

extern "C" int ohp_picpart_extract(ohp_context ctx, int dim, int rank, int with_overlap, int32_t nE, int32_t nV,
                                   const int32_t *ev, const double *coords, const int32_t *eown,
                                   const int32_t *vown, const int64_t *vgid, ohp_core_counts *counts,
                                   int32_t *core_ev, double *core_coords, int32_t *part2core,
                                   int32_t *ghost_core_idx, int64_t *ghost_gid, int32_t *ghost_owner,
                                   int64_t *core_gid) {
  if (!ctx || !ev || !coords || !eown || !vown || !counts) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_picpart_extract: null argument");
  if (dim != 2 && dim != 3) OHP_FAIL(OHP_ERR_SUP, "ohp_picpart_extract: dim %d", dim);
  const int nc = dim + 1, BS = 256;
  int rc = 0;
  cudaStream_t st = ctx->stream;
  int32_t *d_ev = nullptr, *d_eown = nullptr, *d_vown = nullptr, *d_keep = nullptr, *d_escan = nullptr,
          *d_core = nullptr, *d_vscan = nullptr, *d_p2c = nullptr, *d_gmask = nullptr, *d_omask = nullptr,
          *d_gscan = nullptr, *d_cev = nullptr, *d_gcore = nullptr, *d_gown = nullptr;
  double *d_coords = nullptr, *d_ccoords = nullptr;
  int64_t *d_gid = nullptr, *d_cgid = nullptr, *d_ggid = nullptr;
  int64_t nCE = 0, nCV = 0, nG = 0, nO = 0;
  PP_CUDA(cudaSetDevice(ctx->device));
  PP_CUDA(cudaMalloc(&d_ev, sizeof(int32_t) * std::max<int64_t>(1, (int64_t)nE * nc)));
  PP_CUDA(cudaMalloc(&d_eown, sizeof(int32_t) * std::max(1, nE)));
  PP_CUDA(cudaMalloc(&d_vown, sizeof(int32_t) * std::max(1, nV)));
  PP_CUDA(cudaMalloc(&d_coords, sizeof(double) * std::max<int64_t>(1, (int64_t)nV * dim)));
  PP_CUDA(cudaMalloc(&d_keep, sizeof(int32_t) * (nE + 1)));
  PP_CUDA(cudaMalloc(&d_escan, sizeof(int32_t) * (nE + 1)));
  PP_CUDA(cudaMalloc(&d_core, sizeof(int32_t) * (nV + 1)));
  PP_CUDA(cudaMalloc(&d_vscan, sizeof(int32_t) * (nV + 1)));
  PP_CUDA(cudaMalloc(&d_p2c, sizeof(int32_t) * (nV + 1)));
  PP_CUDA(cudaMalloc(&d_gmask, sizeof(int32_t) * (nV + 1)));
  PP_CUDA(cudaMalloc(&d_omask, sizeof(int32_t) * (nV + 1)));
  PP_CUDA(cudaMalloc(&d_gscan, sizeof(int32_t) * (nV + 1)));
  PP_CUDA(cudaMalloc(&d_cev, sizeof(int32_t) * std::max<int64_t>(1, (int64_t)nE * nc)));
  PP_CUDA(cudaMalloc(&d_ccoords, sizeof(double) * std::max<int64_t>(1, (int64_t)nV * dim)));
  PP_CUDA(cudaMalloc(&d_cgid, sizeof(int64_t) * std::max(1, nV)));
  PP_CUDA(cudaMalloc(&d_gcore, sizeof(int32_t) * std::max(1, nV)));
  PP_CUDA(cudaMalloc(&d_ggid, sizeof(int64_t) * std::max(1, nV)));
  PP_CUDA(cudaMalloc(&d_gown, sizeof(int32_t) * std::max(1, nV)));
  if (vgid) {
    PP_CUDA(cudaMalloc(&d_gid, sizeof(int64_t) * std::max(1, nV)));
    PP_CUDA(cudaMemcpyAsync(d_gid, vgid, sizeof(int64_t) * nV, cudaMemcpyHostToDevice, st));
  }
  PP_CUDA(cudaMemcpyAsync(d_ev, ev, sizeof(int32_t) * (int64_t)nE * nc, cudaMemcpyHostToDevice, st));
  PP_CUDA(cudaMemcpyAsync(d_eown, eown, sizeof(int32_t) * nE, cudaMemcpyHostToDevice, st));
  PP_CUDA(cudaMemcpyAsync(d_vown, vown, sizeof(int32_t) * nV, cudaMemcpyHostToDevice, st));
  PP_CUDA(cudaMemcpyAsync(d_coords, coords, sizeof(double) * (int64_t)nV * dim, cudaMemcpyHostToDevice, st));
  PP_CUDA(cudaMemsetAsync(d_core, 0, sizeof(int32_t) * (nV + 1), st));

  if (nc == 3) k_keep_elems<3><<<nblk(nE, BS), BS, 0, st>>>(d_ev, d_eown, d_vown, nE, rank, with_overlap, d_keep);
  else k_keep_elems<4><<<nblk(nE, BS), BS, 0, st>>>(d_ev, d_eown, d_vown, nE, rank, with_overlap, d_keep);
  PP_LAUNCH();
  PP_TRY(ohp_scan_exclusive_i32(ctx, d_keep, d_escan, nE, &nCE));
  if (nc == 3) k_mark_core<3><<<nblk(nE, BS), BS, 0, st>>>(d_ev, d_keep, nE, d_core);
  else k_mark_core<4><<<nblk(nE, BS), BS, 0, st>>>(d_ev, d_keep, nE, d_core);
  PP_LAUNCH();
  PP_TRY(ohp_scan_exclusive_i32(ctx, d_core, d_vscan, nV, &nCV));
  if (dim == 2) k_compact_verts<2><<<nblk(nV, BS), BS, 0, st>>>(d_core, d_vscan, d_coords, d_gid, d_vown, nV, rank, d_ccoords, d_cgid, d_p2c, d_gmask, d_omask);
  else k_compact_verts<3><<<nblk(nV, BS), BS, 0, st>>>(d_core, d_vscan, d_coords, d_gid, d_vown, nV, rank, d_ccoords, d_cgid, d_p2c, d_gmask, d_omask);
  PP_LAUNCH();
  if (nc == 3) k_compact_elems<3><<<nblk(nE, BS), BS, 0, st>>>(d_ev, d_keep, d_escan, d_p2c, nE, d_cev);
  else k_compact_elems<4><<<nblk(nE, BS), BS, 0, st>>>(d_ev, d_keep, d_escan, d_p2c, nE, d_cev);
  PP_LAUNCH();
  PP_TRY(ohp_scan_exclusive_i32(ctx, d_omask, d_gscan, nV, &nO));
  PP_TRY(ohp_scan_exclusive_i32(ctx, d_gmask, d_gscan, nV, &nG));
  k_compact_ghosts<<<nblk(nV, BS), BS, 0, st>>>(d_gmask, d_gscan, d_p2c, d_gid, d_vown, nV, d_gcore, d_ggid, d_gown);
  PP_LAUNCH();

  counts->n_core_elems = (int32_t)nCE; counts->n_core_verts = (int32_t)nCV;
  counts->n_owned_verts = (int32_t)nO; counts->n_ghost_verts = (int32_t)nG;
  if (core_ev) PP_CUDA(cudaMemcpyAsync(core_ev, d_cev, sizeof(int32_t) * nCE * nc, cudaMemcpyDeviceToHost, st));
  if (core_coords) PP_CUDA(cudaMemcpyAsync(core_coords, d_ccoords, sizeof(double) * nCV * dim, cudaMemcpyDeviceToHost, st));
  if (part2core) PP_CUDA(cudaMemcpyAsync(part2core, d_p2c, sizeof(int32_t) * nV, cudaMemcpyDeviceToHost, st));
  if (core_gid) PP_CUDA(cudaMemcpyAsync(core_gid, d_cgid, sizeof(int64_t) * nCV, cudaMemcpyDeviceToHost, st));
  if (ghost_core_idx) PP_CUDA(cudaMemcpyAsync(ghost_core_idx, d_gcore, sizeof(int32_t) * nG, cudaMemcpyDeviceToHost, st));
  if (ghost_gid) PP_CUDA(cudaMemcpyAsync(ghost_gid, d_ggid, sizeof(int64_t) * nG, cudaMemcpyDeviceToHost, st));
  if (ghost_owner) PP_CUDA(cudaMemcpyAsync(ghost_owner, d_gown, sizeof(int32_t) * nG, cudaMemcpyDeviceToHost, st));
  PP_CUDA(cudaStreamSynchronize(st));
cleanup:
  cudaStreamSynchronize(st);
  cudaFree(d_ev); cudaFree(d_eown); cudaFree(d_vown); cudaFree(d_keep); cudaFree(d_escan); cudaFree(d_core);
  cudaFree(d_vscan); cudaFree(d_p2c); cudaFree(d_gmask); cudaFree(d_omask); cudaFree(d_gscan); cudaFree(d_cev);
  cudaFree(d_gcore); cudaFree(d_gown); cudaFree(d_coords); cudaFree(d_ccoords); cudaFree(d_gid); cudaFree(d_cgid);
  cudaFree(d_ggid);
  return rc;
}
