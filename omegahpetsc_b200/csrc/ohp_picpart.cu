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
 *   K6 findIdxOfGid             ex12.cpp:748-761   O(nVerts x nGhosts) linear search: replaced by
 *                                                  ohp_ghost_roots_from_gids below, an open-addressing
 *                                                  hash join on the device between the two sparse
 *                                                  exchanges (Dist::exch, ex12.cpp:740,765)
 * together with the Omega_h primitives each_eq_to / offset_scan / get_sum they call
 * (ex12.cpp:587-588,629-630,647,717).
 */
#include <algorithm>
#include <vector>

#include "ohp_internal.h"

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)std::max<int64_t>(1, (n + bs - 1) / bs); }

/* a bad picpart file must not turn into out-of-bounds device writes (k_mark_core, k_compact_elems) */
__global__ void k_check_part(const int32_t *ev, int64_t n, int nV, int *flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (ev[i] < 0 || ev[i] >= nV)) atomicExch(flag, 1);
}

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
  PP_CUDA(OHP_MALLOC(ctx, &d_ev, sizeof(int32_t) * std::max<int64_t>(1, (int64_t)nE * nc)));
  PP_CUDA(OHP_MALLOC(ctx, &d_eown, sizeof(int32_t) * std::max(1, nE)));
  PP_CUDA(OHP_MALLOC(ctx, &d_vown, sizeof(int32_t) * std::max(1, nV)));
  PP_CUDA(OHP_MALLOC(ctx, &d_coords, sizeof(double) * std::max<int64_t>(1, (int64_t)nV * dim)));
  PP_CUDA(OHP_MALLOC(ctx, &d_keep, sizeof(int32_t) * (nE + 1)));
  PP_CUDA(OHP_MALLOC(ctx, &d_escan, sizeof(int32_t) * (nE + 1)));
  PP_CUDA(OHP_MALLOC(ctx, &d_core, sizeof(int32_t) * (nV + 1)));
  PP_CUDA(OHP_MALLOC(ctx, &d_vscan, sizeof(int32_t) * (nV + 1)));
  PP_CUDA(OHP_MALLOC(ctx, &d_p2c, sizeof(int32_t) * (nV + 1)));
  PP_CUDA(OHP_MALLOC(ctx, &d_gmask, sizeof(int32_t) * (nV + 1)));
  PP_CUDA(OHP_MALLOC(ctx, &d_omask, sizeof(int32_t) * (nV + 1)));
  PP_CUDA(OHP_MALLOC(ctx, &d_gscan, sizeof(int32_t) * (nV + 1)));
  PP_CUDA(OHP_MALLOC(ctx, &d_cev, sizeof(int32_t) * std::max<int64_t>(1, (int64_t)nE * nc)));
  PP_CUDA(OHP_MALLOC(ctx, &d_ccoords, sizeof(double) * std::max<int64_t>(1, (int64_t)nV * dim)));
  PP_CUDA(OHP_MALLOC(ctx, &d_cgid, sizeof(int64_t) * std::max(1, nV)));
  PP_CUDA(OHP_MALLOC(ctx, &d_gcore, sizeof(int32_t) * std::max(1, nV)));
  PP_CUDA(OHP_MALLOC(ctx, &d_ggid, sizeof(int64_t) * std::max(1, nV)));
  PP_CUDA(OHP_MALLOC(ctx, &d_gown, sizeof(int32_t) * std::max(1, nV)));
  if (vgid) {
    PP_CUDA(OHP_MALLOC(ctx, &d_gid, sizeof(int64_t) * std::max(1, nV)));
    PP_CUDA(cudaMemcpyAsync(d_gid, vgid, sizeof(int64_t) * nV, cudaMemcpyHostToDevice, st));
  }
  PP_CUDA(cudaMemcpyAsync(d_ev, ev, sizeof(int32_t) * (int64_t)nE * nc, cudaMemcpyHostToDevice, st));
  PP_CUDA(cudaMemcpyAsync(d_eown, eown, sizeof(int32_t) * nE, cudaMemcpyHostToDevice, st));
  PP_CUDA(cudaMemcpyAsync(d_vown, vown, sizeof(int32_t) * nV, cudaMemcpyHostToDevice, st));
  PP_CUDA(cudaMemcpyAsync(d_coords, coords, sizeof(double) * (int64_t)nV * dim, cudaMemcpyHostToDevice, st));
  PP_CUDA(cudaMemsetAsync(d_core, 0, sizeof(int32_t) * (nV + 1), st));
  if (nE) {
    int bad = 0;
    k_check_part<<<nblk((int64_t)nE * nc, BS), BS, 0, st>>>(d_ev, (int64_t)nE * nc, nV, ctx->d_flag);
    PP_LAUNCH();
    PP_CUDA(cudaMemcpyAsync(&bad, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaStreamSynchronize(st));
    if (bad) {
      cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), st);
      ohp_set_error("ohp_picpart_extract: an element references a vertex outside [0,%d)", nV);
      rc = OHP_ERR_ARG_OUTOFRANGE;
      goto cleanup;
    }
  }

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
  OHP_FREE(ctx, d_ev); OHP_FREE(ctx, d_eown); OHP_FREE(ctx, d_vown); OHP_FREE(ctx, d_keep); OHP_FREE(ctx, d_escan); OHP_FREE(ctx, d_core);
  OHP_FREE(ctx, d_vscan); OHP_FREE(ctx, d_p2c); OHP_FREE(ctx, d_gmask); OHP_FREE(ctx, d_omask); OHP_FREE(ctx, d_gscan); OHP_FREE(ctx, d_cev);
  OHP_FREE(ctx, d_gcore); OHP_FREE(ctx, d_gown); OHP_FREE(ctx, d_coords); OHP_FREE(ctx, d_ccoords); OHP_FREE(ctx, d_gid); OHP_FREE(ctx, d_cgid);
  OHP_FREE(ctx, d_ggid);
  return rc;
}


/* ------------------------------------------------------------------------------------ */
/* gid -> owner index: getPicPartCoreVtxOwnerIdx's round trip (ex12.cpp:731-777).         */
/*   1. every rank sends the gids of its ghosts to their owners   (Dist::exch, :740)      */
/*   2. the owner finds each gid among ITS vertices               (K6 findIdxOfGid, :748) */
/*   3. and sends the local indices back                          (distInv.exch, :765)    */
/* Step 2 is an open-addressing hash join on the device: the owned (gid -> index) pairs   */
/* are inserted once (linear probing, 64-bit atomicCAS on the key), every received gid is  */
/* looked up by one thread: O(n_core + n_received) instead of O(n_core x n_received).     */
/* ------------------------------------------------------------------------------------ */
namespace {
constexpr unsigned long long kEmptyKey = 0xffffffffffffffffull;
__device__ __forceinline__ uint64_t hash64(uint64_t x) { /* splitmix64 finaliser */
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
  return x;
}
__global__ void k_hash_insert(const int64_t *gid, const uint8_t *owned, int n, unsigned long long *keys, int32_t *vals,
                              uint64_t mask, int *dup) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (owned && !owned[i])) return;
  const unsigned long long key = (unsigned long long)gid[i];
  uint64_t h = hash64(key) & mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(keys + h, kEmptyKey, key);
    if (prev == kEmptyKey) { vals[h] = i; return; }
    if (prev == key) { atomicExch(dup, 1); return; } /* two owned vertices with the same gid */
    h = (h + 1) & mask;
  }
}
__global__ void k_hash_lookup(const int64_t *query, int n, const unsigned long long *keys, const int32_t *vals,
                              uint64_t mask, int32_t *out, int *missing) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = (unsigned long long)query[i];
  uint64_t h = hash64(key) & mask;
  for (;;) {
    const unsigned long long k = keys[h];
    if (k == key) { out[i] = vals[h]; return; }
    if (k == kEmptyKey) { out[i] = -1; atomicExch(missing, 1); return; }
    h = (h + 1) & mask;
  }
}
}  // namespace

extern "C" int ohp_ghost_roots_from_gids(ohp_context ctx, int32_t n_ghost, const int64_t *ghost_gid,
                                         const int32_t *ghost_owner, int32_t n_core, const int64_t *core_gid,
                                         const uint8_t *core_owned, int32_t *roots) {
  if (!ctx || (n_ghost && (!ghost_gid || !ghost_owner || !roots)) || (n_core && !core_gid))
    OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_ghost_roots_from_gids: null argument");
  if (n_ghost < 0 || n_core < 0) OHP_FAIL(OHP_ERR_ARG_SIZ, "ohp_ghost_roots_from_gids: negative size");
  const int P = ctx->nranks, BS = 256;
  if (P == 1) {
    if (n_ghost) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_ghost_roots_from_gids: %d ghosts but no communicator attached", n_ghost);
    return 0;
  }
  int rc = 0, status = 0;
  cudaStream_t st = ctx->stream;
  unsigned long long *d_keys = nullptr;
  int32_t *d_vals = nullptr, *d_ans = nullptr, *d_back = nullptr;
  int64_t *d_cgid = nullptr, *d_req = nullptr, *d_in = nullptr;
  uint8_t *d_owned = nullptr;
  std::vector<int64_t> scount(P, 0), rcount;
  std::vector<int64_t> soff(P + 1, 0);
  std::vector<int32_t> order(n_ghost);   /* ghosts grouped by owner rank (counting sort, stable) */
  std::vector<int64_t> req(n_ghost);
  std::vector<int32_t> back;
  int64_t n_in = 0;
  uint64_t cap = 64;
  PP_CUDA(cudaSetDevice(ctx->device));
  for (int i = 0; i < n_ghost; ++i) {
    if (ghost_owner[i] < 0 || ghost_owner[i] >= P || ghost_owner[i] == ctx->rank) status = 1; /* reported collectively below */
    else scount[ghost_owner[i]]++;
  }
  for (int q = 0; q < P; ++q) soff[q + 1] = soff[q] + scount[q];
  if (!status) {
    std::vector<int64_t> cur(soff.begin(), soff.end() - 1);
    for (int i = 0; i < n_ghost; ++i) { const int64_t k = cur[ghost_owner[i]]++; order[k] = i; req[k] = ghost_gid[i]; }
  } else {
    std::fill(scount.begin(), scount.end(), 0); /* take part in the collectives with nothing to send */
    std::fill(soff.begin(), soff.end(), 0);
  }
  /* hash table of my owned vertices */
  while (cap < (uint64_t)2 * std::max(1, n_core)) cap <<= 1;
  PP_CUDA(OHP_MALLOC(ctx, &d_keys, sizeof(unsigned long long) * cap));
  PP_CUDA(OHP_MALLOC(ctx, &d_vals, sizeof(int32_t) * cap));
  PP_CUDA(cudaMemsetAsync(d_keys, 0xff, sizeof(unsigned long long) * cap, st));
  PP_CUDA(OHP_MALLOC(ctx, &d_cgid, sizeof(int64_t) * std::max(1, n_core)));
  PP_CUDA(cudaMemcpyAsync(d_cgid, core_gid, sizeof(int64_t) * n_core, cudaMemcpyHostToDevice, st));
  if (core_owned) {
    PP_CUDA(OHP_MALLOC(ctx, &d_owned, std::max(1, n_core)));
    PP_CUDA(cudaMemcpyAsync(d_owned, core_owned, n_core, cudaMemcpyHostToDevice, st));
  }
  if (n_core) { k_hash_insert<<<nblk(n_core, BS), BS, 0, st>>>(d_cgid, d_owned, n_core, d_keys, d_vals, cap - 1, ctx->d_flag); PP_LAUNCH(); }
  /* 1. gids to the owners */
  PP_CUDA(OHP_MALLOC(ctx, &d_req, sizeof(int64_t) * std::max<int64_t>(1, soff[P])));
  PP_CUDA(cudaMemcpyAsync(d_req, req.data(), sizeof(int64_t) * soff[P], cudaMemcpyHostToDevice, st));
  PP_TRY(ohp_sparse_exchange(ctx, 8, d_req, scount, (void **)&d_in, rcount));
  for (int q = 0; q < P; ++q) n_in += rcount[q];
  /* 2. the join */
  PP_CUDA(OHP_MALLOC(ctx, &d_ans, sizeof(int32_t) * std::max<int64_t>(1, n_in)));
  if (n_in) { k_hash_lookup<<<nblk(n_in, BS), BS, 0, st>>>(d_in, (int)n_in, d_keys, d_vals, cap - 1, d_ans, ctx->d_flag + 0); PP_LAUNCH(); }
  {
    int flag = 0;
    PP_CUDA(cudaMemcpyAsync(&flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaStreamSynchronize(st));
    if (flag) { cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), st); status = 2; }
  }
  /* 3. indices back (the reverse exchange: what I received from q goes back to q) */
  {
    std::vector<int64_t> rc2;
    PP_TRY(ohp_sparse_exchange(ctx, 4, d_ans, rcount, (void **)&d_back, rc2));
    for (int q = 0; q < P; ++q)
      if (rc2[q] != scount[q]) { ohp_set_error("ohp_ghost_roots_from_gids: reply count mismatch from rank %d", q); rc = OHP_ERR_LIB; goto cleanup; }
  }
  back.resize(soff[P]);
  PP_CUDA(cudaMemcpyAsync(back.data(), d_back, sizeof(int32_t) * soff[P], cudaMemcpyDeviceToHost, st));
  PP_CUDA(cudaStreamSynchronize(st));
  for (int64_t k = 0; k < soff[P]; ++k) {
    roots[order[k]] = back[k];
    if (back[k] < 0) status = std::max(status, 3);
  }
  /* agree on the outcome: every rank returns the same code, none is left waiting in a collective */
  {
    double s = (double)status;
    PP_CUDA(cudaMemcpyAsync(ctx->d_scalars + 9, &s, sizeof(double), cudaMemcpyHostToDevice, st));
    PP_TRY(ohp_allreduce_sum(ctx, ctx->d_scalars + 9, 1));
    PP_CUDA(cudaMemcpyAsync(&s, ctx->d_scalars + 9, sizeof(double), cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaStreamSynchronize(st));
    if (s != 0.0) {
      ohp_set_error("ohp_ghost_roots_from_gids: inconsistent point SF (local status %d: 1 = ghost owner out of range or self, "
                    "2 = a rank asked this rank for a gid it does not own or two owned vertices share a gid, 3 = an owner does not hold one of this rank's ghosts)", status);
      rc = OHP_ERR_ARG_WRONG;
    }
  }
cleanup:
  cudaStreamSynchronize(st);
  OHP_FREE(ctx, d_keys); OHP_FREE(ctx, d_vals); OHP_FREE(ctx, d_ans); OHP_FREE(ctx, d_back); OHP_FREE(ctx, d_cgid); OHP_FREE(ctx, d_req); OHP_FREE(ctx, d_in); OHP_FREE(ctx, d_owned);
  return rc;
}
