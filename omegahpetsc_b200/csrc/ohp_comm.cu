/*
 * ohp_comm.cu -- communicator, halo plans and halo exchange for the partitioned path.
 *
 * Replaces MPI as used through Omega_h::Comm/Dist (ex12.cpp:530-563,731-765) and PETSc's PetscSF
 * (ex12.cpp:873-876,937-939): the point SF given to ohp_mesh_set_ghosts becomes
 *   (1) a vertex-level plan (ghost vertex <- owner vertex) used during setup to fetch the owners'
 *       boundary labels and dof numbers (PetscSFBcast of the section, PETSc-internal), and
 *   (2) a dof-level plan for vectors: ghost dofs are numbered after the owned ones, grouped by
 *       owner rank and ordered by the owner's dof number, so every receive lands directly in the
 *       vector's ghost tail (DMGlobalToLocal / the MPIAIJ MatMult scatter, PETSc-internal);
 * dot products are summed with an all-reduce (VecDot's MPI_Allreduce).
 * One process per GPU; transport is NCCL over NVLink/NVSwitch, loaded with dlopen so that the
 * library shares whatever libnccl the host process (e.g. torch) already mapped.
 */
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>

#include "ohp_internal.h"

namespace {
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int nccl_load() {
  if (g_nccl.lib) return 0;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) OHP_FAIL(OHP_ERR_LIB, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                        \
  do {                                                                          \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                 \
    if (!g_nccl.field) OHP_FAIL(OHP_ERR_LIB, "libnccl: missing symbol %s", name); \
  } while (0)
  SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce"); SYM(AllGather, "ncclAllGather"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.lib = h;
  return 0;
}
}  // namespace

#define OHP_NCCL(expr)                                                                       \
  do {                                                                                       \
    ncclResult_t r__ = (expr);                                                               \
    if (r__ != ncclSuccess)                                                                  \
      OHP_FAIL(OHP_ERR_LIB, "%s:%d NCCL error %s: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(r__)); \
  } while (0)

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)std::max<int64_t>(1, (n + bs - 1) / bs); }

/* an NCCL group that is closed on EVERY exit path: an error raised between GroupStart and GroupEnd must not leave the
 * group open (every later NCCL call of this process would be queued into it) */
namespace {
struct NcclGroup {
  bool open = false;
  ncclResult_t start() { const ncclResult_t r = g_nccl.GroupStart(); open = (r == ncclSuccess); return r; }
  ncclResult_t end() { open = false; return g_nccl.GroupEnd(); }
  ~NcclGroup() { if (open) g_nccl.GroupEnd(); }
};
}  // namespace

extern "C" int ohp_comm_unique_id(void *out) {
  if (!out) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_comm_unique_id: null buffer");
  OHP_TRY(nccl_load());
  ncclUniqueId id;
  OHP_NCCL(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(out, &id, sizeof(id));
  return 0;
}

extern "C" int ohp_comm_init(ohp_context ctx, int nranks, int rank, const void *uid) {
  if (!ctx || !uid) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_comm_init: null argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "ohp_comm_init: rank %d of %d", rank, nranks);
  if (ctx->comm) OHP_FAIL(OHP_ERR_ORDER, "ohp_comm_init: communicator already attached");
  OHP_TRY(nccl_load());
  OHP_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, uid, sizeof(id));
  ncclComm_t comm;
  OHP_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
  ctx->comm = comm; ctx->rank = rank; ctx->nranks = nranks;
  return 0;
}

extern "C" int ohp_comm_rank_size(ohp_context ctx, int *rank, int *nranks) {
  *rank = ctx->rank; *nranks = ctx->nranks;
  return 0;
}

void ohp_comm_destroy(ohp_context ctx) {
  ohp_p2p_cache_drop(ctx);
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
}

int ohp_allreduce_sum(ohp_context ctx, double *d_vals, int n) {
  if (ctx->nranks == 1) return 0;
  OHP_NCCL(g_nccl.AllReduce(d_vals, d_vals, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
  return 0;
}

int ohp_allgather_i64(ohp_context ctx, int64_t mine, std::vector<int64_t> &all) {
  all.assign(ctx->nranks, mine);
  if (ctx->nranks == 1) return 0;
  int64_t *d;
  OHP_CUDA(OHP_MALLOC(ctx, &d, sizeof(int64_t) * (ctx->nranks + 1)));
  OHP_CUDA(cudaMemcpyAsync(d + ctx->nranks, &mine, sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  OHP_NCCL(g_nccl.AllGather(d + ctx->nranks, d, 1, ncclInt64, (ncclComm_t)ctx->comm, ctx->stream));
  OHP_CUDA(cudaMemcpyAsync(all.data(), d, sizeof(int64_t) * ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream));
  OHP_CUDA(cudaStreamSynchronize(ctx->stream));
  OHP_FREE(ctx, d);
  return 0;
}

/* MPI_Barrier / MPI_Allgather / MPI_Allreduce as the adapter uses them (ex12.cpp:573-575,814,844,953) */
extern "C" int ohp_comm_barrier(ohp_context ctx) {
  if (!ctx) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_comm_barrier: null context");
  OHP_CUDA(cudaSetDevice(ctx->device));
  if (ctx->nranks > 1) {
    OHP_CUDA(cudaMemsetAsync(ctx->d_scalars + 11, 0, sizeof(double), ctx->stream));
    OHP_NCCL(g_nccl.AllReduce(ctx->d_scalars + 11, ctx->d_scalars + 11, 1, ncclFloat64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
  }
  OHP_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int ohp_comm_allgather_i64(ohp_context ctx, int64_t mine, int64_t *all) {
  if (!ctx || !all) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_comm_allgather_i64: null argument");
  OHP_CUDA(cudaSetDevice(ctx->device));
  std::vector<int64_t> v;
  OHP_TRY(ohp_allgather_i64(ctx, mine, v));
  std::copy(v.begin(), v.end(), all);
  return 0;
}
extern "C" int ohp_comm_allreduce_sum_i64(ohp_context ctx, int64_t *vals, int n) {
  if (!ctx || (n > 0 && !vals)) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_comm_allreduce_sum_i64: null argument");
  if (ctx->nranks == 1 || n <= 0) return 0;
  OHP_CUDA(cudaSetDevice(ctx->device));
  int64_t *d;
  OHP_CUDA(OHP_MALLOC(ctx, &d, sizeof(int64_t) * n));
  OHP_CUDA(cudaMemcpyAsync(d, vals, sizeof(int64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  ncclResult_t r = g_nccl.AllReduce(d, d, (size_t)n, ncclInt64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream);
  if (r == ncclSuccess) {
    cudaMemcpyAsync(vals, d, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
  }
  OHP_FREE(ctx, d);
  OHP_NCCL(r);
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* sparse neighbour exchange of int32 lists (setup only): the role of Omega_h's           */
/* Dist::exch (ex12.cpp:562,740,765).  send[q] goes to rank q; returns recv[q] from q.    */
/* ------------------------------------------------------------------------------------ */
static int exchange_lists(ohp_context ctx, const std::vector<std::vector<int32_t>> &send,
                          std::vector<std::vector<int32_t>> &recv) {
  const int P = ctx->nranks;
  cudaStream_t st = ctx->stream;
  std::vector<int64_t> counts(P), matrix((size_t)P * P);
  int64_t *d_counts;
  OHP_CUDA(OHP_MALLOC(ctx, &d_counts, sizeof(int64_t) * ((size_t)P * P + P)));
  for (int q = 0; q < P; ++q) counts[q] = (int64_t)send[q].size();
  OHP_CUDA(cudaMemcpyAsync(d_counts + (size_t)P * P, counts.data(), sizeof(int64_t) * P, cudaMemcpyHostToDevice, st));
  OHP_NCCL(g_nccl.AllGather(d_counts + (size_t)P * P, d_counts, P, ncclInt64, (ncclComm_t)ctx->comm, st));
  OHP_CUDA(cudaMemcpyAsync(matrix.data(), d_counts, sizeof(int64_t) * P * P, cudaMemcpyDeviceToHost, st));
  OHP_CUDA(cudaStreamSynchronize(st));
  OHP_FREE(ctx, d_counts);
  int64_t ns = 0, nr = 0;
  std::vector<int64_t> soff(P + 1, 0), roff(P + 1, 0);
  for (int q = 0; q < P; ++q) { soff[q + 1] = soff[q] + matrix[(size_t)ctx->rank * P + q]; roff[q + 1] = roff[q] + matrix[(size_t)q * P + ctx->rank]; }
  ns = soff[P]; nr = roff[P];
  int32_t *d_s, *d_r;
  OHP_CUDA(OHP_MALLOC(ctx, &d_s, sizeof(int32_t) * std::max<int64_t>(1, ns)));
  OHP_CUDA(OHP_MALLOC(ctx, &d_r, sizeof(int32_t) * std::max<int64_t>(1, nr)));
  std::vector<int32_t> flat((size_t)ns);
  for (int q = 0; q < P; ++q) std::copy(send[q].begin(), send[q].end(), flat.begin() + soff[q]);
  OHP_CUDA(cudaMemcpyAsync(d_s, flat.data(), sizeof(int32_t) * ns, cudaMemcpyHostToDevice, st));
  NcclGroup grp;
  OHP_NCCL(grp.start());
  for (int q = 0; q < P; ++q) {
    if (q == ctx->rank) continue;
    if (soff[q + 1] > soff[q]) OHP_NCCL(g_nccl.Send(d_s + soff[q], (size_t)(soff[q + 1] - soff[q]), ncclInt32, q, (ncclComm_t)ctx->comm, st));
    if (roff[q + 1] > roff[q]) OHP_NCCL(g_nccl.Recv(d_r + roff[q], (size_t)(roff[q + 1] - roff[q]), ncclInt32, q, (ncclComm_t)ctx->comm, st));
  }
  OHP_NCCL(grp.end());
  std::vector<int32_t> rflat((size_t)nr);
  OHP_CUDA(cudaMemcpyAsync(rflat.data(), d_r, sizeof(int32_t) * nr, cudaMemcpyDeviceToHost, st));
  OHP_CUDA(cudaStreamSynchronize(st));
  OHP_FREE(ctx, d_s); OHP_FREE(ctx, d_r);
  recv.assign(P, {});
  for (int q = 0; q < P; ++q) recv[q].assign(rflat.begin() + roff[q], rflat.begin() + roff[q + 1]);
  return 0;
}

/* The same exchange on DEVICE buffers of `elem_bytes`-sized items: d_send holds the items grouped by
 * destination rank (scount[q] items for rank q, rank order); *d_recv is allocated here (cudaFree it) and holds
 * what the ranks sent to me, grouped by source rank (rcount[q] items from q).  Collective. */
int ohp_sparse_exchange(ohp_context ctx, int elem_bytes, const void *d_send, const std::vector<int64_t> &scount,
                        void **d_recv, std::vector<int64_t> &rcount) {
  const int P = ctx->nranks;
  cudaStream_t st = ctx->stream;
  std::vector<int64_t> matrix((size_t)P * P);
  int64_t *d_counts;
  OHP_CUDA(OHP_MALLOC(ctx, &d_counts, sizeof(int64_t) * ((size_t)P * P + P)));
  OHP_CUDA(cudaMemcpyAsync(d_counts + (size_t)P * P, scount.data(), sizeof(int64_t) * P, cudaMemcpyHostToDevice, st));
  ncclResult_t r = g_nccl.AllGather(d_counts + (size_t)P * P, d_counts, P, ncclInt64, (ncclComm_t)ctx->comm, st);
  if (r == ncclSuccess) {
    cudaMemcpyAsync(matrix.data(), d_counts, sizeof(int64_t) * P * P, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
  }
  OHP_FREE(ctx, d_counts);
  OHP_NCCL(r);
  rcount.assign(P, 0);
  std::vector<int64_t> soff(P + 1, 0), roff(P + 1, 0);
  for (int q = 0; q < P; ++q) {
    rcount[q] = matrix[(size_t)q * P + ctx->rank];
    soff[q + 1] = soff[q] + scount[q];
    roff[q + 1] = roff[q] + rcount[q];
  }
  char *recv = nullptr;
  OHP_CUDA(OHP_MALLOC(ctx, &recv, (size_t)std::max<int64_t>(1, roff[P]) * elem_bytes));
  *d_recv = recv;
  const char *send = (const char *)d_send;
  ncclResult_t first = ncclSuccess;
  g_nccl.GroupStart();
  for (int q = 0; q < P; ++q) {
    if (q == ctx->rank) {
      if (scount[q]) cudaMemcpyAsync(recv + roff[q] * elem_bytes, send + soff[q] * elem_bytes, (size_t)scount[q] * elem_bytes, cudaMemcpyDeviceToDevice, st);
      continue;
    }
    if (scount[q]) { r = g_nccl.Send(send + soff[q] * elem_bytes, (size_t)scount[q] * elem_bytes, ncclChar, q, (ncclComm_t)ctx->comm, st); if (r != ncclSuccess && first == ncclSuccess) first = r; }
    if (rcount[q]) { r = g_nccl.Recv(recv + roff[q] * elem_bytes, (size_t)rcount[q] * elem_bytes, ncclChar, q, (ncclComm_t)ctx->comm, st); if (r != ncclSuccess && first == ncclSuccess) first = r; }
  }
  r = g_nccl.GroupEnd(); /* always closed, whatever happened inside */
  if (first == ncclSuccess) first = r;
  OHP_NCCL(first);
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* halo plans                                                                            */
/* ------------------------------------------------------------------------------------ */
struct ohp_halo {
  /* vertex level: neighbour q sends me the values of its vertices v_req[q]; I send v_serve[q] */
  std::vector<int> nbr;                      /* ranks I exchange with (either direction) */
  std::vector<int64_t> v_soff, v_roff;       /* per neighbour offsets into the flat lists */
  int32_t *d_v_send_idx = nullptr;           /* my vertices other ranks ghost */
  int32_t *d_v_recv_idx = nullptr;           /* my ghost vertices, grouped by owner */
  int32_t *d_v_sbuf = nullptr, *d_v_rbuf = nullptr;
  int64_t v_ns = 0, v_nr = 0;
  /* dof level */
  std::vector<int> d_nbr;
  std::vector<int64_t> d_soff, d_roff;       /* send offsets into d_send_ldof; recv offsets into the ghost tail */
  int32_t *d_send_ldof = nullptr;
  double *d_sbuf = nullptr;
  int64_t d_ns = 0, d_nr = 0;
  std::vector<int64_t> d_dst_off;            /* per neighbour: first index of my block in ITS vector */
};

void ohp_halo_free(ohp_context ctx, ohp_halo *h) {
  if (!h) return;
  OHP_FREE(ctx, h->d_v_send_idx); OHP_FREE(ctx, h->d_v_recv_idx); OHP_FREE(ctx, h->d_v_sbuf); OHP_FREE(ctx, h->d_v_rbuf);
  OHP_FREE(ctx, h->d_send_ldof); OHP_FREE(ctx, h->d_sbuf);
  delete h;
}

static int vertex_plan_build(ohp_mesh m) {
  ohp_context ctx = m->ctx;
  const int P = ctx->nranks;
  ohp_halo *h = new ohp_halo();
  m->halo = h;
  const int nleaf = (int)m->ghost_vertex.size();
  /* my requests, grouped by owner rank, ordered by the owner's vertex index */
  std::vector<int> order(nleaf);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    if (m->ghost_owner_rank[a] != m->ghost_owner_rank[b]) return m->ghost_owner_rank[a] < m->ghost_owner_rank[b];
    if (m->ghost_owner_vertex[a] != m->ghost_owner_vertex[b]) return m->ghost_owner_vertex[a] < m->ghost_owner_vertex[b];
    return m->ghost_vertex[a] < m->ghost_vertex[b];
  });
  std::vector<std::vector<int32_t>> req(P), serve;
  std::vector<int32_t> recv_idx;
  recv_idx.reserve(nleaf);
  for (int i : order) { req[m->ghost_owner_rank[i]].push_back(m->ghost_owner_vertex[i]); recv_idx.push_back(m->ghost_vertex[i]); }
  OHP_TRY(exchange_lists(ctx, req, serve));
  std::vector<int32_t> send_idx;
  h->v_soff.assign(1, 0); h->v_roff.assign(1, 0);
  for (int q = 0; q < P; ++q) {
    if (req[q].empty() && serve[q].empty()) continue;
    for (int32_t v : serve[q]) {
      if (v < 0 || v >= m->nV) OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "point SF: rank %d asks for vertex %d of rank %d (has %d)", q, v, ctx->rank, m->nV);
      send_idx.push_back(v);
    }
    h->nbr.push_back(q);
    h->v_soff.push_back((int64_t)send_idx.size());
    h->v_roff.push_back(h->v_roff.back() + (int64_t)req[q].size());
  }
  h->v_ns = (int64_t)send_idx.size(); h->v_nr = (int64_t)recv_idx.size();
  OHP_CUDA(OHP_MALLOC(ctx, &h->d_v_send_idx, sizeof(int32_t) * std::max<int64_t>(1, h->v_ns)));
  OHP_CUDA(OHP_MALLOC(ctx, &h->d_v_recv_idx, sizeof(int32_t) * std::max<int64_t>(1, h->v_nr)));
  OHP_CUDA(OHP_MALLOC(ctx, &h->d_v_sbuf, sizeof(int32_t) * std::max<int64_t>(1, h->v_ns)));
  OHP_CUDA(OHP_MALLOC(ctx, &h->d_v_rbuf, sizeof(int32_t) * std::max<int64_t>(1, h->v_nr)));
  OHP_CUDA(cudaMemcpyAsync(h->d_v_send_idx, send_idx.data(), sizeof(int32_t) * h->v_ns, cudaMemcpyHostToDevice, ctx->stream));
  OHP_CUDA(cudaMemcpyAsync(h->d_v_recv_idx, recv_idx.data(), sizeof(int32_t) * h->v_nr, cudaMemcpyHostToDevice, ctx->stream));
  OHP_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

__global__ void k_gather_i32(const int32_t *src, const int32_t *idx, int64_t n, int32_t *out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}
__global__ void k_scatter_i32(int32_t *dst, const int32_t *idx, int64_t n, const int32_t *in) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = in[i];
}
__global__ void k_gather_f64(const double *__restrict__ src, const int32_t *__restrict__ idx, int64_t n, double *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}

/* ghost entries of a per-vertex int32 array <- the owners' entries */
int ohp_vertex_halo_i32(ohp_mesh m, int32_t *data) {
  ohp_context ctx = m->ctx;
  if (ctx->nranks == 1) return 0;
  if (!m->halo) OHP_TRY(vertex_plan_build(m));
  ohp_halo *h = m->halo;
  cudaStream_t st = ctx->stream;
  if (h->v_ns) { k_gather_i32<<<nblk(h->v_ns, 256), 256, 0, st>>>(data, h->d_v_send_idx, h->v_ns, h->d_v_sbuf); OHP_CHECK_LAUNCH(ctx); }
  NcclGroup grp;
  OHP_NCCL(grp.start());
  for (size_t k = 0; k < h->nbr.size(); ++k) {
    const int64_t ns = h->v_soff[k + 1] - h->v_soff[k], nr = h->v_roff[k + 1] - h->v_roff[k];
    if (ns) OHP_NCCL(g_nccl.Send(h->d_v_sbuf + h->v_soff[k], (size_t)ns, ncclInt32, h->nbr[k], (ncclComm_t)ctx->comm, st));
    if (nr) OHP_NCCL(g_nccl.Recv(h->d_v_rbuf + h->v_roff[k], (size_t)nr, ncclInt32, h->nbr[k], (ncclComm_t)ctx->comm, st));
  }
  OHP_NCCL(grp.end());
  if (h->v_nr) { k_scatter_i32<<<nblk(h->v_nr, 256), 256, 0, st>>>(data, h->d_v_recv_idx, h->v_nr, h->d_v_rbuf); OHP_CHECK_LAUNCH(ctx); }
  return 0;
}

__global__ void k_ldof_gid(int64_t *gid, int n_owned, int64_t start, const int64_t *ghost_gid, int n_ghost) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_owned) gid[i] = start + i;
  else if (i < n_owned + n_ghost) gid[i] = ghost_gid[i - n_owned];
}

/* global numbering across ranks + ghost dofs + the dof-level plan (called from ohp_mesh_setup
 * once lidx holds the owned rows) */
int ohp_halo_build(ohp_mesh m) {
  ohp_context ctx = m->ctx;
  const int P = ctx->nranks;
  cudaStream_t st = ctx->stream;
  if (!m->halo) OHP_TRY(vertex_plan_build(m));
  ohp_halo *h = m->halo;
  /* rank offsets: exclusive scan of the owned counts (global section, ranks concatenated) */
  std::vector<int64_t> all;
  OHP_TRY(ohp_allgather_i64(ctx, m->n_owned, all));
  std::vector<int64_t> start(P + 1, 0);
  for (int q = 0; q < P; ++q) start[q + 1] = start[q] + all[q];
  m->row_start = start[ctx->rank]; m->n_global = start[P];
  /* owners' local dof numbers for my ghost vertices */
  int32_t *d_tmp;
  OHP_CUDA(OHP_MALLOC(ctx, &d_tmp, sizeof(int32_t) * std::max(1, m->nV)));
  OHP_CUDA(cudaMemcpyAsync(d_tmp, m->lidx, sizeof(int32_t) * m->nV, cudaMemcpyDeviceToDevice, st));
  OHP_TRY(ohp_vertex_halo_i32(m, d_tmp));
  const int nleaf = (int)m->ghost_vertex.size();
  std::vector<int32_t> owner_ldof_all(m->nV);
  OHP_CUDA(cudaMemcpyAsync(owner_ldof_all.data(), d_tmp, sizeof(int32_t) * m->nV, cudaMemcpyDeviceToHost, st));
  OHP_CUDA(cudaStreamSynchronize(st));
  OHP_FREE(ctx, d_tmp);
  {
    /* consistency of the point SF: a ghost that is FREE here (its label came from the owner) must be a free, OWNED vertex on
     * its owner.  A root that points at a vertex the owner does not own has no row there (-1) and would silently be treated
     * as constrained.  Every rank learns the global count before any of them returns, so that nobody is left waiting in the
     * next collective. */
    std::vector<uint8_t> bnd(std::max(1, m->nV));
    OHP_CUDA(cudaMemcpyAsync(bnd.data(), m->bnd, m->nV, cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
    double bad = 0.0;
    int first_bad = -1;
    for (int i = 0; i < nleaf; ++i) {
      const int32_t gv = m->ghost_vertex[i];
      if (owner_ldof_all[gv] < 0 && bnd[gv] == 0) { bad += 1.0; if (first_bad < 0) first_bad = i; }
    }
    OHP_CUDA(cudaMemcpyAsync(ctx->d_scalars + 13, &bad, sizeof(double), cudaMemcpyHostToDevice, st));
    OHP_TRY(ohp_allreduce_sum(ctx, ctx->d_scalars + 13, 1));
    double total = 0.0;
    OHP_CUDA(cudaMemcpyAsync(&total, ctx->d_scalars + 13, sizeof(double), cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
    if (total > 0.0) {
      if (first_bad >= 0)
        OHP_FAIL(OHP_ERR_ARG_WRONG, "point SF: ghost vertex %d is free but its root (rank %d, vertex %d) is not an owned free vertex there (%g such ghosts over all ranks)",
                 m->ghost_vertex[first_bad], m->ghost_owner_rank[first_bad], m->ghost_owner_vertex[first_bad], total);
      OHP_FAIL(OHP_ERR_ARG_WRONG, "point SF: %g ghosts on other ranks point at roots that are not owned free vertices", total);
    }
  }
  /* unconstrained ghosts, ordered by (owner rank, owner dof): contiguous receive ranges */
  struct G { int rank; int32_t ldof; int32_t vertex; };
  std::vector<G> gh;
  for (int i = 0; i < nleaf; ++i) {
    const int32_t l = owner_ldof_all[m->ghost_vertex[i]];
    if (l >= 0) gh.push_back({m->ghost_owner_rank[i], l, m->ghost_vertex[i]});
  }
  std::sort(gh.begin(), gh.end(), [](const G &a, const G &b) { return a.rank != b.rank ? a.rank < b.rank : a.ldof < b.ldof; });
  m->n_ghost = (int32_t)gh.size();
  std::vector<int32_t> lidx(m->nV);
  OHP_CUDA(cudaMemcpyAsync(lidx.data(), m->lidx, sizeof(int32_t) * m->nV, cudaMemcpyDeviceToHost, st));
  OHP_CUDA(cudaStreamSynchronize(st));
  std::vector<int64_t> ghost_gid(gh.size());
  std::vector<std::vector<int32_t>> req(P), serve;
  for (size_t k = 0; k < gh.size(); ++k) {
    lidx[gh[k].vertex] = m->n_owned + (int32_t)k;
    ghost_gid[k] = start[gh[k].rank] + gh[k].ldof;
    req[gh[k].rank].push_back(gh[k].ldof);
  }
  OHP_CUDA(cudaMemcpyAsync(m->lidx, lidx.data(), sizeof(int32_t) * m->nV, cudaMemcpyHostToDevice, st));
  int64_t *d_ggid;
  OHP_CUDA(OHP_MALLOC(ctx, &d_ggid, sizeof(int64_t) * std::max<size_t>(1, gh.size())));
  OHP_CUDA(cudaMemcpyAsync(d_ggid, ghost_gid.data(), sizeof(int64_t) * gh.size(), cudaMemcpyHostToDevice, st));
  OHP_CUDA(OHP_MALLOC(ctx, &m->ldof_gid, sizeof(int64_t) * std::max(1, m->n_owned + m->n_ghost)));
  k_ldof_gid<<<nblk(m->n_owned + m->n_ghost, 256), 256, 0, st>>>(m->ldof_gid, m->n_owned, m->row_start, d_ggid, m->n_ghost);
  OHP_CHECK_LAUNCH(ctx);
  OHP_CUDA(cudaStreamSynchronize(st));
  OHP_FREE(ctx, d_ggid);
  /* who needs which of my dofs */
  OHP_TRY(exchange_lists(ctx, req, serve));
  std::vector<int32_t> send_ldof;
  h->d_soff.assign(1, 0); h->d_roff.assign(1, 0);
  for (int q = 0; q < P; ++q) {
    if (req[q].empty() && serve[q].empty()) continue;
    for (int32_t l : serve[q]) {
      if (l < 0 || l >= m->n_owned) OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "halo: rank %d asks for dof %d of rank %d", q, l, ctx->rank);
      send_ldof.push_back(l);
    }
    h->d_nbr.push_back(q);
    h->d_soff.push_back((int64_t)send_ldof.size());
    h->d_roff.push_back(h->d_roff.back() + (int64_t)req[q].size());
  }
  h->d_ns = (int64_t)send_ldof.size(); h->d_nr = m->n_ghost;
  OHP_CUDA(OHP_MALLOC(ctx, &h->d_send_ldof, sizeof(int32_t) * std::max<int64_t>(1, h->d_ns)));
  OHP_CUDA(OHP_MALLOC(ctx, &h->d_sbuf, sizeof(double) * std::max<int64_t>(1, h->d_ns)));
  OHP_CUDA(cudaMemcpyAsync(h->d_send_ldof, send_ldof.data(), sizeof(int32_t) * h->d_ns, cudaMemcpyHostToDevice, st));
  OHP_CUDA(cudaStreamSynchronize(st));
  /* tell every rank that serves me where its block starts in my vector (used by the P2P push) */
  std::vector<std::vector<int32_t>> tell(P), told;
  for (size_t k = 0; k < h->d_nbr.size(); ++k)
    if (h->d_roff[k + 1] > h->d_roff[k]) tell[h->d_nbr[k]].push_back(m->n_owned + (int32_t)h->d_roff[k]);
  OHP_TRY(exchange_lists(ctx, tell, told));
  h->d_dst_off.assign(h->d_nbr.size(), -1);
  for (size_t k = 0; k < h->d_nbr.size(); ++k)
    if (!told[h->d_nbr[k]].empty()) h->d_dst_off[k] = told[h->d_nbr[k]][0];
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* NVLink peer-to-peer mapping (CUDA IPC between the per-GPU processes)                  */
/* ------------------------------------------------------------------------------------ */
/* Collective when the heap was attached: an exporter must not free memory a peer still maps (CUDA IPC),
 * so every rank first closes its mappings, then all ranks meet (a 1-element all-reduce as the barrier), then
 * each frees its own heap.  ohp_mesh_destroy therefore has to be called by all ranks, like DMDestroy on a
 * parallel DM. */
static void p2p_release(ohp_context ctx, ohp_p2p *p) {
  for (int q = 0; q < kMaxRanks; ++q)
    if (p->peer[q] && p->peer[q] != p->heap) cudaIpcCloseMemHandle(p->peer[q]);
  if (p->attached && ctx->comm && ctx->nranks > 1) {
    cudaMemsetAsync(ctx->d_scalars + 8, 0, sizeof(double), ctx->stream);
    g_nccl.AllReduce(ctx->d_scalars + 8, ctx->d_scalars + 8, 1, ncclFloat64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(p->heap);
  delete p;
}
/* An attached heap is not released with its mesh but kept in the context for the next mesh (ex12 builds one DM per run, a
 * time-stepping or adaptive caller rebuilds it every step): cudaMalloc + cudaIpcOpenMemHandle for 7 peers + cudaFree cost
 * ~25 ms per mesh at 8 GPUs (measured with OHP_BENCH_E2E_BREAKDOWN), 8 % of a 0.29 s Newton step.  All ranks keep / reuse
 * / drop the cache in lockstep: the decisions depend only on quantities that are equal on every rank. */
void ohp_p2p_free(ohp_context ctx, ohp_p2p *p) {
  if (!p) return;
  static const char *nocache = getenv("OHP_P2P_NOCACHE");
  if (p->attached && ctx->comm && ctx->nranks > 1 && !nocache) {
    if (ctx->p2p_cache) p2p_release(ctx, ctx->p2p_cache); /* keep the most recent one */
    ctx->p2p_cache = p;
    return;
  }
  p2p_release(ctx, p);
}
void ohp_p2p_cache_drop(ohp_context ctx) {
  if (ctx->p2p_cache) p2p_release(ctx, ctx->p2p_cache);
  ctx->p2p_cache = nullptr;
}

extern "C" int ohp_mesh_p2p_export(ohp_mesh m, void *handle64) {
  if (!m || !m->setup_done || !handle64) OHP_FAIL(OHP_ERR_ORDER, "ohp_mesh_p2p_export: mesh not set up");
  ohp_context ctx = m->ctx;
  if (ctx->nranks < 2) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_mesh_p2p_export: no communicator attached");
  if (ctx->nranks > kMaxRanks) OHP_FAIL(OHP_ERR_SUP, "ohp_mesh_p2p_export: more than %d ranks", kMaxRanks);
  if (m->p2p) OHP_FAIL(OHP_ERR_ORDER, "ohp_mesh_p2p_export: already exported");
  OHP_CUDA(cudaSetDevice(ctx->device));
  std::vector<int64_t> lens;
  OHP_TRY(ohp_allgather_i64(ctx, (int64_t)m->n_owned + m->n_ghost, lens));
  const int64_t need_cap = (*std::max_element(lens.begin(), lens.end()) + 63) / 64 * 64 + 64;
  if (ctx->p2p_cache && ctx->p2p_cache->pcap >= need_cap) {
    /* reuse the cached heap and the peers' mappings (every rank sees the same need_cap and the same cached capacity, so
     * all of them reuse together).  No peer writes here any more: its last solve on the old mesh ended with `done`, after
     * which its kernels return without pushing.  The control words are cleared in stream order, before the next solve's
     * entry barrier lets anybody push. */
    ohp_p2p *p = ctx->p2p_cache;
    ctx->p2p_cache = nullptr;
    p->reused = true;
    p->attached = false;
    OHP_CUDA(cudaMemsetAsync(p->heap, 0, p->off_p0, ctx->stream));
    OHP_CUDA(cudaMemsetAsync((char *)p->heap + p->bytes - 256, 0, 256, ctx->stream));
    memcpy(handle64, p->handle, 64);
    m->p2p = p;
    return 0;
  }
  if (ctx->p2p_cache) ohp_p2p_cache_drop(ctx); /* too small for this mesh (on every rank alike) */
  ohp_p2p *p = new ohp_p2p();
  p->pcap = need_cap;
  p->off_slots = 0;
  p->off_hflags = 2 * kMaxRanks * sizeof(ohp_ar_slot);
  p->off_p0 = 4096;
  p->off_p1 = p->off_p0 + (size_t)p->pcap * sizeof(double);
  p->bytes = p->off_p1 + (size_t)p->pcap * sizeof(double) + 256; /* + local sequence counters at the end */
  OHP_CUDA(cudaMalloc(&p->heap, p->bytes));
  OHP_CUDA(cudaMemset(p->heap, 0, p->bytes));
  cudaIpcMemHandle_t hnd;
  OHP_CUDA(cudaIpcGetMemHandle(&hnd, p->heap));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  memcpy(handle64, &hnd, 64);
  memcpy(p->handle, &hnd, 64);
  m->p2p = p;
  return 0;
}

extern "C" int ohp_mesh_p2p_attach(ohp_mesh m, const void *handles) {
  if (!m || !m->p2p || !handles) OHP_FAIL(OHP_ERR_ORDER, "ohp_mesh_p2p_attach: call ohp_mesh_p2p_export first");
  ohp_context ctx = m->ctx;
  ohp_p2p *p = m->p2p;
  ohp_halo *h = m->halo;
  OHP_CUDA(cudaSetDevice(ctx->device));
  for (int q = 0; q < ctx->nranks; ++q) {
    if (q == ctx->rank) { p->peer[q] = p->heap; continue; }
    if (p->reused && p->peer[q]) continue; /* mapping kept from the previous mesh */
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, (const char *)handles + 64 * q, 64);
    OHP_CUDA(cudaIpcOpenMemHandle(&p->peer[q], hnd, cudaIpcMemLazyEnablePeerAccess));
  }
  ohp_p2p_dev &d = p->dev;
  memset(&d, 0, sizeof(d));
  d.rank = ctx->rank; d.nranks = ctx->nranks;
  if ((int)h->d_nbr.size() > kMaxRanks) OHP_FAIL(OHP_ERR_SUP, "p2p: too many halo neighbours");
  d.n_nbr = 0;
  for (size_t k = 0; k < h->d_nbr.size(); ++k) {
    /* every neighbour gets a (possibly empty) push so that flags advance in lockstep */
    d.nbr[d.n_nbr] = h->d_nbr[k];
    d.soff[d.n_nbr] = (int)h->d_soff[k];
    d.soff[d.n_nbr + 1] = (int)h->d_soff[k + 1];
    d.dst_off[d.n_nbr] = h->d_dst_off[k];
    if (h->d_soff[k + 1] > h->d_soff[k] && h->d_dst_off[k] < 0) OHP_FAIL(OHP_ERR_LIB, "p2p: missing destination offset for rank %d", h->d_nbr[k]);
    d.n_nbr++;
  }
  d.send_ldof = h->d_send_ldof;
  for (int q = 0; q < ctx->nranks; ++q) {
    char *base = (char *)p->peer[q];
    d.slots[q] = (ohp_ar_slot *)(base + p->off_slots);
    d.hflags[q] = (unsigned long long *)(base + p->off_hflags);
    d.p[0][q] = (double *)(base + p->off_p0);
    d.p[1][q] = (double *)(base + p->off_p1);
  }
  d.seq = (unsigned long long *)((char *)p->heap + p->bytes - 256);
  {
    const char *t = getenv("OHP_P2P_TIMEOUT_S");
    double sec = t ? atof(t) : 60.0;
    if (!(sec > 0.0)) sec = 60.0;
    d.timeout_ns = (unsigned long long)(sec * 1e9);
  }
  p->attached = true;
  return 0;
}

int ohp_p2p_entry_barrier(ohp_mesh m) {
  ohp_context ctx = m->ctx;
  if (ctx->nranks < 2 || !ctx->comm) return 0;
  OHP_CUDA(cudaMemsetAsync(ctx->d_scalars + 10, 0, sizeof(double), ctx->stream));
  OHP_NCCL(g_nccl.AllReduce(ctx->d_scalars + 10, ctx->d_scalars + 10, 1, ncclFloat64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
  return 0;
}

int ohp_p2p_check(ohp_mesh m, const char *what) {
  if (!m->p2p || !m->p2p->attached) return 0;
  ohp_context ctx = m->ctx;
  unsigned long long err = 0;
  unsigned long long *flag = m->p2p->dev.seq + 2;
  OHP_CUDA(cudaMemcpyAsync(&err, flag, sizeof(err), cudaMemcpyDeviceToHost, ctx->stream));
  OHP_CUDA(cudaStreamSynchronize(ctx->stream));
  if (err) {
    cudaMemsetAsync(flag, 0, sizeof(err), ctx->stream); /* reported once; the next solve starts clean */
    OHP_FAIL(OHP_ERR_LIB, "%s: a peer-to-peer wait timed out after %.0f s (a neighbouring rank stopped responding; OHP_P2P_TIMEOUT_S raises the limit)",
             what, m->p2p->dev.timeout_ns * 1e-9);
  }
  return 0;
}

/* export + all-gather of the handles over the communicator + attach, in one collective call (what a C/C++ host
 * uses; the two-step form exists for launchers that carry the handles themselves) */
extern "C" int ohp_mesh_p2p_enable(ohp_mesh m) {
  if (!m || !m->setup_done) OHP_FAIL(OHP_ERR_ORDER, "ohp_mesh_p2p_enable: mesh not set up");
  ohp_context ctx = m->ctx;
  if (ctx->nranks < 2) return 0;
  if (m->p2p && m->p2p->attached) return 0;
  unsigned char mine[64];
  OHP_TRY(ohp_mesh_p2p_export(m, mine));
  const int P = ctx->nranks;
  unsigned char *d;
  OHP_CUDA(OHP_MALLOC(ctx, &d, 64 * (size_t)(P + 1)));
  OHP_CUDA(cudaMemcpyAsync(d + 64 * (size_t)P, mine, 64, cudaMemcpyHostToDevice, ctx->stream));
  ncclResult_t r = g_nccl.AllGather(d + 64 * (size_t)P, d, 64, ncclChar, (ncclComm_t)ctx->comm, ctx->stream);
  std::vector<unsigned char> all(64 * (size_t)P);
  if (r == ncclSuccess) {
    cudaMemcpyAsync(all.data(), d, all.size(), cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
  }
  OHP_FREE(ctx, d);
  OHP_NCCL(r);
  return ohp_mesh_p2p_attach(m, all.data());
}

/* DMGlobalToLocal / MatMult scatter: x[n_owned + k] <- owner's value of ghost dof k */
int ohp_halo_exchange(ohp_mesh m, double *x) {
  ohp_context ctx = m->ctx;
  if (ctx->nranks == 1) return 0;
  ohp_halo *h = m->halo;
  if (!h) OHP_FAIL(OHP_ERR_ORDER, "ohp_halo_exchange: no halo plan");
  cudaStream_t st = ctx->stream;
  if (h->d_ns) { k_gather_f64<<<nblk(h->d_ns, 256), 256, 0, st>>>(x, h->d_send_ldof, h->d_ns, h->d_sbuf); OHP_CHECK_LAUNCH(ctx); }
  if (h->d_nbr.empty()) return 0;
  NcclGroup grp;
  OHP_NCCL(grp.start());
  for (size_t k = 0; k < h->d_nbr.size(); ++k) {
    const int64_t ns = h->d_soff[k + 1] - h->d_soff[k], nr = h->d_roff[k + 1] - h->d_roff[k];
    if (ns) OHP_NCCL(g_nccl.Send(h->d_sbuf + h->d_soff[k], (size_t)ns, ncclFloat64, h->d_nbr[k], (ncclComm_t)ctx->comm, st));
    if (nr) OHP_NCCL(g_nccl.Recv(x + m->n_owned + h->d_roff[k], (size_t)nr, ncclFloat64, h->d_nbr[k], (ncclComm_t)ctx->comm, st));
  }
  OHP_NCCL(grp.end());
  return 0;
}
