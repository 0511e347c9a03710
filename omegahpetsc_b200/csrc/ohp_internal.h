/* Internal declarations of libohp.so (not part of the C-ABI). */
#ifndef OHP_INTERNAL_H
#define OHP_INTERNAL_H

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

#include "ohp.h"

void ohp_set_error(const char *fmt, ...);

#define OHP_FAIL(code, ...)      \
  do {                           \
    ohp_set_error(__VA_ARGS__);  \
    return (code);               \
  } while (0)

#define OHP_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      OHP_FAIL(OHP_ERR_LIB, "%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,             \
               cudaGetErrorString(e__));                                                      \
  } while (0)

/* Device memory comes from the device's stream-ordered pool (cudaMallocAsync on the context stream, release threshold
 * raised to "never" in ohp_init): after warm-up neither an allocation nor a free enters the driver's resource manager,
 * which serialises with every NVML / nvidia-smi query on the box (measured in round 2: with a clock sampler running,
 * single cudaMalloc/cudaFree calls inside a Newton step stalled it by 0.2-1.5 s).  The exchange heap of the NVLink peer
 * path stays cudaMalloc'ed (CUDA IPC cannot export pool memory). */
#define OHP_MALLOC(ctx_, pptr, bytes) cudaMallocAsync((void **)(pptr), (size_t)(bytes), (ctx_)->stream)
#define OHP_FREE(ctx_, ptr) do { if (ptr) cudaFreeAsync((void *)(ptr), (ctx_)->stream); } while (0)

#define OHP_TRY(expr)                 \
  do {                                \
    int rc__ = (expr);                \
    if (rc__) return rc__;            \
  } while (0)

#define OHP_CHECK_LAUNCH(ctx)                                                                 \
  do {                                                                                        \
    (ctx)->launches++;                                                                        \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess)                                                                   \
      OHP_FAIL(OHP_ERR_LIB, "%s:%d kernel launch failed: %s", __FILE__, __LINE__,             \
               cudaGetErrorString(e__));                                                      \
  } while (0)

/* device-side scalar block of the CG recurrence (lives in HBM, polled by the host) */
struct ohp_cg_state {
  double beta, betaold, dpi, alpha, dp, rnorm0, ttol;
  double rtol, abstol, dtol;
  int its, reason, maxit, done;
  int x_pending, pad; /* 1: x += alpha p of the last completed iteration has not been applied yet */
};

struct ohp_halo;  /* ohp_comm.cu */

/* tables of the cell-parallel assembly (ohp_patch.cu) */
struct ohp_patch_plan {
  int R, n_patches, max_cells, max_verts;
  int64_t total_star, total_cells, total_verts;
  int32_t *prow, *pstar_ptr, *pcell_ptr, *pvert_ptr, *pvert;
  uint16_t *pstar, *pconn;
  uint8_t *pslot;
};

/* device-side timeline of the CG loop (OHP_TRACE_CG, see ohp_dev.cuh) */
struct ohp_trace_rec { unsigned long long t; int ev, it, blk, pad; };
struct ohp_trace { ohp_trace_rec *buf; unsigned *count; unsigned cap; int it0, it1; };

/* ---- peer-to-peer (NVLink) view of the communicator, passed BY VALUE to the CG kernels ---- */
constexpr int kMaxRanks = 8;
struct ohp_ar_slot { double v[3]; unsigned long long seq; }; /* 32 B: payload, then the sequence number (release-stored last) */
struct ohp_p2p_dev {
  int rank, nranks;
  int n_nbr;                         /* halo neighbours */
  int nbr[kMaxRanks];
  int soff[kMaxRanks + 1];           /* send-list offsets per neighbour */
  long long dst_off[kMaxRanks];      /* where my values land in the neighbour's vector (its ghost tail) */
  const int32_t *send_ldof;          /* my dofs to push, grouped by neighbour */
  ohp_ar_slot *slots[kMaxRanks];     /* slots[q]: rank q's all-reduce mailbox (2 parities x kMaxRanks sources) */
  unsigned long long *hflags[kMaxRanks]; /* hflags[q][src]: sequence number of src's last halo push into q */
  double *p[2][kMaxRanks];           /* p[b][q]: rank q's search-direction buffer b (owned dofs + ghost tail) */
  unsigned long long *seq;           /* local counters: [0] all-reduces issued, [1] halo pushes issued, [2] wait-timeout flag */
  unsigned long long timeout_ns;     /* how long a peer wait may spin before it raises seq[2] */
};
struct ohp_p2p {
  void *heap = nullptr;              /* cudaMalloc'ed, exported with cudaIpcGetMemHandle */
  size_t bytes = 0, off_slots = 0, off_hflags = 0, off_p0 = 0, off_p1 = 0;
  int64_t pcap = 0;                  /* capacity (doubles) of each p buffer: max over ranks */
  void *peer[kMaxRanks] = {nullptr};
  unsigned char handle[64] = {0};    /* this rank's CUDA IPC handle of `heap` */
  bool attached = false;
  bool reused = false;               /* taken from the context's cache: the peers' mappings are already open */
  ohp_p2p_dev dev;
};

struct ohp_context_s {
  int device;
  int num_sms;
  cudaStream_t stream;
  int64_t launches;
  /* reduction scratch */
  double *d_partials;   /* kMaxPartials * 2 */
  unsigned *d_ticket;   /* last-block ticket counters */
  double *d_scalars;    /* results of reductions */
  double *h_scalars;    /* pinned */
  ohp_cg_state *d_cg;
  ohp_cg_state *h_cg;   /* pinned */
  int *d_flag;          /* device error flag for setup kernels */
  ohp_trace trace;      /* buf == NULL unless OHP_TRACE_CG is set */
  struct ohp_p2p *p2p_cache; /* exchange heap (with the peers' IPC mappings) of the last mesh destroyed, kept for the next one */
  /* communicator (NULL = serial) */
  void *nccl_lib;
  void *comm;
  int rank, nranks;
};

constexpr int kMaxPartials = 1 << 16;
constexpr int kSpmvTile = 4096;     /* nnz per SpMV tile (both kernels) */
constexpr int kSpmvThreads = 256;   /* block size of the fallback CSR-stream kernel */
constexpr int kSpmvRowCap = 2048;   /* rows starting in one tile that the TMA-pipelined kernel stages */
/* vals / colidx are allocated to a whole number of tiles (zero-filled tail) so that every tile is
 * one fixed-size, 16-byte aligned bulk copy */
static inline int64_t ohp_padded_nnz(int64_t nnz) { return ((nnz + kSpmvTile - 1) / kSpmvTile) * kSpmvTile + kSpmvTile; }

struct ohp_mesh_s {
  ohp_context ctx;
  int dim, nc;
  int32_t nC, nV;
  int32_t *cells;   /* device */
  double *coords;   /* device */
  /* point SF (host copies) */
  std::vector<int32_t> ghost_vertex, ghost_owner_rank, ghost_owner_vertex;
  uint8_t *owned;   /* device, per vertex */
  bool setup_done;
  /* topology */
  int32_t *v2c_ptr, *v2c;
  uint8_t *bnd;
  /* section */
  int32_t *lidx;        /* per vertex local dof (-1 constrained) */
  int32_t *row_vertex;  /* per owned row */
  int64_t *ldof_gid;    /* per local dof (owned + ghost): global id */
  int32_t n_owned, n_ghost, n_bnd;
  int64_t n_global, row_start;
  /* CSR */
  int32_t *rowptr, *colidx;
  int16_t *col16;       /* column - row per entry, or NULL when some offset does not fit 16 bits */
  uint8_t *slot;
  int64_t nnz;
  int32_t max_row;
  /* SpMV tiling */
  int32_t *tile_row; /* n_tiles+1: first row of each nnz tile */
  int32_t n_tiles;
  int32_t max_tile_rows; /* most rows starting inside one tile */
  bool all_tma;          /* every rank's matrix fits the TMA-pipelined SpMV (max_tile_rows below its row cap) */
  ohp_halo *halo;    /* dof-level halo plan (NULL when serial) */
  ohp_p2p *p2p;      /* NVLink peer mapping (NULL until ohp_mesh_p2p_attach) */
  /* facets (DMPlexInterpolate's edges / faces), built on demand for Plex queries */
  bool facets_built;
  int32_t n_facets, *facet_verts, *facet_support, *cell_facets;
  /* "marker" label set point by point by the caller (DMSetLabelValue) instead of detected here */
  bool has_user_label;
  std::vector<int32_t> user_label;
  int bc_type;                  /* OHP_BC_ESSENTIAL (0, default) | OHP_BC_NATURAL | OHP_BC_NONE: the last two constrain no dof */
  ohp_patch_plan *patch;        /* NULL: the row kernels are used (limits exceeded, OHP_ASM=rows, or not built yet) */
  struct ohp_pf_plan *pfplan;   /* tables of the one-launch pipelined CG (ohp_linalg.cu), built on first use */
  struct ohp_fused_plan *fused; /* tables of the fused persistent CG kernel, built on first use (ohp_cg_fused.cu) */
  int32_t *cta_tile0;  /* weighted tile partition of the SpMV CTAs (G + 1 entries; NULL when serial), with its own ghost flags: */
  uint8_t *ghost_cta_w;
  uint8_t *ghost_tile; /* per SpMV tile: 1 if it references ghost columns (NULL when serial) */
  uint8_t *ghost_cta; /* per persistent-SpMV CTA: 1 if its tiles reference ghost columns (NULL when serial) */
};

struct ohp_vec_s {
  ohp_mesh mesh;
  double *d;    /* n_owned + n_ghost */
  int32_t n;    /* n_owned */
};

struct ohp_mat_s {
  ohp_mesh mesh;
  double *vals;
  /* CG work vectors, allocated on first solve; the search direction is double-buffered
   * (p[its & 1] is written from p[(its & 1) ^ 1]) so that a halo push never races its own update */
  double *r, *z, *p, *p1, *w, *dinv, *s;
  double *r1, *z1, *s1; /* second buffers of the single-reduction recurrence (r, u, s are double-buffered there) */
  double *pm0, *pm1, *pw0, *pw1; /* pipelined recurrence: SpMV inputs m = M^-1 w (serial runs) and w, double-buffered */
  double *hist;
  int32_t hist_cap;
  int32_t nullspace;            /* 1: the constants are in the kernel (ohp_mat_set_nullspace) */
};

/* ohp_core.cu */
int ohp_scan_exclusive_i32(ohp_context ctx, const int32_t *in, int32_t *out, int64_t n, int64_t *total);
int ohp_assembly_view_make(ohp_mesh mesh, const double *u, double *out, ohp_assembly_view *vw);

/* ohp_patch.cu */
int ohp_patch_plan_build(ohp_mesh mesh);
void ohp_patch_plan_free(ohp_context ctx, ohp_patch_plan *pl);

/* ohp_comm.cu */
int ohp_halo_build(ohp_mesh mesh);
void ohp_halo_free(ohp_context ctx, ohp_halo *h);
/* fill the ghost tail x[n_owned ..] from the owners (on the context stream) */
int ohp_halo_exchange(ohp_mesh mesh, double *x);
int ohp_allreduce_sum(ohp_context ctx, double *d_vals, int n);
int ohp_vertex_halo_i32(ohp_mesh mesh, int32_t *vertex_data_dev);
int ohp_allgather_i64(ohp_context ctx, int64_t mine, std::vector<int64_t> &all);
int ohp_sparse_exchange(ohp_context ctx, int elem_bytes, const void *d_send, const std::vector<int64_t> &scount,
                        void **d_recv, std::vector<int64_t> &rcount);
void ohp_comm_destroy(ohp_context ctx);
void ohp_p2p_free(ohp_context ctx, ohp_p2p *p);
/* really release the cached exchange heap (collective; context teardown) */
void ohp_p2p_cache_drop(ohp_context ctx);
/* solve entry in the peer-to-peer modes: a stream-ordered barrier over the communicator, so that the first fused
 * all-reduce / halo wait of a solve never measures start-up skew (lazy module load, allocation, a profiler) */
int ohp_p2p_entry_barrier(ohp_mesh mesh);
/* after a solve: did a peer wait time out?  Clears the flag once it has been reported. */
int ohp_p2p_check(ohp_mesh mesh, const char *what);

/* ohp_linalg.cu */
ohp_trace ohp_cg_trace_open(ohp_context ctx);
void ohp_cg_trace_dump(ohp_context ctx, const char *algo);
int ohp_spmv_launch(ohp_mat A, const double *x, double *y);
int ohp_extract_dinv(ohp_mat A);
int ohp_spmv_cheb_launch(ohp_mat A, const double *yk, const double *ykm1, const double *r, const double *dinv, double *ynew,
                         double a, double b, double c, int *fused);

void ohp_pf_plan_free(ohp_context ctx, struct ohp_pf_plan *pl);

/* ohp_cg_fused.cu */
int ohp_cg_fused_solve(ohp_mat A, ohp_vec b, ohp_vec x, const ohp_ksp_options *o, ohp_ksp_result *res, bool p2p);
void ohp_fused_plan_free(ohp_context ctx, struct ohp_fused_plan *p);

/* ohp_cg_persistent.cu */
int ohp_cg_persistent_solve(ohp_mat A, ohp_vec b, ohp_vec x, const ohp_ksp_options *o, ohp_ksp_result *res, bool p2p);

/* ohp_pcg.cu */
int ohp_pcg_solve(ohp_mat A, ohp_vec b, ohp_vec x, const ohp_ksp_options *o, ohp_ksp_result *res);

/* ohp_problems.cu */
int ohp_problem_lookup(int id, ohp_assembly_launcher *res, ohp_assembly_launcher *jac,
                       ohp_assembly_launcher *proj, ohp_assembly_launcher *l2);
int ohp_problem_lookup_bd(int id, ohp_assembly_launcher *bd);

#endif
