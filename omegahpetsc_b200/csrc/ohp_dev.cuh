/* Device-side building blocks shared by the Vec/Mat/CG translation units of libohp.so:
 * deterministic two-stage reductions, NVLink peer-to-peer flag / mailbox primitives, mbarrier +
 * bulk-async-copy (TMA) wrappers and the row dot product of the SpMV kernels. */
#ifndef OHP_DEV_CUH
#define OHP_DEV_CUH

#include "ohp_internal.h"

/* Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization
 * attribute may start while its predecessor drains; pdl_wait() blocks until the predecessor has
 * completed and its writes are visible, pdl_launch() lets the successor's CTAs be scheduled.
 * Both are no-ops for a kernel launched without the attribute. */
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t ohp_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

/* ------------------------------------------------------------------------------------ */
/* block reduction + last-block ticket                                                   */
/* ------------------------------------------------------------------------------------ */
template <int NV> __device__ __forceinline__ void block_sum(double (&v)[NV], double *s_red /* NV*32 */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) s_red[k * 32 + wid] = v[k];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double t = lane < nw ? s_red[k * 32 + lane] : 0.0;
#pragma unroll
      for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      v[k] = t;
    }
  }
  __syncthreads();
}

/* returns true in every thread of the last block to arrive, with the grid totals in v
 * (valid in thread 0).  Summation order is fixed => bitwise reproducible. */
template <int NV>
__device__ __forceinline__ bool grid_sum(double (&v)[NV], double *partials, unsigned *ticket) {
  __shared__ double s_red[NV * 32];
  __shared__ bool s_last;
  block_sum<NV>(v, s_red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) partials[(size_t)blockIdx.x * NV + k] = v[k];
    __threadfence();
    const unsigned t = atomicInc(ticket, gridDim.x - 1); /* wraps back to 0: self-resetting */
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = 0.0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] += __ldcg(partials + (size_t)b * NV + k);
  }
  block_sum<NV>(v, s_red);
  return true;
}

/* ------------------------------------------------------------------------------------ */
/* NVLink peer-to-peer all-reduce, fused into the kernel that produces the partial sums.   */
/* Called by every thread of the LAST block (grid_sum returned true); warp 0 does the work: */
/* lane q stores this rank's totals into rank q's mailbox (release store of the sequence    */
/* number after the payload), then spins (acquire) on the mailbox rank q fills here; the    */
/* totals are added in rank order, so every rank gets bitwise the same sum.  Mailboxes are  */
/* double-buffered by sequence parity: a rank can be at most one all-reduce ahead.          */
/* ------------------------------------------------------------------------------------ */
enum { OHP_MODE_SERIAL = 0, OHP_MODE_NCCL = 1, OHP_MODE_P2P = 2 };

/* Device-side timeline of the CG loop (OHP_TRACE_CG=<file>): selected threads append (event, iteration, block,
 * %globaltimer) records to a ring in HBM; the host dumps them after the solve.  This is how the shares of the
 * iteration that are not data movement (launch ramp, halo wait, all-reduce, rank skew) are measured on the
 * multi-GPU runs, where a profiler cannot be attached.  A null buffer costs one predicated branch. */
enum { TR_VEC_BEGIN = 1, TR_VEC_HALO_PUSHED = 2, TR_VEC_END = 3, TR_SPMV_BEGIN = 4, TR_SPMV_HALO_READY = 5, TR_SPMV_CTA_DONE = 6,
       TR_SPMV_LAST_BLOCK = 7, TR_AR_POSTED = 8, TR_AR_DONE = 9, TR_SPMV_END = 10, TR_UPD_BEGIN = 11, TR_UPD_END = 12 };
__device__ __forceinline__ unsigned long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_put(const ohp_trace &tr, int ev, int it) {
  if (!tr.buf || it < tr.it0 || it >= tr.it1) return;
  const unsigned k = atomicAdd(tr.count, 1u);
  if (k < tr.cap) { ohp_trace_rec r; r.t = trace_now(); r.ev = ev; r.it = it; r.blk = (int)blockIdx.x; r.pad = 0; tr.buf[k] = r; }
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
/* spin until *flag >= seq.  A peer that died must not hang this GPU: after `timeout_ns` (ohp_p2p_dev::timeout_ns,
 * OHP_P2P_TIMEOUT_S, default 60 s -- far above any rank skew a healthy run can show: the ranks enter every solve
 * through a stream-ordered barrier) the wait gives up and raises seq[2]; once it is raised every later wait
 * returns at once, so the chunk in flight drains quickly and the host (which reads seq[2] after the chunk) fails
 * the solve with an error, then clears the flag, instead of hanging the box. */
__device__ __forceinline__ void p2p_wait(const unsigned long long *flag, unsigned long long seq, unsigned long long *err,
                                         unsigned long long timeout_ns) {
  if (ld_acquire_sys(flag) >= seq) return;
  if (*(volatile unsigned long long *)err) return;
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(flag) < seq) {
    if (global_ns() - t0 > timeout_ns) { *(volatile unsigned long long *)err = 1ull; return; }
  }
}

/* one full warp; v valid in lane 0 on entry, the totals valid in every lane on return */
template <int NV> __device__ __forceinline__ void p2p_allreduce_warp(const ohp_p2p_dev &pd, double (&v)[NV], int lane) {
  unsigned long long seq = 0;
  if (lane == 0) { seq = pd.seq[0] + 1; pd.seq[0] = seq; }
  seq = __shfl_sync(0xffffffffu, seq, 0);
  double mine[NV], got[NV], tot[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) { mine[k] = __shfl_sync(0xffffffffu, v[k], 0); got[k] = mine[k]; tot[k] = 0.0; }
  const int par = (int)(seq & 1ull);
  const bool peer = lane < pd.nranks && lane != pd.rank;
  if (peer) {
    ohp_ar_slot *s = pd.slots[lane] + par * kMaxRanks + pd.rank;
#pragma unroll
    for (int k = 0; k < NV; ++k) s->v[k] = mine[k];
    st_release_sys(&s->seq, seq);
    const ohp_ar_slot *r = pd.slots[pd.rank] + par * kMaxRanks + lane;
    p2p_wait(&r->seq, seq, pd.seq + 2, pd.timeout_ns);
#pragma unroll
    for (int k = 0; k < NV; ++k) got[k] = r->v[k];
  }
  for (int q = 0; q < pd.nranks; ++q) {
#pragma unroll
    for (int k = 0; k < NV; ++k) tot[k] += __shfl_sync(0xffffffffu, got[k], q);
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = tot[k];
}
template <int NV> __device__ __forceinline__ void p2p_allreduce(const ohp_p2p_dev &pd, double (&v)[NV]) {
  if (threadIdx.x >= 32) return;
  p2p_allreduce_warp<NV>(pd, v, (int)threadIdx.x);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
/* L2 residency control for the matrix stream.  The matrix never changes during a solve and is re-read every
 * iteration, so the tiles that fit a share of the 126 MB L2 are loaded with an evict_last policy (they survive
 * the vector traffic between two SpMVs) and the rest with evict_first (they stream through without displacing
 * the resident ones).  Which tiles are resident is a fixed function of the tile index, so the same lines hit
 * every iteration. */
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_1d_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
/* plain (coherent, L1-cacheable) load: for vectors that peers or this kernel write while it runs, where the
 * read-only path (ld.global.nc, __ldg) is undefined */
__device__ __forceinline__ double ld_weak(const double *p) {
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

/* sum_{k in [a,b)} vs[k] * x[cs[k]], ascending k; loads batched 8 at a time for memory-level parallelism */
/* IDX = int32_t: absolute local column ids.  IDX = int16_t: column = row + delta (the compressed
 * index stream used when every |column - row| fits 16 bits, which mesh-ordered matrices do). */
/* (An escape code for the few offsets that do not fit 16 bits -- ~1 % of the entries of a Morton-ordered box -- was
 * measured in round 2: the data-dependent second load in the gather loop cost 33 % of the SpMV (0.192 vs 0.144 ms on
 * the 16 M-triangle box), far more than the 6 % the 32-bit stream costs, so such matrices use the 32-bit stream.) */
template <class IDX, bool COHERENT = false>
__device__ __forceinline__ double row_dot(const double *__restrict__ vs, const IDX *__restrict__ cs,
                                          int a, int b, const double *x, double s, int row) {
  for (int k = a; k < b; k += 8) {
    double xv[8], vv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool ok = (k + j) < b;
      const int c = ok ? (sizeof(IDX) == 2 ? row + (int)cs[k + j] : (int)cs[k + j]) : 0;
      vv[j] = ok ? vs[k + j] : 0.0;
      xv[j] = ok ? (COHERENT ? ld_weak(x + c) : __ldg(x + c)) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if ((k + j) < b) s += vv[j] * xv[j];
  }
  return s;
}

constexpr int kTmaRpStride = kSpmvRowCap + 4;           /* rowptr slice + (first row, end row) of the tile */
constexpr int tma_smem_bytes(int stages, int idx_bytes = 4) { return stages * (kSpmvTile * (8 + idx_bytes) + kTmaRpStride * 4) + 2 * stages * 8; }

/* scalar recurrences of the single-reduction (Chronopoulos-Gear) CG: g = (r,u), d = (A u,u), nn = (u,u).
 * st->x_pending == 0 marks the first pass (r0, u0, w0 just computed); betaold carries gamma. */
__device__ __forceinline__ void cgcg_scalars(ohp_cg_state *st, double *hist, int hist_cap, double g, double d, double nn) {
  const double dp = sqrt(nn);
  st->dp = dp;
  int reason = 0;
  if (st->x_pending == 0) { /* first pass: r0, u0, w0 are known (x_pending doubles as the phase flag here) */
    st->x_pending = 1;
    st->its = 0;
    st->rnorm0 = dp;
    st->ttol = fmax(st->rtol * dp, st->abstol);
    if (hist) hist[0] = dp;
    if (dp != dp) reason = -9;
    else if (dp <= st->ttol) reason = (dp < st->abstol) ? 3 : 2;
    if (!reason && st->maxit <= 0) reason = -3;
    if (!reason && g == 0.0) reason = 3;
    if (!reason) {
      if (!(d > 0.0)) reason = (d != d) ? -9 : -8;
      else { st->alpha = g / d; st->beta = 0.0; st->betaold = g; }
    }
  } else {
    st->its += 1;
    if (hist && st->its < hist_cap) hist[st->its] = dp;
    if (dp != dp) reason = -9;
    else if (dp <= st->ttol) reason = (dp < st->abstol) ? 3 : 2;
    else if (dp >= st->dtol * st->rnorm0) reason = -4;
    if (!reason && st->its >= st->maxit) reason = -3;
    if (!reason && g == 0.0) reason = 3;
    if (!reason) {
      const double beta = g / st->betaold; /* betaold holds gamma of the previous iteration */
      const double denom = d - beta * g / st->alpha;
      if (!(denom > 0.0)) reason = (denom != denom) ? -9 : -8;
      else { st->beta = beta; st->alpha = g / denom; st->betaold = g; }
    }
  }
  if (reason) { st->reason = reason; st->done = 1; }
}


/* extra inputs of the SpMV epilogue when it closes a single-reduction CG iteration: the partial sums
 * (r,u), (u,u) left by the vector kernel's blocks, summed here in a fixed order */
struct ohp_cgcg_epi {
  const double *vpart; /* nvpart x 2, NULL = classic recurrence */
  int nvpart, jacobi, hist_cap;
  double *hist;
  const uint8_t *ghost_cta; /* non-NULL: the SpMV itself waits for the halo, and only in the CTAs flagged here */
  const int32_t *cta_tile0; /* non-NULL: CTA b owns tiles [cta_tile0[b], cta_tile0[b+1]) instead of the uniform split (the flags above belong to it) */
  int pipe;                 /* pipelined recurrence: no dot product here, the host names the input buffer.  1: an extra
                             * warp of CTA 0 reduces the vector kernel's partials (nvpart x 3), all-reduces them and
                             * advances the scalars WHILE the other warps stream the matrix; 2: nothing to reduce */
  int pin_tiles;            /* > 0: leading tiles of every CTA loaded evict_last, the rest evict_first; < 0: all evict_first; 0: no hints */
  ohp_trace trace;          /* optional device-side timeline (OHP_TRACE_CG) */
};

#endif
