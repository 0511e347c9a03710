/*
 * ohp_cg_fused.cu -- the CG iteration as ONE persistent cooperative kernel with ONE global synchronisation per
 * iteration.  This is the solver of the strong-scaling regime (a few hundred thousand to ~2 M rows per GPU),
 * where a device-side timeline of the two-launch loop (OHP_TRACE_CG, profiles/r02_trace_*.txt) showed that only
 * ~24 of 38 us per iteration are data movement: the rest is two kernel boundaries, the last-block reduction and the
 * ramp of every launch.
 *
 * Replaces [PETSc-internal] KSPSolve_CG's loop (selected by -ksp_type cg, ex12.cpp:1543) including its MatMult
 * scatter (halo) and the MPI_Allreduce of its dot products.  Recurrence: single-reduction CG of Chronopoulos & Gear
 * (u = M^-1 r, w = A u, gamma = (r,u), delta = (w,u); the same iterates as classic CG in exact arithmetic), the same
 * scalar code as the two-launch variant (cgcg_scalars), convergence test = KSPConvergedDefault on ||u||.
 *
 * Structure (one CTA per SM, 19 consumer warps + 1 TMA producer warp, all CTAs resident = cooperative launch):
 *   - CTA b owns the matrix tiles [t0,t1) AND the rows that start in them, for the vector update as well as for the
 *     SpMV, so x, p, s, w never leave "their" SM: only u (the SpMV input) is read across CTAs.
 *   - V phase: p = u + beta p; s = w + beta s; x += alpha p; r -= alpha s; u = M^-1 r on the CTA's rows, with the
 *     partial sums of (r,u), (u,u).  Rows that neighbouring RANKS ghost are stored straight into their vectors over
 *     NVLink by the CTA that owns them; the last contributor per neighbour publishes the halo flag (st.release.sys).
 *   - neighbour hand-off instead of a grid barrier: the CTA publishes "u of iteration j written" (release) and waits
 *     (acquire; this also invalidates its L1) only for the CTAs whose rows its tiles reference -- for a mesh-ordered
 *     matrix its two neighbours; CTAs that read ghost columns also wait for the peers' halo flags.
 *   - S phase: w = A u over the CTA's tiles through the same 3-stage TMA/mbarrier pipeline as k_spmv_tma; the
 *     producer warp never synchronises with anybody but its consumers, so it streams the next iteration's first
 *     tiles during the reduction and the vector update (the matrix is constant).
 *   - the ONE global synchronisation: every CTA writes its three partial sums and takes a ticket; the last CTA adds
 *     the partials in a fixed order and stores the rank's totals into EVERY rank's mailbox (its own included);
 *     every CTA of every rank then polls its LOCAL mailboxes, adds the ranks' totals in rank order and runs the
 *     scalar recurrences itself: bit-identical alpha/beta everywhere, no second "release the grid" hop, and the
 *     all-reduce across GPUs costs one NVLink store latency.
 * Everything is deterministic: fixed-order sums, no floating-point atomics.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "ohp_dev.cuh"
#include "ohp_internal.h"

namespace {

constexpr int FS = 3;      /* TMA stages */
constexpr int FNC = 608;   /* consumer threads (19 warps); + 1 producer warp */

struct FusedArgs {
  const int32_t *rowptr, *tile_row, *colidx32;
  const int16_t *col16;
  const double *vals, *b, *dinv;
  double *x, *r, *u, *p, *s, *w;      /* u == r without a preconditioner */
  ohp_cg_state *st;
  double *hist;
  int hist_cap;
  double *parts;                       /* G x 4 partial sums */
  unsigned *ticket;                    /* arrivals at the global synchronisation (zeroed before every launch) */
  unsigned *vflag;                     /* G: in-launch iteration whose V phase CTA b has published (zeroed before every launch) */
  const int2 *dep;                     /* G: [lo, hi] CTAs whose rows CTA b's tiles reference */
  ohp_ar_slot *self_slots;             /* mailboxes when there is one rank (2 parities x kMaxRanks, slot 0 used) */
  unsigned long long *seqs;            /* [0] all-reduces completed, [1] halo pushes completed (monotonic over solves) */
  const int *hrange;                   /* G x (n_nbr + 1): CTA b pushes send-list entries [hrange[b][k], hrange[b][k+1]) ... see host */
  const int *hcontrib;                 /* n_nbr: how many CTAs push to neighbour k */
  unsigned *hcount;                    /* n_nbr arrival counters (zeroed before every launch) */
  int n, n_tiles, chunk_iters, p2p, jacobi, pin_tiles;
  ohp_p2p_dev pd;
  ohp_trace trace;
};

__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(FNC) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

/* bounded spin on an intra-GPU flag: a bug or a dead CTA must not hang the box (same policy as p2p_wait) */
__device__ __forceinline__ void wait_flag_gpu(const unsigned *flag, unsigned target, unsigned long long *err, unsigned long long timeout_ns) {
  if (ld_acquire_gpu(flag) >= target) return;
  if (*(volatile unsigned long long *)err) return;
  const unsigned long long t0 = global_ns();
  while (ld_acquire_gpu(flag) < target) {
    if (global_ns() - t0 > timeout_ns) { *(volatile unsigned long long *)err = 1ull; return; }
  }
}

/* sum of NV values per consumer thread over the CTA's consumers, fixed tree; result valid in thread 0 */
template <int NV> __device__ __forceinline__ void cons_block_sum(double (&v)[NV], double *s_red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) s_red[k * 32 + wid] = v[k];
  }
  cons_sync();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double t = lane < FNC / 32 ? s_red[k * 32 + lane] : 0.0;
#pragma unroll
      for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      v[k] = t;
    }
  }
  cons_sync();
}

template <class IDX>
__global__ void __launch_bounds__(FNC + 32, 1) k_cg_fused(const FusedArgs a) {
  constexpr int T = kSpmvTile, S = FS, NC = FNC;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *vals_s = reinterpret_cast<double *>(smem_raw);
  IDX *cols_s = reinterpret_cast<IDX *>(vals_s + S * T);
  int32_t *rp_s = reinterpret_cast<int32_t *>(cols_s + S * T);
  uint64_t *bars = reinterpret_cast<uint64_t *>(rp_s + S * kTmaRpStride);
  __shared__ double s_red[3 * 32];
  __shared__ ohp_cg_state s_st;            /* every CTA runs the scalar recurrences on its own copy */
  __shared__ volatile int s_exit, s_consumed;
  __shared__ bool s_last;
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, B = blockIdx.x;
  const IDX *colidx = sizeof(IDX) == 2 ? reinterpret_cast<const IDX *>(a.col16) : reinterpret_cast<const IDX *>(a.colidx32);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 33); mbar_init(empty0 + 8 * s, NC / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_exit = 0; s_consumed = 0;
    s_st = *a.st;
  }
  __syncthreads();
  const int t0 = (int)(((int64_t)a.n_tiles * B) / G);
  const int t1 = (int)(((int64_t)a.n_tiles * (B + 1)) / G);

  if (warp == NC / 32) {
    /* ------------- producer warp: streams the CTA's tiles, iteration after iteration ------------- */
    const uint64_t pol_last = l2_policy_evict_last(), pol_first = l2_policy_evict_first();
    long long gi = 0; /* tiles issued since kernel start: stage = gi % S, parity = (gi / S) & 1 */
    bool exiting = false;
    for (int it = 0; it < a.chunk_iters && !exiting; ++it) {
      for (int t = t0; t < t1; ++t, ++gi) {
        const int s = (int)(gi % S);
        const uint32_t par = (uint32_t)(((gi / S) & 1) ^ 1);
        for (;;) { /* wait for the stage to be free, or for the consumers to leave */
          uint32_t ok;
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                       : "=r"(ok) : "r"(empty0 + 8 * s), "r"(par) : "memory");
          if (ok) break;
          if (s_exit) { exiting = true; break; }
        }
        if (exiting) break;
        const int rs = __ldg(a.tile_row + t), re = __ldg(a.tile_row + t + 1);
        if (lane == 0) {
          mbar_expect_tx(full0 + 8 * s, T * (8 + (int)sizeof(IDX)));
          if (a.pin_tiles > 0) {
            const uint64_t pol = (t - t0 < a.pin_tiles) ? pol_last : pol_first;
            tma_load_1d_hint(smem_u32(vals_s + s * T), a.vals + (size_t)t * T, T * 8, full0 + 8 * s, pol);
            tma_load_1d_hint(smem_u32(cols_s + s * T), colidx + (size_t)t * T, T * (int)sizeof(IDX), full0 + 8 * s, pol);
          } else {
            tma_load_1d(smem_u32(vals_s + s * T), a.vals + (size_t)t * T, T * 8, full0 + 8 * s);
            tma_load_1d(smem_u32(cols_s + s * T), colidx + (size_t)t * T, T * (int)sizeof(IDX), full0 + 8 * s);
          }
        }
        int32_t *rp = rp_s + s * kTmaRpStride;
        for (int j = lane; j <= re - rs; j += 32) cp_async4(smem_u32(rp + j), a.rowptr + rs + j);
        if (lane < 2) cp_async4(smem_u32(rp + kSpmvRowCap + 2 + lane), a.tile_row + t + lane);
        cp_async_mbar_arrive_noinc(full0 + 8 * s);
      }
    }
    /* drain: bulk copies issued for tiles the consumers never took must land before the CTA exits */
    while (!s_exit) { }
    for (long long g = s_consumed; g < gi; ++g) mbar_wait(full0 + 8 * (int)(g % S), (uint32_t)((g / S) & 1));
    return;
  }

  /* ---------------------------------- consumer warps ---------------------------------- */
  const int R0 = t0 < t1 ? __ldg(a.tile_row + t0) : 0, R1 = t0 < t1 ? __ldg(a.tile_row + t1) : 0; /* rows that start in my tiles */
  const bool jac = a.jacobi != 0;
  const ohp_p2p_dev &pd = a.pd;
  const unsigned long long ar_base = a.seqs[0], halo_base = a.seqs[1];
  unsigned long long *errflag = a.seqs + 2;
  const unsigned long long timeout_ns = a.p2p ? pd.timeout_ns : 60000000000ull;
  const int2 dep = a.dep[B];
  const int n_nbr = a.p2p ? pd.n_nbr : 0;
  const bool ghost_cta = a.p2p && a.pd.n_nbr > 0 && a.hrange[(size_t)B * (2 * kMaxRanks + 1) + 2 * kMaxRanks] != 0;
  long long gc = 0;   /* tiles consumed */
  int j = 0;          /* iterations of this launch */
  for (int it = 0; it < a.chunk_iters; ++it) {
    if (s_st.done) break;
    ++j;
    const bool init = (s_st.x_pending == 0);
    const int its_now = s_st.its;
    double sums[3] = {0.0, 0.0, 0.0}; /* (r,u), (w,u), (u,u) */
    if (B == 0 && tid == 0) trace_put(a.trace, TR_VEC_BEGIN, its_now);
    /* ---- V phase ---- */
    if (init) {
      for (int i = R0 + tid; i < R1; i += NC) {
        const double bi = a.b[i];
        a.x[i] = 0.0; a.p[i] = 0.0; a.s[i] = 0.0;
        a.r[i] = bi;
        double ui = bi;
        if (jac) { ui = a.dinv[i] * bi; a.u[i] = ui; }
        sums[0] += bi * ui; sums[2] += ui * ui;
      }
    } else {
      const double alpha = s_st.alpha, beta = s_st.beta;
#ifndef OHP_FUSED_U
#define OHP_FUSED_U 4
#endif
      constexpr int U = OHP_FUSED_U; /* elements in flight per thread and stream */
      for (int i0 = R0 + tid; i0 < R1; i0 += U * NC) {
        double rv[U], uv[U], pv[U], sv[U], wv[U], xv[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int i = i0 + k * NC;
          if (i < R1) {
            rv[k] = ld_weak(a.r + i);
            uv[k] = jac ? ld_weak(a.u + i) : rv[k];
            pv[k] = ld_weak(a.p + i); sv[k] = ld_weak(a.s + i); wv[k] = ld_weak(a.w + i); xv[k] = ld_weak(a.x + i);
          }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int i = i0 + k * NC;
          if (i < R1) {
            const double pi = fma(beta, pv[k], uv[k]);
            const double si = fma(beta, sv[k], wv[k]);
            a.p[i] = pi; a.s[i] = si;
            a.x[i] = fma(alpha, pi, xv[k]);
            const double ri = fma(-alpha, si, rv[k]);
            a.r[i] = ri;
            double un = ri;
            if (jac) { un = __ldg(a.dinv + i) * ri; a.u[i] = un; }
            sums[0] += ri * un; sums[2] += un * un;
          }
        }
      }
    }
    cons_sync(); /* the CTA's u rows are written (CTA scope) */
    /* ---- halo push: the rows of mine that neighbouring ranks ghost, straight into their vectors ---- */
    if (ghost_cta) {
      const int *hr = a.hrange + (size_t)B * (2 * kMaxRanks + 1);
      for (int k = 0; k < n_nbr; ++k) {
        const int e0 = hr[2 * k], e1 = hr[2 * k + 1];
        if (e1 < 0) continue;                     /* this CTA does not contribute to neighbour k */
        double *dst = pd.p[0][pd.nbr[k]] + pd.dst_off[k];
        for (int e = e0 + tid; e < e1; e += NC) dst[e - pd.soff[k]] = ld_weak(a.u + pd.send_ldof[e]);
      }
      __threadfence_system();
      cons_sync();
      if (tid < n_nbr && hr[2 * tid + 1] >= 0) {
        const unsigned old = atomicAdd(a.hcount + tid, 1u);
        if (old + 1u == (unsigned)a.hcontrib[tid] * (unsigned)j) { /* every push of this rank for neighbour `tid` has landed */
          __threadfence_system();
          st_release_sys(pd.hflags[pd.nbr[tid]] + pd.rank, halo_base + (unsigned long long)j);
        }
      }
    }
    /* ---- publish "my u rows of iteration j are written"; wait only for the CTAs (and ranks) my tiles read from ---- */
    if (tid == 0) { __threadfence(); st_release_gpu(a.vflag + B, (unsigned)j); }
    {
      const int ndep = dep.y - dep.x + 1;
      for (int q = tid; q < ndep; q += NC) {
        const int c = dep.x + q;
        if (c != B) wait_flag_gpu(a.vflag + c, (unsigned)j, errflag, timeout_ns);
      }
      if (ghost_cta && tid >= 32 && tid < 32 + n_nbr) /* warp 1 waits for the peers' halo flags */
        p2p_wait(pd.hflags[pd.rank] + pd.nbr[tid - 32], halo_base + (unsigned long long)j, pd.seq + 2, pd.timeout_ns);
    }
    cons_sync();
    if (B == 0 && tid == 0) trace_put(a.trace, TR_SPMV_BEGIN, its_now);
    /* ---- S phase: w = A u over this CTA's tiles, (w,u) fused ---- */
    {
      double carry = 0.0;
      int pend = -1;
      for (int t = t0; t < t1; ++t, ++gc) {
        const int s = (int)(gc % S);
        mbar_wait(full0 + 8 * s, (uint32_t)((gc / S) & 1));
        const int32_t *rp = rp_s + s * kTmaRpStride;
        const double *vs = vals_s + s * T;
        const IDX *cs = cols_s + s * T;
        const int rs = rp[kSpmvRowCap + 2], re = rp[kSpmvRowCap + 3];
        const int base = t * T, lim = base + T;
        if (pend >= 0) {
          const int end = rp[0];
          carry = row_dot<IDX, true>(vs, cs, 0, min(end, lim) - base, a.u, carry, pend);
          if (end <= lim) { a.w[pend] = carry; sums[1] += carry * ld_weak(a.u + pend); pend = -1; }
        }
        for (int r = rs + tid; r < re; r += NC) {
          const int ra = rp[r - rs], rb = rp[r - rs + 1];
          const double sum = row_dot<IDX, true>(vs, cs, ra - base, min(rb, lim) - base, a.u, 0.0, r);
          if (rb > lim) { carry = sum; pend = r; }
          else { a.w[r] = sum; sums[1] += sum * ld_weak(a.u + r); }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
      }
      if (pend >= 0) { /* the row runs past this CTA's last tile: finish it straight from HBM */
        const int end = __ldg(a.rowptr + pend + 1);
        for (int k = t1 * T; k < end; ++k) {
          const int c = sizeof(IDX) == 2 ? pend + (int)colidx[k] : (int)colidx[k];
          carry += __ldg(a.vals + k) * ld_weak(a.u + c);
        }
        a.w[pend] = carry;
        sums[1] += carry * ld_weak(a.u + pend);
      }
    }
    if (tid == 0) trace_put(a.trace, TR_SPMV_CTA_DONE, its_now);
    /* ---- the one global synchronisation: partial sums -> rank totals -> every rank's mailbox ---- */
    cons_block_sum<3>(sums, s_red);
    const unsigned long long seq = ar_base + (unsigned long long)j;
    const int par = (int)(seq & 1ull);
    if (tid == 0) {
      a.parts[B * 4 + 0] = sums[0]; a.parts[B * 4 + 1] = sums[1]; a.parts[B * 4 + 2] = sums[2];
      __threadfence();
      s_last = (atomicAdd(a.ticket, 1u) + 1u == (unsigned)G * (unsigned)j);
    }
    cons_sync();
    if (s_last) {
      if (warp == 0) {
        __threadfence();
        double tot[3] = {0.0, 0.0, 0.0};
        for (int q = lane; q < G; q += 32) { /* fixed assignment and fixed tree => the same bits every run */
          tot[0] += __ldcg(a.parts + q * 4 + 0); tot[1] += __ldcg(a.parts + q * 4 + 1); tot[2] += __ldcg(a.parts + q * 4 + 2);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int o = 16; o; o >>= 1) tot[k] += __shfl_xor_sync(0xffffffffu, tot[k], o);
        if (lane == 0) trace_put(a.trace, TR_AR_POSTED, its_now);
        const int nr = a.p2p ? pd.nranks : 1;
        if (lane < nr) { /* lane q stores this rank's totals into rank q's mailbox (its own included) */
          ohp_ar_slot *slot = (a.p2p ? pd.slots[lane] : a.self_slots) + par * kMaxRanks + (a.p2p ? pd.rank : 0);
          slot->v[0] = tot[0]; slot->v[1] = tot[1]; slot->v[2] = tot[2];
          st_release_sys(&slot->seq, seq);
        }
      }
    }
    /* every CTA: poll the local mailboxes, add the ranks' totals in rank order, advance the scalars */
    if (warp == 0) {
      const int nr = a.p2p ? pd.nranks : 1;
      double got[3] = {0.0, 0.0, 0.0};
      if (lane < nr) {
        const ohp_ar_slot *slot = (a.p2p ? pd.slots[pd.rank] : a.self_slots) + par * kMaxRanks + lane;
        p2p_wait(&slot->seq, seq, errflag, timeout_ns);
        got[0] = slot->v[0]; got[1] = slot->v[1]; got[2] = slot->v[2];
      }
      double tot[3] = {0.0, 0.0, 0.0};
      for (int q = 0; q < nr; ++q) {
#pragma unroll
        for (int k = 0; k < 3; ++k) tot[k] += __shfl_sync(0xffffffffu, got[k], q);
      }
      if (lane == 0) {
        if (B == 0) trace_put(a.trace, TR_AR_DONE, its_now);
        cgcg_scalars(&s_st, B == 0 ? a.hist : nullptr, a.hist_cap, tot[0], tot[1], jac ? tot[2] : tot[0]);
      }
    }
    cons_sync();
  }
  if (tid == 0) {
    if (B == 0) { /* one writer: the host reads the scalar block after the launch; sequence bases advance by the iterations run */
      *a.st = s_st;
      a.seqs[0] = ar_base + (unsigned long long)j;
      a.seqs[1] = halo_base + (unsigned long long)j;
    }
    s_consumed = (int)gc; __threadfence_block(); s_exit = 1;
  }
}

/* [lo, hi] of the CTAs whose rows CTA b's tiles reference: b's columns lie in rows [cmin, cmax]; the owner of a row is
 * found in the CTAs' first-row table */
__global__ void k_fused_deps(const int32_t *colidx, const int32_t *rowptr, const int32_t *tile_row, int n_tiles, int n_owned, int G,
                             int tile, int2 *dep, int *ghost) {
  const int b = blockIdx.x;
  const int t0 = (int)(((int64_t)n_tiles * b) / G), t1 = (int)(((int64_t)n_tiles * (b + 1)) / G);
  __shared__ int s_min[32], s_max[32], s_gh[32];
  int cmin = INT_MAX, cmax = -1, gh = 0;
  if (t0 < t1) {
    const int r0 = tile_row[t0], r1 = tile_row[t1];
    const int64_t k0 = rowptr[r0], k1 = rowptr[r1];
    for (int64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
      const int c = colidx[k];
      if (c < n_owned) { cmin = min(cmin, c); cmax = max(cmax, c); } else gh = 1;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
    cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    gh |= __shfl_xor_sync(0xffffffffu, gh, o);
  }
  if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = cmin; s_max[threadIdx.x >> 5] = cmax; s_gh[threadIdx.x >> 5] = gh; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { cmin = min(cmin, s_min[w]); cmax = max(cmax, s_max[w]); gh |= s_gh[w]; }
    cmin = min(s_min[0], cmin); cmax = max(s_max[0], cmax); gh |= s_gh[0];
    int lo = b, hi = b;
    if (cmax >= 0) {
      /* owner of row c = the CTA q with tile_row[t0(q)] <= c < tile_row[t1(q)] */
      auto owner = [&](int c) {
        int l = 0, h = G - 1;
        while (l < h) {
          const int mid = (l + h + 1) >> 1;
          const int tq = (int)(((int64_t)n_tiles * mid) / G);
          if (tile_row[tq] <= c) l = mid; else h = mid - 1;
        }
        return l;
      };
      lo = min(b, owner(cmin)); hi = max(b, owner(cmax));
    }
    dep[b] = make_int2(lo, hi);
    ghost[b] = gh;
  }
}

}  // namespace

/* per-mesh tables of the fused kernel, built on first use */
struct ohp_fused_plan {
  int G = 0;
  int2 *dep = nullptr;
  int *hrange = nullptr, *hcontrib = nullptr;
  unsigned *sync = nullptr;        /* ticket (1) | vflag (G) | hcount (kMaxRanks) */
  double *parts = nullptr;
  ohp_ar_slot *self_slots = nullptr;
  unsigned long long *self_seqs = nullptr;
};

void ohp_fused_plan_free(ohp_context ctx, ohp_fused_plan *p) {
  if (!p) return;
  OHP_FREE(ctx, p->dep); OHP_FREE(ctx, p->hrange); OHP_FREE(ctx, p->hcontrib); OHP_FREE(ctx, p->sync); OHP_FREE(ctx, p->parts);
  OHP_FREE(ctx, p->self_slots); OHP_FREE(ctx, p->self_seqs);
  delete p;
}

static int fused_plan_build(ohp_mesh m, bool p2p) {
  ohp_context ctx = m->ctx;
  cudaStream_t st = ctx->stream;
  ohp_fused_plan *pl = new ohp_fused_plan();
  m->fused = pl;
  const int G = std::max(1, std::min(ctx->num_sms, m->n_tiles));
  pl->G = G;
  OHP_CUDA(OHP_MALLOC(ctx, &pl->dep, sizeof(int2) * G));
  int *d_ghost;
  OHP_CUDA(OHP_MALLOC(ctx, &d_ghost, sizeof(int) * G));
  k_fused_deps<<<G, 256, 0, st>>>(m->colidx, m->rowptr, m->tile_row, m->n_tiles, m->n_owned, G, kSpmvTile, pl->dep, d_ghost);
  OHP_CHECK_LAUNCH(ctx);
  std::vector<int> ghost(G), tile_row(m->n_tiles + 1);
  OHP_CUDA(cudaMemcpyAsync(ghost.data(), d_ghost, sizeof(int) * G, cudaMemcpyDeviceToHost, st));
  OHP_CUDA(cudaMemcpyAsync(tile_row.data(), m->tile_row, sizeof(int) * (m->n_tiles + 1), cudaMemcpyDeviceToHost, st));
  OHP_CUDA(cudaStreamSynchronize(st));
  OHP_FREE(ctx, d_ghost);
  OHP_CUDA(OHP_MALLOC(ctx, &pl->sync, sizeof(unsigned) * (1 + G + kMaxRanks)));
  OHP_CUDA(OHP_MALLOC(ctx, &pl->parts, sizeof(double) * 4 * G));
  OHP_CUDA(OHP_MALLOC(ctx, &pl->self_slots, sizeof(ohp_ar_slot) * 2 * kMaxRanks));
  OHP_CUDA(OHP_MALLOC(ctx, &pl->self_seqs, sizeof(unsigned long long) * 4));
  OHP_CUDA(cudaMemsetAsync(pl->self_slots, 0, sizeof(ohp_ar_slot) * 2 * kMaxRanks, st));
  OHP_CUDA(cudaMemsetAsync(pl->self_seqs, 0, sizeof(unsigned long long) * 4, st));
  /* halo contributions: which CTA owns which entries of the send list (grouped by neighbour, ascending dof inside a group).
   * Row of table b: [e0, e1) per neighbour k at [2k, 2k+1] (e1 = -1: no contribution), last entry = "reads ghost columns or pushes". */
  const int W = 2 * kMaxRanks + 1;
  std::vector<int> hrange((size_t)G * W, 0), hcontrib(kMaxRanks, 0);
  for (int b = 0; b < G; ++b) for (int k = 0; k < kMaxRanks; ++k) hrange[(size_t)b * W + 2 * k + 1] = -1;
  if (p2p) {
    const ohp_p2p_dev &pd = m->p2p->dev;
    std::vector<int32_t> send(std::max(1, pd.soff[pd.n_nbr]));
    OHP_CUDA(cudaMemcpyAsync(send.data(), pd.send_ldof, sizeof(int32_t) * pd.soff[pd.n_nbr], cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
    std::vector<int> row0(G + 1);
    for (int b = 0; b <= G; ++b) row0[b] = m->n_tiles ? tile_row[(int)(((int64_t)m->n_tiles * b) / G)] : 0;
    row0[G] = m->n_owned;
    for (int k = 0; k < pd.n_nbr; ++k) {
      bool sorted = true;
      for (int e = pd.soff[k] + 1; e < pd.soff[k + 1]; ++e) sorted &= send[e - 1] <= send[e];
      if (!sorted) OHP_FAIL(OHP_ERR_LIB, "fused CG: the send list of neighbour %d is not sorted by dof", pd.nbr[k]);
      int e = pd.soff[k];
      for (int b = 0; b < G; ++b) {
        const int e0 = e;
        while (e < pd.soff[k + 1] && send[e] < row0[b + 1]) ++e;
        if (e > e0) { hrange[(size_t)b * W + 2 * k] = e0; hrange[(size_t)b * W + 2 * k + 1] = e; hcontrib[k]++; }
      }
      if (hcontrib[k] == 0) { /* nothing to send to k, but its flag must still advance: CTA 0 makes the (empty) contribution */
        hrange[2 * k] = pd.soff[k]; hrange[2 * k + 1] = pd.soff[k]; hcontrib[k] = 1;
      }
    }
    for (int b = 0; b < G; ++b) {
      bool any = ghost[b] != 0;
      for (int k = 0; k < pd.n_nbr; ++k) any |= hrange[(size_t)b * W + 2 * k + 1] >= 0;
      hrange[(size_t)b * W + 2 * kMaxRanks] = any ? 1 : 0;
    }
    /* a CTA that pushes but reads no ghost column would still wait for the peers' flags (harmless); a CTA that reads ghost
     * columns but pushes nothing waits too (required) */
  }
  OHP_CUDA(OHP_MALLOC(ctx, &pl->hrange, sizeof(int) * hrange.size()));
  OHP_CUDA(OHP_MALLOC(ctx, &pl->hcontrib, sizeof(int) * kMaxRanks));
  OHP_CUDA(cudaMemcpyAsync(pl->hrange, hrange.data(), sizeof(int) * hrange.size(), cudaMemcpyHostToDevice, st));
  OHP_CUDA(cudaMemcpyAsync(pl->hcontrib, hcontrib.data(), sizeof(int) * kMaxRanks, cudaMemcpyHostToDevice, st));
  OHP_CUDA(cudaStreamSynchronize(st));
  return 0;
}

/* host side: runs the solve in chunks of `check_every` iterations per cooperative launch */
int ohp_cg_fused_solve(ohp_mat A, ohp_vec b, ohp_vec x, const ohp_ksp_options *o, ohp_ksp_result *res, bool p2p) {
  ohp_mesh m = A->mesh;
  ohp_context ctx = m->ctx;
  cudaStream_t st = ctx->stream;
  const int n = m->n_owned;
  const size_t len = (size_t)std::max(1, n + m->n_ghost);
  const bool jac = (o->pc == OHP_PC_JACOBI);
  auto need = [&](double **ptr) -> int {
    if (!*ptr) { OHP_CUDA(OHP_MALLOC(ctx, ptr, sizeof(double) * len)); OHP_CUDA(cudaMemsetAsync(*ptr, 0, sizeof(double) * len, st)); }
    return 0;
  };
  OHP_TRY(need(&A->r)); OHP_TRY(need(&A->p)); OHP_TRY(need(&A->w)); OHP_TRY(need(&A->s));
  if (jac) { OHP_TRY(need(&A->z)); OHP_TRY(need(&A->dinv)); OHP_TRY(ohp_extract_dinv(A)); }
  if (!m->fused) OHP_TRY(fused_plan_build(m, p2p));
  ohp_fused_plan *pl = m->fused;
  const bool idx16 = m->col16 != nullptr;
  const int smem = tma_smem_bytes(FS, idx16 ? 2 : 4);
  if (idx16) OHP_CUDA(cudaFuncSetAttribute(k_cg_fused<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  else OHP_CUDA(cudaFuncSetAttribute(k_cg_fused<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  static const ohp_p2p_dev no_p2p = {};
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.rowptr = m->rowptr; a.tile_row = m->tile_row; a.colidx32 = m->colidx; a.col16 = m->col16; a.vals = A->vals; a.b = b->d; a.dinv = A->dinv;
  a.x = x->d; a.p = A->p; a.s = A->s; a.w = A->w;
  a.pd = p2p ? m->p2p->dev : no_p2p;
  a.p2p = p2p ? 1 : 0;
  /* the SpMV input (u) must sit in the exchange heap when neighbours push into its ghost tail */
  double *uvec = p2p ? a.pd.p[0][ctx->rank] : (jac ? A->z : A->r);
  a.u = uvec;
  a.r = jac ? A->r : uvec;
  a.st = ctx->d_cg;
  a.parts = pl->parts; a.ticket = pl->sync; a.vflag = pl->sync + 1; a.hcount = pl->sync + 1 + pl->G;
  a.dep = pl->dep; a.hrange = pl->hrange; a.hcontrib = pl->hcontrib;
  a.self_slots = pl->self_slots;
  a.seqs = p2p ? a.pd.seq : pl->self_seqs;
  a.n = n; a.n_tiles = m->n_tiles; a.jacobi = jac ? 1 : 0;
  {
    static const char *env = getenv("OHP_L2PIN_MB");
    const double mb = env ? atof(env) : 0.0;
    a.pin_tiles = mb > 0.0 ? std::max(1, (int)(mb * 1e6 / ((double)kSpmvTile * (8 + (idx16 ? 2 : 4))) / pl->G)) : 0;
  }
  double *hist = nullptr;
  if (o->history && o->history_cap > 0) {
    if (A->hist_cap < o->history_cap) {
      OHP_FREE(ctx, A->hist);
      OHP_CUDA(OHP_MALLOC(ctx, &A->hist, sizeof(double) * o->history_cap));
      A->hist_cap = o->history_cap;
    }
    hist = A->hist;
    OHP_CUDA(cudaMemsetAsync(hist, 0, sizeof(double) * o->history_cap, st));
  }
  a.hist = hist; a.hist_cap = hist ? o->history_cap : 0;
  a.trace = ohp_cg_trace_open(ctx);
  ohp_cg_state init;
  memset(&init, 0, sizeof(init));
  init.rtol = o->rtol; init.abstol = o->abstol; init.dtol = o->dtol; init.maxit = o->max_it;
  *ctx->h_cg = init;
  OHP_CUDA(cudaMemcpyAsync(ctx->d_cg, ctx->h_cg, sizeof(init), cudaMemcpyHostToDevice, st));
  a.chunk_iters = o->check_every > 0 ? o->check_every : 500;
  if (p2p) OHP_TRY(ohp_p2p_entry_barrier(m));
  struct Ev { cudaEvent_t e = nullptr; ~Ev() { if (e) cudaEventDestroy(e); } } e0, e1;
  OHP_CUDA(cudaEventCreate(&e0.e)); OHP_CUDA(cudaEventCreate(&e1.e));
  OHP_CUDA(cudaEventRecord(e0.e, st));
  int launched = 0;
  for (;;) {
    OHP_CUDA(cudaMemsetAsync(pl->sync, 0, sizeof(unsigned) * (1 + pl->G + kMaxRanks), st));
    void *params[] = {(void *)&a};
    if (idx16) OHP_CUDA(cudaLaunchCooperativeKernel((const void *)k_cg_fused<int16_t>, dim3(pl->G), dim3(FNC + 32), params, (size_t)smem, st));
    else OHP_CUDA(cudaLaunchCooperativeKernel((const void *)k_cg_fused<int32_t>, dim3(pl->G), dim3(FNC + 32), params, (size_t)smem, st));
    ctx->launches++;
    OHP_CUDA(cudaMemcpyAsync(ctx->h_cg, ctx->d_cg, sizeof(ohp_cg_state), cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
    if (ctx->h_cg->done) break;
    launched += a.chunk_iters;
    if (launched > o->max_it + 2 * a.chunk_iters) OHP_FAIL(OHP_ERR_LIB, "ohp_ksp_cg (fused): iteration control did not terminate");
  }
  OHP_CUDA(cudaEventRecord(e1.e, st));
  OHP_CUDA(cudaEventSynchronize(e1.e));
  res->total_ms = 0.f;
  cudaEventElapsedTime(&res->total_ms, e0.e, e1.e);
  res->its = ctx->h_cg->its; res->reason = ctx->h_cg->reason; res->rnorm = ctx->h_cg->dp; res->rnorm0 = ctx->h_cg->rnorm0;
  res->spmv_ms = 0.f; res->spmv_samples = 0; res->algo = OHP_ALGO_FUSED_PERSISTENT;
  ohp_cg_trace_dump(ctx, "fused");
  if (p2p) OHP_TRY(ohp_p2p_check(m, "ohp_ksp_cg (fused)"));
  else {
    unsigned long long err = 0;
    OHP_CUDA(cudaMemcpyAsync(&err, pl->self_seqs + 2, sizeof(err), cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
    if (err) { cudaMemsetAsync(pl->self_seqs + 2, 0, sizeof(err), st); OHP_FAIL(OHP_ERR_LIB, "ohp_ksp_cg (fused): the grid synchronisation timed out"); }
  }
  if (hist) {
    const int ncopy = std::min(o->history_cap, res->its + 1);
    OHP_CUDA(cudaMemcpyAsync(o->history, hist, sizeof(double) * ncopy, cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}
