/*
 * ohp_linalg.cu -- Vec / Mat kernels and the CG solve.
 *
 * Replaces [PETSc-internal] MatMult_SeqAIJ/MPIAIJ, VecDot/Norm/AXPY/AYPX/Set/Chop (used
 * directly at ex12.cpp:1626-1675 and inside KSP) and KSPSolve_CG (selected by -ksp_type cg
 * at ex12.cpp:1543; tolerances of the reference CTests at CMakeLists.txt:53).
 *
 * Everything is HBM-bandwidth bound, FP64, no tensor cores:
 *   - SpMV is CSR-stream: a block streams one tile of vals/colidx fully coalesced, leaves the
 *     products in shared memory, then each thread sums whole rows in column order (the same
 *     order as a serial CSR loop).  (p, Ap) is fused into the same kernel.
 *   - axpy-type updates are fused with the dot products they feed.
 *   - reductions are two-stage and deterministic: per-block partials, summed in a fixed order by
 *     the last block to arrive (ticket counter); the CG scalars (alpha, beta, norms, iteration
 *     count, converged reason) live in HBM and are updated by that last block, so an
 *     iteration needs no host round trip.  The host polls the `done` flag every
 *     `check_every` iterations; kernels launched past convergence return immediately.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "ohp_dev.cuh"
#include "ohp_internal.h"

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)std::max<int64_t>(1, (n + bs - 1) / bs); }

/* ------------------------------------------------------------------------------------ */
/* CG scalar recurrences (KSPSolve_CG + KSPConvergedDefault), run by one thread          */
/* ------------------------------------------------------------------------------------ */
__device__ __forceinline__ void cg_converged(ohp_cg_state *st, double dp) {
  if (dp != dp) st->reason = -9;                                  /* KSP_DIVERGED_NANORINF */
  else if (dp <= st->ttol) st->reason = (dp < st->abstol) ? 3 : 2; /* ATOL : RTOL */
  else if (st->its > 0 && dp >= st->dtol * st->rnorm0) st->reason = -4;
}
__device__ __forceinline__ void cg_loop_head(ohp_cg_state *st) {
  /* `while (!reason && i < maxit) { if (beta == 0) -> CONVERGED_ATOL ...` */
  if (!st->reason && st->its >= st->maxit) st->reason = -3;
  if (!st->reason && st->beta == 0.0) st->reason = 3;
  if (st->reason) st->done = 1;
}
__device__ void cg_fin_init(ohp_cg_state *st, double zz, double zr, double *hist) {
  const double dp = sqrt(zz);
  st->dp = dp; st->rnorm0 = dp; st->ttol = fmax(st->rtol * dp, st->abstol);
  st->beta = zr; st->betaold = 1.0; st->its = 0; st->reason = 0; st->done = 0; st->x_pending = 0;
  if (hist) hist[0] = dp;
  cg_converged(st, dp);
  cg_loop_head(st);
}
__device__ void cg_fin_spmv(ohp_cg_state *st, double dpi) {
  st->dpi = dpi;
  st->betaold = st->beta;
  if (!(dpi > 0.0)) { st->reason = (dpi != dpi) ? -9 : -8; st->its += 1; st->done = 1; st->x_pending = 0; return; }
  st->alpha = st->beta / dpi;
  st->x_pending = 1; /* applied by the next k_cg_pupdate, or by k_cg_final_x when the iteration stops */
}
__device__ void cg_fin_update(ohp_cg_state *st, double zz, double zr, double *hist, int hist_cap) {
  const double dp = sqrt(zz);
  st->dp = dp;
  st->its += 1;
  if (hist && st->its < hist_cap) hist[st->its] = dp;
  cg_converged(st, dp);
  if (!st->reason) st->beta = zr;
  cg_loop_head(st);
}

__global__ void k_cg_fin(ohp_cg_state *st, const double *sums, int which, double *hist, int hist_cap) {
  if (which != 0 && st->done) return;
  if (which == 0) cg_fin_init(st, sums[0], sums[1], hist);
  else if (which == 1) cg_fin_spmv(st, sums[0]);
  else cg_fin_update(st, sums[0], sums[1], hist, hist_cap);
}

/* which buffer the CG SpMV multiplies: classic recurrence -> P[its & 1]; single-reduction recurrence ->
 * U[0] on the first pass, then the buffer the vector kernel of this iteration wrote, U[(its + 1) & 1] */
__device__ __forceinline__ const double *cg_pick_input(const ohp_cg_state *st, const ohp_cgcg_epi &epi, const double *x0,
                                                       const double *x1) {
  if (epi.pipe) return x0; /* pipelined recurrence: the host names the buffer (it knows the pass number) */
  if (epi.vpart) return (st->x_pending != 0 && ((st->its + 1) & 1)) ? x1 : x0;
  return (st->its & 1) ? x1 : x0;
}

/* Pipelined recurrence: run by ONE extra warp of CTA 0 of the SpMV n = A m, concurrently with the matrix stream.  Adds
 * the partial sums (r,u), (w,u), (u,u) that the vector kernel's blocks left (fixed order), all-reduces them over the
 * ranks' mailboxes and advances the scalar recurrences, so that neither the reduction nor the all-reduce nor rank skew
 * of up to one SpMV duration is on the critical path of the iteration: the next vector kernel finds alpha, beta and
 * `done` ready.  (epi.pipe == 2: a SpMV of the pipelined solver with nothing to reduce -- set-up and replacement passes.) */
__device__ __forceinline__ void pipe_reduce(ohp_cg_state *st, int mode, const ohp_p2p_dev &pd, const ohp_cgcg_epi &epi, int lane) {
  double t[3] = {0.0, 0.0, 0.0};
  for (int b0 = 0; b0 < epi.nvpart; b0 += 32 * 8) { /* 24 loads in flight per lane, added in a fixed order */
    double a[8][3];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = b0 + j * 32 + lane;
#pragma unroll
      for (int k = 0; k < 3; ++k) a[j][k] = b < epi.nvpart ? __ldcg(epi.vpart + 3 * b + k) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int k = 0; k < 3; ++k) t[k] += a[j][k];
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o; o >>= 1) t[k] += __shfl_xor_sync(0xffffffffu, t[k], o);
  }
  const int its_now = st->its;
  if (lane == 0) trace_put(epi.trace, TR_AR_POSTED, its_now);
  if (mode == OHP_MODE_P2P) p2p_allreduce_warp<3>(pd, t, lane);
  if (lane == 0) {
    trace_put(epi.trace, TR_AR_DONE, its_now);
    cgcg_scalars(st, epi.hist, epi.hist_cap, t[0], t[1], epi.jacobi ? t[2] : t[0]);
    trace_put(epi.trace, TR_SPMV_END, its_now);
  }
}

/* run by every thread of the LAST block of a CG SpMV (dot[0] = (x, A x) valid in thread 0) */
__device__ __forceinline__ void cg_spmv_epilogue(ohp_cg_state *st, double (&dot)[1], double *sums, int mode,
                                                 const ohp_p2p_dev &pd, const ohp_cgcg_epi &epi, int its_now = 0) {
  if (threadIdx.x == 0) trace_put(epi.trace, TR_SPMV_LAST_BLOCK, its_now);
  if (!epi.vpart) { /* classic: (p, Ap) -> alpha */
    if (mode == OHP_MODE_P2P) p2p_allreduce<1>(pd, dot);
    if (threadIdx.x == 0) {
      trace_put(epi.trace, TR_AR_DONE, its_now);
      sums[0] = dot[0];
      if (mode != OHP_MODE_NCCL) cg_fin_spmv(st, dot[0]);
    }
    return;
  }
  /* single reduction: add the vector kernel's per-block partials of (r,u), (u,u) in a fixed order */
  __shared__ double s_red3[3 * 32];
  double tot[3] = {0.0, threadIdx.x == 0 ? dot[0] : 0.0, 0.0};
  for (int b = threadIdx.x; b < epi.nvpart; b += blockDim.x) { tot[0] += __ldcg(epi.vpart + 2 * b); tot[2] += __ldcg(epi.vpart + 2 * b + 1); }
  block_sum<3>(tot, s_red3);
  if (threadIdx.x == 0) trace_put(epi.trace, TR_AR_POSTED, its_now);
  if (mode == OHP_MODE_P2P) p2p_allreduce<3>(pd, tot);
  if (threadIdx.x == 0) {
    trace_put(epi.trace, TR_AR_DONE, its_now);
    cgcg_scalars(st, epi.hist, epi.hist_cap, tot[0], tot[1], epi.jacobi ? tot[2] : tot[0]);
    trace_put(epi.trace, TR_SPMV_END, its_now);
  }
}

/* ------------------------------------------------------------------------------------ */
/* SpMV (CSR-stream) with optional fused (x, y) dot                                      */
/* ------------------------------------------------------------------------------------ */
template <bool CG>
__global__ void __launch_bounds__(kSpmvThreads)
k_spmv_stream(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
              const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y,
              const int32_t *__restrict__ tile_row, int n_tiles, ohp_cg_state *st, double *partials,
              unsigned *ticket, double *sums, int mode, const double *__restrict__ x_alt, const ohp_p2p_dev pd,
              const ohp_cgcg_epi epi) {
  extern __shared__ double s_prod[];
  if (CG && st->done) return;
  if (CG) x = cg_pick_input(st, epi, x, x_alt); /* double-buffered SpMV input */
  double dot[1] = {0.0};
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int r0 = __ldg(tile_row + t), r1 = __ldg(tile_row + t + 1);
    if (r0 == r1) continue;
    const int k0 = __ldg(rowptr + r0), k1 = __ldg(rowptr + r1);
    for (int k = k0 + threadIdx.x; k < k1; k += kSpmvThreads)
      s_prod[k - k0] = __ldcs(vals + k) * x[__ldcs(colidx + k)];
    __syncthreads();
    for (int r = r0 + threadIdx.x; r < r1; r += kSpmvThreads) {
      const int a = __ldg(rowptr + r) - k0, b = __ldg(rowptr + r + 1) - k0;
      double s = 0.0;
      for (int i = a; i < b; ++i) s += s_prod[i];
      y[r] = s;
      if (CG) dot[0] += s * x[r];
    }
    __syncthreads();
  }
  if (CG) {
    if (grid_sum<1>(dot, partials, ticket)) cg_spmv_epilogue(st, dot, sums, mode, pd, epi);
  }
}

/* ------------------------------------------------------------------------------------ */
/* SpMV, sm_100a path: persistent, warp-specialised, TMA-pipelined.                      */
/*                                                                                       */
/* One CTA per SM owns a contiguous run of fixed-size nnz tiles.  A producer warp streams  */
/* each tile's vals (8 B) and colidx (4 B) into a ring of shared-memory stages with 1-D   */
/* bulk async copies (cp.async.bulk -> mbarrier complete_tx) and stages the tile's rowptr  */
/* slice with cp.async; consumer warps wait on the stage's `full` mbarrier, sum whole rows */
/* (one thread per row, columns in ascending order: the serial CSR order) gathering x     */
/* through the read-only path, and release the stage through its `empty` mbarrier.  No     */
/* __syncthreads in the steady state; bytes in flight per SM = STAGES x 24 KB, which is    */
/* what hides HBM latency (the first CSR-stream kernel was latency-bound at 47 % of peak). */
/* A row that straddles a tile boundary is carried in a register by the thread that       */
/* started it (same thread finishes it in the next tile), so the order stays serial.      */
/* ------------------------------------------------------------------------------------ */
/* the consumer warps' tile loop; COHERENT selects plain loads of x (CTAs whose columns include ghost dofs that
 * peers write over NVLink while the kernel runs) instead of the read-only path */
template <bool CG, int S, int NC, class IDX, bool SPLIT, bool COHERENT>
__device__ __forceinline__ double spmv_consume(const int32_t *__restrict__ rowptr, const IDX *__restrict__ colidx,
                                               const double *__restrict__ vals,
                                               const double *x, double *__restrict__ y, const double *vals_s, const IDX *cols_s,
                                               const int32_t *rp_s, uint32_t full0, uint32_t empty0, int t0, int t1) {
  constexpr int T = kSpmvTile;
  const int tid = threadIdx.x, lane = tid & 31;
  double dot = 0.0, carry = 0.0;
  int pend = -1;
  for (int t = t0; t < t1; ++t) {
    const int i = t - t0, s = i % S;
    mbar_wait(full0 + 8 * s, (i / S) & 1);
    const int32_t *rp = rp_s + s * kTmaRpStride;
    const double *vs = vals_s + s * T;
    const IDX *cs = cols_s + s * T;
    const int rs = rp[kSpmvRowCap + 2], re = rp[kSpmvRowCap + 3];
    const int base = t * T, lim = base + T;
    if (pend >= 0) { /* finish (or continue) the row this thread carried over the tile boundary */
      const int end = rp[0];
      carry = row_dot<IDX, COHERENT>(vs, cs, 0, min(end, lim) - base, x, carry, pend);
      if (end <= lim) { y[pend] = carry; if (CG) dot += carry * (COHERENT ? ld_weak(x + pend) : __ldg(x + pend)); pend = -1; }
    }
    if (!SPLIT) {
      for (int r = rs + tid; r < re; r += NC) {
        const int a = rp[r - rs], b = rp[r - rs + 1];
        const double sum = row_dot<IDX, COHERENT>(vs, cs, a - base, min(b, lim) - base, x, 0.0, r);
        if (b > lim) { carry = sum; pend = r; }
        else { y[r] = sum; if (CG) dot += sum * (COHERENT ? ld_weak(x + r) : __ldg(x + r)); }
      }
    } else {
      /* SPLIT (matrices with long rows, e.g. 15 non-zeros per row in 3-D): when a tile starts few rows, 2 or 4
       * adjacent lanes share a row so that it still costs one gather latency; their partial sums are combined in
       * a fixed butterfly order.  The choice depends on the tile only, so results stay reproducible. */
      const int nrows = re - rs;
      const int tpr = (nrows * 4 <= NC) ? 4 : (nrows * 2 <= NC) ? 2 : 1;
      const int sub = tid & (tpr - 1), grp = tid / tpr, ngrp = NC / tpr;
      for (int rb = rs; rb < re; rb += ngrp) { /* uniform trip count: the shuffles below need the whole warp */
        const int r = rb + grp;
        const bool valid = r < re;
        int a = 0, b = 0;
        if (valid) { a = rp[r - rs]; b = rp[r - rs + 1]; }
        const int bl = min(b, lim);
        int ca = a, cb = bl;
        if (tpr > 1) {
          if (b > lim) { if (sub != 0) cb = ca; }          /* a straddling row stays with its leader (serial order) */
          else { const int chunk = (bl - a + tpr - 1) / tpr; ca = min(a + sub * chunk, bl); cb = min(ca + chunk, bl); }
        }
        double part = (valid && cb > ca) ? row_dot<IDX, COHERENT>(vs, cs, ca - base, cb - base, x, 0.0, r) : 0.0;
        if (tpr > 1) part += __shfl_xor_sync(0xffffffffu, part, 1); /* tpr is uniform over the CTA */
        if (tpr > 2) part += __shfl_xor_sync(0xffffffffu, part, 2);
        if (valid && sub == 0) {
          if (b > lim) { carry = part; pend = r; }
          else { y[r] = part; if (CG) dot += part * (COHERENT ? ld_weak(x + r) : __ldg(x + r)); }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * s);
  }
  if (pend >= 0) { /* the row runs past this CTA's last tile: finish it straight from HBM */
    const int end = __ldg(rowptr + pend + 1);
    for (int k = t1 * T; k < end; ++k) {
      const int c = sizeof(IDX) == 2 ? pend + (int)colidx[k] : (int)colidx[k];
      carry += vals[k] * (COHERENT ? ld_weak(x + c) : __ldg(x + c));
    }
    y[pend] = carry;
    if (CG) dot += carry * (COHERENT ? ld_weak(x + pend) : __ldg(x + pend));
  }
  return dot;
}


/* spmv_consume with the output step abstracted: emit(row, sum) is called exactly once per row, by one thread, with the
 * row's sum taken in serial CSR order (same carry / lane-sharing rules as above).  Used by the fused pipelined-CG kernel,
 * whose emit applies the whole vector update of the row while n = (A m)[row] is still in a register. */
template <int S, int NC, class IDX, bool SPLIT, bool COHERENT, class Emit>
__device__ __forceinline__ void spmv_consume_emit(const int32_t *__restrict__ rowptr, const IDX *__restrict__ colidx,
                                                  const double *__restrict__ vals, const double *x, const double *vals_s,
                                                  const IDX *cols_s, const int32_t *rp_s, uint32_t full0, uint32_t empty0,
                                                  int t0, int t1, Emit &emit) {
  constexpr int T = kSpmvTile;
  const int tid = threadIdx.x, lane = tid & 31;
  double carry = 0.0;
  int pend = -1;
  for (int t = t0; t < t1; ++t) {
    const int i = t - t0, s = i % S;
    mbar_wait(full0 + 8 * s, (i / S) & 1);
    const int32_t *rp = rp_s + s * kTmaRpStride;
    const double *vs = vals_s + s * T;
    const IDX *cs = cols_s + s * T;
    const int rs = rp[kSpmvRowCap + 2], re = rp[kSpmvRowCap + 3];
    const int base = t * T, lim = base + T;
    if (pend >= 0) {
      const int end = rp[0];
      carry = row_dot<IDX, COHERENT>(vs, cs, 0, min(end, lim) - base, x, carry, pend);
      if (end <= lim) { emit(pend, carry); pend = -1; }
    }
    if (!SPLIT) {
      for (int r = rs + tid; r < re; r += NC) {
        const int a = rp[r - rs], b = rp[r - rs + 1];
        if (b > lim) { carry = row_dot<IDX, COHERENT>(vs, cs, a - base, lim - base, x, 0.0, r); pend = r; }
        else {
          auto opnd = emit.load(r); /* the row's vector operands travel while the row is summed */
          const double sum = row_dot<IDX, COHERENT>(vs, cs, a - base, b - base, x, 0.0, r);
          emit.store(r, sum, opnd);
        }
      }
    } else {
      const int nrows = re - rs;
      const int tpr = (nrows * 4 <= NC) ? 4 : (nrows * 2 <= NC) ? 2 : 1;
      const int sub = tid & (tpr - 1), grp = tid / tpr, ngrp = NC / tpr;
      for (int rb = rs; rb < re; rb += ngrp) {
        const int r = rb + grp;
        const bool valid = r < re;
        int a = 0, b = 0;
        if (valid) { a = rp[r - rs]; b = rp[r - rs + 1]; }
        const int bl = min(b, lim);
        int ca = a, cb = bl;
        if (tpr > 1) {
          if (b > lim) { if (sub != 0) cb = ca; }
          else { const int chunk = (bl - a + tpr - 1) / tpr; ca = min(a + sub * chunk, bl); cb = min(ca + chunk, bl); }
        }
        double part = (valid && cb > ca) ? row_dot<IDX, COHERENT>(vs, cs, ca - base, cb - base, x, 0.0, r) : 0.0;
        if (tpr > 1) part += __shfl_xor_sync(0xffffffffu, part, 1);
        if (tpr > 2) part += __shfl_xor_sync(0xffffffffu, part, 2);
        if (valid && sub == 0) {
          if (b > lim) { carry = part; pend = r; }
          else emit(r, part);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * s);
  }
  if (pend >= 0) {
    const int end = __ldg(rowptr + pend + 1);
    for (int k = t1 * T; k < end; ++k) {
      const int c = sizeof(IDX) == 2 ? pend + (int)colidx[k] : (int)colidx[k];
      carry += vals[k] * (COHERENT ? ld_weak(x + c) : __ldg(x + c));
    }
    emit(pend, carry);
  }
}


/* The same consumer loop with the consumer warps split into GR groups that work on DIFFERENT tiles (group g takes the
 * CTA's tiles g, g + GR, ...): one tile's dependency chain (mbarrier -> rowptr -> columns -> gather -> FMA -> vector
 * loads -> stores, ~2 us in the fused pipelined-CG kernel against 0.9 us per tile at HBM rate) no longer bounds the CTA.
 * A row that runs past its tile is finished by the thread that started it straight from global memory (at most one row
 * per tile), so the next tile's group never waits for a carry and the summation order stays the serial CSR order.  emit.load(row) issues the row's vector operands
 * before the row is summed, emit.store(row, sum, operands) applies the update. */
template <int S, int NCT, int GR, class IDX, bool COHERENT, class Emit>
__device__ __forceinline__ void spmv_consume_groups(const int32_t *__restrict__ rowptr, const IDX *__restrict__ colidx,
                                                    const double *__restrict__ vals, const double *x, const double *vals_s,
                                                    const IDX *cols_s, const int32_t *rp_s, uint32_t full0, uint32_t empty0,
                                                    int t0, int t1, Emit &emit) {
  constexpr int T = kSpmvTile, NG = NCT / GR;
  const int tid = threadIdx.x, lane = tid & 31;
  const int grp = tid / NG, gtid = tid - grp * NG;
  for (int t = t0 + grp; t < t1; t += GR) {
    const int i = t - t0, s = i % S;
    mbar_wait(full0 + 8 * s, (i / S) & 1);
    const int32_t *rp = rp_s + s * kTmaRpStride;
    const double *vs = vals_s + s * T;
    const IDX *cs = cols_s + s * T;
    const int rs = rp[kSpmvRowCap + 2], re = rp[kSpmvRowCap + 3];
    const int base = t * T, lim = base + T;
    for (int r = rs + gtid; r < re; r += NG) {
      const int a = rp[r - rs], b = rp[r - rs + 1];
      auto opnd = emit.load(r);
      double sum;
      if (b <= lim) sum = row_dot<IDX, COHERENT>(vs, cs, a - base, b - base, x, 0.0, r);
      else {
        /* the row's tail lies in the next tile (another group's): add it in order straight from global memory */
        sum = row_dot<IDX, COHERENT>(vs, cs, a - base, lim - base, x, 0.0, r);
        sum = row_dot<IDX, COHERENT>(vals + lim, colidx + lim, 0, b - lim, x, sum, r);
      }
      emit.store(r, sum, opnd);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * s);
  }
}

template <bool CG, int kTmaStages, int kTmaConsumers, class IDX, bool SPLIT, int MINB = 1>
__global__ void __launch_bounds__(kTmaConsumers + 64, MINB)
k_spmv_tma(const int32_t *__restrict__ rowptr, const IDX *__restrict__ colidx, const double *__restrict__ vals, const double *x, double *__restrict__ y, const int32_t *__restrict__ tile_row, int n_tiles,
           ohp_cg_state *st, double *partials, unsigned *ticket, double *sums, int mode,
           const double *x_alt, const ohp_p2p_dev pd, const ohp_cgcg_epi epi) {
  constexpr int T = kSpmvTile, S = kTmaStages, NC = kTmaConsumers;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *vals_s = reinterpret_cast<double *>(smem_raw);
  IDX *cols_s = reinterpret_cast<IDX *>(vals_s + S * T);
  int32_t *rp_s = reinterpret_cast<int32_t *>(cols_s + S * T);
  uint64_t *bars = reinterpret_cast<uint64_t *>(rp_s + S * kTmaRpStride);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = epi.cta_tile0 ? __ldg(epi.cta_tile0 + blockIdx.x) : (int)(((int64_t)n_tiles * blockIdx.x) / gridDim.x);
  const int t1 = epi.cta_tile0 ? __ldg(epi.cta_tile0 + blockIdx.x + 1) : (int)(((int64_t)n_tiles * (blockIdx.x + 1)) / gridDim.x);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 33); mbar_init(empty0 + 8 * s, NC / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  /* The matrix (vals, colidx, rowptr, tile_row) is constant during a solve: the producer warp starts streaming the
   * first stages BEFORE griddepcontrol.wait, i.e. while the predecessor kernel is still draining; only x, the
   * scalars and everything written below depend on the predecessor. */
  const bool is_producer = (warp == NC / 32);
  const bool is_reducer = (warp == NC / 32 + 1); /* only launched (block size NC + 64) by the pipelined recurrence */
  int issued = 0;
  const uint64_t pol_last = l2_policy_evict_last(), pol_first = l2_policy_evict_first();
  auto issue_tile = [&](int t) {
    const int i = t - t0, s = i % S;
    const int rs = __ldg(tile_row + t), re = __ldg(tile_row + t + 1);
    if (lane == 0) {
      mbar_expect_tx(full0 + 8 * s, T * (8 + (int)sizeof(IDX)));
      if (epi.pin_tiles != 0) { /* > 0: that many leading tiles evict_last; < 0: every tile evict_first */
        const uint64_t pol = (i < epi.pin_tiles) ? pol_last : pol_first;
        tma_load_1d_hint(smem_u32(vals_s + s * T), vals + (size_t)t * T, T * 8, full0 + 8 * s, pol);
        tma_load_1d_hint(smem_u32(cols_s + s * T), colidx + (size_t)t * T, T * (int)sizeof(IDX), full0 + 8 * s, pol);
      } else {
        tma_load_1d(smem_u32(vals_s + s * T), vals + (size_t)t * T, T * 8, full0 + 8 * s);
        tma_load_1d(smem_u32(cols_s + s * T), colidx + (size_t)t * T, T * (int)sizeof(IDX), full0 + 8 * s);
      }
    }
    int32_t *rp = rp_s + s * kTmaRpStride;
    for (int j = lane; j <= re - rs; j += 32) cp_async4(smem_u32(rp + j), rowptr + rs + j);
    if (lane < 2) cp_async4(smem_u32(rp + kSpmvRowCap + 2 + lane), tile_row + t + lane);
    cp_async_mbar_arrive_noinc(full0 + 8 * s);
  };
#ifndef OHP_AB_NOPREFETCH
  if (is_producer) {
    for (; issued < S && t0 + issued < t1; ++issued) issue_tile(t0 + issued); /* stages are free at kernel start */
  }
#endif
  pdl_wait();
  pdl_launch();
  if (CG && st->done) {
    /* bulk copies already in flight must land before the CTA's shared memory is released */
    if (is_producer) for (int i = 0; i < issued; ++i) mbar_wait(full0 + 8 * (i % S), 0);
    return;
  }
  if (is_reducer) {
    if (CG && epi.pipe == 1 && blockIdx.x == 0) pipe_reduce(st, mode, pd, epi, lane);
    return;
  }
  const int its_now = CG ? st->its : 0;
  if (CG) x = cg_pick_input(st, epi, x, x_alt);
  if (CG && blockIdx.x == 0 && tid == 0) trace_put(epi.trace, TR_SPMV_BEGIN, its_now);
  /* halo/compute overlap: the vector kernel only PUSHED its halo; the neighbours' values are awaited here, and
   * only by the CTAs whose tiles reference ghost columns (the strips' first and last CTAs): interior CTAs start at once */
#ifdef OHP_AB_NOCOHERENT
  const bool ghost_cta = false;
#else
  const bool ghost_cta = CG && mode == OHP_MODE_P2P && epi.ghost_cta && epi.ghost_cta[blockIdx.x];
#endif
  double dot[1] = {0.0};
  if (is_producer) {
    /* ---------------- producer warp ---------------- */
    for (int t = t0 + issued; t < t1; ++t) {
      const int i = t - t0, s = i % S;
      mbar_wait(empty0 + 8 * s, ((i / S) & 1) ^ 1);
      issue_tile(t);
    }
  } else {
    /* ---------------- consumer warps ---------------- */
    if (ghost_cta) {
      if (tid < pd.n_nbr) p2p_wait(pd.hflags[pd.rank] + pd.nbr[tid], pd.seq[1], pd.seq + 2, pd.timeout_ns);
      asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
      if (tid == 0) trace_put(epi.trace, TR_SPMV_HALO_READY, its_now);
      dot[0] = spmv_consume<CG, S, NC, IDX, SPLIT, true>(rowptr, colidx, vals, x, y, vals_s, cols_s, rp_s, full0, empty0, t0, t1);
    } else {
      dot[0] = spmv_consume<CG, S, NC, IDX, SPLIT, false>(rowptr, colidx, vals, x, y, vals_s, cols_s, rp_s, full0, empty0, t0, t1);
    }
    if (CG && tid == 0) trace_put(epi.trace, TR_SPMV_CTA_DONE, its_now);
  }
  if (CG && !epi.pipe) {
    if (grid_sum<1>(dot, partials, ticket)) cg_spmv_epilogue(st, dot, sums, mode, pd, epi, its_now);
  }
}


/* ------------------------------------------------------------------------------------ */
/* The same SpMV with DYNAMIC tile scheduling (pipelined CG and plain MatMult: no dot      */
/* product here, so which CTA sums which rows cannot change a rounding).  A device-side     */
/* timeline of the static kernel at 12 tiles per CTA showed CTAs finishing between 14.0 and   */
/* 17.3 us with the SAME CTAs slow in every launch (die position, L2 distance): the launch     */
/* ends with its slowest SM.  Here every CTA starts on a static run of tiles (so that the       */
/* producer can prefetch before griddepcontrol.wait, as before) worth ~70 % of its share; the    */
/* rest is handed out in chunks of `chunk` tiles from a counter in HBM that the PRECEDING        */
/* kernel re-arms.  The producer writes the tile number next to the staged rowptr slice; a        */
/* negative number ends the CTA.  A row that runs past the last tile of a run is finished from    */
/* global memory by the thread that carried it (serial CSR order as always).  Halo: a tile that   */
/* references ghost columns (table built at set-up) makes its CTA wait for the neighbours' flags   */
/* once and read x coherently for that tile.                                                    */
/* ------------------------------------------------------------------------------------ */
struct ohp_dyn_sched { unsigned *counter; int static_tiles, chunk; const uint8_t *ghost_tile; };

template <class IDX, bool COHERENT, int NC>
__device__ __forceinline__ void spmv_dyn_tile(const int32_t *rp, const double *vs, const IDX *cs, const double *x, double *__restrict__ y,
                                              int t, double &carry, int &pend) {
  constexpr int T = kSpmvTile;
  const int tid = threadIdx.x;
  const int rs = rp[kSpmvRowCap + 2], re = rp[kSpmvRowCap + 3];
  const int base = t * T, lim = base + T;
  if (pend >= 0) {
    const int end = rp[0];
    carry = row_dot<IDX, COHERENT>(vs, cs, 0, min(end, lim) - base, x, carry, pend);
    if (end <= lim) { y[pend] = carry; pend = -1; }
  }
  for (int r = rs + tid; r < re; r += NC) {
    const int a = rp[r - rs], b = rp[r - rs + 1];
    const double sum = row_dot<IDX, COHERENT>(vs, cs, a - base, min(b, lim) - base, x, 0.0, r);
    if (b > lim) { carry = sum; pend = r; }
    else y[r] = sum;
  }
}

template <int S, int NC, class IDX>
__global__ void __launch_bounds__(NC + 64, 1)
k_spmv_dyn(const int32_t *__restrict__ rowptr, const IDX *__restrict__ colidx, const double *__restrict__ vals, const double *x,
           double *__restrict__ y, const int32_t *__restrict__ tile_row, int n_tiles, ohp_cg_state *st, int mode,
           const ohp_p2p_dev pd, const ohp_cgcg_epi epi, const ohp_dyn_sched ds) {
  constexpr int T = kSpmvTile;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *vals_s = reinterpret_cast<double *>(smem_raw);
  IDX *cols_s = reinterpret_cast<IDX *>(vals_s + S * T);
  int32_t *rp_s = reinterpret_cast<int32_t *>(cols_s + S * T);
  uint64_t *bars = reinterpret_cast<uint64_t *>(rp_s + S * kTmaRpStride);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 33); mbar_init(empty0 + 8 * s, NC / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const bool is_producer = (warp == NC / 32), is_reducer = (warp == NC / 32 + 1);
  const uint64_t pol_first = l2_policy_evict_first();
  /* stage i % S of this CTA's i-th tile; t < 0 ends the CTA */
  auto issue = [&](int i, int t) {
    const int s = i % S;
    int32_t *rp = rp_s + s * kTmaRpStride;
    if (t < 0) {
      if (lane == 0) { rp[kSpmvRowCap] = -1; mbar_arrive(full0 + 8 * s); }
      cp_async_mbar_arrive_noinc(full0 + 8 * s);
      return;
    }
    const int rs = __ldg(tile_row + t), re = __ldg(tile_row + t + 1);
    if (lane == 0) {
      rp[kSpmvRowCap] = t;
      mbar_expect_tx(full0 + 8 * s, T * (8 + (int)sizeof(IDX)));
      tma_load_1d_hint(smem_u32(vals_s + s * T), vals + (size_t)t * T, T * 8, full0 + 8 * s, pol_first);
      tma_load_1d_hint(smem_u32(cols_s + s * T), colidx + (size_t)t * T, T * (int)sizeof(IDX), full0 + 8 * s, pol_first);
    }
    for (int j = lane; j <= re - rs; j += 32) cp_async4(smem_u32(rp + j), rowptr + rs + j);
    if (lane < 2) cp_async4(smem_u32(rp + kSpmvRowCap + 2 + lane), tile_row + t + lane);
    cp_async_mbar_arrive_noinc(full0 + 8 * s);
  };
  /* static run of this CTA: [ts0, ts1) */
  const int ts0 = min(n_tiles, (int)blockIdx.x * ds.static_tiles), ts1 = min(n_tiles, ts0 + ds.static_tiles);
  int issued = 0;
  if (is_producer) {
    for (; issued < S && ts0 + issued < ts1; ++issued) issue(issued, ts0 + issued); /* stages are free at kernel start */
  }
  pdl_wait();
  pdl_launch();
  if (st && st->done) {
    if (is_producer) for (int i = 0; i < issued; ++i) mbar_wait(full0 + 8 * (i % S), 0);
    return;
  }
  if (is_reducer) {
    if (st && epi.pipe == 1 && blockIdx.x == 0) pipe_reduce(st, mode, pd, epi, lane);
    return;
  }
  if (is_producer) {
    int i = issued;
    for (int t = ts0 + issued; t < ts1; ++t, ++i) {
      mbar_wait(empty0 + 8 * (i % S), ((i / S) & 1) ^ 1);
      issue(i, t);
    }
    for (;;) {
      int c0 = 0;
      if (lane == 0) c0 = (int)atomicAdd(ds.counter, (unsigned)ds.chunk);
      c0 = __shfl_sync(0xffffffffu, c0, 0);
      if (c0 >= n_tiles) break;
      for (int t = c0; t < min(n_tiles, c0 + ds.chunk); ++t, ++i) {
        mbar_wait(empty0 + 8 * (i % S), ((i / S) & 1) ^ 1);
        issue(i, t);
      }
    }
    mbar_wait(empty0 + 8 * (i % S), ((i / S) & 1) ^ 1);
    issue(i, -1);
    return;
  }
  /* ---------------- consumer warps ---------------- */
  double carry = 0.0;
  int pend = -1, prev = -2;
  bool halo_ready = false;
  for (int i = 0;; ++i) {
    const int s = i % S;
    mbar_wait(full0 + 8 * s, (i / S) & 1);
    const int32_t *rp = rp_s + s * kTmaRpStride;
    const int t = rp[kSpmvRowCap];
    if (pend >= 0 && t != prev + 1) { /* the run ended: the carried row's tail is not coming through shared memory */
      const int end = __ldg(rowptr + pend + 1);
      carry = row_dot<IDX, true>(vals + (size_t)(prev + 1) * T, colidx + (size_t)(prev + 1) * T, 0, end - (prev + 1) * T, x, carry, pend);
      y[pend] = carry;
      pend = -1;
    }
    if (t < 0) break;
    const bool ghost = mode == OHP_MODE_P2P && ds.ghost_tile && ds.ghost_tile[t];
    if (ghost && !halo_ready) { /* uniform over the consumer warps: the tile number is */
      if (tid < pd.n_nbr) p2p_wait(pd.hflags[pd.rank] + pd.nbr[tid], pd.seq[1], pd.seq + 2, pd.timeout_ns);
      asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
      halo_ready = true;
    }
    if (ghost) spmv_dyn_tile<IDX, true, NC>(rp, vals_s + s * T, cols_s + s * T, x, y, t, carry, pend);
    else spmv_dyn_tile<IDX, false, NC>(rp, vals_s + s * T, cols_s + s * T, x, y, t, carry, pend);
    prev = t;
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * s);
  }
}

constexpr int kPersistentBelowRows = 400000;
constexpr int kSingleReductionBelowRows = 1600000;
constexpr int kPipelinedBelowRows = 1600000;
constexpr int kPipelinedAboveRows = 300000;
static bool spmv_use_tma(ohp_mesh m) { return m->max_tile_rows <= kSpmvRowCap - 1 && m->n_tiles > 0; }

static int spmv_smem_bytes(ohp_mesh m) { return (int)sizeof(double) * (kSpmvTile + std::max(1, m->max_row)); }

static int spmv_configure(ohp_mesh m) {
  const int bytes = spmv_smem_bytes(m);
  if (bytes > 200 * 1024) OHP_FAIL(OHP_ERR_SUP, "SpMV: longest row %d exceeds the shared-memory tile", m->max_row);
  static int configured_bytes[64] = {0}; /* per device id */
  if (bytes > configured_bytes[m->ctx->device & 63]) {
    OHP_CUDA(cudaFuncSetAttribute(k_spmv_stream<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    OHP_CUDA(cudaFuncSetAttribute(k_spmv_stream<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    configured_bytes[m->ctx->device & 63] = bytes;
  }
  return 0;
}

/* L2 policy of the matrix stream.  The matrix is read exactly once per SpMV and is larger than what the L2 can keep next
 * to the vectors, so by default EVERY tile is loaded with evict_first: the stream passes through without displacing the
 * CG vectors, which are re-read within the iteration.  Measured at 1.05 M rows per GPU (the 8-GPU share of the 16 M-triangle
 * box, working set ~ L2 size): 37.4 -> 32.6 us per iteration (two-launch single-reduction loop), 38.9 -> 33.0 (pipelined);
 * at 2.1 M rows 67.2 -> 63.0.  Keeping a share of the tiles with evict_last instead (OHP_L2PIN_MB=<megabytes>) was never
 * better than keeping none (16 MB: 32.8, 32 MB: 34.3, 48 MB: 35.4, 64 MB: 37.0 us): resident matrix lines cost more in
 * evicted vector lines than they save.  OHP_L2PIN_MB=off loads without hints (A/B runs).
 * Return value: number of leading tiles per CTA loaded evict_last, the rest evict_first; 0 = no hints. */
static int spmv_pin_tiles(ohp_mesh m, int grid, int idx_bytes) {
  static const char *env = getenv("OHP_L2PIN_MB");
  if (env && (env[0] == 'o' || env[0] == 'O')) return 0;
  const double mb = env ? atof(env) : 0.0;
  if (grid <= 0) return 0;
  if (!(mb > 0.0)) return -1; /* all tiles evict_first */
  const double tile_bytes = (double)kSpmvTile * (8 + idx_bytes);
  const int64_t total = (int64_t)(mb * 1e6 / tile_bytes);
  const int per_cta = (int)std::min<int64_t>(total / grid, (int64_t)(m->n_tiles + grid - 1) / grid);
  return std::max(per_cta, 1);
}

template <int S, int NC, class IDX, bool SPLIT, int MINB = 1>
static int spmv_tma_launch_ts(ohp_mat A, const IDX *cols, const double *x, double *y, ohp_cg_state *st, double *sums, int mode,
                              const double *x_alt, const ohp_p2p_dev &pd, const ohp_cgcg_epi &epi_in) {
  ohp_mesh m = A->mesh;
  ohp_context ctx = m->ctx;
  constexpr int smem = tma_smem_bytes(S, (int)sizeof(IDX));
  /* once per DEVICE, not per process: the attribute belongs to the context of the device (one bit per device id) */
  static unsigned long long configured_devices = 0ull;
  if (!((configured_devices >> (ctx->device & 63)) & 1ull)) {
    OHP_CUDA(cudaFuncSetAttribute(k_spmv_tma<true, S, NC, IDX, SPLIT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    OHP_CUDA(cudaFuncSetAttribute(k_spmv_tma<false, S, NC, IDX, SPLIT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_devices |= 1ull << (ctx->device & 63);
  }
  const int grid = std::max(1, std::min(m->n_tiles, ctx->num_sms * MINB));
  ohp_cgcg_epi epi = epi_in;
  epi.pin_tiles = spmv_pin_tiles(m, grid, (int)sizeof(IDX));
  if (st) {
    static const char *nopdl = getenv("OHP_NOPDL");
    OHP_CUDA(ohp_launch(k_spmv_tma<true, S, NC, IDX, SPLIT, MINB>, dim3(grid), dim3(NC + (epi.pipe ? 64 : 32)), (size_t)smem, ctx->stream, mode != OHP_MODE_NCCL && !nopdl,
                        m->rowptr, cols, A->vals, x, y, m->tile_row, m->n_tiles, st, ctx->d_partials, ctx->d_ticket + 1, sums, mode, x_alt, pd, epi));
  }
  else k_spmv_tma<false, S, NC, IDX, SPLIT, MINB><<<grid, NC + 32, smem, ctx->stream>>>(
      m->rowptr, cols, A->vals, x, y, m->tile_row, m->n_tiles, nullptr, nullptr, nullptr, nullptr, 0, nullptr, pd, epi);
  OHP_CHECK_LAUNCH(ctx);
  return 0;
}
/* long rows (more than 8 non-zeros on average: 3-D meshes) share a row among 2-4 lanes */
template <int S, int NC, class IDX>
static int spmv_tma_launch_t(ohp_mat A, const IDX *cols, const double *x, double *y, ohp_cg_state *st, double *sums, int mode,
                             const double *x_alt, const ohp_p2p_dev &pd, const ohp_cgcg_epi &epi) {
  ohp_mesh m = A->mesh;
  if (m->nnz > (int64_t)8 * std::max(1, m->n_owned)) return spmv_tma_launch_ts<S, NC, IDX, true>(A, cols, x, y, st, sums, mode, x_alt, pd, epi);
  return spmv_tma_launch_ts<S, NC, IDX, false>(A, cols, x, y, st, sums, mode, x_alt, pd, epi);
}
template <int S, int NC>
static int spmv_tma_launch(ohp_mat A, const double *x, double *y, ohp_cg_state *st, double *sums, int mode,
                           const double *x_alt, const ohp_p2p_dev &pd, const ohp_cgcg_epi &epi) {
  ohp_mesh m = A->mesh;
  static const char *idx = getenv("OHP_IDX"); /* "32" keeps absolute 32-bit column ids (A/B runs) */
  /* Long rows (3-D meshes, 15 non-zeros per row: ~273 rows per tile).  The ncu capture (profiles/r02_spmv_tma_3d_ncu.txt)
   * shows this case bound by instruction issue and gather latency, not DRAM (42 % of peak, 1.46 warp-instructions per
   * non-zero against 1.10 in 2-D): sharing a row among 2-4 lanes of 19 warps costs more instructions (predicated batches,
   * shuffles) than its parallelism gains.  One thread per row with 14 consumer warps is faster: 53.3 vs 61.5 us per launch
   * on the 7.99 M-tet box (0.73 vs 0.64 of the HBM roofline; 384 consumers 54.2, 320: 63.0, lane sharing with 512: 58.6). */
  static const char *force_cfg = getenv("OHP_SPMV");
  if (NC == 608 && S == 3 && !force_cfg && m->nnz > (int64_t)8 * std::max(1, m->n_owned) && m->col16 && !(idx && idx[0] == '3'))
    return spmv_tma_launch_ts<3, 448, int16_t, false>(A, m->col16, x, y, st, sums, mode, x_alt, pd, epi);
  if (m->col16 && !(idx && idx[0] == '3')) return spmv_tma_launch_t<S, NC, int16_t>(A, m->col16, x, y, st, sums, mode, x_alt, pd, epi);
  return spmv_tma_launch_t<S, NC, int32_t>(A, m->colidx, x, y, st, sums, mode, x_alt, pd, epi);
}


/* dynamic schedule of k_spmv_dyn: ~70 % of every CTA's share as a static run, the rest in chunks from the counter */
static void dyn_sched_params(ohp_mesh m, int num_sms, int *G, int *static_tiles, int *chunk) {
  *G = std::max(1, std::min(m->n_tiles, num_sms));
  static const char *fr = getenv("OHP_DYN_STATIC"); /* static share in percent (A/B runs) */
  const double share = fr ? atof(fr) / 100.0 : 0.7;
  *static_tiles = (int)(share * (double)m->n_tiles / (double)*G);
  const int rest = m->n_tiles - *static_tiles * *G;
  *chunk = std::max(1, std::min(4, rest / (*G * 4)));
}
template <int S, int NC, class IDX>
static int spmv_dyn_launch_t(ohp_mat A, const IDX *cols, const double *x, double *y, ohp_cg_state *st, int mode, const ohp_p2p_dev &pd,
                             const ohp_cgcg_epi &epi) {
  ohp_mesh m = A->mesh;
  ohp_context ctx = m->ctx;
  constexpr int smem = tma_smem_bytes(S, (int)sizeof(IDX));
  static unsigned long long configured_devices = 0ull;
  if (!((configured_devices >> (ctx->device & 63)) & 1ull)) {
    OHP_CUDA(cudaFuncSetAttribute(k_spmv_dyn<S, NC, IDX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_devices |= 1ull << (ctx->device & 63);
  }
  ohp_dyn_sched ds;
  int G;
  dyn_sched_params(m, ctx->num_sms, &G, &ds.static_tiles, &ds.chunk);
  ds.counter = ctx->d_ticket + 3;
  ds.ghost_tile = (mode == OHP_MODE_P2P && epi.ghost_cta) ? m->ghost_tile : nullptr;
  static const char *nopdl = getenv("OHP_NOPDL");
  if (!st) { /* plain MatMult: nobody armed the counter -> everything comes from it */
    OHP_CUDA(cudaMemsetAsync(ds.counter, 0, sizeof(unsigned), ctx->stream));
    ds.static_tiles = 0;
  }
  OHP_CUDA(ohp_launch(k_spmv_dyn<S, NC, IDX>, dim3(G), dim3(NC + 64), (size_t)smem, ctx->stream, st != nullptr && mode != OHP_MODE_NCCL && !nopdl,
                      (const int32_t *)m->rowptr, cols, (const double *)A->vals, x, y, (const int32_t *)m->tile_row, m->n_tiles, st, mode, pd, epi, ds));
  OHP_CHECK_LAUNCH(ctx);
  return 0;
}
static int spmv_dyn_launch(ohp_mat A, const double *x, double *y, ohp_cg_state *st, int mode, const ohp_p2p_dev &pd, const ohp_cgcg_epi &epi) {
  ohp_mesh m = A->mesh;
  static const char *idx = getenv("OHP_IDX");
  if (m->col16 && !(idx && idx[0] == '3')) return spmv_dyn_launch_t<3, 608, int16_t>(A, m->col16, x, y, st, mode, pd, epi);
  return spmv_dyn_launch_t<3, 608, int32_t>(A, m->colidx, x, y, st, mode, pd, epi);
}

/* y = A x on the context stream; with `st` the (x, y) dot and the CG scalar update are fused in */
static int spmv_enqueue(ohp_mat A, const double *x, double *y, ohp_cg_state *st, double *sums, int mode,
                        const double *x_alt, const ohp_p2p_dev &pd, const ohp_cgcg_epi &epi = ohp_cgcg_epi{}) {
  ohp_mesh m = A->mesh;
  ohp_context ctx = m->ctx;
  static const char *force = getenv("OHP_SPMV"); /* "stream" forces the fallback kernel; "0".."5" pick a TMA config (A/B runs) */
  /* a rank that owns no rows still runs the pipelined solver's reducer warp, which lives in the TMA kernel's CTA 0 */
  const bool tma = (spmv_use_tma(m) || (epi.pipe && m->n_tiles == 0)) && !(force && force[0] == 's');
  /* OHP_DYN=1: dynamically scheduled tiles in the pipelined solver.  Measured SLOWER than the static partition (1.05 M rows:
   * 36.2 vs 32.5 us per iteration with a 70 % static share, 34.8 with 85 %): the scattered tail of tiles and the counter
   * round trips cost more than the 14.0-17.3 us spread between CTAs that it removes.  Kept for A/B runs. */
  static const char *dyn = getenv("OHP_DYN");
  if (tma && epi.pipe && dyn && dyn[0] == '1' && !(force && force[0] >= '0' && force[0] <= '9')) return spmv_dyn_launch(A, x, y, st, mode, pd, epi);
  if (tma) {
    const int cfg = (force && force[0] >= '0' && force[0] <= '9') ? force[0] - '0' : 0;
    switch (cfg) {
      /* 4, 5: two CTAs per SM with two stages -- measured slower (0.151 / 0.177 ms vs 0.146 on the 16 M box), kept for A/B */
      case 4: if (m->col16 && ctx->nranks == 1) return spmv_tma_launch_ts<2, 352, int16_t, false, 2>(A, m->col16, x, y, st, sums, mode, x_alt, pd, epi); break;
      case 5: if (m->col16 && ctx->nranks == 1) return spmv_tma_launch_ts<2, 288, int16_t, false, 2>(A, m->col16, x, y, st, sums, mode, x_alt, pd, epi); break;
      /* 6-8: one thread per row whatever the row length (no lane sharing), fewer consumer warps (long-row matrices, A/B) */
      case 6: if (m->col16) return spmv_tma_launch_ts<3, 384, int16_t, false>(A, m->col16, x, y, st, sums, mode, x_alt, pd, epi); break;
      case 7: if (m->col16) return spmv_tma_launch_ts<3, 320, int16_t, false>(A, m->col16, x, y, st, sums, mode, x_alt, pd, epi); break;
      case 8: if (m->col16) return spmv_tma_launch_ts<3, 448, int16_t, false>(A, m->col16, x, y, st, sums, mode, x_alt, pd, epi); break;
      case 9: if (m->col16) return spmv_tma_launch_ts<3, 512, int16_t, true>(A, m->col16, x, y, st, sums, mode, x_alt, pd, epi); break;
      case 1: return spmv_tma_launch<3, 384>(A, x, y, st, sums, mode, x_alt, pd, epi);
      case 2: return spmv_tma_launch<3, 768>(A, x, y, st, sums, mode, x_alt, pd, epi);
      case 3: return spmv_tma_launch<2, 640>(A, x, y, st, sums, mode, x_alt, pd, epi);
      default: return spmv_tma_launch<3, 608>(A, x, y, st, sums, mode, x_alt, pd, epi);
    }
  }
  OHP_TRY(spmv_configure(m));
  const int grid = std::max(1, std::min(m->n_tiles, kMaxPartials));
  if (st) k_spmv_stream<true><<<grid, kSpmvThreads, spmv_smem_bytes(m), ctx->stream>>>(
      m->rowptr, m->colidx, A->vals, x, y, m->tile_row, m->n_tiles, st, ctx->d_partials, ctx->d_ticket + 1, sums, mode, x_alt, pd, epi);
  else k_spmv_stream<false><<<grid, kSpmvThreads, spmv_smem_bytes(m), ctx->stream>>>(
      m->rowptr, m->colidx, A->vals, x, y, m->tile_row, m->n_tiles, nullptr, nullptr, nullptr, nullptr, 0, nullptr, pd, epi);
  OHP_CHECK_LAUNCH(ctx);
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* SpMV with the Chebyshev update in its consumer loop (general preconditioned CG,          */
/* ohp_pcg.cu): y_{k+1}[row] = a y_{k-1}[row] + b y_k[row] + c D^-1[row] (r[row] - (A y_k)[row])     */
/* while (A y_k)[row] is still in a register.  Against SpMV + k_pcg_cheb_step this saves the    */
/* store and the re-load of A y_k (16 of 56 bytes per row next to the matrix stream) and one     */
/* kernel boundary per application; the row's four vector operands are loaded before the row is  */
/* summed, so their latency hides behind the gather (the emit.load / emit.store protocol of the   */
/* fused pipelined-CG kernel).  Same producer warp, stages, tiles and summation order as           */
/* k_spmv_tma; the per-row arithmetic is the expression of k_pcg_cheb_step.                        */
/* ------------------------------------------------------------------------------------ */
struct ChebRowUpdate {
  const double *ykm1, *yk, *r, *dinv;
  double *ynew;
  double a, b, c;
  struct Opnd { double km1, k, r, d; };
  __device__ __forceinline__ Opnd load(int i) const {
    Opnd o;
    o.km1 = ykm1 ? ykm1[i] : 0.0; o.k = yk[i]; o.r = r[i]; o.d = dinv[i];
    return o;
  }
  __device__ __forceinline__ void store(int i, double n, const Opnd &o) const {
    ynew[i] = (ykm1 ? a * o.km1 : 0.0) + b * o.k + c * (o.d * (o.r - n));
  }
  __device__ __forceinline__ void operator()(int i, double n) const { store(i, n, load(i)); }
};

template <int S, int NC, class IDX, bool SPLIT>
__global__ void __launch_bounds__(NC + 32, 1)
k_spmv_cheb(const int32_t *__restrict__ rowptr, const IDX *__restrict__ colidx, const double *__restrict__ vals,
            const int32_t *__restrict__ tile_row, int n_tiles, const double *x, ChebRowUpdate up, int pin_tiles) {
  constexpr int T = kSpmvTile;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *vals_s = reinterpret_cast<double *>(smem_raw);
  IDX *cols_s = reinterpret_cast<IDX *>(vals_s + S * T);
  int32_t *rp_s = reinterpret_cast<int32_t *>(cols_s + S * T);
  uint64_t *bars = reinterpret_cast<uint64_t *>(rp_s + S * kTmaRpStride);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = (int)(((int64_t)n_tiles * blockIdx.x) / gridDim.x);
  const int t1 = (int)(((int64_t)n_tiles * (blockIdx.x + 1)) / gridDim.x);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 33); mbar_init(empty0 + 8 * s, NC / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == NC / 32) {
    /* ---------------- producer warp ---------------- */
    const uint64_t pol_last = l2_policy_evict_last(), pol_first = l2_policy_evict_first();
    for (int t = t0; t < t1; ++t) {
      const int i = t - t0, s = i % S;
      if (i >= S) mbar_wait(empty0 + 8 * s, ((i / S) & 1) ^ 1); /* the first S stages are free at kernel start */
      const int rs = __ldg(tile_row + t), re = __ldg(tile_row + t + 1);
      if (lane == 0) {
        mbar_expect_tx(full0 + 8 * s, T * (8 + (int)sizeof(IDX)));
        if (pin_tiles != 0) {
          const uint64_t pol = (i < pin_tiles) ? pol_last : pol_first;
          tma_load_1d_hint(smem_u32(vals_s + s * T), vals + (size_t)t * T, T * 8, full0 + 8 * s, pol);
          tma_load_1d_hint(smem_u32(cols_s + s * T), colidx + (size_t)t * T, T * (int)sizeof(IDX), full0 + 8 * s, pol);
        } else {
          tma_load_1d(smem_u32(vals_s + s * T), vals + (size_t)t * T, T * 8, full0 + 8 * s);
          tma_load_1d(smem_u32(cols_s + s * T), colidx + (size_t)t * T, T * (int)sizeof(IDX), full0 + 8 * s);
        }
      }
      int32_t *rp = rp_s + s * kTmaRpStride;
      for (int j = lane; j <= re - rs; j += 32) cp_async4(smem_u32(rp + j), rowptr + rs + j);
      if (lane < 2) cp_async4(smem_u32(rp + kSpmvRowCap + 2 + lane), tile_row + t + lane);
      cp_async_mbar_arrive_noinc(full0 + 8 * s);
    }
    return;
  }
  /* ---------------- consumer warps ---------------- */
  spmv_consume_emit<S, NC, IDX, SPLIT, false>(rowptr, colidx, vals, x, vals_s, cols_s, rp_s, full0, empty0, t0, t1, up);
}

template <int S, int NC, class IDX, bool SPLIT>
static int spmv_cheb_launch_t(ohp_mat A, const IDX *cols, const double *x, const ChebRowUpdate &up) {
  ohp_mesh m = A->mesh;
  ohp_context ctx = m->ctx;
  constexpr int smem = tma_smem_bytes(S, (int)sizeof(IDX));
  static unsigned long long configured_devices = 0ull;
  if (!((configured_devices >> (ctx->device & 63)) & 1ull)) {
    OHP_CUDA(cudaFuncSetAttribute(k_spmv_cheb<S, NC, IDX, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_devices |= 1ull << (ctx->device & 63);
  }
  const int grid = std::max(1, std::min(m->n_tiles, ctx->num_sms));
  k_spmv_cheb<S, NC, IDX, SPLIT><<<grid, NC + 32, smem, ctx->stream>>>(m->rowptr, cols, A->vals, m->tile_row, m->n_tiles, x, up,
                                                                     spmv_pin_tiles(m, grid, (int)sizeof(IDX)));
  OHP_CHECK_LAUNCH(ctx);
  return 0;
}

/* ynew = a ykm1 + b yk + c D^-1 (r - A yk) in ONE launch when the matrix fits the TMA-pipelined SpMV (*fused = 1); otherwise
 * *fused = 0 and nothing was launched: the caller runs ohp_spmv_launch + the element-wise update.  ykm1 may be NULL (taken
 * as zero).  yk's ghost tail must be current.  OHP_CHEB_FUSED=0 switches the fused kernel off, =1 also uses it for long-row
 * matrices (A/B runs). */
int ohp_spmv_cheb_launch(ohp_mat A, const double *yk, const double *ykm1, const double *r, const double *dinv, double *ynew,
                         double a, double b, double c, int *fused) {
  ohp_mesh m = A->mesh;
  static const char *off = getenv("OHP_CHEB_FUSED");
  static const char *force = getenv("OHP_SPMV");
  static const char *idx = getenv("OHP_IDX");
  *fused = 0;
  if ((off && off[0] == '0') || force || !spmv_use_tma(m)) return 0;
  /* Measured on B200 (profiles/r02_chebyshev_time_to_solution*.json, k = 16): 16 M-triangle box 1766 -> 1591 ms per solve
   * with the fused kernel (plain CG: 1717); the 8 M-tet box (15 non-zeros per row, 1.3 M rows, latency-bound) 29.1 -> 33.2 ms:
   * there one thread per row already waits on the gather and four more operand loads per row lengthen the chain.  So
   * long-row matrices keep SpMV + element-wise update unless OHP_CHEB_FUSED=1 asks for the fused kernel (A/B runs). */
  const bool long_rows = m->nnz > (int64_t)8 * std::max(1, m->n_owned);
  if (long_rows && !(off && off[0] == '1')) return 0;
  *fused = 1;
  if (m->n_owned == 0) return 0;
  ChebRowUpdate up;
  up.ykm1 = ykm1; up.yk = yk; up.r = r; up.dinv = dinv; up.ynew = ynew; up.a = a; up.b = b; up.c = c;
  const bool i16 = m->col16 && !(idx && idx[0] == '3');
  /* long rows: one thread per row, 14 consumer warps (see spmv_tma_launch) */
  if (i16) return long_rows ? spmv_cheb_launch_t<3, 448, int16_t, false>(A, m->col16, yk, up) : spmv_cheb_launch_t<3, 608, int16_t, false>(A, m->col16, yk, up);
  return long_rows ? spmv_cheb_launch_t<3, 448, int32_t, false>(A, m->colidx, yk, up) : spmv_cheb_launch_t<3, 608, int32_t, false>(A, m->colidx, yk, up);
}

int ohp_spmv_launch(ohp_mat A, const double *x, double *y) {
  if (A->mesh->n_owned == 0) return 0;
  static const ohp_p2p_dev none = {};
  return spmv_enqueue(A, x, y, nullptr, nullptr, 0, nullptr, none);
}

/* ------------------------------------------------------------------------------------ */
/* Vec                                                                                   */
/* ------------------------------------------------------------------------------------ */
__global__ void k_set(double *x, int n, double a) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = a;
}
__global__ void k_axpy(double *y, const double *x, int n, double a) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] += a * x[i];
}
__global__ void k_aypx(double *y, const double *x, int n, double a) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = x[i] + a * y[i];
}
__global__ void k_chop(double *x, int n, double tol) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (fabs(x[i]) < tol) x[i] = 0.0;
}
__global__ void __launch_bounds__(256) k_dot(const double *x, const double *y, int n, double *partials,
                                             unsigned *ticket, double *out) {
  double s[1] = {0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s[0] += x[i] * y[i];
  if (grid_sum<1>(s, partials, ticket) && threadIdx.x == 0) out[0] = s[0];
}

static int vec_grid(ohp_context ctx, int n) {
  return (int)std::min<int64_t>(nblk(n, 256), (int64_t)ctx->num_sms * 8);
}

extern "C" int ohp_vec_create(ohp_mesh m, ohp_vec *out) {
  if (!m || !m->setup_done) OHP_FAIL(OHP_ERR_ORDER, "ohp_vec_create: mesh not set up");
  OHP_CUDA(cudaSetDevice(m->ctx->device));
  ohp_vec v = new ohp_vec_s();
  v->mesh = m; v->n = m->n_owned;
  const size_t len = (size_t)std::max(1, m->n_owned + m->n_ghost);
  OHP_CUDA(OHP_MALLOC(m->ctx, &v->d, sizeof(double) * len));
  OHP_CUDA(cudaMemsetAsync(v->d, 0, sizeof(double) * len, m->ctx->stream));
  *out = v;
  return 0;
}
extern "C" int ohp_vec_duplicate(ohp_vec v, ohp_vec *w) { return ohp_vec_create(v->mesh, w); }
extern "C" int ohp_vec_destroy(ohp_vec v) {
  if (!v) return 0;
  cudaSetDevice(v->mesh->ctx->device);
  cudaStreamSynchronize(v->mesh->ctx->stream);
  OHP_FREE(v->mesh->ctx, v->d);
  delete v;
  return 0;
}
extern "C" int ohp_vec_size(ohp_vec v, int32_t *n) { *n = v->n; return 0; }
extern "C" void *ohp_vec_device_ptr(ohp_vec v) { return v ? (void *)v->d : nullptr; }

extern "C" int ohp_vec_set(ohp_vec v, double a) {
  ohp_context ctx = v->mesh->ctx;
  OHP_CUDA(cudaSetDevice(ctx->device));
  k_set<<<vec_grid(ctx, v->n), 256, 0, ctx->stream>>>(v->d, v->n, a);
  OHP_CHECK_LAUNCH(ctx);
  return 0;
}
extern "C" int ohp_vec_copy(ohp_vec src, ohp_vec dst) {
  if (src->n != dst->n) OHP_FAIL(OHP_ERR_ARG_SIZ, "ohp_vec_copy: sizes %d vs %d", src->n, dst->n);
  OHP_CUDA(cudaSetDevice(src->mesh->ctx->device));
  OHP_CUDA(cudaMemcpyAsync(dst->d, src->d, sizeof(double) * src->n, cudaMemcpyDeviceToDevice, src->mesh->ctx->stream));
  return 0;
}
extern "C" int ohp_vec_from_host(ohp_vec v, const double *h) {
  OHP_CUDA(cudaSetDevice(v->mesh->ctx->device));
  OHP_CUDA(cudaMemcpyAsync(v->d, h, sizeof(double) * v->n, cudaMemcpyHostToDevice, v->mesh->ctx->stream));
  OHP_CUDA(cudaStreamSynchronize(v->mesh->ctx->stream));
  return 0;
}
extern "C" int ohp_vec_to_host(ohp_vec v, double *h) {
  OHP_CUDA(cudaSetDevice(v->mesh->ctx->device));
  OHP_CUDA(cudaMemcpyAsync(h, v->d, sizeof(double) * v->n, cudaMemcpyDeviceToHost, v->mesh->ctx->stream));
  OHP_CUDA(cudaStreamSynchronize(v->mesh->ctx->stream));
  return 0;
}
extern "C" int ohp_vec_axpy(ohp_vec y, double a, ohp_vec x) {
  if (x->n != y->n) OHP_FAIL(OHP_ERR_ARG_SIZ, "ohp_vec_axpy: sizes %d vs %d", x->n, y->n);
  ohp_context ctx = y->mesh->ctx;
  OHP_CUDA(cudaSetDevice(ctx->device));
  k_axpy<<<vec_grid(ctx, y->n), 256, 0, ctx->stream>>>(y->d, x->d, y->n, a);
  OHP_CHECK_LAUNCH(ctx);
  return 0;
}
extern "C" int ohp_vec_aypx(ohp_vec y, double a, ohp_vec x) {
  if (x->n != y->n) OHP_FAIL(OHP_ERR_ARG_SIZ, "ohp_vec_aypx: sizes %d vs %d", x->n, y->n);
  ohp_context ctx = y->mesh->ctx;
  OHP_CUDA(cudaSetDevice(ctx->device));
  k_aypx<<<vec_grid(ctx, y->n), 256, 0, ctx->stream>>>(y->d, x->d, y->n, a);
  OHP_CHECK_LAUNCH(ctx);
  return 0;
}
extern "C" int ohp_vec_chop(ohp_vec v, double tol) {
  ohp_context ctx = v->mesh->ctx;
  OHP_CUDA(cudaSetDevice(ctx->device));
  k_chop<<<vec_grid(ctx, v->n), 256, 0, ctx->stream>>>(v->d, v->n, tol);
  OHP_CHECK_LAUNCH(ctx);
  return 0;
}
extern "C" int ohp_vec_dot(ohp_vec x, ohp_vec y, double *result) {
  if (x->n != y->n) OHP_FAIL(OHP_ERR_ARG_SIZ, "ohp_vec_dot: sizes %d vs %d", x->n, y->n);
  ohp_context ctx = x->mesh->ctx;
  OHP_CUDA(cudaSetDevice(ctx->device));
  k_dot<<<vec_grid(ctx, x->n), 256, 0, ctx->stream>>>(x->d, y->d, x->n, ctx->d_partials, ctx->d_ticket, ctx->d_scalars);
  OHP_CHECK_LAUNCH(ctx);
  if (ctx->nranks > 1) OHP_TRY(ohp_allreduce_sum(ctx, ctx->d_scalars, 1));
  OHP_CUDA(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  OHP_CUDA(cudaStreamSynchronize(ctx->stream));
  *result = ctx->h_scalars[0];
  return 0;
}
extern "C" int ohp_vec_norm2(ohp_vec x, double *result) {
  double d = 0.0;
  OHP_TRY(ohp_vec_dot(x, x, &d));
  *result = sqrt(d);
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Mat                                                                                   */
/* ------------------------------------------------------------------------------------ */
extern "C" int ohp_mat_create(ohp_mesh m, ohp_mat *out) {
  if (!m || !m->setup_done) OHP_FAIL(OHP_ERR_ORDER, "ohp_mat_create: mesh not set up");
  OHP_CUDA(cudaSetDevice(m->ctx->device));
  ohp_mat A = new ohp_mat_s();
  A->mesh = m; A->r = A->z = A->p = A->p1 = A->w = A->dinv = A->hist = A->s = A->r1 = A->z1 = A->s1 = nullptr; A->hist_cap = 0; A->nullspace = 0;
  OHP_CUDA(OHP_MALLOC(m->ctx, &A->vals, sizeof(double) * ohp_padded_nnz(m->nnz)));
  OHP_CUDA(cudaMemsetAsync(A->vals, 0, sizeof(double) * ohp_padded_nnz(m->nnz), m->ctx->stream));
  *out = A;
  return 0;
}
extern "C" int ohp_mat_destroy(ohp_mat A) {
  if (!A) return 0;
  cudaSetDevice(A->mesh->ctx->device);
  cudaStreamSynchronize(A->mesh->ctx->stream);
  OHP_FREE(A->mesh->ctx, A->vals); OHP_FREE(A->mesh->ctx, A->r); OHP_FREE(A->mesh->ctx, A->z); OHP_FREE(A->mesh->ctx, A->p); OHP_FREE(A->mesh->ctx, A->p1); OHP_FREE(A->mesh->ctx, A->s); OHP_FREE(A->mesh->ctx, A->r1); OHP_FREE(A->mesh->ctx, A->z1); OHP_FREE(A->mesh->ctx, A->s1); OHP_FREE(A->mesh->ctx, A->pm0); OHP_FREE(A->mesh->ctx, A->pm1); OHP_FREE(A->mesh->ctx, A->pw0); OHP_FREE(A->mesh->ctx, A->pw1); OHP_FREE(A->mesh->ctx, A->w); OHP_FREE(A->mesh->ctx, A->dinv); OHP_FREE(A->mesh->ctx, A->hist);
  delete A;
  return 0;
}
extern "C" int ohp_mat_zero(ohp_mat A) {
  OHP_CUDA(cudaSetDevice(A->mesh->ctx->device));
  OHP_CUDA(cudaMemsetAsync(A->vals, 0, sizeof(double) * A->mesh->nnz, A->mesh->ctx->stream));
  return 0;
}
extern "C" int ohp_mat_get_values(ohp_mat A, double *vals) {
  OHP_CUDA(cudaSetDevice(A->mesh->ctx->device));
  OHP_CUDA(cudaMemcpyAsync(vals, A->vals, sizeof(double) * A->mesh->nnz, cudaMemcpyDeviceToHost, A->mesh->ctx->stream));
  OHP_CUDA(cudaStreamSynchronize(A->mesh->ctx->stream));
  return 0;
}
extern "C" int ohp_mat_mult(ohp_mat A, ohp_vec x, ohp_vec y) {
  if (x->n != A->mesh->n_owned || y->n != A->mesh->n_owned) OHP_FAIL(OHP_ERR_ARG_SIZ, "ohp_mat_mult: size mismatch");
  if (x == y) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_mat_mult: x and y must differ");
  OHP_CUDA(cudaSetDevice(A->mesh->ctx->device));
  if (A->mesh->ctx->nranks > 1) OHP_TRY(ohp_halo_exchange(A->mesh, x->d));
  return ohp_spmv_launch(A, x->d, y->d);
}

/* ------------------------------------------------------------------------------------ */
/* CG                                                                                    */
/* ------------------------------------------------------------------------------------ */
template <bool JACOBI>
__global__ void __launch_bounds__(256)
k_cg_init(const double *__restrict__ b, double *__restrict__ x, double *__restrict__ r, double *__restrict__ z,
          const double *__restrict__ dinv, int n, ohp_cg_state *st, double *partials, unsigned *ticket,
          double *sums, double *hist, int mode, const ohp_p2p_dev pd) {
  double s[2] = {0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double bi = b[i];
    x[i] = 0.0;
    r[i] = bi;
    double zi = bi;
    if (JACOBI) { zi = dinv[i] * bi; z[i] = zi; }
    s[0] += zi * zi;
    s[1] += zi * bi;
  }
  if (grid_sum<2>(s, partials, ticket)) {
    if (mode == OHP_MODE_P2P) p2p_allreduce<2>(pd, s);
    if (threadIdx.x == 0) {
      sums[0] = s[0]; sums[1] = s[1];
      if (mode != OHP_MODE_NCCL) cg_fin_init(st, s[0], s[1], hist);
    }
  }
}

/* x += alpha_prev p_old (the x update of the PREVIOUS iteration, deferred so that p is read once);
 * p_new = z + (beta/betaold) p_old  (first iteration: p_new = z).  The search direction is
 * double-buffered: p_new = P[its & 1], p_old = P[(its & 1) ^ 1].
 * P2P mode: block 0 is the halo block.  It recomputes p_new for the dofs the neighbours ghost and
 * stores them straight into the neighbours' vectors over NVLink (no pack buffer, no extra launch),
 * publishes a sequence flag per neighbour, then waits for the neighbours' flags, so that when this
 * kernel retires the ghost tail of P[its & 1] is complete: halo exchange fused into the p update. */
template <int MODE>
__global__ void __launch_bounds__(256)
k_cg_pupdate(double *__restrict__ p0, double *__restrict__ p1, const double *__restrict__ z, double *__restrict__ x,
             int n, const ohp_cg_state *st, const ohp_p2p_dev pd, int n_halo_blocks, unsigned *halo_ticket, int wait_here) {
  pdl_wait();
  pdl_launch();
  if (st->done) return;
  const int its = st->its;
  const bool first = (its == 0);
  const double bb = first ? 0.0 : st->beta / st->betaold;
  const double a = st->alpha;
  double *__restrict__ pn = (its & 1) ? p1 : p0;
  const double *__restrict__ po = (its & 1) ? p0 : p1;
  int b = blockIdx.x, nb = gridDim.x;
  if (MODE == OHP_MODE_P2P) {
    if (b < n_halo_blocks) {
      const int nsend = pd.soff[pd.n_nbr];
      for (int i = b * blockDim.x + threadIdx.x; i < nsend; i += n_halo_blocks * blockDim.x) {
        const int l = pd.send_ldof[i];
        const double val = first ? z[l] : fma(bb, po[l], z[l]);
        int k = 0;
        while (i >= pd.soff[k + 1]) ++k;
        pd.p[its & 1][pd.nbr[k]][pd.dst_off[k] + (i - pd.soff[k])] = val;
      }
      __threadfence_system();
      __syncthreads();
      __shared__ unsigned long long s_seq;
      __shared__ bool s_last;
      if (threadIdx.x == 0) {
        s_last = (atomicInc(halo_ticket, n_halo_blocks - 1) == (unsigned)(n_halo_blocks - 1));
        if (s_last) { __threadfence(); s_seq = pd.seq[1] + 1; pd.seq[1] = s_seq; }
      }
      __syncthreads();
      if (s_last && threadIdx.x < pd.n_nbr) { /* every push of this rank has landed: publish, then wait for the neighbours */
        const int q = pd.nbr[threadIdx.x];
        st_release_sys(pd.hflags[q] + pd.rank, s_seq);
        if (wait_here) p2p_wait(pd.hflags[pd.rank] + q, s_seq, pd.seq + 2, pd.timeout_ns);
      }
      return;
    }
    b -= n_halo_blocks; nb -= n_halo_blocks;
  }
  if (first) {
    for (int i = b * blockDim.x + threadIdx.x; i < n; i += nb * blockDim.x) pn[i] = z[i];
  } else {
    for (int i = b * blockDim.x + threadIdx.x; i < n; i += nb * blockDim.x) {
      const double pi = po[i];
      x[i] += a * pi;
      pn[i] = fma(bb, pi, z[i]);
    }
  }
}

/* the x update still pending when the iteration stops */
__global__ void __launch_bounds__(256)
k_cg_final_x(double *__restrict__ x, const double *__restrict__ p0, const double *__restrict__ p1, int n,
             ohp_cg_state *st) {
  if (!st->x_pending) return;
  const double a = st->alpha;
  const double *__restrict__ p = ((st->its - 1) & 1) ? p1 : p0; /* direction of the last completed iteration */
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] += a * p[i];
}

/* r -= a w; z = M^-1 r; (z,z), (z,r) */
template <bool JACOBI>
__global__ void __launch_bounds__(256)
k_cg_update(double *__restrict__ r, double *__restrict__ z, const double *__restrict__ w,
            const double *__restrict__ dinv, int n, ohp_cg_state *st,
            double *partials, unsigned *ticket, double *sums, double *hist, int hist_cap, int mode,
            const ohp_p2p_dev pd) {
  pdl_wait();
  pdl_launch();
  if (st->done) return;
  const double a = st->alpha;
  double s[2] = {0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double ri = r[i] - a * w[i];
    r[i] = ri;
    double zi = ri;
    if (JACOBI) { zi = dinv[i] * ri; z[i] = zi; }
    s[0] += zi * zi;
    s[1] += zi * ri;
  }
  if (grid_sum<2>(s, partials, ticket)) {
    if (mode == OHP_MODE_P2P) p2p_allreduce<2>(pd, s);
    if (threadIdx.x == 0) {
      sums[0] = s[0]; sums[1] = s[1];
      if (mode != OHP_MODE_NCCL) cg_fin_update(st, s[0], s[1], hist, hist_cap);
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* Single-reduction CG (Chronopoulos-Gear) as two launches per iteration:                 */
/*   k_cgcg_vec : p = u + beta p; s = w + beta s; x += alpha p; r -= alpha s; u = M^-1 r,  */
/*                partial (r,u), (u,u) per block; halo blocks push the new u to neighbours */
/*   k_spmv_*   : w = A u, (w,u); its last block adds the partials, all-reduces ONCE and   */
/*                advances alpha/beta/norm/its.                                           */
/* r, u and s are double-buffered by iteration parity so that the halo blocks can recompute */
/* the values they push from buffers nobody writes during the kernel.                     */
/* ------------------------------------------------------------------------------------ */
template <int MODE, bool JACOBI>
__global__ void __launch_bounds__(256)
k_cgcg_vec(double *R0, double *R1, double *U0, double *U1, double *S0, double *S1, double *__restrict__ p,
           const double *__restrict__ w, double *__restrict__ x, const double *__restrict__ b,
           const double *__restrict__ dinv, int n, const ohp_cg_state *st, double *vpart, const ohp_p2p_dev pd,
           int n_halo_blocks, unsigned *halo_ticket, int wait_here, const ohp_trace trace) {
  pdl_wait();
  pdl_launch();
  if (st->done) return;
  if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) trace_put(trace, TR_VEC_BEGIN, st->its);
  const bool init = (st->x_pending == 0);
  const int k = init ? 1 : st->its; /* init writes the index-0 buffers, like an iteration of odd parity */
  const double alpha = st->alpha, beta = st->beta;
  const double *ro = (k & 1) ? R1 : R0, *uo = (k & 1) ? U1 : U0, *so = (k & 1) ? S1 : S0;
  double *rn = (k & 1) ? R0 : R1, *un = (k & 1) ? U0 : U1, *sn = (k & 1) ? S0 : S1;
  int blk = blockIdx.x, nb = gridDim.x;
  if (MODE == OHP_MODE_P2P) {
    if (blk < n_halo_blocks) {
      const int nsend = pd.soff[pd.n_nbr];
      const int nbuf = (k & 1) ? 0 : 1; /* index of the buffer being written */
      for (int i = blk * blockDim.x + threadIdx.x; i < nsend; i += n_halo_blocks * blockDim.x) {
        const int l = pd.send_ldof[i];
        double val;
        if (init) val = JACOBI ? dinv[l] * b[l] : b[l];
        else {
          const double si = fma(beta, so[l], w[l]);
          const double ri = fma(-alpha, si, ro[l]);
          val = JACOBI ? dinv[l] * ri : ri;
        }
        int q = 0;
        while (i >= pd.soff[q + 1]) ++q;
        pd.p[nbuf][pd.nbr[q]][pd.dst_off[q] + (i - pd.soff[q])] = val;
      }
      __threadfence_system();
      __syncthreads();
      __shared__ unsigned long long s_seq;
      __shared__ bool s_last;
      if (threadIdx.x == 0) {
        s_last = (atomicInc(halo_ticket, n_halo_blocks - 1) == (unsigned)(n_halo_blocks - 1));
        if (s_last) { __threadfence(); s_seq = pd.seq[1] + 1; pd.seq[1] = s_seq; }
      }
      __syncthreads();
      if (s_last && threadIdx.x < pd.n_nbr) {
        const int q = pd.nbr[threadIdx.x];
        st_release_sys(pd.hflags[q] + pd.rank, s_seq);
        if (wait_here) p2p_wait(pd.hflags[pd.rank] + q, s_seq, pd.seq + 2, pd.timeout_ns);
      }
      if (s_last && threadIdx.x == 0) trace_put(trace, TR_VEC_HALO_PUSHED, st->its);
      return;
    }
    blk -= n_halo_blocks; nb -= n_halo_blocks;
  }
  double sums[2] = {0.0, 0.0};
  if (init) {
    for (int i = blk * blockDim.x + threadIdx.x; i < n; i += nb * blockDim.x) {
      const double bi = b[i];
      x[i] = 0.0; p[i] = 0.0; sn[i] = 0.0;
      rn[i] = bi;
      double ui = bi;
      if (JACOBI) { ui = dinv[i] * bi; un[i] = ui; }
      sums[0] += bi * ui; sums[1] += ui * ui;
    }
  } else {
    for (int i = blk * blockDim.x + threadIdx.x; i < n; i += nb * blockDim.x) {
      const double ui = uo[i];
      const double pi = fma(beta, p[i], ui);
      const double si = fma(beta, so[i], w[i]);
      p[i] = pi; sn[i] = si;
      x[i] = fma(alpha, pi, x[i]);
      const double ri = fma(-alpha, si, ro[i]);
      rn[i] = ri;
      double uu = ri;
      if (JACOBI) { uu = dinv[i] * ri; un[i] = uu; }
      sums[0] += ri * uu; sums[1] += uu * uu;
    }
  }
  __shared__ double s_red2[2 * 32];
  block_sum<2>(sums, s_red2);
  if (threadIdx.x == 0) {
    vpart[2 * blk] = sums[0]; vpart[2 * blk + 1] = sums[1];
    if (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1 || blockIdx.x == gridDim.x / 2) trace_put(trace, TR_VEC_END, st->its);
  }
}

__global__ void k_extract_dinv(const int32_t *rowptr, const int32_t *colidx, const double *vals, int n, double *dinv) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double d = 0.0;
  for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) if (colidx[k] == r) d = vals[k];
  dinv[r] = (d != 0.0) ? 1.0 / d : 1.0;
}

int ohp_extract_dinv(ohp_mat A) {
  ohp_mesh m = A->mesh;
  k_extract_dinv<<<nblk(m->n_owned, 256), 256, 0, m->ctx->stream>>>(m->rowptr, m->colidx, A->vals, m->n_owned, A->dinv);
  OHP_CHECK_LAUNCH(m->ctx);
  return 0;
}

extern "C" int ohp_ksp_default_options(ohp_ksp_options *o) {
  o->rtol = 1e-5; o->abstol = 1e-50; o->dtol = 1e5; o->max_it = 10000; o->pc = OHP_PC_NONE;
  o->check_every = 0; o->history_cap = 0; o->history = nullptr; o->profile = 0; o->ksp_type = OHP_KSP_CG;
  o->cheb_its = 0; o->reserved0 = 0; o->cheb_emin = 0.0; o->cheb_emax = 0.0;
  return 0;
}

namespace {
/* CUDA events that are destroyed on every exit path */
struct EventBag {
  std::vector<cudaEvent_t> ev;
  ~EventBag() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
  cudaError_t record(cudaStream_t st, cudaEvent_t *out = nullptr) {
    cudaEvent_t e;
    cudaError_t rc = cudaEventCreate(&e);
    if (rc != cudaSuccess) return rc;
    ev.push_back(e);
    if (out) *out = e;
    return cudaEventRecord(e, st);
  }
};
}  // namespace

/* OHP_TRACE_CG=<path>: record a device-side timeline of iterations [OHP_TRACE_IT0 (default 200), +OHP_TRACE_N
 * (default 40)) of every solve and write it to <path>.rank<r>.csv (columns: event, iteration, block, ns) */
ohp_trace ohp_cg_trace_open(ohp_context ctx) {
  static const char *path = getenv("OHP_TRACE_CG");
  ohp_trace none = {};
  if (!path) return none;
  if (!ctx->trace.buf) {
    const unsigned cap = 1u << 18;
    if (OHP_MALLOC(ctx, &ctx->trace.buf, sizeof(ohp_trace_rec) * cap) != cudaSuccess) return none;
    if (OHP_MALLOC(ctx, &ctx->trace.count, sizeof(unsigned)) != cudaSuccess) { OHP_FREE(ctx, ctx->trace.buf); ctx->trace.buf = nullptr; return none; }
    ctx->trace.cap = cap;
    const char *i0 = getenv("OHP_TRACE_IT0"), *n = getenv("OHP_TRACE_N");
    ctx->trace.it0 = i0 ? atoi(i0) : 200;
    ctx->trace.it1 = ctx->trace.it0 + (n ? atoi(n) : 40);
  }
  cudaMemsetAsync(ctx->trace.count, 0, sizeof(unsigned), ctx->stream);
  return ctx->trace;
}
void ohp_cg_trace_dump(ohp_context ctx, const char *algo) {
  static const char *path = getenv("OHP_TRACE_CG");
  if (!path || !ctx->trace.buf) return;
  unsigned n = 0;
  cudaMemcpyAsync(&n, ctx->trace.count, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  n = std::min(n, ctx->trace.cap);
  std::vector<ohp_trace_rec> rec(n);
  cudaMemcpy(rec.data(), ctx->trace.buf, sizeof(ohp_trace_rec) * n, cudaMemcpyDeviceToHost);
  std::sort(rec.begin(), rec.end(), [](const ohp_trace_rec &a, const ohp_trace_rec &b) { return a.t < b.t; });
  char name[1200];
  snprintf(name, sizeof(name), "%s.rank%d.csv", path, ctx->rank);
  FILE *f = fopen(name, "w");
  if (!f) return;
  fprintf(f, "# algo=%s ranks=%d events: 1 vec_begin 2 vec_halo_pushed 3 vec_end 4 spmv_begin 5 spmv_halo_ready 6 spmv_cta_done 7 spmv_last_block "
             "8 allreduce_posted 9 allreduce_done 10 spmv_end 11 upd_begin 12 upd_end\nevent,iteration,block,ns\n", algo, ctx->nranks);
  for (const ohp_trace_rec &r : rec) fprintf(f, "%d,%d,%d,%llu\n", r.ev, r.it, r.blk, r.t);
  fclose(f);
}

static int cgcg_solve(ohp_mat A, ohp_vec b, ohp_vec x, const ohp_ksp_options *o, ohp_ksp_result *res, bool p2p) {
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
  OHP_TRY(need(&A->p)); OHP_TRY(need(&A->w)); OHP_TRY(need(&A->s)); OHP_TRY(need(&A->s1));
  static const ohp_p2p_dev no_p2p = {};
  const ohp_p2p_dev &pd = p2p ? m->p2p->dev : no_p2p;
  double *U0, *U1, *R0, *R1;
  if (p2p) { U0 = pd.p[0][ctx->rank]; U1 = pd.p[1][ctx->rank]; }
  else if (jac) { OHP_TRY(need(&A->z)); OHP_TRY(need(&A->z1)); U0 = A->z; U1 = A->z1; }
  else { OHP_TRY(need(&A->r)); OHP_TRY(need(&A->r1)); U0 = A->r; U1 = A->r1; }
  if (jac) {
    OHP_TRY(need(&A->r)); OHP_TRY(need(&A->r1)); OHP_TRY(need(&A->dinv));
    R0 = A->r; R1 = A->r1;
    OHP_TRY(ohp_extract_dinv(A));
  } else { R0 = U0; R1 = U1; }
  double *hist = nullptr;
  int hist_cap = 0;
  if (o->history && o->history_cap > 0) {
    if (A->hist_cap < o->history_cap) {
      OHP_FREE(ctx, A->hist);
      OHP_CUDA(OHP_MALLOC(ctx, &A->hist, sizeof(double) * o->history_cap));
      A->hist_cap = o->history_cap;
    }
    hist = A->hist; hist_cap = o->history_cap;
    OHP_CUDA(cudaMemsetAsync(hist, 0, sizeof(double) * hist_cap, st));
  }
  ohp_cg_state init;
  memset(&init, 0, sizeof(init));
  init.rtol = o->rtol; init.abstol = o->abstol; init.dtol = o->dtol; init.maxit = o->max_it;
  *ctx->h_cg = init;
  OHP_CUDA(cudaMemcpyAsync(ctx->d_cg, ctx->h_cg, sizeof(init), cudaMemcpyHostToDevice, st));
  const int mode = p2p ? OHP_MODE_P2P : OHP_MODE_SERIAL;
  const int vgrid = vec_grid(ctx, n);
  const int nhalo = p2p ? std::max(1, std::min(16, (pd.soff[pd.n_nbr] + 255) / 256)) : 0;
  double *vpart = ctx->d_partials + kMaxPartials; /* second half of the scratch: the SpMV uses the first */
  ohp_cgcg_epi epi;
  memset(&epi, 0, sizeof(epi));
  epi.vpart = vpart; epi.nvpart = vgrid; epi.jacobi = jac ? 1 : 0; epi.hist = hist; epi.hist_cap = hist_cap;
  epi.trace = ohp_cg_trace_open(ctx);
  if (p2p) OHP_TRY(ohp_p2p_entry_barrier(m));
  static const char *nooverlap = getenv("OHP_NOOVERLAP");
  const bool overlap = p2p && spmv_use_tma(m) && m->ghost_cta && !nooverlap && !getenv("OHP_SPMV");
  const int wait_in_vec = overlap ? 0 : 1;
  epi.ghost_cta = overlap ? (m->cta_tile0 ? m->ghost_cta_w : m->ghost_cta) : nullptr;
  epi.cta_tile0 = overlap ? m->cta_tile0 : nullptr;
  static const char *nopdl = getenv("OHP_NOPDL");
  const bool pdl = !nopdl;
  const int check_every = o->check_every > 0 ? o->check_every : 50;
  EventBag bag, profbag;
  cudaEvent_t ev0, ev1;
  OHP_CUDA(bag.record(st, &ev0));
  int launched = 0;
  std::vector<cudaEvent_t> &prof = profbag.ev; /* optional per-launch timing of the SpMV, skipping the first (cold) chunk */
  const int nprof = std::max(0, o->profile);
  for (;;) {
    for (int it = 0; it < check_every; ++it) {
      if (p2p) {
        if (jac) OHP_CUDA(ohp_launch(k_cgcg_vec<OHP_MODE_P2P, true>, dim3(vgrid + nhalo), dim3(256), 0, st, pdl, R0, R1, U0, U1, A->s, A->s1, A->p, (const double *)A->w, x->d, (const double *)b->d, (const double *)A->dinv, n, (const ohp_cg_state *)ctx->d_cg, vpart, pd, nhalo, ctx->d_ticket + 2, wait_in_vec, epi.trace));
        else OHP_CUDA(ohp_launch(k_cgcg_vec<OHP_MODE_P2P, false>, dim3(vgrid + nhalo), dim3(256), 0, st, pdl, R0, R1, U0, U1, A->s, A->s1, A->p, (const double *)A->w, x->d, (const double *)b->d, (const double *)A->dinv, n, (const ohp_cg_state *)ctx->d_cg, vpart, pd, nhalo, ctx->d_ticket + 2, wait_in_vec, epi.trace));
      } else {
        if (jac) OHP_CUDA(ohp_launch(k_cgcg_vec<OHP_MODE_SERIAL, true>, dim3(vgrid), dim3(256), 0, st, pdl, R0, R1, U0, U1, A->s, A->s1, A->p, (const double *)A->w, x->d, (const double *)b->d, (const double *)A->dinv, n, (const ohp_cg_state *)ctx->d_cg, vpart, pd, 0, (unsigned *)nullptr, 1, epi.trace));
        else OHP_CUDA(ohp_launch(k_cgcg_vec<OHP_MODE_SERIAL, false>, dim3(vgrid), dim3(256), 0, st, pdl, R0, R1, U0, U1, A->s, A->s1, A->p, (const double *)A->w, x->d, (const double *)b->d, (const double *)A->dinv, n, (const ohp_cg_state *)ctx->d_cg, vpart, pd, 0, (unsigned *)nullptr, 1, epi.trace));
      }
      ctx->launches++;
      const bool timed = (int)prof.size() < 2 * nprof && (launched > 0 || o->max_it <= check_every);
      if (timed) OHP_CUDA(profbag.record(st));
      OHP_TRY(spmv_enqueue(A, U0, A->w, ctx->d_cg, ctx->d_scalars, mode, U1, pd, epi));
      if (timed) OHP_CUDA(profbag.record(st));
    }
    launched += check_every;
    OHP_CUDA(cudaMemcpyAsync(ctx->h_cg, ctx->d_cg, sizeof(ohp_cg_state), cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
    if (ctx->h_cg->done) break;
    if (launched > o->max_it + 2 * check_every) OHP_FAIL(OHP_ERR_LIB, "ohp_ksp_cg (single-reduction): iteration control did not terminate");
  }
  OHP_CUDA(bag.record(st, &ev1));
  OHP_CUDA(cudaEventSynchronize(ev1));
  res->total_ms = 0.f;
  cudaEventElapsedTime(&res->total_ms, ev0, ev1);
  res->its = ctx->h_cg->its; res->reason = ctx->h_cg->reason; res->rnorm = ctx->h_cg->dp; res->rnorm0 = ctx->h_cg->rnorm0;
  res->spmv_ms = 0.f; res->spmv_samples = 0;
  ohp_cg_trace_dump(ctx, "cgcg");
  res->algo = OHP_ALGO_SINGLE_REDUCTION;
  {
    /* SpMV launch number j closes iteration j - 1 (launch 0 is the initial pass): count the ones that did real work */
    double acc = 0.0;
    int used = 0;
    const int first_timed = (o->max_it <= check_every) ? 0 : check_every;
    for (size_t i = 0; i + 1 < prof.size(); i += 2) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, prof[i], prof[i + 1]);
      if (first_timed + (int)(i / 2) <= res->its) { acc += ms; used++; }
    }
    if (used) { res->spmv_ms = (float)(acc / used); res->spmv_samples = used; }
  }
  if (p2p) OHP_TRY(ohp_p2p_check(m, "ohp_ksp_cg (single-reduction)"));
  if (hist) {
    const int ncopy = std::min(hist_cap, res->its + 1);
    OHP_CUDA(cudaMemcpyAsync(o->history, hist, sizeof(double) * ncopy, cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}


/* ------------------------------------------------------------------------------------ */
/* Pipelined CG (Ghysels & Vanroose; PETSc's -ksp_type pipecg) as two launches per iteration: */
/*   k_spmv_*      : n = A m  (m = M^-1 w); an extra warp of its CTA 0 meanwhile reduces the     */
/*                   previous vector kernel's partial sums, all-reduces them and advances       */
/*                   alpha / beta / norm / its (pipe_reduce)                                    */
/*   k_pipecg_vec  : z = n + beta z; s = w + beta s; p = u + beta p; x += alpha p; r -= alpha s; */
/*                   w -= alpha z; u = M^-1 r; m = M^-1 w; partial (r,u), (w,u), (u,u) per block; */
/*                   halo blocks push the new m to the neighbours                               */
/* Same iterates as CG in exact arithmetic, same scalar formulas as the single-reduction         */
/* recurrence (cgcg_scalars), same convergence test (KSPConvergedDefault on ||u||).  What it     */
/* buys in the strong-scaling regime: the dot products of iteration i are not needed before the  */
/* vector kernel of iteration i+1, so the grid-wide reduction, the 8-rank all-reduce and up to   */
/* one SpMV duration of rank skew overlap the matrix stream instead of sitting between the two   */
/* kernels (device timeline of the two-launch single-reduction loop at 1 M rows: 16 of 43 us).   */
/* Price: 104 instead of 88 bytes of vector traffic per row and two extra SpMVs per solve.       */
/*                                                                                               */
/* Attainable accuracy.  The recurrences for w = A u, s = A p, z = A q drift from the products    */
/* they stand for, and r from b - A x (measured: on the 2048^2 box the TRUE residual stalled at   */
/* 2.8e-8 |b| after 5454 iterations although the recurred norm met rtol = 1e-9).  So every         */
/* kPipeReplaceEvery iterations one pass RECOMPUTES them (the residual-replacement form of         */
/* Cools & Vanroose): p = u + beta p and x += alpha p as usual, then s = A p, z = A M^-1 s,         */
/* r = b - A x, w = A M^-1 r with four extra SpMVs -- 1 % more SpMVs at a period of 400, after       */
/* which iteration counts and true residuals match classic CG (tests/, bench check object).       */
/*                                                                                               */
/* The host knows the pass number L (the device only knows `done`), so each launch is told what    */
/* to do: kind 0 sets up r = b, u, m = u; kind 1 takes w = A u from the SpMV output; kind 2 is a    */
/* regular iteration.  m is double-buffered: every halo push alternates between the two exported   */
/* buffers and every pass makes an odd number of pushes (1, or 5 in a replacement pass), so pass L  */
/* leaves m in buffer L & 1; w (== m without a preconditioner) and z follow the same parity, which  */
/* lets the halo blocks recompute what they push from buffers nobody writes during the kernel.    */
/* A push into a buffer is safe because the neighbour's SpMV that last read it completed before the */
/* neighbour's previous push, which this rank's previous SpMV has waited for.                      */
/* ------------------------------------------------------------------------------------ */
constexpr int kPipeReplaceEvery = 400;
enum { PK_INIT = 0, PK_TAKE_W = 1, PK_ITER = 2, PK_PX = 3, PK_STAGE_S = 4, PK_STAGE_X = 5, PK_NEW_R = 6, PK_STAGE_P = 7 };
struct ohp_pipe_vecs {
  double *M0, *M1;   /* SpMV input m = M^-1 w (the NVLink-exported buffers on a partitioned run) */
  double *W0, *W1;   /* w (== M when there is no preconditioner) */
  double *Z0, *Z1;
  double *r, *s, *p, *x;
  const double *nb, *b, *dinv;
  unsigned *tcnt;    /* tile counter of the dynamically scheduled SpMV that follows: re-armed here (NULL = none) */
  unsigned tinit;
};

/* the value pass `kind` stages into the SpMV input buffer for dof l: computed from arrays that no block writes in that
 * pass, so the halo blocks and the row blocks get bitwise the same number */
template <bool JACOBI>
__device__ __forceinline__ double pipe_staged(const ohp_pipe_vecs &v, int kind, int l, double alpha, double beta,
                                              const double *wo, const double *zo) {
  double t;
  switch (kind) {
    case PK_INIT: t = v.b[l]; break;
    case PK_TAKE_W: t = v.nb[l]; break;
    case PK_ITER: { const double zi = fma(beta, zo[l], v.nb[l]); t = fma(-alpha, zi, wo[l]); break; }
    case PK_STAGE_S: t = v.s[l]; break;
    case PK_STAGE_X: return v.x[l];
    case PK_NEW_R: t = v.b[l] - v.nb[l]; break;
    default: return v.p[l]; /* PK_STAGE_P */
  }
  return JACOBI ? v.dinv[l] * t : t;
}

template <int MODE, bool JACOBI>
__global__ void __launch_bounds__(256)
k_pipecg_vec(const ohp_pipe_vecs v, int n, const ohp_cg_state *st, double *vpart, const ohp_p2p_dev pd, int n_halo_blocks,
             unsigned *halo_ticket, int wait_here, const ohp_trace trace, int kind, int cur) {
  pdl_wait();
  pdl_launch();
  if (v.tcnt && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *v.tcnt = v.tinit; /* the previous SpMV has completed */
  if (st->done) return;
  if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) trace_put(trace, TR_VEC_BEGIN, st->its);
  const double alpha = st->alpha, beta = st->beta;
  const double *__restrict__ nb = v.nb;
  const double *wo = cur ? v.W0 : v.W1, *zo = cur ? v.Z0 : v.Z1;
  double *wn = cur ? v.W1 : v.W0, *zn = cur ? v.Z1 : v.Z0, *mn = cur ? v.M1 : v.M0;
  int blk = blockIdx.x, nblocks = gridDim.x;
  if (MODE == OHP_MODE_P2P) {
    if (blk < n_halo_blocks) {
      if (kind == PK_PX) return; /* p and x are updated in place: staged by the next launch (PK_STAGE_P) */
      const int nsend = pd.soff[pd.n_nbr];
      for (int i = blk * blockDim.x + threadIdx.x; i < nsend; i += n_halo_blocks * blockDim.x) {
        const int l = pd.send_ldof[i];
        const double val = pipe_staged<JACOBI>(v, kind, l, alpha, beta, wo, zo);
        int q = 0;
        while (i >= pd.soff[q + 1]) ++q;
        pd.p[cur][pd.nbr[q]][pd.dst_off[q] + (i - pd.soff[q])] = val;
      }
      __threadfence_system();
      __syncthreads();
      __shared__ unsigned long long s_seq;
      __shared__ bool s_last;
      if (threadIdx.x == 0) {
        s_last = (atomicInc(halo_ticket, n_halo_blocks - 1) == (unsigned)(n_halo_blocks - 1));
        if (s_last) { __threadfence(); s_seq = pd.seq[1] + 1; pd.seq[1] = s_seq; }
      }
      __syncthreads();
      if (s_last && threadIdx.x < pd.n_nbr) {
        const int q = pd.nbr[threadIdx.x];
        st_release_sys(pd.hflags[q] + pd.rank, s_seq);
        if (wait_here) p2p_wait(pd.hflags[pd.rank] + q, s_seq, pd.seq + 2, pd.timeout_ns);
      }
      if (s_last && threadIdx.x == 0) trace_put(trace, TR_VEC_HALO_PUSHED, st->its);
      return;
    }
    blk -= n_halo_blocks; nblocks -= n_halo_blocks;
  }
  double sums[3] = {0.0, 0.0, 0.0};
  const int i0 = blk * blockDim.x + threadIdx.x, stride = nblocks * blockDim.x;
  switch (kind) {
    case PK_INIT: /* r = b, u = M^-1 r, m = u (first SpMV: w0 = A u0); x = p = s = z = 0 */
      for (int i = i0; i < n; i += stride) {
        v.r[i] = v.b[i]; v.x[i] = 0.0; v.p[i] = 0.0; v.s[i] = 0.0; v.Z0[i] = 0.0; v.Z1[i] = 0.0;
        mn[i] = pipe_staged<JACOBI>(v, kind, i, alpha, beta, wo, zo);
      }
      return;
    case PK_TAKE_W: /* w = A u is in the SpMV output: (r,u), (w,u), (u,u); m = M^-1 w */
      for (int i = i0; i < n; i += stride) {
        const double wi = nb[i], ri = v.r[i];
        const double ui = JACOBI ? v.dinv[i] * ri : ri;
        wn[i] = wi;
        if (JACOBI) mn[i] = v.dinv[i] * wi;
        sums[0] += ri * ui; sums[1] += wi * ui; sums[2] += ui * ui;
      }
      break;
    case PK_ITER:
      for (int i = i0; i < n; i += stride) {
        const double ro = v.r[i], woi = wo[i];
        const double ui = JACOBI ? v.dinv[i] * ro : ro;
        const double zi = fma(beta, zo[i], nb[i]);
        const double si = fma(beta, v.s[i], woi);
        const double pi = fma(beta, v.p[i], ui);
        zn[i] = zi; v.s[i] = si; v.p[i] = pi;
        v.x[i] = fma(alpha, pi, v.x[i]);
        const double ri = fma(-alpha, si, ro);
        const double wi = fma(-alpha, zi, woi);
        v.r[i] = ri; wn[i] = wi;
        const double un = JACOBI ? v.dinv[i] * ri : ri;
        if (JACOBI) mn[i] = v.dinv[i] * wi;
        sums[0] += ri * un; sums[1] += wi * un; sums[2] += un * un;
      }
      break;
    case PK_PX: /* replacement pass, first launch: p = u + beta p; x += alpha p */
      for (int i = i0; i < n; i += stride) {
        const double ro = v.r[i];
        const double pi = fma(beta, v.p[i], JACOBI ? v.dinv[i] * ro : ro);
        v.p[i] = pi;
        v.x[i] = fma(alpha, pi, v.x[i]);
      }
      return;
    case PK_NEW_R: /* r = b - A x (A x is in the SpMV output); m = u = M^-1 r */
      for (int i = i0; i < n; i += stride) {
        v.r[i] = v.b[i] - nb[i];
        mn[i] = pipe_staged<JACOBI>(v, kind, i, alpha, beta, wo, zo);
      }
      return;
    default: /* PK_STAGE_S / _X / _P: copy into the SpMV input buffer */
      for (int i = i0; i < n; i += stride) mn[i] = pipe_staged<JACOBI>(v, kind, i, alpha, beta, wo, zo);
      return;
  }
  __shared__ double s_red3[3 * 32];
  block_sum<3>(sums, s_red3);
  if (threadIdx.x == 0) {
    vpart[3 * blk] = sums[0]; vpart[3 * blk + 1] = sums[1]; vpart[3 * blk + 2] = sums[2];
    if (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1 || blockIdx.x == gridDim.x / 2) trace_put(trace, TR_VEC_END, st->its);
  }
}

static int pipecg_solve(ohp_mat A, ohp_vec b, ohp_vec x, const ohp_ksp_options *o, ohp_ksp_result *res, bool p2p) {
  ohp_mesh m = A->mesh;
  ohp_context ctx = m->ctx;
  cudaStream_t st = ctx->stream;
  const int n = m->n_owned;
  const size_t len = (size_t)std::max(1, n + m->n_ghost);
  const bool jac = (o->pc == OHP_PC_JACOBI);
  if (!(spmv_use_tma(m) || m->n_tiles == 0) || getenv("OHP_SPMV")) OHP_FAIL(OHP_ERR_SUP, "ohp_ksp_cg (pipelined): needs the TMA-pipelined SpMV (its CTA 0 carries the reducer warp)");
  auto need = [&](double **ptr) -> int {
    if (!*ptr) { OHP_CUDA(OHP_MALLOC(ctx, ptr, sizeof(double) * len)); OHP_CUDA(cudaMemsetAsync(*ptr, 0, sizeof(double) * len, st)); }
    return 0;
  };
  OHP_TRY(need(&A->p)); OHP_TRY(need(&A->w)); OHP_TRY(need(&A->s)); OHP_TRY(need(&A->r)); OHP_TRY(need(&A->z)); OHP_TRY(need(&A->z1));
  static const ohp_p2p_dev no_p2p = {};
  const ohp_p2p_dev &pd = p2p ? m->p2p->dev : no_p2p;
  ohp_pipe_vecs v;
  memset(&v, 0, sizeof(v));
  if (p2p) { v.M0 = pd.p[0][ctx->rank]; v.M1 = pd.p[1][ctx->rank]; }
  else { OHP_TRY(need(&A->pm0)); OHP_TRY(need(&A->pm1)); v.M0 = A->pm0; v.M1 = A->pm1; }
  if (jac) {
    OHP_TRY(need(&A->pw0)); OHP_TRY(need(&A->pw1)); OHP_TRY(need(&A->dinv));
    v.W0 = A->pw0; v.W1 = A->pw1;
    OHP_TRY(ohp_extract_dinv(A));
  } else { v.W0 = v.M0; v.W1 = v.M1; }
  v.Z0 = A->z; v.Z1 = A->z1;
  v.r = A->r; v.s = A->s; v.p = A->p; v.x = x->d; v.nb = A->w; v.b = b->d; v.dinv = A->dinv;
  {
    int G_, st_, ch_;
    dyn_sched_params(m, ctx->num_sms, &G_, &st_, &ch_);
    v.tcnt = (getenv("OHP_DYN") && getenv("OHP_DYN")[0] == '1') ? ctx->d_ticket + 3 : nullptr;
    v.tinit = (unsigned)(G_ * st_);
  }
  double *hist = nullptr;
  int hist_cap = 0;
  if (o->history && o->history_cap > 0) {
    if (A->hist_cap < o->history_cap) {
      OHP_FREE(ctx, A->hist);
      OHP_CUDA(OHP_MALLOC(ctx, &A->hist, sizeof(double) * o->history_cap));
      A->hist_cap = o->history_cap;
    }
    hist = A->hist; hist_cap = o->history_cap;
    OHP_CUDA(cudaMemsetAsync(hist, 0, sizeof(double) * hist_cap, st));
  }
  ohp_cg_state init;
  memset(&init, 0, sizeof(init));
  init.rtol = o->rtol; init.abstol = o->abstol; init.dtol = o->dtol; init.maxit = o->max_it;
  *ctx->h_cg = init;
  OHP_CUDA(cudaMemcpyAsync(ctx->d_cg, ctx->h_cg, sizeof(init), cudaMemcpyHostToDevice, st));
  const int mode = p2p ? OHP_MODE_P2P : OHP_MODE_SERIAL;
  /* one resident wave: 48 registers x 256 threads -> 5 blocks per SM (a second wave only adds a tail) */
  static const char *vps_env = getenv("OHP_VGRID_PER_SM");
  const int vps = vps_env ? std::max(1, atoi(vps_env)) : 5;
  const int vgrid = (int)std::min<int64_t>(std::min<int64_t>(nblk(n, 256), (int64_t)ctx->num_sms * vps), kMaxPartials / 3);
  const int nhalo = p2p ? std::max(1, std::min(16, (pd.soff[pd.n_nbr] + 255) / 256)) : 0;
  double *vpart = ctx->d_partials + kMaxPartials; /* second half of the scratch */
  ohp_cgcg_epi epi;
  memset(&epi, 0, sizeof(epi));
  epi.vpart = vpart; epi.nvpart = vgrid; epi.jacobi = jac ? 1 : 0; epi.hist = hist; epi.hist_cap = hist_cap;
  epi.trace = ohp_cg_trace_open(ctx);
  if (p2p) OHP_TRY(ohp_p2p_entry_barrier(m));
  static const char *nooverlap = getenv("OHP_NOOVERLAP");
  const bool overlap = p2p && m->ghost_cta && !nooverlap;
  const int wait_in_vec = overlap ? 0 : 1;
  epi.ghost_cta = overlap ? (m->cta_tile0 ? m->ghost_cta_w : m->ghost_cta) : nullptr;
  epi.cta_tile0 = overlap ? m->cta_tile0 : nullptr;
  static const char *nopdl = getenv("OHP_NOPDL");
  const bool pdl = !nopdl;
  static const char *rr_env = getenv("OHP_PIPE_RR"); /* replacement period (A/B runs); 0 = never */
  const int rr = rr_env ? atoi(rr_env) : kPipeReplaceEvery;
  const int check_every = o->check_every > 0 ? o->check_every : 50;
  EventBag bag, profbag;
  cudaEvent_t ev0, ev1;
  OHP_CUDA(bag.record(st, &ev0));
  std::vector<cudaEvent_t> &prof = profbag.ev;
  const int nprof = std::max(0, o->profile);
  int pushes = 0; /* halo pushes so far: the next one goes to buffer pushes & 1 */
  auto vec = [&](int kind, int cur) -> int {
    if (p2p) {
      if (jac) OHP_CUDA(ohp_launch(k_pipecg_vec<OHP_MODE_P2P, true>, dim3(vgrid + nhalo), dim3(256), 0, st, pdl, v, n, (const ohp_cg_state *)ctx->d_cg, vpart, pd, nhalo, ctx->d_ticket + 2, wait_in_vec, epi.trace, kind, cur));
      else OHP_CUDA(ohp_launch(k_pipecg_vec<OHP_MODE_P2P, false>, dim3(vgrid + nhalo), dim3(256), 0, st, pdl, v, n, (const ohp_cg_state *)ctx->d_cg, vpart, pd, nhalo, ctx->d_ticket + 2, wait_in_vec, epi.trace, kind, cur));
    } else {
      if (jac) OHP_CUDA(ohp_launch(k_pipecg_vec<OHP_MODE_SERIAL, true>, dim3(vgrid), dim3(256), 0, st, pdl, v, n, (const ohp_cg_state *)ctx->d_cg, vpart, pd, 0, (unsigned *)nullptr, 1, epi.trace, kind, cur));
      else OHP_CUDA(ohp_launch(k_pipecg_vec<OHP_MODE_SERIAL, false>, dim3(vgrid), dim3(256), 0, st, pdl, v, n, (const ohp_cg_state *)ctx->d_cg, vpart, pd, 0, (unsigned *)nullptr, 1, epi.trace, kind, cur));
    }
    ctx->launches++;
    return 0;
  };
  /* stage-and-push launch followed by the SpMV of the staged buffer into `out`; reduce = CTA 0's extra warp closes the pass */
  auto stage_mult = [&](int kind, double *out, bool reduce, bool timed) -> int {
    const int cur = pushes & 1;
    OHP_TRY(vec(kind, cur));
    pushes++;
    ohp_cgcg_epi e = epi;
    e.pipe = reduce ? 1 : 2;
    if (timed) OHP_CUDA(profbag.record(st));
    OHP_TRY(spmv_enqueue(A, cur ? v.M1 : v.M0, out, ctx->d_cg, ctx->d_scalars, mode, nullptr, pd, e));
    if (timed) OHP_CUDA(profbag.record(st));
    return 0;
  };
  int pass = 0; /* == the device's pass count while it is not done */
  for (;;) {
    for (int it = 0; it < check_every; ++it, ++pass) {
      const bool timed = (int)prof.size() < 2 * nprof && (pass >= check_every || o->max_it <= check_every);
      if (pass == 0) OHP_TRY(stage_mult(PK_INIT, A->w, false, false));
      else if (pass == 1) OHP_TRY(stage_mult(PK_TAKE_W, A->w, true, false));
      else if (rr > 0 && (pass - 2) > 0 && (pass - 2) % rr == 0) {
        /* residual replacement in place of iteration pass - 2's recurrences */
        double *znew = (pass & 1) ? v.Z1 : v.Z0;
        OHP_TRY(vec(PK_PX, pushes & 1));
        OHP_TRY(stage_mult(PK_STAGE_P, A->s, false, false));   /* s = A p */
        OHP_TRY(stage_mult(PK_STAGE_S, znew, false, false));   /* z = A M^-1 s */
        OHP_TRY(stage_mult(PK_STAGE_X, A->w, false, false));   /* A x */
        OHP_TRY(stage_mult(PK_NEW_R, A->w, false, false));     /* r = b - A x; w = A M^-1 r (into the SpMV output) */
        OHP_TRY(stage_mult(PK_TAKE_W, A->w, true, false));     /* w, dot products, m = M^-1 w; n = A m */
      } else OHP_TRY(stage_mult(PK_ITER, A->w, true, timed));
    }
    OHP_CUDA(cudaMemcpyAsync(ctx->h_cg, ctx->d_cg, sizeof(ohp_cg_state), cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
    if (ctx->h_cg->done) break;
    if (pass > o->max_it + 3 * check_every) OHP_FAIL(OHP_ERR_LIB, "ohp_ksp_cg (pipelined): iteration control did not terminate");
  }
  OHP_CUDA(bag.record(st, &ev1));
  OHP_CUDA(cudaEventSynchronize(ev1));
  res->total_ms = 0.f;
  cudaEventElapsedTime(&res->total_ms, ev0, ev1);
  res->its = ctx->h_cg->its; res->reason = ctx->h_cg->reason; res->rnorm = ctx->h_cg->dp; res->rnorm0 = ctx->h_cg->rnorm0;
  res->spmv_ms = 0.f; res->spmv_samples = 0;
  ohp_cg_trace_dump(ctx, "pipecg");
  res->algo = OHP_ALGO_PIPELINED_2L;
  {
    /* the timed launches are the regular passes from `check_every` on: pass L works while L - 2 < its */
    double acc = 0.0;
    int used = 0, L = (o->max_it <= check_every) ? 2 : check_every;
    for (size_t i = 0; i + 1 < prof.size(); i += 2, ++L) {
      while (rr > 0 && L > 2 && (L - 2) % rr == 0) ++L; /* replacement passes were not timed */
      float ms = 0.f;
      cudaEventElapsedTime(&ms, prof[i], prof[i + 1]);
      if (L - 2 < res->its) { acc += ms; used++; }
    }
    if (used) { res->spmv_ms = (float)(acc / used); res->spmv_samples = used; }
  }
  if (p2p) OHP_TRY(ohp_p2p_check(m, "ohp_ksp_cg (pipelined)"));
  if (hist) {
    const int ncopy = std::min(hist_cap, res->its + 1);
    OHP_CUDA(cudaMemcpyAsync(o->history, hist, sizeof(double) * ncopy, cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}


/* ------------------------------------------------------------------------------------ */
/* Pipelined CG, ONE launch per iteration (k_pipecg_fused): the vector update is applied   */
/* row by row inside the SpMV's consumer loop, while n = (A m)[row] is still in a register. */
/*   per row:  z = n + beta z;  s = w + beta s;  p = u + beta p;  x += alpha p;            */
/*             r -= alpha s;  w' = w - alpha z;  m' = M^-1 w'   (w' into the OTHER m buffer)  */
/* Every update is row-local once n is known, and n only needs the PREVIOUS iteration's m,  */
/* so the kernel boundary is the only grid-wide synchronisation: no vector kernel, no n in  */
/* memory (16 B/row less than the two-launch form: 96 B/row of vector traffic), one kernel   */
/* boundary per iteration instead of two.                                                  */
/* Scalars.  Each CTA leaves its partial (r,u), (w,u), (u,u) in parts[parity][cta]; at the    */
/* start of the NEXT launch warp 0 of every CTA adds the G partials in a fixed order (the     */
/* boundary made them visible: no ticket, no last block on one GPU), adds the other ranks'   */
/* totals from its local mailboxes in rank order, and runs the scalar recurrences on the      */
/* CTA's OWN replica of the solver state -- all replicas on all ranks stay bit-identical, CTA  */
/* 0 mirrors its replica for the host.  Across ranks the total is posted by the last CTA of    */
/* the rank to finish (ticket) straight into the peers' mailboxes; the poll overlaps the next  */
/* launch's prologue and TMA prefetch.                                                       */
/* Halo.  The CTAs that own rows some neighbour ghosts store w' for those rows into the        */
/* neighbours' vectors when their tiles are done; the last of them (and of the CTAs that read   */
/* ghost columns: a neighbour must not overwrite a tail somebody here still reads) publishes    */
/* the flags.                                                                                 */
/* Set-up passes, the scalar step of a replacement pass and the replacement passes themselves   */
/* reuse the two-launch pieces above (k_pipecg_vec kinds + plain SpMVs).                        */
/* ------------------------------------------------------------------------------------ */
struct ohp_pf_plan {
  int G = 0, n_contrib = 1;
  int32_t *hptr = nullptr, *hidx = nullptr; /* CSR over CTAs of the send-list entries whose row the CTA owns */
  uint8_t *contrib = nullptr;               /* per CTA: takes part in the halo arrival count */
  ohp_cg_state *rep = nullptr;              /* G replicas of the solver state */
  double *parts = nullptr;                  /* 2 parities x G x 3 */
  unsigned *cnt = nullptr;                  /* [0] halo arrivals, [1] rank-total ticket */
};
void ohp_pf_plan_free(ohp_context ctx, ohp_pf_plan *pl) {
  if (!pl) return;
  OHP_FREE(ctx, pl->hptr); OHP_FREE(ctx, pl->hidx); OHP_FREE(ctx, pl->contrib); OHP_FREE(ctx, pl->rep); OHP_FREE(ctx, pl->parts); OHP_FREE(ctx, pl->cnt);
  delete pl;
}
struct ohp_pf_dev {
  int G, n_contrib, rpar; /* rpar: parity of the partials this launch READS (it writes the other one) */
  const int32_t *hptr, *hidx;
  const uint8_t *contrib, *ghost_cta;
  ohp_cg_state *rep, *st_host;
  double *parts;
  unsigned *cnt;
  double *hist;
  int hist_cap, jacobi;
  ohp_trace trace;
};

/* G partial triples -> their sum in lane 0..31 (every lane): lane-strided loads, in-lane adds in ascending order, fixed
 * butterfly.  Used by every consumer of the partials so that all of them get bitwise the same totals. */
__device__ __forceinline__ void pf_sum_parts(const double *parts, int G, int lane, double (&t)[3]) {
  t[0] = t[1] = t[2] = 0.0;
  for (int b0 = 0; b0 < G; b0 += 32 * 8) {
    double a[8][3];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = b0 + j * 32 + lane;
#pragma unroll
      for (int k = 0; k < 3; ++k) a[j][k] = b < G ? __ldcg(parts + 3 * b + k) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int k = 0; k < 3; ++k) t[k] += a[j][k];
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o; o >>= 1) t[k] += __shfl_xor_sync(0xffffffffu, t[k], o);
  }
}
/* this rank's total -> the peers' mailboxes (one warp; t valid in every lane) */
__device__ __forceinline__ void pf_post_total(const ohp_p2p_dev &pd, const double (&t)[3], int lane) {
  unsigned long long seq = 0;
  if (lane == 0) { seq = pd.seq[0] + 1; pd.seq[0] = seq; }
  seq = __shfl_sync(0xffffffffu, seq, 0);
  if (lane < pd.nranks && lane != pd.rank) {
    ohp_ar_slot *s = pd.slots[lane] + (int)(seq & 1ull) * kMaxRanks + pd.rank;
#pragma unroll
    for (int k = 0; k < 3; ++k) s->v[k] = t[k];
    st_release_sys(&s->seq, seq);
  }
}
/* own total + the peers' posted totals, added in rank order (one warp; result in every lane) */
__device__ __forceinline__ void pf_gather_totals(const ohp_p2p_dev &pd, double (&t)[3], int lane) {
  const unsigned long long seq = pd.seq[0]; /* bumped by the poster of the previous launch */
  double got[3] = {t[0], t[1], t[2]};
  if (lane < pd.nranks && lane != pd.rank) {
    const ohp_ar_slot *r = pd.slots[pd.rank] + (int)(seq & 1ull) * kMaxRanks + lane;
    p2p_wait(&r->seq, seq, pd.seq + 2, pd.timeout_ns);
#pragma unroll
    for (int k = 0; k < 3; ++k) got[k] = r->v[k];
  }
  double tot[3] = {0.0, 0.0, 0.0};
  for (int q = 0; q < pd.nranks; ++q) {
#pragma unroll
    for (int k = 0; k < 3; ++k) tot[k] += __shfl_sync(0xffffffffu, got[k], q);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) t[k] = tot[k];
}

/* the per-row update of the consumer loop: load(row) issues the row's vector operands (before the row is summed, so
 * that their latency overlaps the gather), store(row, n, operands) applies the update with n = (A m)[row] */
template <bool JACOBI>
struct PipeRowUpdate {
  const double *wo;
  double *wn, *mn, *z, *s, *p, *r, *x;
  const double *dinv;
  double alpha, beta;
  double g, d, nn;
  struct Opnd { double ro, wo, z, s, p, x, di; };
  __device__ __forceinline__ Opnd load(int i) const {
    Opnd o;
    o.ro = r[i]; o.wo = wo[i]; o.z = z[i]; o.s = s[i]; o.p = p[i]; o.x = x[i];
    o.di = JACOBI ? dinv[i] : 1.0;
    return o;
  }
  __device__ __forceinline__ void store(int i, double n, const Opnd &o) {
    const double ui = JACOBI ? o.di * o.ro : o.ro;
    const double zi = fma(beta, o.z, n);
    const double si = fma(beta, o.s, o.wo);
    const double pi = fma(beta, o.p, ui);
    z[i] = zi; s[i] = si; p[i] = pi;
    x[i] = fma(alpha, pi, o.x);
    const double ri = fma(-alpha, si, o.ro);
    const double wi = fma(-alpha, zi, o.wo);
    r[i] = ri; wn[i] = wi;
    const double un = JACOBI ? o.di * ri : ri;
    if (JACOBI) mn[i] = o.di * wi;
    g += ri * un; d += wi * un; nn += un * un;
  }
  __device__ __forceinline__ void operator()(int i, double n) { const Opnd o = load(i); store(i, n, o); }
};

template <int S, int NC, int GR, class IDX, bool SPLIT, bool JACOBI, int MODE>
__global__ void __launch_bounds__(NC + 32, 1)
k_pipecg_fused(const int32_t *__restrict__ rowptr, const IDX *__restrict__ colidx, const double *__restrict__ vals,
               const int32_t *__restrict__ tile_row, int n_tiles, const ohp_pipe_vecs v, const ohp_pf_dev f,
               const ohp_p2p_dev pd, int cur, int pin_tiles) {
  constexpr int T = kSpmvTile;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *vals_s = reinterpret_cast<double *>(smem_raw);
  IDX *cols_s = reinterpret_cast<IDX *>(vals_s + S * T);
  int32_t *rp_s = reinterpret_cast<int32_t *>(cols_s + S * T);
  uint64_t *bars = reinterpret_cast<uint64_t *>(rp_s + S * kTmaRpStride);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
  __shared__ double s_scal[2];
  __shared__ int s_flag[2];          /* [0] done, [1] this CTA took the last rank-total ticket */
  __shared__ double s_red[3 * 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = (int)(((int64_t)n_tiles * blockIdx.x) / gridDim.x);
  const int t1 = (int)(((int64_t)n_tiles * (blockIdx.x + 1)) / gridDim.x);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 33); mbar_init(empty0 + 8 * s, NC / 32 / GR); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const bool is_producer = (warp == NC / 32);
  int issued = 0;
  const uint64_t pol_last = l2_policy_evict_last(), pol_first = l2_policy_evict_first();
  auto issue_tile = [&](int t) {
    const int i = t - t0, s = i % S;
    const int rs = __ldg(tile_row + t), re = __ldg(tile_row + t + 1);
    if (lane == 0) {
      mbar_expect_tx(full0 + 8 * s, T * (8 + (int)sizeof(IDX)));
      if (pin_tiles != 0) {
        const uint64_t pol = (i < pin_tiles) ? pol_last : pol_first;
        tma_load_1d_hint(smem_u32(vals_s + s * T), vals + (size_t)t * T, T * 8, full0 + 8 * s, pol);
        tma_load_1d_hint(smem_u32(cols_s + s * T), colidx + (size_t)t * T, T * (int)sizeof(IDX), full0 + 8 * s, pol);
      } else {
        tma_load_1d(smem_u32(vals_s + s * T), vals + (size_t)t * T, T * 8, full0 + 8 * s);
        tma_load_1d(smem_u32(cols_s + s * T), colidx + (size_t)t * T, T * (int)sizeof(IDX), full0 + 8 * s);
      }
    }
    int32_t *rp = rp_s + s * kTmaRpStride;
    for (int j = lane; j <= re - rs; j += 32) cp_async4(smem_u32(rp + j), rowptr + rs + j);
    if (lane < 2) cp_async4(smem_u32(rp + kSpmvRowCap + 2 + lane), tile_row + t + lane);
    cp_async_mbar_arrive_noinc(full0 + 8 * s);
  };
  /* the matrix is constant: stream the first stages while the previous launch is still draining */
  if (is_producer) {
    for (; issued < S && t0 + issued < t1; ++issued) issue_tile(t0 + issued);
  }
  pdl_wait();
  pdl_launch();
  /* ---- scalars of this iteration, redundantly in every CTA ---- */
  ohp_cg_state *me = f.rep + blockIdx.x;
  if (warp == 0) {
    int done = me->done;
    if (lane == 0 && f.trace.buf) trace_put(f.trace, TR_VEC_BEGIN, me->its + (me->x_pending ? 1 : 0));
    if (!done) {
      double t[3];
      pf_sum_parts(f.parts + (size_t)f.rpar * 3 * f.G, f.G, lane, t);
      if (MODE == OHP_MODE_P2P) pf_gather_totals(pd, t, lane);
      if (lane == 0) {
        ohp_cg_state loc = *me; /* one round trip for the whole replica, the recurrences then run in registers */
        cgcg_scalars(&loc, blockIdx.x == 0 ? f.hist : nullptr, f.hist_cap, t[0], t[1], f.jacobi ? t[2] : t[0]);
        *me = loc;
        done = loc.done;
        s_scal[0] = loc.alpha; s_scal[1] = loc.beta;
        if (blockIdx.x == 0) *f.st_host = loc;
        trace_put(f.trace, TR_SPMV_BEGIN, loc.its);
      }
    }
    if (lane == 0) { s_flag[0] = done; s_flag[1] = 0; }
  }
  __syncthreads();
  if (s_flag[0]) {
    if (is_producer) for (int i = 0; i < issued; ++i) mbar_wait(full0 + 8 * (i % S), 0);
    return;
  }
  if (is_producer) {
    for (int t = t0 + issued; t < t1; ++t) {
      const int i = t - t0, s = i % S;
      mbar_wait(empty0 + 8 * s, ((i / S) & 1) ^ 1);
      issue_tile(t);
    }
    return;
  }
  /* ---------------- consumer warps ---------------- */
  const bool ghost_cta = MODE == OHP_MODE_P2P && f.ghost_cta && f.ghost_cta[blockIdx.x];
  const bool contrib = MODE == OHP_MODE_P2P && f.contrib[blockIdx.x];
  unsigned long long halo_seq = 0;
  if (MODE == OHP_MODE_P2P) halo_seq = pd.seq[1]; /* the neighbours' pushes of the previous launch carry this number */
  PipeRowUpdate<JACOBI> up;
  up.wo = cur ? v.W0 : v.W1; up.wn = cur ? v.W1 : v.W0; up.mn = cur ? v.M1 : v.M0;
  up.z = v.Z0; up.s = v.s; up.p = v.p; up.r = v.r; up.x = v.x; up.dinv = v.dinv;
  up.alpha = s_scal[0]; up.beta = s_scal[1];
  up.g = up.d = up.nn = 0.0;
  const double *xin = cur ? v.M0 : v.M1; /* m of the previous iteration (owned part + ghost tail) */
  if (ghost_cta) {
    if (tid < pd.n_nbr) p2p_wait(pd.hflags[pd.rank] + pd.nbr[tid], halo_seq, pd.seq + 2, pd.timeout_ns);
    asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
    if (GR > 1) spmv_consume_groups<S, NC, GR, IDX, true>(rowptr, colidx, vals, xin, vals_s, cols_s, rp_s, full0, empty0, t0, t1, up);
    else spmv_consume_emit<S, NC, IDX, SPLIT, true>(rowptr, colidx, vals, xin, vals_s, cols_s, rp_s, full0, empty0, t0, t1, up);
  } else {
    if (GR > 1) spmv_consume_groups<S, NC, GR, IDX, false>(rowptr, colidx, vals, xin, vals_s, cols_s, rp_s, full0, empty0, t0, t1, up);
    else spmv_consume_emit<S, NC, IDX, SPLIT, false>(rowptr, colidx, vals, xin, vals_s, cols_s, rp_s, full0, empty0, t0, t1, up);
  }
  /* ---- CTA partials (consumer warps only: named barrier 1) ---- */
  {
    double sum[3] = {up.g, up.d, up.nn};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int o = 16; o; o >>= 1) sum[k] += __shfl_xor_sync(0xffffffffu, sum[k], o);
      if (lane == 0) s_red[k * 32 + warp] = sum[k];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double t = lane < NC / 32 ? s_red[k * 32 + lane] : 0.0;
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) f.parts[(size_t)(f.rpar ^ 1) * 3 * f.G + 3 * blockIdx.x + k] = t;
      }
    }
  }
  if (tid == 0) trace_put(f.trace, TR_SPMV_CTA_DONE, me->its);
  if (MODE != OHP_MODE_P2P) return;
  /* ---- rank total for the other ranks: the last CTA to arrive adds the G partials and posts them.  Taken BEFORE the halo push:
   * the push's system-scope fence and the NVLink round trip (5-8 us in the boundary CTAs) then overlap the all-reduce ---- */
  if (warp == 0) {
    int last = 0;
    if (lane == 0) { __threadfence(); last = (atomicInc(f.cnt + 1, (unsigned)(f.G - 1)) == (unsigned)(f.G - 1)); }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      __threadfence();
      double t[3];
      pf_sum_parts(f.parts + (size_t)(f.rpar ^ 1) * 3 * f.G, f.G, lane, t);
      pf_post_total(pd, t, lane);
      if (lane == 0) trace_put(f.trace, TR_AR_POSTED, me->its);
    }
  }
  /* ---- halo: push w' (m') of the rows the neighbours ghost, publish when every contributor has arrived ---- */
  if (contrib) {
    /* a neighbour may still be READING the tail this push overwrites unless its own previous push has been seen */
    if (tid < pd.n_nbr) p2p_wait(pd.hflags[pd.rank] + pd.nbr[tid], halo_seq, pd.seq + 2, pd.timeout_ns);
    asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory"); /* also: every row of this CTA is stored */
    const double *src = up.mn;
    for (int j = f.hptr[blockIdx.x] + tid; j < f.hptr[blockIdx.x + 1]; j += NC) {
      const int i = f.hidx[j];
      const int l = pd.send_ldof[i];
      int q = 0;
      while (i >= pd.soff[q + 1]) ++q;
      pd.p[cur][pd.nbr[q]][pd.dst_off[q] + (i - pd.soff[q])] = ld_weak(src + l);
    }
    __threadfence_system();
    asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
    if (warp == 0) {
      int last = 0;
      if (lane == 0) last = (atomicInc(f.cnt, (unsigned)(f.n_contrib - 1)) == (unsigned)(f.n_contrib - 1));
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        unsigned long long sq = 0;
        if (lane == 0) { __threadfence(); sq = pd.seq[1] + 1; pd.seq[1] = sq; }
        sq = __shfl_sync(0xffffffffu, sq, 0);
        if (lane < pd.n_nbr) st_release_sys(pd.hflags[pd.nbr[lane]] + pd.rank, sq);
      }
    }
  }
}

/* after a k_pipecg_vec pass that left vgrid x 3 partials: their sum becomes "CTA 0's partial" of the parity the next
 * fused launch reads (the other CTAs' entries are zero), and the rank total goes to the peers' mailboxes */
template <int MODE>
__global__ void k_pipe_post(const double *vpart, int nvpart, ohp_pf_dev f, const ohp_p2p_dev pd) {
  pdl_wait();
  pdl_launch();
  if (f.st_host->done) return;
  const int lane = threadIdx.x;
  double t[3];
  pf_sum_parts(vpart, nvpart, lane, t);
  double *dst = f.parts + (size_t)f.rpar * 3 * f.G;
  for (int i = lane; i < 3 * f.G; i += 32) dst[i] = i < 3 ? t[i] : 0.0;
  if (MODE == OHP_MODE_P2P) pf_post_total(pd, t, lane);
}
/* the scalar step on its own (first launch of a replacement pass): what every CTA of k_pipecg_fused does at its start,
 * written to the host-visible state and to all replicas */
template <int MODE>
__global__ void k_pipe_scalars(ohp_pf_dev f, const ohp_p2p_dev pd) {
  pdl_wait();
  pdl_launch();
  __shared__ ohp_cg_state s_st;
  if (f.st_host->done) return;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    double t[3];
    pf_sum_parts(f.parts + (size_t)f.rpar * 3 * f.G, f.G, lane, t);
    if (MODE == OHP_MODE_P2P) pf_gather_totals(pd, t, lane);
    if (lane == 0) {
      cgcg_scalars(f.st_host, f.hist, f.hist_cap, t[0], t[1], f.jacobi ? t[2] : t[0]);
      s_st = *f.st_host;
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < f.G; b += blockDim.x) f.rep[b] = s_st;
}
__global__ void k_pipe_replicate(const ohp_cg_state *src, ohp_cg_state *rep, int G) {
  for (int b = threadIdx.x; b < G; b += blockDim.x) rep[b] = *src;
}

static int pf_plan_build(ohp_mesh m, bool p2p) {
  ohp_context ctx = m->ctx;
  cudaStream_t st = ctx->stream;
  ohp_pf_plan *pl = new ohp_pf_plan();
  const int G = std::max(1, std::min(m->n_tiles, ctx->num_sms));
  pl->G = G;
  std::vector<int32_t> hptr(G + 1, 0), hidx;
  std::vector<uint8_t> contrib(G, 0);
  if (p2p) {
    const ohp_p2p_dev &pd = m->p2p->dev;
    const int nsend = pd.soff[pd.n_nbr];
    std::vector<int32_t> tr(m->n_tiles + 1, 0), send(std::max(1, nsend));
    std::vector<uint8_t> gc(G, 0);
    if (m->n_tiles > 0) OHP_CUDA(cudaMemcpyAsync(tr.data(), m->tile_row, sizeof(int32_t) * (m->n_tiles + 1), cudaMemcpyDeviceToHost, st));
    if (nsend > 0) OHP_CUDA(cudaMemcpyAsync(send.data(), pd.send_ldof, sizeof(int32_t) * nsend, cudaMemcpyDeviceToHost, st));
    if (m->ghost_cta) OHP_CUDA(cudaMemcpyAsync(gc.data(), m->ghost_cta, G, cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
    std::vector<int32_t> row0(G + 1); /* CTA b owns the rows that start in its tiles */
    for (int b = 0; b <= G; ++b) row0[b] = tr[(int)(((int64_t)m->n_tiles * b) / G)];
    std::vector<int32_t> owner(std::max(1, nsend));
    for (int i = 0; i < nsend; ++i) {
      const int b = (int)(std::upper_bound(row0.begin(), row0.end(), send[i]) - row0.begin()) - 1;
      owner[i] = std::min(std::max(b, 0), G - 1);
      hptr[owner[i] + 1]++;
    }
    for (int b = 0; b < G; ++b) hptr[b + 1] += hptr[b];
    hidx.resize(std::max(1, nsend));
    std::vector<int32_t> cursor(hptr.begin(), hptr.end() - 1);
    for (int i = 0; i < nsend; ++i) hidx[cursor[owner[i]]++] = i;
    pl->n_contrib = 0;
    for (int b = 0; b < G; ++b) { contrib[b] = (hptr[b + 1] > hptr[b] || gc[b]) ? 1 : 0; pl->n_contrib += contrib[b]; }
    if (pl->n_contrib == 0) { contrib[0] = 1; pl->n_contrib = 1; } /* the flags still advance in lockstep */
  }
  if (hidx.empty()) hidx.push_back(0);
  OHP_CUDA(OHP_MALLOC(ctx, &pl->hptr, sizeof(int32_t) * (G + 1)));
  OHP_CUDA(OHP_MALLOC(ctx, &pl->hidx, sizeof(int32_t) * hidx.size()));
  OHP_CUDA(OHP_MALLOC(ctx, &pl->contrib, G));
  OHP_CUDA(OHP_MALLOC(ctx, &pl->rep, sizeof(ohp_cg_state) * G));
  OHP_CUDA(OHP_MALLOC(ctx, &pl->parts, sizeof(double) * 2 * 3 * G));
  OHP_CUDA(OHP_MALLOC(ctx, &pl->cnt, sizeof(unsigned) * 4));
  OHP_CUDA(cudaMemcpyAsync(pl->hptr, hptr.data(), sizeof(int32_t) * (G + 1), cudaMemcpyHostToDevice, st));
  OHP_CUDA(cudaMemcpyAsync(pl->hidx, hidx.data(), sizeof(int32_t) * hidx.size(), cudaMemcpyHostToDevice, st));
  OHP_CUDA(cudaMemcpyAsync(pl->contrib, contrib.data(), G, cudaMemcpyHostToDevice, st));
  OHP_CUDA(cudaMemsetAsync(pl->parts, 0, sizeof(double) * 2 * 3 * G, st));
  OHP_CUDA(cudaMemsetAsync(pl->cnt, 0, sizeof(unsigned) * 4, st));
  OHP_CUDA(cudaStreamSynchronize(st)); /* the host vectors above go out of scope */
  m->pfplan = pl;
  return 0;
}

template <int S, int NC, int GR, class IDX, bool SPLIT, bool JACOBI, int MODE>
static int pf_launch_t(ohp_mat A, const IDX *cols, const ohp_pipe_vecs &v, const ohp_pf_dev &f, const ohp_p2p_dev &pd, int cur, bool pdl) {
  ohp_mesh m = A->mesh;
  ohp_context ctx = m->ctx;
  constexpr int smem = tma_smem_bytes(S, (int)sizeof(IDX));
  static unsigned long long configured_devices = 0ull;
  if (!((configured_devices >> (ctx->device & 63)) & 1ull)) {
    OHP_CUDA(cudaFuncSetAttribute(k_pipecg_fused<S, NC, GR, IDX, SPLIT, JACOBI, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_devices |= 1ull << (ctx->device & 63);
  }
  OHP_CUDA(ohp_launch(k_pipecg_fused<S, NC, GR, IDX, SPLIT, JACOBI, MODE>, dim3(f.G), dim3(NC + 32), (size_t)smem, ctx->stream, pdl,
                      (const int32_t *)m->rowptr, cols, (const double *)A->vals, (const int32_t *)m->tile_row, m->n_tiles, v, f, pd, cur,
                      spmv_pin_tiles(m, f.G, (int)sizeof(IDX))));
  ctx->launches++;
  return 0;
}
template <bool JACOBI, int MODE>
static int pf_launch(ohp_mat A, const ohp_pipe_vecs &v, const ohp_pf_dev &f, const ohp_p2p_dev &pd, int cur, bool pdl) {
  ohp_mesh m = A->mesh;
  static const char *idx = getenv("OHP_IDX");
  static const char *grp = getenv("OHP_PF_GROUPS"); /* "2" / "3": consumer groups on different tiles (A/B runs) */
  const bool i16 = m->col16 && !(idx && idx[0] == '3');
  /* Consumer groups on different tiles: the number of stages must be a multiple of the number of groups, so that the
   * tile that used a stage before (i - S) belongs to the SAME group -- a group that waited for phase k of a stage's
   * `full` barrier while the barrier was still in phase k - 1 would pass on the parity of phase k - 2.
   * Measured (1.05 M rows): one group 29.0 us per iteration, two groups 30.7: once the row's vector operands are loaded
   * before the row is summed the tile loop runs at the L2 throughput limit (~10 TB/s), not at the per-tile latency
   * chain, so the default stays one group. */
  if (grp && grp[0] == '2' && i16) return pf_launch_t<4, 576, 2, int16_t, false, JACOBI, MODE>(A, m->col16, v, f, pd, cur, pdl);
  if (grp && grp[0] == '3') {
    if (i16) return pf_launch_t<3, 576, 3, int16_t, false, JACOBI, MODE>(A, m->col16, v, f, pd, cur, pdl);
    return pf_launch_t<3, 576, 3, int32_t, false, JACOBI, MODE>(A, m->colidx, v, f, pd, cur, pdl);
  }
  const bool split = m->nnz > (int64_t)8 * std::max(1, m->n_owned);
  if (i16) return split ? pf_launch_t<3, 608, 1, int16_t, true, JACOBI, MODE>(A, m->col16, v, f, pd, cur, pdl)
                        : pf_launch_t<3, 608, 1, int16_t, false, JACOBI, MODE>(A, m->col16, v, f, pd, cur, pdl);
  return split ? pf_launch_t<3, 608, 1, int32_t, true, JACOBI, MODE>(A, m->colidx, v, f, pd, cur, pdl)
               : pf_launch_t<3, 608, 1, int32_t, false, JACOBI, MODE>(A, m->colidx, v, f, pd, cur, pdl);
}

static int pipefused_solve(ohp_mat A, ohp_vec b, ohp_vec x, const ohp_ksp_options *o, ohp_ksp_result *res, bool p2p) {
  ohp_mesh m = A->mesh;
  ohp_context ctx = m->ctx;
  cudaStream_t st = ctx->stream;
  const int n = m->n_owned;
  const size_t len = (size_t)std::max(1, n + m->n_ghost);
  const bool jac = (o->pc == OHP_PC_JACOBI);
  if (!(spmv_use_tma(m) || m->n_tiles == 0) || getenv("OHP_SPMV")) OHP_FAIL(OHP_ERR_SUP, "ohp_ksp_cg (pipelined, fused): needs the TMA-pipelined SpMV");
  if (!m->pfplan) OHP_TRY(pf_plan_build(m, p2p));
  ohp_pf_plan *pl = m->pfplan;
  auto need = [&](double **ptr) -> int {
    if (!*ptr) { OHP_CUDA(OHP_MALLOC(ctx, ptr, sizeof(double) * len)); OHP_CUDA(cudaMemsetAsync(*ptr, 0, sizeof(double) * len, st)); }
    return 0;
  };
  OHP_TRY(need(&A->p)); OHP_TRY(need(&A->w)); OHP_TRY(need(&A->s)); OHP_TRY(need(&A->r)); OHP_TRY(need(&A->z));
  static const ohp_p2p_dev no_p2p = {};
  const ohp_p2p_dev &pd = p2p ? m->p2p->dev : no_p2p;
  ohp_pipe_vecs v;
  memset(&v, 0, sizeof(v));
  if (p2p) { v.M0 = pd.p[0][ctx->rank]; v.M1 = pd.p[1][ctx->rank]; }
  else { OHP_TRY(need(&A->pm0)); OHP_TRY(need(&A->pm1)); v.M0 = A->pm0; v.M1 = A->pm1; }
  if (jac) {
    OHP_TRY(need(&A->pw0)); OHP_TRY(need(&A->pw1)); OHP_TRY(need(&A->dinv));
    v.W0 = A->pw0; v.W1 = A->pw1;
    OHP_TRY(ohp_extract_dinv(A));
  } else { v.W0 = v.M0; v.W1 = v.M1; }
  v.Z0 = v.Z1 = A->z; /* z is updated in place here (row-local) */
  v.r = A->r; v.s = A->s; v.p = A->p; v.x = x->d; v.nb = A->w; v.b = b->d; v.dinv = A->dinv;
  {
    int G_, st_, ch_;
    dyn_sched_params(m, ctx->num_sms, &G_, &st_, &ch_);
    v.tcnt = (getenv("OHP_DYN") && getenv("OHP_DYN")[0] == '1') ? ctx->d_ticket + 3 : nullptr;
    v.tinit = (unsigned)(G_ * st_);
  }
  double *hist = nullptr;
  int hist_cap = 0;
  if (o->history && o->history_cap > 0) {
    if (A->hist_cap < o->history_cap) {
      OHP_FREE(ctx, A->hist);
      OHP_CUDA(OHP_MALLOC(ctx, &A->hist, sizeof(double) * o->history_cap));
      A->hist_cap = o->history_cap;
    }
    hist = A->hist; hist_cap = o->history_cap;
    OHP_CUDA(cudaMemsetAsync(hist, 0, sizeof(double) * hist_cap, st));
  }
  ohp_cg_state init;
  memset(&init, 0, sizeof(init));
  init.rtol = o->rtol; init.abstol = o->abstol; init.dtol = o->dtol; init.maxit = o->max_it;
  *ctx->h_cg = init;
  OHP_CUDA(cudaMemcpyAsync(ctx->d_cg, ctx->h_cg, sizeof(init), cudaMemcpyHostToDevice, st));
  k_pipe_replicate<<<1, 256, 0, st>>>(ctx->d_cg, pl->rep, pl->G); OHP_CHECK_LAUNCH(ctx);
  const int mode = p2p ? OHP_MODE_P2P : OHP_MODE_SERIAL;
  const int vgrid = (int)std::min<int64_t>(std::min<int64_t>(nblk(n, 256), (int64_t)ctx->num_sms * 5), kMaxPartials / 3);
  const int nhalo = p2p ? std::max(1, std::min(16, (pd.soff[pd.n_nbr] + 255) / 256)) : 0;
  double *vpart = ctx->d_partials + kMaxPartials;
  ohp_cgcg_epi epi;
  memset(&epi, 0, sizeof(epi));
  epi.vpart = vpart; epi.nvpart = vgrid; epi.jacobi = jac ? 1 : 0; epi.hist = hist; epi.hist_cap = hist_cap; epi.pipe = 2;
  epi.trace = ohp_cg_trace_open(ctx);
  if (p2p) OHP_TRY(ohp_p2p_entry_barrier(m));
  static const char *nooverlap = getenv("OHP_NOOVERLAP");
  const bool overlap = p2p && m->ghost_cta && !nooverlap;
  const int wait_in_vec = overlap ? 0 : 1;
  epi.ghost_cta = overlap ? (m->cta_tile0 ? m->ghost_cta_w : m->ghost_cta) : nullptr;
  epi.cta_tile0 = overlap ? m->cta_tile0 : nullptr;
  static const char *nopdl = getenv("OHP_NOPDL");
  const bool pdl = !nopdl;
  static const char *rr_env = getenv("OHP_PIPE_RR");
  const int rr = rr_env ? atoi(rr_env) : kPipeReplaceEvery;
  const int check_every = o->check_every > 0 ? o->check_every : 50;
  ohp_pf_dev f;
  memset(&f, 0, sizeof(f));
  f.G = pl->G; f.n_contrib = pl->n_contrib; f.hptr = pl->hptr; f.hidx = pl->hidx; f.contrib = pl->contrib;
  f.ghost_cta = p2p ? m->ghost_cta : nullptr; /* without the flags table every CTA of a partitioned run would have to wait: see below */
  f.rep = pl->rep; f.st_host = ctx->d_cg; f.parts = pl->parts; f.cnt = pl->cnt; f.hist = hist; f.hist_cap = hist_cap; f.jacobi = jac ? 1 : 0;
  f.trace = epi.trace;
  if (p2p && !m->ghost_cta && m->n_tiles > 0) OHP_FAIL(OHP_ERR_LIB, "ohp_ksp_cg (pipelined, fused): ghost-column table missing");
  EventBag bag, profbag;
  cudaEvent_t ev0, ev1;
  OHP_CUDA(bag.record(st, &ev0));
  std::vector<cudaEvent_t> &prof = profbag.ev;
  const int nprof = std::max(0, o->profile);
  int pushes = 0, rpar = 0; /* next push goes to buffer pushes & 1; the next consumer of partials reads parity rpar */
  auto vec = [&](int kind, int cur) -> int {
    if (p2p) {
      if (jac) OHP_CUDA(ohp_launch(k_pipecg_vec<OHP_MODE_P2P, true>, dim3(vgrid + nhalo), dim3(256), 0, st, pdl, v, n, (const ohp_cg_state *)ctx->d_cg, vpart, pd, nhalo, ctx->d_ticket + 2, wait_in_vec, epi.trace, kind, cur));
      else OHP_CUDA(ohp_launch(k_pipecg_vec<OHP_MODE_P2P, false>, dim3(vgrid + nhalo), dim3(256), 0, st, pdl, v, n, (const ohp_cg_state *)ctx->d_cg, vpart, pd, nhalo, ctx->d_ticket + 2, wait_in_vec, epi.trace, kind, cur));
    } else {
      if (jac) OHP_CUDA(ohp_launch(k_pipecg_vec<OHP_MODE_SERIAL, true>, dim3(vgrid), dim3(256), 0, st, pdl, v, n, (const ohp_cg_state *)ctx->d_cg, vpart, pd, 0, (unsigned *)nullptr, 1, epi.trace, kind, cur));
      else OHP_CUDA(ohp_launch(k_pipecg_vec<OHP_MODE_SERIAL, false>, dim3(vgrid), dim3(256), 0, st, pdl, v, n, (const ohp_cg_state *)ctx->d_cg, vpart, pd, 0, (unsigned *)nullptr, 1, epi.trace, kind, cur));
    }
    ctx->launches++;
    return 0;
  };
  auto stage = [&](int kind) -> int { const int cur = pushes & 1; OHP_TRY(vec(kind, cur)); pushes++; return 0; };
  auto stage_mult = [&](int kind, double *out) -> int {
    const int cur = pushes & 1;
    OHP_TRY(stage(kind));
    OHP_TRY(spmv_enqueue(A, cur ? v.M1 : v.M0, out, ctx->d_cg, ctx->d_scalars, mode, nullptr, pd, epi));
    return 0;
  };
  auto post = [&]() -> int { /* vpart -> parts[rpar] (+ mailboxes) for the next consumer */
    f.rpar = rpar;
    if (p2p) OHP_CUDA(ohp_launch(k_pipe_post<OHP_MODE_P2P>, dim3(1), dim3(32), 0, st, pdl, (const double *)vpart, vgrid, f, pd));
    else OHP_CUDA(ohp_launch(k_pipe_post<OHP_MODE_SERIAL>, dim3(1), dim3(32), 0, st, pdl, (const double *)vpart, vgrid, f, pd));
    ctx->launches++;
    return 0;
  };
  /* set-up: r = b, u, m = u; w = A u; partials of (r,u), (w,u), (u,u) */
  OHP_TRY(stage_mult(PK_INIT, A->w));
  OHP_TRY(stage(PK_TAKE_W));
  OHP_TRY(post());
  int iter = 0; /* == the device's iteration count while it is not done */
  for (;;) {
    for (int it = 0; it < check_every; ++it, ++iter) {
      f.rpar = rpar;
      if (rr > 0 && iter > 0 && iter % rr == 0) {
        /* residual replacement in place of iteration `iter`'s recurrences */
        if (p2p) OHP_CUDA(ohp_launch(k_pipe_scalars<OHP_MODE_P2P>, dim3(1), dim3(256), 0, st, pdl, f, pd));
        else OHP_CUDA(ohp_launch(k_pipe_scalars<OHP_MODE_SERIAL>, dim3(1), dim3(256), 0, st, pdl, f, pd));
        ctx->launches++;
        OHP_TRY(vec(PK_PX, pushes & 1));
        OHP_TRY(stage_mult(PK_STAGE_P, A->s));   /* s = A p */
        OHP_TRY(stage_mult(PK_STAGE_S, A->z));   /* z = A M^-1 s */
        OHP_TRY(stage_mult(PK_STAGE_X, A->w));   /* A x */
        OHP_TRY(stage_mult(PK_NEW_R, A->w));     /* r = b - A x; A M^-1 r */
        OHP_TRY(stage(PK_TAKE_W));               /* w, partials, m = M^-1 w */
        OHP_TRY(post());                         /* (rpar unchanged: the scalar step consumed it, the post refills it) */
      } else {
        const bool timed = (int)prof.size() < 2 * nprof && (iter >= check_every || o->max_it <= check_every);
        const int cur = pushes & 1;
        if (timed) OHP_CUDA(profbag.record(st));
        if (p2p) { if (jac) OHP_TRY((pf_launch<true, OHP_MODE_P2P>(A, v, f, pd, cur, pdl))); else OHP_TRY((pf_launch<false, OHP_MODE_P2P>(A, v, f, pd, cur, pdl))); }
        else { if (jac) OHP_TRY((pf_launch<true, OHP_MODE_SERIAL>(A, v, f, pd, cur, pdl))); else OHP_TRY((pf_launch<false, OHP_MODE_SERIAL>(A, v, f, pd, cur, pdl))); }
        if (timed) OHP_CUDA(profbag.record(st));
        pushes++;
        rpar ^= 1;
      }
    }
    OHP_CUDA(cudaMemcpyAsync(ctx->h_cg, ctx->d_cg, sizeof(ohp_cg_state), cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
    if (ctx->h_cg->done) break;
    if (iter > o->max_it + 3 * check_every) OHP_FAIL(OHP_ERR_LIB, "ohp_ksp_cg (pipelined, fused): iteration control did not terminate");
  }
  OHP_CUDA(bag.record(st, &ev1));
  OHP_CUDA(cudaEventSynchronize(ev1));
  res->total_ms = 0.f;
  cudaEventElapsedTime(&res->total_ms, ev0, ev1);
  res->its = ctx->h_cg->its; res->reason = ctx->h_cg->reason; res->rnorm = ctx->h_cg->dp; res->rnorm0 = ctx->h_cg->rnorm0;
  res->spmv_ms = 0.f; res->spmv_samples = 0;
  ohp_cg_trace_dump(ctx, "pipecg-fused");
  res->algo = OHP_ALGO_PIPELINED_FUSED;
  {
    double acc = 0.0;
    int used = 0, L = (o->max_it <= check_every) ? 0 : check_every;
    for (size_t i = 0; i + 1 < prof.size(); i += 2, ++L) {
      while (rr > 0 && L > 0 && L % rr == 0) ++L;
      float ms = 0.f;
      cudaEventElapsedTime(&ms, prof[i], prof[i + 1]);
      if (L < res->its) { acc += ms; used++; }
    }
    if (used) { res->spmv_ms = (float)(acc / used); res->spmv_samples = used; }
  }
  if (p2p) OHP_TRY(ohp_p2p_check(m, "ohp_ksp_cg (pipelined, fused)"));
  if (hist) {
    const int ncopy = std::min(hist_cap, res->its + 1);
    OHP_CUDA(cudaMemcpyAsync(o->history, hist, sizeof(double) * ncopy, cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}

extern "C" int ohp_ksp_cg(ohp_mat A, ohp_vec b, ohp_vec x, const ohp_ksp_options *o, ohp_ksp_result *res) {
  if (!A || !b || !x || !o || !res) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_ksp_cg: null argument");
  ohp_mesh m = A->mesh;
  ohp_context ctx = m->ctx;
  const int n = m->n_owned;
  if (b->n != n || x->n != n) OHP_FAIL(OHP_ERR_ARG_SIZ, "ohp_ksp_cg: vector sizes do not match the matrix");
  if (b == x) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_ksp_cg: b and x must differ");
  if (o->pc != OHP_PC_NONE && o->pc != OHP_PC_JACOBI && o->pc != OHP_PC_CHEBYSHEV) OHP_FAIL(OHP_ERR_SUP, "ohp_ksp_cg: pc %d", o->pc);
  OHP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  res->cheb_emin = res->cheb_emax = 0.0; res->spmvs = 0; res->reserved0 = 0;
  /* the polynomial preconditioner and the null-space projection run in the general loop of ohp_pcg.cu */
  if (o->pc == OHP_PC_CHEBYSHEV || A->nullspace) return ohp_pcg_solve(A, b, x, o, res);
  {
    /* algorithm selection (the role of -ksp_type among PETSc's CG variants): the three-launch classic
     * recurrence moves the fewest bytes and wins when the per-GPU problem is large; the persistent
     * single-reduction kernel wins when kernel boundaries and reductions dominate (partitioned runs
     * over NVLink peer memory, small meshes).  OHP_CG=classic|persistent overrides. */
    static const char *force = getenv("OHP_CG");
    static const char *force_comm = getenv("OHP_COMM");
    const bool can_p2p = ctx->nranks > 1 && m->p2p && m->p2p->attached && !(force_comm && force_comm[0] == 'n');
    const bool feasible = spmv_use_tma(m) && n > 0 && (ctx->nranks == 1 || can_p2p);
    bool persistent = feasible && ctx->nranks == 1 && n < kPersistentBelowRows; /* measured: see DESIGN.md section 5 */
    if (force && force[0] == 'c') persistent = false; /* classic or cgcg */
    if (force && force[0] == 'p' && force[1] == 'e') persistent = feasible;
    /* OHP_CG=fused: the one-synchronisation persistent kernel (ohp_cg_fused.cu) */
    if (force && force[0] == 'f' && feasible) return ohp_cg_fused_solve(A, b, x, o, res, can_p2p);
    /* OHP_CG=pipecg: the pipelined recurrence (reduction and all-reduce overlapped with the SpMV) */
    const bool feasible_pipe = (spmv_use_tma(m) || m->n_tiles == 0) && (ctx->nranks == 1 ? n > 0 : can_p2p);
    if (o->ksp_type == OHP_KSP_PIPECG && n == 0 && ctx->nranks == 1) { res->its = 0; res->reason = 3; res->rnorm = res->rnorm0 = 0.0; res->spmv_ms = res->total_ms = 0.f; res->spmv_samples = 0; return 0; }
    if (o->ksp_type == OHP_KSP_PIPECG && !feasible_pipe) OHP_FAIL(OHP_ERR_SUP, "ohp_ksp_cg: -ksp_type pipecg needs the TMA SpMV and, on a partitioned run, the NVLink peer transport");
    /* the pipelined recurrence has two forms: one launch per iteration (vector update fused into the SpMV: fewest bytes and
     * kernel boundaries, but the dot products are needed at the start of the next launch, so across ranks the all-reduce is
     * exposed) and two launches (the all-reduce hides behind the SpMV).  Measured on the 16 M-triangle box: one GPU at
     * 1.05 M rows 29.0 vs 32.5 us per iteration; 8 GPUs 49.5 vs 38.4.  -ksp_type pipecg picks by the number of ranks;
     * OHP_CG=pipecg1 / pipecg2 force a form (A/B runs). */
    const bool want_pipe = o->ksp_type == OHP_KSP_PIPECG || (force && force[0] == 'p' && force[1] == 'i');
    if (want_pipe && feasible_pipe) {
      bool one_launch = ctx->nranks == 1;
      if (force && !strcmp(force, "pipecg1")) one_launch = true;
      if (force && !strcmp(force, "pipecg2")) one_launch = false;
      return one_launch ? pipefused_solve(A, b, x, o, res, can_p2p) : pipecg_solve(A, b, x, o, res, can_p2p);
    }
    if (persistent) return ohp_cg_persistent_solve(A, b, x, o, res, can_p2p);
    /* single-reduction recurrence in two launches: default for partitioned runs over NVLink peer memory (one
     * all-reduce and one kernel boundary fewer per iteration); OHP_CG=cgcg forces it on one GPU */
    /* every rank must take the same branch: decide on GLOBAL quantities only (a rank may own no rows at all) */
    const int64_t rows_per_rank = m->n_global / std::max(1, ctx->nranks);
    /* strong-scaling regime (measured at 1.05 M rows per rank, 8 GPUs: classic 3 launches 58, single-reduction 48.8,
     * pipelined two-launch 38.4 us per iteration): reductions and rank skew must be off the critical path */
    /* ... and only where it was measured to win: below kPipelinedAboveRows per rank (the reference's own 24k-element XGC
     * fixtures) the single-reduction recurrence stays -- it keeps CG's iterates to rounding, while the pipelined recurrence
     * takes 4-7 % more iterations on that ill-conditioned mesh (535 against 505 at 2 GPUs) */
    bool pipe2 = can_p2p && m->all_tma && rows_per_rank < kPipelinedBelowRows && rows_per_rank >= kPipelinedAboveRows;
    if (force && force[0] == 'c') pipe2 = false;
    if (pipe2) return pipecg_solve(A, b, x, o, res, can_p2p);
    bool cgcg = can_p2p && rows_per_rank < kSingleReductionBelowRows; /* measured crossover: 0.145 vs 0.141 ms at 4.2 M rows/rank, 0.056 vs 0.068 at 1.05 M */
    if (force && force[0] == 'c' && force[1] == 'l') cgcg = false;                          /* classic */
    if (force && force[0] == 'c' && force[1] == 'g') cgcg = (ctx->nranks == 1 || can_p2p);  /* cgcg */
    if (cgcg) return cgcg_solve(A, b, x, o, res, can_p2p);
  }
  const size_t len = (size_t)std::max(1, n + m->n_ghost);
  if (!A->r) {
    OHP_CUDA(OHP_MALLOC(ctx, &A->r, sizeof(double) * len));
    OHP_CUDA(OHP_MALLOC(ctx, &A->p, sizeof(double) * len));
    OHP_CUDA(OHP_MALLOC(ctx, &A->p1, sizeof(double) * len));
    OHP_CUDA(OHP_MALLOC(ctx, &A->w, sizeof(double) * len));
    OHP_CUDA(cudaMemsetAsync(A->p, 0, sizeof(double) * len, st));
    OHP_CUDA(cudaMemsetAsync(A->p1, 0, sizeof(double) * len, st));
  }
  /* transport: serial | NCCL (pack + send/recv + all-reduce + scalar kernel) | NVLink P2P fused into the kernels */
  static const char *force_nccl = getenv("OHP_COMM"); /* "nccl" disables the P2P path (A/B runs) */
  const bool p2p = ctx->nranks > 1 && m->p2p && m->p2p->attached && !(force_nccl && force_nccl[0] == 'n');
  const int mode = ctx->nranks == 1 ? OHP_MODE_SERIAL : (p2p ? OHP_MODE_P2P : OHP_MODE_NCCL);
  static const ohp_p2p_dev no_p2p = {};
  const ohp_p2p_dev &pd = p2p ? m->p2p->dev : no_p2p;
  double *P0 = p2p ? pd.p[0][ctx->rank] : A->p, *P1 = p2p ? pd.p[1][ctx->rank] : A->p1;
  const bool jac = (o->pc == OHP_PC_JACOBI);
  if (jac) {
    if (!A->z) OHP_CUDA(OHP_MALLOC(ctx, &A->z, sizeof(double) * len));
    if (!A->dinv) OHP_CUDA(OHP_MALLOC(ctx, &A->dinv, sizeof(double) * len)); /* another recurrence may have made it already */
    k_extract_dinv<<<nblk(n, 256), 256, 0, st>>>(m->rowptr, m->colidx, A->vals, n, A->dinv);
    OHP_CHECK_LAUNCH(ctx);
  }
  double *z = jac ? A->z : A->r;
  double *hist = nullptr;
  int hist_cap = 0;
  if (o->history && o->history_cap > 0) {
    if (A->hist_cap < o->history_cap) {
      OHP_FREE(ctx, A->hist);
      OHP_CUDA(OHP_MALLOC(ctx, &A->hist, sizeof(double) * o->history_cap));
      A->hist_cap = o->history_cap;
    }
    hist = A->hist; hist_cap = o->history_cap;
    OHP_CUDA(cudaMemsetAsync(hist, 0, sizeof(double) * hist_cap, st));
  }
  /* scalar block */
  ohp_cg_state init;
  memset(&init, 0, sizeof(init));
  init.rtol = o->rtol; init.abstol = o->abstol; init.dtol = o->dtol; init.maxit = o->max_it;
  *ctx->h_cg = init;
  OHP_CUDA(cudaMemcpyAsync(ctx->d_cg, ctx->h_cg, sizeof(init), cudaMemcpyHostToDevice, st));
  const bool nccl = (mode == OHP_MODE_NCCL);
  static const char *nopdl = getenv("OHP_NOPDL"); /* A/B switch for programmatic dependent launch */
  const bool pdl = !nccl && !nopdl;
  /* with the persistent TMA SpMV the halo wait moves into the SpMV's boundary CTAs (overlap); otherwise it stays here */
  static const char *nooverlap = getenv("OHP_NOOVERLAP");
  const bool overlap = p2p && spmv_use_tma(m) && m->ghost_cta && !nooverlap && !getenv("OHP_SPMV");
  const int wait_in_vec = overlap ? 0 : 1;
  ohp_cgcg_epi cepi;
  memset(&cepi, 0, sizeof(cepi));
  cepi.ghost_cta = overlap ? (m->cta_tile0 ? m->ghost_cta_w : m->ghost_cta) : nullptr;
  cepi.cta_tile0 = overlap ? m->cta_tile0 : nullptr;
  cepi.trace = ohp_cg_trace_open(ctx);
  if (p2p) OHP_TRY(ohp_p2p_entry_barrier(m));
  const int nhalo = p2p ? std::max(1, std::min(16, (pd.soff[pd.n_nbr] + 255) / 256)) : 0;
  const int vgrid = vec_grid(ctx, n);
  double *sums = ctx->d_scalars;

  if (jac) k_cg_init<true><<<vgrid, 256, 0, st>>>(b->d, x->d, A->r, z, A->dinv, n, ctx->d_cg, ctx->d_partials, ctx->d_ticket, sums, hist, mode, pd);
  else k_cg_init<false><<<vgrid, 256, 0, st>>>(b->d, x->d, A->r, z, A->dinv, n, ctx->d_cg, ctx->d_partials, ctx->d_ticket, sums, hist, mode, pd);
  OHP_CHECK_LAUNCH(ctx);
  if (nccl) {
    OHP_TRY(ohp_allreduce_sum(ctx, sums, 2));
    k_cg_fin<<<1, 1, 0, st>>>(ctx->d_cg, sums, 0, hist, hist_cap); OHP_CHECK_LAUNCH(ctx);
  }
  const int check_every = o->check_every > 0 ? o->check_every : 50;
  int launched = 0;
  /* optional per-launch timing of the dominant kernel, skipping the first (cold) chunk */
  EventBag bag, profbag;
  std::vector<cudaEvent_t> &prof = profbag.ev;
  const int nprof = std::max(0, o->profile);
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  OHP_CUDA(bag.record(st, &ev_t0));
  for (;;) {
    OHP_CUDA(cudaMemcpyAsync(ctx->h_cg, ctx->d_cg, sizeof(ohp_cg_state), cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
    if (ctx->h_cg->done) break;
    if (launched > o->max_it + check_every) OHP_FAIL(OHP_ERR_LIB, "ohp_ksp_cg: device-side iteration control did not terminate");
    for (int it = 0; it < check_every; ++it) {
      if (p2p) OHP_CUDA(ohp_launch(k_cg_pupdate<OHP_MODE_P2P>, dim3(vgrid + nhalo), dim3(256), 0, st, pdl, P0, P1, z, x->d, n, ctx->d_cg, pd, nhalo, ctx->d_ticket + 2, wait_in_vec));
      else OHP_CUDA(ohp_launch(k_cg_pupdate<OHP_MODE_SERIAL>, dim3(vgrid), dim3(256), 0, st, pdl, P0, P1, z, x->d, n, ctx->d_cg, pd, 0, (unsigned *)nullptr, 1));
      OHP_CHECK_LAUNCH(ctx);
      if (nccl) {
        /* the host does not know the iteration parity on the device: exchange both buffers' tails is
         * wasteful, so the NCCL path tracks parity on the host (it never runs ahead of `done`) */
        OHP_TRY(ohp_halo_exchange(m, ((launched + it) & 1) ? P1 : P0));
      }
      const bool timed = (int)prof.size() < 2 * nprof && (launched > 0 || o->max_it <= check_every);
      if (timed) OHP_CUDA(profbag.record(st));
      OHP_TRY(spmv_enqueue(A, P0, A->w, ctx->d_cg, sums, mode, P1, pd, cepi));
      if (timed) OHP_CUDA(profbag.record(st));
      if (nccl) {
        OHP_TRY(ohp_allreduce_sum(ctx, sums, 1));
        k_cg_fin<<<1, 1, 0, st>>>(ctx->d_cg, sums, 1, hist, hist_cap); OHP_CHECK_LAUNCH(ctx);
      }
      if (jac) OHP_CUDA(ohp_launch(k_cg_update<true>, dim3(vgrid), dim3(256), 0, st, pdl, A->r, z, A->w, A->dinv, n, ctx->d_cg, ctx->d_partials, ctx->d_ticket, sums, hist, hist_cap, mode, pd));
      else OHP_CUDA(ohp_launch(k_cg_update<false>, dim3(vgrid), dim3(256), 0, st, pdl, A->r, z, A->w, A->dinv, n, ctx->d_cg, ctx->d_partials, ctx->d_ticket, sums, hist, hist_cap, mode, pd));
      OHP_CHECK_LAUNCH(ctx);
      if (nccl) {
        OHP_TRY(ohp_allreduce_sum(ctx, sums, 2));
        k_cg_fin<<<1, 1, 0, st>>>(ctx->d_cg, sums, 2, hist, hist_cap); OHP_CHECK_LAUNCH(ctx);
      }
    }
    launched += check_every;
  }
  k_cg_final_x<<<vgrid, 256, 0, st>>>(x->d, P0, P1, n, ctx->d_cg); OHP_CHECK_LAUNCH(ctx);
  res->its = ctx->h_cg->its; res->reason = ctx->h_cg->reason; res->rnorm = ctx->h_cg->dp; res->rnorm0 = ctx->h_cg->rnorm0;
  if (p2p) OHP_TRY(ohp_p2p_check(m, "ohp_ksp_cg"));
  OHP_CUDA(bag.record(st, &ev_t1));
  OHP_CUDA(cudaEventSynchronize(ev_t1));
  res->total_ms = 0.f;
  cudaEventElapsedTime(&res->total_ms, ev_t0, ev_t1);
  res->spmv_ms = 0.f; res->spmv_samples = 0;
  ohp_cg_trace_dump(ctx, "classic");
  res->algo = OHP_ALGO_CLASSIC;
  {
    /* only launches that ran before convergence did real work: iteration k is sample k - first_timed */
    double acc = 0.0;
    int used = 0;
    const int first_timed_it = (o->max_it <= check_every) ? 0 : check_every;
    for (size_t i = 0; i + 1 < prof.size(); i += 2) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, prof[i], prof[i + 1]);
      if (first_timed_it + (int)(i / 2) < res->its) { acc += ms; used++; }
    }
    if (used) { res->spmv_ms = (float)(acc / used); res->spmv_samples = used; }
  }
  if (hist) {
    const int ncopy = std::min(hist_cap, res->its + 1);
    OHP_CUDA(cudaMemcpyAsync(o->history, hist, sizeof(double) * ncopy, cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}
