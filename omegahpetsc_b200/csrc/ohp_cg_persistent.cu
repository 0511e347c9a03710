/*
 * ohp_cg_persistent.cu -- the whole CG iteration in ONE persistent cooperative kernel.
 *
 * Why: with three launches per iteration the fixed cost of the kernel boundaries (drain, launch,
 * ramp: ~14 us per iteration on B200) and of two dot-product reductions dominates as soon as the
 * per-GPU problem is small -- the strong-scaling regime of the partitioned runs and the 24k-triangle
 * XGC meshes.  This kernel keeps one CTA per SM resident for a whole chunk of iterations and uses
 *   - the single-reduction CG recurrence of Chronopoulos & Gear (one fused reduction per iteration
 *     instead of two; mathematically the same iterates as KSPSolve_CG, selected like PETSc's
 *     communication-reducing variants would be with -ksp_type),
 *   - two grid barriers per iteration (arrive counter + generation flag in HBM); the LAST CTA to
 *     arrive is the leader of that barrier and does the serial work before releasing the grid:
 *       barrier A (after the vector update): pushes the halo of the new direction into the
 *         neighbours' vectors over NVLink and waits for theirs,
 *       barrier B (after the SpMV): sums the per-CTA partials in a fixed order, all-reduces them
 *         through the peer mailboxes, advances alpha/beta/norm/iteration count/converged reason,
 *   - the same warp-specialised TMA pipeline as k_spmv_tma for the matrix stream; the producer warp
 *     never takes part in the barriers, so it prefetches the next iteration's first tiles while the
 *     consumers are still in the vector update.
 * Vectors rewritten inside the kernel are read with plain (coherent) loads; the acquire at each
 * barrier invalidates L1, which is what a kernel boundary did in the three-launch version.
 *
 * Per iteration (u = M^-1 r, w = A u, gamma = (r,u), delta = (w,u)):
 *   beta = gamma/gamma_old; alpha = gamma / (delta - beta*gamma/alpha_old)
 *   p = u + beta p; s = w + beta s; x += alpha p; r -= alpha s        (one pass, 9 vector streams)
 * Convergence test = KSPConvergedDefault on ||u|| (preconditioned norm), as in ohp_ksp_cg's classic path.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "ohp_dev.cuh"
#include "ohp_internal.h"

namespace {

constexpr int PS = 3;            /* TMA stages */
constexpr int PNC = 608;         /* consumer threads (19 warps) + 1 producer warp */

struct PcgArgs {
  const int32_t *rowptr, *colidx, *tile_row;
  const double *vals, *b, *dinv;
  double *x, *r, *u, *p, *s, *w;   /* u == r when there is no preconditioner */
  ohp_cg_state *st;
  double *partials;                /* gridDim.x * 3 */
  unsigned *sync;                  /* [0] arrive counter, [1] release generation (zeroed before every launch) */
  double *hist;
  int hist_cap, n, n_tiles, chunk_iters, mode, jacobi;
  ohp_p2p_dev pd;
};

__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(PNC) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

/* consumers only.  Returns true in every consumer thread of the last CTA to arrive. */
__device__ __forceinline__ bool gbar_arrive(unsigned *count, unsigned target, bool *s_flag) {
  cons_sync();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned old = atomicAdd(count, 1u);
    *s_flag = (old + 1u == target);
    __threadfence();
  }
  cons_sync();
  return *s_flag;
}
__device__ __forceinline__ void gbar_release(unsigned *gen, unsigned value) {
  cons_sync();
  if (threadIdx.x == 0) { __threadfence(); st_release_gpu(gen, value); }
}
__device__ __forceinline__ void gbar_wait(const unsigned *gen, unsigned value) {
  if (threadIdx.x == 0) {
    while (ld_acquire_gpu(gen) < value) { }
    __threadfence();
  }
  cons_sync();
}

/* sum of `NV` values per consumer thread over the CTA, fixed tree; result valid in thread 0 */
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
      double t = lane < PNC / 32 ? s_red[k * 32 + lane] : 0.0;
#pragma unroll
      for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      v[k] = t;
    }
  }
  cons_sync();
}

/* sum_{k in [a,b)} vs[k] * x[cs[k]] with coherent gathers (x is rewritten inside this kernel) */
__device__ __forceinline__ double row_dot_weak(const double *vs, const int32_t *cs, int a, int b, const double *x, double s) {
  for (int k = a; k < b; k += 8) {
    double xv[8], vv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool ok = (k + j) < b;
      const int c = ok ? cs[k + j] : 0;
      vv[j] = ok ? vs[k + j] : 0.0;
      xv[j] = ok ? ld_weak(x + c) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if ((k + j) < b) s += vv[j] * xv[j];
  }
  return s;
}

__global__ void __launch_bounds__(PNC + 32, 1) k_cg_persistent(const PcgArgs a) {
  constexpr int T = kSpmvTile, S = PS, NC = PNC;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *vals_s = reinterpret_cast<double *>(smem_raw);
  int32_t *cols_s = reinterpret_cast<int32_t *>(vals_s + S * T);
  int32_t *rp_s = cols_s + S * T;
  uint64_t *bars = reinterpret_cast<uint64_t *>(rp_s + S * kTmaRpStride);
  __shared__ double s_red[3 * 32];
  __shared__ bool s_flag;
  __shared__ volatile int s_exit, s_consumed;
  __shared__ unsigned long long s_seq;
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 33); mbar_init(empty0 + 8 * s, NC / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_exit = 0; s_consumed = 0;
  }
  __syncthreads();
  const int t0 = (int)(((int64_t)a.n_tiles * blockIdx.x) / G);
  const int t1 = (int)(((int64_t)a.n_tiles * (blockIdx.x + 1)) / G);
  const int ntile = t1 - t0;

  if (warp == NC / 32) {
    /* ------------- producer warp: streams the CTA's tiles, iteration after iteration ------------- */
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
          mbar_expect_tx(full0 + 8 * s, T * 12);
          tma_load_1d(smem_u32(vals_s + s * T), a.vals + (size_t)t * T, T * 8, full0 + 8 * s);
          tma_load_1d(smem_u32(cols_s + s * T), a.colidx + (size_t)t * T, T * 4, full0 + 8 * s);
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
  const int e0 = (int)(((int64_t)a.n * blockIdx.x) / G), e1 = (int)(((int64_t)a.n * (blockIdx.x + 1)) / G);
  long long gc = 0;     /* tiles consumed */
  unsigned bar_k = 0;   /* grid barriers passed in this launch */
  const bool jac = a.jacobi != 0;
  for (int it = 0; it < a.chunk_iters; ++it) {
    if (*(volatile int *)&a.st->done) break;
    const bool init = (*(volatile int *)&a.st->x_pending == 0);
    double sums[3] = {0.0, 0.0, 0.0}; /* (r,u), (w,u), (u,u) */
    /* ---- vector update ---- */
    if (init) {
      for (int i = e0 + tid; i < e1; i += NC) {
        const double bi = a.b[i];
        a.x[i] = 0.0; a.p[i] = 0.0; a.s[i] = 0.0;
        a.r[i] = bi;
        double ui = bi;
        if (jac) { ui = a.dinv[i] * bi; a.u[i] = ui; }
        sums[0] += bi * ui; sums[2] += ui * ui;
      }
    } else {
      const double alpha = *(volatile double *)&a.st->alpha, beta = *(volatile double *)&a.st->beta;
      constexpr int U = 4; /* elements in flight per thread: 608 threads x 4 x 5 streams x 8 B = 97 KB per SM */
      for (int i0 = e0 + tid; i0 < e1; i0 += U * NC) {
        double rv[U], uv[U], pv[U], sv[U], wv[U], xv[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int i = i0 + k * NC;
          if (i < e1) {
            rv[k] = ld_weak(a.r + i);
            uv[k] = jac ? ld_weak(a.u + i) : rv[k];
            pv[k] = ld_weak(a.p + i); sv[k] = ld_weak(a.s + i); wv[k] = ld_weak(a.w + i); xv[k] = ld_weak(a.x + i);
          }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int i = i0 + k * NC;
          if (i < e1) {
            const double pi = fma(beta, pv[k], uv[k]);
            const double si = fma(beta, sv[k], wv[k]);
            a.p[i] = pi; a.s[i] = si;
            a.x[i] = fma(alpha, pi, xv[k]);
            const double ri = fma(-alpha, si, rv[k]);
            a.r[i] = ri;
            double un = ri;
            if (jac) { un = a.dinv[i] * ri; a.u[i] = un; }
            sums[0] += ri * un; sums[2] += un * un;
          }
        }
      }
    }
    /* ---- barrier A: everybody's u is written; the leader exchanges the halo ---- */
    ++bar_k;
    if (gbar_arrive(a.sync, (unsigned)G * bar_k, &s_flag)) {
      if (a.mode == OHP_MODE_P2P) {
        const ohp_p2p_dev &pd = a.pd;
        const int nsend = pd.soff[pd.n_nbr];
        for (int i = tid; i < nsend; i += NC) {
          const int l = pd.send_ldof[i];
          int k = 0;
          while (i >= pd.soff[k + 1]) ++k;
          pd.p[0][pd.nbr[k]][pd.dst_off[k] + (i - pd.soff[k])] = ld_weak(a.u + l);
        }
        __threadfence_system();
        cons_sync();
        if (tid == 0) { s_seq = pd.seq[1] + 1; pd.seq[1] = s_seq; }
        cons_sync();
        if (tid < pd.n_nbr) {
          const int q = pd.nbr[tid];
          st_release_sys(pd.hflags[q] + pd.rank, s_seq);
          p2p_wait(pd.hflags[pd.rank] + q, s_seq, pd.seq + 2, pd.timeout_ns);
        }
      }
      gbar_release(a.sync + 1, bar_k);
    } else {
      gbar_wait(a.sync + 1, bar_k);
    }
    /* ---- SpMV w = A u over this CTA's tiles, (w,u) fused ---- */
    {
      double carry = 0.0;
      int pend = -1;
      for (int t = t0; t < t1; ++t, ++gc) {
        const int s = (int)(gc % S);
        mbar_wait(full0 + 8 * s, (uint32_t)((gc / S) & 1));
        const int32_t *rp = rp_s + s * kTmaRpStride;
        const double *vs = vals_s + s * T;
        const int32_t *cs = cols_s + s * T;
        const int rs = rp[kSpmvRowCap + 2], re = rp[kSpmvRowCap + 3];
        const int base = t * T, lim = base + T;
        if (pend >= 0) {
          const int end = rp[0];
          carry = row_dot_weak(vs, cs, 0, min(end, lim) - base, a.u, carry);
          if (end <= lim) { a.w[pend] = carry; sums[1] += carry * ld_weak(a.u + pend); pend = -1; }
        }
        for (int r = rs + tid; r < re; r += NC) {
          const int ra = rp[r - rs], rb = rp[r - rs + 1];
          const double sum = row_dot_weak(vs, cs, ra - base, min(rb, lim) - base, a.u, 0.0);
          if (rb > lim) { carry = sum; pend = r; }
          else { a.w[r] = sum; sums[1] += sum * ld_weak(a.u + r); }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
      }
      if (pend >= 0) {
        const int end = __ldg(a.rowptr + pend + 1);
        for (int k = t1 * T; k < end; ++k) carry += __ldg(a.vals + k) * ld_weak(a.u + __ldg(a.colidx + k));
        a.w[pend] = carry;
        sums[1] += carry * ld_weak(a.u + pend);
      }
    }
    /* ---- barrier B: reduction, all-reduce, scalar recurrences ---- */
    cons_block_sum<3>(sums, s_red);
    if (tid == 0) { a.partials[blockIdx.x * 3 + 0] = sums[0]; a.partials[blockIdx.x * 3 + 1] = sums[1]; a.partials[blockIdx.x * 3 + 2] = sums[2]; }
    ++bar_k;
    if (gbar_arrive(a.sync, (unsigned)G * bar_k, &s_flag)) {
      double tot[3] = {0.0, 0.0, 0.0};
      for (int bq = tid; bq < G; bq += NC) { /* G <= 608: one partial triple per thread, fixed tree below */
        tot[0] += __ldcg(a.partials + bq * 3 + 0); tot[1] += __ldcg(a.partials + bq * 3 + 1); tot[2] += __ldcg(a.partials + bq * 3 + 2);
      }
      cons_block_sum<3>(tot, s_red);
      if (a.mode == OHP_MODE_P2P) p2p_allreduce<3>(a.pd, tot); /* warp 0; totals in thread 0 */
      if (tid == 0) cgcg_scalars(a.st, a.hist, a.hist_cap, tot[0], tot[1], jac ? tot[2] : tot[0]);
      gbar_release(a.sync + 1, bar_k);
    } else {
      gbar_wait(a.sync + 1, bar_k);
    }
  }
  cons_sync();
  if (tid == 0) { s_consumed = (int)gc; __threadfence_block(); s_exit = 1; }
  (void)ntile;
}

}  // namespace

/* host side: runs the solve in chunks of `check_every` iterations per cooperative launch */
int ohp_cg_persistent_solve(ohp_mat A, ohp_vec b, ohp_vec x, const ohp_ksp_options *o, ohp_ksp_result *res, bool p2p) {
  ohp_mesh m = A->mesh;
  ohp_context ctx = m->ctx;
  cudaStream_t st = ctx->stream;
  const int n = m->n_owned;
  const size_t len = (size_t)std::max(1, n + m->n_ghost);
  const bool jac = (o->pc == OHP_PC_JACOBI);
  if (!A->r) {
    OHP_CUDA(OHP_MALLOC(ctx, &A->r, sizeof(double) * len));
    OHP_CUDA(OHP_MALLOC(ctx, &A->p, sizeof(double) * len));
    OHP_CUDA(OHP_MALLOC(ctx, &A->p1, sizeof(double) * len));
    OHP_CUDA(OHP_MALLOC(ctx, &A->w, sizeof(double) * len));
  }
  if (!A->s) OHP_CUDA(OHP_MALLOC(ctx, &A->s, sizeof(double) * len));
  if (jac && !A->z) OHP_CUDA(OHP_MALLOC(ctx, &A->z, sizeof(double) * len));
  if (jac && !A->dinv) OHP_CUDA(OHP_MALLOC(ctx, &A->dinv, sizeof(double) * len)); /* another recurrence may have made it already */
  if (jac) OHP_TRY(ohp_extract_dinv(A));
  const int smem = tma_smem_bytes(PS);
  OHP_CUDA(cudaFuncSetAttribute(k_cg_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); /* per device */
  static const ohp_p2p_dev no_p2p = {};
  PcgArgs a;
  memset(&a, 0, sizeof(a));
  a.rowptr = m->rowptr; a.colidx = m->colidx; a.tile_row = m->tile_row; a.vals = A->vals; a.b = b->d; a.dinv = A->dinv;
  a.x = x->d; a.p = A->p; a.s = A->s; a.w = A->w;
  a.pd = p2p ? m->p2p->dev : no_p2p;
  /* the SpMV input (u) must sit in the exchange heap when neighbours push into its ghost tail */
  double *uvec = p2p ? a.pd.p[0][ctx->rank] : (jac ? A->z : A->r);
  a.u = uvec;
  a.r = jac ? A->r : uvec;
  a.st = ctx->d_cg; a.partials = ctx->d_partials; a.sync = ctx->d_ticket + 4;
  a.n = n; a.n_tiles = m->n_tiles; a.jacobi = jac ? 1 : 0;
  a.mode = ctx->nranks == 1 ? OHP_MODE_SERIAL : OHP_MODE_P2P;
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
  ohp_cg_state init;
  memset(&init, 0, sizeof(init));
  init.rtol = o->rtol; init.abstol = o->abstol; init.dtol = o->dtol; init.maxit = o->max_it;
  *ctx->h_cg = init;
  OHP_CUDA(cudaMemcpyAsync(ctx->d_cg, ctx->h_cg, sizeof(init), cudaMemcpyHostToDevice, st));
  const int grid = std::min(ctx->num_sms, PNC);
  a.chunk_iters = o->check_every > 0 ? o->check_every : 200;
  if (p2p) OHP_TRY(ohp_p2p_entry_barrier(m));
  struct Ev { cudaEvent_t e = nullptr; ~Ev() { if (e) cudaEventDestroy(e); } } e0, e1;
  OHP_CUDA(cudaEventCreate(&e0.e)); OHP_CUDA(cudaEventCreate(&e1.e));
  cudaEvent_t ev0 = e0.e, ev1 = e1.e;
  OHP_CUDA(cudaEventRecord(ev0, st));
  int launched = 0;
  for (;;) {
    OHP_CUDA(cudaMemsetAsync(a.sync, 0, sizeof(unsigned) * 2, st));
    void *params[] = {(void *)&a};
    OHP_CUDA(cudaLaunchCooperativeKernel((const void *)k_cg_persistent, dim3(grid), dim3(PNC + 32), params, (size_t)smem, st));
    ctx->launches++;
    OHP_CUDA(cudaMemcpyAsync(ctx->h_cg, ctx->d_cg, sizeof(ohp_cg_state), cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
    if (ctx->h_cg->done) break;
    launched += a.chunk_iters;
    if (launched > o->max_it + 2 * a.chunk_iters) OHP_FAIL(OHP_ERR_LIB, "ohp_ksp_cg (persistent): iteration control did not terminate");
  }
  OHP_CUDA(cudaEventRecord(ev1, st));
  OHP_CUDA(cudaEventSynchronize(ev1));
  res->total_ms = 0.f;
  cudaEventElapsedTime(&res->total_ms, ev0, ev1);
  res->its = ctx->h_cg->its; res->reason = ctx->h_cg->reason; res->rnorm = ctx->h_cg->dp; res->rnorm0 = ctx->h_cg->rnorm0;
  res->spmv_ms = 0.f; res->spmv_samples = 0; res->algo = OHP_ALGO_PERSISTENT;
  if (p2p) OHP_TRY(ohp_p2p_check(m, "ohp_ksp_cg (persistent)"));
  if (hist) {
    const int ncopy = std::min(o->history_cap, res->its + 1);
    OHP_CUDA(cudaMemcpyAsync(o->history, hist, sizeof(double) * ncopy, cudaMemcpyDeviceToHost, st));
    OHP_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}
