/*
 * ohp_fem.cuh -- sm_100a assembly kernels of the P1 Poisson path, templated on the problem
 * (the PetscDS pointwise callbacks).  libohp.so instantiates them for ex12's built-in
 * callback sets; a user translation unit compiled by nvcc instantiates them for its own
 * __host__ __device__ callbacks through include/ohp_petsc.h and registers the launchers
 * with ohp_problem_register().
 *
 * What is restated here [PETSc-internal, reached from DMPlexSetSNESLocalFEM ex12.cpp:1540]:
 *   DMPlexComputeCellGeometryFEM (affine simplex: v0, J, invJ, detJ on the biunit reference
 *   cell), PetscFE P1 tabulation, PetscFEIntegrateResidual_Basic / IntegrateJacobian_Basic,
 *   DMPlexInsertBoundaryValues (essential values from the BC function at the vertices),
 *   DMPlexVecSetClosure / DMPlexMatSetClosure (ADD, constrained dofs dropped).
 *
 * Parallelisation: "owner computes", colour-free and atomic-free.  One thread owns one matrix
 * row (= one unconstrained owned vertex v) and walks star(v) in ascending cell order, so the
 * per-slot summation order equals PETSc's cell-loop order and results are bitwise reproducible
 * run to run.  Row values are accumulated in shared memory (each thread touches only its own
 * row's slots) and flushed to HBM as one coalesced stream per block.
 */
#ifndef OHP_FEM_CUH
#define OHP_FEM_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "ohp.h"

#ifndef OHP_HD
#define OHP_HD __host__ __device__ __forceinline__
#endif

namespace ohp {

/* PetscDS pointwise-callback argument lists (ex12.cpp:147-150 and :209-212) */
#define OHP_PF_ARGS                                                                              \
  int dim, int Nf, int NfAux, const int uOff[], const int uOff_x[], const double u[],            \
      const double u_t[], const double u_x[], const int aOff[], const int aOff_x[],              \
      const double a[], const double a_t[], const double a_x[], double t, const double x[],      \
      int numConstants, const double constants[]
#define OHP_PJ_ARGS                                                                              \
  int dim, int Nf, int NfAux, const int uOff[], const int uOff_x[], const double u[],            \
      const double u_t[], const double u_x[], const int aOff[], const int aOff_x[],              \
      const double a[], const double a_t[], const double a_x[], double t, double u_tShift,       \
      const double x[], int numConstants, const double constants[]

/* Stroud conical quadrature, 2 Gauss-Jacobi points per direction, on the biunit simplex
 * [PETSc-internal: PetscFECreateDefault(..., qorder=-1) at ex12.cpp:1309 ->
 * PetscDTStroudConicalQuadrature]; sum of weights 2 (2-D), 4/3 (3-D). */
template <int DIM> struct Quadrature;
template <> struct Quadrature<2> {
  static constexpr int NQ = 4;
  OHP_HD static double xi(int q, int d) {
    constexpr double P[4][2] = {{-6.42882543472767187254e-01, -6.89897948556635665085e-01},
                                {-8.49937779554783778835e-01, 2.89897948556635642881e-01},
                                {3.32780492029402796827e-01, -6.89897948556635665085e-01},
                                {-4.39960169001851864046e-01, 2.89897948556635642881e-01}};
    return P[q][d];
  }
  OHP_HD static double w(int q) {
    constexpr double W[4] = {6.36082763487954339077e-01, 3.63917236512045660923e-01,
                             6.36082763487954339077e-01, 3.63917236512045660923e-01};
    return W[q];
  }
};
template <> struct Quadrature<3> {
  static constexpr int NQ = 8;
  OHP_HD static double xi(int q, int d) {
    constexpr double P[8][3] = {
        {-6.86634725326363382081e-01, -7.27890046394307987931e-01, -7.54970354689117217895e-01},
        {-8.37208665970659460243e-01, -8.58640551681206232182e-01, 8.83036880224505743575e-02},
        {-8.68322625879911158542e-01, 1.31866330145601756696e-01, -7.54970354689117217895e-01},
        {-9.31594413526467213238e-01, -4.12002398736754260611e-01, 8.83036880224505743575e-02},
        {1.69495126409788587907e-01, -7.27890046394307987931e-01, -7.54970354689117217895e-01},
        {-3.92454470370584895811e-01, -8.58640551681206232182e-01, 8.83036880224505743575e-02},
        {-5.08573349576573296993e-01, 1.31866330145601756696e-01, -7.54970354689117217895e-01},
        {-7.44706875759229114387e-01, -4.12002398736754260611e-01, 8.83036880224505743575e-02}};
    return P[q][d];
  }
  OHP_HD static double w(int q) {
    constexpr double W[8] = {2.95838850870823288908e-01, 1.28216324787812918640e-01,
                             1.69256051636192478282e-01, 7.33554393718379577072e-02,
                             2.95838850870823288908e-01, 1.28216324787812918640e-01,
                             1.69256051636192478282e-01, 7.33554393718379577072e-02};
    return W[q];
  }
};

/* Affine cell geometry on the biunit reference simplex: J[d][f] = (x_{f+1}[d]-x_0[d])/2,
 * real-space P1 gradients G[b][d] = sum_e Dref[b][e] invJ[e][d], Dref row 0 = -1/2,
 * row b>0 = +1/2 at e = b-1. */
template <int DIM> struct CellGeom {
  double v0[DIM], J[DIM * DIM], detJ, G[DIM + 1][DIM];
};

template <int DIM> OHP_HD void cell_geometry(const double (&X)[DIM + 1][DIM], CellGeom<DIM> &g) {
  double I[DIM * DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    g.v0[d] = X[0][d];
#pragma unroll
    for (int f = 0; f < DIM; ++f) g.J[d * DIM + f] = 0.5 * (X[f + 1][d] - X[0][d]);
  }
  const double *J = g.J;
  if constexpr (DIM == 2) {
    g.detJ = J[0] * J[3] - J[1] * J[2];
    const double id = 1.0 / g.detJ;
    I[0] = id * J[3]; I[1] = -id * J[1]; I[2] = -id * J[2]; I[3] = id * J[0];
  } else {
    const double c0 = J[4] * J[8] - J[5] * J[7];
    const double c1 = J[5] * J[6] - J[3] * J[8];
    const double c2 = J[3] * J[7] - J[4] * J[6];
    g.detJ = J[0] * c0 + J[1] * c1 + J[2] * c2;
    const double id = 1.0 / g.detJ;
    I[0] = id * c0;
    I[1] = id * (J[2] * J[7] - J[1] * J[8]);
    I[2] = id * (J[1] * J[5] - J[2] * J[4]);
    I[3] = id * c1;
    I[4] = id * (J[0] * J[8] - J[2] * J[6]);
    I[5] = id * (J[2] * J[3] - J[0] * J[5]);
    I[6] = id * c2;
    I[7] = id * (J[1] * J[6] - J[0] * J[7]);
    I[8] = id * (J[0] * J[4] - J[1] * J[3]);
  }
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    double s = 0.0;
#pragma unroll
    for (int e = 0; e < DIM; ++e) s += I[e * DIM + d];
    g.G[0][d] = -0.5 * s;
#pragma unroll
    for (int b = 1; b <= DIM; ++b) g.G[b][d] = 0.5 * I[(b - 1) * DIM + d];
  }
}

template <int DIM> OHP_HD void p1_basis(int q, double (&phi)[DIM + 1]) {
  double s = 0.0;
#pragma unroll
  for (int e = 0; e < DIM; ++e) {
    phi[e + 1] = 0.5 * (1.0 + Quadrature<DIM>::xi(q, e));
    s += Quadrature<DIM>::xi(q, e);
  }
  phi[0] = -0.5 * (s + DIM - 2.0);
}

/* closure of one cell: vertex ids, coordinates and (optionally) local solution values with
 * the essential boundary values inserted from the problem's BC function at the vertex
 * (registered with DMAddBoundary at ex12.cpp:1237-1239) */
template <int DIM, class P, bool WITH_U>
__device__ __forceinline__ void load_cell(const ohp_assembly_view &vw, int c, int (&cv)[DIM + 1],
                                          double (&X)[DIM + 1][DIM], double (&ue)[DIM + 1]) {
  constexpr int NC = DIM + 1;
#pragma unroll
  for (int j = 0; j < NC; ++j) cv[j] = __ldg(vw.cells_dev + (size_t)c * NC + j);
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    if constexpr (DIM == 2) { /* one 16-byte load per vertex (cudaMalloc'ed base, 16-byte stride) */
      const double2 xy = __ldg(reinterpret_cast<const double2 *>(vw.coords_dev) + cv[j]);
      X[j][0] = xy.x; X[j][1] = xy.y;
    } else {
#pragma unroll
      for (int d = 0; d < DIM; ++d) X[j][d] = __ldg(vw.coords_dev + (size_t)cv[j] * DIM + d);
    }
  }
  if (WITH_U) {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int l = __ldg(vw.lidx_dev + cv[j]);
      if (l >= 0) ue[j] = __ldg(vw.u_dev + l);
      else { double val = 0.0; P::bc(DIM, 0.0, X[j], 1, &val, nullptr); ue[j] = val; }
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* Per-incidence work: what cell c contributes to the row of its corner b.                */
/* ------------------------------------------------------------------------------------ */
/* Fe_c[b] = sum_q w_q detJ ( phi_b f0 + grad(phi_b) . f1 ) */
template <int DIM, class P>
__device__ __forceinline__ double cell_residual_entry(const ohp_assembly_view &vw, int c, int b) {
  constexpr int NC = DIM + 1;
  const int uOff[2] = {0, 1}, uOff_x[2] = {0, DIM};
  int cv[NC];
  double X[NC][DIM], ue[NC];
  load_cell<DIM, P, true>(vw, c, cv, X, ue);
  CellGeom<DIM> g;
  cell_geometry<DIM>(X, g);
  double ux[DIM], Gb[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < NC; ++j) s += g.G[j][d] * ue[j];
    ux[d] = s;
    double gb = g.G[0][d];
#pragma unroll
    for (int j = 1; j < NC; ++j) gb = (b == j) ? g.G[j][d] : gb;
    Gb[d] = gb;
  }
  double fe = 0.0;
#pragma unroll
  for (int q = 0; q < Quadrature<DIM>::NQ; ++q) {
    double phi[NC], xq[DIM], uq = 0.0;
    p1_basis<DIM>(q, phi);
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double s = g.v0[d];
#pragma unroll
      for (int f = 0; f < DIM; ++f) s += g.J[d * DIM + f] * (Quadrature<DIM>::xi(q, f) + 1.0);
      xq[d] = s;
    }
    double phib = phi[0];
#pragma unroll
    for (int j = 0; j < NC; ++j) { uq += phi[j] * ue[j]; phib = (b == j) ? phi[j] : phib; }
    double f0[1] = {0.0}, f1[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) f1[d] = 0.0;
    P::f0(DIM, 1, 0, uOff, uOff_x, &uq, nullptr, ux, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, xq, 0, nullptr, f0);
    P::f1(DIM, 1, 0, uOff, uOff_x, &uq, nullptr, ux, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, xq, 0, nullptr, f1);
    const double w = g.detJ * Quadrature<DIM>::w(q);
    double s = phib * (f0[0] * w);
#pragma unroll
    for (int d = 0; d < DIM; ++d) s += Gb[d] * (f1[d] * w);
    fe += s;
  }
  return fe;
}

/* Ke_c[b][j] = detJ * G_b^T (sum_q w_q g3_q) G_j   (P1: gradients constant per cell) */
template <int DIM, class P>
__device__ __forceinline__ void cell_jacobian_row(const ohp_assembly_view &vw, int c, int b, double (&ke)[DIM + 1]) {
  constexpr int NC = DIM + 1;
  const int uOff[2] = {0, 1}, uOff_x[2] = {0, DIM};
  int cv[NC];
  double X[NC][DIM], ue[NC];
  load_cell<DIM, P, P::kJacobianUsesU>(vw, c, cv, X, ue);
  CellGeom<DIM> g;
  cell_geometry<DIM>(X, g);
  double ux[DIM], Gb[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    double s = 0.0;
    if (P::kJacobianUsesU) {
#pragma unroll
      for (int j = 0; j < NC; ++j) s += g.G[j][d] * ue[j];
    }
    ux[d] = s;
    double gb = g.G[0][d];
#pragma unroll
    for (int j = 1; j < NC; ++j) gb = (b == j) ? g.G[j][d] : gb;
    Gb[d] = gb;
  }
  double S[DIM * DIM];
#pragma unroll
  for (int i = 0; i < DIM * DIM; ++i) S[i] = 0.0;
#pragma unroll
  for (int q = 0; q < Quadrature<DIM>::NQ; ++q) {
    double xq[DIM], uq = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double s = g.v0[d];
#pragma unroll
      for (int f = 0; f < DIM; ++f) s += g.J[d * DIM + f] * (Quadrature<DIM>::xi(q, f) + 1.0);
      xq[d] = s;
    }
    if (P::kJacobianUsesU) {
      double phi[NC];
      p1_basis<DIM>(q, phi);
#pragma unroll
      for (int j = 0; j < NC; ++j) uq += phi[j] * ue[j];
    }
    double g3[DIM * DIM];
#pragma unroll
    for (int i = 0; i < DIM * DIM; ++i) g3[i] = 0.0; /* the integrator zeroes g3: g3_uu writes only the diagonal (ex12.cpp:215) */
    P::g3(DIM, 1, 0, uOff, uOff_x, &uq, nullptr, ux, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, 0.0, xq, 0, nullptr, g3);
    const double w = Quadrature<DIM>::w(q);
#pragma unroll
    for (int i = 0; i < DIM * DIM; ++i) S[i] += w * g3[i];
  }
  double t[DIM];
#pragma unroll
  for (int e2 = 0; e2 < DIM; ++e2) {
    double s = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) s += Gb[d] * S[d * DIM + e2];
    t[e2] = g.detJ * s;
  }
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    double s = 0.0;
#pragma unroll
    for (int e2 = 0; e2 < DIM; ++e2) s += t[e2] * g.G[j][e2];
    ke[j] = s;
  }
}

/* ------------------------------------------------------------------------------------ */
/* Assembly kernels: "owner computes".  One thread owns one matrix row and walks the star   */
/* of its vertex IN ASCENDING CELL ORDER (PETSc's cell-loop order): no atomics, no           */
/* colouring, bitwise reproducible.  Jacobian rows accumulate in shared memory (a thread     */
/* only touches its own row) and leave as one coalesced stream per block; a block whose      */
/* rows do not fit accumulates in place in HBM.                                             */
/* (An incidence-parallel variant that staged every (vertex, cell) contribution through      */
/* shared memory first was measured 2x SLOWER on B200: the kernels are bound by L1 gather    */
/* wavefronts, not by per-thread latency, and the staging buffers halve the occupancy.)      */
/* ------------------------------------------------------------------------------------ */
constexpr int kAsmThreads = 256;
constexpr int kAsmAccCap = 6144; /* row entries accumulated per block: 48 KB */

template <int DIM, class P>
__global__ void __launch_bounds__(kAsmThreads) k_residual_rows(const ohp_assembly_view vw) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= vw.n_rows) return;
  const int v = __ldg(vw.row_vertex_dev + r);
  const int k0 = __ldg(vw.v2c_ptr_dev + v), k1 = __ldg(vw.v2c_ptr_dev + v + 1);
  double acc = 0.0;
  for (int k = k0; k < k1; ++k) {
    const int e = __ldg(vw.v2c_dev + k);
    acc += cell_residual_entry<DIM, P>(vw, e >> 2, e & 3);
  }
  vw.out_dev[r] = acc;
}

template <int DIM, class P>
__global__ void __launch_bounds__(kAsmThreads) k_jacobian_rows(const ohp_assembly_view vw) {
  constexpr int NC = DIM + 1;
  extern __shared__ double acc_s[];
  const int r0 = blockIdx.x * blockDim.x;
  const int r1 = min(r0 + (int)blockDim.x, vw.n_rows);
  const int s0 = __ldg(vw.rowptr_dev + r0), s1 = __ldg(vw.rowptr_dev + r1);
  const bool acc_in_smem = (s1 - s0) <= kAsmAccCap;
  double *acc_base = acc_in_smem ? acc_s : (vw.out_dev + s0);
  for (int i = threadIdx.x; i < s1 - s0; i += blockDim.x) acc_base[i] = 0.0;
  __syncthreads();
  const int r = r0 + threadIdx.x;
  if (r < r1) {
    const int v = __ldg(vw.row_vertex_dev + r);
    const int k0 = __ldg(vw.v2c_ptr_dev + v), k1 = __ldg(vw.v2c_ptr_dev + v + 1);
    double *row = acc_base + (__ldg(vw.rowptr_dev + r) - s0);
    for (int k = k0; k < k1; ++k) {
      const int e = __ldg(vw.v2c_dev + k);
      double ke[NC];
      cell_jacobian_row<DIM, P>(vw, e >> 2, e & 3, ke);
      const uint8_t *sl = vw.slot_dev + (size_t)k * NC;
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int slot = __ldg(sl + j);
        if (slot != 255) row[slot] += ke[j];
      }
    }
  }
  if (acc_in_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < s1 - s0; i += blockDim.x) vw.out_dev[s0 + i] = acc_s[i];
  }
}

/* DMProjectFunction of the exact solution onto P1 = point evaluation at the vertices */
template <int DIM, class P> __global__ void k_project_exact(const ohp_assembly_view vw) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= vw.n_rows) return;
  const int v = __ldg(vw.row_vertex_dev + r);
  double x[DIM], val = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) x[d] = __ldg(vw.coords_dev + (size_t)v * DIM + d);
  P::exact(DIM, 0.0, x, 1, &val, nullptr);
  vw.out_dev[r] = val;
}

/* DMProjectFunction of any point function F::eval(dim, time, x, Nc, u, ctx) compiled for the device */
template <int DIM, class F> __global__ void k_project_fn(const ohp_assembly_view vw) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= vw.n_rows) return;
  const int v = __ldg(vw.row_vertex_dev + r);
  double x[DIM], val = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) x[d] = __ldg(vw.coords_dev + (size_t)v * DIM + d);
  F::eval(DIM, 0.0, x, 1, &val, nullptr);
  vw.out_dev[r] = val;
}

/* DMComputeL2Diff [PETSc-internal, reached from ex12.cpp:1611,1637]: sum_c sum_q w_q |detJ| (u_h - u_exact)^2 with the
 * problem's own exact solution and the assembly quadrature; one partial sum per block, added in block order by the
 * host (deterministic).  Partitioned meshes overlap by one layer of cells: a cell is integrated by the rank that
 * owns its first corner. */
template <int DIM, class P>
__global__ void __launch_bounds__(256) k_l2_diff(const ohp_assembly_view vw) {
  constexpr int NC = DIM + 1;
  double acc = 0.0;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < vw.n_cells; c += gridDim.x * blockDim.x) {
    if (!vw.owned_dev[vw.cells_dev[(size_t)c * NC]]) continue;
    int cv[NC];
    double X[NC][DIM], ue[NC];
    load_cell<DIM, P, true>(vw, c, cv, X, ue);
    CellGeom<DIM> g;
    cell_geometry<DIM>(X, g);
    for (int q = 0; q < Quadrature<DIM>::NQ; ++q) {
      double phi[NC], xq[DIM], uq = 0.0, ex = 0.0;
      p1_basis<DIM>(q, phi);
      for (int d = 0; d < DIM; ++d) {
        double s = g.v0[d];
        for (int f = 0; f < DIM; ++f) s += g.J[d * DIM + f] * (Quadrature<DIM>::xi(q, f) + 1.0);
        xq[d] = s;
      }
      for (int j = 0; j < NC; ++j) uq += phi[j] * ue[j];
      P::exact(DIM, 0.0, xq, 1, &ex, nullptr);
      acc += Quadrature<DIM>::w(q) * fabs(g.detJ) * (uq - ex) * (uq - ex);
    }
  }
  __shared__ double red[8];
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    vw.out_dev[blockIdx.x] = t;
  }
}


/* ------------------------------------------------------------------------------------ */
/* Cell-parallel assembly over "patches" (tables built by csrc/ohp_patch.cu at set-up).     */
/* The row kernels above recompute a cell once per incident row (3x per triangle, 4x per     */
/* tetrahedron) and were FP64-issue bound at 0.20 of the HBM roofline.  A patch is a block of */
/* rows that are close in space together with the cells of their stars and the vertices of    */
/* those cells, each listed ONCE: the block stages the patch's vertex coordinates and          */
/* solution values in shared memory (one coalesced pass over the vertex list), one thread per   */
/* cell computes the whole element vector / matrix into shared memory (connectivity read as     */
/* 16-bit patch-local ids), and one thread per row adds up its star from there in ascending     */
/* cell order -- PETSc's cell-loop order, the same order and the same per-entry arithmetic as   */
/* the row kernels, so the values are bit-identical to theirs: no atomics, no colouring.        */
/* ------------------------------------------------------------------------------------ */
constexpr int kPatchThreads = 128;

/* all rows of the element vector at once: Fe[b] = sum_q w_q detJ ( phi_b f0 + grad(phi_b) . f1 ), entry by entry the
 * arithmetic of cell_residual_entry */
template <int DIM, class P>
__device__ __forceinline__ void cell_residual_all(const double (&X)[DIM + 1][DIM], const double (&ue)[DIM + 1], double (&fe)[DIM + 1]) {
  constexpr int NC = DIM + 1;
  const int uOff[2] = {0, 1}, uOff_x[2] = {0, DIM};
  CellGeom<DIM> g;
  cell_geometry<DIM>(X, g);
  double ux[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < NC; ++j) s += g.G[j][d] * ue[j];
    ux[d] = s;
  }
#pragma unroll
  for (int b = 0; b < NC; ++b) fe[b] = 0.0;
#pragma unroll
  for (int q = 0; q < Quadrature<DIM>::NQ; ++q) {
    double phi[NC], xq[DIM], uq = 0.0;
    p1_basis<DIM>(q, phi);
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double s = g.v0[d];
#pragma unroll
      for (int f = 0; f < DIM; ++f) s += g.J[d * DIM + f] * (Quadrature<DIM>::xi(q, f) + 1.0);
      xq[d] = s;
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) uq += phi[j] * ue[j];
    double f0[1] = {0.0}, f1[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) f1[d] = 0.0;
    P::f0(DIM, 1, 0, uOff, uOff_x, &uq, nullptr, ux, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, xq, 0, nullptr, f0);
    P::f1(DIM, 1, 0, uOff, uOff_x, &uq, nullptr, ux, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, xq, 0, nullptr, f1);
    const double w = g.detJ * Quadrature<DIM>::w(q);
#pragma unroll
    for (int b = 0; b < NC; ++b) {
      double s = phi[b] * (f0[0] * w);
#pragma unroll
      for (int d = 0; d < DIM; ++d) s += g.G[b][d] * (f1[d] * w);
      fe[b] += s;
    }
  }
}

/* the whole element matrix: Ke[b][j] = detJ * G_b^T (sum_q w_q g3_q) G_j, entry by entry the arithmetic of cell_jacobian_row */
template <int DIM, class P>
__device__ __forceinline__ void cell_jacobian_all(const double (&X)[DIM + 1][DIM], const double (&ue)[DIM + 1], double *ke /* NC*NC */) {
  constexpr int NC = DIM + 1;
  const int uOff[2] = {0, 1}, uOff_x[2] = {0, DIM};
  CellGeom<DIM> g;
  cell_geometry<DIM>(X, g);
  double ux[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    double s = 0.0;
    if (P::kJacobianUsesU) {
#pragma unroll
      for (int j = 0; j < NC; ++j) s += g.G[j][d] * ue[j];
    }
    ux[d] = s;
  }
  double S[DIM * DIM];
#pragma unroll
  for (int i = 0; i < DIM * DIM; ++i) S[i] = 0.0;
#pragma unroll
  for (int q = 0; q < Quadrature<DIM>::NQ; ++q) {
    double xq[DIM], uq = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double s = g.v0[d];
#pragma unroll
      for (int f = 0; f < DIM; ++f) s += g.J[d * DIM + f] * (Quadrature<DIM>::xi(q, f) + 1.0);
      xq[d] = s;
    }
    if (P::kJacobianUsesU) {
      double phi[NC];
      p1_basis<DIM>(q, phi);
#pragma unroll
      for (int j = 0; j < NC; ++j) uq += phi[j] * ue[j];
    }
    double g3[DIM * DIM];
#pragma unroll
    for (int i = 0; i < DIM * DIM; ++i) g3[i] = 0.0;
    P::g3(DIM, 1, 0, uOff, uOff_x, &uq, nullptr, ux, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, 0.0, xq, 0, nullptr, g3);
    const double w = Quadrature<DIM>::w(q);
#pragma unroll
    for (int i = 0; i < DIM * DIM; ++i) S[i] += w * g3[i];
  }
#pragma unroll
  for (int b = 0; b < NC; ++b) {
    double t[DIM];
#pragma unroll
    for (int e2 = 0; e2 < DIM; ++e2) {
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) s += g.G[b][d] * S[d * DIM + e2];
      t[e2] = g.detJ * s;
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      double s = 0.0;
#pragma unroll
      for (int e2 = 0; e2 < DIM; ++e2) s += t[e2] * g.G[j][e2];
      ke[b * NC + j] = s;
    }
  }
}

template <int DIM, class P, bool JAC>
__global__ void __launch_bounds__(kPatchThreads) k_assemble_patch(const ohp_assembly_view vw) {
  constexpr int NC = DIM + 1, EN = JAC ? NC * NC : NC;
  constexpr bool NEED_U = JAC ? P::kJacobianUsesU : true;
  extern __shared__ double sm[];
  double *Xs = sm;                                   /* patch_max_verts x DIM */
  double *us = Xs + (size_t)vw.patch_max_verts * DIM; /* patch_max_verts */
  double *E = us + vw.patch_max_verts;               /* patch_max_cells x EN */
  double *acc = E + (size_t)vw.patch_max_cells * EN; /* JAC: patch_rows x max_row */
  const int p = blockIdx.x, R = vw.patch_rows, tid = threadIdx.x;
  const int w0 = __ldg(vw.pvert_ptr_dev + p), nv = __ldg(vw.pvert_ptr_dev + p + 1) - w0;
  const int c0 = __ldg(vw.pcell_ptr_dev + p), nu = __ldg(vw.pcell_ptr_dev + p + 1) - c0;
  /* 1. the patch's vertices, once */
  for (int i = tid; i < nv; i += kPatchThreads) {
    const int gv = __ldg(vw.pvert_dev + w0 + i);
    double x[DIM];
    if constexpr (DIM == 2) {
      const double2 xy = __ldg(reinterpret_cast<const double2 *>(vw.coords_dev) + gv);
      x[0] = xy.x; x[1] = xy.y;
    } else {
#pragma unroll
      for (int d = 0; d < DIM; ++d) x[d] = __ldg(vw.coords_dev + (size_t)gv * DIM + d);
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d) Xs[i * DIM + d] = x[d];
    if (NEED_U) {
      const int l = __ldg(vw.lidx_dev + gv);
      double val = 0.0;
      if (l >= 0) val = __ldg(vw.u_dev + l);
      else P::bc(DIM, 0.0, x, 1, &val, nullptr);
      us[i] = val;
    }
  }
  __syncthreads();
  /* 2. every cell of the patch, once */
  for (int c = tid; c < nu; c += kPatchThreads) {
    double X[NC][DIM], ue[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int lv = (int)__ldg(vw.pconn_dev + (size_t)(c0 + c) * NC + j);
#pragma unroll
      for (int d = 0; d < DIM; ++d) X[j][d] = Xs[lv * DIM + d];
      ue[j] = NEED_U ? us[lv] : 0.0;
    }
    if (JAC) {
      double ke[NC * NC];
      cell_jacobian_all<DIM, P>(X, ue, ke);
#pragma unroll
      for (int k = 0; k < NC * NC; ++k) E[(size_t)c * EN + k] = ke[k];
    } else {
      double fe[NC];
      cell_residual_all<DIM, P>(X, ue, fe);
#pragma unroll
      for (int k = 0; k < NC; ++k) E[(size_t)c * EN + k] = fe[k];
    }
  }
  __syncthreads();
  /* 3. every row adds up its star in ascending cell order */
  if (tid >= R) return;
  const int slot_i = p * R + tid;
  const int row = __ldg(vw.prow_dev + slot_i);
  if (row < 0) return;
  const int e0 = __ldg(vw.pstar_ptr_dev + slot_i), e1 = __ldg(vw.pstar_ptr_dev + slot_i + 1);
  if (!JAC) {
    double a = 0.0;
    for (int e = e0; e < e1; ++e) {
      const int code = (int)__ldg(vw.pstar_dev + e);
      a += E[(size_t)(code >> 2) * NC + (code & 3)];
    }
    vw.out_dev[row] = a;
  } else {
    const int k0 = __ldg(vw.rowptr_dev + row), len = __ldg(vw.rowptr_dev + row + 1) - k0;
    double *a = acc + (size_t)tid * vw.max_row;
    for (int k = 0; k < len; ++k) a[k] = 0.0;
    for (int e = e0; e < e1; ++e) {
      const int code = (int)__ldg(vw.pstar_dev + e);
      const double *ke = E + ((size_t)(code >> 2) * NC + (code & 3)) * NC;
      const uint8_t *sl = vw.pslot_dev + (size_t)e * NC;
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int slot = __ldg(sl + j);
        if (slot != 255) a[slot] += ke[j];
      }
    }
    for (int k = 0; k < len; ++k) vw.out_dev[k0 + k] = a[k];
  }
}

template <int DIM, class P, bool JAC> int launch_patch(const ohp_assembly_view *vw, bool *used) {
  constexpr int NC = DIM + 1, EN = JAC ? NC * NC : NC;
  *used = false;
  if (vw->n_patches <= 0 || vw->patch_rows > kPatchThreads) return 0;
  const size_t smem = sizeof(double) * ((size_t)vw->patch_max_verts * (DIM + 1) + (size_t)vw->patch_max_cells * EN +
                                        (JAC ? (size_t)vw->patch_rows * vw->max_row : 0));
  if (smem > 200 * 1024) return 0; /* very irregular patches: the row kernels handle them */
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(k_assemble_patch<DIM, P, JAC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return OHP_ERR_LIB;
  }
  k_assemble_patch<DIM, P, JAC><<<vw->n_patches, kPatchThreads, smem, (cudaStream_t)vw->stream>>>(*vw);
  *used = true;
  return cudaGetLastError() == cudaSuccess ? 0 : OHP_ERR_LIB;
}

template <int DIM, class P> int launch_residual(const ohp_assembly_view *vw) {
  if (vw->n_rows <= 0) return 0;
  bool used = false;
  const int rc = launch_patch<DIM, P, false>(vw, &used);
  if (rc || used) return rc;
  const int bs = 128;
  k_residual_rows<DIM, P><<<(vw->n_rows + bs - 1) / bs, bs, 0, (cudaStream_t)vw->stream>>>(*vw);
  return cudaGetLastError() == cudaSuccess ? 0 : OHP_ERR_LIB;
}
template <int DIM, class P> int launch_jacobian(const ohp_assembly_view *vw) {
  if (vw->n_rows <= 0) return 0;
  bool used = false;
  const int rc = launch_patch<DIM, P, true>(vw, &used);
  if (rc || used) return rc;
  k_jacobian_rows<DIM, P><<<(vw->n_rows + kAsmThreads - 1) / kAsmThreads, kAsmThreads, kAsmAccCap * sizeof(double),
                            (cudaStream_t)vw->stream>>>(*vw);
  return cudaGetLastError() == cudaSuccess ? 0 : OHP_ERR_LIB;
}
template <int DIM, class P> int launch_project_exact(const ohp_assembly_view *vw) {
  if (vw->n_rows <= 0) return 0;
  const int bs = 256;
  k_project_exact<DIM, P><<<(vw->n_rows + bs - 1) / bs, bs, 0, (cudaStream_t)vw->stream>>>(*vw);
  return cudaGetLastError() == cudaSuccess ? 0 : OHP_ERR_LIB;
}

template <int DIM, class P> int launch_l2_diff(const ohp_assembly_view *vw) {
  if (vw->n_partials <= 0) return OHP_ERR_ARG_SIZ;
  k_l2_diff<DIM, P><<<vw->n_partials, 256, 0, (cudaStream_t)vw->stream>>>(*vw);
  return cudaGetLastError() == cudaSuccess ? 0 : OHP_ERR_LIB;
}
template <class P> int launch_l2_diff_any(const ohp_assembly_view *vw) {
  return vw->dim == 2 ? launch_l2_diff<2, P>(vw) : vw->dim == 3 ? launch_l2_diff<3, P>(vw) : OHP_ERR_SUP;
}
template <class F> int launch_project_fn(const ohp_assembly_view *vw) {
  if (vw->n_rows <= 0) return 0;
  const int bs = 256, grid = (vw->n_rows + bs - 1) / bs;
  if (vw->dim == 2) k_project_fn<2, F><<<grid, bs, 0, (cudaStream_t)vw->stream>>>(*vw);
  else if (vw->dim == 3) k_project_fn<3, F><<<grid, bs, 0, (cudaStream_t)vw->stream>>>(*vw);
  else return OHP_ERR_SUP;
  return cudaGetLastError() == cudaSuccess ? 0 : OHP_ERR_LIB;
}

/* ------------------------------------------------------------------------------------ */
/* Natural boundary residual (-bc_type neumann, ex12.cpp:1226-1238).    */
/* [PETSc-internal DMPlexComputeBdResidual_Internal -> PetscFEIntegrateBdResidual_Basic]:    */
/* over every facet F with one supporting cell (label "boundary", ex12.cpp:1088-1094), with   */
/* the face rule of PetscFECreateDefault (2 Gauss points on an edge, the 4-point conical rule  */
/* on a triangle; weights sum to 2, detJ_F = |F| / 2):                                        */
/*     Fe[b] += w_q detJ_F ( phi_b(x_q) f0_bd(u, grad u, x_q, n) + grad(phi_b) . f1_bd(...) )  */
/* n = unit normal pointing out of the supporting cell.  "Owner computes" like the volume      */
/* kernels: the thread of row v walks star(v); a facet of a star cell that contains v is on    */
/* the boundary iff no other cell of star(v) contains it (any cell sharing the facet contains  */
/* v, so the star is the whole support -- also on a partitioned mesh, where every rank holds   */
/* all cells touching the vertices it owns).  Contributions are added in star order: no        */
/* atomics, bitwise reproducible.  Only rows of "marker" vertices do any work.                 */
/* The per-row function is __host__ __device__ and free of CUDA intrinsics so that the CPU     */
/* unit test (tests/host/bd_host.cu) can run exactly this code on host arrays.                 */
/* ------------------------------------------------------------------------------------ */
#define OHP_BD_ARGS                                                                              \
  int dim, int Nf, int NfAux, const int uOff[], const int uOff_x[], const double u[],            \
      const double u_t[], const double u_x[], const int aOff[], const int aOff_x[],              \
      const double a[], const double a_t[], const double a_x[], double t, const double x[],      \
      const double n[], int numConstants, const double constants[]

#if defined(__CUDA_ARCH__)
#define OHP_LD(p) __ldg(p)
#else
#define OHP_LD(p) (*(p))
#endif

/* barycentric coordinates of the face quadrature points with respect to the face's vertices, and the weights */
template <int DIM> struct FaceQuadrature;
template <> struct FaceQuadrature<2> {
  static constexpr int NQ = 2;
  OHP_HD static void lam(int q, double (&l)[2]) {
    const double s = (q == 0) ? -5.77350269189625764509e-01 : 5.77350269189625764509e-01; /* Gauss-Legendre on [-1,1] */
    l[0] = 0.5 * (1.0 - s); l[1] = 0.5 * (1.0 + s);
  }
  OHP_HD static double w(int) { return 1.0; }
};
template <> struct FaceQuadrature<3> {
  static constexpr int NQ = Quadrature<2>::NQ;
  OHP_HD static void lam(int q, double (&l)[3]) {
    l[1] = 0.5 * (1.0 + Quadrature<2>::xi(q, 0)); l[2] = 0.5 * (1.0 + Quadrature<2>::xi(q, 1)); l[0] = 1.0 - l[1] - l[2];
  }
  OHP_HD static double w(int q) { return Quadrature<2>::w(q); }
};

template <int DIM, class P>
OHP_HD double bd_residual_row(const ohp_assembly_view &vw, int k0, int k1) {
  constexpr int NC = DIM + 1;
  const int uOff[2] = {0, 1}, uOff_x[2] = {0, DIM};
  double acc = 0.0;
  for (int k = k0; k < k1; ++k) {
    const int e = OHP_LD(vw.v2c_dev + k), c = e >> 2, b = e & 3;
    int cv[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) cv[j] = OHP_LD(vw.cells_dev + (size_t)c * NC + j);
    for (int skip = 0; skip < NC; ++skip) {
      if (skip == b) continue; /* the facet opposite corner `skip` contains this row's vertex iff skip != b */
      bool shared = false;
      for (int k2 = k0; k2 < k1 && !shared; ++k2) {
        if (k2 == k) continue;
        const int c2 = OHP_LD(vw.v2c_dev + k2) >> 2;
        int ov[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) ov[j] = OHP_LD(vw.cells_dev + (size_t)c2 * NC + j);
        bool all = true;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          bool found = (j == skip);
#pragma unroll
          for (int i = 0; i < NC; ++i) found |= (ov[i] == cv[j]);
          all &= found;
        }
        shared = all;
      }
      if (shared) continue;
      /* closure of the supporting cell */
      double X[NC][DIM], ue[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) X[j][d] = OHP_LD(vw.coords_dev + (size_t)cv[j] * DIM + d);
        const int l = OHP_LD(vw.lidx_dev + cv[j]);
        if (l >= 0) ue[j] = OHP_LD(vw.u_dev + l);
        else { double val = 0.0; P::bc(DIM, 0.0, X[j], 1, &val, nullptr); ue[j] = val; }
      }
      CellGeom<DIM> g;
      cell_geometry<DIM>(X, g);
      double ux[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NC; ++j) s += g.G[j][d] * ue[j];
        ux[d] = s;
      }
      /* the face: corners of the cell in cell order, `skip` left out */
      int fj[DIM], nf = 0, av = 0;
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (j != skip) { if (j == b) av = nf; fj[nf++] = j; }
      double e1[DIM], n[DIM], meas;
#pragma unroll
      for (int d = 0; d < DIM; ++d) e1[d] = X[fj[1]][d] - X[fj[0]][d];
      if constexpr (DIM == 2) {
        meas = sqrt(e1[0] * e1[0] + e1[1] * e1[1]);
        n[0] = e1[1] / meas; n[1] = -e1[0] / meas;
      } else {
        double e2[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) e2[d] = X[fj[2]][d] - X[fj[0]][d];
        n[0] = e1[1] * e2[2] - e1[2] * e2[1]; n[1] = e1[2] * e2[0] - e1[0] * e2[2]; n[2] = e1[0] * e2[1] - e1[1] * e2[0];
        const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        meas = 0.5 * nn;
#pragma unroll
        for (int d = 0; d < DIM; ++d) n[d] /= nn;
      }
      double side = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) side += n[d] * (X[skip][d] - X[fj[0]][d]);
      if (side > 0.0) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) n[d] = -n[d];
      }
      const double detJf = 0.5 * meas;
      for (int q = 0; q < FaceQuadrature<DIM>::NQ; ++q) {
        double lam[DIM], xq[DIM], uq = 0.0;
        FaceQuadrature<DIM>::lam(q, lam);
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < DIM; ++a) s += lam[a] * X[fj[a]][d];
          xq[d] = s;
        }
#pragma unroll
        for (int a = 0; a < DIM; ++a) uq += lam[a] * ue[fj[a]];
        double f0[1] = {0.0}, f1[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) f1[d] = 0.0;
        P::f0_bd(DIM, 1, 0, uOff, uOff_x, &uq, nullptr, ux, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, xq, n, 0, nullptr, f0);
        P::f1_bd(DIM, 1, 0, uOff, uOff_x, &uq, nullptr, ux, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, xq, n, 0, nullptr, f1);
        double s = lam[av] * f0[0];
#pragma unroll
        for (int d = 0; d < DIM; ++d) s += g.G[b][d] * f1[d];
        acc += (detJf * FaceQuadrature<DIM>::w(q)) * s;
      }
    }
  }
  return acc;
}

template <int DIM, class P>
__global__ void __launch_bounds__(128) k_bd_residual_rows(const ohp_assembly_view vw) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= vw.n_rows) return;
  const int v = __ldg(vw.row_vertex_dev + r);
  if (!__ldg(vw.bnd_dev + v)) return;
  vw.out_dev[r] += bd_residual_row<DIM, P>(vw, __ldg(vw.v2c_ptr_dev + v), __ldg(vw.v2c_ptr_dev + v + 1));
}

template <int DIM, class P> int launch_bd_residual(const ohp_assembly_view *vw) {
  if (vw->n_rows <= 0) return 0;
  const int bs = 128;
  k_bd_residual_rows<DIM, P><<<(vw->n_rows + bs - 1) / bs, bs, 0, (cudaStream_t)vw->stream>>>(*vw);
  return cudaGetLastError() == cudaSuccess ? 0 : OHP_ERR_LIB;
}
template <class P> int launch_bd_residual_any(const ohp_assembly_view *vw) {
  return vw->dim == 2 ? launch_bd_residual<2, P>(vw) : vw->dim == 3 ? launch_bd_residual<3, P>(vw) : OHP_ERR_SUP;
}

/* dimension dispatch used by both the built-in problems and user registrations */
template <class P> int launch_residual_any(const ohp_assembly_view *vw) {
  return vw->dim == 2 ? launch_residual<2, P>(vw) : vw->dim == 3 ? launch_residual<3, P>(vw) : OHP_ERR_SUP;
}
template <class P> int launch_jacobian_any(const ohp_assembly_view *vw) {
  return vw->dim == 2 ? launch_jacobian<2, P>(vw) : vw->dim == 3 ? launch_jacobian<3, P>(vw) : OHP_ERR_SUP;
}
template <class P> int launch_project_exact_any(const ohp_assembly_view *vw) {
  return vw->dim == 2 ? launch_project_exact<2, P>(vw) : vw->dim == 3 ? launch_project_exact<3, P>(vw) : OHP_ERR_SUP;
}

}  // namespace ohp
#endif
