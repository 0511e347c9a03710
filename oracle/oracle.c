/*
 * ORACLE -- test infrastructure, NOT product code.
 *
 * Plain-C CPU restatement of the ex12 P1 Poisson hot path of cwsmith/omegahPetsc:
 * boundary marking, section/global numbering, CSR preallocation, FEM residual and
 * Jacobian integration through the PetscDS pointwise-callback ABI, KSP CG and the
 * SNES newtonls driver.  Each function cites the reference file:line it follows
 * (paths relative to /root/reference).  Rules that live inside PETSc itself
 * (un-vendored dependency, API level ~3.14, README.md:364-365 pins 7a013d0) are
 * restated from PETSc's published algorithm and marked [PETSc-internal].
 *
 * PARITY UNPINNED: the reference holds no golden vectors for this path (CTests
 * check exit status only, CMakeLists.txt:66-120) and cannot be built here (no
 * PETSc / Omega_h / MPI).  The oracle is anchored instead on (i) the manufactured
 * solution u = x^2+y^2, f = 4 (ex12.cpp:88-96,113-117,147-153), (ii) an independent
 * scipy.sparse assembly + solve in tests/, (iii) structural invariants (symmetry,
 * zero row sums of the unconstrained operator, A u + F(0) = F(u), ex12.cpp:1651-1661).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg may load this library.  The product never links or calls it.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXDIM 3
#define ORC_MAXNC 4
#define ORC_MAXQ 8

/* ------------------------------------------------------------------------- */
/* Pointwise-callback ABI (ex12.cpp:147-150 residual form, :209-212 Jacobian  */
/* form, :113 boundary/exact form).  PetscInt = int32 (hazard H4), scalar=f64 */
/* ------------------------------------------------------------------------- */
typedef void (*orc_pointfn)(int dim, int Nf, int NfAux, const int uOff[], const int uOff_x[],
                            const double u[], const double u_t[], const double u_x[],
                            const int aOff[], const int aOff_x[], const double a[],
                            const double a_t[], const double a_x[], double t, const double x[],
                            int numConstants, const double constants[], double f[]);
typedef void (*orc_pointjac)(int dim, int Nf, int NfAux, const int uOff[], const int uOff_x[],
                             const double u[], const double u_t[], const double u_x[],
                             const int aOff[], const int aOff_x[], const double a[],
                             const double a_t[], const double a_x[], double t, double u_tShift,
                             const double x[], int numConstants, const double constants[],
                             double g[]);
typedef int (*orc_bcfn)(int dim, double time, const double x[], int Nc, double *u, void *ctx);

#define PF_ARGS                                                                                  \
  int dim, int Nf, int NfAux, const int uOff[], const int uOff_x[], const double u[],            \
      const double u_t[], const double u_x[], const int aOff[], const int aOff_x[],              \
      const double a[], const double a_t[], const double a_x[], double t, const double x[],      \
      int numConstants, const double constants[]
#define PJ_ARGS                                                                                  \
  int dim, int Nf, int NfAux, const int uOff[], const int uOff_x[], const double u[],            \
      const double u_t[], const double u_x[], const int aOff[], const int aOff_x[],              \
      const double a[], const double a_t[], const double a_x[], double t, double u_tShift,       \
      const double x[], int numConstants, const double constants[]

/* constant-coefficient Poisson: ex12.cpp:147-153 (f0_u), :198-205 (f1_u), :209-216 (g3_uu) */
static void cb_f0_const(PF_ARGS, double f0[]) { f0[0] = 4.0; }
static void cb_f1_grad(PF_ARGS, double f1[]) {
  for (int d = 0; d < dim; ++d) f1[d] = u_x[d];
}
static void cb_g3_ident(PJ_ARGS, double g3[]) {
  for (int d = 0; d < dim; ++d) g3[d * dim + d] = 1.0;
}
/* analytic coefficient nu = x+y: ex12.cpp:283-289, :292-299, :313-320 */
static void cb_f0_analytic(PF_ARGS, double f0[]) { f0[0] = 6.0 * (x[0] + x[1]); }
static void cb_f1_analytic(PF_ARGS, double f1[]) {
  for (int d = 0; d < dim; ++d) f1[d] = (x[0] + x[1]) * u_x[d];
}
static void cb_g3_analytic(PJ_ARGS, double g3[]) {
  for (int d = 0; d < dim; ++d) g3[d * dim + d] = x[0] + x[1];
}
/* p-Laplacian (p=4): ex12.cpp:341-347, :350-359, :369-383 */
static void cb_f0_nonlinear(PF_ARGS, double f0[]) { f0[0] = 16.0 * (x[0] * x[0] + x[1] * x[1]); }
static void cb_f1_nonlinear(PF_ARGS, double f1[]) {
  double nu = 0.0;
  for (int d = 0; d < dim; ++d) nu += u_x[d] * u_x[d];
  for (int d = 0; d < dim; ++d) f1[d] = 0.5 * nu * u_x[d];
}
static void cb_g3_nonlinear(PJ_ARGS, double g3[]) {
  double nu = 0.0;
  for (int d = 0; d < dim; ++d) nu += u_x[d] * u_x[d];
  for (int d = 0; d < dim; ++d) {
    g3[d * dim + d] = 0.5 * nu;
    for (int e = 0; e < dim; ++e) g3[d * dim + e] += u_x[d] * u_x[e];
  }
}
/* circle / cross forcing terms: ex12.cpp:155-166 (f0_circle_u), :168-177 (f0_cross_u); flux and g3 are f1_u, g3_uu */
static void cb_f0_circle(PF_ARGS, double f0[]) {
  const double alpha = 500., radius2 = 0.15 * 0.15;
  const double r2 = (x[0] - 0.5) * (x[0] - 0.5) + (x[1] - 0.5) * (x[1] - 0.5);
  const double xi = alpha * (radius2 - r2);
  f0[0] = (-4.0 * alpha - 8.0 * alpha * alpha * r2 * tanh(xi)) * (1.0 / cosh(xi)) * (1.0 / cosh(xi));
}
static void cb_f0_cross(PF_ARGS, double f0[]) {
  const double alpha = 50 * 4;
  const double xy = (x[0] - 0.5) * (x[1] - 0.5);
  f0[0] = sin(alpha * xy) * (alpha * fabs(xy) < 2 * M_PI ? 1 : 0.01);
}
/* their Dirichlet / exact data: ex12.cpp:127-136 (circle_u_2d), :138-145 (cross_u_2d) */
static int cb_circle_2d(int dim, double time, const double x[], int Nc, double *u, void *ctx) {
  (void)dim; (void)time; (void)Nc; (void)ctx;
  const double alpha = 500., radius2 = 0.15 * 0.15;
  const double r2 = (x[0] - 0.5) * (x[0] - 0.5) + (x[1] - 0.5) * (x[1] - 0.5);
  *u = tanh(alpha * (radius2 - r2)) + 1.0;
  return 0;
}
static int cb_cross_2d(int dim, double time, const double x[], int Nc, double *u, void *ctx) {
  (void)dim; (void)time; (void)Nc; (void)ctx;
  const double alpha = 50 * 4;
  const double xy = (x[0] - 0.5) * (x[1] - 0.5);
  *u = sin(alpha * xy) * (alpha * fabs(xy) < 2 * M_PI ? 1 : 0.01);
  return 0;
}
/* Dirichlet / exact data: ex12.cpp:113-117 (2-D), :408-412 (3-D); guess :76-80 */
static int cb_quadratic_2d(int dim, double time, const double x[], int Nc, double *u, void *ctx) {
  (void)dim; (void)time; (void)Nc; (void)ctx;
  *u = x[0] * x[0] + x[1] * x[1];
  return 0;
}
static int cb_quadratic_3d(int dim, double time, const double x[], int Nc, double *u, void *ctx) {
  (void)dim; (void)time; (void)Nc; (void)ctx;
  *u = 2.0 * (x[0] * x[0] + x[1] * x[1] + x[2] * x[2]) / 3.0;
  return 0;
}

typedef struct {
  orc_pointfn f0, f1;
  orc_pointjac g3;
  orc_bcfn bc;
} orc_problem;

/* problem ids mirror SetupProblem's switch (ex12.cpp:1170-1204): 0 COEFF_NONE, 1 ANALYTIC, 2 NONLINEAR, 3 CIRCLE,
 * 4 CROSS; the exact solution / Dirichlet function follows ex12.cpp:1206-1225 */
static int orc_problem_get(int id, int dim, orc_problem *p) {
  p->bc = (dim == 3) ? cb_quadratic_3d : cb_quadratic_2d; /* ex12.cpp:1222,1230 */
  switch (id) {
    case 0: p->f0 = cb_f0_const; p->f1 = cb_f1_grad; p->g3 = cb_g3_ident; return 0;
    case 1: p->f0 = cb_f0_analytic; p->f1 = cb_f1_analytic; p->g3 = cb_g3_analytic; return 0;
    case 2: p->f0 = cb_f0_nonlinear; p->f1 = cb_f1_nonlinear; p->g3 = cb_g3_nonlinear; return 0;
    case 3: p->f0 = cb_f0_circle; p->f1 = cb_f1_grad; p->g3 = cb_g3_ident; if (dim == 2) p->bc = cb_circle_2d; return 0;
    case 4: p->f0 = cb_f0_cross; p->f1 = cb_f1_grad; p->g3 = cb_g3_ident; if (dim == 2) p->bc = cb_cross_2d; return 0;
  }
  return 1;
}

/* ------------------------------------------------------------------------- */
/* Quadrature [PETSc-internal: PetscFECreateDefault(qorder=-1) at             */
/* ex12.cpp:1309 -> PetscDTStroudConicalQuadrature, 2 Gauss-Jacobi points per */
/* direction, biunit simplex].  Gauss-Jacobi(alpha,0) two-point rules are     */
/* derived here from the moments of (1-x)^alpha on [-1,1].                    */
/* ------------------------------------------------------------------------- */
static void gauss_jacobi2(int alpha, double x[2], double w[2]) {
  /* m_k = int_{-1}^{1} (1-x)^alpha x^k dx, substitution s = 1-x in [0,2] */
  double m[4];
  for (int k = 0; k < 4; ++k) {
    /* int_0^2 s^alpha (1-s)^k ds = sum_j C(k,j)(-1)^j 2^(alpha+j+1)/(alpha+j+1) */
    double acc = 0.0, binom = 1.0;
    for (int j = 0; j <= k; ++j) {
      if (j > 0) binom = binom * (k - j + 1) / j;
      acc += ((j & 1) ? -1.0 : 1.0) * binom * ldexp(1.0, alpha + j + 1) / (alpha + j + 1);
    }
    m[k] = acc;
  }
  /* monic orthogonal quadratic x^2 + b x + c:  [m1 m0; m2 m1] [b c]^T = -[m2 m3]^T */
  double det = m[1] * m[1] - m[0] * m[2];
  double b = (-m[2] * m[1] + m[0] * m[3]) / det;
  double c = (-m[1] * m[3] + m[2] * m[2]) / det;
  double disc = sqrt(b * b - 4.0 * c);
  x[0] = 0.5 * (-b - disc);
  x[1] = 0.5 * (-b + disc);
  w[1] = (m[1] - x[0] * m[0]) / (x[1] - x[0]);
  w[0] = m[0] - w[1];
}

int orc_quadrature(int dim, int *nq, double *xi /* nq*dim */, double *w) {
  double px[2], wx[2], py[2], wy[2], pz[2], wz[2];
  gauss_jacobi2(0, px, wx);
  gauss_jacobi2(1, py, wy);
  gauss_jacobi2(2, pz, wz);
  if (dim == 2) {
    *nq = 4;
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) {
        int q = i * 2 + j;
        xi[q * 2 + 0] = 0.5 * (1.0 + px[i]) * (1.0 - py[j]) - 1.0;
        xi[q * 2 + 1] = py[j];
        w[q] = 0.5 * wx[i] * wy[j];
      }
    return 0;
  }
  if (dim == 3) {
    *nq = 8;
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j)
        for (int k = 0; k < 2; ++k) {
          int q = (i * 2 + j) * 2 + k;
          xi[q * 3 + 0] = 0.25 * (1.0 + px[i]) * (1.0 - py[j]) * (1.0 - pz[k]) - 1.0;
          xi[q * 3 + 1] = 0.5 * (1.0 + py[j]) * (1.0 - pz[k]) - 1.0;
          xi[q * 3 + 2] = pz[k];
          w[q] = 0.125 * wx[i] * wy[j] * wz[k];
        }
    return 0;
  }
  return 1;
}

/* ------------------------------------------------------------------------- */
/* vertex -> cell adjacency (support of the Plex vertices after               */
/* DMPlexCreateFromCellListPetsc, ex12.cpp:859,933) [PETSc-internal]          */
/* ------------------------------------------------------------------------- */
static void build_v2c(int nc, int64_t nC, int nV, const int *cells, int64_t **ptr_out, int **adj_out) {
  int64_t *ptr = (int64_t *)calloc((size_t)nV + 1, sizeof(int64_t));
  for (int64_t i = 0; i < nC * nc; ++i) ptr[cells[i] + 1]++;
  for (int v = 0; v < nV; ++v) ptr[v + 1] += ptr[v];
  int *adj = (int *)malloc((size_t)(nC * nc) * sizeof(int));
  int64_t *cur = (int64_t *)malloc((size_t)nV * sizeof(int64_t));
  memcpy(cur, ptr, (size_t)nV * sizeof(int64_t));
  for (int64_t c = 0; c < nC; ++c)
    for (int j = 0; j < nc; ++j) adj[cur[cells[c * nc + j]]++] = (int)c;
  free(cur);
  *ptr_out = ptr;
  *adj_out = adj;
}

/* Boundary marking: a facet (edge in 2-D, face in 3-D) is on the boundary iff its
 * support size is 1 (ex12.cpp:986-997); boundary vertices are the union of the cones
 * of those facets (ex12.cpp:1001-1010), labelled "marker"=1 (ex12.cpp:1023-1026).
 * Evaluated on the GLOBAL mesh (the reference's per-rank count is hazard H2). */
int orc_mark_boundary(int dim, int64_t nC, int nV, const int *cells, uint8_t *bnd) {
  const int nc = dim + 1;
  int64_t *ptr; int *adj;
  build_v2c(nc, nC, nV, cells, &ptr, &adj);
  memset(bnd, 0, (size_t)nV);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nC; ++c) {
    const int *cv = cells + c * nc;
    for (int skip = 0; skip < nc; ++skip) { /* facet = cell minus vertex `skip` */
      int fv[ORC_MAXNC], nf = 0;
      for (int j = 0; j < nc; ++j) if (j != skip) fv[nf++] = cv[j];
      int support = 0;
      for (int64_t k = ptr[fv[0]]; k < ptr[fv[0] + 1]; ++k) {
        const int *ov = cells + (int64_t)adj[k] * nc;
        int all = 1;
        for (int a = 1; a < nf && all; ++a) {
          int found = 0;
          for (int j = 0; j < nc; ++j) found |= (ov[j] == fv[a]);
          all = found;
        }
        support += all;
      }
      if (support == 1)
        for (int a = 0; a < nf; ++a) bnd[fv[a]] = 1; /* benign same-value race */
    }
  }
  free(ptr); free(adj);
  return 0;
}

/* Global section [PETSc-internal, reached from DMCreateGlobalVector ex12.cpp:1505]:
 * one dof per vertex (P1, ex12.cpp:1309), constrained on "marker" vertices
 * (DM_BC_ESSENTIAL, ex12.cpp:1237); unconstrained dofs numbered consecutively in
 * ascending local point order.  gidx[v] = -1 for constrained vertices. */
int orc_numbering(int nV, const uint8_t *bnd, int *gidx) {
  int n = 0;
  for (int v = 0; v < nV; ++v) gidx[v] = bnd[v] ? -1 : n++;
  return n;
}

/* CSR preallocation [PETSc-internal DMPlexPreallocateOperator, reached from
 * DMCreateMatrix ex12.cpp:1508]: row(v) holds col(w) iff v,w unconstrained and share a
 * cell (closure(star(v))), self included; columns ascending; explicit zeros kept.
 * Call with colidx == NULL to size (returns nnz, fills rowptr). */
int64_t orc_pattern(int dim, int64_t nC, int nV, const int *cells, const int *gidx, int N,
                    int *rowptr, int *colidx) {
  const int nc = dim + 1;
  int64_t *ptr; int *adj;
  build_v2c(nc, nC, nV, cells, &ptr, &adj);
  int *v_of_row = (int *)malloc((size_t)(N > 0 ? N : 1) * sizeof(int));
  for (int v = 0; v < nV; ++v) if (gidx[v] >= 0) v_of_row[gidx[v]] = v;
  int fill = (colidx != NULL);
  if (!fill) rowptr[0] = 0;
#pragma omp parallel
  {
    int cap = 64, *buf = (int *)malloc((size_t)cap * sizeof(int));
#pragma omp for schedule(static)
    for (int r = 0; r < N; ++r) {
      int v = v_of_row[r], n = 0;
      int64_t need = (ptr[v + 1] - ptr[v]) * nc;
      if (need > cap) { cap = (int)need * 2; buf = (int *)realloc(buf, (size_t)cap * sizeof(int)); }
      for (int64_t k = ptr[v]; k < ptr[v + 1]; ++k) {
        const int *ov = cells + (int64_t)adj[k] * nc;
        for (int j = 0; j < nc; ++j) {
          int g = gidx[ov[j]];
          if (g < 0) continue;
          int pos = 0; /* sorted insert, unique */
          while (pos < n && buf[pos] < g) ++pos;
          if (pos < n && buf[pos] == g) continue;
          for (int t = n; t > pos; --t) buf[t] = buf[t - 1];
          buf[pos] = g; ++n;
        }
      }
      if (!fill) rowptr[r + 1] = n;
      else memcpy(colidx + rowptr[r], buf, (size_t)n * sizeof(int));
    }
    free(buf);
  }
  if (!fill) for (int r = 0; r < N; ++r) rowptr[r + 1] += rowptr[r];
  free(ptr); free(adj); free(v_of_row);
  return rowptr[N];
}

/* ------------------------------------------------------------------------- */
/* Cell geometry + P1 tabulation [PETSc-internal DMPlexComputeCellGeometryFEM */
/* and PetscFE tabulation on the biunit simplex]                              */
/* ------------------------------------------------------------------------- */
typedef struct {
  double v0[ORC_MAXDIM], J[ORC_MAXDIM * ORC_MAXDIM], invJ[ORC_MAXDIM * ORC_MAXDIM], detJ;
  double G[ORC_MAXNC][ORC_MAXDIM]; /* real-space gradients of the P1 basis */
} cell_geom;

static void cell_geometry(int dim, const double *X[ORC_MAXNC], cell_geom *g) {
  for (int d = 0; d < dim; ++d) {
    g->v0[d] = X[0][d];
    for (int f = 0; f < dim; ++f) g->J[d * dim + f] = 0.5 * (X[f + 1][d] - X[0][d]);
  }
  const double *J = g->J;
  double *I = g->invJ;
  if (dim == 2) {
    g->detJ = J[0] * J[3] - J[1] * J[2];
    double id = 1.0 / g->detJ;
    I[0] = id * J[3]; I[1] = -id * J[1]; I[2] = -id * J[2]; I[3] = id * J[0];
  } else {
    g->detJ = J[0] * (J[4] * J[8] - J[5] * J[7]) + J[1] * (J[5] * J[6] - J[3] * J[8]) +
              J[2] * (J[3] * J[7] - J[4] * J[6]);
    double id = 1.0 / g->detJ;
    I[0] = id * (J[4] * J[8] - J[5] * J[7]); I[1] = id * (J[2] * J[7] - J[1] * J[8]); I[2] = id * (J[1] * J[5] - J[2] * J[4]);
    I[3] = id * (J[5] * J[6] - J[3] * J[8]); I[4] = id * (J[0] * J[8] - J[2] * J[6]); I[5] = id * (J[2] * J[3] - J[0] * J[5]);
    I[6] = id * (J[3] * J[7] - J[4] * J[6]); I[7] = id * (J[1] * J[6] - J[0] * J[7]); I[8] = id * (J[0] * J[4] - J[1] * J[3]);
  }
  /* Dref[b][e]: row 0 = -1/2 each, row b>0 = +1/2 at e=b-1;  G[b][d] = sum_e Dref[b][e] invJ[e][d] */
  for (int d = 0; d < dim; ++d) {
    double s = 0.0;
    for (int e = 0; e < dim; ++e) s += I[e * dim + d];
    g->G[0][d] = -0.5 * s;
    for (int b = 1; b <= dim; ++b) g->G[b][d] = 0.5 * I[(b - 1) * dim + d];
  }
}

static inline void p1_basis(int dim, const double *xi, double *phi) {
  double s = 0.0;
  for (int e = 0; e < dim; ++e) { phi[e + 1] = 0.5 * (1.0 + xi[e]); s += xi[e]; }
  phi[0] = -0.5 * (s + dim - 2.0);
}

/* local closure values: uloc = G2L(u) with essential BC values inserted
 * [PETSc-internal DMPlexSNESComputeBoundaryFEM -> DMPlexInsertBoundaryValues], BC
 * function called with the vertex coordinates (registered at ex12.cpp:1237-1239). */
static void local_values(const orc_problem *p, int dim, int nV, const double *coords, const int *gidx,
                         const double *u, double *uloc) {
#pragma omp parallel for schedule(static)
  for (int v = 0; v < nV; ++v) {
    if (gidx[v] >= 0) uloc[v] = u[gidx[v]];
    else p->bc(dim, 0.0, coords + (size_t)v * dim, 1, &uloc[v], NULL);
  }
}

/* Residual [PETSc-internal DMPlexSNESComputeResidualFEM ->
 * PetscFEIntegrateResidual_Basic, installed by DMPlexSetSNESLocalFEM ex12.cpp:1540] */
int orc_residual(int problem, int dim, int64_t nC, int nV, const int *cells, const double *coords,
                 const int *gidx, int N, const double *u, double *F) {
  orc_problem p;
  if (orc_problem_get(problem, dim, &p)) return 1;
  const int nc = dim + 1;
  int nq; double xi[ORC_MAXQ * ORC_MAXDIM], wq[ORC_MAXQ];
  if (orc_quadrature(dim, &nq, xi, wq)) return 1;
  double *uloc = (double *)malloc((size_t)nV * sizeof(double));
  local_values(&p, dim, nV, coords, gidx, u, uloc);
  for (int i = 0; i < N; ++i) F[i] = 0.0;
  const int uOff[2] = {0, 1}, uOff_x[2] = {0, dim};
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nC; ++c) {
    const int *cv = cells + c * nc;
    const double *X[ORC_MAXNC];
    double ue[ORC_MAXNC], Fe[ORC_MAXNC] = {0, 0, 0, 0};
    for (int b = 0; b < nc; ++b) { X[b] = coords + (size_t)cv[b] * dim; ue[b] = uloc[cv[b]]; }
    cell_geom g;
    cell_geometry(dim, X, &g);
    for (int q = 0; q < nq; ++q) {
      double phi[ORC_MAXNC], xq[ORC_MAXDIM], uq = 0.0, ux[ORC_MAXDIM] = {0, 0, 0};
      p1_basis(dim, xi + q * dim, phi);
      for (int d = 0; d < dim; ++d) {
        xq[d] = g.v0[d];
        for (int f = 0; f < dim; ++f) xq[d] += g.J[d * dim + f] * (xi[q * dim + f] + 1.0);
      }
      for (int b = 0; b < nc; ++b) {
        uq += phi[b] * ue[b];
        for (int d = 0; d < dim; ++d) ux[d] += g.G[b][d] * ue[b];
      }
      double f0[1] = {0.0}, f1[ORC_MAXDIM] = {0, 0, 0};
      p.f0(dim, 1, 0, uOff, uOff_x, &uq, NULL, ux, NULL, NULL, NULL, NULL, NULL, 0.0, xq, 0, NULL, f0);
      p.f1(dim, 1, 0, uOff, uOff_x, &uq, NULL, ux, NULL, NULL, NULL, NULL, NULL, 0.0, xq, 0, NULL, f1);
      const double w = g.detJ * wq[q];
      f0[0] *= w;
      for (int d = 0; d < dim; ++d) f1[d] *= w;
      for (int b = 0; b < nc; ++b) {
        double s = phi[b] * f0[0];
        for (int d = 0; d < dim; ++d) s += g.G[b][d] * f1[d];
        Fe[b] += s;
      }
    }
    for (int b = 0; b < nc; ++b) { /* L2G(ADD): constrained rows dropped */
      int r = gidx[cv[b]];
      if (r < 0) continue;
#pragma omp atomic
      F[r] += Fe[b];
    }
  }
  free(uloc);
  return 0;
}

static inline int64_t find_slot(const int *rowptr, const int *colidx, int r, int c) {
  int lo = rowptr[r], hi = rowptr[r + 1] - 1;
  while (lo <= hi) {
    int mid = (lo + hi) >> 1;
    if (colidx[mid] < c) lo = mid + 1; else if (colidx[mid] > c) hi = mid - 1; else return mid;
  }
  return -1;
}

/* Jacobian [PETSc-internal DMPlexSNESComputeJacobianFEM -> PetscFEIntegrateJacobian_Basic
 * -> DMPlexMatSetClosure / MatSetValues(ADD_VALUES)]; only g3 is registered
 * (PetscDSSetJacobian(prob,0,0,NULL,NULL,NULL,g3), ex12.cpp:1182); g3 is zeroed by the
 * integrator before each call (g3_uu only writes the diagonal, ex12.cpp:215). */
int orc_jacobian(int problem, int dim, int64_t nC, int nV, const int *cells, const double *coords,
                 const int *gidx, int N, const double *u, const int *rowptr, const int *colidx,
                 double *vals) {
  orc_problem p;
  if (orc_problem_get(problem, dim, &p)) return 1;
  const int nc = dim + 1;
  int nq; double xi[ORC_MAXQ * ORC_MAXDIM], wq[ORC_MAXQ];
  if (orc_quadrature(dim, &nq, xi, wq)) return 1;
  double *uloc = (double *)malloc((size_t)nV * sizeof(double));
  local_values(&p, dim, nV, coords, gidx, u, uloc);
  for (int64_t i = 0; i < rowptr[N]; ++i) vals[i] = 0.0;
  const int uOff[2] = {0, 1}, uOff_x[2] = {0, dim};
  int bad = 0;
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nC; ++c) {
    const int *cv = cells + c * nc;
    const double *X[ORC_MAXNC];
    double ue[ORC_MAXNC], Ke[ORC_MAXNC][ORC_MAXNC];
    memset(Ke, 0, sizeof(Ke));
    for (int b = 0; b < nc; ++b) { X[b] = coords + (size_t)cv[b] * dim; ue[b] = uloc[cv[b]]; }
    cell_geom g;
    cell_geometry(dim, X, &g);
    for (int q = 0; q < nq; ++q) {
      double phi[ORC_MAXNC], xq[ORC_MAXDIM], uq = 0.0, ux[ORC_MAXDIM] = {0, 0, 0};
      p1_basis(dim, xi + q * dim, phi);
      for (int d = 0; d < dim; ++d) {
        xq[d] = g.v0[d];
        for (int f = 0; f < dim; ++f) xq[d] += g.J[d * dim + f] * (xi[q * dim + f] + 1.0);
      }
      for (int b = 0; b < nc; ++b) {
        uq += phi[b] * ue[b];
        for (int d = 0; d < dim; ++d) ux[d] += g.G[b][d] * ue[b];
      }
      double g3[ORC_MAXDIM * ORC_MAXDIM] = {0};
      p.g3(dim, 1, 0, uOff, uOff_x, &uq, NULL, ux, NULL, NULL, NULL, NULL, NULL, 0.0, 0.0, xq, 0, NULL, g3);
      const double w = g.detJ * wq[q];
      for (int b = 0; b < nc; ++b)
        for (int cc = 0; cc < nc; ++cc) {
          double s = 0.0;
          for (int d = 0; d < dim; ++d)
            for (int e = 0; e < dim; ++e) s += g.G[b][d] * (w * g3[d * dim + e]) * g.G[cc][e];
          Ke[b][cc] += s;
        }
    }
    for (int b = 0; b < nc; ++b) {
      int r = gidx[cv[b]];
      if (r < 0) continue;
      for (int cc = 0; cc < nc; ++cc) {
        int col = gidx[cv[cc]];
        if (col < 0) continue;
        int64_t s = find_slot(rowptr, colidx, r, col);
        if (s < 0) { bad = 1; continue; }
#pragma omp atomic
        vals[s] += Ke[b][cc];
      }
    }
  }
  free(uloc);
  return bad;
}

/* ------------------------------------------------------------------------- */
/* Mat/Vec kernels [PETSc-internal MatMult_SeqAIJ, VecDot/Norm/AXPY/AYPX];     */
/* used directly at ex12.cpp:1655-1675 and inside KSP                         */
/* ------------------------------------------------------------------------- */
void orc_spmv(int N, const int *rowptr, const int *colidx, const double *vals, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) s += vals[k] * x[colidx[k]];
    y[i] = s;
  }
}
static double vdot(int N, const double *a, const double *b) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int i = 0; i < N; ++i) s += a[i] * b[i];
  return s;
}

/* KSP CG [PETSc-internal KSPSolve_CG, selected by -ksp_type cg at ex12.cpp:1543];
 * zero initial guess, preconditioned norm, KSPConvergedDefault (rtol/abstol/dtol).
 * pc: 0 = none, 1 = Jacobi.  reason codes follow KSPConvergedReason:
 * 2 RTOL, 3 ATOL, -3 DIVERGED_ITS, -4 DIVERGED_DTOL, -8 INDEFINITE_MAT, -9 NANORINF.
 * hist (optional, size maxit+1) receives the monitored norm at each iteration. */
int orc_cg(int N, const int *rowptr, const int *colidx, const double *vals, const double *b, double *x,
           double rtol, double abstol, double dtol, int maxit, int pc, int *its_out, int *reason_out,
           double *rnorm_out, double *hist) {
  double *r = (double *)malloc(sizeof(double) * (size_t)N), *z = (double *)malloc(sizeof(double) * (size_t)N);
  double *p = (double *)malloc(sizeof(double) * (size_t)N), *w = (double *)malloc(sizeof(double) * (size_t)N);
  double *dinv = NULL;
  if (pc == 1) {
    dinv = (double *)malloc(sizeof(double) * (size_t)N);
    for (int i = 0; i < N; ++i) {
      int64_t s = find_slot(rowptr, colidx, i, i);
      double d = (s >= 0) ? vals[s] : 0.0;
      dinv[i] = (d != 0.0) ? 1.0 / d : 1.0;
    }
  }
  int reason = 0, i = 0;
#pragma omp parallel for schedule(static)
  for (int k = 0; k < N; ++k) { x[k] = 0.0; r[k] = b[k]; z[k] = dinv ? dinv[k] * b[k] : b[k]; }
  double dp = sqrt(vdot(N, z, z));
  double beta = vdot(N, z, r), betaold = 1.0;
  const double rnorm0 = dp, ttol = fmax(rtol * rnorm0, abstol);
  if (hist) hist[0] = dp;
  if (dp != dp) reason = -9;
  else if (dp <= ttol) reason = (dp < abstol) ? 3 : 2;
  while (!reason && i < maxit) {
    if (beta == 0.0) { reason = 3; break; }
    if (i == 0) {
#pragma omp parallel for schedule(static)
      for (int k = 0; k < N; ++k) p[k] = z[k];
    } else {
      const double bb = beta / betaold;
#pragma omp parallel for schedule(static)
      for (int k = 0; k < N; ++k) p[k] = z[k] + bb * p[k];
    }
    orc_spmv(N, rowptr, colidx, vals, p, w);
    const double dpi = vdot(N, p, w);
    betaold = beta;
    if (!(dpi > 0.0)) { reason = (dpi != dpi) ? -9 : -8; ++i; break; }
    const double a = beta / dpi;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < N; ++k) { x[k] += a * p[k]; r[k] -= a * w[k]; z[k] = dinv ? dinv[k] * r[k] : r[k]; }
    dp = sqrt(vdot(N, z, z));
    ++i;
    if (hist) hist[i] = dp;
    if (dp != dp) reason = -9;
    else if (dp <= ttol) reason = (dp < abstol) ? 3 : 2;
    else if (dp >= dtol * rnorm0) reason = -4;
    if (reason) break;
    beta = vdot(N, z, r);
  }
  if (!reason) reason = -3;
  *its_out = i; *reason_out = reason; *rnorm_out = dp;
  free(r); free(z); free(p); free(w); free(dinv);
  return 0;
}

/* SNES newtonls with full step (bt line search accepts lambda = 1 when the sufficient
 * decrease test holds) [PETSc-internal SNESSolve_NEWTONLS, reached from SNESSolve
 * ex12.cpp:1609]; defaults snes_rtol 1e-8, atol 1e-50, stol 1e-8, max_it 50.
 * reason: 2 FNORM_ABS, 3 FNORM_RELATIVE, 4 SNORM_RELATIVE, -5 MAX_IT, -3 LINEAR_SOLVE.
 * out[0]=snes its, out[1]=total ksp its, out[2]=reason; norms[k] = ||F_k||. */
int orc_newton(int problem, int dim, int64_t nC, int nV, const int *cells, const double *coords,
               const int *gidx, int N, const int *rowptr, const int *colidx, double *u,
               double snes_rtol, double snes_atol, double snes_stol, int snes_maxit,
               double ksp_rtol, double ksp_atol, int ksp_maxit, int pc, int *out, double *norms) {
  double *F = (double *)malloc(sizeof(double) * (size_t)N), *y = (double *)malloc(sizeof(double) * (size_t)N);
  double *vals = (double *)malloc(sizeof(double) * (size_t)rowptr[N]);
  double *w = (double *)malloc(sizeof(double) * (size_t)N), *Fw = (double *)malloc(sizeof(double) * (size_t)N);
  int its = 0, kits = 0, reason = 0;
  orc_residual(problem, dim, nC, nV, cells, coords, gidx, N, u, F);
  double fnorm = sqrt(vdot(N, F, F));
  const double fnorm0 = fnorm;
  norms[0] = fnorm;
  if (fnorm < snes_atol) reason = 2;
  while (!reason && its < snes_maxit) {
    orc_jacobian(problem, dim, nC, nV, cells, coords, gidx, N, u, rowptr, colidx, vals);
    int k, kr; double kn;
    orc_cg(N, rowptr, colidx, vals, F, y, ksp_rtol, ksp_atol, 1e5, ksp_maxit, pc, &k, &kr, &kn, NULL);
    kits += k;
    if (kr < 0) { reason = -3; break; }
    /* backtracking on lambda: 0.5 g^2 <= 0.5 f^2 + lambda*alpha*initslope, alpha = 1e-4 */
    orc_spmv(N, rowptr, colidx, vals, y, w);
    double initslope = -vdot(N, F, w);
    if (initslope > 0.0) initslope = -initslope;
    if (initslope == 0.0) initslope = -1.0;
    double lambda = 1.0, gnorm = 0.0;
    for (int bt = 0; bt < 40; ++bt) {
      for (int i = 0; i < N; ++i) w[i] = u[i] - lambda * y[i];
      orc_residual(problem, dim, nC, nV, cells, coords, gidx, N, w, Fw);
      gnorm = sqrt(vdot(N, Fw, Fw));
      if (0.5 * gnorm * gnorm <= 0.5 * fnorm * fnorm + lambda * 1e-4 * initslope) break;
      lambda *= 0.5;
    }
    double ynorm = lambda * sqrt(vdot(N, y, y));
    memcpy(u, w, sizeof(double) * (size_t)N);
    memcpy(F, Fw, sizeof(double) * (size_t)N);
    fnorm = gnorm;
    ++its;
    norms[its] = fnorm;
    double xnorm = sqrt(vdot(N, u, u));
    if (fnorm < snes_atol) reason = 2;
    else if (fnorm <= snes_rtol * fnorm0) reason = 3;
    else if (ynorm < snes_stol * xnorm) reason = 4;
  }
  if (!reason) reason = -5;
  out[0] = its; out[1] = kits; out[2] = reason;
  free(F); free(y); free(vals); free(w); free(Fw);
  return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its children; the CPU baseline must still use every host core */
int orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

/* ========================================================================= */
/* natural (Neumann) boundary term, null space, a          */
/* polynomial preconditioner.  Same status as the rest of this file: test       */
/* infrastructure, parity unpinned.                                             */
/* ========================================================================= */

/* boundary pointwise callbacks: f0_bd_u (ex12.cpp:179-186), f1_bd_zero (ex12.cpp:188-195); the argument list of the
 * volume callbacks plus the unit outward normal n[] after x[] */
typedef void (*orc_bdfn)(int dim, int Nf, int NfAux, const int uOff[], const int uOff_x[], const double u[],
                         const double u_t[], const double u_x[], const int aOff[], const int aOff_x[], const double a[],
                         const double a_t[], const double a_x[], double t, const double x[], const double n[],
                         int numConstants, const double constants[], double f[]);
#define BD_ARGS                                                                                  \
  int dim, int Nf, int NfAux, const int uOff[], const int uOff_x[], const double u[],            \
      const double u_t[], const double u_x[], const int aOff[], const int aOff_x[],              \
      const double a[], const double a_t[], const double a_x[], double t, const double x[],      \
      const double n[], int numConstants, const double constants[]
static void cb_f0_bd_u(BD_ARGS, double f0[]) {
  f0[0] = 0.0;
  for (int d = 0; d < dim; ++d) f0[0] += -n[d] * 2.0 * x[d];
}
static void cb_f1_bd_zero(BD_ARGS, double f1[]) {
  for (int c = 0; c < dim; ++c) f1[c] = 0.0;
}
/* a second, synthetic boundary callback set (not in ex12): reads EVERY argument the integrator supplies -- u, grad u, x, n --
 * and has a non-zero f1 term, so that the checks of the boundary integrator do not rest on the one callback ex12 ships */
static void cb_f0_bd_probe(BD_ARGS, double f0[]) {
  f0[0] = 0.5 * u[0];
  for (int d = 0; d < dim; ++d) f0[0] += n[d] * u_x[d] + (d + 1) * x[d] * n[d];
}
static void cb_f1_bd_probe(BD_ARGS, double f1[]) {
  for (int d = 0; d < dim; ++d) f1[d] = u[0] * n[d] + 0.25 * x[d];
}

/* Natural boundary residual [PETSc-internal DMPlexComputeBdResidual_Internal -> PetscFEIntegrateBdResidual_Basic,
 * reached from DMPlexSetSNESLocalFEM ex12.cpp:1540 when SetupProblem registered PetscDSSetBdResidual(prob, 0, f0_bd_u,
 * f1_bd_zero) (ex12.cpp:1226,1231) and DMAddBoundary(DM_BC_NATURAL, "wall", "boundary", ...) (ex12.cpp:1237)].
 * Label "boundary" = the facets with one supporting cell (DMPlexMarkBoundaryFaces, ex12.cpp:1088-1094).  Per boundary
 * facet F of cell c, per face quadrature point q (face rule of PetscFECreateDefault: 2 Gauss points on an edge, the
 * 4-point conical rule on a triangle; weights sum to 2):
 *     Fe[b] += w_q detJ_F ( phi_b(x_q) f0_bd(u, grad u, x_q, n) + grad(phi_b) . f1_bd(...) ),   detJ_F = |F| / 2,
 * phi_b = the cell's basis restricted to the face, n = unit normal pointing out of c.  ADDS into F (call after
 * orc_residual); constrained rows (gidx < 0) are dropped like everywhere else.  bd_set 0 = ex12's callbacks, 1 = the
 * synthetic probe set above. */
int orc_bd_residual(int bd_set, int dim, int64_t nC, int nV, const int *cells, const double *coords, const int *gidx, int N,
                    const double *u, double *F) {
  (void)N;
  const int nc = dim + 1, nf = dim;
  orc_problem p;
  if (orc_problem_get(0, dim, &p)) return 1;
  const orc_bdfn f0_bd = bd_set ? cb_f0_bd_probe : cb_f0_bd_u, f1_bd = bd_set ? cb_f1_bd_probe : cb_f1_bd_zero;
  int64_t *ptr; int *adj;
  build_v2c(nc, nC, nV, cells, &ptr, &adj);
  double *uloc = (double *)malloc((size_t)nV * sizeof(double));
  local_values(&p, dim, nV, coords, gidx, u, uloc);
  /* face rule on the biunit (dim-1)-simplex */
  int nq = 0; double fxi[ORC_MAXQ * 2], fw[ORC_MAXQ];
  if (dim == 2) {
    double gx[2], gw[2];
    gauss_jacobi2(0, gx, gw);
    nq = 2;
    for (int q = 0; q < 2; ++q) { fxi[q] = gx[q]; fw[q] = gw[q]; }
  } else if (orc_quadrature(2, &nq, fxi, fw)) return 1;
  const int uOff[2] = {0, 1}, uOff_x[2] = {0, dim};
  for (int64_t c = 0; c < nC; ++c) { /* serial: the order of the additions into F is then fixed */
    const int *cv = cells + c * nc;
    for (int skip = 0; skip < nc; ++skip) {
      int fv[ORC_MAXNC], fb[ORC_MAXNC], k = 0;
      for (int j = 0; j < nc; ++j) if (j != skip) { fv[k] = cv[j]; fb[k] = j; ++k; }
      int support = 0;
      for (int64_t s = ptr[fv[0]]; s < ptr[fv[0] + 1]; ++s) {
        const int *ov = cells + (int64_t)adj[s] * nc;
        int all = 1;
        for (int a = 1; a < nf && all; ++a) {
          int found = 0;
          for (int j = 0; j < nc; ++j) found |= (ov[j] == fv[a]);
          all = found;
        }
        support += all;
      }
      if (support != 1) continue;
      /* cell gradient (constant for P1) */
      const double *X[ORC_MAXNC];
      double ue[ORC_MAXNC];
      for (int b = 0; b < nc; ++b) { X[b] = coords + (size_t)cv[b] * dim; ue[b] = uloc[cv[b]]; }
      cell_geom g;
      cell_geometry(dim, X, &g);
      double ux[ORC_MAXDIM] = {0, 0, 0};
      for (int b = 0; b < nc; ++b) for (int d = 0; d < dim; ++d) ux[d] += g.G[b][d] * ue[b];
      /* face geometry */
      const double *P0 = coords + (size_t)fv[0] * dim, *P1 = coords + (size_t)fv[1] * dim;
      const double *Po = coords + (size_t)cv[skip] * dim;
      double e1[3] = {0, 0, 0}, e2[3] = {0, 0, 0}, n[3] = {0, 0, 0}, meas;
      for (int d = 0; d < dim; ++d) e1[d] = P1[d] - P0[d];
      if (dim == 2) {
        meas = sqrt(e1[0] * e1[0] + e1[1] * e1[1]);
        n[0] = e1[1] / meas; n[1] = -e1[0] / meas;
      } else {
        const double *P2 = coords + (size_t)fv[2] * dim;
        for (int d = 0; d < 3; ++d) e2[d] = P2[d] - P0[d];
        n[0] = e1[1] * e2[2] - e1[2] * e2[1]; n[1] = e1[2] * e2[0] - e1[0] * e2[2]; n[2] = e1[0] * e2[1] - e1[1] * e2[0];
        const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        meas = 0.5 * nn;
        for (int d = 0; d < 3; ++d) n[d] /= nn;
      }
      double side = 0.0;
      for (int d = 0; d < dim; ++d) side += n[d] * (Po[d] - P0[d]);
      if (side > 0.0) for (int d = 0; d < dim; ++d) n[d] = -n[d];
      const double detJf = 0.5 * meas;
      double Fe[ORC_MAXNC] = {0, 0, 0, 0};
      for (int q = 0; q < nq; ++q) {
        double lam[3], xq[ORC_MAXDIM], uq = 0.0;
        if (dim == 2) { lam[0] = 0.5 * (1.0 - fxi[q]); lam[1] = 0.5 * (1.0 + fxi[q]); }
        else { lam[1] = 0.5 * (1.0 + fxi[q * 2]); lam[2] = 0.5 * (1.0 + fxi[q * 2 + 1]); lam[0] = 1.0 - lam[1] - lam[2]; }
        for (int d = 0; d < dim; ++d) {
          xq[d] = 0.0;
          for (int a = 0; a < nf; ++a) xq[d] += lam[a] * coords[(size_t)fv[a] * dim + d];
        }
        for (int a = 0; a < nf; ++a) uq += lam[a] * uloc[fv[a]];
        double f0[1] = {0.0}, f1[ORC_MAXDIM] = {0, 0, 0};
        f0_bd(dim, 1, 0, uOff, uOff_x, &uq, NULL, ux, NULL, NULL, NULL, NULL, NULL, 0.0, xq, n, 0, NULL, f0);
        f1_bd(dim, 1, 0, uOff, uOff_x, &uq, NULL, ux, NULL, NULL, NULL, NULL, NULL, 0.0, xq, n, 0, NULL, f1);
        const double w = detJf * fw[q];
        for (int a = 0; a < nf; ++a) {
          double s = lam[a] * f0[0];
          for (int d = 0; d < dim; ++d) s += g.G[fb[a]][d] * f1[d];
          Fe[a] += w * s;
        }
      }
      for (int a = 0; a < nf; ++a) {
        const int r = gidx[fv[a]];
        if (r >= 0) F[r] += Fe[a];
      }
    }
  }
  free(uloc); free(ptr); free(adj);
  return 0;
}

/* largest eigenvalue of the symmetric tridiagonal (d[0..n), e[0..n-1)) by bisection on the Sturm count */
static double tridiag_lambda_max(int n, const double *d, const double *e) {
  double lo = d[0], hi = d[0];
  for (int i = 0; i < n; ++i) {
    const double r = (i > 0 ? fabs(e[i - 1]) : 0.0) + (i < n - 1 ? fabs(e[i]) : 0.0);
    if (d[i] - r < lo) lo = d[i] - r;
    if (d[i] + r > hi) hi = d[i] + r;
  }
  for (int it = 0; it < 200 && hi - lo > 1e-14 * fmax(fabs(hi), fabs(lo)); ++it) {
    const double mid = 0.5 * (lo + hi);
    int below = 0; /* number of eigenvalues < mid */
    double q = d[0] - mid;
    if (q < 0.0) ++below;
    for (int i = 1; i < n; ++i) {
      if (q == 0.0) q = 1e-300;
      q = d[i] - mid - e[i - 1] * e[i - 1] / q;
      if (q < 0.0) ++below;
    }
    if (below >= n) hi = mid; else lo = mid;
  }
  return 0.5 * (lo + hi);
}

/* deterministic "noise" in [-1,1) from the global row number (PETSc seeds KSPChebyshev's eigenvalue estimate with a random
 * vector, -ksp_chebyshev_esteig_noisy; a hash keeps runs and partitions reproducible) */
double orc_hash_noise(int64_t i) {
  uint64_t z = (uint64_t)i + 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  z ^= z >> 31;
  return (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

/* lambda_max(D^-1 A) from `steps` iterations of Jacobi-preconditioned CG on the noise vector (the Lanczos tridiagonal of
 * CG: T_jj = 1/alpha_j + beta_{j-1}/alpha_{j-1}, T_{j,j+1} = sqrt(beta_j)/alpha_j) [PETSc-internal
 * KSPChebyshevEstEigSet / KSPComputeExtremeSingularValues; PETSc runs GMRES there, whose Ritz values equal Lanczos' for
 * a symmetric operator] */
double orc_estimate_emax(int N, const int *rowptr, const int *colidx, const double *vals, const double *dinv, int steps) {
  double *r = malloc(sizeof(double) * (size_t)N), *z = malloc(sizeof(double) * (size_t)N);
  double *p = malloc(sizeof(double) * (size_t)N), *w = malloc(sizeof(double) * (size_t)N);
  double td[64], te[64];
  if (steps > 64) steps = 64;
  for (int i = 0; i < N; ++i) { r[i] = orc_hash_noise(i); z[i] = dinv[i] * r[i]; p[i] = z[i]; }
  double beta = vdot(N, z, r), alpha_old = 1.0, bb_old = 0.0;
  int n = 0;
  for (int j = 0; j < steps && beta > 0.0; ++j) {
    orc_spmv(N, rowptr, colidx, vals, p, w);
    const double pw = vdot(N, p, w);
    if (!(pw > 0.0)) break;
    const double alpha = beta / pw;
    td[n] = 1.0 / alpha + (j > 0 ? bb_old / alpha_old : 0.0);
    for (int i = 0; i < N; ++i) { r[i] -= alpha * w[i]; z[i] = dinv[i] * r[i]; }
    const double beta_new = vdot(N, z, r);
    const double bb = beta_new / beta;
    te[n] = sqrt(bb) / alpha;
    ++n;
    for (int i = 0; i < N; ++i) p[i] = z[i] + bb * p[i];
    beta = beta_new; alpha_old = alpha; bb_old = bb;
  }
  const double lam = n ? tridiag_lambda_max(n, td, te) : 1.0;
  free(r); free(z); free(p); free(w);
  return lam;
}

/* z = p_k(D^-1 A) D^-1 r: `its` iterations of the Chebyshev semi-iteration with Jacobi inside, zero initial guess
 * [PETSc-internal KSPSolve_Chebyshev as -pc_type ksp -ksp_ksp_type chebyshev -ksp_pc_type jacobi -ksp_ksp_max_it its
 * runs it: three rotating iterates, scale = 2/(emax+emin), alpha = 1 - scale*emin, mu = 1/alpha, c_{k+1} = 2 mu c_k -
 * c_{k-1}, omega = (2/alpha) c_k / c_{k+1},  y_{k+1} = omega (y_k - y_{k-1} + scale B^-1 (r - A y_k)) + y_{k-1}] */
static void cheb_apply(int N, const int *rowptr, const int *colidx, const double *vals, const double *dinv, int its,
                       double emin, double emax, const double *r, double *z, double *y0, double *y1, double *t) {
  const double scale = 2.0 / (emax + emin), alpha = 1.0 - scale * emin, mu = 1.0 / alpha, omegaprod = 2.0 / alpha;
  double c_km1 = 1.0, c_k = mu;
  double *ykm1 = y0, *yk = y1, *ykp1 = z;
  for (int i = 0; i < N; ++i) { ykm1[i] = 0.0; yk[i] = scale * dinv[i] * r[i]; }
  for (int k = 1; k < its; ++k) {
    orc_spmv(N, rowptr, colidx, vals, yk, t);
    const double c_kp1 = 2.0 * mu * c_k - c_km1, omega = omegaprod * c_k / c_kp1;
    for (int i = 0; i < N; ++i)
      ykp1[i] = (1.0 - omega) * ykm1[i] + omega * yk[i] + scale * omega * (dinv[i] * (r[i] - t[i]));
    double *tmp = ykm1; ykm1 = yk; yk = ykp1; ykp1 = tmp;
    c_km1 = c_k; c_k = c_kp1;
  }
  if (yk != z) memcpy(z, yk, sizeof(double) * (size_t)N);
}

/* Preconditioned CG with the options this row adds to orc_cg: pc 2 = the Chebyshev/Jacobi polynomial above
 * (emax <= 0: 1.1 x the estimate of 10 Lanczos steps, PETSc's safety factor; emin <= 0: emax / (4 k^2 + 36), an
 * interval that widens with the degree because the polynomial preconditions CG here instead of smoothing for multigrid --
 * PETSc's smoother default emin = 0.1 x estimate is reached by passing both ends), and
 * `nullspace` = 1: the constants are projected out of every preconditioned residual [PETSc-internal KSP_PCApply ->
 * KSP_RemoveNullSpace, with the MatNullSpaceCreate(comm, PETSC_TRUE, 0, NULL) that ex12.cpp:1534-1538 attaches when
 * -bc_type is not dirichlet].  Iteration and convergence test as orc_cg.  eig[0..1] returns the interval used. */
int orc_pcg(int N, const int *rowptr, const int *colidx, const double *vals, const double *b, double *x, double rtol,
            double abstol, double dtol, int maxit, int pc, int cheb_its, double emin, double emax, int nullspace,
            int *its_out, int *reason_out, double *rnorm_out, double *hist, double *eig) {
  const size_t nb = sizeof(double) * (size_t)(N > 0 ? N : 1);
  double *r = malloc(nb), *z = malloc(nb), *p = malloc(nb), *w = malloc(nb), *dinv = malloc(nb);
  double *y0 = malloc(nb), *y1 = malloc(nb), *t = malloc(nb);
  for (int i = 0; i < N; ++i) {
    const int64_t s = find_slot(rowptr, colidx, i, i);
    const double d = (s >= 0) ? vals[s] : 0.0;
    dinv[i] = (pc != 0 && d != 0.0) ? 1.0 / d : 1.0;
  }
  if (pc == 2) {
    if (!(emax > 0.0)) emax = 1.1 * orc_estimate_emax(N, rowptr, colidx, vals, dinv, 10);
    if (!(emin > 0.0)) emin = emax / (4.0 * cheb_its * cheb_its + 36.0);
  }
  if (eig) { eig[0] = emin; eig[1] = emax; }
#define ORC_APPLY_PC()                                                                              \
  do {                                                                                              \
    if (pc == 2) cheb_apply(N, rowptr, colidx, vals, dinv, cheb_its, emin, emax, r, z, y0, y1, t);  \
    else for (int k = 0; k < N; ++k) z[k] = dinv[k] * r[k];                                         \
    if (nullspace && N > 0) {                                                                       \
      double s = 0.0;                                                                               \
      for (int k = 0; k < N; ++k) s += z[k];                                                        \
      s /= N;                                                                                       \
      for (int k = 0; k < N; ++k) z[k] -= s;                                                        \
    }                                                                                               \
  } while (0)
  int reason = 0, i = 0;
  for (int k = 0; k < N; ++k) { x[k] = 0.0; r[k] = b[k]; }
  ORC_APPLY_PC();
  double dp = sqrt(vdot(N, z, z));
  double beta = vdot(N, z, r), betaold = 1.0;
  const double rnorm0 = dp, ttol = fmax(rtol * rnorm0, abstol);
  if (hist) hist[0] = dp;
  if (dp != dp) reason = -9;
  else if (dp <= ttol) reason = (dp < abstol) ? 3 : 2;
  while (!reason && i < maxit) {
    if (beta == 0.0) { reason = 3; break; }
    if (i == 0) memcpy(p, z, nb);
    else {
      const double bb = beta / betaold;
      for (int k = 0; k < N; ++k) p[k] = z[k] + bb * p[k];
    }
    orc_spmv(N, rowptr, colidx, vals, p, w);
    const double dpi = vdot(N, p, w);
    betaold = beta;
    if (!(dpi > 0.0)) { reason = (dpi != dpi) ? -9 : -8; ++i; break; }
    const double a = beta / dpi;
    for (int k = 0; k < N; ++k) { x[k] += a * p[k]; r[k] -= a * w[k]; }
    ORC_APPLY_PC();
    dp = sqrt(vdot(N, z, z));
    ++i;
    if (hist) hist[i] = dp;
    if (dp != dp) reason = -9;
    else if (dp <= ttol) reason = (dp < abstol) ? 3 : 2;
    else if (dp >= dtol * rnorm0) reason = -4;
    if (reason) break;
    beta = vdot(N, z, r);
  }
#undef ORC_APPLY_PC
  if (!reason) reason = -3;
  *its_out = i; *reason_out = reason; *rnorm_out = dp;
  free(r); free(z); free(p); free(w); free(dinv); free(y0); free(y1); free(t);
  return 0;
}
