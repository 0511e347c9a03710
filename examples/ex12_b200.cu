/*
 * ex12_b200.cu -- the reference's main() (ex12.cpp:1472-1705) for the P1 Poisson hot path, written
 * against include/ohp_petsc.h + include/ohp_omega_h.h.  What differs from ex12.cpp, and why, is
 * listed in INTEGRATION.md; everything else (option names, call order, callback bodies, the
 * CHKERRQ style) is kept so the two files can be read side by side.
 *
 *   ./ex12_b200 -mesh <dir.osh | box> [-cells nx,ny] [-run_type full|test|perf] [-variable_coefficient none|analytic|nonlinear]
 *               [-ksp_rtol 1e-9] [-ksp_max_it N] [-pc_type none|jacobi] [-snes_monitor_short] [-snes_converged_reason]
 */
#include <assert.h>

#include "ohp_omega_h.h"
#include "ohp_petsc.h"

static char help[] = "Poisson Problem in 2d and 3d with simplicial finite elements (B200 path).\n";

typedef enum { NEUMANN, DIRICHLET, NONE } BCType;
typedef enum { RUN_FULL, RUN_EXACT, RUN_TEST, RUN_PERF } RunType;
typedef enum { COEFF_NONE, COEFF_ANALYTIC, COEFF_FIELD, COEFF_NONLINEAR } CoeffType;

typedef struct { /* the fields of AppCtx the hot path reads (ex12.cpp:43-74) */
  PetscInt debug;
  RunType runType;
  PetscInt dim;
  PetscInt cells[3];
  BCType bcType;
  CoeffType variableCoefficient;
  PetscSimplePointFunc *exactFuncs;
  char mesh_type[1024];
  char picpart_path[1024];
} AppCtx;

/* ---- pointwise callbacks: bodies as in ex12.cpp, qualified PETSC_HOSTDEVICE ---- */
PETSC_HOSTDEVICE static PetscErrorCode zero(PetscInt dim, PetscReal time, const PetscReal x[], PetscInt Nc, PetscScalar *u, void *ctx)
{ /* ex12.cpp:76-80 */
  u[0] = 0.0;
  return 0;
}
PETSC_HOSTDEVICE static PetscErrorCode quadratic_u_2d(PetscInt dim, PetscReal time, const PetscReal x[], PetscInt Nc, PetscScalar *u, void *ctx)
{ /* ex12.cpp:113-117 */
  *u = x[0]*x[0] + x[1]*x[1];
  return 0;
}
PETSC_HOSTDEVICE static void f0_u(PetscInt dim, PetscInt Nf, PetscInt NfAux,
                 const PetscInt uOff[], const PetscInt uOff_x[], const PetscScalar u[], const PetscScalar u_t[], const PetscScalar u_x[],
                 const PetscInt aOff[], const PetscInt aOff_x[], const PetscScalar a[], const PetscScalar a_t[], const PetscScalar a_x[],
                 PetscReal t, const PetscReal x[], PetscInt numConstants, const PetscScalar constants[], PetscScalar f0[])
{ /* ex12.cpp:147-153 */
  f0[0] = 4.0;
}
PETSC_HOSTDEVICE static void f1_u(PetscInt dim, PetscInt Nf, PetscInt NfAux,
                 const PetscInt uOff[], const PetscInt uOff_x[], const PetscScalar u[], const PetscScalar u_t[], const PetscScalar u_x[],
                 const PetscInt aOff[], const PetscInt aOff_x[], const PetscScalar a[], const PetscScalar a_t[], const PetscScalar a_x[],
                 PetscReal t, const PetscReal x[], PetscInt numConstants, const PetscScalar constants[], PetscScalar f1[])
{ /* ex12.cpp:198-205 */
  PetscInt d;
  for (d = 0; d < dim; ++d) f1[d] = u_x[d];
}
PETSC_HOSTDEVICE static void g3_uu(PetscInt dim, PetscInt Nf, PetscInt NfAux,
                  const PetscInt uOff[], const PetscInt uOff_x[], const PetscScalar u[], const PetscScalar u_t[], const PetscScalar u_x[],
                  const PetscInt aOff[], const PetscInt aOff_x[], const PetscScalar a[], const PetscScalar a_t[], const PetscScalar a_x[],
                  PetscReal t, PetscReal u_tShift, const PetscReal x[], PetscInt numConstants, const PetscScalar constants[], PetscScalar g3[])
{ /* ex12.cpp:209-216 */
  PetscInt d;
  for (d = 0; d < dim; ++d) g3[d*dim+d] = 1.0;
}
PETSC_HOSTDEVICE static void f0_analytic_u(PetscInt dim, PetscInt Nf, PetscInt NfAux,
                   const PetscInt uOff[], const PetscInt uOff_x[], const PetscScalar u[], const PetscScalar u_t[], const PetscScalar u_x[],
                   const PetscInt aOff[], const PetscInt aOff_x[], const PetscScalar a[], const PetscScalar a_t[], const PetscScalar a_x[],
                   PetscReal t, const PetscReal x[], PetscInt numConstants, const PetscScalar constants[], PetscScalar f0[])
{ /* ex12.cpp:283-289 */
  f0[0] = 6.0*(x[0] + x[1]);
}
PETSC_HOSTDEVICE static void f1_analytic_u(PetscInt dim, PetscInt Nf, PetscInt NfAux,
                   const PetscInt uOff[], const PetscInt uOff_x[], const PetscScalar u[], const PetscScalar u_t[], const PetscScalar u_x[],
                   const PetscInt aOff[], const PetscInt aOff_x[], const PetscScalar a[], const PetscScalar a_t[], const PetscScalar a_x[],
                   PetscReal t, const PetscReal x[], PetscInt numConstants, const PetscScalar constants[], PetscScalar f1[])
{ /* ex12.cpp:292-299 */
  PetscInt d;
  for (d = 0; d < dim; ++d) f1[d] = (x[0] + x[1])*u_x[d];
}
PETSC_HOSTDEVICE static void g3_analytic_uu(PetscInt dim, PetscInt Nf, PetscInt NfAux,
                    const PetscInt uOff[], const PetscInt uOff_x[], const PetscScalar u[], const PetscScalar u_t[], const PetscScalar u_x[],
                    const PetscInt aOff[], const PetscInt aOff_x[], const PetscScalar a[], const PetscScalar a_t[], const PetscScalar a_x[],
                    PetscReal t, PetscReal u_tShift, const PetscReal x[], PetscInt numConstants, const PetscScalar constants[], PetscScalar g3[])
{ /* ex12.cpp:313-320 */
  PetscInt d;
  for (d = 0; d < dim; ++d) g3[d*dim+d] = x[0] + x[1];
}

/* device instantiations of the callback sets SetupProblem can select (one line each) */
OHP_DEFINE_PROBLEM(coeff_none, f0_u, f1_u, g3_uu, quadratic_u_2d, quadratic_u_2d)
OHP_DEFINE_PROBLEM(coeff_analytic, f0_analytic_u, f1_analytic_u, g3_analytic_uu, quadratic_u_2d, quadratic_u_2d)

static PetscErrorCode ProcessOptions(MPI_Comm comm, AppCtx *options)
{ /* the options of ex12.cpp:430-513 that the hot path reads */
  char buf[64];
  PetscInt n = 3;
  PetscBool set;
  PetscErrorCode ierr;

  PetscFunctionBeginUser;
  options->debug = 0; options->runType = RUN_FULL; options->dim = 2;
  options->cells[0] = options->cells[1] = options->cells[2] = 2;
  options->bcType = DIRICHLET; options->variableCoefficient = COEFF_NONE;
  strcpy(options->mesh_type, "box"); options->picpart_path[0] = 0;
  ierr = PetscOptionsGetInt(NULL, NULL, "-debug", &options->debug, NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetString(NULL, NULL, "-run_type", buf, sizeof(buf), &set);CHKERRQ(ierr);
  if (set) options->runType = !strcmp(buf, "full") ? RUN_FULL : !strcmp(buf, "test") ? RUN_TEST : !strcmp(buf, "perf") ? RUN_PERF : RUN_EXACT;
  ierr = PetscOptionsGetInt(NULL, NULL, "-dim", &options->dim, NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetIntArray(NULL, NULL, "-cells", options->cells, &n, NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetString(NULL, NULL, "-variable_coefficient", buf, sizeof(buf), &set);CHKERRQ(ierr);
  if (set) options->variableCoefficient = !strcmp(buf, "analytic") ? COEFF_ANALYTIC : COEFF_NONE;
  ierr = PetscOptionsGetString(NULL, NULL, "-mesh", options->mesh_type, sizeof(options->mesh_type), NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetString(NULL, NULL, "-picpart_path", options->picpart_path, sizeof(options->picpart_path), NULL);CHKERRQ(ierr);
  PetscFunctionReturn(0);
}

static std::vector<int> getPtnMeshElmToVtxArray(Omega_h::Mesh &mesh)
{ /* ex12.cpp:781-791 */
  const auto dim = mesh.dim();
  const auto cells2verts = mesh.ask_elem_verts();
  const auto cells2verts_h = Omega_h::HostRead<Omega_h::LO>(cells2verts);
  assert(cells2verts.size() == (mesh.nelems() * (dim + 1)));
  std::vector<int> c2v;
  c2v.reserve(cells2verts_h.size());
  for (int i = 0; i < cells2verts_h.size(); i++) c2v.push_back(cells2verts_h[i]);
  return c2v;
}

static PetscErrorCode CreateQuadMesh(MPI_Comm comm, DM *dm, AppCtx *options, Omega_h::Library &lib)
{ /* serial branch of ex12.cpp:793-1036 */
  PetscErrorCode ierr;
  std::vector<int> cells2verts;
  std::vector<double> vertexCoords;
  int numCells, numVertices, dim = 2;

  PetscFunctionBeginUser;
  if (strcmp(options->mesh_type, "box") == 0) {
    /* Omega_h::build_box (ex12.cpp:805) is replaced by the same-diagonal box of SURVEY 8d */
    const int nx = options->cells[0], ny = options->cells[1];
    numCells = 2 * nx * ny; numVertices = (nx + 1) * (ny + 1);
    vertexCoords.resize(2 * (size_t)numVertices);
    cells2verts.resize(3 * (size_t)numCells);
    for (int j = 0; j <= ny; ++j) for (int i = 0; i <= nx; ++i) {
      vertexCoords[2 * (size_t)(j * (nx + 1) + i)] = i * (1.0 / nx);
      vertexCoords[2 * (size_t)(j * (nx + 1) + i) + 1] = j * (1.0 / ny);
    }
    for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      const int v00 = j * (nx + 1) + i, v10 = v00 + 1, v01 = v00 + nx + 1, v11 = v01 + 1;
      int *c = &cells2verts[6 * (size_t)(j * nx + i)];
      c[0] = v00; c[1] = v10; c[2] = v11; c[3] = v00; c[4] = v11; c[5] = v01;
    }
  } else {
    Omega_h::Mesh mesh(&lib);
    Omega_h::binary::read(options->mesh_type, lib.world(), &mesh, false); /* ex12.cpp:818 */
    dim = mesh.dim();
    cells2verts = getPtnMeshElmToVtxArray(mesh);                          /* ex12.cpp:887 */
    const auto coords = Omega_h::HostRead<Omega_h::Real>(mesh.coords());
    vertexCoords.assign(coords.data(), coords.data() + coords.size());
    numCells = mesh.nelems(); numVertices = mesh.nverts();
  }
  options->dim = dim;
  ierr = DMPlexCreateFromCellListPetsc(comm, dim, numCells, numVertices, dim + 1, PETSC_FALSE,
                                       cells2verts.data(), dim, vertexCoords.data(), dm);CHKERRQ(ierr); /* ex12.cpp:933 */
  {
    DM idm;
    ierr = DMPlexInterpolate(*dm, &idm);CHKERRQ(ierr);                    /* ex12.cpp:958 */
    ierr = DMDestroy(dm);CHKERRQ(ierr);
    *dm = idm;
  }
  { /* boundary: CreateBCLabel's two calls (ex12.cpp:515-526) instead of the host loops of ex12.cpp:986-1033 */
    DMLabel label;
    ierr = DMCreateLabel(*dm, "marker");CHKERRQ(ierr);
    ierr = DMGetLabel(*dm, "marker", &label);CHKERRQ(ierr);
    ierr = DMPlexMarkBoundaryFaces(*dm, 1, label);CHKERRQ(ierr);
    ierr = DMPlexLabelComplete(*dm, label);CHKERRQ(ierr);
  }
  PetscFunctionReturn(0);
}

static PetscErrorCode SetupProblem(DM dm, AppCtx *user)
{ /* ex12.cpp:1162-1243, COEFF_NONE / COEFF_ANALYTIC, Dirichlet */
  PetscDS prob;
  const PetscInt id = 1;
  PetscErrorCode ierr;

  PetscFunctionBeginUser;
  ierr = DMGetDS(dm, &prob);CHKERRQ(ierr);
  switch (user->variableCoefficient) {
  case COEFF_NONE:
    ierr = PetscDSSetResidual(prob, 0, f0_u, f1_u);CHKERRQ(ierr);
    ierr = PetscDSSetJacobian(prob, 0, 0, NULL, NULL, NULL, g3_uu);CHKERRQ(ierr);
    break;
  case COEFF_ANALYTIC:
    ierr = PetscDSSetResidual(prob, 0, f0_analytic_u, f1_analytic_u);CHKERRQ(ierr);
    ierr = PetscDSSetJacobian(prob, 0, 0, NULL, NULL, NULL, g3_analytic_uu);CHKERRQ(ierr);
    break;
  default: SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, "Invalid variable coefficient type %d", user->variableCoefficient);
  }
  user->exactFuncs[0] = quadratic_u_2d;
  ierr = DMAddBoundary(dm, DM_BC_ESSENTIAL, "wall", "marker", 0, 0, NULL, (void (*)(void)) user->exactFuncs[0], NULL, 1, &id, user);CHKERRQ(ierr);
  ierr = PetscDSSetExactSolution(prob, 0, user->exactFuncs[0], user);CHKERRQ(ierr);
  PetscFunctionReturn(0);
}

static PetscErrorCode SetupDiscretization(DM dm, AppCtx *user)
{ /* ex12.cpp:1297-1336 */
  PetscFE fe;
  PetscErrorCode ierr;

  PetscFunctionBeginUser;
  ierr = PetscFECreateDefault(PETSC_COMM_WORLD, user->dim, 1, PETSC_TRUE, NULL, -1, &fe);CHKERRQ(ierr);
  ierr = PetscObjectSetName((PetscObject) fe, "potential");CHKERRQ(ierr);
  ierr = DMSetField(dm, 0, NULL, (PetscObject) fe);CHKERRQ(ierr);
  ierr = DMCreateDS(dm);CHKERRQ(ierr);
  ierr = SetupProblem(dm, user);CHKERRQ(ierr);
  ierr = PetscFEDestroy(&fe);CHKERRQ(ierr);
  PetscFunctionReturn(0);
}

int main(int argc, char **argv)
{
  DM             dm;
  SNES           snes;
  Vec            u;
  Mat            A, J;
  AppCtx         user;
  PetscReal      error = 0.0;
  PetscSimplePointFunc exact[1], initialGuess[1] = {zero};
  PetscErrorCode ierr;

  ierr = PetscInitialize(&argc, &argv, NULL, help);if (ierr) { fprintf(stderr, "%s\n", ohp_last_error()); return ierr; }
  ierr = ProcessOptions(PETSC_COMM_WORLD, &user);CHKERRQ(ierr);
  ierr = SNESCreate(PETSC_COMM_WORLD, &snes);CHKERRQ(ierr);
  {
    auto ohlib = Omega_h::Library(&argc, &argv);
    ierr = CreateQuadMesh(PETSC_COMM_WORLD, &dm, &user, ohlib);CHKERRQ(ierr);
  }
  ierr = SNESSetDM(snes, dm);CHKERRQ(ierr);
  ierr = DMSetApplicationContext(dm, &user);CHKERRQ(ierr);
  user.exactFuncs = exact;
  ierr = SetupDiscretization(dm, &user);CHKERRQ(ierr);

  ierr = DMCreateGlobalVector(dm, &u);CHKERRQ(ierr);
  ierr = PetscObjectSetName((PetscObject) u, "potential");CHKERRQ(ierr);
  ierr = DMCreateMatrix(dm, &J);CHKERRQ(ierr);
  A = J;
  ierr = DMPlexSetSNESLocalFEM(dm, &user, &user, &user);CHKERRQ(ierr);
  ierr = SNESSetJacobian(snes, A, J, NULL, NULL);CHKERRQ(ierr);
  ierr = SNESSetFromOptions(snes);CHKERRQ(ierr);

  ierr = DMProjectFunction(dm, 0.0, user.exactFuncs, NULL, INSERT_ALL_VALUES, u);CHKERRQ(ierr);
  if (user.runType == RUN_FULL) {
    ierr = DMProjectFunction(dm, 0.0, initialGuess, NULL, INSERT_VALUES, u);CHKERRQ(ierr);  /* ex12.cpp:1598-1602 */
    ierr = SNESSolve(snes, NULL, u);CHKERRQ(ierr);                                            /* ex12.cpp:1609 */
    ierr = SNESGetSolution(snes, &u);CHKERRQ(ierr);
    ierr = SNESGetDM(snes, &dm);CHKERRQ(ierr);
    ierr = DMComputeL2Diff(dm, 0.0, user.exactFuncs, NULL, u, &error);CHKERRQ(ierr);
    PetscInt its, lits, N;
    ierr = SNESGetIterationNumber(snes, &its);CHKERRQ(ierr);
    ierr = SNESGetLinearSolveIterations(snes, &lits);CHKERRQ(ierr);
    ierr = VecGetSize(u, &N);CHKERRQ(ierr);
    ierr = PetscPrintf(PETSC_COMM_WORLD, "N: %d Newton its: %d CG its: %d L_2 Error: %.6g\n", N, its, lits, (double)error);CHKERRQ(ierr);
    ierr = VecViewFromOptions(u, NULL, "-vec_view");CHKERRQ(ierr);                             /* ex12.cpp:1618 */
  } else if (user.runType == RUN_PERF) {                                                       /* ex12.cpp:1619-1628 */
    Vec r;
    PetscReal res = 0.0;
    ierr = DMProjectFunction(dm, 0.0, initialGuess, NULL, INSERT_VALUES, u);CHKERRQ(ierr);
    ierr = VecDuplicate(u, &r);CHKERRQ(ierr);
    ierr = SNESComputeFunction(snes, u, r);CHKERRQ(ierr);
    ierr = VecChop(r, 1.0e-10);CHKERRQ(ierr);
    ierr = VecNorm(r, NORM_2, &res);CHKERRQ(ierr);
    ierr = PetscPrintf(PETSC_COMM_WORLD, "Initial Residual\n L_2 Residual: %g\n", (double)res);CHKERRQ(ierr);
    ierr = VecDestroy(&r);CHKERRQ(ierr);
  } else {                                                                                     /* RUN_TEST, ex12.cpp:1629-1680 */
    Vec r, b;
    PetscReal res = 0.0;
    ierr = DMComputeL2Diff(dm, 0.0, user.exactFuncs, NULL, u, &error);CHKERRQ(ierr);
    ierr = PetscPrintf(PETSC_COMM_WORLD, "L_2 Error: %g\n", (double)error);CHKERRQ(ierr);
    ierr = VecDuplicate(u, &r);CHKERRQ(ierr);
    ierr = SNESComputeFunction(snes, u, r);CHKERRQ(ierr);
    ierr = VecChop(r, 1.0e-10);CHKERRQ(ierr);
    ierr = VecNorm(r, NORM_2, &res);CHKERRQ(ierr);
    ierr = PetscPrintf(PETSC_COMM_WORLD, "L_2 Residual: %g\n", (double)res);CHKERRQ(ierr);
    /* Jacobian consistency  A u + F(0) = F(u)  (ex12.cpp:1647-1661) */
    ierr = SNESComputeJacobian(snes, u, A, A);CHKERRQ(ierr);
    ierr = VecDuplicate(u, &b);CHKERRQ(ierr);
    ierr = VecSet(r, 0.0);CHKERRQ(ierr);
    ierr = SNESComputeFunction(snes, r, b);CHKERRQ(ierr);
    ierr = MatMult(A, u, r);CHKERRQ(ierr);
    ierr = VecAXPY(r, 1.0, b);CHKERRQ(ierr);
    ierr = VecChop(r, 1.0e-10);CHKERRQ(ierr);
    ierr = VecNorm(r, NORM_2, &res);CHKERRQ(ierr);
    ierr = PetscPrintf(PETSC_COMM_WORLD, "Linear L_2 Residual: %g\n", (double)res);CHKERRQ(ierr);
    ierr = VecDestroy(&b);CHKERRQ(ierr);
    ierr = VecDestroy(&r);CHKERRQ(ierr);
  }
  if (A != J) {ierr = MatDestroy(&A);CHKERRQ(ierr);}
  ierr = MatDestroy(&J);CHKERRQ(ierr);
  ierr = VecDestroy(&u);CHKERRQ(ierr);
  ierr = SNESDestroy(&snes);CHKERRQ(ierr);
  ierr = DMDestroy(&dm);CHKERRQ(ierr);
  ierr = PetscFinalize();
  return ierr;
}
