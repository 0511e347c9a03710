/*
 * ex12_b200.cu -- an ex12-style driver for the P1 Poisson hot path on top of include/ohp_petsc.h and
 * include/ohp_omega_h.h.  It shows that the reference's PETSc/Omega_h call surface (ex12.cpp:1472-1705)
 * carries over: the same option names, the same DM/DS/SNES calls in the same order, pointwise callbacks
 * with the PetscDS ABI -- now compiled for the device and bound with OHP_DEFINE_PROBLEM.
 * INTEGRATION.md lists what a maintainer edits in ex12.cpp itself.
 *
 *   ./ex12_b200 -mesh <dir.osh | box> [-cells nx,ny] [-run_type full|test|perf]
 *               [-variable_coefficient none|analytic] [-ksp_rtol 1e-9] [-ksp_max_it N] [-pc_type none|jacobi]
 *               [-snes_monitor_short] [-snes_converged_reason] [-vec_view vtk:out.vtu]
 */
#include <assert.h>

#include "ohp_omega_h.h"
#include "ohp_petsc.h"

/* ---------------------------------------------------------------------------------------------
 * Pointwise callbacks.  OHP_PF_ARGS / OHP_PJ_ARGS expand to the 17 / 18 leading parameters of a
 * PetscPointFunc / PetscPointJac (dim, Nf, NfAux, uOff, uOff_x, u, u_t, u_x, aOff, aOff_x, a, a_t, a_x, t,
 * [u_tShift,] x, numConstants, constants); the last parameter is the output array.
 * --------------------------------------------------------------------------------------------- */
namespace poisson {

/* u = 0: initial guess (the reference's zero(), ex12.cpp:76) */
PETSC_HOSTDEVICE static PetscErrorCode guess(PetscInt, PetscReal, const PetscReal[], PetscInt, PetscScalar *u, void *) { u[0] = 0.0; return 0; }

/* manufactured solution and Dirichlet data u = x^2 + y^2 (ex12.cpp:113) */
PETSC_HOSTDEVICE static PetscErrorCode exact2d(PetscInt, PetscReal, const PetscReal x[], PetscInt, PetscScalar *u, void *) {
  u[0] = x[0] * x[0] + x[1] * x[1];
  return 0;
}

/* constant coefficient: -lap u + 4 = 0  ->  f0 = 4, f1 = grad u, g3 = I (ex12.cpp:147,198,209) */
PETSC_HOSTDEVICE static void source(OHP_PF_ARGS, PetscScalar f0[]) { f0[0] = 4.0; }
PETSC_HOSTDEVICE static void flux(OHP_PF_ARGS, PetscScalar f1[]) {
  for (PetscInt d = 0; d < dim; ++d) f1[d] = u_x[d];
}
PETSC_HOSTDEVICE static void dflux(OHP_PJ_ARGS, PetscScalar g3[]) {
  for (PetscInt d = 0; d < dim; ++d) g3[d * dim + d] = 1.0;
}

/* nu = x + y: -div(nu grad u) + 6 (x + y) = 0 (ex12.cpp:283,292,313) */
PETSC_HOSTDEVICE static void source_nu(OHP_PF_ARGS, PetscScalar f0[]) { f0[0] = 6.0 * (x[0] + x[1]); }
PETSC_HOSTDEVICE static void flux_nu(OHP_PF_ARGS, PetscScalar f1[]) {
  const PetscReal nu = x[0] + x[1];
  for (PetscInt d = 0; d < dim; ++d) f1[d] = nu * u_x[d];
}
PETSC_HOSTDEVICE static void dflux_nu(OHP_PJ_ARGS, PetscScalar g3[]) {
  const PetscReal nu = x[0] + x[1];
  for (PetscInt d = 0; d < dim; ++d) g3[d * dim + d] = nu;
}

}  // namespace poisson

/* one line per callback set: instantiates the assembly kernels with the callbacks inlined */
OHP_DEFINE_PROBLEM(constant, poisson::source, poisson::flux, poisson::dflux, poisson::exact2d, poisson::exact2d)
OHP_DEFINE_PROBLEM(variable, poisson::source_nu, poisson::flux_nu, poisson::dflux_nu, poisson::exact2d, poisson::exact2d)

/* ---------------------------------------------------------------------------------------------
 * Run configuration: the AppCtx fields the hot path reads (ex12.cpp:43-74, options at :430-513)
 * --------------------------------------------------------------------------------------------- */
struct Config {
  enum Run { FULL, TEST, PERF } run = FULL;
  bool variable_coefficient = false;
  PetscInt dim = 2;
  PetscInt cells[3] = {2, 2, 2};
  char mesh[1024] = "box";
  PetscSimplePointFunc exact[1] = {poisson::exact2d};
};

static PetscErrorCode read_options(Config *cfg) {
  char word[64];
  PetscBool given;
  PetscInt n = 3;
  PetscErrorCode ierr;
  ierr = PetscOptionsGetString(NULL, NULL, "-run_type", word, sizeof(word), &given);CHKERRQ(ierr);
  if (given) cfg->run = !strcmp(word, "test") ? Config::TEST : !strcmp(word, "perf") ? Config::PERF : Config::FULL;
  ierr = PetscOptionsGetString(NULL, NULL, "-variable_coefficient", word, sizeof(word), &given);CHKERRQ(ierr);
  if (given) {
    if (!strcmp(word, "analytic")) cfg->variable_coefficient = true;
    else if (strcmp(word, "none")) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, "-variable_coefficient %s is not built", word);
  }
  ierr = PetscOptionsGetIntArray(NULL, NULL, "-cells", cfg->cells, &n, NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetString(NULL, NULL, "-mesh", cfg->mesh, sizeof(cfg->mesh), NULL);CHKERRQ(ierr);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Mesh -> DM: the serial branch of CreateQuadMesh (ex12.cpp:793-1036)
 * --------------------------------------------------------------------------------------------- */
static void unit_square(int nx, int ny, std::vector<int> &cells, std::vector<double> &coords) {
  /* stands in for Omega_h::build_box (ex12.cpp:805): quads cut on the same diagonal, vertices row-major */
  const int stride = nx + 1;
  coords.resize(2 * (size_t)stride * (ny + 1));
  for (int j = 0; j <= ny; ++j)
    for (int i = 0; i <= nx; ++i) {
      coords[2 * ((size_t)j * stride + i)] = i * (1.0 / nx);
      coords[2 * ((size_t)j * stride + i) + 1] = j * (1.0 / ny);
    }
  cells.clear();
  cells.reserve(6 * (size_t)nx * ny);
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      const int sw = j * stride + i, se = sw + 1, nw = sw + stride, ne = nw + 1;
      const int two[6] = {sw, se, ne, sw, ne, nw};
      cells.insert(cells.end(), two, two + 6);
    }
}

static PetscErrorCode create_dm(Config *cfg, Omega_h::Library &lib, DM *dm) {
  std::vector<int> cells;
  std::vector<double> coords;
  PetscErrorCode ierr;
  if (!strcmp(cfg->mesh, "box")) {
    unit_square(cfg->cells[0], cfg->cells[1], cells, coords);
  } else {
    Omega_h::Mesh mesh(&lib);
    try {
      Omega_h::binary::read(cfg->mesh, lib.world(), &mesh, false);     /* ex12.cpp:818 */
    } catch (const std::exception &e) {
      SETERRQ1(PETSC_COMM_SELF, OHP_ERR_FILE_OPEN, "%s", e.what());
    }
    cfg->dim = mesh.dim();
    const Omega_h::HostRead<Omega_h::LO> ev(mesh.ask_elem_verts());    /* ex12.cpp:784-790 */
    const Omega_h::HostRead<Omega_h::Real> xyz(mesh.coords());
    assert(ev.size() == mesh.nelems() * (cfg->dim + 1));
    cells.assign(ev.data(), ev.data() + ev.size());
    coords.assign(xyz.data(), xyz.data() + xyz.size());
  }
  const PetscInt nc = cfg->dim + 1;
  ierr = DMPlexCreateFromCellListPetsc(PETSC_COMM_WORLD, cfg->dim, (PetscInt)(cells.size() / nc), (PetscInt)(coords.size() / cfg->dim),
                                       nc, PETSC_FALSE, cells.data(), cfg->dim, coords.data(), dm);CHKERRQ(ierr);   /* ex12.cpp:933 */
  DM interpolated;
  ierr = DMPlexInterpolate(*dm, &interpolated);CHKERRQ(ierr);                                                        /* ex12.cpp:958 */
  ierr = DMDestroy(dm);CHKERRQ(ierr);
  *dm = interpolated;
  /* "marker" label on the boundary: the two calls of CreateBCLabel (ex12.cpp:515-526) instead of the
   * host loops over edge supports (ex12.cpp:986-1033); evaluated on the device */
  DMLabel marker;
  ierr = DMCreateLabel(*dm, "marker");CHKERRQ(ierr);
  ierr = DMGetLabel(*dm, "marker", &marker);CHKERRQ(ierr);
  ierr = DMPlexMarkBoundaryFaces(*dm, 1, marker);CHKERRQ(ierr);
  ierr = DMPlexLabelComplete(*dm, marker);CHKERRQ(ierr);
  return 0;
}

/* P1 element + callbacks + Dirichlet condition: SetupDiscretization / SetupProblem (ex12.cpp:1297-1336, 1162-1243) */
static PetscErrorCode setup_problem(DM dm, Config *cfg) {
  const PetscInt wall = 1;
  PetscFE fe;
  PetscDS ds;
  PetscErrorCode ierr;
  ierr = PetscFECreateDefault(PETSC_COMM_WORLD, cfg->dim, 1, PETSC_TRUE, NULL, -1, &fe);CHKERRQ(ierr);
  ierr = DMSetField(dm, 0, NULL, (PetscObject)fe);CHKERRQ(ierr);
  ierr = DMCreateDS(dm);CHKERRQ(ierr);
  ierr = DMGetDS(dm, &ds);CHKERRQ(ierr);
  if (cfg->variable_coefficient) {
    ierr = PetscDSSetResidual(ds, 0, poisson::source_nu, poisson::flux_nu);CHKERRQ(ierr);
    ierr = PetscDSSetJacobian(ds, 0, 0, NULL, NULL, NULL, poisson::dflux_nu);CHKERRQ(ierr);
  } else {
    ierr = PetscDSSetResidual(ds, 0, poisson::source, poisson::flux);CHKERRQ(ierr);
    ierr = PetscDSSetJacobian(ds, 0, 0, NULL, NULL, NULL, poisson::dflux);CHKERRQ(ierr);
  }
  ierr = DMAddBoundary(dm, DM_BC_ESSENTIAL, "wall", "marker", 0, 0, NULL, (void (*)(void))cfg->exact[0], NULL, 1, &wall, cfg);CHKERRQ(ierr);
  ierr = PetscDSSetExactSolution(ds, 0, cfg->exact[0], cfg);CHKERRQ(ierr);
  ierr = PetscFEDestroy(&fe);CHKERRQ(ierr);
  return 0;
}

/* -run_type full: zero guess, one Newton solve, error of the result (ex12.cpp:1597-1618) */
static PetscErrorCode run_full(SNES snes, DM dm, Vec u, Config *cfg) {
  PetscSimplePointFunc zero_guess[1] = {poisson::guess};
  PetscReal l2 = 0.0;
  PetscInt newton = 0, linear = 0, ndof = 0;
  PetscErrorCode ierr;
  ierr = DMProjectFunction(dm, 0.0, zero_guess, NULL, INSERT_VALUES, u);CHKERRQ(ierr);
  ierr = SNESSolve(snes, NULL, u);CHKERRQ(ierr);
  ierr = SNESGetSolution(snes, &u);CHKERRQ(ierr);
  ierr = DMComputeL2Diff(dm, 0.0, cfg->exact, NULL, u, &l2);CHKERRQ(ierr);
  ierr = SNESGetIterationNumber(snes, &newton);CHKERRQ(ierr);
  ierr = SNESGetLinearSolveIterations(snes, &linear);CHKERRQ(ierr);
  ierr = VecGetSize(u, &ndof);CHKERRQ(ierr);
  ierr = PetscPrintf(PETSC_COMM_WORLD, "N: %d Newton its: %d CG its: %d L_2 Error: %.6g\n", ndof, newton, linear, (double)l2);CHKERRQ(ierr);
  ierr = VecViewFromOptions(u, NULL, "-vec_view");CHKERRQ(ierr);
  return 0;
}

/* -run_type perf: one residual evaluation (ex12.cpp:1619-1628) */
static PetscErrorCode run_perf(SNES snes, DM dm, Vec u) {
  PetscSimplePointFunc zero_guess[1] = {poisson::guess};
  Vec residual;
  PetscReal norm = 0.0;
  PetscErrorCode ierr;
  ierr = DMProjectFunction(dm, 0.0, zero_guess, NULL, INSERT_VALUES, u);CHKERRQ(ierr);
  ierr = VecDuplicate(u, &residual);CHKERRQ(ierr);
  ierr = SNESComputeFunction(snes, u, residual);CHKERRQ(ierr);
  ierr = VecChop(residual, 1.0e-10);CHKERRQ(ierr);
  ierr = VecNorm(residual, NORM_2, &norm);CHKERRQ(ierr);
  ierr = PetscPrintf(PETSC_COMM_WORLD, "Initial Residual\n L_2 Residual: %g\n", (double)norm);CHKERRQ(ierr);
  ierr = VecDestroy(&residual);CHKERRQ(ierr);
  return 0;
}

/* -run_type test: error and residual of the projected exact solution, and A u + F(0) = F(u) (ex12.cpp:1629-1661) */
static PetscErrorCode run_test(SNES snes, DM dm, Vec u, Mat A, Config *cfg) {
  Vec work, f0;
  PetscReal value = 0.0;
  PetscErrorCode ierr;
  ierr = DMComputeL2Diff(dm, 0.0, cfg->exact, NULL, u, &value);CHKERRQ(ierr);
  ierr = PetscPrintf(PETSC_COMM_WORLD, "L_2 Error: %g\n", (double)value);CHKERRQ(ierr);
  ierr = VecDuplicate(u, &work);CHKERRQ(ierr);
  ierr = VecDuplicate(u, &f0);CHKERRQ(ierr);
  ierr = SNESComputeFunction(snes, u, work);CHKERRQ(ierr);
  ierr = VecChop(work, 1.0e-10);CHKERRQ(ierr);
  ierr = VecNorm(work, NORM_2, &value);CHKERRQ(ierr);
  ierr = PetscPrintf(PETSC_COMM_WORLD, "L_2 Residual: %g\n", (double)value);CHKERRQ(ierr);
  ierr = SNESComputeJacobian(snes, u, A, A);CHKERRQ(ierr);
  ierr = VecSet(work, 0.0);CHKERRQ(ierr);
  ierr = SNESComputeFunction(snes, work, f0);CHKERRQ(ierr);   /* F(0) */
  ierr = MatMult(A, u, work);CHKERRQ(ierr);                   /* A u  */
  ierr = VecAXPY(work, 1.0, f0);CHKERRQ(ierr);
  ierr = VecChop(work, 1.0e-10);CHKERRQ(ierr);
  ierr = VecNorm(work, NORM_2, &value);CHKERRQ(ierr);
  ierr = PetscPrintf(PETSC_COMM_WORLD, "Linear L_2 Residual: %g\n", (double)value);CHKERRQ(ierr);
  ierr = VecDestroy(&f0);CHKERRQ(ierr);
  ierr = VecDestroy(&work);CHKERRQ(ierr);
  return 0;
}

int main(int argc, char **argv) {
  Config cfg;
  DM dm;
  SNES snes;
  Vec u;
  Mat J;
  PetscErrorCode ierr;

  ierr = PetscInitialize(&argc, &argv, NULL, "P1 Poisson on simplices, B200 path\n");
  if (ierr) { fprintf(stderr, "%s\n", ohp_last_error()); return ierr; }
  ierr = read_options(&cfg);CHKERRQ(ierr);
  ierr = SNESCreate(PETSC_COMM_WORLD, &snes);CHKERRQ(ierr);
  {
    Omega_h::Library lib(&argc, &argv);
    ierr = create_dm(&cfg, lib, &dm);CHKERRQ(ierr);
  }
  ierr = SNESSetDM(snes, dm);CHKERRQ(ierr);
  ierr = DMSetApplicationContext(dm, &cfg);CHKERRQ(ierr);
  ierr = setup_problem(dm, &cfg);CHKERRQ(ierr);

  ierr = DMCreateGlobalVector(dm, &u);CHKERRQ(ierr);
  ierr = VecSetName(u, "potential");CHKERRQ(ierr);
  ierr = DMCreateMatrix(dm, &J);CHKERRQ(ierr);
  ierr = DMPlexSetSNESLocalFEM(dm, &cfg, &cfg, &cfg);CHKERRQ(ierr);
  ierr = SNESSetJacobian(snes, J, J, NULL, NULL);CHKERRQ(ierr);
  ierr = SNESSetFromOptions(snes);CHKERRQ(ierr);
  ierr = DMProjectFunction(dm, 0.0, cfg.exact, NULL, INSERT_ALL_VALUES, u);CHKERRQ(ierr);

  switch (cfg.run) {
    case Config::FULL: ierr = run_full(snes, dm, u, &cfg);CHKERRQ(ierr); break;
    case Config::PERF: ierr = run_perf(snes, dm, u);CHKERRQ(ierr); break;
    case Config::TEST: ierr = run_test(snes, dm, u, J, &cfg);CHKERRQ(ierr); break;
  }

  ierr = MatDestroy(&J);CHKERRQ(ierr);
  ierr = VecDestroy(&u);CHKERRQ(ierr);
  ierr = SNESDestroy(&snes);CHKERRQ(ierr);
  ierr = DMDestroy(&dm);CHKERRQ(ierr);
  return PetscFinalize();
}
