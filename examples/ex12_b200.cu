/*
 * ex12_b200.cu -- an ex12-style driver for the P1 Poisson hot path on top of include/ohp_petsc.h and
 * include/ohp_omega_h.h.  It shows that the reference's PETSc/Omega_h call surface (ex12.cpp:1472-1705)
 * carries over: the same option names, the same DM/DS/SNES calls in the same order, pointwise callbacks
 * with the PetscDS ABI -- now compiled for the device and bound with OHP_DEFINE_PROBLEM.
 * INTEGRATION.md lists what a maintainer edits in ex12.cpp itself.
 *
 *   ./ex12_b200 -mesh <dir.osh | box | picpart> [-picpart_path <prefix>] [-cells nx,ny] [-run_type full|test|perf]
 *               [-variable_coefficient none|analytic|circle|cross] [-bc_type dirichlet|neumann] [-ksp_rtol 1e-9] [-ksp_max_it N]
 *               [-pc_type none|jacobi|ksp [-ksp_ksp_max_it k] [-ksp_ksp_chebyshev_eigenvalues emin,emax]]
 *               [-boundary device|plex] [-snes_monitor_short] [-ksp_monitor_short] [-snes_converged_reason]
 *               [-ksp_converged_reason] [-dm_view] [-vec_view vtk:out.vtu]
 * Partitioned runs (one process per GPU, no MPI needed):
 *   python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 --no-python \
 *       ./ex12_b200 -mesh picpart -picpart_path <dir>/picpart     (rank r reads <dir>/picpart<r>.osh, ex12.cpp:808-813)
 */
#include <assert.h>

#include <set>
#include <string>

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
/* natural boundary data of -bc_type neumann: -du/dn of the manufactured solution, and no boundary flux term
 * (f0_bd_u / f1_bd_zero, ex12.cpp:179-195); OHP_BD_ARGS = OHP_PF_ARGS with the unit outward normal n[] after x[] */
PETSC_HOSTDEVICE static void bd_source(OHP_BD_ARGS, PetscScalar f0[]) {
  f0[0] = 0.0;
  for (PetscInt d = 0; d < dim; ++d) f0[0] += -n[d] * 2.0 * x[d];
}
PETSC_HOSTDEVICE static void bd_flux(OHP_BD_ARGS, PetscScalar f1[]) {
  for (PetscInt d = 0; d < dim; ++d) f1[d] = 0.0;
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

/* "circle" and "cross" forcing terms with their own Dirichlet data (ex12.cpp:127-145,155-177) */
PETSC_HOSTDEVICE static PetscErrorCode circle2d(PetscInt, PetscReal, const PetscReal x[], PetscInt, PetscScalar *u, void *) {
  const PetscReal alpha = 500., r2 = (x[0] - 0.5) * (x[0] - 0.5) + (x[1] - 0.5) * (x[1] - 0.5);
  u[0] = tanh(alpha * (0.15 * 0.15 - r2)) + 1.0;
  return 0;
}
PETSC_HOSTDEVICE static void source_circle(OHP_PF_ARGS, PetscScalar f0[]) {
  const PetscReal alpha = 500., r2 = (x[0] - 0.5) * (x[0] - 0.5) + (x[1] - 0.5) * (x[1] - 0.5);
  const PetscReal xi = alpha * (0.15 * 0.15 - r2), sech = 1.0 / cosh(xi);
  f0[0] = (-4.0 * alpha - 8.0 * alpha * alpha * r2 * tanh(xi)) * sech * sech;
}
PETSC_HOSTDEVICE static PetscReal cross_value(const PetscReal x[]) {
  const PetscReal alpha = 50 * 4, xy = (x[0] - 0.5) * (x[1] - 0.5);
  return sin(alpha * xy) * (alpha * fabs(xy) < 2 * 3.14159265358979323846 ? 1 : 0.01);
}
PETSC_HOSTDEVICE static PetscErrorCode cross2d(PetscInt, PetscReal, const PetscReal x[], PetscInt, PetscScalar *u, void *) { u[0] = cross_value(x); return 0; }
PETSC_HOSTDEVICE static void source_cross(OHP_PF_ARGS, PetscScalar f0[]) { f0[0] = cross_value(x); }

}  // namespace poisson

/* one line per callback set: instantiates the assembly kernels with the callbacks inlined */
OHP_DEFINE_PROBLEM_BD(constant, poisson::source, poisson::flux, poisson::dflux, poisson::exact2d, poisson::exact2d, poisson::bd_source, poisson::bd_flux)
OHP_DEFINE_PROBLEM(variable, poisson::source_nu, poisson::flux_nu, poisson::dflux_nu, poisson::exact2d, poisson::exact2d)
OHP_DEFINE_PROBLEM(circle, poisson::source_circle, poisson::flux, poisson::dflux, poisson::circle2d, poisson::circle2d)
OHP_DEFINE_PROBLEM(cross, poisson::source_cross, poisson::flux, poisson::dflux, poisson::cross2d, poisson::cross2d)
/* any other function handed to DMProjectFunction (here the zero initial guess, ex12.cpp:76,1598-1602) */
OHP_DEFINE_FUNCTION(guess, poisson::guess)

/* ---------------------------------------------------------------------------------------------
 * Run configuration: the AppCtx fields the hot path reads (ex12.cpp:43-74, options at :430-513)
 * --------------------------------------------------------------------------------------------- */
struct Config {
  enum Run { FULL, TEST, PERF } run = FULL;
  enum Coeff { NONE, ANALYTIC, CIRCLE, CROSS } coeff = NONE;
  enum BC { NEUMANN, DIRICHLET } bc = DIRICHLET;   /* -bc_type (ex12.cpp:39,432) */
  bool plex_boundary = false;   /* -boundary plex: walk the edges as ex12.cpp:986-1033 does, instead of the device detection */
  PetscInt dim = 2;
  PetscInt cells[3] = {2, 2, 2};
  char mesh[1024] = "box";
  char picpart_path[1024] = "";
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
    if (!strcmp(word, "analytic")) cfg->coeff = Config::ANALYTIC;
    else if (!strcmp(word, "circle")) { cfg->coeff = Config::CIRCLE; cfg->exact[0] = poisson::circle2d; }
    else if (!strcmp(word, "cross")) { cfg->coeff = Config::CROSS; cfg->exact[0] = poisson::cross2d; }
    else if (strcmp(word, "none")) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, "-variable_coefficient %s is not built", word);
  }
  ierr = PetscOptionsGetString(NULL, NULL, "-bc_type", word, sizeof(word), &given);CHKERRQ(ierr);
  if (given) {
    if (!strcmp(word, "neumann")) cfg->bc = Config::NEUMANN;
    else if (strcmp(word, "dirichlet")) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, "-bc_type %s is not built", word);
  }
  if (cfg->bc == Config::NEUMANN && cfg->coeff != Config::NONE) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_SUP, "-bc_type neumann is wired for -variable_coefficient none here");
  ierr = PetscOptionsGetString(NULL, NULL, "-boundary", word, sizeof(word), &given);CHKERRQ(ierr);
  if (given) cfg->plex_boundary = !strcmp(word, "plex");
  ierr = PetscOptionsGetIntArray(NULL, NULL, "-cells", cfg->cells, &n, NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetString(NULL, NULL, "-mesh", cfg->mesh, sizeof(cfg->mesh), NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetString(NULL, NULL, "-picpart_path", cfg->picpart_path, sizeof(cfg->picpart_path), NULL);CHKERRQ(ierr);
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

/* -mesh picpart (ex12.cpp:808-813, 830-882): every rank reads ITS picpart file, extracts the part of the mesh it
 * works on (its owned elements plus the layer of elements touching a vertex it owns, so that every owned row is
 * assembled locally), finds for each ghost vertex the index it has on its owner, and hands PETSc-style arrays to
 * DMPlexCreateFromCellListPetsc / PetscSFSetGraph exactly as the reference does. */
static PetscErrorCode create_dm_picpart(Config *cfg, Omega_h::Library &lib, DM *dm) {
  PetscErrorCode ierr;
  int rank = 0, size = 1;
  MPI_Comm_rank(PETSC_COMM_WORLD, &rank);
  MPI_Comm_size(PETSC_COMM_WORLD, &size);
  Omega_h::Mesh mesh(&lib);
  const std::string path = std::string(cfg->picpart_path) + std::to_string(rank) + ".osh";
  try {
    Omega_h::binary::read(path, lib.self(), &mesh);                                        /* ex12.cpp:813 */
  } catch (const std::exception &e) {
    SETERRQ1(PETSC_COMM_SELF, OHP_ERR_FILE_OPEN, "%s", e.what());
  }
  ierr = MPI_Barrier(PETSC_COMM_WORLD);CHKERRQ(ierr);                                       /* ex12.cpp:814 */
  cfg->dim = mesh.dim();
  const PetscInt dim = cfg->dim, nc = dim + 1, nPartV = mesh.nverts(), nPartE = mesh.nelems();
  const Omega_h::HostRead<Omega_h::LO> ev(mesh.ask_elem_verts());
  const Omega_h::HostRead<Omega_h::Real> xyz(mesh.coords());
  const Omega_h::HostRead<Omega_h::LO> eown(mesh.get_array<Omega_h::LO>(dim, "ownership"));  /* ex12.cpp:586 */
  const Omega_h::HostRead<Omega_h::LO> vown(mesh.get_array<Omega_h::LO>(0, "ownership"));    /* ex12.cpp:627 */
  const Omega_h::HostRead<Omega_h::GO> gids(mesh.get_array<Omega_h::GO>(0, "gids"));         /* ex12.cpp:704 */
  /* getPicPartCoreVtxCoords + getPicPartCoreElmToVtxArray + the ghost list of getPicPartCoreVtxOwnerIdx
   * (ex12.cpp:584-730: six device lambdas and three scans) in one device call */
  ohp_core_counts n;
  std::vector<int> cells((size_t)nPartE * nc), part2core(nPartV), ghostLocIdx(nPartV), ghostOwner(nPartV), roots(nPartV);
  std::vector<double> coords((size_t)nPartV * dim);
  std::vector<int64_t> ghostGid(nPartV), coreGid(nPartV);
  ierr = ohp_picpart_extract(ohp::global_context(), dim, rank, /*with_overlap=*/1, nPartE, nPartV, ev.data(), xyz.data(), eown.data(),
                             vown.data(), gids.data(), &n, cells.data(), coords.data(), part2core.data(), ghostLocIdx.data(),
                             ghostGid.data(), ghostOwner.data(), coreGid.data());CHKERRQ(ierr);
  assert(n.n_core_elems > 0);                                                                /* ex12.cpp:840 */
  /* the gid round trip (ex12.cpp:731-777): index of every ghost in its owner's vertex array */
  std::vector<uint8_t> owned(n.n_core_verts, 1);
  for (int i = 0; i < n.n_ghost_verts; ++i) owned[ghostLocIdx[i]] = 0;
  ierr = ohp_ghost_roots_from_gids(ohp::global_context(), n.n_ghost_verts, ghostGid.data(), ghostOwner.data(), n.n_core_verts,
                                   coreGid.data(), owned.data(), roots.data());CHKERRQ(ierr);
  /* Plex points of the roots are offset by the OWNER's cell count (ex12.cpp:757, getNeighborElmCounts :528-577) */
  std::vector<int64_t> cellCount(size);
  ierr = ohp_comm_allgather_i64(ohp::global_context(), n.n_core_elems, cellCount.data());CHKERRQ(ierr);
  ierr = DMPlexCreateFromCellListPetsc(PETSC_COMM_WORLD, dim, n.n_core_elems, n.n_core_verts, nc, PETSC_FALSE, cells.data(), dim,
                                       coords.data(), dm);CHKERRQ(ierr);                    /* ex12.cpp:859 */
  PetscInt *localVertex;
  PetscSFNode *remoteVertex;
  ierr = PetscMalloc1(n.n_ghost_verts, &localVertex);CHKERRQ(ierr);
  ierr = PetscMalloc1(n.n_ghost_verts, &remoteVertex);CHKERRQ(ierr);
  for (int i = 0; i < n.n_ghost_verts; ++i) {                                                /* ex12.cpp:867-871 */
    localVertex[i] = n.n_core_elems + ghostLocIdx[i];
    remoteVertex[i].rank = ghostOwner[i];
    remoteVertex[i].index = (PetscInt)cellCount[ghostOwner[i]] + roots[i];
  }
  PetscSF pointSF;
  ierr = DMGetPointSF(*dm, &pointSF);CHKERRQ(ierr);
  ierr = PetscSFSetGraph(pointSF, n.n_core_elems + n.n_core_verts, n.n_ghost_verts, localVertex, PETSC_OWN_POINTER, remoteVertex,
                         PETSC_OWN_POINTER);CHKERRQ(ierr);                                  /* ex12.cpp:875 */
  return 0;
}

/* boundary vertices the way CreateQuadMesh finds them (ex12.cpp:966-1033): edges with one supporting cell, their cones,
 * DMSetLabelValue point by point.  With the overlap layer every edge at an OWNED vertex has its full support locally, and
 * ghost vertices take their owner's label inside the library, so the label is the global boundary (hazard H2 of the
 * reference, interface edges marked as boundary, does not arise). */
static PetscErrorCode mark_boundary_like_ex12(DM dm) {
  PetscInt cStart, cEnd, vStart, vEnd, eStart, eEnd;
  PetscErrorCode ierr;
  ierr = DMPlexGetHeightStratum(dm, 0, &cStart, &cEnd);CHKERRQ(ierr);
  ierr = DMPlexGetHeightStratum(dm, 1, &eStart, &eEnd);CHKERRQ(ierr);
  ierr = DMPlexGetHeightStratum(dm, 2, &vStart, &vEnd);CHKERRQ(ierr);
  std::vector<int> boundary_edge;
  for (int i = eStart; i < eEnd; i++) {
    int support_size;
    ierr = DMPlexGetSupportSize(dm, i, &support_size);CHKERRQ(ierr);
    if (support_size == 1) boundary_edge.push_back(i);
  }
  std::set<int> boundary_vert;
  for (unsigned int i = 0; i < boundary_edge.size(); i++) {
    const int *verts;
    ierr = DMPlexGetCone(dm, boundary_edge[i], &verts);CHKERRQ(ierr);
    boundary_vert.insert(verts[0]);
    boundary_vert.insert(verts[1]);
  }
  for (auto i = boundary_vert.begin(); i != boundary_vert.end(); i++) { ierr = DMSetLabelValue(dm, "marker", *i, 1);CHKERRQ(ierr); }
  for (unsigned int i = 0; i < boundary_edge.size(); i++) { ierr = DMSetLabelValue(dm, "marker", boundary_edge[i], 1);CHKERRQ(ierr); }
  return 0;
}

static PetscErrorCode create_dm(Config *cfg, Omega_h::Library &lib, DM *dm) {
  std::vector<int> cells;
  std::vector<double> coords;
  PetscErrorCode ierr;
  if (!strcmp(cfg->mesh, "picpart")) {
    ierr = create_dm_picpart(cfg, lib, dm);CHKERRQ(ierr);
  } else if (!strcmp(cfg->mesh, "box")) {
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
  if (strcmp(cfg->mesh, "picpart")) {
    const PetscInt nc = cfg->dim + 1;
    ierr = DMPlexCreateFromCellListPetsc(PETSC_COMM_WORLD, cfg->dim, (PetscInt)(cells.size() / nc), (PetscInt)(coords.size() / cfg->dim),
                                         nc, PETSC_FALSE, cells.data(), cfg->dim, coords.data(), dm);CHKERRQ(ierr);   /* ex12.cpp:933 */
  }
  DM interpolated;
  ierr = DMPlexInterpolate(*dm, &interpolated);CHKERRQ(ierr);                                                        /* ex12.cpp:958 */
  ierr = DMDestroy(dm);CHKERRQ(ierr);
  *dm = interpolated;
  if (cfg->plex_boundary) {
    ierr = mark_boundary_like_ex12(*dm);CHKERRQ(ierr);                                                               /* ex12.cpp:966-1033 */
  } else {
    /* "marker" label on the boundary: the two calls of CreateBCLabel (ex12.cpp:515-526) instead of the
     * host loops over edge supports; evaluated on the device */
    DMLabel marker;
    ierr = DMCreateLabel(*dm, "marker");CHKERRQ(ierr);
    ierr = DMGetLabel(*dm, "marker", &marker);CHKERRQ(ierr);
    ierr = DMPlexMarkBoundaryFaces(*dm, 1, marker);CHKERRQ(ierr);
    ierr = DMPlexLabelComplete(*dm, marker);CHKERRQ(ierr);
  }
  ierr = DMViewFromOptions(*dm, NULL, "-dm_view");CHKERRQ(ierr);                                                     /* ex12.cpp:962 */
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
  switch (cfg->coeff) {                                                         /* ex12.cpp:1170-1204 */
    case Config::ANALYTIC:
      ierr = PetscDSSetResidual(ds, 0, poisson::source_nu, poisson::flux_nu);CHKERRQ(ierr);
      ierr = PetscDSSetJacobian(ds, 0, 0, NULL, NULL, NULL, poisson::dflux_nu);CHKERRQ(ierr);
      break;
    case Config::CIRCLE:
      ierr = PetscDSSetResidual(ds, 0, poisson::source_circle, poisson::flux);CHKERRQ(ierr);
      ierr = PetscDSSetJacobian(ds, 0, 0, NULL, NULL, NULL, poisson::dflux);CHKERRQ(ierr);
      break;
    case Config::CROSS:
      ierr = PetscDSSetResidual(ds, 0, poisson::source_cross, poisson::flux);CHKERRQ(ierr);
      ierr = PetscDSSetJacobian(ds, 0, 0, NULL, NULL, NULL, poisson::dflux);CHKERRQ(ierr);
      break;
    default:
      ierr = PetscDSSetResidual(ds, 0, poisson::source, poisson::flux);CHKERRQ(ierr);
      ierr = PetscDSSetJacobian(ds, 0, 0, NULL, NULL, NULL, poisson::dflux);CHKERRQ(ierr);
  }
  if (cfg->bc == Config::NEUMANN) {                                             /* ex12.cpp:1088-1094,1226,1237 */
    DMLabel label;
    ierr = DMCreateLabel(dm, "boundary");CHKERRQ(ierr);
    ierr = DMGetLabel(dm, "boundary", &label);CHKERRQ(ierr);
    ierr = DMPlexMarkBoundaryFaces(dm, 1, label);CHKERRQ(ierr);
    ierr = PetscDSSetBdResidual(ds, 0, poisson::bd_source, poisson::bd_flux);CHKERRQ(ierr);
  }
  ierr = DMAddBoundary(dm, cfg->bc == Config::DIRICHLET ? DM_BC_ESSENTIAL : DM_BC_NATURAL, "wall", cfg->bc == Config::DIRICHLET ? "marker" : "boundary",
                       0, 0, NULL, (void (*)(void))cfg->exact[0], NULL, 1, &wall, cfg);CHKERRQ(ierr);
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
  MatNullSpace nullSpace = NULL;                                /* "May be necessary for Neumann conditions", ex12.cpp:1478,1534-1538 */
  if (cfg.bc != Config::DIRICHLET) {
    ierr = MatNullSpaceCreate(PETSC_COMM_WORLD, PETSC_TRUE, 0, NULL, &nullSpace);CHKERRQ(ierr);
    ierr = MatSetNullSpace(J, nullSpace);CHKERRQ(ierr);
  }
  ierr = DMPlexSetSNESLocalFEM(dm, &cfg, &cfg, &cfg);CHKERRQ(ierr);
  ierr = SNESSetJacobian(snes, J, J, NULL, NULL);CHKERRQ(ierr);
  ierr = SNESSetFromOptions(snes);CHKERRQ(ierr);
  ierr = DMProjectFunction(dm, 0.0, cfg.exact, NULL, INSERT_ALL_VALUES, u);CHKERRQ(ierr);

  switch (cfg.run) {
    case Config::FULL: ierr = run_full(snes, dm, u, &cfg);CHKERRQ(ierr); break;
    case Config::PERF: ierr = run_perf(snes, dm, u);CHKERRQ(ierr); break;
    case Config::TEST: ierr = run_test(snes, dm, u, J, &cfg);CHKERRQ(ierr); break;
  }

  ierr = MatNullSpaceDestroy(&nullSpace);CHKERRQ(ierr);
  ierr = MatDestroy(&J);CHKERRQ(ierr);
  ierr = VecDestroy(&u);CHKERRQ(ierr);
  ierr = SNESDestroy(&snes);CHKERRQ(ierr);
  ierr = DMDestroy(&dm);CHKERRQ(ierr);
  return PetscFinalize();
}
