/*
 * ohp_petsc.h -- PETSc-shaped C++ facade over the C-ABI (include/ohp.h), so that an ex12-style
 * main() keeps its call sequence and its pointwise callbacks (ex12.cpp:1472-1703) while every
 * numerical step runs in libohp.so's sm_100a kernels.
 *
 * Only the calls on the reference's hot path are provided (SURVEY 8b "object calls on the path"):
 *   SNESCreate/SetDM/SetJacobian/SetFromOptions/Solve/GetSolution/GetDM/ComputeFunction/ComputeJacobian,
 *   DMPlexCreateFromCellListPetsc, DMGetPointSF + PetscSFSetGraph, DMPlexInterpolate (identity),
 *   DMCreateLabel + DMPlexMarkBoundaryFaces (the reference's CreateBCLabel, ex12.cpp:515-526; replaces
 *   the host loops over edge supports at ex12.cpp:986-1033), PetscFECreateDefault, DMSetField, DMCreateDS,
 *   DMGetDS, PetscDSSetResidual/SetJacobian/SetExactSolution, DMAddBoundary, DMCreateGlobalVector,
 *   DMCreateMatrix, DMPlexSetSNESLocalFEM, DMProjectFunction, DMComputeL2Diff, MatMult,
 *   VecSet/AXPY/AYPX/Norm/Dot/Chop/Copy/Duplicate/GetSize, destroy calls, a PetscOptions subset.
 * Same names, argument order and meaning, PetscErrorCode returns, CHKERRQ propagation.
 *
 * Callbacks run ON THE DEVICE, so a translation unit that defines them is compiled by nvcc (the
 * reference already compiles ex12.cpp as CUDA, CMakeLists.txt:33-37) and needs two additions:
 *   1. qualify each callback with PETSC_HOSTDEVICE (= __host__ __device__);
 *   2. one line per callback set, after the definitions:
 *        OHP_DEFINE_PROBLEM(name, f0, f1, g3, bc, exact)
 *      which instantiates the assembly kernels of include/ohp_fem.cuh with the callbacks inlined and
 *      registers them under the tuple of host function pointers; PetscDSSetResidual/SetJacobian/
 *      DMAddBoundary then find the instantiation from the pointers they are given at run time.
 * A callback set that was not declared this way is refused with PETSC_ERR_SUP (no silent CPU path).
 */
#ifndef OHP_PETSC_H
#define OHP_PETSC_H

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "ohp.h"
#ifdef __CUDACC__
#include "ohp_fem.cuh"
#define PETSC_HOSTDEVICE __host__ __device__
#else
#define PETSC_HOSTDEVICE
#endif

/* ---- scalar types (PETSc configured with 32-bit indices and real double scalars; hazard H4) ---- */
typedef int PetscErrorCode;
typedef int32_t PetscInt;
typedef double PetscReal;
typedef double PetscScalar;
typedef int PetscMPIInt;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef int MPI_Comm;
#define PETSC_COMM_WORLD 0
#define PETSC_COMM_SELF 1
#ifndef MPI_COMM_WORLD
#define MPI_COMM_WORLD PETSC_COMM_WORLD
#define MPI_COMM_SELF PETSC_COMM_SELF
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_INT 1
#define MPI_SUM 1
#endif
#define PETSC_DEFAULT (-2)
#define PETSC_ERR_SUP OHP_ERR_SUP
#define PETSC_ERR_ARG_WRONG OHP_ERR_ARG_WRONG
#define PETSC_ERR_ARG_OUTOFRANGE OHP_ERR_ARG_OUTOFRANGE
#define PETSC_ERR_ORDER OHP_ERR_ORDER
#define PETSC_ERR_LIB OHP_ERR_LIB
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_FROBENIUS = 2, NORM_INFINITY = 3 } NormType;
typedef enum { INSERT_VALUES = 1, ADD_VALUES = 2, INSERT_ALL_VALUES = 5, INSERT_BC_VALUES = 7 } InsertMode;
typedef enum { DM_BC_ESSENTIAL = 1, DM_BC_ESSENTIAL_FIELD = 5, DM_BC_NATURAL = 2 } DMBoundaryConditionType;
typedef enum { PETSC_COPY_VALUES, PETSC_OWN_POINTER, PETSC_USE_POINTER } PetscCopyMode;
typedef struct { PetscInt rank, index; } PetscSFNode;

/* ---- error handling: `ierr = f();CHKERRQ(ierr);` (ex12.cpp:1491-1501) ---- */
#define PetscFunctionBeginUser do { } while (0)
#define PetscFunctionReturn(x) return (x)
#define CHKERRQ(ierr)                                                                               \
  do {                                                                                              \
    if (ierr) {                                                                                     \
      fprintf(stderr, "[0]PETSC ERROR: %s() line %d in %s: error %d: %s\n", __func__, __LINE__, __FILE__, (int)(ierr), \
              ohp_last_error());                                                                    \
      return (ierr);                                                                                \
    }                                                                                               \
  } while (0)
#define SETERRQ(comm, code, msg) do { fprintf(stderr, "[0]PETSC ERROR: %s\n", msg); return (code); } while (0)
#define SETERRQ1(comm, code, fmt, a) do { fprintf(stderr, "[0]PETSC ERROR: " fmt "\n", a); return (code); } while (0)
#define SETERRQ2(comm, code, fmt, a, b) do { fprintf(stderr, "[0]PETSC ERROR: " fmt "\n", a, b); return (code); } while (0)

/* ---- callback ABI (ex12.cpp:147-150, :209-212, :113) ---- */
typedef void (*PetscPointFunc)(PetscInt, PetscInt, PetscInt, const PetscInt[], const PetscInt[], const PetscScalar[],
                               const PetscScalar[], const PetscScalar[], const PetscInt[], const PetscInt[],
                               const PetscScalar[], const PetscScalar[], const PetscScalar[], PetscReal, const PetscReal[],
                               PetscInt, const PetscScalar[], PetscScalar[]);
typedef void (*PetscPointJac)(PetscInt, PetscInt, PetscInt, const PetscInt[], const PetscInt[], const PetscScalar[],
                              const PetscScalar[], const PetscScalar[], const PetscInt[], const PetscInt[],
                              const PetscScalar[], const PetscScalar[], const PetscScalar[], PetscReal, PetscReal,
                              const PetscReal[], PetscInt, const PetscScalar[], PetscScalar[]);
typedef PetscErrorCode (*PetscSimplePointFunc)(PetscInt, PetscReal, const PetscReal[], PetscInt, PetscScalar *, void *);
/* PetscBdPointFunc: the volume form plus the unit outward normal n[] after x[] (f0_bd_u, ex12.cpp:179-182) */
typedef void (*PetscBdPointFunc)(PetscInt, PetscInt, PetscInt, const PetscInt[], const PetscInt[], const PetscScalar[], const PetscScalar[],
                                 const PetscScalar[], const PetscInt[], const PetscInt[], const PetscScalar[], const PetscScalar[],
                                 const PetscScalar[], PetscReal, const PetscReal[], const PetscReal[], PetscInt, const PetscScalar[], PetscScalar[]);

namespace ohp {

/* ---- options database (the -key value subset SNESSetFromOptions / ProcessOptions read) ---- */
inline std::map<std::string, std::string> &options_db() { static std::map<std::string, std::string> db; return db; }
inline ohp_context &global_context() { static ohp_context c = nullptr; return c; }

struct ProblemKey {
  void *f0, *f1, *g3, *bc;
  bool operator<(const ProblemKey &o) const {
    if (f0 != o.f0) return f0 < o.f0;
    if (f1 != o.f1) return f1 < o.f1;
    if (g3 != o.g3) return g3 < o.g3;
    return bc < o.bc;
  }
};
struct ProblemRecord { int id; void *exact; void *f0_bd = nullptr, *f1_bd = nullptr; };
inline std::map<ProblemKey, ProblemRecord> &problem_registry() { static std::map<ProblemKey, ProblemRecord> r; return r; }

/* point functions compiled for the device so that DMProjectFunction can evaluate them at the vertices */
inline std::map<void *, ohp_assembly_launcher> &function_registry() { static std::map<void *, ohp_assembly_launcher> r; return r; }

#ifdef __CUDACC__
template <class F> int register_function(void *fn) { function_registry()[fn] = launch_project_fn<F>; return 0; }
template <class P> int register_problem(void *f0, void *f1, void *g3, void *bc, void *exact) {
  int id = -1;
  if (ohp_problem_register(launch_residual_any<P>, launch_jacobian_any<P>, launch_project_exact_any<P>, launch_l2_diff_any<P>, &id)) return -1;
  problem_registry()[ProblemKey{f0, f1, g3, bc}] = ProblemRecord{id, exact};
  return id;
}
/* the same with boundary callbacks (PetscDSSetBdResidual): also instantiates the natural-boundary kernel */
template <class P> int register_problem_bd(void *f0, void *f1, void *g3, void *bc, void *exact, void *f0_bd, void *f1_bd) {
  const int id = register_problem<P>(f0, f1, g3, bc, exact);
  if (id < 0 || ohp_problem_set_bd_residual(id, launch_bd_residual_any<P>)) return -1;
  ProblemRecord &r = problem_registry()[ProblemKey{f0, f1, g3, bc}];
  r.f0_bd = f0_bd; r.f1_bd = f1_bd;
  return id;
}
#endif

}  // namespace ohp

/* Instantiates the device kernels for one callback set and registers them (file scope, nvcc only). */
#define OHP_DEFINE_PROBLEM(name, F0, F1, G3, BC, EXACT)                                                         \
  struct ohp_problem_##name {                                                                                   \
    static constexpr bool kJacobianUsesU = true;                                                                \
    OHP_HD static void f0(OHP_PF_ARGS, double f[]) {                                                            \
      F0(dim, Nf, NfAux, uOff, uOff_x, u, u_t, u_x, aOff, aOff_x, a, a_t, a_x, t, x, numConstants, constants, f); \
    }                                                                                                           \
    OHP_HD static void f1(OHP_PF_ARGS, double f[]) {                                                            \
      F1(dim, Nf, NfAux, uOff, uOff_x, u, u_t, u_x, aOff, aOff_x, a, a_t, a_x, t, x, numConstants, constants, f); \
    }                                                                                                           \
    OHP_HD static void g3(OHP_PJ_ARGS, double g[]) {                                                            \
      G3(dim, Nf, NfAux, uOff, uOff_x, u, u_t, u_x, aOff, aOff_x, a, a_t, a_x, t, u_tShift, x, numConstants, constants, g); \
    }                                                                                                           \
    OHP_HD static int bc(int dim, double time, const double x[], int Nc, double *u, void *ctx) {                \
      return BC(dim, time, x, Nc, u, ctx);                                                                      \
    }                                                                                                           \
    OHP_HD static int exact(int dim, double time, const double x[], int Nc, double *u, void *ctx) {             \
      return EXACT(dim, time, x, Nc, u, ctx);                                                                   \
    }                                                                                                           \
  };                                                                                                            \
  static const int ohp_problem_id_##name = ohp::register_problem<ohp_problem_##name>(                           \
      (void *)(PetscPointFunc)F0, (void *)(PetscPointFunc)F1, (void *)(PetscPointJac)G3,                        \
      (void *)(PetscSimplePointFunc)BC, (void *)(PetscSimplePointFunc)EXACT);

/* The same for a callback set that also registers boundary callbacks with PetscDSSetBdResidual (-bc_type neumann,
 * ex12.cpp:1226,1231): F0_BD / F1_BD have the PetscBdPointFunc signature (OHP_BD_ARGS, then the output array). */
#define OHP_DEFINE_PROBLEM_BD(name, F0, F1, G3, BC, EXACT, F0_BD, F1_BD)                                        \
  struct ohp_problem_##name {                                                                                   \
    static constexpr bool kJacobianUsesU = true;                                                                \
    OHP_HD static void f0(OHP_PF_ARGS, double f[]) {                                                            \
      F0(dim, Nf, NfAux, uOff, uOff_x, u, u_t, u_x, aOff, aOff_x, a, a_t, a_x, t, x, numConstants, constants, f); \
    }                                                                                                           \
    OHP_HD static void f1(OHP_PF_ARGS, double f[]) {                                                            \
      F1(dim, Nf, NfAux, uOff, uOff_x, u, u_t, u_x, aOff, aOff_x, a, a_t, a_x, t, x, numConstants, constants, f); \
    }                                                                                                           \
    OHP_HD static void g3(OHP_PJ_ARGS, double g[]) {                                                            \
      G3(dim, Nf, NfAux, uOff, uOff_x, u, u_t, u_x, aOff, aOff_x, a, a_t, a_x, t, u_tShift, x, numConstants, constants, g); \
    }                                                                                                           \
    OHP_HD static void f0_bd(OHP_BD_ARGS, double f[]) {                                                         \
      F0_BD(dim, Nf, NfAux, uOff, uOff_x, u, u_t, u_x, aOff, aOff_x, a, a_t, a_x, t, x, n, numConstants, constants, f); \
    }                                                                                                           \
    OHP_HD static void f1_bd(OHP_BD_ARGS, double f[]) {                                                         \
      F1_BD(dim, Nf, NfAux, uOff, uOff_x, u, u_t, u_x, aOff, aOff_x, a, a_t, a_x, t, x, n, numConstants, constants, f); \
    }                                                                                                           \
    OHP_HD static int bc(int dim, double time, const double x[], int Nc, double *u, void *ctx) {                \
      return BC(dim, time, x, Nc, u, ctx);                                                                      \
    }                                                                                                           \
    OHP_HD static int exact(int dim, double time, const double x[], int Nc, double *u, void *ctx) {             \
      return EXACT(dim, time, x, Nc, u, ctx);                                                                   \
    }                                                                                                           \
  };                                                                                                            \
  static const int ohp_problem_id_##name = ohp::register_problem_bd<ohp_problem_##name>(                        \
      (void *)(PetscPointFunc)F0, (void *)(PetscPointFunc)F1, (void *)(PetscPointJac)G3,                        \
      (void *)(PetscSimplePointFunc)BC, (void *)(PetscSimplePointFunc)EXACT, (void *)(PetscBdPointFunc)F0_BD,    \
      (void *)(PetscBdPointFunc)F1_BD);

/* Makes a PetscSimplePointFunc (initial guess, any function handed to DMProjectFunction) available on the device */
#define OHP_DEFINE_FUNCTION(name, FN)                                                                           \
  struct ohp_function_##name {                                                                                  \
    OHP_HD static int eval(int dim, double time, const double x[], int Nc, double *u, void *ctx) {              \
      return FN(dim, time, x, Nc, u, ctx);                                                                      \
    }                                                                                                           \
  };                                                                                                            \
  static const int ohp_function_id_##name = ohp::register_function<ohp_function_##name>((void *)(PetscSimplePointFunc)FN);

/* ------------------------------------------------------------------------------------------------ */
/* objects                                                                                          */
/* ------------------------------------------------------------------------------------------------ */
struct _p_PetscSF {
  struct _p_DM *dm;
};
struct _p_PetscDS {
  PetscPointFunc f0 = nullptr, f1 = nullptr;
  PetscPointJac g3 = nullptr;
  PetscSimplePointFunc bc = nullptr, exact = nullptr;
  PetscBdPointFunc f0_bd = nullptr, f1_bd = nullptr;
  void *exact_ctx = nullptr;
  int problem = -1;
};
struct _p_PetscFE { PetscInt dim, Nc; PetscBool simplex; };
struct _p_DMLabel { std::string name; };
struct _p_DM {
  ohp_mesh mesh = nullptr;
  int refs = 1;
  bool set_up = false;
  PetscInt nC = 0, nV = 0;
  /* Plex queries (ex12.cpp:966-1010): host copies of the facets, fetched once */
  bool facets = false;
  PetscInt nF = 0;
  std::vector<PetscInt> facet_cone, facet_support, cell_cone; /* cones as Plex POINT numbers */
  /* DMSetLabelValue("marker", vertex point, 1) calls (ex12.cpp:1023-1026) */
  bool label_by_caller = false;
  std::vector<PetscInt> label_verts;
  void *app_ctx = nullptr;
  _p_PetscSF sf;
  _p_PetscDS ds;
  _p_DMLabel marker, boundary;
  bool has_marker = false, has_field = false, has_bc = false;
  int bc_type = OHP_BC_ESSENTIAL; /* DMAddBoundary's type: DM_BC_NATURAL (-bc_type neumann) constrains nothing */
  PetscInt dim = 0;
};
struct _p_Vec { ohp_vec v = nullptr; _p_DM *dm = nullptr; std::string name; };
struct _p_Mat { ohp_mat m = nullptr; _p_DM *dm = nullptr; };
struct _p_MatNullSpace { PetscBool has_cnst; };
struct _p_SNES {
  _p_DM *dm = nullptr;
  _p_Mat *A = nullptr, *J = nullptr;
  _p_Vec *sol = nullptr;
  ohp_snes_options opts;
  ohp_snes_result last;
  bool monitor = false, ksp_monitor = false, converged_reason = false, ksp_converged_reason = false;
  std::string ksp_type = "cg";
  int last_ksp_reason = 0, last_ksp_its = 0;
};
typedef _p_DM *DM;
typedef _p_Vec *Vec;
typedef _p_Mat *Mat;
typedef _p_SNES *SNES;
typedef _p_PetscDS *PetscDS;
typedef _p_PetscFE *PetscFE;
typedef _p_PetscSF *PetscSF;
typedef _p_DMLabel *DMLabel;
typedef void *PetscObject;
typedef struct _p_MatNullSpace *MatNullSpace;

/* ---- the few MPI calls of the adapter (ex12.cpp:573-575,800-801,814,844,953,978), over the library's communicator ---- */
inline int MPI_Comm_rank(MPI_Comm comm, int *rank) {
  int r = 0, n = 1;
  if (comm == PETSC_COMM_WORLD && ohp::global_context()) ohp_comm_rank_size(ohp::global_context(), &r, &n);
  *rank = r;
  return 0;
}
inline int MPI_Comm_size(MPI_Comm comm, int *size) {
  int r = 0, n = 1;
  if (comm == PETSC_COMM_WORLD && ohp::global_context()) ohp_comm_rank_size(ohp::global_context(), &r, &n);
  *size = n;
  return 0;
}
inline int MPI_Barrier(MPI_Comm comm) { return comm == PETSC_COMM_WORLD ? ohp_comm_barrier(ohp::global_context()) : 0; }
inline int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm) {
  if (type != MPI_INT || op != MPI_SUM) return PETSC_ERR_SUP;
  std::vector<int64_t> v(count);
  for (int i = 0; i < count; ++i) v[i] = ((const int *)sendbuf)[i];
  const int rc = comm == PETSC_COMM_WORLD ? ohp_comm_allreduce_sum_i64(ohp::global_context(), v.data(), count) : 0;
  for (int i = 0; i < count; ++i) ((int *)recvbuf)[i] = (int)v[i];
  return rc;
}

/* ---- init / options ---- */
/* One process per GPU.  Launched by torch.distributed.run (or anything that exports RANK / WORLD_SIZE / LOCAL_RANK /
 * MASTER_ADDR / MASTER_PORT) the ranks form the NCCL communicator here, the way PetscInitialize calls MPI_Init. */
inline PetscErrorCode PetscInitialize(int *argc, char ***argv, const char *, const char *) {
  for (int i = 1; i < *argc; ++i) {
    const char *a = (*argv)[i];
    if (a[0] == '-' && !(a[1] >= '0' && a[1] <= '9')) {
      std::string key(a + 1), val;
      if (i + 1 < *argc && !((*argv)[i + 1][0] == '-' && !((*argv)[i + 1][1] >= '0' && (*argv)[i + 1][1] <= '9') && (*argv)[i + 1][1] != '.')) val = (*argv)[++i];
      ohp::options_db()[key] = val;
    }
  }
  int rank = 0, nranks = 1, device = 0;
  PetscErrorCode ierr = ohp_boot_env(&rank, &nranks, &device);
  if (ierr) return ierr;
  ierr = ohp_init(device, &ohp::global_context());
  if (ierr) return ierr;
  if (nranks > 1) ierr = ohp_comm_init_from_env(ohp::global_context());
  return ierr;
}
inline PetscErrorCode PetscFinalize(void) {
  PetscErrorCode ierr = ohp_finalize(ohp::global_context());
  ohp::global_context() = nullptr;
  return ierr;
}
inline PetscErrorCode PetscOptionsGetString(void *, const char *, const char *name, char *out, size_t len, PetscBool *set) {
  auto it = ohp::options_db().find(name + 1);
  if (set) *set = it != ohp::options_db().end() ? PETSC_TRUE : PETSC_FALSE;
  if (it != ohp::options_db().end()) { strncpy(out, it->second.c_str(), len); out[len - 1] = 0; }
  return 0;
}
inline PetscErrorCode PetscOptionsGetInt(void *, const char *, const char *name, PetscInt *v, PetscBool *set) {
  auto it = ohp::options_db().find(name + 1);
  if (set) *set = it != ohp::options_db().end() ? PETSC_TRUE : PETSC_FALSE;
  if (it != ohp::options_db().end()) *v = atoi(it->second.c_str());
  return 0;
}
inline PetscErrorCode PetscOptionsGetReal(void *, const char *, const char *name, PetscReal *v, PetscBool *set) {
  auto it = ohp::options_db().find(name + 1);
  if (set) *set = it != ohp::options_db().end() ? PETSC_TRUE : PETSC_FALSE;
  if (it != ohp::options_db().end()) *v = atof(it->second.c_str());
  return 0;
}
inline PetscErrorCode PetscOptionsHasName(void *, const char *, const char *name, PetscBool *set) {
  *set = ohp::options_db().count(name + 1) ? PETSC_TRUE : PETSC_FALSE;
  return 0;
}
inline PetscErrorCode PetscOptionsGetIntArray(void *, const char *, const char *name, PetscInt v[], PetscInt *n, PetscBool *set) {
  auto it = ohp::options_db().find(name + 1);
  if (set) *set = it != ohp::options_db().end() ? PETSC_TRUE : PETSC_FALSE;
  if (it == ohp::options_db().end()) { *n = 0; return 0; }
  PetscInt k = 0;
  const char *s = it->second.c_str();
  while (*s && k < *n) { v[k++] = (PetscInt)strtol(s, (char **)&s, 10); if (*s == ',') ++s; }
  *n = k;
  return 0;
}
inline bool ohp_prints(MPI_Comm comm) { /* PETSC_COMM_WORLD prints on rank 0 only */
  int r = 0;
  if (comm == PETSC_COMM_WORLD) MPI_Comm_rank(comm, &r);
  return r == 0;
}
#define PetscPrintf(comm, ...) (ohp_prints(comm) ? (printf(__VA_ARGS__), fflush(stdout), 0) : 0)
inline PetscErrorCode PetscObjectSetName(PetscObject, const char *) { return 0; } /* names: see VecSetName */

/* ---- DM: mesh ---- */
inline PetscErrorCode DMPlexCreateFromCellListPetsc(MPI_Comm, PetscInt dim, PetscInt numCells, PetscInt numVertices,
                                                    PetscInt numCorners, PetscBool /*interpolate*/, const PetscInt cells[],
                                                    PetscInt spaceDim, const PetscReal vertexCoords[], DM *dm) {
  if (numCorners != dim + 1 || spaceDim != dim) {
    fprintf(stderr, "[0]PETSC ERROR: DMPlexCreateFromCellListPetsc: only simplices with spaceDim == dim are built\n");
    return PETSC_ERR_SUP;
  }
  DM d = new _p_DM();
  d->dim = dim;
  d->nC = numCells; d->nV = numVertices;
  d->sf.dm = d;
  PetscErrorCode ierr = ohp_mesh_create(ohp::global_context(), dim, numCells, numVertices, cells, vertexCoords, &d->mesh);
  if (ierr) { delete d; return ierr; }
  *dm = d;
  return 0;
}
inline PetscErrorCode DMGetPointSF(DM dm, PetscSF *sf) { *sf = &dm->sf; return 0; }
/* ex12.cpp:864-876: leaves and roots are Plex POINTS (local: nCells + vertex; remote: the owner's
 * nCells + its vertex index, ex12.cpp:868-870,924-926).  ohp_mesh_set_point_sf converts them to vertex
 * indices (it all-gathers the ranks' cell counts).  PETSC_OWN_POINTER: the arrays are released here. */
inline PetscErrorCode PetscSFSetGraph(PetscSF sf, PetscInt /*nroots*/, PetscInt nleaves, PetscInt *ilocal, PetscCopyMode lmode,
                                      PetscSFNode *iremote, PetscCopyMode rmode) {
  std::vector<int32_t> orank(nleaves), opoint(nleaves);
  for (PetscInt i = 0; i < nleaves; ++i) { orank[i] = iremote[i].rank; opoint[i] = iremote[i].index; }
  PetscErrorCode ierr = ohp_mesh_set_point_sf(sf->dm->mesh, nleaves, ilocal, orank.data(), opoint.data());
  if (lmode == PETSC_OWN_POINTER) free(ilocal);
  if (rmode == PETSC_OWN_POINTER) free(iremote);
  return ierr;
}
inline PetscErrorCode PetscMalloc1(size_t n, void *out, size_t elem) { *(void **)out = malloc(n * elem); return *(void **)out ? 0 : OHP_ERR_MEM; }
#define PetscMalloc1(n, pp) PetscMalloc1((size_t)(n), (void *)(pp), sizeof(**(pp)))
inline PetscErrorCode DMPlexInterpolate(DM dm, DM *idm) { dm->refs++; *idm = dm; return 0; } /* edges are implicit here */
/* two labels exist on this path: "marker" (vertices, ex12.cpp:521,1025) and "boundary" (facets, -bc_type neumann,
 * ex12.cpp:1091); both describe the facets with one supporting cell, which the library finds itself */
inline PetscErrorCode DMCreateLabel(DM dm, const char *name) {
  if (!strcmp(name, "boundary")) dm->boundary.name = name; else dm->marker.name = name;
  return 0;
}
inline PetscErrorCode DMGetLabel(DM dm, const char *name, DMLabel *label) { *label = !strcmp(name, "boundary") ? &dm->boundary : &dm->marker; return 0; }
inline PetscErrorCode DMHasLabel(DM dm, const char *, PetscBool *has) { *has = dm->has_marker ? PETSC_TRUE : PETSC_FALSE; return 0; }

inline PetscErrorCode ohp_dm_setup(DM dm) {
  if (dm->set_up) return 0;
  PetscErrorCode ierr = 0;
  if (dm->label_by_caller) ierr = ohp_mesh_set_label(dm->mesh, (PetscInt)dm->label_verts.size(), dm->label_verts.data());
  if (!ierr && dm->bc_type != OHP_BC_ESSENTIAL) ierr = ohp_mesh_set_bc_type(dm->mesh, dm->bc_type);
  if (!ierr) ierr = ohp_mesh_setup(dm->mesh);
  /* partitioned run: NVLink peer mapping for the fused halo / all-reduce kernels (OHP_COMM=nccl keeps plain NCCL) */
  if (!ierr) {
    int r = 0, n = 1;
    ohp_comm_rank_size(ohp::global_context(), &r, &n);
    const char *c = getenv("OHP_COMM");
    if (n > 1 && !(c && c[0] == 'n')) ierr = ohp_mesh_p2p_enable(dm->mesh);
  }
  if (!ierr) dm->set_up = true;
  return ierr;
}

/* ---- Plex queries used by CreateQuadMesh after DMPlexInterpolate (ex12.cpp:966-1033).  Point numbering as in the
 * reference: cells [0,nC), vertices [nC,nC+nV), then the facets DMPlexInterpolate creates (edges in 2-D). ---- */
inline PetscErrorCode ohp_dm_facets(DM dm) {
  if (dm->facets) return 0;
  PetscErrorCode ierr = ohp_mesh_build_facets(dm->mesh, &dm->nF); if (ierr) return ierr;
  const PetscInt nc = dm->dim + 1, fv = dm->dim;
  dm->facet_cone.resize((size_t)dm->nF * fv); dm->facet_support.resize(dm->nF); dm->cell_cone.resize((size_t)dm->nC * nc);
  ierr = ohp_mesh_get_facets(dm->mesh, dm->facet_cone.data(), dm->facet_support.data(), dm->cell_cone.data()); if (ierr) return ierr;
  for (auto &p : dm->facet_cone) p += dm->nC;             /* vertex index -> Plex point */
  for (auto &p : dm->cell_cone) p += dm->nC + dm->nV;     /* facet index  -> Plex point */
  dm->facets = true;
  return 0;
}
inline PetscErrorCode DMPlexGetHeightStratum(DM dm, PetscInt height, PetscInt *start, PetscInt *end) {
  if (height == 0) { *start = 0; *end = dm->nC; return 0; }
  if (height == dm->dim) { *start = dm->nC; *end = dm->nC + dm->nV; return 0; }
  if (height == 1) {
    PetscErrorCode ierr = ohp_dm_facets(dm); if (ierr) return ierr;
    *start = dm->nC + dm->nV; *end = dm->nC + dm->nV + dm->nF;
    return 0;
  }
  fprintf(stderr, "[0]PETSC ERROR: DMPlexGetHeightStratum: height %d of a %d-D mesh is not built (cells, facets, vertices only)\n", (int)height, (int)dm->dim);
  return PETSC_ERR_SUP;
}
inline PetscErrorCode DMPlexGetDepthStratum(DM dm, PetscInt depth, PetscInt *start, PetscInt *end) {
  return DMPlexGetHeightStratum(dm, dm->dim - depth, start, end);
}
inline PetscErrorCode DMPlexGetSupportSize(DM dm, PetscInt p, PetscInt *size) {
  PetscErrorCode ierr = ohp_dm_facets(dm); if (ierr) return ierr;
  const PetscInt f0 = dm->nC + dm->nV;
  if (p < f0 || p >= f0 + dm->nF) { fprintf(stderr, "[0]PETSC ERROR: DMPlexGetSupportSize: point %d is not a facet\n", (int)p); return PETSC_ERR_ARG_OUTOFRANGE; }
  *size = dm->facet_support[p - f0];
  return 0;
}
inline PetscErrorCode DMPlexGetConeSize(DM dm, PetscInt p, PetscInt *size) {
  const PetscInt f0 = dm->nC + dm->nV;
  *size = p < dm->nC ? dm->dim + 1 : p >= f0 ? dm->dim : 0;
  return 0;
}
inline PetscErrorCode DMPlexGetCone(DM dm, PetscInt p, const PetscInt **cone) {
  PetscErrorCode ierr = ohp_dm_facets(dm); if (ierr) return ierr;
  const PetscInt f0 = dm->nC + dm->nV;
  if (p >= 0 && p < dm->nC) { *cone = dm->cell_cone.data() + (size_t)p * (dm->dim + 1); return 0; }
  if (p >= f0 && p < f0 + dm->nF) { *cone = dm->facet_cone.data() + (size_t)(p - f0) * dm->dim; return 0; }
  fprintf(stderr, "[0]PETSC ERROR: DMPlexGetCone: point %d has no cone\n", (int)p);
  return PETSC_ERR_ARG_OUTOFRANGE;
}
/* the reference labels the boundary vertices (and edges) one point at a time (ex12.cpp:1023-1033): vertex points are
 * collected and handed to the library at set-up; edge points carry no P1 dof and are accepted and ignored */
inline PetscErrorCode DMSetLabelValue(DM dm, const char *name, PetscInt point, PetscInt value) {
  if (strcmp(name, "marker") || value != 1) { fprintf(stderr, "[0]PETSC ERROR: DMSetLabelValue: only label \"marker\" = 1 is built\n"); return PETSC_ERR_SUP; }
  if (dm->set_up) return PETSC_ERR_ORDER;
  dm->label_by_caller = true; dm->has_marker = true;
  if (point >= dm->nC && point < dm->nC + dm->nV) dm->label_verts.push_back(point - dm->nC);
  else if (point < dm->nC) return PETSC_ERR_ARG_OUTOFRANGE;
  return 0;
}
inline PetscErrorCode DMPlexUninterpolate(DM dm, DM *udm) { dm->refs++; *udm = dm; return 0; }
inline PetscErrorCode DMPlexCopyCoordinates(DM, DM) { return 0; }
inline PetscErrorCode DMSetFromOptions(DM) { return 0; }
inline PetscErrorCode PetscSFView(PetscSF, void *) { return 0; }
/* -dm_view (CMakeLists.txt:44-62 pass it through -dm_view in the TEST block, ex12.cpp:881,962): PETSc's one-line-per-stratum summary */
inline PetscErrorCode DMViewFromOptions(DM dm, PetscObject, const char *name) {
  PetscBool set;
  PetscOptionsHasName(NULL, NULL, name, &set);
  if (!set) return 0;
  int64_t tot[3] = {dm->nV, dm->facets ? dm->nF : -1, dm->nC};
  PetscPrintf(PETSC_COMM_WORLD, "DM Object: Mesh in %d dimensions (this rank):\n  0-cells: %lld\n", (int)dm->dim, (long long)tot[0]);
  if (tot[1] >= 0) PetscPrintf(PETSC_COMM_WORLD, "  %d-cells: %lld\n", (int)dm->dim - 1, (long long)tot[1]);
  PetscPrintf(PETSC_COMM_WORLD, "  %d-cells: %lld\nLabels:\n  marker: %s\n", (int)dm->dim, (long long)tot[2], dm->has_marker ? "1 strata (boundary)" : "not set");
  return 0;
}
/* CreateBCLabel (ex12.cpp:515-526): faces with one supporting cell, completed to their vertices */
inline PetscErrorCode DMPlexMarkBoundaryFaces(DM dm, PetscInt value, DMLabel) {
  if (value != 1) return PETSC_ERR_SUP;
  dm->has_marker = true;
  return 0; /* evaluated on the device by ohp_mesh_setup, once ghosts are known */
}
inline PetscErrorCode DMPlexLabelComplete(DM, DMLabel) { return 0; }
inline PetscErrorCode DMDestroy(DM *dm) {
  if (!*dm) return 0;
  if (--(*dm)->refs == 0) { ohp_mesh_destroy((*dm)->mesh); delete *dm; }
  *dm = nullptr;
  return 0;
}
inline PetscErrorCode DMSetApplicationContext(DM dm, void *ctx) { dm->app_ctx = ctx; return 0; }
inline PetscErrorCode DMGetApplicationContext(DM dm, void *ctx) { *(void **)ctx = dm->app_ctx; return 0; }
inline PetscErrorCode DMGetDimension(DM dm, PetscInt *dim) { *dim = dm->dim; return 0; }

/* ---- discretisation ---- */
inline PetscErrorCode PetscFECreateDefault(MPI_Comm, PetscInt dim, PetscInt Nc, PetscBool simplex, const char *, PetscInt qorder, PetscFE *fe) {
  if (!simplex || Nc != 1 || (qorder != -1 && qorder != PETSC_DEFAULT)) return PETSC_ERR_SUP; /* P1 simplices, default quadrature (ex12.cpp:1309) */
  PetscInt degree = 1;
  PetscOptionsGetInt(NULL, NULL, "-petscspace_degree", &degree, NULL);
  if (degree != 1) { fprintf(stderr, "[0]PETSC ERROR: only -petscspace_degree 1 is built\n"); return PETSC_ERR_SUP; }
  *fe = new _p_PetscFE{dim, Nc, simplex};
  return 0;
}
inline PetscErrorCode PetscFEDestroy(PetscFE *fe) { delete *fe; *fe = nullptr; return 0; }
inline PetscErrorCode DMSetField(DM dm, PetscInt f, DMLabel, PetscObject) { if (f != 0) return PETSC_ERR_SUP; dm->has_field = true; return 0; }
inline PetscErrorCode DMCreateDS(DM) { return 0; }
inline PetscErrorCode DMGetDS(DM dm, PetscDS *ds) { *ds = &dm->ds; return 0; }
inline PetscErrorCode PetscDSSetResidual(PetscDS ds, PetscInt f, PetscPointFunc f0, PetscPointFunc f1) {
  if (f != 0) return PETSC_ERR_SUP;
  ds->f0 = f0; ds->f1 = f1; ds->problem = -1;
  return 0;
}
inline PetscErrorCode PetscDSSetJacobian(PetscDS ds, PetscInt f, PetscInt g, PetscPointJac g0, PetscPointJac g1, PetscPointJac g2, PetscPointJac g3) {
  if (f != 0 || g != 0 || g0 || g1 || g2) return PETSC_ERR_SUP; /* only g3 is used on this path (ex12.cpp:1182) */
  ds->g3 = g3; ds->problem = -1;
  return 0;
}
/* PetscDSSetBdResidual(prob, 0, f0_bd_u, f1_bd_zero) -- ex12.cpp:1226,1231 */
inline PetscErrorCode PetscDSSetBdResidual(PetscDS ds, PetscInt f, PetscBdPointFunc f0, PetscBdPointFunc f1) {
  if (f != 0) return PETSC_ERR_SUP;
  ds->f0_bd = f0; ds->f1_bd = f1; ds->problem = -1;
  return 0;
}
inline PetscErrorCode PetscDSSetExactSolution(PetscDS ds, PetscInt f, PetscSimplePointFunc sol, void *ctx) {
  if (f != 0) return PETSC_ERR_SUP;
  ds->exact = sol; ds->exact_ctx = ctx;
  return 0;
}
inline PetscErrorCode DMAddBoundary(DM dm, DMBoundaryConditionType type, const char *, const char *labelname, PetscInt field,
                                    PetscInt numcomps, const PetscInt *, void (*bcFunc)(void), void (*)(void), PetscInt numids,
                                    const PetscInt *ids, void *) {
  /* ex12.cpp:1237-1239: DM_BC_ESSENTIAL on label "marker" (-bc_type dirichlet) or DM_BC_NATURAL on label "boundary"
   * (-bc_type neumann: nothing is constrained, the PetscDSSetBdResidual callbacks are integrated over the boundary facets) */
  const bool essential = type == DM_BC_ESSENTIAL && !strcmp(labelname, "marker");
  const bool natural = type == DM_BC_NATURAL && !strcmp(labelname, "boundary");
  if (!(essential || natural) || field != 0 || numcomps != 0 || numids != 1 || ids[0] != 1) {
    fprintf(stderr, "[0]PETSC ERROR: only DM_BC_ESSENTIAL on label \"marker\" = 1 and DM_BC_NATURAL on label \"boundary\" = 1 are built (ex12.cpp:1237-1239)\n");
    return PETSC_ERR_SUP;
  }
  if (dm->set_up) { fprintf(stderr, "[0]PETSC ERROR: DMAddBoundary after the DM was set up (vectors or matrices exist)\n"); return OHP_ERR_ORDER; }
  dm->bc_type = natural ? OHP_BC_NATURAL : OHP_BC_ESSENTIAL;
  dm->ds.bc = (PetscSimplePointFunc)bcFunc;
  dm->ds.problem = -1;
  dm->has_bc = true;
  return 0;
}
/* resolves the registered device instantiation of the DS's callback set */
inline PetscErrorCode ohp_ds_problem(PetscDS ds, int *problem) {
  if (ds->problem < 0) {
    auto &reg = ohp::problem_registry();
    auto it = reg.find(ohp::ProblemKey{(void *)ds->f0, (void *)ds->f1, (void *)ds->g3, (void *)ds->bc});
    if (it == reg.end()) {
      fprintf(stderr, "[0]PETSC ERROR: this (f0, f1, g3, bc) callback set has no device instantiation: declare it with "
                      "OHP_DEFINE_PROBLEM(...) in the translation unit that defines the callbacks\n");
      return PETSC_ERR_SUP;
    }
    if (ds->f0_bd && ((void *)ds->f0_bd != it->second.f0_bd || (void *)ds->f1_bd != it->second.f1_bd)) {
      fprintf(stderr, "[0]PETSC ERROR: the boundary callbacks given to PetscDSSetBdResidual have no device instantiation: declare the "
                      "callback set with OHP_DEFINE_PROBLEM_BD(name, f0, f1, g3, bc, exact, f0_bd, f1_bd)\n");
      return PETSC_ERR_SUP;
    }
    ds->problem = it->second.id;
  }
  *problem = ds->problem;
  return 0;
}

/* ---- Vec / Mat ---- */
inline PetscErrorCode DMCreateGlobalVector(DM dm, Vec *v) {
  PetscErrorCode ierr = ohp_dm_setup(dm); if (ierr) return ierr;
  Vec x = new _p_Vec(); x->dm = dm;
  ierr = ohp_vec_create(dm->mesh, &x->v);
  if (ierr) { delete x; return ierr; }
  *v = x;
  return 0;
}
inline PetscErrorCode VecDuplicate(Vec v, Vec *w) {
  Vec x = new _p_Vec(); x->dm = v->dm;
  PetscErrorCode ierr = ohp_vec_duplicate(v->v, &x->v);
  if (ierr) { delete x; return ierr; }
  *w = x;
  return 0;
}
inline PetscErrorCode VecDestroy(Vec *v) { if (*v) { ohp_vec_destroy((*v)->v); delete *v; *v = nullptr; } return 0; }
inline PetscErrorCode VecSetName(Vec v, const char *name) { v->name = name; return 0; }
inline PetscErrorCode VecSet(Vec v, PetscScalar a) { return ohp_vec_set(v->v, a); }
inline PetscErrorCode VecCopy(Vec x, Vec y) { return ohp_vec_copy(x->v, y->v); }
inline PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x) { return ohp_vec_axpy(y->v, a, x->v); }
inline PetscErrorCode VecAYPX(Vec y, PetscScalar a, Vec x) { return ohp_vec_aypx(y->v, a, x->v); }
inline PetscErrorCode VecDot(Vec x, Vec y, PetscScalar *r) { return ohp_vec_dot(x->v, y->v, r); }
inline PetscErrorCode VecChop(Vec v, PetscReal tol) { return ohp_vec_chop(v->v, tol); }
inline PetscErrorCode VecNorm(Vec v, NormType t, PetscReal *r) { if (t != NORM_2) return PETSC_ERR_SUP; return ohp_vec_norm2(v->v, r); }
inline PetscErrorCode VecGetLocalSize(Vec v, PetscInt *n) { return ohp_vec_size(v->v, n); }
inline PetscErrorCode VecGetSize(Vec v, PetscInt *n) {
  ohp_mesh_info info;
  PetscErrorCode ierr = ohp_mesh_get_info(v->dm->mesh, &info);
  *n = (PetscInt)info.n_global;
  return ierr;
}
/* host copy of the local part (VecGetArrayRead + copy) */
inline PetscErrorCode VecGetValuesHost(Vec v, std::vector<PetscScalar> &out) {
  PetscInt n; ohp_vec_size(v->v, &n);
  out.resize(n);
  return ohp_vec_to_host(v->v, out.data());
}
inline PetscErrorCode DMCreateMatrix(DM dm, Mat *J) {
  PetscErrorCode ierr = ohp_dm_setup(dm); if (ierr) return ierr;
  Mat A = new _p_Mat(); A->dm = dm;
  ierr = ohp_mat_create(dm->mesh, &A->m);
  if (ierr) { delete A; return ierr; }
  *J = A;
  return 0;
}
inline PetscErrorCode MatDestroy(Mat *A) { if (*A) { ohp_mat_destroy((*A)->m); delete *A; *A = nullptr; } return 0; }
inline PetscErrorCode MatMult(Mat A, Vec x, Vec y) { return ohp_mat_mult(A->m, x->v, y->v); }
inline PetscErrorCode MatGetSize(Mat A, PetscInt *M, PetscInt *N) {
  ohp_mesh_info info;
  PetscErrorCode ierr = ohp_mesh_get_info(A->dm->mesh, &info);
  if (M) *M = (PetscInt)info.n_global;
  if (N) *N = (PetscInt)info.n_global;
  return ierr;
}
/* MatNullSpaceCreate(comm, PETSC_TRUE, 0, NULL, &nullSpace) + MatSetNullSpace(A, nullSpace) (ex12.cpp:1534-1538);
 * MatNullSpaceRemove(nullSpace, u) (ex12.cpp:1667); MatNullSpaceDestroy (ex12.cpp:1695).  Only the constant null space. */
inline PetscErrorCode MatNullSpaceCreate(MPI_Comm, PetscBool has_cnst, PetscInt n, const Vec *, MatNullSpace *sp) {
  if (n != 0) { fprintf(stderr, "[0]PETSC ERROR: MatNullSpaceCreate: only the constant null space is built (ex12.cpp:1536)\n"); return PETSC_ERR_SUP; }
  *sp = new _p_MatNullSpace{has_cnst};
  return 0;
}
inline PetscErrorCode MatNullSpaceDestroy(MatNullSpace *sp) { if (sp && *sp) { delete *sp; *sp = nullptr; } return 0; }
inline PetscErrorCode MatSetNullSpace(Mat A, MatNullSpace sp) { return ohp_mat_set_nullspace(A->m, sp && sp->has_cnst ? 1 : 0); }
inline PetscErrorCode MatNullSpaceRemove(MatNullSpace sp, Vec u) { return (sp && sp->has_cnst) ? ohp_vec_remove_constant(u->v) : 0; }

/* ---- projection / error ---- */
inline PetscErrorCode DMProjectFunction(DM dm, PetscReal, PetscSimplePointFunc *funcs, void **, InsertMode, Vec u) {
  int problem;
  PetscErrorCode ierr = ohp_dm_setup(dm); if (ierr) return ierr;
  ierr = ohp_ds_problem(&dm->ds, &problem); if (ierr) return ierr;
  if (funcs[0] == dm->ds.exact) return ohp_project(dm->mesh, problem, 1, u->v);
  auto it = ohp::function_registry().find((void *)funcs[0]);
  if (it != ohp::function_registry().end()) return ohp_project_function(dm->mesh, it->second, u->v); /* OHP_DEFINE_FUNCTION */
  fprintf(stderr, "[0]PETSC ERROR: DMProjectFunction: this function has no device instantiation: declare it with "
                  "OHP_DEFINE_FUNCTION(name, fn) next to its definition (the exact solution of the DS needs no declaration)\n");
  return PETSC_ERR_SUP;
}
inline PetscErrorCode DMComputeL2Diff(DM dm, PetscReal, PetscSimplePointFunc *, void **, Vec u, PetscReal *err) {
  int problem;
  PetscErrorCode ierr = ohp_ds_problem(&dm->ds, &problem); if (ierr) return ierr;
  return ohp_l2_error(dm->mesh, problem, u->v, err); /* the registered problem's own exact solution */
}

/* ---- viewers: `-vec_view vtk:<file>.vtu` (what every reference CTest writes, CMakeLists.txt:75,81,87,93,100,107) ----
 * ASCII VTK XML UnstructuredGrid of this rank's mesh with the nodal field: unconstrained vertices from
 * the vector, "marker" vertices from the Dirichlet function (the values PETSc inserts into the local vector). */
inline PetscErrorCode VecViewVTU(Vec u, const char *path) {
  DM dm = u->dm;
  ohp_mesh_info info;
  PetscErrorCode ierr = ohp_mesh_get_info(dm->mesh, &info); if (ierr) return ierr;
  const int dim = info.dim, nc = dim + 1;
  std::vector<int32_t> cells((size_t)info.n_cells * nc);
  std::vector<double> X((size_t)info.n_verts * dim), vals;
  std::vector<int64_t> gidx(info.n_verts);
  ierr = ohp_mesh_get_cells(dm->mesh, cells.data()); if (ierr) return ierr;
  ierr = ohp_mesh_get_coords(dm->mesh, X.data()); if (ierr) return ierr;
  ierr = ohp_mesh_get_numbering(dm->mesh, gidx.data()); if (ierr) return ierr;
  ierr = VecGetValuesHost(u, vals); if (ierr) return ierr;
  FILE *f = fopen(path, "w");
  if (!f) { fprintf(stderr, "[0]PETSC ERROR: cannot open %s\n", path); return OHP_ERR_FILE_OPEN; }
  fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n <UnstructuredGrid>\n");
  fprintf(f, "  <Piece NumberOfPoints=\"%d\" NumberOfCells=\"%d\">\n   <Points>\n    <DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n", info.n_verts, info.n_cells);
  for (int v = 0; v < info.n_verts; ++v) fprintf(f, "%.17g %.17g %.17g\n", X[(size_t)v * dim], X[(size_t)v * dim + 1], dim == 3 ? X[(size_t)v * dim + 2] : 0.0);
  fprintf(f, "    </DataArray>\n   </Points>\n   <Cells>\n    <DataArray type=\"Int32\" Name=\"connectivity\" format=\"ascii\">\n");
  for (int c = 0; c < info.n_cells; ++c) { for (int j = 0; j < nc; ++j) fprintf(f, "%d ", cells[(size_t)c * nc + j]); fprintf(f, "\n"); }
  fprintf(f, "    </DataArray>\n    <DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">\n");
  for (int c = 0; c < info.n_cells; ++c) fprintf(f, "%d\n", (c + 1) * nc);
  fprintf(f, "    </DataArray>\n    <DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n");
  for (int c = 0; c < info.n_cells; ++c) fprintf(f, "%d\n", dim == 2 ? 5 : 10); /* VTK_TRIANGLE / VTK_TETRA */
  fprintf(f, "    </DataArray>\n   </Cells>\n   <PointData Scalars=\"%s\">\n    <DataArray type=\"Float64\" Name=\"%s\" format=\"ascii\">\n",
          u->name.empty() ? "potential" : u->name.c_str(), u->name.empty() ? "potential" : u->name.c_str());
  for (int v = 0; v < info.n_verts; ++v) {
    double val = 0.0;
    const int64_t g = gidx[v] - info.row_start;
    if (gidx[v] >= 0 && g >= 0 && g < info.n_owned) val = vals[(size_t)g];
    else if (gidx[v] < 0 && dm->ds.bc) dm->ds.bc(dim, 0.0, &X[(size_t)v * dim], 1, &val, NULL);
    fprintf(f, "%.17g\n", val);
  }
  fprintf(f, "    </DataArray>\n   </PointData>\n  </Piece>\n </UnstructuredGrid>\n</VTKFile>\n");
  fclose(f);
  return 0;
}
inline PetscErrorCode VecViewFromOptions(Vec u, PetscObject, const char *name) {
  char buf[1024];
  PetscBool set;
  PetscOptionsGetString(NULL, NULL, name, buf, sizeof(buf), &set);
  if (!set) return 0;
  if (strncmp(buf, "vtk:", 4)) { fprintf(stderr, "[0]PETSC ERROR: %s %s: only vtk:<file>.vtu is built\n", name, buf); return PETSC_ERR_SUP; }
  return VecViewVTU(u, buf + 4);
}

/* ---- SNES ---- */
inline PetscErrorCode SNESCreate(MPI_Comm, SNES *snes) {
  SNES s = new _p_SNES();
  ohp_snes_default_options(&s->opts);
  memset(&s->last, 0, sizeof(s->last));
  *snes = s;
  return 0;
}
inline PetscErrorCode SNESDestroy(SNES *snes) {
  if (*snes) { delete *snes; *snes = nullptr; }
  return 0;
}
inline PetscErrorCode SNESSetDM(SNES snes, DM dm) { snes->dm = dm; return 0; }
inline PetscErrorCode SNESGetDM(SNES snes, DM *dm) { *dm = snes->dm; return 0; }
inline PetscErrorCode DMPlexSetSNESLocalFEM(DM dm, void *, void *, void *) {
  int problem;
  return ohp_ds_problem(&dm->ds, &problem); /* fail early if the callbacks have no device instantiation */
}
inline PetscErrorCode SNESSetJacobian(SNES snes, Mat A, Mat J, void *, void *) { snes->A = A; snes->J = J; return 0; }
inline PetscErrorCode SNESSetFromOptions(SNES snes) {
  char buf[64];
  PetscBool set;
  PetscOptionsGetReal(NULL, NULL, "-snes_rtol", &snes->opts.rtol, NULL);
  PetscOptionsGetReal(NULL, NULL, "-snes_atol", &snes->opts.abstol, NULL);
  PetscOptionsGetReal(NULL, NULL, "-snes_stol", &snes->opts.stol, NULL);
  PetscOptionsGetInt(NULL, NULL, "-snes_max_it", &snes->opts.max_it, NULL);
  PetscOptionsGetReal(NULL, NULL, "-ksp_rtol", &snes->opts.ksp.rtol, NULL);
  PetscOptionsGetReal(NULL, NULL, "-ksp_atol", &snes->opts.ksp.abstol, NULL);
  PetscOptionsGetReal(NULL, NULL, "-ksp_divtol", &snes->opts.ksp.dtol, NULL);
  PetscOptionsGetInt(NULL, NULL, "-ksp_max_it", &snes->opts.ksp.max_it, NULL);
  PetscOptionsGetString(NULL, NULL, "-ksp_type", buf, sizeof(buf), &set);
  snes->opts.ksp.ksp_type = OHP_KSP_CG;
  if (set && !strcmp(buf, "pipecg")) snes->opts.ksp.ksp_type = OHP_KSP_PIPECG; /* KSPPIPECG: reductions overlap the SpMV */
  else if (set && strcmp(buf, "cg")) { fprintf(stderr, "[0]PETSC ERROR: -ksp_type %s: only cg and pipecg are built (north_star)\n", buf); return PETSC_ERR_SUP; }
  snes->opts.ksp.pc = OHP_PC_NONE;
  PetscOptionsGetString(NULL, NULL, "-pc_type", buf, sizeof(buf), &set);
  if (set) {
    if (!strcmp(buf, "none")) snes->opts.ksp.pc = OHP_PC_NONE;
    else if (!strcmp(buf, "jacobi")) snes->opts.ksp.pc = OHP_PC_JACOBI;
    else if (!strcmp(buf, "ksp") || !strcmp(buf, "chebyshev")) {
      /* -pc_type ksp -ksp_ksp_type chebyshev -ksp_pc_type jacobi -ksp_ksp_max_it k [-ksp_ksp_chebyshev_eigenvalues emin,emax]:
       * PETSc's way of using a Chebyshev/Jacobi polynomial as the preconditioner of the outer CG (PCKSP) */
      char inner[64] = "chebyshev", ipc[64] = "jacobi";
      PetscBool iset;
      PetscOptionsGetString(NULL, NULL, "-ksp_ksp_type", inner, sizeof(inner), &iset);
      PetscOptionsGetString(NULL, NULL, "-ksp_pc_type", ipc, sizeof(ipc), &iset);
      if (strcmp(inner, "chebyshev") || strcmp(ipc, "jacobi")) {
        fprintf(stderr, "[0]PETSC ERROR: -pc_type ksp: only -ksp_ksp_type chebyshev with -ksp_pc_type jacobi is built\n");
        return PETSC_ERR_SUP;
      }
      snes->opts.ksp.pc = OHP_PC_CHEBYSHEV;
      PetscInt k = 0;
      PetscOptionsGetInt(NULL, NULL, "-ksp_ksp_max_it", &k, NULL);
      snes->opts.ksp.cheb_its = k;
      char eb[128];
      PetscOptionsGetString(NULL, NULL, "-ksp_ksp_chebyshev_eigenvalues", eb, sizeof(eb), &iset);
      if (iset) {
        double lo = 0.0, hi = 0.0;
        if (sscanf(eb, "%lf,%lf", &lo, &hi) != 2 || !(hi > lo) || !(lo > 0.0)) {
          fprintf(stderr, "[0]PETSC ERROR: -ksp_ksp_chebyshev_eigenvalues %s: expected emin,emax with 0 < emin < emax\n", eb);
          return PETSC_ERR_ARG_WRONG;
        }
        snes->opts.ksp.cheb_emin = lo; snes->opts.ksp.cheb_emax = hi;
      }
    } else { fprintf(stderr, "[0]PETSC ERROR: -pc_type %s: none, jacobi and ksp (Chebyshev/Jacobi polynomial) are built\n", buf); return PETSC_ERR_SUP; }
  }
  PetscBool b;
  PetscOptionsHasName(NULL, NULL, "-snes_monitor_short", &b); snes->monitor = b;
  PetscOptionsHasName(NULL, NULL, "-snes_monitor", &b); snes->monitor = snes->monitor || b;
  PetscOptionsHasName(NULL, NULL, "-ksp_monitor_short", &b); snes->ksp_monitor = b;
  PetscOptionsHasName(NULL, NULL, "-ksp_monitor", &b); snes->ksp_monitor = snes->ksp_monitor || b;
  PetscOptionsHasName(NULL, NULL, "-snes_converged_reason", &b); snes->converged_reason = b;
  PetscOptionsHasName(NULL, NULL, "-ksp_converged_reason", &b); snes->ksp_converged_reason = b;
  return 0;
}
inline PetscErrorCode SNESComputeFunction(SNES snes, Vec u, Vec F) {
  int problem;
  PetscErrorCode ierr = ohp_ds_problem(&snes->dm->ds, &problem); if (ierr) return ierr;
  return ohp_residual(snes->dm->mesh, problem, u->v, F->v);
}
inline PetscErrorCode SNESComputeJacobian(SNES snes, Vec u, Mat A, Mat) {
  int problem;
  PetscErrorCode ierr = ohp_ds_problem(&snes->dm->ds, &problem); if (ierr) return ierr;
  return ohp_jacobian(snes->dm->mesh, problem, u->v, A->m);
}
inline const char *ohp_snes_reason_name(int r) {
  switch (r) {
    case 2: return "CONVERGED_FNORM_ABS"; case 3: return "CONVERGED_FNORM_RELATIVE"; case 4: return "CONVERGED_SNORM_RELATIVE";
    case -3: return "DIVERGED_LINEAR_SOLVE"; case -5: return "DIVERGED_MAX_IT"; default: return "UNKNOWN";
  }
}
inline const char *ohp_ksp_reason_name(int r) {
  switch (r) {
    case 2: return "CONVERGED_RTOL"; case 3: return "CONVERGED_ATOL"; case -3: return "DIVERGED_ITS"; case -4: return "DIVERGED_DTOL";
    case -8: return "DIVERGED_INDEFINITE_MAT"; case -9: return "DIVERGED_NANORINF"; default: return "UNKNOWN";
  }
}
/* -snes_monitor_short / -ksp_monitor_short / -ksp_converged_reason (CMakeLists.txt:56-59), in PETSc's wording and
 * indentation (the KSP is nested one level inside the SNES) */
inline void ohp_snes_monitor_print(void *ctx, int32_t it, double fnorm) {
  SNES snes = (SNES)ctx;
  if (snes->monitor) PetscPrintf(PETSC_COMM_WORLD, "  %d SNES Function norm %g \n", (int)it, fnorm);
}
inline void ohp_ksp_monitor_print(void *ctx, int32_t, int32_t n, const double *rnorm, int32_t reason) {
  SNES snes = (SNES)ctx;
  if (snes->ksp_monitor && rnorm)
    for (int k = 0; k <= n; ++k) PetscPrintf(PETSC_COMM_WORLD, "    %d KSP Residual norm %g \n", k, rnorm[k] < 1e-11 ? 0.0 : rnorm[k]);
  if (snes->ksp_converged_reason)
    PetscPrintf(PETSC_COMM_WORLD, "    Linear solve %s due to %s iterations %d\n", reason > 0 ? "converged" : "did not converge", ohp_ksp_reason_name(reason), (int)n);
}
inline PetscErrorCode SNESSolve(SNES snes, Vec b, Vec u) {
  if (b) return PETSC_ERR_SUP;
  if (!snes->J) return PETSC_ERR_ORDER;
  int problem;
  PetscErrorCode ierr = ohp_ds_problem(&snes->dm->ds, &problem); if (ierr) return ierr;
  snes->opts.monitor_ctx = snes;
  snes->opts.snes_monitor = snes->monitor ? ohp_snes_monitor_print : nullptr;
  snes->opts.ksp_monitor = (snes->ksp_monitor || snes->ksp_converged_reason) ? ohp_ksp_monitor_print : nullptr;
  if (snes->ksp_converged_reason && !snes->ksp_monitor) { snes->opts.ksp.history = nullptr; snes->opts.ksp.history_cap = 0; }
  ierr = ohp_snes_solve(snes->dm->mesh, problem, snes->J->m, u->v, &snes->opts, &snes->last); if (ierr) return ierr;
  if (snes->converged_reason)
    PetscPrintf(PETSC_COMM_WORLD, "Nonlinear solve %s due to %s iterations %d\n", snes->last.reason > 0 ? "converged" : "did not converge",
                ohp_snes_reason_name(snes->last.reason), snes->last.its);
  snes->sol = u; /* borrowed reference, as SNESGetSolution returns (ex12.cpp:1610) */
  return 0;
}
inline PetscErrorCode SNESGetSolution(SNES snes, Vec *u) { *u = snes->sol; return 0; }
inline PetscErrorCode SNESGetIterationNumber(SNES snes, PetscInt *its) { *its = snes->last.its; return 0; }
inline PetscErrorCode SNESGetLinearSolveIterations(SNES snes, PetscInt *its) { *its = snes->last.ksp_its; return 0; }
inline PetscErrorCode SNESGetConvergedReason(SNES snes, int *reason) { *reason = snes->last.reason; return 0; }

#endif /* OHP_PETSC_H */
