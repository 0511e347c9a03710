/*
 * ohp.h -- C-ABI of the B200-native P1 Poisson hot path (omegahPetsc ex12 pipeline).
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes / opaque handles and
 * returns an int error code (0 = success, PetscErrorCode convention, ex12.cpp:1491-1501
 * `ierr = f();CHKERRQ(ierr);`).  ohp_last_error() gives the message of the last failure on
 * the calling thread.  Nothing here is a torch type; data handed in is HOST memory unless the
 * parameter name ends in `_dev`.
 *
 * Each function cites the reference interface (paths relative to the reference checkout)
 * that it replaces.  "PETSc-internal"/"Omega_h-internal" = behaviour of the un-vendored
 * library reached from the cited call site.
 *
 * Types: PetscInt = int32 (ex12.cpp:915,939 pass int* as PetscInt*), PetscScalar =
 * PetscReal = double, Omega_h::LO = int32, GO = int64.
 */
#ifndef OHP_H
#define OHP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OHP_VERSION 1

typedef struct ohp_context_s *ohp_context; /* device, stream, optional NCCL communicator   */
typedef struct ohp_osh_s *ohp_osh;         /* one decoded `.osh` stream, host memory        */
typedef struct ohp_mesh_s *ohp_mesh;       /* DM: Plex-shaped mesh + section + CSR pattern  */
typedef struct ohp_vec_s *ohp_vec;         /* Vec (device resident, FP64)                    */
typedef struct ohp_mat_s *ohp_mat;         /* Mat (AIJ / CSR values on the DM's pattern)     */

/* error codes (subset of PETSc's petscerror.h numbering) */
enum {
  OHP_OK = 0,
  OHP_ERR_MEM = 55,          /* PETSC_ERR_MEM */
  OHP_ERR_SUP = 56,          /* PETSC_ERR_SUP */
  OHP_ERR_ARG_SIZ = 60,      /* PETSC_ERR_ARG_SIZ */
  OHP_ERR_ARG_WRONG = 62,    /* PETSC_ERR_ARG_WRONG */
  OHP_ERR_ARG_OUTOFRANGE = 63,
  OHP_ERR_FILE_OPEN = 65,
  OHP_ERR_FILE_READ = 66,
  OHP_ERR_FILE_UNEXPECTED = 79,
  OHP_ERR_LIB = 76,          /* PETSC_ERR_LIB: CUDA / NCCL / zlib failure */
  OHP_ERR_ORDER = 58,        /* PETSC_ERR_ORDER: object used before set up */
  OHP_ERR_GPU = 97           /* PETSC_ERR_GPU: no usable device (there is NO CPU fallback) */
};

const char *ohp_last_error(void);
int ohp_version(void);

/* ------------------------------------------------------------------------------------ */
/* Context.  Replaces PetscInitialize / Omega_h::Library / MPI_Comm plumbing              */
/* (ex12.cpp:1485,1494).  One context per process = one GPU = one rank.                  */
/* ------------------------------------------------------------------------------------ */
int ohp_init(int device, ohp_context *ctx);
int ohp_finalize(ohp_context ctx);
int ohp_device_name(ohp_context ctx, char *buf, int buflen);
int ohp_sync(ohp_context ctx);
/* cudaStream_t the context launches on, as a void* (for CUDA-event timing by the caller) */
int ohp_stream(ohp_context ctx, void **stream);
/* cumulative number of kernel launches issued by this library on this context */
int ohp_launch_count(ohp_context ctx, int64_t *count);

/* Multi-GPU: MPI_Comm_rank/size (ex12.cpp:800-801) + the communicator PETSc/Omega_h use
 * for halo exchange and dot products.  `unique_id` is the 128-byte ncclUniqueId made by
 * ohp_comm_unique_id() on rank 0 and broadcast by the host launcher. */
int ohp_comm_unique_id(void *unique_id128);
int ohp_comm_init(ohp_context ctx, int nranks, int rank, const void *unique_id128);
int ohp_comm_rank_size(ohp_context ctx, int *rank, int *nranks);

/* MPI-free start-up (this image has no MPI; PetscInitialize -> MPI_Init, ex12.cpp:1485).  ohp_boot_env reads rank,
 * size and local rank from the launcher's environment (RANK / WORLD_SIZE / LOCAL_RANK of torch.distributed.run;
 * OMPI_*, PMI_*, SLURM_* as fall-backs).  ohp_boot_bcast sends `bytes` bytes from rank 0 to every rank over TCP
 * at MASTER_ADDR, port OHP_BOOT_PORT (default MASTER_PORT + 307): the one exchange needed before NCCL exists.
 * ohp_comm_init_from_env = both, applied to the ncclUniqueId, then ohp_comm_init. */
int ohp_boot_env(int *rank, int *nranks, int *local_rank);
int ohp_boot_bcast(int rank, int nranks, void *buf, int bytes);
int ohp_comm_init_from_env(ohp_context ctx);
/* MPI_Barrier (ex12.cpp:573,575,814,953,978), MPI_Allgather of one count per rank (the element counts that
 * getNeighborElmCounts exchanges, ex12.cpp:528-577), MPI_Allreduce(SUM) of integers (ex12.cpp:844) */
int ohp_comm_barrier(ohp_context ctx);
int ohp_comm_allgather_i64(ohp_context ctx, int64_t mine, int64_t *all /* nranks */);
int ohp_comm_allreduce_sum_i64(ohp_context ctx, int64_t *vals, int n);

/* ------------------------------------------------------------------------------------ */
/* `.osh` loading -- Omega_h::binary::read (ex12.cpp:813,818) + Mesh::ask_elem_verts     */
/* (ex12.cpp:591,634,784) + mesh.coords() / get_array<>(dim,name) (ex12.cpp:593-594,     */
/* 627-628,703-705).  Host only: needs no GPU.                                           */
/* ------------------------------------------------------------------------------------ */
int ohp_osh_read(const char *dir, int part, ohp_osh *osh);
int ohp_osh_free(ohp_osh osh);
int ohp_osh_info(ohp_osh osh, int *dim, int *nverts, int *nelems, int *version, int *nparts);
/* element->vertex connectivity, derived from the stored downward adjacencies + codes;
 * borrowed pointer, (dim+1)*nelems int32, valid until ohp_osh_free */
int ohp_osh_elem_verts(ohp_osh osh, const int32_t **ev);
int ohp_osh_coords(ohp_osh osh, const double **coords); /* dim*nverts doubles, borrowed */
/* tag lookup: type code 0=i8 2=i32 3=i64 5=f64 (Omega_h_defines.h); returns
 * OHP_ERR_ARG_WRONG if the tag is absent */
int ohp_osh_tag(ohp_osh osh, int ent_dim, const char *name, int *ncomps, int *type,
                int64_t *count, const void **data);
int ohp_osh_nents(ohp_osh osh, int ent_dim, int *n);

/* ------------------------------------------------------------------------------------ */
/* picpart core extraction -- getPicPartCoreVtxCoords (ex12.cpp:623-672, kernels K2,K3), */
/* getPicPartCoreElmToVtxArray (ex12.cpp:584-619, K1), ghost list of                     */
/* getPicPartCoreVtxOwnerIdx (ex12.cpp:694-730, K4,K5).  Runs on the device (mask,       */
/* scan, compaction); results returned in caller-provided host arrays sized by the       */
/* part's counts.  `with_overlap` = 0 reproduces the reference's core (owned elements);  */
/* 1 adds every element touching an owned vertex (the assembly closure this library      */
/* needs for communication-free row assembly).                            */
/* ------------------------------------------------------------------------------------ */
typedef struct {
  int32_t n_core_elems;   /* elements kept                                     */
  int32_t n_core_verts;   /* vertices of kept elements                         */
  int32_t n_owned_verts;  /* core vertices with owner == rank                  */
  int32_t n_ghost_verts;  /* core vertices with owner != rank                  */
} ohp_core_counts;

int ohp_picpart_extract(ohp_context ctx, int dim, int rank, int with_overlap, int32_t n_elems,
                        int32_t n_verts, const int32_t *elem_verts, const double *coords,
                        const int32_t *elem_owner, const int32_t *vert_owner,
                        const int64_t *vert_gid, ohp_core_counts *counts,
                        int32_t *core_elem_verts /* (dim+1)*n_elems cap */,
                        double *core_coords /* dim*n_verts cap */,
                        int32_t *part2core /* n_verts, -1 if not core */,
                        int32_t *ghost_core_idx /* n_verts cap */,
                        int64_t *ghost_gid /* n_verts cap */,
                        int32_t *ghost_owner /* n_verts cap */,
                        int64_t *core_gid /* n_verts cap */);

/* The rest of getPicPartCoreVtxOwnerIdx (ex12.cpp:731-777): for every ghost vertex, the index it has in
 * its OWNER's vertex array.  Restates the gid round trip of the reference -- ghosts' gids to the owners
 * (Dist::exch, :740), findIdxOfGid on the owner (K6, :748-761, here a hash join on the device instead of the
 * O(n_verts x n_ghosts) search), indices back (distInv.exch, :765) -- over the NCCL communicator.
 * `core_gid[n_core]` are the global ids of THIS rank's vertex array (the one given to ohp_mesh_create),
 * `core_owned[n_core]` flags the vertices this rank owns (NULL = all of them are entered in the join table).  roots[i] = index of ghost i in rank ghost_owner[i]'s array; add the owner's cell count
 * to get the Plex point ex12 stores (:757) or hand the indices to ohp_mesh_set_ghosts as they are.
 * Collective; every rank returns the same error code when any rank finds an inconsistency. */
int ohp_ghost_roots_from_gids(ohp_context ctx, int32_t n_ghost, const int64_t *ghost_gid, const int32_t *ghost_owner,
                              int32_t n_core, const int64_t *core_gid, const uint8_t *core_owned, int32_t *roots);

/* ------------------------------------------------------------------------------------ */
/* DM.  ohp_mesh_create = DMPlexCreateFromCellListPetsc(comm,dim,nC,nV,nCorners,interp,  */
/* cells,spaceDim,coords,&dm) (ex12.cpp:859,933): cells and coords are COPIED (to HBM).  */
/* Points follow Plex numbering: cells [0,nC), vertices [nC,nC+nV).                      */
/* ------------------------------------------------------------------------------------ */
int ohp_mesh_create(ohp_context ctx, int dim, int32_t n_cells, int32_t n_verts,
                    const int32_t *cells, const double *coords, ohp_mesh *mesh);
int ohp_mesh_destroy(ohp_mesh mesh); /* collective once ohp_mesh_p2p_attach was called (DMDestroy of a parallel DM) */

/* PetscSFSetGraph on the DM's point SF (ex12.cpp:873-876, 937-939): leaf i = local vertex
 * ghost_vertex[i] (0-based VERTEX index, i.e. Plex point minus nC), root = vertex
 * owner_vertex[i] on rank owner_rank[i].  Arrays are copied (the caller keeps ownership,
 * unlike PETSC_OWN_POINTER). Collective when a communicator is attached. */
int ohp_mesh_set_ghosts(ohp_mesh mesh, int32_t n_leaves, const int32_t *ghost_vertex,
                        const int32_t *owner_rank, const int32_t *owner_vertex);

/* The same graph given exactly as ex12 gives it to PetscSFSetGraph: Plex POINT numbers.  Leaf =
 * n_cells + local vertex (ex12.cpp:868,924); root = owner's n_cells + owner's vertex index
 * (ex12.cpp:870,926, which the reference obtains with getNeighborElmCounts, ex12.cpp:528-577).
 * Collective: all-gathers the ranks' cell counts to undo the offsets. */
int ohp_mesh_set_point_sf(ohp_mesh mesh, int32_t n_leaves, const int32_t *leaf_point,
                          const int32_t *root_rank, const int32_t *root_point);

/* Everything ex12 does between mesh creation and the first solve, on the device:
 *   DMPlexInterpolate + boundary marking, facets with support size 1 -> label "marker"
 *     (ex12.cpp:958-1033),
 *   PetscFECreateDefault(dim,1,simplex,-1) / DMCreateDS / DMAddBoundary(ESSENTIAL,"marker",1)
 *     -> local + global section (ex12.cpp:1309-1321,1237-1239),
 *   DMCreateMatrix preallocation -> CSR pattern (ex12.cpp:1508),
 *   and the row->element gather map used by the assembly kernels.
 * Collective when a communicator is attached. */
int ohp_mesh_setup(ohp_mesh mesh);

/* DMAddBoundary's type (ex12.cpp:1237): OHP_BC_ESSENTIAL (-bc_type dirichlet, default) constrains the dofs of the
 * "marker" vertices; OHP_BC_NATURAL (-bc_type neumann) constrains nothing -- every owned vertex gets a row -- and
 * ohp_residual adds the boundary integral of the problem's boundary callbacks (PetscDSSetBdResidual, ex12.cpp:1226,1231)
 * over the facets with one supporting cell (label "boundary", DMPlexMarkBoundaryFaces ex12.cpp:1088-1094); OHP_BC_NONE
 * (-bc_type none) constrains nothing and integrates nothing.  Call before ohp_mesh_setup. */
enum { OHP_BC_ESSENTIAL = 0, OHP_BC_NATURAL = 1, OHP_BC_NONE = 2 };
int ohp_mesh_set_bc_type(ohp_mesh mesh, int bc_type);

/* Plex queries for callers that walk the interpolated mesh themselves, as CreateQuadMesh does between
 * DMPlexInterpolate and the solve (ex12.cpp:958-1033).  Facets are the points DMPlexInterpolate creates between
 * cells and vertices: edges in 2-D (Plex points [nC+nV, nC+nV+nF)), triangles in 3-D.  Built on the device on first
 * use (support counted inside vertex stars), numbered by creating cell and local facet like PETSc's interpolation.
 *   facet_verts[(dim)*nF]      cone of each facet as 0-based vertex indices      (DMPlexGetCone,        ex12.cpp:1006)
 *   support[nF]                number of cells on each facet                      (DMPlexGetSupportSize, ex12.cpp:990)
 *   cell_facets[(dim+1)*nC]    cone of each cell as 0-based facet indices         (may be NULL)
 * ohp_mesh_set_label = the caller's DMSetLabelValue(dm,"marker",nC+v,1) loop (ex12.cpp:1023-1026): when given,
 * the library uses this label for the vertices the rank owns instead of detecting the boundary itself (ghost
 * vertices still take the label from their owner, so an incomplete local star cannot mislabel them). */
int ohp_mesh_build_facets(ohp_mesh mesh, int32_t *n_facets);
int ohp_mesh_get_facets(ohp_mesh mesh, int32_t *facet_verts, int32_t *support, int32_t *cell_facets);
int ohp_mesh_set_label(ohp_mesh mesh, int32_t n, const int32_t *vertices);

typedef struct {
  int32_t dim, n_cells, n_verts;
  int32_t n_boundary_verts; /* vertices labelled "marker"=1                              */
  int32_t n_owned;          /* owned unconstrained dofs = local rows (DMCreateGlobalVector) */
  int32_t n_ghost;          /* unconstrained ghost dofs = off-process columns              */
  int64_t n_global;         /* global unconstrained dofs                                   */
  int64_t row_start;        /* first global row of this rank                               */
  int64_t nnz;              /* local CSR entries (diag + off-diag block)                   */
  int32_t max_row;          /* longest row                                                 */
} ohp_mesh_info;
int ohp_mesh_get_info(ohp_mesh mesh, ohp_mesh_info *info);

/* NVLink peer-to-peer fast path for the partitioned CG solve (optional; without it halos and dot
 * products go through NCCL).  After ohp_mesh_setup every rank exports a 64-byte CUDA IPC handle of
 * its exchange heap, the launcher all-gathers the handles (nranks x 64 bytes, rank order) and every
 * rank attaches.  The CG kernels then push halo values straight into the neighbours' vectors and
 * all-reduce their dot products through peer mailboxes, inside the kernels that produce them.
 * Plays the role PETSc's PetscSF/VecScatter + MPI_Allreduce play under SNESSolve (ex12.cpp:1609). */
int ohp_mesh_p2p_export(ohp_mesh mesh, void *handle64);
int ohp_mesh_p2p_attach(ohp_mesh mesh, const void *handles /* nranks x 64 bytes */);
/* the same in one collective call: the handles are all-gathered over the communicator itself */
int ohp_mesh_p2p_enable(ohp_mesh mesh);

/* the DM's copy of the cell list ((dim+1)*n_cells int32) and coordinates (dim*n_verts doubles), for viewers */
int ohp_mesh_get_cells(ohp_mesh mesh, int32_t *cells);
int ohp_mesh_get_coords(ohp_mesh mesh, double *coords);
/* label "marker" (ex12.cpp:1023-1026): bnd[v] in {0,1}, n_verts bytes */
int ohp_mesh_get_boundary(ohp_mesh mesh, uint8_t *bnd);
/* global section: gidx[v] = global dof of vertex v, -1 if constrained; n_verts int64 */
int ohp_mesh_get_numbering(ohp_mesh mesh, int64_t *gidx);
/* CSR pattern: rowptr[n_owned+1]; colidx[nnz] in GLOBAL dof numbers.  Serial: ascending per row (PETSc's AIJ
 * order).  Partitioned: per row the owned (diagonal-block) columns ascending, then the ghost (off-diagonal-block)
 * columns ascending, the MPIAIJ layout. */
int ohp_mesh_get_csr(ohp_mesh mesh, int32_t *rowptr, int64_t *colidx);

/* ------------------------------------------------------------------------------------ */
/* Problems = the PetscDS registration of SetupProblem (ex12.cpp:1162-1243).             */
/* Built-in ids restate ex12's callback sets with the callbacks inlined into the kernels: */
/*   OHP_PROBLEM_COEFF_NONE : f0_u, f1_u, g3_uu            (ex12.cpp:147,198,209)          */
/*   OHP_PROBLEM_ANALYTIC   : f0/f1/g3_analytic_u          (ex12.cpp:283,292,313)          */
/*   OHP_PROBLEM_NONLINEAR  : f0/f1/g3_analytic_nonlinear_u (ex12.cpp:341,350,369)         */
/*   OHP_PROBLEM_CIRCLE/CROSS : f0_circle_u / f0_cross_u, -bc_type dirichlet with           */
/*                              -variable_coefficient circle|cross (ex12.cpp:155-177,1185-1204) */
/* Dirichlet / exact data: quadratic_u_2d / quadratic_u_3d (ex12.cpp:113,408); circle_u_2d /  */
/* cross_u_2d for the last two (ex12.cpp:127-145,1222-1225).                                */
/* User callbacks compiled by nvcc are registered through ohp_problem_register (see      */
/* include/ohp_petsc.h, which turns PetscDSSetResidual/Jacobian into this call).         */
/* ------------------------------------------------------------------------------------ */
enum { OHP_PROBLEM_COEFF_NONE = 0, OHP_PROBLEM_ANALYTIC = 1, OHP_PROBLEM_NONLINEAR = 2,
       OHP_PROBLEM_CIRCLE = 3,   /* f0_circle_u (ex12.cpp:155-166) with f1_u, g3_uu; Dirichlet/exact circle_u_2d (:127-136) */
       OHP_PROBLEM_CROSS = 4 };  /* f0_cross_u (ex12.cpp:168-177) with f1_u, g3_uu; Dirichlet/exact cross_u_2d (:138-145) */

/* device view handed to registered assembly launchers (all pointers are DEVICE pointers) */
typedef struct {
  int32_t dim, n_cells, n_verts, n_rows;
  const int32_t *cells_dev;    /* (dim+1)*n_cells                                       */
  const double *coords_dev;    /* dim*n_verts                                           */
  const int32_t *lidx_dev;     /* n_verts: local dof (owned rows first, then ghosts), -1 constrained */
  const int32_t *row_vertex_dev; /* n_rows: vertex of each local row                    */
  const int32_t *v2c_ptr_dev;  /* n_verts+1: star(v) as a CSR over vertices              */
  const int32_t *v2c_dev;      /* entries cell*4 + (position of v in the cell), ascending */
  const uint8_t *slot_dev;     /* (dim+1) per v2c entry: column slot within the row, 255 = constrained */
  const int32_t *rowptr_dev;   /* n_rows+1                                              */
  const double *u_dev;         /* local solution: n_rows owned + ghost dofs             */
  double *out_dev;             /* residual: F[n_rows]; Jacobian: vals[nnz]; projection: u[n_rows]; L2 diff: partial sums */
  void *stream;                /* cudaStream_t                                          */
  const uint8_t *owned_dev;    /* n_verts: 1 if this rank owns the vertex               */
  int32_t n_partials;          /* L2 diff: out_dev receives this many per-block partial sums of the squared error */
  /* cell-parallel assembly ("patches", csrc/ohp_patch.cu): blocks of patch_rows rows that are close in space, with the
   * cells of their stars and the vertices of those cells listed once in patch-local 16-bit ids.  n_patches == 0: the
   * row kernels walk the vertex stars instead. */
  int32_t n_patches, patch_rows, patch_max_cells, patch_max_verts, max_row;
  const int32_t *prow_dev;      /* n_patches*patch_rows: row of each patch slot (-1 = padding)                       */
  const int32_t *pstar_ptr_dev; /* n_patches*patch_rows + 1: (row, cell) incidences of each slot, ascending cell order */
  const uint16_t *pstar_dev;    /* per incidence: patch-local cell << 2 | corner                                     */
  const uint8_t *pslot_dev;     /* (dim+1) per incidence: column slot of the cell's corners in the row, 255 = none   */
  const int32_t *pcell_ptr_dev; /* n_patches + 1                                                                      */
  const uint16_t *pconn_dev;    /* (dim+1) per patch cell: patch-local vertex ids                                    */
  const int32_t *pvert_ptr_dev; /* n_patches + 1                                                                      */
  const int32_t *pvert_dev;     /* global (rank-local) vertex of each patch vertex                                   */
  const uint8_t *bnd_dev;       /* n_verts: label "marker" (1 on vertices of facets with one supporting cell)        */
} ohp_assembly_view;

typedef int (*ohp_assembly_launcher)(const ohp_assembly_view *view);
/* `project_exact` and `l2_diff` may be NULL (the problem then has no exact solution to project / compare with) */
int ohp_problem_register(ohp_assembly_launcher residual, ohp_assembly_launcher jacobian,
                         ohp_assembly_launcher project_exact, ohp_assembly_launcher l2_diff, int *problem_id);
/* PetscDSSetBdResidual(prob, 0, f0_bd, f1_bd) (ex12.cpp:1226,1231): the launcher ADDS the boundary integral into
 * out_dev[n_rows] (ohp::launch_bd_residual_any<P>); used by ohp_residual on OHP_BC_NATURAL meshes.  The built-in problems
 * carry ex12's f0_bd_u / f1_bd_zero (ex12.cpp:179-195). */
int ohp_problem_set_bd_residual(int problem_id, ohp_assembly_launcher bd_residual);
/* DMProjectFunction of an arbitrary point function compiled for the device (ohp::launch_project_fn<F>) */
int ohp_project_function(ohp_mesh mesh, ohp_assembly_launcher project, ohp_vec u);

/* ------------------------------------------------------------------------------------ */
/* Vec (VecCreate via DMCreateGlobalVector ex12.cpp:1505; VecSet/AXPY/Norm/Chop/Duplicate */
/* ex12.cpp:1626-1660).  Length = n_owned (+ ghost tail kept internally).                */
/* ------------------------------------------------------------------------------------ */
int ohp_vec_create(ohp_mesh mesh, ohp_vec *v);
int ohp_vec_duplicate(ohp_vec v, ohp_vec *w);
int ohp_vec_destroy(ohp_vec v);
int ohp_vec_size(ohp_vec v, int32_t *n_local);
int ohp_vec_set(ohp_vec v, double alpha);
int ohp_vec_copy(ohp_vec src, ohp_vec dst);
int ohp_vec_from_host(ohp_vec v, const double *host);
int ohp_vec_to_host(ohp_vec v, double *host);
int ohp_vec_axpy(ohp_vec y, double alpha, ohp_vec x);              /* y += alpha x   */
int ohp_vec_aypx(ohp_vec y, double alpha, ohp_vec x);              /* y = x + alpha y */
int ohp_vec_dot(ohp_vec x, ohp_vec y, double *result);             /* global         */
int ohp_vec_norm2(ohp_vec x, double *result);                      /* NORM_2, global */
int ohp_vec_chop(ohp_vec v, double tol);                           /* VecChop        */
void *ohp_vec_device_ptr(ohp_vec v);

/* ------------------------------------------------------------------------------------ */
/* Mat (DMCreateMatrix ex12.cpp:1508; MatMult ex12.cpp:1655)                             */
/* ------------------------------------------------------------------------------------ */
int ohp_mat_create(ohp_mesh mesh, ohp_mat *A);
int ohp_mat_destroy(ohp_mat A);
int ohp_mat_zero(ohp_mat A);
int ohp_mat_get_values(ohp_mat A, double *vals /* nnz, pattern order */);
int ohp_mat_mult(ohp_mat A, ohp_vec x, ohp_vec y); /* y = A x (halo exchange inside) */
/* MatNullSpaceCreate(comm, PETSC_TRUE, 0, NULL) + MatSetNullSpace(A, nullSpace) (ex12.cpp:1534-1538, done when -bc_type is
 * not dirichlet): the operator annihilates the constants.  ohp_ksp_cg then projects them out of every preconditioned
 * residual [PETSc-internal KSP_PCApply -> KSP_RemoveNullSpace].  ohp_vec_remove_constant = MatNullSpaceRemove(nullSpace, u)
 * (ex12.cpp:1667): u -= mean(u) over all ranks. */
int ohp_mat_set_nullspace(ohp_mat A, int has_constant);
int ohp_vec_remove_constant(ohp_vec v);

/* ------------------------------------------------------------------------------------ */
/* FEM evaluators installed by DMPlexSetSNESLocalFEM (ex12.cpp:1540) and called through   */
/* SNESComputeFunction / SNESComputeJacobian (ex12.cpp:1624,1641,1651).                  */
/* ------------------------------------------------------------------------------------ */
int ohp_residual(ohp_mesh mesh, int problem, ohp_vec u, ohp_vec F);
int ohp_jacobian(ohp_mesh mesh, int problem, ohp_vec u, ohp_mat J);
/* DMProjectFunction (ex12.cpp:1546,1602): which = 0 zero (ex12.cpp:76), 1 the problem's
 * exact solution sampled at the vertices (P1 dual basis = point evaluation) */
int ohp_project(ohp_mesh mesh, int problem, int which, ohp_vec u);
/* DMComputeL2Diff against the exact solution (ex12.cpp:1611,1637) */
int ohp_l2_error(ohp_mesh mesh, int problem, ohp_vec u, double *err);

/* ------------------------------------------------------------------------------------ */
/* KSP CG (PETSc-internal KSPSolve_CG selected by -ksp_type cg at ex12.cpp:1543).        */
/* ------------------------------------------------------------------------------------ */
/* OHP_PC_CHEBYSHEV ("preconditioning beyond Jacobi"): the Chebyshev polynomial in D^-1 A that PETSc
 * builds with -pc_type ksp -ksp_ksp_type chebyshev -ksp_pc_type jacobi -ksp_ksp_max_it <cheb_its> [PETSc-internal
 * KSPSolve_Chebyshev inside PCApply_KSP], applied without any reduction; the outer CG then runs in the general
 * host-driven loop (csrc/ohp_pcg.cu) instead of the fused recurrences (so does every solve on a matrix with a null
 * space; `ksp_type` is not consulted there: the loop is plain preconditioned CG). */
enum { OHP_PC_NONE = 0, OHP_PC_JACOBI = 1, OHP_PC_CHEBYSHEV = 2 };
/* -ksp_type: cg (the library picks the recurrence -- classic, single-reduction or pipelined -- from the global problem
 * size and the transport; all give CG's iterates in exact arithmetic) or pipecg (PETSc's KSPPIPECG, Ghysels-Vanroose:
 * forces the pipelined recurrence, whose reductions overlap the SpMV) */
enum { OHP_KSP_CG = 0, OHP_KSP_PIPECG = 1 };
typedef struct {
  double rtol, abstol, dtol; /* -ksp_rtol (1e-5; CTests 1e-9 CMakeLists.txt:53) -ksp_atol 1e-50 -ksp_divtol 1e5 */
  int32_t max_it;            /* -ksp_max_it 10000 */
  int32_t pc;                /* -pc_type none|jacobi */
  int32_t check_every;       /* iterations between host polls of the device-side converged flag (0 = default) */
  int32_t history_cap;       /* length of `history` (0 = none) */
  double *history;           /* host array receiving the monitored norm per iteration (-ksp_monitor) */
  int32_t profile;           /* >0: bracket up to this many SpMV launches with CUDA events (-log_view analogue) */
  int32_t ksp_type;          /* OHP_KSP_CG (default) | OHP_KSP_PIPECG */
  int32_t cheb_its;          /* OHP_PC_CHEBYSHEV: Chebyshev iterations per application, -ksp_ksp_max_it (0 = 8) */
  int32_t reserved0;
  double cheb_emin, cheb_emax; /* -ksp_ksp_chebyshev_eigenvalues emin,emax.  emax <= 0: 1.1 x lambda_max(D^-1 A) estimated with
                                * 10 Lanczos steps on a reproducible noise vector (KSPChebyshevEstEigSet's role); emin <= 0:
                                * emax / (4 cheb_its^2 + 36) (see csrc/ohp_pcg.hpp; PETSc's smoother default is 0.1 x estimate) */
} ohp_ksp_options;
/* the kernel the profiled launches belong to: k_spmv_tma (classic three-launch CG, single-reduction two-launch CG, two-launch
 * pipelined CG), the persistent cooperative kernel, or k_pipecg_fused (SpMV + the whole vector update in one launch) */
enum { OHP_ALGO_CLASSIC = 0, OHP_ALGO_SINGLE_REDUCTION = 1, OHP_ALGO_PERSISTENT = 2, OHP_ALGO_PIPELINED_2L = 3, OHP_ALGO_PIPELINED_FUSED = 4, OHP_ALGO_FUSED_PERSISTENT = 5,
       OHP_ALGO_GENERAL_PCG = 6 /* host-driven loop of csrc/ohp_pcg.cu: Chebyshev preconditioner and / or null space */ };
typedef struct {
  int32_t its;
  int32_t reason; /* KSPConvergedReason: 2 RTOL, 3 ATOL, -3 ITS, -4 DTOL, -8 INDEFINITE_MAT, -9 NANORINF */
  double rnorm;
  double rnorm0;
  float spmv_ms;        /* mean device time of the profiled SpMV launches (0 if profile == 0) */
  int32_t spmv_samples; /* how many launches that mean is over */
  float total_ms;       /* device time of the whole solve (CUDA events) */
  int32_t algo;         /* which recurrence ran (OHP_ALGO_*): names the kernel that `spmv_ms` times */
  double cheb_emin, cheb_emax; /* OHP_PC_CHEBYSHEV: the interval the polynomial was built for (0 otherwise) */
  int32_t spmvs;        /* OHP_ALGO_GENERAL_PCG: operator applications, eigenvalue estimate included (0 otherwise) */
  int32_t reserved0;
} ohp_ksp_result;
int ohp_ksp_default_options(ohp_ksp_options *o);
int ohp_ksp_cg(ohp_mat A, ohp_vec b, ohp_vec x, const ohp_ksp_options *opts, ohp_ksp_result *res);

/* ------------------------------------------------------------------------------------ */
/* SNES newtonls + bt (PETSc-internal SNESSolve_NEWTONLS reached from SNESSolve          */
/* ex12.cpp:1609)                                                                        */
/* ------------------------------------------------------------------------------------ */
/* -snes_monitor / -ksp_monitor (CMakeLists.txt:56-57): called on the host, after Newton iteration `it` produced
 * ||F|| = fnorm, and after the linear solve of Newton iteration `it` with the n+1 monitored norms of its CG run */
typedef void (*ohp_snes_monitor_fn)(void *ctx, int32_t it, double fnorm);
typedef void (*ohp_ksp_monitor_fn)(void *ctx, int32_t newton_it, int32_t n, const double *rnorm, int32_t reason);
typedef struct {
  double rtol, abstol, stol; /* -snes_rtol 1e-8 -snes_atol 1e-50 -snes_stol 1e-8 */
  int32_t max_it;            /* -snes_max_it 50 */
  ohp_ksp_options ksp;
  ohp_snes_monitor_fn snes_monitor; /* may be NULL */
  ohp_ksp_monitor_fn ksp_monitor;   /* may be NULL; when set the CG history is recorded on the device and copied back per solve */
  void *monitor_ctx;
} ohp_snes_options;
typedef struct {
  int32_t its;          /* Newton iterations */
  int32_t ksp_its;      /* total linear iterations */
  int32_t reason;       /* SNESConvergedReason: 2 FNORM_ABS 3 FNORM_RELATIVE 4 SNORM_RELATIVE -3 LINEAR_SOLVE -5 MAX_IT */
  int32_t n_residual;   /* residual evaluations */
  int32_t n_jacobian;   /* Jacobian evaluations */
  double fnorm0, fnorm; /* ||F(u0)||, ||F(u_final)|| */
  float ms_residual, ms_jacobian, ms_ksp; /* device time per phase (CUDA events), summed */
  float spmv_ms;        /* from the last linear solve, see ohp_ksp_result */
  int32_t spmv_samples;
  int32_t algo;         /* of the last linear solve */
} ohp_snes_result;
int ohp_snes_default_options(ohp_snes_options *o);
int ohp_snes_solve(ohp_mesh mesh, int problem, ohp_mat J, ohp_vec u, const ohp_snes_options *opts,
                   ohp_snes_result *res);

#ifdef __cplusplus
}
#endif
#endif /* OHP_H */
