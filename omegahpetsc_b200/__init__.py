"""omegahpetsc_b200 -- ctypes binding of libohp.so, the B200-native P1 Poisson hot path.

This module is a thin handle layer over the C-ABI declared in include/ohp.h (the product is
the CUDA library; Python is only the harness language of tests/ and bench.py).  Method names
follow the PETSc calls of the reference's main() (ex12.cpp:1491-1618) that each entry point
replaces.  There is NO CPU fallback: every compute call raises OhpError when the library or
a CUDA device is missing.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OHP_LIB") or os.path.join(_HERE, "lib", "libohp.so")  # OHP_LIB: A/B builds of the same ABI

PROBLEM_COEFF_NONE, PROBLEM_ANALYTIC, PROBLEM_NONLINEAR, PROBLEM_CIRCLE, PROBLEM_CROSS = 0, 1, 2, 3, 4
PC_NONE, PC_JACOBI, PC_CHEBYSHEV = 0, 1, 2
BC_ESSENTIAL, BC_NATURAL, BC_NONE = 0, 1, 2
ALGO_GENERAL_PCG = 6
KSP_CG, KSP_PIPECG = 0, 1

_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_vp = ctypes.c_void_p


class OhpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ohp error %d: %s" % (code, msg))
        self.code = code


class CoreCounts(ctypes.Structure):
    _fields_ = [("n_core_elems", ctypes.c_int32), ("n_core_verts", ctypes.c_int32),
                ("n_owned_verts", ctypes.c_int32), ("n_ghost_verts", ctypes.c_int32)]


class MeshInfo(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int32), ("n_cells", ctypes.c_int32), ("n_verts", ctypes.c_int32),
                ("n_boundary_verts", ctypes.c_int32), ("n_owned", ctypes.c_int32), ("n_ghost", ctypes.c_int32),
                ("n_global", ctypes.c_int64), ("row_start", ctypes.c_int64), ("nnz", ctypes.c_int64),
                ("max_row", ctypes.c_int32)]


class KspOptions(ctypes.Structure):
    _fields_ = [("rtol", ctypes.c_double), ("abstol", ctypes.c_double), ("dtol", ctypes.c_double),
                ("max_it", ctypes.c_int32), ("pc", ctypes.c_int32), ("check_every", ctypes.c_int32),
                ("history_cap", ctypes.c_int32), ("history", _f64p), ("profile", ctypes.c_int32),
                ("ksp_type", ctypes.c_int32), ("cheb_its", ctypes.c_int32), ("reserved0", ctypes.c_int32),
                ("cheb_emin", ctypes.c_double), ("cheb_emax", ctypes.c_double)]


class KspResult(ctypes.Structure):
    _fields_ = [("its", ctypes.c_int32), ("reason", ctypes.c_int32), ("rnorm", ctypes.c_double),
                ("rnorm0", ctypes.c_double), ("spmv_ms", ctypes.c_float), ("spmv_samples", ctypes.c_int32),
                ("total_ms", ctypes.c_float), ("algo", ctypes.c_int32), ("cheb_emin", ctypes.c_double),
                ("cheb_emax", ctypes.c_double), ("spmvs", ctypes.c_int32), ("reserved0", ctypes.c_int32)]


SNES_MONITOR_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int32, ctypes.c_double)
KSP_MONITOR_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, _f64p, ctypes.c_int32)


class SnesOptions(ctypes.Structure):
    _fields_ = [("rtol", ctypes.c_double), ("abstol", ctypes.c_double), ("stol", ctypes.c_double),
                ("max_it", ctypes.c_int32), ("ksp", KspOptions),
                ("snes_monitor", SNES_MONITOR_FN), ("ksp_monitor", KSP_MONITOR_FN), ("monitor_ctx", ctypes.c_void_p)]


class SnesResult(ctypes.Structure):
    _fields_ = [("its", ctypes.c_int32), ("ksp_its", ctypes.c_int32), ("reason", ctypes.c_int32),
                ("n_residual", ctypes.c_int32), ("n_jacobian", ctypes.c_int32), ("fnorm0", ctypes.c_double),
                ("fnorm", ctypes.c_double), ("ms_residual", ctypes.c_float), ("ms_jacobian", ctypes.c_float),
                ("ms_ksp", ctypes.c_float), ("spmv_ms", ctypes.c_float), ("spmv_samples", ctypes.c_int32),
                ("algo", ctypes.c_int32)]


_lib = None


def build(verbose=False):
    """Compile libohp.so in-tree (nvcc, sm_100a only)."""
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"] + ([] if verbose else ["-s"]), check=True)


def lib():
    """The loaded C-ABI library.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OhpError(76, "libohp.so is not built (%s); run __graft_entry__.build() -- "
                               "this package has no CPU or PyTorch fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.ohp_last_error.restype = ctypes.c_char_p
        L.ohp_vec_device_ptr.restype = ctypes.c_void_p
        L.ohp_vec_device_ptr.argtypes = [_vp]
        for name in ("ohp_vec_set", "ohp_vec_chop"):
            getattr(L, name).argtypes = [_vp, ctypes.c_double]
        for name in ("ohp_vec_axpy", "ohp_vec_aypx"):
            getattr(L, name).argtypes = [_vp, ctypes.c_double, _vp]
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise OhpError(rc, lib().ohp_last_error().decode(errors="replace"))


def _ptr(a, t):
    return a.ctypes.data_as(t)


class Context:
    """PetscInitialize + Omega_h::Library (ex12.cpp:1485,1494): one GPU, one stream."""

    def __init__(self, device=0):
        self.h = _vp()
        _chk(lib().ohp_init(int(device), ctypes.byref(self.h)))

    def close(self):
        if self.h:
            lib().ohp_finalize(self.h)
            self.h = _vp()

    def sync(self):
        _chk(lib().ohp_sync(self.h))

    def device_name(self):
        buf = ctypes.create_string_buffer(256)
        _chk(lib().ohp_device_name(self.h, buf, 256))
        return buf.value.decode()

    def launch_count(self):
        n = ctypes.c_int64()
        _chk(lib().ohp_launch_count(self.h, ctypes.byref(n)))
        return n.value

    def stream(self):
        s = _vp()
        _chk(lib().ohp_stream(self.h, ctypes.byref(s)))
        return s.value

    # communicator ---------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = (ctypes.c_uint8 * 128)()
        _chk(lib().ohp_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, nranks, rank, unique_id):
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        _chk(lib().ohp_comm_init(self.h, int(nranks), int(rank), buf))

    def comm_init_from_env(self):
        """PetscInitialize for a multi-rank launch without MPI: rank/size from the launcher's environment, the NCCL id
        from rank 0 over TCP next to MASTER_PORT (ohp_boot_env + ohp_boot_bcast + ohp_comm_init)."""
        _chk(lib().ohp_comm_init_from_env(self.h))

    def barrier(self):
        _chk(lib().ohp_comm_barrier(self.h))

    def allgather_i64(self, mine):
        n = self.rank_size()[1]
        out = np.zeros(n, dtype=np.int64)
        _chk(lib().ohp_comm_allgather_i64(self.h, ctypes.c_int64(int(mine)), _ptr(out, _i64p)))
        return out

    def allreduce_sum_i64(self, vals):
        v = np.ascontiguousarray(vals, dtype=np.int64).copy()
        _chk(lib().ohp_comm_allreduce_sum_i64(self.h, _ptr(v, _i64p), v.size))
        return v

    def rank_size(self):
        r, n = ctypes.c_int(), ctypes.c_int()
        _chk(lib().ohp_comm_rank_size(self.h, ctypes.byref(r), ctypes.byref(n)))
        return r.value, n.value

    def ghost_roots_from_gids(self, ghost_gid, ghost_owner, core_gid, core_owned=None):
        """getPicPartCoreVtxOwnerIdx's gid round trip (ex12.cpp:731-777) on the device: index of every ghost in
        its owner's vertex array.  Collective."""
        gg = np.ascontiguousarray(ghost_gid, dtype=np.int64)
        go = np.ascontiguousarray(ghost_owner, dtype=np.int32)
        cg = np.ascontiguousarray(core_gid, dtype=np.int64)
        co = None if core_owned is None else np.ascontiguousarray(core_owned, dtype=np.uint8)
        roots = np.empty(max(gg.size, 1), dtype=np.int32)
        _chk(lib().ohp_ghost_roots_from_gids(self.h, gg.size, _ptr(gg, _i64p), _ptr(go, _i32p), cg.size, _ptr(cg, _i64p),
                                             _ptr(co, _u8p) if co is not None else None, _ptr(roots, _i32p)))
        return roots[:gg.size]

    def picpart_extract(self, dim, rank, elem_verts, coords, elem_owner, vert_owner, vert_gid=None, with_overlap=False):
        """getPicPartCoreVtxCoords / CoreElmToVtxArray / ghost list (ex12.cpp:584-730) on the device."""
        ev = np.ascontiguousarray(elem_verts, dtype=np.int32)
        X = np.ascontiguousarray(coords, dtype=np.float64)
        eo = np.ascontiguousarray(elem_owner, dtype=np.int32)
        vo = np.ascontiguousarray(vert_owner, dtype=np.int32)
        nE, nc = ev.shape
        nV = X.shape[0]
        gid = None if vert_gid is None else np.ascontiguousarray(vert_gid, dtype=np.int64)
        cnt = CoreCounts()
        cev = np.empty((max(nE, 1), nc), dtype=np.int32)
        cX = np.empty((max(nV, 1), dim), dtype=np.float64)
        p2c = np.empty(max(nV, 1), dtype=np.int32)
        gci = np.empty(max(nV, 1), dtype=np.int32)
        ggid = np.empty(max(nV, 1), dtype=np.int64)
        gown = np.empty(max(nV, 1), dtype=np.int32)
        cgid = np.empty(max(nV, 1), dtype=np.int64)
        _chk(lib().ohp_picpart_extract(self.h, int(dim), int(rank), int(bool(with_overlap)), nE, nV, _ptr(ev, _i32p),
                                       _ptr(X, _f64p), _ptr(eo, _i32p), _ptr(vo, _i32p),
                                       _ptr(gid, _i64p) if gid is not None else None, ctypes.byref(cnt),
                                       _ptr(cev, _i32p), _ptr(cX, _f64p), _ptr(p2c, _i32p), _ptr(gci, _i32p),
                                       _ptr(ggid, _i64p), _ptr(gown, _i32p), _ptr(cgid, _i64p)))
        return dict(n_core_elems=cnt.n_core_elems, n_core_verts=cnt.n_core_verts, n_owned_verts=cnt.n_owned_verts,
                    n_ghost_verts=cnt.n_ghost_verts, core_elem_verts=cev[:cnt.n_core_elems].copy(),
                    core_coords=cX[:cnt.n_core_verts].copy(), part2core=p2c[:nV].copy(),
                    ghost_core_idx=gci[:cnt.n_ghost_verts].copy(), ghost_gid=ggid[:cnt.n_ghost_verts].copy(),
                    ghost_owner=gown[:cnt.n_ghost_verts].copy(), core_gid=cgid[:cnt.n_core_verts].copy())


def boot_env():
    """(rank, nranks, local_rank) from the launcher's environment (ohp_boot_env); needs no GPU."""
    r, n, l = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _chk(lib().ohp_boot_env(ctypes.byref(r), ctypes.byref(n), ctypes.byref(l)))
    return r.value, n.value, l.value


def boot_bcast(rank, nranks, payload, nbytes):
    """Rank 0's `payload` (bytes) to every rank over TCP (ohp_boot_bcast); needs no GPU."""
    buf = (ctypes.c_uint8 * nbytes)()
    if rank == 0:
        buf[:] = payload[:nbytes]
    _chk(lib().ohp_boot_bcast(int(rank), int(nranks), buf, int(nbytes)))
    return bytes(buf)


class Osh:
    """Omega_h::binary::read (ex12.cpp:813,818).  Host only."""
    _DT = {0: np.int8, 2: np.int32, 3: np.int64, 5: np.float64}

    def __init__(self, path, part=0):
        self.h = _vp()
        _chk(lib().ohp_osh_read(os.fsencode(path), int(part), ctypes.byref(self.h)))
        d, nv, ne, ver, npart = (ctypes.c_int() for _ in range(5))
        _chk(lib().ohp_osh_info(self.h, ctypes.byref(d), ctypes.byref(nv), ctypes.byref(ne), ctypes.byref(ver),
                                ctypes.byref(npart)))
        self.dim, self.nverts, self.nelems, self.version, self.nparts = d.value, nv.value, ne.value, ver.value, npart.value

    def close(self):
        if self.h:
            lib().ohp_osh_free(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def nents(self, d):
        n = ctypes.c_int()
        _chk(lib().ohp_osh_nents(self.h, int(d), ctypes.byref(n)))
        return n.value

    def elem_verts(self):
        """Mesh::ask_elem_verts (ex12.cpp:591,634,784)."""
        p = _i32p()
        _chk(lib().ohp_osh_elem_verts(self.h, ctypes.byref(p)))
        n = self.nelems * (self.dim + 1)
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(self.nelems, self.dim + 1).copy() if n else np.zeros((0, self.dim + 1), np.int32)

    def coords(self):
        p = _f64p()
        _chk(lib().ohp_osh_coords(self.h, ctypes.byref(p)))
        n = self.nverts * self.dim
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(self.nverts, self.dim).copy()

    def tag(self, ent_dim, name):
        """mesh.get_array<T>(dim, name) (ex12.cpp:593-594,627-628,703-705)."""
        nc, ty, cnt, data = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64(), _vp()
        _chk(lib().ohp_osh_tag(self.h, int(ent_dim), name.encode(), ctypes.byref(nc), ctypes.byref(ty),
                               ctypes.byref(cnt), ctypes.byref(data)))
        dt = np.dtype(self._DT[ty.value])
        if cnt.value == 0:
            return np.zeros(0, dt)
        buf = (ctypes.c_uint8 * (cnt.value * dt.itemsize)).from_address(data.value)
        a = np.frombuffer(buf, dtype=dt).copy()
        return a.reshape(-1, nc.value) if nc.value > 1 else a

    def has_tag(self, ent_dim, name):
        rc = lib().ohp_osh_tag(self.h, int(ent_dim), name.encode(), None, None, None, None)
        return rc == 0


class Vec:
    def __init__(self, mesh, handle=None):
        self.mesh = mesh
        self.h = _vp()
        if handle is None:
            _chk(lib().ohp_vec_create(mesh.h, ctypes.byref(self.h)))
        else:
            self.h = handle

    @property
    def n(self):
        k = ctypes.c_int32()
        _chk(lib().ohp_vec_size(self.h, ctypes.byref(k)))
        return k.value

    def destroy(self):
        if self.h:
            lib().ohp_vec_destroy(self.h)
            self.h = _vp()

    def duplicate(self):
        return Vec(self.mesh)

    def set(self, a):
        _chk(lib().ohp_vec_set(self.h, float(a)))
        return self

    def from_host(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.size != self.n:
            raise OhpError(60, "Vec.from_host: size %d, expected %d" % (a.size, self.n))
        _chk(lib().ohp_vec_from_host(self.h, _ptr(a, _f64p)))
        return self

    def to_host(self):
        a = np.empty(max(self.n, 1), dtype=np.float64)
        _chk(lib().ohp_vec_to_host(self.h, _ptr(a, _f64p)))
        return a[:self.n]

    def copy_from(self, src):
        _chk(lib().ohp_vec_copy(src.h, self.h))
        return self

    def axpy(self, alpha, x):
        _chk(lib().ohp_vec_axpy(self.h, float(alpha), x.h))
        return self

    def aypx(self, alpha, x):
        _chk(lib().ohp_vec_aypx(self.h, float(alpha), x.h))
        return self

    def dot(self, y):
        r = ctypes.c_double()
        _chk(lib().ohp_vec_dot(self.h, y.h, ctypes.byref(r)))
        return r.value

    def norm(self):
        r = ctypes.c_double()
        _chk(lib().ohp_vec_norm2(self.h, ctypes.byref(r)))
        return r.value

    def chop(self, tol):
        _chk(lib().ohp_vec_chop(self.h, float(tol)))
        return self

    def device_ptr(self):
        return lib().ohp_vec_device_ptr(self.h)

    def remove_constant(self):
        """MatNullSpaceRemove for the constant null space (ex12.cpp:1667): subtract the global mean."""
        _chk(lib().ohp_vec_remove_constant(self.h))
        return self


class Mat:
    def __init__(self, mesh):
        self.mesh = mesh
        self.h = _vp()
        _chk(lib().ohp_mat_create(mesh.h, ctypes.byref(self.h)))

    def destroy(self):
        if self.h:
            lib().ohp_mat_destroy(self.h)
            self.h = _vp()

    def zero(self):
        _chk(lib().ohp_mat_zero(self.h))

    def values(self):
        v = np.empty(max(self.mesh.info().nnz, 1), dtype=np.float64)
        _chk(lib().ohp_mat_get_values(self.h, _ptr(v, _f64p)))
        return v[:self.mesh.info().nnz]

    def mult(self, x, y):
        _chk(lib().ohp_mat_mult(self.h, x.h, y.h))
        return y

    def set_nullspace(self, has_constant=True):
        """MatNullSpaceCreate(comm, PETSC_TRUE, 0, NULL) + MatSetNullSpace (ex12.cpp:1534-1538)."""
        _chk(lib().ohp_mat_set_nullspace(self.h, int(bool(has_constant))))
        return self


def ksp_options(rtol=1e-5, abstol=1e-50, dtol=1e5, max_it=10000, pc=PC_NONE, check_every=0, profile=0, ksp_type=KSP_CG,
                cheb_its=0, cheb_emin=0.0, cheb_emax=0.0):
    o = KspOptions()
    _chk(lib().ohp_ksp_default_options(ctypes.byref(o)))
    o.rtol, o.abstol, o.dtol, o.max_it, o.pc, o.check_every = rtol, abstol, dtol, int(max_it), int(pc), int(check_every)
    o.profile = int(profile)
    o.ksp_type = int(ksp_type)
    o.cheb_its, o.cheb_emin, o.cheb_emax = int(cheb_its), float(cheb_emin), float(cheb_emax)
    return o


def ksp_cg(A, b, x, opts=None, history=False):
    """KSPSolve with -ksp_type cg (ex12.cpp:1543)."""
    o = opts if opts is not None else ksp_options()
    hist = None
    if history:
        hist = np.zeros(o.max_it + 2)
        o.history_cap = hist.size
        o.history = _ptr(hist, _f64p)
    res = KspResult()
    try:
        _chk(lib().ohp_ksp_cg(A.h, b.h, x.h, ctypes.byref(o), ctypes.byref(res)))
    finally:
        o.history_cap = 0
        o.history = None
    out = dict(its=res.its, reason=res.reason, rnorm=res.rnorm, rnorm0=res.rnorm0, spmv_ms=res.spmv_ms,
               spmv_samples=res.spmv_samples, total_ms=res.total_ms, algo=res.algo, cheb_emin=res.cheb_emin,
               cheb_emax=res.cheb_emax, spmvs=res.spmvs)
    if history:
        out["hist"] = hist[:res.its + 1].copy()
    return out


class Mesh:
    """DM.  DMPlexCreateFromCellListPetsc (ex12.cpp:859,933) + everything ex12 does up to
    DMCreateMatrix (ex12.cpp:958-1033,1309-1321,1505-1508), on the device."""

    def __init__(self, ctx, cells, coords):
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        if cells.ndim != 2 or coords.ndim != 2:
            raise OhpError(62, "Mesh: cells must be (nC, dim+1), coords (nV, dim)")
        self.ctx = ctx
        self.dim = coords.shape[1]
        if cells.shape[1] != self.dim + 1:
            raise OhpError(60, "Mesh: %d corners for dim %d" % (cells.shape[1], self.dim))
        self.h = _vp()
        self._nC = cells.shape[0]
        _chk(lib().ohp_mesh_create(ctx.h, self.dim, cells.shape[0], coords.shape[0], _ptr(cells, _i32p),
                                   _ptr(coords, _f64p), ctypes.byref(self.h)))
        self._info = None

    @classmethod
    def from_osh(cls, ctx, path, part=0):
        """-mesh <dir> (ex12.cpp:816-820): Omega_h::binary::read + ask_elem_verts + coords."""
        o = Osh(path, part)
        try:
            return cls(ctx, o.elem_verts(), o.coords())
        finally:
            o.close()

    def destroy(self):
        if self.h:
            lib().ohp_mesh_destroy(self.h)
            self.h = _vp()

    def set_ghosts(self, ghost_vertex, owner_rank, owner_vertex):
        """PetscSFSetGraph on the point SF (ex12.cpp:873-876,937-939)."""
        gv = np.ascontiguousarray(ghost_vertex, dtype=np.int32)
        orank = np.ascontiguousarray(owner_rank, dtype=np.int32)
        overt = np.ascontiguousarray(owner_vertex, dtype=np.int32)
        _chk(lib().ohp_mesh_set_ghosts(self.h, gv.size, _ptr(gv, _i32p), _ptr(orank, _i32p), _ptr(overt, _i32p)))

    def set_ghosts_from_gids(self, ghost_vertex, owner_rank, vertex_gid, vertex_owned):
        """Point SF from global vertex ids: the roots are resolved on the device (ohp_ghost_roots_from_gids)."""
        gv = np.ascontiguousarray(ghost_vertex, dtype=np.int32)
        gid = np.ascontiguousarray(vertex_gid, dtype=np.int64)
        roots = self.ctx.ghost_roots_from_gids(gid[gv], owner_rank, gid, vertex_owned)
        self.set_ghosts(gv, owner_rank, roots)
        return roots

    def set_point_sf(self, leaf_point, root_rank, root_point):
        """The same graph in Plex POINT numbers, exactly as ex12 passes it (ex12.cpp:867-871,924-926)."""
        lp = np.ascontiguousarray(leaf_point, dtype=np.int32)
        rr = np.ascontiguousarray(root_rank, dtype=np.int32)
        rp = np.ascontiguousarray(root_point, dtype=np.int32)
        _chk(lib().ohp_mesh_set_point_sf(self.h, lp.size, _ptr(lp, _i32p), _ptr(rr, _i32p), _ptr(rp, _i32p)))

    def set_bc_type(self, bc_type):
        """DMAddBoundary's type (ex12.cpp:1237): BC_ESSENTIAL (default), BC_NATURAL (-bc_type neumann), BC_NONE."""
        _chk(lib().ohp_mesh_set_bc_type(self.h, int(bc_type)))
        return self

    def setup(self):
        _chk(lib().ohp_mesh_setup(self.h))
        self._info = None
        return self

    def p2p_enable_native(self):
        """ohp_mesh_p2p_enable: the IPC handles travel over the library's own communicator."""
        _chk(lib().ohp_mesh_p2p_enable(self.h))

    def facets(self):
        """What DMPlexInterpolate adds between cells and vertices (edges in 2-D, faces in 3-D): (facet_verts [nF, dim],
        support [nF], cell_facets [nC, dim+1]) as 0-based indices (DMPlexGetCone / GetSupportSize, ex12.cpp:990,1006)."""
        n = ctypes.c_int32()
        _chk(lib().ohp_mesh_build_facets(self.h, ctypes.byref(n)))
        nF, nC = n.value, self._nC
        fv = np.empty((max(nF, 1), self.dim), dtype=np.int32)
        sup = np.empty(max(nF, 1), dtype=np.int32)
        cf = np.empty((max(nC, 1), self.dim + 1), dtype=np.int32)
        _chk(lib().ohp_mesh_get_facets(self.h, _ptr(fv, _i32p), _ptr(sup, _i32p), _ptr(cf, _i32p)))
        return fv[:nF], sup[:nF], cf[:nC]

    def set_label(self, vertices):
        """DMSetLabelValue(dm, "marker", nC + v, 1) for every v (ex12.cpp:1023-1026)."""
        v = np.ascontiguousarray(vertices, dtype=np.int32)
        _chk(lib().ohp_mesh_set_label(self.h, v.size, _ptr(v, _i32p)))

    def p2p_enable(self, all_gather_object, world):
        """ohp_mesh_p2p_export + all-gather of the CUDA IPC handles + ohp_mesh_p2p_attach."""
        buf = (ctypes.c_uint8 * 64)()
        _chk(lib().ohp_mesh_p2p_export(self.h, buf))
        handles = [None] * world
        all_gather_object(handles, bytes(buf))
        blob = (ctypes.c_uint8 * (64 * world)).from_buffer_copy(b"".join(handles))
        _chk(lib().ohp_mesh_p2p_attach(self.h, blob))

    def info(self):
        if self._info is None:
            i = MeshInfo()
            _chk(lib().ohp_mesh_get_info(self.h, ctypes.byref(i)))
            self._info = i
        return self._info

    def boundary(self):
        b = np.empty(max(self.info().n_verts, 1), dtype=np.uint8)
        _chk(lib().ohp_mesh_get_boundary(self.h, _ptr(b, _u8p)))
        return b[:self.info().n_verts]

    def numbering(self):
        g = np.empty(max(self.info().n_verts, 1), dtype=np.int64)
        _chk(lib().ohp_mesh_get_numbering(self.h, _ptr(g, _i64p)))
        return g[:self.info().n_verts]

    def csr(self):
        i = self.info()
        rowptr = np.empty(i.n_owned + 1, dtype=np.int32)
        colidx = np.empty(max(i.nnz, 1), dtype=np.int64)
        _chk(lib().ohp_mesh_get_csr(self.h, _ptr(rowptr, _i32p), _ptr(colidx, _i64p)))
        return rowptr, colidx[:i.nnz]

    # DM-level factories ----------------------------------------------------------------
    def create_global_vector(self):
        return Vec(self)

    def create_matrix(self):
        return Mat(self)

    # FEM evaluators (DMPlexSetSNESLocalFEM, ex12.cpp:1540) -------------------------------
    def residual(self, problem, u, F):
        _chk(lib().ohp_residual(self.h, int(problem), u.h, F.h))
        return F

    def jacobian(self, problem, u, J):
        _chk(lib().ohp_jacobian(self.h, int(problem), u.h, J.h))
        return J

    def project(self, problem, which, u):
        _chk(lib().ohp_project(self.h, int(problem), int(which), u.h))
        return u

    def l2_error(self, problem, u):
        e = ctypes.c_double()
        _chk(lib().ohp_l2_error(self.h, int(problem), u.h, ctypes.byref(e)))
        return e.value

    def snes_solve(self, problem, J, u, snes_rtol=1e-8, snes_atol=1e-50, snes_stol=1e-8, snes_max_it=50, ksp=None, monitor=None):
        """SNESSolve (ex12.cpp:1609): newtonls + bt over CG.  `monitor` (a dict) receives the -snes_monitor norms under
        "snes" and one array of -ksp_monitor norms per Newton iteration under "ksp"."""
        o = SnesOptions()
        _chk(lib().ohp_snes_default_options(ctypes.byref(o)))
        o.rtol, o.abstol, o.stol, o.max_it = snes_rtol, snes_atol, snes_stol, int(snes_max_it)
        if ksp is not None:
            o.ksp = ksp
        keep = []
        if monitor is not None:
            monitor.setdefault("snes", []); monitor.setdefault("ksp", []); monitor.setdefault("ksp_reason", [])
            sm = SNES_MONITOR_FN(lambda ctx, it, fnorm: monitor["snes"].append((it, fnorm)))
            km = KSP_MONITOR_FN(lambda ctx, nit, n, rn, reason: (monitor["ksp"].append(np.array([rn[k] for k in range(n + 1)])),
                                                                  monitor["ksp_reason"].append(reason)) and None)
            keep = [sm, km]
            o.snes_monitor, o.ksp_monitor = sm, km
        r = SnesResult()
        _chk(lib().ohp_snes_solve(self.h, int(problem), J.h, u.h, ctypes.byref(o), ctypes.byref(r)))
        del keep
        return {f[0]: getattr(r, f[0]) for f in SnesResult._fields_}
