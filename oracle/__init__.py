"""ORACLE package (test infrastructure, not product). See oracle/oracle.c and
oracle/osh_oracle.py for the restated algorithms and their reference citations.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this package.  PARITY UNPINNED (no golden vectors in the reference).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_u8p = ctypes.POINTER(ctypes.c_uint8)


def build():
    """Compile oracle.c (serial + OpenMP) into oracle/_build/."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib(omp=False):
    name = "liboracle_omp.so" if omp else "liboracle.so"
    if name not in _LIBS:
        path = os.path.join(_HERE, "_build", name)
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.orc_pattern.restype = ctypes.c_int64
        _LIBS[name] = L
    return _LIBS[name]


def _p(a, t):
    return a.ctypes.data_as(t)


def quadrature(dim):
    nq = ctypes.c_int()
    xi = np.zeros(8 * 3)
    w = np.zeros(8)
    assert lib().orc_quadrature(dim, ctypes.byref(nq), _p(xi, c_f64p), _p(w, c_f64p)) == 0
    return xi[: nq.value * dim].reshape(nq.value, dim).copy(), w[: nq.value].copy()


def mark_boundary(cells, nV, omp=False):
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    nC, nc = cells.shape
    bnd = np.zeros(nV, dtype=np.uint8)
    lib(omp).orc_mark_boundary(nc - 1, ctypes.c_int64(nC), nV, _p(cells, c_i32p), _p(bnd, c_u8p))
    return bnd


def numbering(bnd):
    gidx = np.zeros(bnd.size, dtype=np.int32)
    N = lib().orc_numbering(bnd.size, _p(np.ascontiguousarray(bnd, dtype=np.uint8), c_u8p), _p(gidx, c_i32p))
    return gidx, N


def pattern(cells, gidx, N, omp=False):
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    nC, nc = cells.shape
    rowptr = np.zeros(N + 1, dtype=np.int32)
    L = lib(omp)
    nnz = L.orc_pattern(nc - 1, ctypes.c_int64(nC), gidx.size, _p(cells, c_i32p), _p(gidx, c_i32p), N,
                        _p(rowptr, c_i32p), None)
    colidx = np.zeros(nnz, dtype=np.int32)
    L.orc_pattern(nc - 1, ctypes.c_int64(nC), gidx.size, _p(cells, c_i32p), _p(gidx, c_i32p), N,
                  _p(rowptr, c_i32p), _p(colidx, c_i32p))
    return rowptr, colidx


def residual(problem, cells, coords, gidx, N, u, omp=False):
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    nC, nc = cells.shape
    F = np.zeros(N)
    u = np.ascontiguousarray(u, dtype=np.float64)
    rc = lib(omp).orc_residual(problem, nc - 1, ctypes.c_int64(nC), gidx.size, _p(cells, c_i32p),
                               _p(coords, c_f64p), _p(gidx, c_i32p), N, _p(u, c_f64p), _p(F, c_f64p))
    assert rc == 0
    return F


def jacobian(problem, cells, coords, gidx, N, u, rowptr, colidx, omp=False):
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    nC, nc = cells.shape
    vals = np.zeros(colidx.size)
    u = np.ascontiguousarray(u, dtype=np.float64)
    rc = lib(omp).orc_jacobian(problem, nc - 1, ctypes.c_int64(nC), gidx.size, _p(cells, c_i32p),
                               _p(coords, c_f64p), _p(gidx, c_i32p), N, _p(u, c_f64p),
                               _p(rowptr, c_i32p), _p(colidx, c_i32p), _p(vals, c_f64p))
    assert rc == 0
    return vals


def spmv(rowptr, colidx, vals, x, omp=False):
    N = rowptr.size - 1
    y = np.zeros(N)
    x = np.ascontiguousarray(x, dtype=np.float64)
    lib(omp).orc_spmv(N, _p(rowptr, c_i32p), _p(colidx, c_i32p), _p(vals, c_f64p), _p(x, c_f64p), _p(y, c_f64p))
    return y


def cg(rowptr, colidx, vals, b, rtol=1e-5, abstol=1e-50, dtol=1e5, maxit=10000, pc=0, omp=False, history=False):
    N = rowptr.size - 1
    x = np.zeros(N)
    b = np.ascontiguousarray(b, dtype=np.float64)
    its = ctypes.c_int()
    reason = ctypes.c_int()
    rnorm = ctypes.c_double()
    hist = np.zeros(maxit + 1) if history else None
    lib(omp).orc_cg(N, _p(rowptr, c_i32p), _p(colidx, c_i32p), _p(vals, c_f64p), _p(b, c_f64p), _p(x, c_f64p),
                    ctypes.c_double(rtol), ctypes.c_double(abstol), ctypes.c_double(dtol), maxit, pc,
                    ctypes.byref(its), ctypes.byref(reason), ctypes.byref(rnorm),
                    _p(hist, c_f64p) if history else None)
    out = dict(x=x, its=its.value, reason=reason.value, rnorm=rnorm.value)
    if history:
        out["hist"] = hist[: its.value + 1]
    return out


def newton(problem, cells, coords, gidx, N, rowptr, colidx, u0=None, snes_rtol=1e-8, snes_atol=1e-50,
           snes_stol=1e-8, snes_maxit=50, ksp_rtol=1e-5, ksp_atol=1e-50, ksp_maxit=10000, pc=0, omp=False):
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    nC, nc = cells.shape
    u = np.zeros(N) if u0 is None else np.array(u0, dtype=np.float64)
    out = np.zeros(3, dtype=np.int32)
    norms = np.zeros(snes_maxit + 2)
    lib(omp).orc_newton(problem, nc - 1, ctypes.c_int64(nC), gidx.size, _p(cells, c_i32p), _p(coords, c_f64p),
                        _p(gidx, c_i32p), N, _p(rowptr, c_i32p), _p(colidx, c_i32p), _p(u, c_f64p),
                        ctypes.c_double(snes_rtol), ctypes.c_double(snes_atol), ctypes.c_double(snes_stol),
                        snes_maxit, ctypes.c_double(ksp_rtol), ctypes.c_double(ksp_atol), ksp_maxit, pc,
                        _p(out, c_i32p), _p(norms, c_f64p))
    return dict(u=u, snes_its=int(out[0]), ksp_its=int(out[1]), reason=int(out[2]), norms=norms[: out[0] + 1].copy())


def bd_residual(cells, coords, gidx, N, u, F, bd_set=0):
    """F += natural-boundary term of ex12's -bc_type neumann (f0_bd_u, f1_bd_zero); returns F.
    bd_set=1: a synthetic callback pair that reads u, grad u, x and n and has a non-zero f1."""
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    nC, nc = cells.shape
    u = np.ascontiguousarray(u, dtype=np.float64)
    F = np.ascontiguousarray(F, dtype=np.float64)
    rc = lib().orc_bd_residual(int(bd_set), nc - 1, ctypes.c_int64(nC), gidx.size, _p(cells, c_i32p), _p(coords, c_f64p),
                               _p(gidx, c_i32p), N, _p(u, c_f64p), _p(F, c_f64p))
    assert rc == 0
    return F


def hash_noise(n, start=0):
    L = lib()
    L.orc_hash_noise.restype = ctypes.c_double
    return np.array([L.orc_hash_noise(ctypes.c_int64(start + i)) for i in range(n)])


def estimate_emax(rowptr, colidx, vals, dinv, steps=10):
    L = lib()
    L.orc_estimate_emax.restype = ctypes.c_double
    dinv = np.ascontiguousarray(dinv, dtype=np.float64)
    return L.orc_estimate_emax(rowptr.size - 1, _p(rowptr, c_i32p), _p(colidx, c_i32p), _p(vals, c_f64p),
                               _p(dinv, c_f64p), steps)


def pcg(rowptr, colidx, vals, b, rtol=1e-5, abstol=1e-50, dtol=1e5, maxit=10000, pc=0, cheb_its=4, emin=0.0,
        emax=0.0, nullspace=False, history=False):
    """orc_pcg: pc 0 none, 1 Jacobi, 2 Chebyshev(cheb_its)/Jacobi; nullspace = project the constants out."""
    N = rowptr.size - 1
    x = np.zeros(N)
    b = np.ascontiguousarray(b, dtype=np.float64)
    its = ctypes.c_int()
    reason = ctypes.c_int()
    rnorm = ctypes.c_double()
    eig = np.zeros(2)
    hist = np.zeros(maxit + 1) if history else None
    lib().orc_pcg(N, _p(rowptr, c_i32p), _p(colidx, c_i32p), _p(vals, c_f64p), _p(b, c_f64p), _p(x, c_f64p),
                  ctypes.c_double(rtol), ctypes.c_double(abstol), ctypes.c_double(dtol), maxit, pc, cheb_its,
                  ctypes.c_double(emin), ctypes.c_double(emax), int(bool(nullspace)),
                  ctypes.byref(its), ctypes.byref(reason), ctypes.byref(rnorm),
                  _p(hist, c_f64p) if history else None, _p(eig, c_f64p))
    out = dict(x=x, its=its.value, reason=reason.value, rnorm=rnorm.value, emin=eig[0], emax=eig[1])
    if history:
        out["hist"] = hist[: its.value + 1]
    return out


def num_threads(omp=True):
    return lib(omp).orc_num_threads()


def use_all_cores():
    """OpenMP thread count := every core this process may run on, whatever OMP_NUM_THREADS says
    (torch.distributed.run exports OMP_NUM_THREADS=1).  Returns the count in use."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return lib(True).orc_set_num_threads(int(n))


# ---------------------------------------------------------------------------
# synthetic inputs (C4 box of nx*ny quads split on the same diagonal,
# vertices row-major; C5 cube of n^3 hexes x 6 Kuhn tets)
# ---------------------------------------------------------------------------

def box2d(nx, ny, lx=1.0, ly=1.0):
    i, j = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    coords = np.stack([i.reshape(-1) * (lx / nx), j.reshape(-1) * (ly / ny)], axis=1).astype(np.float64)
    qi, qj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v00 = (qj * (nx + 1) + qi).reshape(-1)
    v10, v01, v11 = v00 + 1, v00 + nx + 1, v00 + nx + 2
    cells = np.empty((nx * ny * 2, 3), dtype=np.int32)
    cells[0::2] = np.stack([v00, v10, v11], axis=1)
    cells[1::2] = np.stack([v00, v11, v01], axis=1)
    return cells, coords


_KUHN = [(0, 1, 3, 7), (0, 3, 2, 7), (0, 2, 6, 7), (0, 6, 4, 7), (0, 4, 5, 7), (0, 5, 1, 7)]


def box3d(nx, ny, nz):
    k, j, i = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    coords = np.stack([i.reshape(-1) / nx, j.reshape(-1) / ny, k.reshape(-1) / nz], axis=1).astype(np.float64)
    qk, qj, qi = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = ((qk * (ny + 1) + qj) * (nx + 1) + qi).reshape(-1)
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    corner = [base + (c & 1) * sx + ((c >> 1) & 1) * sy + ((c >> 2) & 1) * sz for c in range(8)]
    cells = np.empty((base.size * 6, 4), dtype=np.int32)
    for t, tet in enumerate(_KUHN):
        cells[t::6] = np.stack([corner[c] for c in tet], axis=1)
    return cells, coords
