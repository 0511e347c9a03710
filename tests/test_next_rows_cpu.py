"""CPU checks of the SURVEY 8(f) rank-4 slice (natural boundary term, constant null space, Chebyshev/Jacobi polynomial
preconditioner): the oracle against independent numpy / scipy restatements, and the product's host-side control flow
(omegahpetsc_b200/csrc/ohp_pcg.hpp, instantiated over host arrays by tests/host/pcg_host.cpp) against the oracle.
No GPU is needed; the device kernels behind the same control flow are covered by tests/test_zz_next_rows_gpu.py."""
import ctypes
import os
import subprocess
from collections import defaultdict

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from conftest import ROOT, synthetic_meshes

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_f64p = ctypes.POINTER(ctypes.c_double)


def _p(a, t):
    return a.ctypes.data_as(t)


@pytest.fixture(scope="module")
def pcg_host():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libpcg_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "omegahpetsc_b200", "csrc"),
                    os.path.join(ROOT, "tests", "host", "pcg_host.cpp"), "-o", so], check=True)
    L = ctypes.CDLL(so)
    L.pcg_host_tridiag_max.restype = ctypes.c_double

    def solve(rowptr, colidx, vals, b, rtol=1e-9, abstol=1e-50, dtol=1e5, maxit=10000, pc=0, cheb_its=4, emin=0.0, emax=0.0,
              nullspace=False):
        N = rowptr.size - 1
        x = np.zeros(N)
        its, reason, spmvs = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        rnorm = ctypes.c_double()
        eig = np.zeros(2)
        hist = np.zeros(maxit + 1)
        b = np.ascontiguousarray(b, dtype=np.float64)
        rc = L.pcg_host_solve(N, _p(rowptr, c_i32p), _p(colidx, c_i32p), _p(vals, c_f64p), _p(b, c_f64p), _p(x, c_f64p),
                              ctypes.c_double(rtol), ctypes.c_double(abstol), ctypes.c_double(dtol), maxit, pc, cheb_its,
                              ctypes.c_double(emin), ctypes.c_double(emax), int(nullspace), ctypes.byref(its),
                              ctypes.byref(reason), ctypes.byref(rnorm), _p(eig, c_f64p), _p(hist, c_f64p), hist.size,
                              ctypes.byref(spmvs))
        assert rc == 0
        return dict(x=x, its=its.value, reason=reason.value, rnorm=rnorm.value, emin=eig[0], emax=eig[1],
                    hist=hist[: its.value + 1], spmvs=spmvs.value)

    solve.lib = L
    return solve


def _system(oracle_mod, cells, coords, dirichlet=True, problem=0):
    nV = coords.shape[0]
    bnd = oracle_mod.mark_boundary(cells, nV) if dirichlet else np.zeros(nV, np.uint8)
    gidx, N = oracle_mod.numbering(bnd)
    rp, ci = oracle_mod.pattern(cells, gidx, N)
    vals = oracle_mod.jacobian(problem, cells, coords, gidx, N, np.zeros(N), rp, ci)
    F = oracle_mod.residual(problem, cells, coords, gidx, N, np.zeros(N))
    return gidx, N, rp, ci, vals, F


# ---------------------------------------------------------------------------------------------------------------
# natural boundary term (ex12.cpp:179-195, 1226-1238)
# ---------------------------------------------------------------------------------------------------------------
def _independent_bd(cells, coords):
    """boundary facets by a dictionary of sorted vertex tuples; integral of phi_a * (-2 x.n) by rules that are exact for
    quadratics (Simpson on edges, the edge-midpoint rule on triangles) -- shares nothing with oracle.c"""
    nC, nc = cells.shape
    dim = nc - 1
    sup = defaultdict(list)
    for c in range(nC):
        for skip in range(nc):
            sup[tuple(sorted(np.delete(cells[c], skip)))].append((c, skip))
    out = np.zeros(coords.shape[0])
    for f, lst in sup.items():
        if len(lst) != 1:
            continue
        c, skip = lst[0]
        P, Po = coords[list(f)], coords[cells[c][skip]]
        if dim == 2:
            t = P[1] - P[0]
            L = np.linalg.norm(t)
            n = np.array([t[1], -t[0]]) / L
            if n @ (Po - P[0]) > 0:
                n = -n
            g = lambda x: -2 * (x @ n)
            xm = 0.5 * (P[0] + P[1])
            for a in range(2):
                out[f[a]] += L / 6 * ((a == 0) * g(P[0]) + 4 * 0.5 * g(xm) + (a == 1) * g(P[1]))
        else:
            nn = np.cross(P[1] - P[0], P[2] - P[0])
            area = 0.5 * np.linalg.norm(nn)
            n = nn / np.linalg.norm(nn)
            if n @ (Po - P[0]) > 0:
                n = -n
            g = lambda x: -2 * (x @ n)
            for a in range(3):
                s = 0.0
                for (i, j) in ((0, 1), (1, 2), (0, 2)):
                    s += 0.5 * ((a == i) + (a == j)) * g(0.5 * (P[i] + P[j]))
                out[f[a]] += area / 3 * s
    return out


@pytest.mark.parametrize("name", ["fan4", "box2d_8x8", "sheared_21x13", "box3d_5x4x3", "warped3d_6x5x7"])
def test_oracle_boundary_residual_vs_independent(oracle_mod, name):
    cells, coords = synthetic_meshes()[name]
    nV = coords.shape[0]
    gidx = np.arange(nV, dtype=np.int32)
    F = oracle_mod.bd_residual(cells, coords, gidx, nV, np.zeros(nV), np.zeros(nV))
    G = _independent_bd(cells, coords)
    assert np.abs(F - G).max() <= 1e-13 * max(1.0, np.abs(G).max())
    # interior vertices receive nothing
    bnd = oracle_mod.mark_boundary(cells, nV)
    assert np.all(F[bnd == 0] == 0.0)


def test_oracle_neumann_total_flux_on_boxes(oracle_mod):
    """test function 1: sum_i F_i(0) = int f0 dx + oint f0_bd ds = 4|Omega| - int div(2x) dx: 0 in 2-D, -2|Omega| in 3-D
    (ex12's f0_bd_u keeps the factor 2 in 3-D although quadratic_u_3d carries 2/3: restated as it is)"""
    for (cells, coords), expect in ((oracle_mod.box2d(9, 7, 2.0, 1.5), 0.0), (oracle_mod.box3d(4, 3, 5), -2.0)):
        nV = coords.shape[0]
        gidx = np.arange(nV, dtype=np.int32)
        F = oracle_mod.residual(0, cells, coords, gidx, nV, np.zeros(nV))
        F = oracle_mod.bd_residual(cells, coords, gidx, nV, np.zeros(nV), F)
        assert abs(F.sum() - expect) < 1e-12


def test_oracle_neumann_residual_of_exact_solution_converges(oracle_mod):
    """F(I_h u) -> 0 like O(h) in the l2 norm scaled by the dof count for u = x^2 + y^2 with the natural data of
    f0_bd_u, and the unconstrained operator annihilates the constants"""
    errs = []
    for n in (8, 16, 32):
        cells, coords = oracle_mod.box2d(n, n)
        nV = coords.shape[0]
        gidx = np.arange(nV, dtype=np.int32)
        u = (coords ** 2).sum(axis=1)
        F = oracle_mod.residual(0, cells, coords, gidx, nV, u)
        F = oracle_mod.bd_residual(cells, coords, gidx, nV, u, F)
        errs.append(np.abs(F).sum())
        rp, ci = oracle_mod.pattern(cells, gidx, nV)
        vals = oracle_mod.jacobian(0, cells, coords, gidx, nV, u, rp, ci)
        assert np.abs(oracle_mod.spmv(rp, ci, vals, np.ones(nV))).max() < 1e-13
    assert errs[1] < 0.6 * errs[0] and errs[2] < 0.6 * errs[1]


# ---------------------------------------------------------------------------------------------------------------
# Chebyshev / Jacobi polynomial, eigenvalue estimate
# ---------------------------------------------------------------------------------------------------------------
def test_tridiag_lambda_max(pcg_host):
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 10):
        d, e = rng.normal(size=n) + 3.0, rng.normal(size=max(n - 1, 1))
        T = np.diag(d) + (np.diag(e[: n - 1], 1) + np.diag(e[: n - 1], -1) if n > 1 else 0)
        got = pcg_host.lib.pcg_host_tridiag_max(n, _p(d, c_f64p), _p(e, c_f64p))
        assert abs(got - np.linalg.eigvalsh(T).max()) < 1e-12 * abs(got)


def test_oracle_eigen_estimate_and_chebyshev_polynomial(oracle_mod):
    cells, coords = oracle_mod.box2d(24, 24)
    gidx, N, rp, ci, vals, F = _system(oracle_mod, cells, coords)
    A = sp.csr_matrix((vals, ci, rp), shape=(N, N))
    d = A.diagonal()
    S = sp.diags(1 / np.sqrt(d)) @ A @ sp.diags(1 / np.sqrt(d))
    lam = np.linalg.eigvalsh(S.toarray())
    est = oracle_mod.estimate_emax(rp, ci, vals, 1 / d, 10)
    assert 0.9 * lam[-1] < est <= lam[-1] * (1 + 1e-12)  # a Ritz value: never above, close after 10 steps
    # the Chebyshev iteration with the TRUE spectral interval is a convergent solver whose error contracts at the
    # textbook rate 2 rho^k / (1 + rho^2k), rho = (sqrt(kappa) - 1) / (sqrt(kappa) + 1), in the D^-1 A energy norm
    emin, emax = lam[0], lam[-1]
    xs = spl.spsolve(A.tocsc(), F)
    kappa = emax / emin
    rho = (np.sqrt(kappa) - 1) / (np.sqrt(kappa) + 1)
    for k in (5, 20, 60):
        # one application of the polynomial to b = one Chebyshev "solve" with k iterations and zero guess
        r = oracle_mod.pcg(rp, ci, vals, F, rtol=1e-300, maxit=1, pc=2, cheb_its=k, emin=emin, emax=emax)
        # after ONE outer CG step x = alpha * p_k(D^-1 A) D^-1 b with alpha the exact line search: at least as good as the
        # plain Chebyshev iterate in the A norm
        e = r["x"] - xs
        bound = 2 * rho ** k / (1 + rho ** (2 * k))
        assert np.sqrt(e @ (A @ e)) <= 1.0001 * bound * np.sqrt(xs @ (A @ xs))


@pytest.mark.parametrize("name", ["box2d_33x17", "sheared_21x13", "box3d_12x11x10"])
def test_oracle_pcg_against_scipy(oracle_mod, name):
    cells, coords = synthetic_meshes()[name]
    gidx, N, rp, ci, vals, F = _system(oracle_mod, cells, coords)
    A = sp.csr_matrix((vals, ci, rp), shape=(N, N))
    xs = spl.spsolve(A.tocsc(), F)
    base = oracle_mod.cg(rp, ci, vals, F, rtol=1e-10, pc=1)
    same = oracle_mod.pcg(rp, ci, vals, F, rtol=1e-10, pc=1)
    assert same["its"] == base["its"] and np.abs(same["x"] - base["x"]).max() <= 1e-12 * np.abs(xs).max()
    prev = base["its"]
    for k in (2, 4, 8):
        r = oracle_mod.pcg(rp, ci, vals, F, rtol=1e-10, pc=2, cheb_its=k)
        assert r["reason"] == 2
        assert np.abs(r["x"] - xs).max() <= 1e-7 * np.abs(xs).max()
        assert r["its"] <= prev  # the outer iterations (global reductions) drop with the degree
        prev = r["its"]
        if k == 2:
            assert r["its"] < 0.6 * base["its"]


def test_oracle_pcg_null_space(oracle_mod):
    """pure Neumann operator: singular with the constants in the kernel; CG with the projection converges to the
    minimum-norm-in-the-mean solution scipy gives for the same consistent right-hand side"""
    cells, coords = oracle_mod.box2d(20, 14)
    nV = coords.shape[0]
    gidx, N, rp, ci, vals, F = _system(oracle_mod, cells, coords, dirichlet=False)
    F = oracle_mod.bd_residual(cells, coords, gidx, N, np.zeros(N), F)
    assert abs(F.sum()) < 1e-12  # consistent (total flux test above)
    A = sp.csr_matrix((vals, ci, rp), shape=(N, N))
    # reference: bordered system [A 1; 1^T 0]
    K = sp.bmat([[A, np.ones((N, 1))], [np.ones((1, N)), None]]).tocsc()
    xs = spl.spsolve(K, np.concatenate([F, [0.0]]))[:N]
    for pc, k in ((0, 0), (1, 0), (2, 4)):
        r = oracle_mod.pcg(rp, ci, vals, F, rtol=1e-10, pc=pc, cheb_its=k, nullspace=True)
        assert r["reason"] == 2
        x = r["x"] - r["x"].mean()
        assert np.abs(x - xs).max() <= 1e-7 * np.abs(xs).max()


# ---------------------------------------------------------------------------------------------------------------
# the product's control flow (ohp_pcg.hpp over host arrays) against the oracle
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["fan4", "box2d_8x8", "box2d_100x70", "sheared_21x13", "box3d_12x11x10"])
def test_product_control_flow_matches_oracle(oracle_mod, pcg_host, name):
    cells, coords = synthetic_meshes()[name]
    gidx, N, rp, ci, vals, F = _system(oracle_mod, cells, coords)
    for pc, k, emin, emax in ((0, 0, 0, 0), (1, 0, 0, 0), (2, 1, 0, 0), (2, 2, 0, 0), (2, 5, 0, 0), (2, 8, 0, 0), (2, 4, 0.2, 2.2)):
        o = oracle_mod.pcg(rp, ci, vals, F, rtol=1e-9, pc=pc, cheb_its=k, emin=emin, emax=emax, history=True)
        h = pcg_host(rp, ci, vals, F, rtol=1e-9, pc=pc, cheb_its=k, emin=emin, emax=emax)
        assert (h["its"], h["reason"]) == (o["its"], o["reason"]), (pc, k)
        assert abs(h["emin"] - o["emin"]) <= 1e-12 * max(o["emax"], 1) and abs(h["emax"] - o["emax"]) <= 1e-12 * max(o["emax"], 1)
        xs = max(np.abs(o["x"]).max(), 1e-300)
        assert np.abs(h["x"] - o["x"]).max() <= 1e-10 * xs
        assert np.abs(h["hist"] - o["hist"]).max() <= 1e-9 * max(o["hist"][0], 1e-300)
        if pc == 2 and N > 50:  # (on a handful of dofs the Lanczos run of the estimate breaks down early)
            est = 10 if emax == 0 else 0
            assert h["spmvs"] == est + h["its"] * k + (k - 1)  # k - 1 inside every application, 1 outside


def test_product_control_flow_null_space(oracle_mod, pcg_host):
    cells, coords = oracle_mod.box3d(6, 5, 4)
    gidx, N, rp, ci, vals, F = _system(oracle_mod, cells, coords, dirichlet=False)
    F = F - F.mean()  # make the right-hand side consistent (ex12's 3-D natural data is not, see the flux test)
    for pc, k in ((0, 0), (1, 0), (2, 3)):
        o = oracle_mod.pcg(rp, ci, vals, F, rtol=1e-9, pc=pc, cheb_its=k, nullspace=True)
        h = pcg_host(rp, ci, vals, F, rtol=1e-9, pc=pc, cheb_its=k, nullspace=True)
        assert (h["its"], h["reason"]) == (o["its"], o["reason"]) and o["reason"] == 2
        assert np.abs(h["x"] - o["x"]).max() <= 1e-10 * np.abs(o["x"]).max()


def test_product_control_flow_edge_cases(oracle_mod, pcg_host):
    # zero right-hand side: converged at once with reason ATOL (dp = 0 < abstol), as KSPConvergedDefault
    cells, coords = oracle_mod.box2d(5, 5)
    gidx, N, rp, ci, vals, F = _system(oracle_mod, cells, coords)
    for pc in (0, 1, 2):
        h = pcg_host(rp, ci, vals, np.zeros(N), pc=pc, cheb_its=3)
        o = oracle_mod.pcg(rp, ci, vals, np.zeros(N), rtol=1e-9, pc=pc, cheb_its=3)
        assert (h["its"], h["reason"]) == (o["its"], o["reason"]) == (0, 3)
    # iteration limit
    h = pcg_host(rp, ci, vals, F, rtol=1e-14, maxit=3, pc=2, cheb_its=2)
    assert (h["its"], h["reason"]) == (3, -3)


def test_product_control_flow_rejects_an_empty_interval(oracle_mod, pcg_host):
    cells, coords = oracle_mod.box2d(5, 5)
    gidx, N, rp, ci, vals, F = _system(oracle_mod, cells, coords)
    L = pcg_host.lib
    x, eig, hist = np.zeros(N), np.zeros(2), np.zeros(8)
    its, reason, spmvs, rnorm = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
    for emin, emax in ((2.0, 1.0), (1.0, 1.0), (5.0, 0.0)):  # reversed, empty, lower end above the estimated upper end
        rc = L.pcg_host_solve(N, _p(rp, c_i32p), _p(ci, c_i32p), _p(vals, c_f64p), _p(F, c_f64p), _p(x, c_f64p),
                              ctypes.c_double(1e-9), ctypes.c_double(1e-50), ctypes.c_double(1e5), 100, 2, 4,
                              ctypes.c_double(emin), ctypes.c_double(emax), 0, ctypes.byref(its), ctypes.byref(reason),
                              ctypes.byref(rnorm), _p(eig, c_f64p), _p(hist, c_f64p), hist.size, ctypes.byref(spmvs))
        assert rc == 62, (emin, emax, rc)


# ---------------------------------------------------------------------------------------------------------------
# the product's natural-boundary row function (include/ohp_fem.cuh: bd_residual_row, __host__ __device__) on host arrays
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def bd_host():
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libbd_host.so")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "--expt-relaxed-constexpr",
                    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "host", "bd_host.cu"), "-o", so], check=True)
    L = ctypes.CDLL(so)

    def rows(bd_set, cells, coords, lidx, u, bnd):
        """the data layout ohp_mesh_setup builds: stars as (cell*4 + corner) ascending, rows = vertices with lidx >= 0"""
        cells = np.ascontiguousarray(cells, np.int32)
        coords = np.ascontiguousarray(coords, np.float64)
        nC, nc = cells.shape
        nV = coords.shape[0]
        code = (np.arange(nC, dtype=np.int64)[:, None] * 4 + np.arange(nc)[None, :]).reshape(-1)
        vert = cells.reshape(-1)
        order = np.lexsort((code, vert))
        v2c = code[order].astype(np.int32)
        v2c_ptr = np.zeros(nV + 1, np.int32)
        np.add.at(v2c_ptr, vert + 1, 1)
        v2c_ptr = np.cumsum(v2c_ptr).astype(np.int32)
        lidx = np.ascontiguousarray(lidx, np.int32)
        n_rows = int((lidx >= 0).sum())
        row_vertex = np.zeros(max(n_rows, 1), np.int32)
        row_vertex[lidx[lidx >= 0]] = np.nonzero(lidx >= 0)[0]
        u = np.ascontiguousarray(u, np.float64)
        bnd = np.ascontiguousarray(bnd, np.uint8)
        outv = np.zeros(max(n_rows, 1))
        c_u8p = ctypes.POINTER(ctypes.c_uint8)
        rc = L.bd_host_rows(int(bd_set), nc - 1, nC, nV, _p(cells, c_i32p), _p(coords, c_f64p), _p(lidx, c_i32p), _p(u, c_f64p),
                            _p(v2c_ptr, c_i32p), _p(v2c, c_i32p), n_rows, _p(row_vertex, c_i32p), _p(bnd, c_u8p), _p(outv, c_f64p))
        assert rc == 0
        return outv[:n_rows]

    return rows


@pytest.mark.parametrize("name", ["one_tri", "fan4", "box2d_8x8", "box2d_33x17", "sheared_21x13", "box3d_5x4x3", "warped3d_6x5x7"])
@pytest.mark.parametrize("bd_set", [0, 1])
def test_product_boundary_rows_match_oracle(oracle_mod, bd_host, name, bd_set):
    cells, coords = synthetic_meshes()[name]
    nV = coords.shape[0]
    bnd = oracle_mod.mark_boundary(cells, nV)
    gidx = np.arange(nV, dtype=np.int32)  # natural condition: nothing is constrained
    u = np.sin(3 * coords[:, 0]) * np.cos(2 * coords[:, 1]) + 0.3
    want = oracle_mod.bd_residual(cells, coords, gidx, nV, u, np.zeros(nV), bd_set=bd_set)
    got = bd_host(bd_set, cells, coords, gidx, u, bnd)
    assert np.abs(got - want).max() <= 1e-13 * max(1.0, np.abs(want).max())
    assert np.abs(want).max() > 0


def test_product_boundary_rows_with_constrained_and_permuted_dofs(oracle_mod, bd_host):
    """rows in another order than the vertices, and some vertices constrained (their value comes from the Dirichlet function
    inside the row function): the mixed situation of a user who integrates a boundary term on an essential-BC mesh"""
    cells, coords = synthetic_meshes()["sheared_21x13"]
    nV = coords.shape[0]
    bnd = oracle_mod.mark_boundary(cells, nV)
    rng = np.random.default_rng(5)
    constrained = (rng.random(nV) < 0.2) & (bnd == 1)
    free = np.nonzero(~constrained)[0]
    perm = rng.permutation(free.size)
    gidx = np.full(nV, -1, np.int32)
    gidx[free] = perm
    u = rng.normal(size=free.size)
    want = oracle_mod.bd_residual(cells, coords, gidx, free.size, u, np.zeros(free.size), bd_set=1)
    got = bd_host(1, cells, coords, gidx, u, bnd)
    assert np.abs(got - want).max() <= 1e-13 * max(1.0, np.abs(want).max())


# ---------------------------------------------------------------------------------------------------------------
# golden values of this slice (tests/golden/expected_next_rows.json, written by tests/golden/make_golden_next_rows.py)
# ---------------------------------------------------------------------------------------------------------------
def test_oracle_against_recorded_next_row_values(oracle_mod):
    import hashlib
    import json
    from conftest import GOLDEN, load_osh_oracle
    with open(os.path.join(GOLDEN, "expected_next_rows.json")) as f:
        exp = json.load(f)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    for name in ("T23", "T24k", "box2d_33x17", "box3d_5x4x3"):
        if name.startswith("T"):
            cells, coords, _ = load_osh_oracle({"T23": "23elm.osh", "T24k": "24k.osh"}[name])
        else:
            cells, coords = synthetic_meshes()[name]
        e = exp[name]
        nV = coords.shape[0]
        gid = np.arange(nV, dtype=np.int32)
        rowptr, colidx = oracle_mod.pattern(cells, gid, nV)
        assert (nV, colidx.size) == (e["natural_N"], e["natural_nnz"])
        assert sha(rowptr.astype(np.int32)) == e["natural_sha_rowptr"] and sha(colidx.astype(np.int64)) == e["natural_sha_colidx"]
        F0 = oracle_mod.bd_residual(cells, coords, gid, nV, np.zeros(nV), oracle_mod.residual(0, cells, coords, gid, nV, np.zeros(nV)))
        assert np.linalg.norm(F0) == pytest.approx(e["natural_F0_norm"], rel=1e-12)
        gidx, N, rp, ci, vals, F = _system(oracle_mod, cells, coords)
        for k in (1, 4, 8, 16):
            r = oracle_mod.pcg(rp, ci, vals, F, rtol=1e-9, pc=2, cheb_its=k, maxit=100000)
            g = e["chebyshev_k%d" % k]
            assert (r["its"], r["reason"]) == (g["its"], g["reason"]) and r["emax"] == pytest.approx(g["emax"], rel=1e-12)
            assert np.linalg.norm(r["x"]) == pytest.approx(g["x_norm"], rel=1e-9)
    # T24k with the polynomial of degree 1 is Jacobi-preconditioned CG: the count recorded in expected.json
    with open(os.path.join(GOLDEN, "expected.json")) as f:
        assert exp["T24k"]["chebyshev_k1"]["its"] == json.load(f)["T24k"]["cg_its_jacobi"]


@pytest.mark.parametrize("dim,seed", [(2, 1), (2, 2), (3, 3), (3, 4)])
def test_product_boundary_rows_on_unstructured_delaunay_meshes(oracle_mod, bd_host, dim, seed):
    """irregular stars, boundary vertices with many boundary facets, cells of either orientation (the facet normal is taken
    from the opposite vertex, so the orientation of the cell must not matter): three independent computations agree"""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = rng.random((60 if dim == 2 else 40, dim))
    cells = Delaunay(pts).simplices.astype(np.int32)
    used = np.unique(cells)
    remap = np.full(pts.shape[0], -1, np.int32)
    remap[used] = np.arange(used.size, dtype=np.int32)
    cells, coords = remap[cells], pts[used]
    nV = coords.shape[0]
    bnd = oracle_mod.mark_boundary(cells, nV)
    gidx = np.arange(nV, dtype=np.int32)
    want = oracle_mod.bd_residual(cells, coords, gidx, nV, np.zeros(nV), np.zeros(nV))
    assert np.abs(want - _independent_bd(cells, coords)).max() <= 1e-12 * np.abs(want).max()
    for bd_set in (0, 1):
        u = np.cos(2 * coords[:, 0]) + coords[:, -1] ** 2
        ref = oracle_mod.bd_residual(cells, coords, gidx, nV, u, np.zeros(nV), bd_set=bd_set)
        got = bd_host(bd_set, cells, coords, gidx, u, bnd)
        assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
