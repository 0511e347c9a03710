"""CPU tests of the ORACLE itself (-m "not gpu"): golden vectors, an independent scipy assembly,
and the analytic anchors the reference source provides (ex12.cpp:88-96,113-117,1651-1661)."""
import hashlib

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import load_osh_oracle, synthetic_meshes


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def setup_oracle(oracle, cells, coords):
    bnd = oracle.mark_boundary(cells, coords.shape[0])
    gidx, N = oracle.numbering(bnd)
    rowptr, colidx = oracle.pattern(cells, gidx, N)
    return bnd, gidx, N, rowptr, colidx


def independent_stiffness(cells, coords):
    """Textbook P1 stiffness K_ij = |K| grad(l_i).grad(l_j) and load b_i = f |K|/(d+1), written
    with barycentric gradients (no reference element, no quadrature): independent of oracle.c."""
    nV, dim = coords.shape
    nc = dim + 1
    rows, cols, vals = [], [], []
    load = np.zeros(nV)
    for cv in cells:
        X = coords[cv]
        M = np.hstack([np.ones((nc, 1)), X])
        vol = abs(np.linalg.det(M)) / (2.0 if dim == 2 else 6.0)
        grads = np.linalg.inv(M)[1:, :].T  # (nc, dim): gradient of each barycentric coordinate
        K = vol * grads @ grads.T
        for a in range(nc):
            load[cv[a]] += vol / nc
            for b in range(nc):
                rows.append(cv[a]); cols.append(cv[b]); vals.append(K[a, b])
    return sp.csr_matrix((vals, (rows, cols)), shape=(nV, nV)), load


def boundary_by_edge_count(cells, nV):
    dim = cells.shape[1] - 1
    from collections import Counter
    cnt = Counter()
    for cv in cells:
        for skip in range(dim + 1):
            cnt[tuple(sorted(np.delete(cv, skip)))] += 1
    bnd = np.zeros(nV, np.uint8)
    for f, k in cnt.items():
        if k == 1:
            bnd[list(f)] = 1
    return bnd


@pytest.mark.parametrize("name,file", [("T23", "23elm.osh"), ("T24k", "24k.osh")])
def test_golden_fixture_meshes(oracle_mod, expected, name, file):
    cells, coords, m = load_osh_oracle(file)
    assert m["consumed"] == m["total"], "fixture must decode to the last byte"
    e = expected[name]
    assert sha(cells.astype(np.int32)) == e["sha_elem_verts"] and sha(coords) == e["sha_coords"]
    bnd, gidx, N, rowptr, colidx = setup_oracle(oracle_mod, cells, coords)
    assert (coords.shape[0], cells.shape[0], int(bnd.sum()), N, colidx.size) == (
        e["n_verts"], e["n_cells"], e["n_boundary"], e["N"], e["nnz"])
    assert sha(bnd) == e["sha_bnd"] and sha(gidx.astype(np.int64)) == e["sha_gidx"]
    assert sha(rowptr) == e["sha_rowptr"] and sha(colidx.astype(np.int64)) == e["sha_colidx"]
    F0 = oracle_mod.residual(0, cells, coords, gidx, N, np.zeros(N))
    assert np.linalg.norm(F0) == pytest.approx(e["F0_norm"], rel=1e-13)
    # positive orientation and total area recorded by the survey probe
    X = coords[cells]
    area = 0.5 * ((X[:, 1, 0] - X[:, 0, 0]) * (X[:, 2, 1] - X[:, 0, 1]) - (X[:, 1, 1] - X[:, 0, 1]) * (X[:, 2, 0] - X[:, 0, 0]))
    assert (area > 0).all()
    assert area.sum() == pytest.approx(3.03019611 if name == "T24k" else 0.98877, rel=1e-5)


def test_t23_pattern_literal(oracle_mod, expected):
    cells, coords, _ = load_osh_oracle("23elm.osh")
    bnd, gidx, N, rowptr, colidx = setup_oracle(oracle_mod, cells, coords)
    assert rowptr.tolist() == [0, 4, 10, 14, 18, 22, 26] == expected["T23"]["rowptr"]
    assert np.nonzero(gidx >= 0)[0].tolist() == [0, 2, 3, 4, 5, 6]
    assert colidx.tolist() == expected["T23"]["colidx"]


@pytest.mark.parametrize("name", ["box2d_8x8", "box2d_33x17", "box3d_5x4x3"])
def test_golden_synthetic(oracle_mod, expected, name):
    cells, coords = synthetic_meshes()[name]
    e = expected[name]
    bnd, gidx, N, rowptr, colidx = setup_oracle(oracle_mod, cells, coords)
    assert (int(bnd.sum()), N, colidx.size) == (e["n_boundary"], e["N"], e["nnz"])
    assert sha(rowptr) == e["sha_rowptr"] and sha(colidx.astype(np.int64)) == e["sha_colidx"]
    vals = oracle_mod.jacobian(0, cells, coords, gidx, N, np.zeros(N), rowptr, colidx)
    assert np.linalg.norm(vals) == pytest.approx(e["vals_fro"], rel=1e-13)
    cg = oracle_mod.cg(rowptr, colidx, vals, oracle_mod.residual(0, cells, coords, gidx, N, np.zeros(N)), rtol=1e-9)
    assert cg["its"] == e["cg_its_none"]


@pytest.mark.parametrize("name", ["T23", "T24k", "box2d_33x17", "sheared_21x13", "box3d_5x4x3", "warped3d_6x5x7", "fan4"])
def test_oracle_vs_independent_assembly(oracle_mod, name):
    if name == "T23":
        cells, coords, _ = load_osh_oracle("23elm.osh")
    elif name == "T24k":
        cells, coords, _ = load_osh_oracle("24k.osh")
    else:
        cells, coords = synthetic_meshes()[name]
    nV, dim = coords.shape
    bnd, gidx, N, rowptr, colidx = setup_oracle(oracle_mod, cells, coords)
    assert np.array_equal(bnd, boundary_by_edge_count(cells, nV))
    K, load = independent_stiffness(cells, coords)
    free = np.nonzero(bnd == 0)[0]
    assert np.array_equal(gidx[free], np.arange(free.size)) and (gidx[bnd == 1] == -1).all()
    Kff = K[free][:, free].tocsr()
    Kff.sort_indices()
    # pattern: structural adjacency (explicit zeros kept) == pattern of the cell-sharing graph
    inc = sp.csr_matrix((np.ones(cells.size), (np.repeat(np.arange(cells.shape[0]), dim + 1), cells.reshape(-1))),
                        shape=(cells.shape[0], nV))
    adj = (inc.T @ inc).tocsr()[free][:, free].tocsr()
    adj.sort_indices()
    assert np.array_equal(adj.indptr, rowptr) and np.array_equal(adj.indices, colidx)
    rng = np.random.default_rng(3)
    u = rng.standard_normal(N)
    vals = oracle_mod.jacobian(0, cells, coords, gidx, N, u, rowptr, colidx)
    A = sp.csr_matrix((vals, colidx, rowptr), shape=(N, N))
    scale = abs(Kff).max() if N else 1.0
    assert abs(A - Kff).max() <= 1e-12 * scale if N else True
    # residual: F(u) = 4*load + K u_full  with Dirichlet data x^2+y^2 (2/3 r^2 in 3-D)
    ufull = (coords ** 2).sum(axis=1) * (1.0 if dim == 2 else 2.0 / 3.0)
    ufull[free] = u
    Fref = (4.0 * load + K @ ufull)[free]
    F = oracle_mod.residual(0, cells, coords, gidx, N, u)
    assert np.abs(F - Fref).max() <= 1e-12 * max(1.0, np.abs(Fref).max())
    if N:
        # Jacobian consistency A u + F(0) = F(u)  (ex12.cpp:1651-1661) and symmetry
        F0 = oracle_mod.residual(0, cells, coords, gidx, N, np.zeros(N))
        assert np.abs(A @ u + F0 - F).max() <= 1e-11 * np.abs(F).max()
        assert abs(A - A.T).max() <= 1e-13 * scale
        x = oracle_mod.cg(rowptr, colidx, vals, F0, rtol=1e-12, maxit=20000)
        xs = spla.spsolve(Kff.tocsc(), F0)
        assert np.abs(x["x"] - xs).max() <= 1e-8 * np.abs(xs).max()


def test_manufactured_solution_is_nodally_exact_on_the_box(oracle_mod):
    """u = x^2+y^2, f = 4 (ex12.cpp:88-96): on the same-diagonal box P1 reproduces it at the nodes."""
    cells, coords = oracle_mod.box2d(32, 32)
    bnd, gidx, N, rowptr, colidx = setup_oracle(oracle_mod, cells, coords)
    out = oracle_mod.newton(0, cells, coords, gidx, N, rowptr, colidx, ksp_rtol=1e-13, ksp_maxit=5000)
    exact = (coords ** 2).sum(axis=1)[gidx >= 0]
    assert out["reason"] == 3 and out["snes_its"] == 1
    assert np.abs(out["u"] - exact).max() < 1e-11
    vals = oracle_mod.jacobian(0, cells, coords, gidx, N, np.zeros(N), rowptr, colidx)
    assert (vals == 0.0).sum() > 0, "hypotenuse couplings are exact zeros that stay in the pattern"


@pytest.mark.parametrize("problem", [1, 2])
def test_variable_and_nonlinear_problems_converge(oracle_mod, problem):
    cells, coords = oracle_mod.box2d(16, 16)
    bnd, gidx, N, rowptr, colidx = setup_oracle(oracle_mod, cells, coords)
    # the p-Laplacian Jacobian is singular where grad u = 0: start from a scaled exact field
    u0 = 0.7 * (coords ** 2).sum(axis=1)[gidx >= 0] if problem == 2 else None
    out = oracle_mod.newton(problem, cells, coords, gidx, N, rowptr, colidx, u0=u0, ksp_rtol=1e-10, ksp_maxit=5000)
    assert out["reason"] in (2, 3, 4)
    exact = (coords ** 2).sum(axis=1)[gidx >= 0]
    assert np.abs(out["u"] - exact).max() < 5e-3
    if problem == 2:
        assert out["snes_its"] > 1


def test_quadrature_rule(oracle_mod):
    for dim, total in ((2, 2.0), (3, 4.0 / 3.0)):
        xi, w = oracle_mod.quadrature(dim)
        assert w.sum() == pytest.approx(total, rel=1e-14)
        # exact to degree 2 on the biunit simplex: int xi_0^2 = 2/3 (2-D), 8/15 (3-D)
        assert (w * xi[:, 0] ** 2).sum() == pytest.approx(2.0 / 3.0 if dim == 2 else 8.0 / 15.0, rel=1e-13)
        assert (w * xi[:, 0]).sum() == pytest.approx(-2.0 / 3.0 if dim == 2 else -2.0 / 3.0, rel=1e-13)


def test_osh_writer_roundtrip(oracle_mod, tmp_path):
    from oracle import osh_oracle as oo
    cells, coords = oracle_mod.box2d(7, 5)
    oo.write_osh(str(tmp_path / "box.osh"), 2, coords, cells)
    m = oo.read_osh(str(tmp_path / "box.osh"))
    assert m["consumed"] == m["total"]
    assert np.array_equal(oo.elem_verts(m), cells) and np.array_equal(oo.coords(m), coords)
