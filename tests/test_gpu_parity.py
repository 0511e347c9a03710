"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI, against the oracle on the
same inputs.  Contract (BASELINE.json north_star): numbering and sparsity pattern BIT-EXACT;
assembled values within 1e-12 relative (relative to the largest magnitude of the row / vector:
exact-zero couplings of right-angled cells make an entrywise ratio meaningless); solution within
the solver tolerance."""
import os

import numpy as np
import pytest

import omegahpetsc_b200 as ohp
from conftest import ROOT, load_osh_oracle, synthetic_meshes

pytestmark = pytest.mark.gpu

VAL_RTOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    c = ohp.Context(0)
    yield c
    c.close()


def all_meshes():
    out = dict(synthetic_meshes())
    for name, file in (("T23", "23elm.osh"), ("T24k", "24k.osh")):
        c, x, _ = load_osh_oracle(file)
        out[name] = (c, x)
    return out


MESHES = None


def get_mesh(name):
    global MESHES
    if MESHES is None:
        MESHES = all_meshes()
    return MESHES[name]


NAMES = ["one_tri", "fan4", "T23", "box2d_8x8", "box2d_33x17", "box2d_100x70", "sheared_21x13", "T24k",
         "box3d_5x4x3", "box3d_12x11x10", "warped3d_6x5x7"]


def oracle_setup(oracle, cells, coords):
    bnd = oracle.mark_boundary(cells, coords.shape[0])
    gidx, N = oracle.numbering(bnd)
    rowptr, colidx = oracle.pattern(cells, gidx, N)
    return bnd, gidx, N, rowptr, colidx


def probe_field(coords, gidx):
    x = coords[gidx >= 0]
    return np.sin(3.0 * x[:, 0]) * np.cos(2.0 * x[:, 1]) + (0.3 * x[:, 2] if x.shape[1] == 3 else 0.0)


def rowwise_close(vals, ref, rowptr, rtol):
    if vals.size == 0:
        return True
    rowmax = np.maximum.reduceat(np.abs(ref), rowptr[:-1][np.diff(rowptr) > 0])
    scale = np.repeat(rowmax, np.diff(rowptr)[np.diff(rowptr) > 0])
    return bool((np.abs(vals - ref) <= rtol * scale).all())


@pytest.mark.parametrize("name", NAMES)
def test_numbering_and_pattern_bit_exact(ctx, oracle_mod, name):
    cells, coords = get_mesh(name)
    bnd, gidx, N, rowptr, colidx = oracle_setup(oracle_mod, cells, coords)
    m = ohp.Mesh(ctx, cells, coords).setup()
    info = m.info()
    assert (info.n_cells, info.n_verts, info.n_boundary_verts, info.n_owned, info.n_global, info.nnz) == (
        cells.shape[0], coords.shape[0], int(bnd.sum()), N, N, colidx.size)
    assert np.array_equal(m.boundary(), bnd)
    assert np.array_equal(m.numbering(), gidx.astype(np.int64))
    rp, ci = m.csr()
    assert np.array_equal(rp, rowptr)
    assert np.array_equal(ci, colidx.astype(np.int64))
    assert info.max_row == (int(np.diff(rowptr).max()) if N else 0)
    m.destroy()


def test_golden_hashes_T24k(ctx, expected):
    import hashlib
    m = ohp.Mesh.from_osh(ctx, __import__("conftest").golden_mesh_path("24k.osh")).setup()
    e = expected["T24k"]
    rp, ci = m.csr()
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert sha(m.boundary()) == e["sha_bnd"] and sha(m.numbering()) == e["sha_gidx"]
    assert sha(rp) == e["sha_rowptr"] and sha(ci) == e["sha_colidx"]
    u, F = m.create_global_vector(), m.create_global_vector()
    m.residual(0, u.set(0.0), F)
    assert F.norm() == pytest.approx(e["F0_norm"], rel=1e-13)
    m.destroy()


@pytest.mark.parametrize("problem", [0, 1, 2])
@pytest.mark.parametrize("name", NAMES)
def test_residual_and_jacobian_values(ctx, oracle_mod, name, problem):
    cells, coords = get_mesh(name)
    if problem != 0 and coords.shape[1] == 3 and name != "box3d_5x4x3":
        pytest.skip("variable-coefficient callbacks are 2-D formulas; one 3-D case is enough")
    bnd, gidx, N, rowptr, colidx = oracle_setup(oracle_mod, cells, coords)
    m = ohp.Mesh(ctx, cells, coords).setup()
    u, F, J = m.create_global_vector(), m.create_global_vector(), m.create_matrix()
    for uh in (np.zeros(N), probe_field(coords, gidx)):
        u.from_host(uh)
        Fd = m.residual(problem, u, F).to_host()
        Fo = oracle_mod.residual(problem, cells, coords, gidx, N, uh)
        assert np.abs(Fd - Fo).max(initial=0.0) <= VAL_RTOL * max(np.abs(Fo).max(initial=0.0), 1e-300)
        vd = m.jacobian(problem, u, J).values()
        vo = oracle_mod.jacobian(problem, cells, coords, gidx, N, uh, rowptr, colidx)
        assert rowwise_close(vd, vo, rowptr, VAL_RTOL)
        if problem == 0 and name.startswith("box2d"):
            assert np.array_equal(vd == 0.0, vo == 0.0), "explicit zeros must stay exact zeros"
    # bitwise reproducible: colour-free, atomic-free assembly
    v1 = m.jacobian(problem, u, J).values().copy()
    v2 = m.jacobian(problem, u, J).values()
    F1 = m.residual(problem, u, F).to_host().copy()
    F2 = m.residual(problem, u, F).to_host()
    assert np.array_equal(v1, v2) and np.array_equal(F1, F2)
    for o in (u, F):
        o.destroy()
    J.destroy()
    m.destroy()


@pytest.mark.parametrize("name", ["T23", "box2d_100x70", "T24k", "box3d_12x11x10", "warped3d_6x5x7"])
def test_spmv_and_vec_ops(ctx, oracle_mod, name):
    cells, coords = get_mesh(name)
    bnd, gidx, N, rowptr, colidx = oracle_setup(oracle_mod, cells, coords)
    m = ohp.Mesh(ctx, cells, coords).setup()
    J = m.create_matrix()
    u, x, y = (m.create_global_vector() for _ in range(3))
    m.jacobian(0, u, J)
    vals = J.values()
    rng = np.random.default_rng(11)
    xh, yh = rng.standard_normal(N), rng.standard_normal(N)
    x.from_host(xh)
    yd = J.mult(x, y).to_host()
    yo = oracle_mod.spmv(rowptr, colidx, vals, xh)
    assert np.abs(yd - yo).max() <= 1e-13 * np.abs(yo).max()
    y.from_host(yh)
    assert x.dot(y) == pytest.approx(float(xh @ yh), rel=1e-12, abs=1e-12)
    assert x.norm() == pytest.approx(float(np.linalg.norm(xh)), rel=1e-13)
    assert np.allclose(y.axpy(-0.75, x).to_host(), yh - 0.75 * xh, rtol=0, atol=1e-15 * 4)
    y.from_host(yh)
    assert np.allclose(y.aypx(0.5, x).to_host(), xh + 0.5 * yh, rtol=0, atol=1e-15 * 4)
    y.from_host(yh)
    ch = y.chop(0.5).to_host()
    assert np.array_equal(ch, np.where(np.abs(yh) < 0.5, 0.0, yh))
    assert np.array_equal(y.copy_from(x).to_host(), xh)
    assert x.dot(x) == x.dot(x), "deterministic reduction"
    for o in (u, x, y):
        o.destroy()
    J.destroy()
    m.destroy()


@pytest.mark.parametrize("ksp_type", [ohp.KSP_CG, ohp.KSP_PIPECG])
@pytest.mark.parametrize("pc", [0, 1])
@pytest.mark.parametrize("name", ["T23", "box2d_33x17", "sheared_21x13", "T24k", "box3d_12x11x10"])
def test_cg_matches_oracle(ctx, oracle_mod, name, pc, ksp_type):
    """-ksp_type cg and pipecg (the pipelined recurrence: same iterates in exact arithmetic) against the oracle's CG."""
    cells, coords = get_mesh(name)
    bnd, gidx, N, rowptr, colidx = oracle_setup(oracle_mod, cells, coords)
    m = ohp.Mesh(ctx, cells, coords).setup()
    J = m.create_matrix()
    u, b, x = (m.create_global_vector() for _ in range(3))
    m.jacobian(0, u, J)
    m.residual(0, u, b)
    out = ohp.ksp_cg(J, b, x, ohp.ksp_options(rtol=1e-9, max_it=20000, pc=pc, ksp_type=ksp_type), history=True)
    ref = oracle_mod.cg(rowptr, colidx, J.values(), b.to_host(), rtol=1e-9, maxit=20000, pc=pc, history=True)
    assert out["reason"] == ref["reason"] == 2
    # rounding-order sensitivity of CG; the pipelined recurrence (more recurred vectors) loses more on the ill-conditioned
    # XGC mesh: 4-7 % more iterations, depending on the form and the replacement period
    pipelined = ksp_type == ohp.KSP_PIPECG or os.environ.get("OHP_CG", "").startswith("pipecg")
    assert abs(out["its"] - ref["its"]) <= max(2, ref["its"] // (10 if pipelined else 40))  # XGC mesh: 524-539 against 506
    # CG amplifies rounding differences (summation order, FMA contraction) exponentially on an
    # ill-conditioned operator: the monitored norms agree to many digits only for the first iterations
    k = min(out["its"], ref["its"], 8)
    assert np.allclose(out["hist"][:k], ref["hist"][:k], rtol=1e-9)
    xs = x.to_host()
    assert np.abs(xs - ref["x"]).max() <= 1e-7 * np.abs(ref["x"]).max()
    # true residual meets the tolerance
    r = b.to_host() - oracle_mod.spmv(rowptr, colidx, J.values(), xs)
    assert np.linalg.norm(r) <= 2e-9 * np.linalg.norm(b.to_host()) * (1 if pc == 0 else 50)
    for o in (u, b, x):
        o.destroy()
    J.destroy()
    m.destroy()


def test_cg_limits(ctx, oracle_mod):
    cells, coords = get_mesh("T24k")
    m = ohp.Mesh(ctx, cells, coords).setup()
    J = m.create_matrix()
    u, b, x = (m.create_global_vector() for _ in range(3))
    m.jacobian(0, u, J)
    m.residual(0, u, b)
    out = ohp.ksp_cg(J, b, x, ohp.ksp_options(rtol=1e-9, max_it=37, check_every=7))
    assert (out["its"], out["reason"]) == (37, -3)
    out = ohp.ksp_cg(J, b.set(0.0), x, ohp.ksp_options(rtol=1e-9))
    assert (out["its"], out["reason"]) == (0, 3)
    # the same limits through the pipelined recurrence
    m.residual(0, u, b)
    out = ohp.ksp_cg(J, b, x, ohp.ksp_options(rtol=1e-9, max_it=37, check_every=7, ksp_type=ohp.KSP_PIPECG))
    assert (out["its"], out["reason"]) == (37, -3)
    out = ohp.ksp_cg(J, b.set(0.0), x, ohp.ksp_options(rtol=1e-9, ksp_type=ohp.KSP_PIPECG))
    assert (out["its"], out["reason"]) == (0, 3)
    m.destroy()


@pytest.mark.parametrize("name", ["T23", "box2d_33x17", "T24k", "box3d_12x11x10"])
def test_newton_matches_oracle(ctx, oracle_mod, expected, name):
    cells, coords = get_mesh(name)
    bnd, gidx, N, rowptr, colidx = oracle_setup(oracle_mod, cells, coords)
    m = ohp.Mesh(ctx, cells, coords).setup()
    J, u = m.create_matrix(), m.create_global_vector()
    m.project(0, 0, u)
    res = m.snes_solve(0, J, u, ksp=ohp.ksp_options(rtol=1e-9, max_it=100000))
    ref = oracle_mod.newton(0, cells, coords, gidx, N, rowptr, colidx, ksp_rtol=1e-9, ksp_maxit=100000)
    assert (res["its"], res["reason"]) == (ref["snes_its"], ref["reason"]) == (1, 3)
    assert (res["n_residual"], res["n_jacobian"]) == (2, 1)
    assert abs(res["ksp_its"] - ref["ksp_its"]) <= max(2, ref["ksp_its"] // 40)
    assert res["fnorm0"] == pytest.approx(ref["norms"][0], rel=1e-12)
    uh = u.to_host()
    assert np.abs(uh - ref["u"]).max() <= 1e-7 * np.abs(ref["u"]).max()
    if name in expected:
        assert np.linalg.norm(uh) == pytest.approx(expected[name]["u_norm"], rel=1e-7)
    exact = m.project(0, 1, m.create_global_vector()).to_host()
    full_exact = (coords ** 2).sum(axis=1) * (1.0 if coords.shape[1] == 2 else 2.0 / 3.0)
    assert np.allclose(exact, full_exact[gidx >= 0], rtol=4e-16, atol=0)  # x*x+y*y may contract to FMA
    m.destroy()


def test_nonlinear_newton(ctx, oracle_mod):
    cells, coords = get_mesh("box2d_33x17")
    bnd, gidx, N, rowptr, colidx = oracle_setup(oracle_mod, cells, coords)
    u0 = 0.7 * (coords ** 2).sum(axis=1)[gidx >= 0]
    m = ohp.Mesh(ctx, cells, coords).setup()
    J, u = m.create_matrix(), m.create_global_vector()
    u.from_host(u0)
    res = m.snes_solve(2, J, u, ksp=ohp.ksp_options(rtol=1e-10, max_it=5000))
    ref = oracle_mod.newton(2, cells, coords, gidx, N, rowptr, colidx, u0=u0, ksp_rtol=1e-10, ksp_maxit=5000)
    assert res["reason"] == ref["reason"] and res["its"] == ref["snes_its"] > 1
    assert np.abs(u.to_host() - ref["u"]).max() <= 1e-6 * np.abs(ref["u"]).max()
    m.destroy()


def test_error_paths(ctx):
    with pytest.raises(ohp.OhpError):
        ohp.Mesh(ctx, np.array([[0, 1, 7]], np.int32), np.zeros((3, 2)))
    m = ohp.Mesh(ctx, *get_mesh("fan4"))
    with pytest.raises(ohp.OhpError) as e:
        m.create_global_vector()
    assert e.value.code == 58
    m.setup()
    u = m.create_global_vector()
    with pytest.raises(ohp.OhpError):
        m.residual(99, u, u)
    with pytest.raises(ohp.OhpError):
        ohp.Context(1234)
    m.destroy()


def test_vertex_star_limit_is_reported(ctx):
    """A vertex with more incident cells than the pattern sort stages (512 candidate columns = 170 triangles)
    must be refused with PETSC_ERR_SUP, not mis-assembled."""
    k = 200
    ang = np.linspace(0.0, 2.0 * np.pi, k, endpoint=False)
    coords = np.vstack([[0.0, 0.0], np.stack([np.cos(ang), np.sin(ang)], axis=1), 2.0 * np.stack([np.cos(ang), np.sin(ang)], axis=1)])
    inner = np.stack([np.zeros(k, int), 1 + np.arange(k), 1 + (np.arange(k) + 1) % k], axis=1)
    a, b = 1 + np.arange(k), 1 + (np.arange(k) + 1) % k
    outer = np.vstack([np.stack([a, a + k, b], axis=1), np.stack([b, a + k, b + k], axis=1)])
    m = ohp.Mesh(ctx, np.vstack([inner, outer]).astype(np.int32), coords)
    with pytest.raises(ohp.OhpError) as e:
        m.setup()
    assert e.value.code == 56 and "star too large" in str(e.value)
    m.destroy()
    # the same fan with 150 triangles is inside the limit and must match the oracle
    import oracle
    k = 150
    ang = np.linspace(0.0, 2.0 * np.pi, k, endpoint=False)
    coords = np.vstack([[0.0, 0.0], np.stack([np.cos(ang), np.sin(ang)], axis=1), 2.0 * np.stack([np.cos(ang), np.sin(ang)], axis=1)])
    a, b = 1 + np.arange(k), 1 + (np.arange(k) + 1) % k
    cells = np.vstack([np.stack([np.zeros(k, int), a, b], axis=1), np.stack([a, a + k, b], axis=1), np.stack([b, a + k, b + k], axis=1)]).astype(np.int32)
    bnd, gidx, N, rowptr, colidx = oracle_setup(oracle, cells, coords)
    m = ohp.Mesh(ctx, cells, coords).setup()
    rp, ci = m.csr()
    assert np.array_equal(rp, rowptr) and np.array_equal(ci, colidx) and m.info().max_row == k + 1
    J, u = m.create_matrix(), m.create_global_vector()
    vo = oracle.jacobian(0, cells, coords, gidx, N, np.zeros(N), rowptr, colidx)
    assert rowwise_close(m.jacobian(0, u, J).values(), vo, rowptr, VAL_RTOL)
    m.destroy()


def test_picpart_extract_matches_reference_semantics(ctx, expected):
    """getPicPartCore* (ex12.cpp:584-730) on the 4-part XGC fixture: counts from the survey
    probe, arrays against a numpy restatement."""
    from conftest import golden_mesh_path
    for r in range(4):
        o = ohp.Osh(golden_mesh_path("24k_4p_%d.osh" % r))
        ev, X = o.elem_verts(), o.coords()
        eo, vo, gid = o.tag(2, "ownership"), o.tag(0, "ownership"), o.tag(0, "gids")
        got = ctx.picpart_extract(2, r, ev, X, eo, vo, gid)
        e = expected["T24k_4p"][str(r)]
        assert (got["n_core_elems"], got["n_core_verts"], got["n_owned_verts"], got["n_ghost_verts"]) == (
            e["owned_elems"], e["core_verts"], e["owned_verts"], e["ghosts"])
        keep = eo == r
        core = np.zeros(X.shape[0], bool)
        core[ev[keep].reshape(-1)] = True
        p2c = np.where(core, np.cumsum(core) - 1, -1)
        assert np.array_equal(got["part2core"], p2c)
        assert np.array_equal(got["core_elem_verts"], p2c[ev[keep]])
        assert np.array_equal(got["core_coords"], X[core])
        g = core & (vo != r)
        assert np.array_equal(got["ghost_core_idx"], p2c[g]) and np.array_equal(got["ghost_gid"], gid[g])
        assert np.array_equal(got["ghost_owner"], vo[g]) and np.array_equal(got["core_gid"], gid[core])
        assert (got["ghost_owner"] == r - 1).all()
        o.close()


# ----------------------------------------------------------------------------------------------
# size-independent properties at (and near) BASELINE.json's full sizes
# ----------------------------------------------------------------------------------------------
def test_large_box_properties(ctx, oracle_mod):
    """1024^2 box (N ~ 1M): pattern equals the oracle's (hash), A u + F(0) = F(u)
    (ex12.cpp:1651-1661), symmetry via x'Ay = y'Ax, explicit zeros, nodal exactness after the solve."""
    n = 1024
    cells, coords = oracle_mod.box2d(n, n)
    m = ohp.Mesh(ctx, cells, coords).setup()
    info = m.info()
    N = (n - 1) ** 2
    assert (info.n_owned, info.n_boundary_verts, info.nnz) == (N, 4 * n, N + 2 * (2 * (n - 1) * (n - 2) + (n - 2) ** 2))
    bnd = oracle_mod.mark_boundary(cells, coords.shape[0], omp=True)
    gidx, _ = oracle_mod.numbering(bnd)
    rowptr, colidx = oracle_mod.pattern(cells, gidx, N, omp=True)
    rp, ci = m.csr()
    assert np.array_equal(rp, rowptr) and np.array_equal(ci, colidx)
    J = m.create_matrix()
    u, F0, Fu, Au, y = (m.create_global_vector() for _ in range(5))
    rng = np.random.default_rng(5)
    uh, yh = rng.standard_normal(N), rng.standard_normal(N)
    m.residual(0, u.set(0.0), F0)
    m.jacobian(0, u, J)
    u.from_host(uh)
    y.from_host(yh)
    m.residual(0, u, Fu)
    J.mult(u, Au)
    # direct SpMV parity at a size where CTAs own many tiles and rows straddle tile boundaries
    Ao = oracle_mod.spmv(rowptr, colidx, J.values(), uh, omp=True)
    assert np.abs(Au.to_host() - Ao).max() <= 1e-13 * np.abs(Ao).max()
    lin = Au.to_host() + F0.to_host() - Fu.to_host()
    assert np.abs(lin).max() <= 1e-11 * np.abs(Fu.to_host()).max()
    yAu = y.dot(Au)
    J.mult(y, Au)
    assert u.dot(Au) == pytest.approx(yAu, rel=1e-11)
    vals = J.values()
    assert (vals == 0.0).sum() == 2 * (n - 2) ** 2, "hypotenuse couplings: exact zeros kept in the pattern"
    res = m.snes_solve(0, J, u.set(0.0), ksp=ohp.ksp_options(rtol=1e-12, max_it=100000))
    assert res["reason"] == 3 and res["its"] == 1
    exact = (coords ** 2).sum(axis=1)[gidx >= 0]
    assert np.abs(u.to_host() - exact).max() < 1e-8
    m.destroy()


def test_full_size_16M_triangles(ctx):
    """BASELINE config 4 (2896 x 2896 box, 16 773 632 triangles): counts in closed form, linearity of
    the assembled operator, symmetry, and CG residual reduction over a fixed number of iterations."""
    n = 2896
    import oracle  # input generator only
    cells, coords = oracle.box2d(n, n)
    m = ohp.Mesh(ctx, cells, coords).setup()
    info = m.info()
    N = (n - 1) ** 2
    assert info.n_cells == 16773632 and info.n_verts == 8392609
    assert (info.n_owned, info.n_boundary_verts, info.nnz, info.max_row) == (N, 4 * n, 58644017, 7)
    del cells
    J = m.create_matrix()
    u, F0, Fu, Au, y = (m.create_global_vector() for _ in range(5))
    rng = np.random.default_rng(9)
    m.residual(0, u.set(0.0), F0)
    m.jacobian(0, u, J)
    u.from_host(rng.standard_normal(N))
    y.from_host(rng.standard_normal(N))
    m.residual(0, u, Fu)
    J.mult(u, Au)
    Au.axpy(1.0, F0).axpy(-1.0, Fu)
    assert Au.norm() <= 1e-11 * Fu.norm()
    J.mult(u, Au)
    yAu = y.dot(Au)
    J.mult(y, Au)
    assert u.dot(Au) == pytest.approx(yAu, rel=1e-10)
    out = ohp.ksp_cg(J, F0, u, ohp.ksp_options(rtol=1e-9, max_it=300), history=True)
    assert out["reason"] == -3 and out["its"] == 300
    J.mult(u, Au)
    Au.aypx(-1.0, F0)  # r = b - A x
    assert Au.norm() == pytest.approx(out["rnorm"], rel=1e-6)
    # the oracle itself at FULL size (OpenMP): numbering and pattern bit-exact, F(0) and J(0) to 1e-12, 300 CG iterations
    full_size_oracle_compare(m, J, cells_coords=(oracle.box2d(n, n)), cg_hist=out["hist"])
    m.destroy()


def full_size_oracle_compare(m, J, cells_coords, cg_hist=None, exact_zeros=True):
    import oracle
    oracle.use_all_cores()
    cells, coords = cells_coords
    bnd = oracle.mark_boundary(cells, coords.shape[0], omp=True)
    gidx, N = oracle.numbering(bnd)
    rowptr, colidx = oracle.pattern(cells, gidx, N, omp=True)
    assert np.array_equal(m.boundary(), bnd) and np.array_equal(m.numbering(), gidx.astype(np.int64))
    rp, ci = m.csr()
    assert np.array_equal(rp, rowptr) and np.array_equal(ci, colidx.astype(np.int64))
    u0, F0 = m.create_global_vector(), m.create_global_vector()
    m.residual(0, u0.set(0.0), F0)
    m.jacobian(0, u0, J)
    Fo = oracle.residual(0, cells, coords, gidx, N, np.zeros(N), omp=True)
    vo = oracle.jacobian(0, cells, coords, gidx, N, np.zeros(N), rowptr, colidx, omp=True)
    assert np.abs(F0.to_host() - Fo).max() <= VAL_RTOL * np.abs(Fo).max()
    vd = J.values()
    assert rowwise_close(vd, vo, rowptr, VAL_RTOL)
    if exact_zeros:  # 2-D same-diagonal box: every hypotenuse coupling is a sum of two exact zeros (SURVEY 8c anchor 1)
        assert np.array_equal(vd == 0.0, vo == 0.0)
    if cg_hist is not None:
        ref = oracle.cg(rowptr, colidx, vo, Fo, rtol=1e-9, maxit=len(cg_hist) - 1, omp=True, history=True)
        assert np.allclose(cg_hist, ref["hist"], rtol=1e-6), "residual history of the first iterations"
    u0.destroy(); F0.destroy()


def test_persistent_cg_kernel_small_mesh(ctx, oracle_mod):
    """The persistent single-reduction kernel (auto-selected below 400k rows) against the oracle on the
    24k XGC mesh: same converged solution, iteration count within the rounding-order band, history."""
    cells, coords = get_mesh("T24k")
    bnd, gidx, N, rowptr, colidx = oracle_setup(oracle_mod, cells, coords)
    m = ohp.Mesh(ctx, cells, coords).setup()
    J = m.create_matrix()
    u, b, x = (m.create_global_vector() for _ in range(3))
    m.jacobian(0, u, J)
    m.residual(0, u, b)
    for pc in (0, 1):
        out = ohp.ksp_cg(J, b, x, ohp.ksp_options(rtol=1e-9, max_it=20000, pc=pc, check_every=64), history=True)
        ref = oracle_mod.cg(rowptr, colidx, J.values(), b.to_host(), rtol=1e-9, maxit=20000, pc=pc, history=True)
        assert out["reason"] == 2 and abs(out["its"] - ref["its"]) <= ref["its"] // 40
        assert np.allclose(out["hist"][:6], ref["hist"][:6], rtol=1e-9)
        assert np.abs(x.to_host() - ref["x"]).max() <= 1e-7 * np.abs(ref["x"]).max()
    out = ohp.ksp_cg(J, b, x, ohp.ksp_options(rtol=1e-9, max_it=41, check_every=7))
    assert (out["its"], out["reason"]) == (41, -3)
    m.destroy()


def test_full_size_3d_8M_tets(ctx):
    """BASELINE config 5 (110^3 x 6 Kuhn tets = 7 986 000 tets): closed-form counts, linearity, symmetry,
    and the manufactured solution u = 2/3 r^2 reproduced at the nodes after the Newton solve."""
    n = 110
    from omegahpetsc_b200 import meshgen
    cells, coords = meshgen.box3d(n, n, n)
    m = ohp.Mesh(ctx, cells, coords).setup()
    info = m.info()
    N = (n - 1) ** 3
    assert (info.n_cells, info.n_verts, info.n_owned) == (7986000, 1367631, N)
    assert info.n_boundary_verts == (n + 1) ** 3 - N and info.max_row == 15
    J = m.create_matrix()
    u, F0, Fu, Au, y = (m.create_global_vector() for _ in range(5))
    rng = np.random.default_rng(4)
    m.residual(0, u.set(0.0), F0)
    m.jacobian(0, u, J)
    u.from_host(rng.standard_normal(N))
    y.from_host(rng.standard_normal(N))
    m.residual(0, u, Fu)
    J.mult(u, Au)
    yAu = y.dot(Au)
    Au.axpy(1.0, F0).axpy(-1.0, Fu)
    assert Au.norm() <= 1e-11 * Fu.norm()
    J.mult(y, Au)
    assert u.dot(Au) == pytest.approx(yAu, rel=1e-10)
    res = m.snes_solve(0, J, u.set(0.0), ksp=ohp.ksp_options(rtol=1e-10, max_it=100000))
    assert res["reason"] == 3 and res["its"] == 1
    exact = m.project(0, 1, m.create_global_vector()).to_host()
    assert np.abs(u.to_host() - exact).max() < 1e-7
    # (3-D Kuhn tets: the zero couplings are cancellations ACROSS cells, so whether they come out as 0.0 or 1e-18
    #  depends on the summation order; they are covered by the 1e-12 row-relative bound)
    full_size_oracle_compare(m, J, (cells, coords), exact_zeros=False)
    m.destroy()


# ------------------------------------------------------------------------------------------------
# round 2: circle / cross problems, the problem's own exact solution, Plex facets, caller-set labels,
# orderings whose column offsets do not fit 16 bits, monitors
# ------------------------------------------------------------------------------------------------
def exact_of(problem, coords):
    x, y = coords[:, 0], coords[:, 1]
    if problem == 3:
        return np.tanh(500.0 * (0.15 ** 2 - ((x - 0.5) ** 2 + (y - 0.5) ** 2))) + 1.0
    if problem == 4:
        xy = (x - 0.5) * (y - 0.5)
        return np.sin(200.0 * xy) * np.where(200.0 * np.abs(xy) < 2 * np.pi, 1.0, 0.01)
    return (coords ** 2).sum(axis=1)


def numpy_l2_error(oracle, problem, cells, coords, gidx, u):
    xi, w = oracle.quadrature(2)
    full = exact_of(problem, coords)
    full[gidx >= 0] = u
    X = coords[cells]
    J = 0.5 * np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=2)
    detJ = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
    err = 0.0
    for q in range(4):
        phi = np.array([-(xi[q, 0] + xi[q, 1]) / 2, (1 + xi[q, 0]) / 2, (1 + xi[q, 1]) / 2])
        xq = X[:, 0] + J @ (xi[q] + 1.0)
        err += (w[q] * np.abs(detJ) * (full[cells] @ phi - exact_of(problem, xq)) ** 2).sum()
    return float(np.sqrt(err))


@pytest.mark.parametrize("problem", [3, 4])
@pytest.mark.parametrize("name", ["box2d_100x70", "sheared_21x13", "T24k"])
def test_circle_and_cross_forcing(ctx, oracle_mod, name, problem):
    """f0_circle_u / f0_cross_u with their own Dirichlet data (ex12.cpp:127-177, selected at :1197-1213)."""
    cells, coords = get_mesh(name)
    bnd, gidx, N, rowptr, colidx = oracle_setup(oracle_mod, cells, coords)
    m = ohp.Mesh(ctx, cells, coords).setup()
    J, u, F = m.create_matrix(), m.create_global_vector(), m.create_global_vector()
    uh = probe_field(coords, gidx)
    u.from_host(uh)
    Fo = oracle_mod.residual(problem, cells, coords, gidx, N, uh)
    vo = oracle_mod.jacobian(problem, cells, coords, gidx, N, uh, rowptr, colidx)
    assert np.abs(m.residual(problem, u, F).to_host() - Fo).max() <= VAL_RTOL * np.abs(Fo).max()
    assert rowwise_close(m.jacobian(problem, u, J).values(), vo, rowptr, VAL_RTOL)
    res = m.snes_solve(problem, J, u.set(0.0), ksp=ohp.ksp_options(rtol=1e-10, max_it=100000))
    ref = oracle_mod.newton(problem, cells, coords, gidx, N, rowptr, colidx, ksp_rtol=1e-10, ksp_maxit=100000)
    # on the XGC coordinates (r > 1) sech^2 underflows: F(0) == 0 exactly and SNESConvergedDefault stops at
    # iteration 0 with FNORM_ABS, in the oracle as well as here
    assert (res["its"], res["reason"]) == (ref["snes_its"], ref["reason"])
    assert (res["its"], res["reason"]) in ((1, 3), (0, 2))
    assert np.abs(u.to_host() - ref["u"]).max() <= 1e-7 * max(1.0, np.abs(ref["u"]).max())
    # DMComputeL2Diff uses THIS problem's exact solution (it used to be the quadratic for every problem)
    assert m.l2_error(problem, u) == pytest.approx(numpy_l2_error(oracle_mod, problem, cells, coords, gidx, ref["u"]), rel=1e-6)
    assert np.abs(m.project(problem, 1, F).to_host() - exact_of(problem, coords)[gidx >= 0]).max() <= 1e-14
    m.destroy()


@pytest.mark.parametrize("name", ["one_tri", "fan4", "T23", "box2d_33x17", "T24k", "box3d_5x4x3", "warped3d_6x5x7"])
def test_facets_like_plex_interpolate(ctx, name):
    """DMPlexGetHeightStratum(1) / GetSupportSize / GetCone (ex12.cpp:966-1010) against a numpy construction."""
    cells, coords = get_mesh(name)
    nc = cells.shape[1]
    loc = [(0, 1), (1, 2), (2, 0)] if nc == 3 else [(0, 1, 2), (0, 3, 1), (0, 2, 3), (2, 1, 3)]
    first, support, order = {}, {}, []
    for c, cv in enumerate(cells):
        for f in loc:
            key = tuple(sorted(int(cv[k]) for k in f))
            if key not in first:
                first[key] = (len(order), tuple(int(cv[k]) for k in f))
                order.append(key)
            support[key] = support.get(key, 0) + 1
    m = ohp.Mesh(ctx, cells, coords)
    fv, sup, cf = m.facets()
    assert fv.shape[0] == len(order)
    for i, key in enumerate(order):        # numbered by creating cell, then local facet; cone in that cell's orientation
        assert tuple(fv[i]) == first[key][1] and sup[i] == support[key]
    for c, cv in enumerate(cells):
        for j, f in enumerate(loc):
            assert cf[c, j] == first[tuple(sorted(int(cv[k]) for k in f))][0]
    # facets with one supporting cell carry exactly the boundary vertices (ex12.cpp:986-1010)
    m.setup()
    bverts = np.zeros(coords.shape[0], np.uint8)
    bverts[np.unique(fv[sup == 1])] = 1
    assert np.array_equal(bverts, m.boundary())
    m.destroy()


def test_label_set_by_the_caller(ctx, oracle_mod):
    """DMSetLabelValue point by point (ex12.cpp:1023-1026): the caller's label replaces the library's detection."""
    cells, coords = get_mesh("box2d_33x17")
    bnd = oracle_mod.mark_boundary(cells, coords.shape[0])
    extra = np.nonzero(bnd == 0)[0][[5, 77, 200]]
    lab = bnd.copy()
    lab[extra] = 1
    gidx, N = oracle_mod.numbering(lab)
    rowptr, colidx = oracle_mod.pattern(cells, gidx, N)
    m = ohp.Mesh(ctx, cells, coords)
    m.set_label(np.nonzero(lab)[0])
    m.setup()
    assert np.array_equal(m.boundary(), lab) and np.array_equal(m.numbering(), gidx.astype(np.int64))
    rp, ci = m.csr()
    assert np.array_equal(rp, rowptr) and np.array_equal(ci, colidx.astype(np.int64))
    with pytest.raises(ohp.OhpError):
        m.set_label([0])  # after setup
    m.destroy()


def test_morton_ordering_uses_the_32bit_column_stream(ctx, oracle_mod):
    """A space-filling-curve ordering (what Omega_h::build_box produces, ex12.cpp:805): |column - row| no longer fits
    16 bits everywhere, so the SpMV streams the 32-bit column ids.  Same parity bar as every mesh."""
    from omegahpetsc_b200 import meshgen
    n = 700
    c0, x0 = meshgen.box2d(n, n)
    cells, coords = meshgen.morton_reorder(c0, x0, (n, n))
    bnd, gidx, N, rowptr, colidx = oracle_setup(oracle_mod, cells, coords)
    rows = np.repeat(np.arange(N), np.diff(rowptr))
    far = np.abs(colidx - rows) > 32767
    assert far.any(), "this mesh must have offsets beyond 16 bits"
    m = ohp.Mesh(ctx, cells, coords).setup()
    rp, ci = m.csr()
    assert np.array_equal(m.numbering(), gidx.astype(np.int64)) and np.array_equal(rp, rowptr) and np.array_equal(ci, colidx.astype(np.int64))
    J, u, y = m.create_matrix(), m.create_global_vector(), m.create_global_vector()
    m.jacobian(0, u, J)
    xh = np.random.default_rng(3).standard_normal(N)
    J.mult(u.from_host(xh), y)
    yo = oracle_mod.spmv(rowptr, colidx, J.values(), xh)
    assert np.abs(y.to_host() - yo).max() <= 1e-13 * np.abs(yo).max()
    res = m.snes_solve(0, J, u.set(0.0), ksp=ohp.ksp_options(rtol=1e-10, max_it=100000))
    assert res["reason"] == 3 and np.abs(u.to_host() - (coords ** 2).sum(axis=1)[gidx >= 0]).max() < 1e-7
    m.destroy()


def test_monitors_report_every_iteration(ctx, oracle_mod):
    """-snes_monitor / -ksp_monitor (CMakeLists.txt:56-57): norms per Newton iteration and per CG iteration."""
    cells, coords = get_mesh("box2d_100x70")
    bnd, gidx, N, rowptr, colidx = oracle_setup(oracle_mod, cells, coords)
    m = ohp.Mesh(ctx, cells, coords).setup()
    J, u = m.create_matrix(), m.create_global_vector()
    mon = {}
    # nonlinear problem (nu = |grad u|^2-dependent, ex12.cpp:292-383): start away from u = 0, where J is singular
    u.from_host(0.7 * (coords ** 2).sum(axis=1)[gidx >= 0])
    res = m.snes_solve(2, J, u, ksp=ohp.ksp_options(rtol=1e-9, max_it=100000), monitor=mon)
    assert res["reason"] > 0 and res["its"] >= 2
    assert [it for it, _ in mon["snes"]] == list(range(res["its"] + 1))
    assert mon["snes"][0][1] == pytest.approx(res["fnorm0"]) and mon["snes"][-1][1] == pytest.approx(res["fnorm"])
    assert len(mon["ksp"]) == res["its"] and sum(len(h) - 1 for h in mon["ksp"]) == res["ksp_its"]
    assert all(r == 2 for r in mon["ksp_reason"]) and all(h[-1] <= 1e-9 * h[0] for h in mon["ksp"])
    m.destroy()


def test_patch_assembly_path_opt_in():
    """OHP_ASM=patch: the cell-parallel assembly (every cell computed once per patch, csrc/ohp_patch.cu + k_assemble_patch)
    passes the same residual / Jacobian parity cases as the default row kernels.  The variable is read once per process,
    so the cases run in a child pytest."""
    import subprocess
    import sys
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x",
                        "-k", "residual_and_jacobian_values or newton_matches_oracle"],
                       env=dict(os.environ, OHP_ASM="patch"), capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]


@pytest.mark.parametrize("variant", ["classic", "cgcg", "pipecg1", "pipecg2", "persistent"])
def test_every_cg_recurrence_on_one_gpu(variant):
    """The library picks the recurrence from the problem size and the transport, so on one GPU the default suite only meets
    some of them.  OHP_CG forces each one (classic three-launch, single-reduction two-launch, pipelined one- and two-launch --
    the latter is the 8-GPU default of the bench -- and the persistent cooperative kernel) through the CG parity cases:
    oracle iteration counts, residual histories, solutions, true residuals, iteration limits.  The variable is read once per
    process, so the cases run in a child pytest."""
    import subprocess
    import sys
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x",
                        "-k", "cg_matches_oracle or cg_limits"],
                       env=dict(os.environ, OHP_CG=variant), capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
