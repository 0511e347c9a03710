"""GPU parity tests (-m gpu) of the SURVEY 8(f) rank-4 slice: the natural (-bc_type neumann) boundary condition, the
constant null space and the Chebyshev/Jacobi polynomial preconditioner, through the C-ABI against the oracle.
(The file sorts last on purpose: the hot-path parity tests run first.)
Tolerances: numbering / pattern bit-exact; assembled values 1e-12 of the largest magnitude; iteration counts of the
general CG equal to the oracle's up to rounding-order effects (+-2 % or 2); solutions within the solver tolerance."""
import os
import subprocess

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


def get_mesh(name):
    if name == "T24k":
        c, x, _ = load_osh_oracle("24k.osh")
        return c, x
    return synthetic_meshes()[name]


NAT_NAMES = ["one_tri", "fan4", "box2d_8x8", "box2d_33x17", "sheared_21x13", "T24k", "box3d_5x4x3", "warped3d_6x5x7"]


def probe(coords):
    return np.sin(3.0 * coords[:, 0]) * np.cos(2.0 * coords[:, 1]) + (0.3 * coords[:, 2] if coords.shape[1] == 3 else 0.0)


@pytest.mark.parametrize("name", NAT_NAMES)
def test_natural_bc_section_pattern_and_assembly(ctx, oracle_mod, name):
    """DM_BC_NATURAL: no dof is constrained (N = nV, the numbering is the identity), the pattern is the full vertex
    adjacency, F carries the boundary integral of f0_bd_u / f1_bd_zero, J is the unconstrained operator."""
    cells, coords = get_mesh(name)
    nV = coords.shape[0]
    bnd = oracle_mod.mark_boundary(cells, nV)
    gidx = np.arange(nV, dtype=np.int32)
    rowptr, colidx = oracle_mod.pattern(cells, gidx, nV)
    m = ohp.Mesh(ctx, cells, coords).set_bc_type(ohp.BC_NATURAL).setup()
    info = m.info()
    assert (info.n_owned, info.n_global, info.nnz, info.n_boundary_verts) == (nV, nV, colidx.size, int(bnd.sum()))
    assert np.array_equal(m.boundary(), bnd)  # the label is still there: it selects the rows that integrate facets
    assert np.array_equal(m.numbering(), gidx.astype(np.int64))
    rp, ci = m.csr()
    assert np.array_equal(rp, rowptr) and np.array_equal(ci, colidx.astype(np.int64))
    u, F, J = m.create_global_vector(), m.create_global_vector(), m.create_matrix()
    for uh in (np.zeros(nV), probe(coords)):
        u.from_host(uh)
        want = oracle_mod.residual(0, cells, coords, gidx, nV, uh)
        vol = want.copy()
        want = oracle_mod.bd_residual(cells, coords, gidx, nV, uh, want)
        got = m.residual(0, u, F).to_host()
        assert np.abs(got - want).max() <= VAL_RTOL * max(np.abs(want).max(), 1e-300)
        assert np.abs(want - vol).max() > 0  # the boundary term is really in there
        # bitwise reproducible
        assert np.array_equal(m.residual(0, u, F).to_host(), got)
        vals = m.jacobian(0, u, J).values()
        ref = oracle_mod.jacobian(0, cells, coords, gidx, nV, uh, rowptr, colidx)
        rowmax = np.maximum.reduceat(np.abs(ref), rowptr[:-1])
        assert (np.abs(vals - ref) <= VAL_RTOL * np.repeat(rowmax, np.diff(rowptr))).all()
    # the constants are in the kernel of the unconstrained operator
    one, y = m.create_global_vector().set(1.0), m.create_global_vector()
    assert np.abs(J.mult(one, y).to_host()).max() <= 1e-12 * np.abs(vals).max()
    m.destroy()


def test_bc_none_integrates_nothing_and_order_errors(ctx, oracle_mod):
    cells, coords = get_mesh("box2d_8x8")
    nV = coords.shape[0]
    gidx = np.arange(nV, dtype=np.int32)
    m = ohp.Mesh(ctx, cells, coords).set_bc_type(ohp.BC_NONE).setup()
    u, F = m.create_global_vector().set(0.0), m.create_global_vector()
    want = oracle_mod.residual(0, cells, coords, gidx, nV, np.zeros(nV))
    assert np.abs(m.residual(0, u, F).to_host() - want).max() <= VAL_RTOL * np.abs(want).max()
    with pytest.raises(ohp.OhpError) as e:
        m.set_bc_type(ohp.BC_NATURAL)  # after setup
    assert e.value.code == 58
    m.destroy()
    m2 = ohp.Mesh(ctx, cells, coords)
    with pytest.raises(ohp.OhpError) as e:
        m2.set_bc_type(7)
    assert e.value.code == 63
    m2.destroy()


def _dirichlet_system(ctx, oracle_mod, name):
    cells, coords = get_mesh(name)
    bnd = oracle_mod.mark_boundary(cells, coords.shape[0])
    gidx, N = oracle_mod.numbering(bnd)
    rowptr, colidx = oracle_mod.pattern(cells, gidx, N)
    vals = oracle_mod.jacobian(0, cells, coords, gidx, N, np.zeros(N), rowptr, colidx)
    F0 = oracle_mod.residual(0, cells, coords, gidx, N, np.zeros(N))
    m = ohp.Mesh(ctx, cells, coords).setup()
    u, F, J = m.create_global_vector().set(0.0), m.create_global_vector(), m.create_matrix()
    m.residual(0, u, F)
    m.jacobian(0, u, J)
    return m, J, F, (rowptr, colidx, vals, F0)


def _close_its(a, b):
    return abs(a - b) <= max(2, int(0.02 * b))


@pytest.mark.parametrize("name", ["box2d_33x17", "box2d_100x70", "T24k", "box3d_12x11x10"])
def test_chebyshev_preconditioned_cg_matches_oracle(ctx, oracle_mod, name):
    m, J, F, (rowptr, colidx, vals, F0) = _dirichlet_system(ctx, oracle_mod, name)
    x = m.create_global_vector()
    base = ohp.ksp_cg(J, F, x, ohp.ksp_options(rtol=1e-9, pc=ohp.PC_JACOBI))
    for k, emin, emax in ((1, 0.0, 0.0), (2, 0.0, 0.0), (4, 0.0, 0.0), (8, 0.0, 0.0), (4, 0.2, 2.2)):
        want = oracle_mod.pcg(rowptr, colidx, vals, F0, rtol=1e-9, pc=2, cheb_its=k, emin=emin, emax=emax, history=True)
        got = ohp.ksp_cg(J, F, x, ohp.ksp_options(rtol=1e-9, pc=ohp.PC_CHEBYSHEV, cheb_its=k, cheb_emin=emin, cheb_emax=emax), history=True)
        assert got["algo"] == ohp.ALGO_GENERAL_PCG and got["reason"] == want["reason"] == 2
        assert _close_its(got["its"], want["its"]), (k, got["its"], want["its"])
        assert abs(got["cheb_emax"] - want["emax"]) <= 1e-8 * want["emax"] and abs(got["cheb_emin"] - want["emin"]) <= 1e-8 * want["emax"]
        xs = np.abs(want["x"]).max()
        assert np.abs(x.to_host() - want["x"]).max() <= 1e-6 * xs
        n = min(got["its"], want["its"], 10) + 1  # the first norms agree to rounding (later ones drift with the summation order)
        assert np.abs(got["hist"][:n] - want["hist"][:n]).max() <= 1e-9 * want["hist"][0]
        est = 10 if emax == 0.0 else 0
        assert got["spmvs"] == est + got["its"] * k + (k - 1)
        if k >= 2 and emax == 0.0:
            assert got["its"] < base["its"]  # fewer outer iterations (global reductions) than Jacobi-preconditioned CG
    # option errors
    with pytest.raises(ohp.OhpError):
        ohp.ksp_cg(J, F, x, ohp.ksp_options(pc=3))
    m.destroy()


def test_null_space_solve_matches_oracle(ctx, oracle_mod):
    """-bc_type neumann on the box: singular operator, consistent right-hand side; CG with the constants projected out"""
    cells, coords = get_mesh("box2d_33x17")
    nV = coords.shape[0]
    gidx = np.arange(nV, dtype=np.int32)
    rowptr, colidx = oracle_mod.pattern(cells, gidx, nV)
    vals = oracle_mod.jacobian(0, cells, coords, gidx, nV, np.zeros(nV), rowptr, colidx)
    F0 = oracle_mod.bd_residual(cells, coords, gidx, nV, np.zeros(nV), oracle_mod.residual(0, cells, coords, gidx, nV, np.zeros(nV)))
    m = ohp.Mesh(ctx, cells, coords).set_bc_type(ohp.BC_NATURAL).setup()
    u, F, J, x = m.create_global_vector().set(0.0), m.create_global_vector(), m.create_matrix(), m.create_global_vector()
    m.residual(0, u, F)
    m.jacobian(0, u, J)
    J.set_nullspace(True)
    for pc, k in ((ohp.PC_NONE, 0), (ohp.PC_JACOBI, 0), (ohp.PC_CHEBYSHEV, 4)):
        want = oracle_mod.pcg(rowptr, colidx, vals, F0, rtol=1e-9, pc=pc, cheb_its=k, nullspace=True)
        got = ohp.ksp_cg(J, F, x, ohp.ksp_options(rtol=1e-9, pc=pc, cheb_its=k))
        assert got["algo"] == ohp.ALGO_GENERAL_PCG and got["reason"] == want["reason"] == 2
        assert _close_its(got["its"], want["its"])
        xd, xo = x.to_host(), want["x"]
        assert np.abs((xd - xd.mean()) - (xo - xo.mean())).max() <= 1e-6 * np.abs(xo - xo.mean()).max()
    # MatNullSpaceRemove
    x.from_host(probe(coords) + 5.0).remove_constant()
    assert abs(x.to_host().mean()) < 1e-13
    assert np.abs(x.to_host() - (probe(coords) - probe(coords).mean())).max() < 1e-12
    m.destroy()


def test_newton_with_natural_bc(ctx, oracle_mod):
    """-run_type full -bc_type neumann: Newton over the null-space CG; the solution is the manufactured one up to a
    constant (P1 reproduces the nodal values of x^2 + y^2 on the same-diagonal box up to the solver tolerance)"""
    cells, coords = oracle_mod.box2d(48, 40)
    m = ohp.Mesh(ctx, cells, coords).set_bc_type(ohp.BC_NATURAL).setup()
    u, J = m.create_global_vector().set(0.0), m.create_matrix()
    J.set_nullspace(True)
    r = m.snes_solve(0, J, u, ksp=ohp.ksp_options(rtol=1e-10, pc=ohp.PC_JACOBI))
    assert r["reason"] in (2, 3, 4) and r["algo"] == ohp.ALGO_GENERAL_PCG
    uh = u.to_host()
    exact = (coords ** 2).sum(axis=1)
    err = (uh - uh.mean()) - (exact - exact.mean())
    assert np.abs(err).max() < 5e-3  # O(h^2): the natural condition is only weakly imposed, no nodal exactness here
    F = m.create_global_vector()
    assert m.residual(0, u, F).norm() <= 1e-7 * r["fnorm0"]
    m.destroy()


def _example():
    exe = os.path.join(ROOT, "examples", "ex12_b200")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True)
    return exe


def _run(args):
    p = subprocess.run([_example()] + args, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def test_facade_neumann_and_polynomial_preconditioner(ctx, oracle_mod):
    """the C++ drop-in: ex12's -bc_type neumann -run_type test|full, and -pc_type ksp (Chebyshev/Jacobi) on the Dirichlet run"""
    cells, coords = oracle_mod.box2d(24, 20)
    nV = coords.shape[0]
    gidx = np.arange(nV, dtype=np.int32)
    uex = (coords ** 2).sum(axis=1)
    F = oracle_mod.bd_residual(cells, coords, gidx, nV, uex, oracle_mod.residual(0, cells, coords, gidx, nV, uex))
    F[np.abs(F) < 1e-10] = 0.0
    out = _run(["-mesh", "box", "-cells", "24,20", "-bc_type", "neumann", "-run_type", "test"])
    vals = {}
    for line in out.splitlines():
        key, sep, val = line.partition(":")
        if sep and ("Residual" in key or "Error" in key):
            vals[key.strip()] = float(val.replace("<", ""))
    assert vals["L_2 Residual"] == pytest.approx(np.linalg.norm(F), rel=1e-5)  # printed with %g
    # "Au - b = Au + F(0)" (ex12.cpp:1655-1661) is F(u) again for this linear problem
    assert vals["Linear L_2 Residual"] == pytest.approx(vals["L_2 Residual"], rel=1e-5)
    assert 0 < vals["L_2 Error"] < 1e-2  # interpolation error of the quadratic
    out = _run(["-mesh", "box", "-cells", "24,20", "-bc_type", "neumann", "-run_type", "full", "-ksp_rtol", "1e-10", "-pc_type", "jacobi",
                "-snes_converged_reason"])
    assert "CONVERGED" in out and "N: %d " % nV in out
    # Dirichlet run with the polynomial preconditioner: same answer as plain CG, fewer outer iterations
    plain = _run(["-mesh", "box", "-cells", "64,64", "-ksp_rtol", "1e-9", "-pc_type", "jacobi"])
    poly = _run(["-mesh", "box", "-cells", "64,64", "-ksp_rtol", "1e-9", "-pc_type", "ksp", "-ksp_ksp_type", "chebyshev", "-ksp_pc_type", "jacobi",
                 "-ksp_ksp_max_it", "4"])
    get = lambda s, key: float(s.split(key)[1].split()[0])
    assert get(poly, "CG its:") < 0.5 * get(plain, "CG its:")
    assert abs(get(poly, "L_2 Error:") - get(plain, "L_2 Error:")) < 1e-6  # nodally exact up to the solver tolerance, both
    bad = subprocess.run([_example(), "-mesh", "box", "-pc_type", "ksp", "-ksp_ksp_type", "gmres"], capture_output=True, text=True, timeout=600)
    assert bad.returncode == 56


def test_recorded_values_T24k(ctx):
    """the recorded values of tests/golden/expected_next_rows.json on the reference's XGC fixture, read from its .osh file"""
    import hashlib
    import json
    from conftest import GOLDEN, golden_mesh_path
    with open(os.path.join(GOLDEN, "expected_next_rows.json")) as f:
        e = json.load(f)["T24k"]
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    m = ohp.Mesh.from_osh(ctx, golden_mesh_path("24k.osh")).set_bc_type(ohp.BC_NATURAL).setup()
    rp, ci = m.csr()
    assert (m.info().n_owned, m.info().nnz) == (e["natural_N"], e["natural_nnz"])
    assert sha(rp) == e["natural_sha_rowptr"] and sha(ci) == e["natural_sha_colidx"]
    u, F = m.create_global_vector().set(0.0), m.create_global_vector()
    assert m.residual(0, u, F).norm() == pytest.approx(e["natural_F0_norm"], rel=1e-11)
    m.destroy()
    m = ohp.Mesh.from_osh(ctx, golden_mesh_path("24k.osh")).setup()
    u, F, J, x = m.create_global_vector().set(0.0), m.create_global_vector(), m.create_matrix(), m.create_global_vector()
    m.residual(0, u, F)
    m.jacobian(0, u, J)
    for k in (4, 16):
        g = e["chebyshev_k%d" % k]
        r = ohp.ksp_cg(J, F, x, ohp.ksp_options(rtol=1e-9, pc=ohp.PC_CHEBYSHEV, cheb_its=k))
        assert r["reason"] == g["reason"] and _close_its(r["its"], g["its"])
        assert r["cheb_emax"] == pytest.approx(g["emax"], rel=1e-8)
        assert x.norm() == pytest.approx(g["x_norm"], rel=1e-7)
    m.destroy()


def test_hilbert_ordering_like_build_box(ctx, oracle_mod):
    """The ordering Omega_h::build_box itself produces (ex12.cpp:805: vertices sorted along a Hilbert curve): the same bar as
    the Morton-ordered mesh of test_gpu_parity.py -- bit-exact numbering and pattern, SpMV on the 32-bit column stream, the
    Newton solve nodally exact."""
    from omegahpetsc_b200 import meshgen
    n = 700
    c0, x0 = meshgen.box2d(n, n)
    cells, coords = meshgen.hilbert_reorder(c0, x0, (n, n))
    bnd = oracle_mod.mark_boundary(cells, coords.shape[0])
    gidx, N = oracle_mod.numbering(bnd)
    rowptr, colidx = oracle_mod.pattern(cells, gidx, N)
    rows = np.repeat(np.arange(N), np.diff(rowptr))
    assert (np.abs(colidx - rows) > 32767).any(), "this mesh must have offsets beyond 16 bits"
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
