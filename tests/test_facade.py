"""The PETSc-shaped C++ facade (include/ohp_petsc.h, include/ohp_omega_h.h) through the ex12-style
driver examples/ex12_b200.cu: user callbacks compiled by nvcc, registered with OHP_DEFINE_PROBLEM,
found again from the host function pointers PetscDSSetResidual/SetJacobian/DMAddBoundary receive."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_mesh_path, load_osh_oracle

EXE = os.path.join(ROOT, "examples", "ex12_b200")


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "omegahpetsc_b200", "csrc"), "-j8"], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "examples")], check=True)


def test_example_builds_with_user_callbacks_inlined():
    build()
    out = subprocess.run(["cuobjdump", "-elf", "-lelf", EXE], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sym = subprocess.run("cuobjdump -symbols %s | grep -c ohp_problem_constant" % EXE, shell=True, capture_output=True, text=True).stdout
    assert int(sym.strip() or 0) >= 2, "assembly kernels must be instantiated on the user's callback set"


def oracle_l2_error(oracle, cells, coords, gidx, u):
    xi, w = oracle.quadrature(2)
    full = (coords ** 2).sum(axis=1)
    full[gidx >= 0] = u
    X = coords[cells]
    J = 0.5 * np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=2)
    detJ = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
    err = 0.0
    for q in range(4):
        phi = np.array([-(xi[q, 0] + xi[q, 1]) / 2, (1 + xi[q, 0]) / 2, (1 + xi[q, 1]) / 2])
        xq = X[:, 0] + J @ (xi[q] + 1.0)
        uq = full[cells] @ phi
        err += (w[q] * np.abs(detJ) * (uq - (xq ** 2).sum(axis=1)) ** 2).sum()
    return float(np.sqrt(err))


@pytest.mark.gpu
def test_full_run_on_24k_fixture(oracle_mod):
    build()
    p = subprocess.run([EXE, "-mesh", golden_mesh_path("24k.osh"), "-run_type", "full", "-ksp_rtol", "1e-9", "-ksp_max_it", "100000",
                        "-snes_converged_reason", "-snes_monitor_short"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    m = re.search(r"N: (\d+) Newton its: (\d+) CG its: (\d+) L_2 Error: ([0-9.e+-]+)", p.stdout)
    assert m, p.stdout
    cells, coords, _ = load_osh_oracle("24k.osh")
    bnd = oracle_mod.mark_boundary(cells, coords.shape[0])
    gidx, N = oracle_mod.numbering(bnd)
    rowptr, colidx = oracle_mod.pattern(cells, gidx, N)
    ref = oracle_mod.newton(0, cells, coords, gidx, N, rowptr, colidx, ksp_rtol=1e-9, ksp_maxit=100000)
    assert int(m.group(1)) == N and int(m.group(2)) == 1
    assert abs(int(m.group(3)) - ref["ksp_its"]) <= ref["ksp_its"] // 40
    assert float(m.group(4)) == pytest.approx(oracle_l2_error(oracle_mod, cells, coords, gidx, ref["u"]), rel=1e-4)
    assert "Nonlinear solve converged due to CONVERGED_FNORM_RELATIVE iterations 1" in p.stdout
    assert "0 SNES Function norm 181.915" in p.stdout


@pytest.mark.gpu
def test_run_type_test_and_perf_and_analytic():
    build()
    p = subprocess.run([EXE, "-mesh", "box", "-cells", "48,40", "-run_type", "test"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    lin = float(re.search(r"Linear L_2 Residual: ([0-9.e+-]+)", p.stdout).group(1))
    assert lin < 1e-9, p.stdout  # A u + F(0) = F(u), ex12.cpp:1651-1661
    res = float(re.search(r"\nL_2 Residual: ([0-9.e+-]+)", p.stdout).group(1))
    assert res < 1e-9, "the nodal interpolant of x^2+y^2 solves the P1 system on the same-diagonal box"
    p = subprocess.run([EXE, "-mesh", golden_mesh_path("23elm.osh"), "-run_type", "perf"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "L_2 Residual: 10.9345" in p.stdout, p.stdout + p.stderr
    p = subprocess.run([EXE, "-mesh", "box", "-cells", "32,32", "-variable_coefficient", "analytic", "-ksp_rtol", "1e-10", "-pc_type", "jacobi"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert float(re.search(r"L_2 Error: ([0-9.e+-]+)", p.stdout).group(1)) < 2e-3
    # -ksp_type pipecg (KSPPIPECG): the pipelined recurrence through the facade, same answer
    p = subprocess.run([EXE, "-mesh", "box", "-cells", "32,32", "-variable_coefficient", "analytic", "-ksp_rtol", "1e-10", "-ksp_type", "pipecg"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert float(re.search(r"L_2 Error: ([0-9.e+-]+)", p.stdout).group(1)) < 2e-3


@pytest.mark.gpu
def test_unsupported_requests_fail_loudly():
    build()
    p = subprocess.run([EXE, "-mesh", "box", "-cells", "8,8", "-ksp_type", "gmres"], capture_output=True, text=True, timeout=120)
    assert p.returncode != 0 and "only cg and pipecg are built" in p.stderr
    p = subprocess.run([EXE, "-mesh", "/nonexistent.osh"], capture_output=True, text=True, timeout=120)
    assert p.returncode != 0


@pytest.mark.gpu
def test_vec_view_vtu(tmp_path, oracle_mod):
    """-vec_view vtk:<file>.vtu (ex12.cpp:1618; every reference CTest writes one, CMakeLists.txt:75-109)."""
    build()
    out = tmp_path / "sol.vtu"
    p = subprocess.run([EXE, "-mesh", golden_mesh_path("23elm.osh"), "-ksp_rtol", "1e-12", "-vec_view", "vtk:%s" % out],
                       capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    import xml.etree.ElementTree as ET
    root = ET.parse(out).getroot()
    piece = root.find("UnstructuredGrid/Piece")
    assert (piece.get("NumberOfPoints"), piece.get("NumberOfCells")) == ("19", "23")
    pts = np.array(piece.find("Points/DataArray").text.split(), float).reshape(-1, 3)
    conn = np.array(piece.find("Cells/DataArray[@Name='connectivity']").text.split(), int).reshape(-1, 3)
    vals = np.array(piece.find("PointData/DataArray").text.split(), float)
    cells, coords, _ = load_osh_oracle("23elm.osh")
    assert np.array_equal(conn, cells) and np.allclose(pts[:, :2], coords, rtol=0, atol=0)
    bnd = oracle_mod.mark_boundary(cells, coords.shape[0])
    gidx, N = oracle_mod.numbering(bnd)
    rowptr, colidx = oracle_mod.pattern(cells, gidx, N)
    ref = oracle_mod.newton(0, cells, coords, gidx, N, rowptr, colidx, ksp_rtol=1e-12)
    full = (coords ** 2).sum(axis=1)
    full[gidx >= 0] = ref["u"]
    assert np.abs(vals - full).max() < 1e-9


@pytest.mark.gpu
def test_boundary_found_by_walking_the_plex_like_ex12(oracle_mod):
    """`-boundary plex` runs the reference's own loop (ex12.cpp:966-1033: DMPlexGetHeightStratum, DMPlexGetSupportSize,
    DMPlexGetCone, DMSetLabelValue) against the facade; same result as the device-side detection."""
    build()
    outs = []
    for b in ("device", "plex"):
        p = subprocess.run([EXE, "-mesh", golden_mesh_path("24k.osh"), "-ksp_rtol", "1e-9", "-ksp_max_it", "100000", "-boundary", b, "-dm_view"],
                           capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stdout + p.stderr
        outs.append(re.search(r"N: (\d+) Newton its: (\d+) CG its: (\d+) L_2 Error: ([0-9.e+-]+)", p.stdout).groups())
        assert "DM Object: Mesh in 2 dimensions" in p.stdout and "2-cells: 23754" in p.stdout and "0-cells: 12060" in p.stdout
        if b == "plex":
            assert "1-cells: 35813" in p.stdout   # the 24k fixture has 35 813 edges
    assert outs[0] == outs[1] and int(outs[0][0]) == 11696


@pytest.mark.gpu
def test_circle_forcing_and_monitors(oracle_mod):
    build()
    p = subprocess.run([EXE, "-mesh", "box", "-cells", "64,64", "-variable_coefficient", "circle", "-ksp_rtol", "1e-10", "-snes_monitor_short",
                        "-ksp_monitor_short", "-ksp_converged_reason", "-snes_converged_reason"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    c, x = oracle_mod.box2d(64, 64)
    bnd = oracle_mod.mark_boundary(c, x.shape[0])
    gidx, N = oracle_mod.numbering(bnd)
    rowptr, colidx = oracle_mod.pattern(c, gidx, N)
    ref = oracle_mod.newton(3, c, x, gidx, N, rowptr, colidx, ksp_rtol=1e-10, ksp_maxit=10000)
    m = re.search(r"N: (\d+) Newton its: (\d+) CG its: (\d+) L_2 Error: ([0-9.e+-]+)", p.stdout)
    assert int(m.group(1)) == N and int(m.group(2)) == 1 and abs(int(m.group(3)) - ref["ksp_its"]) <= 2
    # PETSc's monitor layout: SNES lines indented 2, KSP lines indented 4, one KSP line per iteration + the initial norm
    assert p.stdout.count(" KSP Residual norm ") == int(m.group(3)) + 1
    assert re.search(r"^  0 SNES Function norm ", p.stdout, re.M) and re.search(r"^  1 SNES Function norm ", p.stdout, re.M)
    assert re.search(r"^    0 KSP Residual norm ", p.stdout, re.M)
    assert "Linear solve converged due to CONVERGED_RTOL iterations %s" % m.group(3) in p.stdout
    # the L2 error is measured against circle_u_2d, the problem's own exact solution (not the quadratic)
    exact = np.tanh(500.0 * (0.15 ** 2 - ((x[:, 0] - 0.5) ** 2 + (x[:, 1] - 0.5) ** 2))) + 1.0
    assert float(m.group(4)) < 0.2 * np.sqrt(np.mean(exact ** 2))
    # a function without a device instantiation is refused, not evaluated on the host
    q = subprocess.run([EXE, "-mesh", "box", "-cells", "8,8", "-variable_coefficient", "cross", "-run_type", "perf"], capture_output=True, text=True, timeout=120)
    assert q.returncode == 0, q.stdout + q.stderr
