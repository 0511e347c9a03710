"""Multi-GPU parity (-m gpu): spawns tests/mgpu_worker.py with one rank per GPU.  Skipped on a
box with a single GPU; `gpurun --gpus 2|4` runs it."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def ngpus():
    import torch
    return torch.cuda.device_count()


def run_worker(world, mesh, port, comm="p2p"):
    cmd = ["timeout", "300", sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"), mesh]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OHP_COMM=comm))
    assert p.returncode == 0 and "MGPU_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]


@pytest.mark.parametrize("comm", ["p2p", "nccl"])
@pytest.mark.parametrize("mesh", ["box2d_40", "24k_4p", "box3d_9"])
def test_two_gpus(mesh, comm):
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    run_worker(2, mesh, 29611, comm)


def test_two_gpus_from_picpart_files():
    """Each rank reads its own pumipic picpart (23elm 2-part fixture) and extracts its core on the device."""
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    run_worker(2, "pic:23elm_2p", 29613)


@pytest.mark.parametrize("mesh", ["24k_4p", "box2d_40", "pic:24k_4p"])
def test_four_gpus(mesh):
    if ngpus() < 4:
        pytest.skip("needs 4 GPUs")
    run_worker(4, mesh, 29612)


@pytest.mark.parametrize("mesh", ["24k_24p", "box3d_12", "box2d_64"])
def test_eight_gpus(mesh):
    """BASELINE config 3 at 8 GPUs (the 24-part XGC partition merged 3 -> 1) and the boxes."""
    if ngpus() < 8:
        pytest.skip("needs 8 GPUs")
    run_worker(8, mesh, 29614)


def test_two_gpus_merged_24p():
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    run_worker(2, "24k_24p", 29615)


def run_cxx(world, args, port, env=None):
    """The C++ drop-in under torch.distributed.run (no Python in the ranks, no MPI): PetscInitialize forms the
    communicator from RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT."""
    exe = os.path.join(ROOT, "examples", "ex12_b200")
    cmd = ["timeout", "300", sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), "--no-python", exe] + args
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-4000:]
    return p.stdout


def serial_reference(oracle, mesh_file):
    from conftest import load_osh_oracle
    cells, coords, _ = load_osh_oracle(mesh_file)
    bnd = oracle.mark_boundary(cells, coords.shape[0])
    gidx, N = oracle.numbering(bnd)
    rowptr, colidx = oracle.pattern(cells, gidx, N)
    ref = oracle.newton(0, cells, coords, gidx, N, rowptr, colidx, ksp_rtol=1e-9, ksp_maxit=100000)
    return cells, coords, gidx, N, ref


@pytest.mark.parametrize("boundary", ["device", "plex"])
def test_cxx_picpart_four_gpus(oracle_mod, boundary):
    """`ex12_b200 -mesh picpart -picpart_path ...` on the reference's 4-part XGC picparts (CMakeLists.txt:104-111):
    every rank reads its own file, extracts its core on the device, resolves the ghost roots with the device gid
    join, builds the SF in Plex points as ex12 does, and the solve matches the serial oracle."""
    import re
    from conftest import golden_mesh_path
    from test_facade import oracle_l2_error
    if ngpus() < 4:
        pytest.skip("needs 4 GPUs")
    out = run_cxx(4, ["-mesh", "picpart", "-picpart_path", golden_mesh_path("24k_4p_"), "-run_type", "full", "-ksp_rtol", "1e-9",
                      "-ksp_max_it", "100000", "-snes_converged_reason", "-snes_monitor_short", "-boundary", boundary], 29621)
    cells, coords, gidx, N, ref = serial_reference(oracle_mod, "24k.osh")
    m = re.findall(r"N: (\d+) Newton its: (\d+) CG its: (\d+) L_2 Error: ([0-9.e+-]+)", out)
    assert len(m) == 1, "PETSC_COMM_WORLD prints once\n" + out
    n, newton, cg, l2 = int(m[0][0]), int(m[0][1]), int(m[0][2]), float(m[0][3])
    assert n == N and newton == 1 and abs(cg - ref["ksp_its"]) <= max(2, ref["ksp_its"] // 40)
    assert l2 == pytest.approx(oracle_l2_error(oracle_mod, cells, coords, gidx, ref["u"]), rel=5e-3)
    assert out.count("Nonlinear solve converged due to CONVERGED_FNORM_RELATIVE iterations 1") == 1
    assert "0 SNES Function norm 181.915" in out


def test_cxx_picpart_two_gpus_23elm(oracle_mod):
    """The 2-part 23-element picpart (one rank owns very few rows): C++ drop-in, NCCL transport and NVLink transport."""
    import re
    from conftest import golden_mesh_path
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    _, _, _, N, ref = serial_reference(oracle_mod, "23elm.osh")
    for comm in ("p2p", "nccl"):
        out = run_cxx(2, ["-mesh", "picpart", "-picpart_path", golden_mesh_path("23elm_2p_"), "-ksp_rtol", "1e-12", "-ksp_monitor_short",
                          "-ksp_converged_reason"], 29622, env={"OHP_COMM": comm})
        m = re.search(r"N: (\d+) Newton its: (\d+) CG its: (\d+)", out)
        assert m and int(m.group(1)) == N and int(m.group(2)) == 1, out
        assert abs(int(m.group(3)) - ref["ksp_its"]) <= 1
        assert out.count("KSP Residual norm") == int(m.group(3)) + 1 and "Linear solve converged due to CONVERGED_RTOL" in out


def test_two_gpus_next_rows():
    """SURVEY 8(f) rank 4 across ranks: Chebyshev/Jacobi-preconditioned CG, natural boundary condition, null-space CG, Newton."""
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = ["timeout", "300", sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "tests", "mgpu_next_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "MGPU_NEXT_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
