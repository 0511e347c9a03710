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
