"""CPU test (-m "not gpu") of the host-side multi-rank plumbing with world_size 2 over gloo: the
strip partitioner, the overlap closure, and the point-SF root resolution every rank feeds to
ohp_mesh_set_ghosts (PetscSFSetGraph, ex12.cpp:864-876)."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, %r)
    from omegahpetsc_b200 import meshgen, partition
    import oracle
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    for cells, coords in (meshgen.box2d(13, 9), meshgen.box3d(4, 3, 5)):
        eo, vo = partition.chain_partition(cells, coords, world)
        assert vo.max() < world and np.array_equal(np.bincount(eo, minlength=world) > 0, np.ones(world, bool))
        part = partition.local_part(cells, coords, eo, vo, rank)
        gv = part["gverts"]
        # closure: every element touching an owned vertex is local, so owned stars are complete
        touch = (vo[cells] == rank).any(axis=1)
        assert np.array_equal(part["elem_global"], np.nonzero(touch)[0])
        assert np.array_equal(gv[part["cells"]], cells[touch]) and np.array_equal(part["coords"], coords[gv])
        # ghosts = local vertices owned elsewhere; chain convention: owner = lowest part touching the vertex
        assert np.array_equal(part["ghost_local"], np.nonzero(vo[gv] != rank)[0])
        roots = partition.resolve_ghost_roots(part, vo, rank, world, dist.all_gather_object)
        tabs = [None] * world
        dist.all_gather_object(tabs, gv)
        for k, (r, root) in enumerate(zip(part["ghost_owner"], roots)):
            assert tabs[r][root] == part["ghost_gvert"][k] and vo[tabs[r][root]] == r
        # every vertex is owned by exactly one rank and appears as owned in that rank's part
        cnt = [None] * world
        dist.all_gather_object(cnt, int((vo[gv] == rank).sum()))
        assert sum(cnt) == coords.shape[0]
        # boundary of an owned vertex can be decided locally (its star is complete): compare with the oracle
        bnd_g = oracle.mark_boundary(cells, coords.shape[0])
        bnd_l = oracle.mark_boundary(part["cells"], gv.size)
        own = vo[gv] == rank
        assert np.array_equal(bnd_l[own], bnd_g[gv][own])
        assert (bnd_l[~own] >= bnd_g[gv][~own]).all()  # ghosts may be over-marked locally: owners' labels win
    dist.barrier()
    if rank == 0:
        print("GLOO_OK")
""")


def test_partition_plumbing_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29655", str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0 and "GLOO_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-3000:]
