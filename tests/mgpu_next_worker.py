"""Multi-GPU parity worker of the SURVEY 8(f) rank-4 slice (run under torch.distributed.run, one rank per GPU): the
Chebyshev/Jacobi-preconditioned CG on a partitioned Dirichlet mesh, and the natural (-bc_type neumann) boundary
condition with the null-space CG on a partitioned mesh.  Reference = the single-process oracle on the unpartitioned mesh,
compared through global vertex ids as in mgpu_worker.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import omegahpetsc_b200 as ohp  # noqa: E402
from omegahpetsc_b200 import meshgen, partition  # noqa: E402


def build(ctx, cells, coords, eo, vo, rank, world, natural):
    part = partition.local_part(cells, coords, eo, vo, rank)
    m = ohp.Mesh(ctx, part["cells"], part["coords"])
    roots = partition.resolve_ghost_roots(part, vo, rank, world, dist.all_gather_object)
    m.set_ghosts(part["ghost_local"], part["ghost_owner"], roots)
    if natural:
        m.set_bc_type(ohp.BC_NATURAL)
    m.setup()
    if os.environ.get("OHP_COMM", "p2p") != "nccl":
        m.p2p_enable_native()
    return m, part


def dof_maps(m, part, vo, rank, world, free_mask_global):
    """global dof -> global vertex (all ranks), and this rank's first row"""
    gv = part["gverts"]
    mine = (vo[gv] == rank) & free_mask_global[gv]
    tabs = [None] * world
    dist.all_gather_object(tabs, gv[mine])
    start = int(sum(len(t) for t in tabs[:rank]))
    return np.concatenate(tabs), start


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ohp.Context(local)
    ids = [ohp.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(world, rank, ids[0])
    import oracle

    cells, coords = meshgen.box2d(40, 43)
    eo, vo = partition.chain_partition(cells, coords, world)
    nV = coords.shape[0]
    bnd = oracle.mark_boundary(cells, nV)

    # ---- 1. Dirichlet mesh, Chebyshev/Jacobi polynomial inside CG ----
    gidx, N = oracle.numbering(bnd)
    rowptr, colidx = oracle.pattern(cells, gidx, N)
    vals = oracle.jacobian(0, cells, coords, gidx, N, np.zeros(N), rowptr, colidx)
    F0 = oracle.residual(0, cells, coords, gidx, N, np.zeros(N))
    m, part = build(ctx, cells, coords, eo, vo, rank, world, natural=False)
    info = m.info()
    dof2vert, start = dof_maps(m, part, vo, rank, world, bnd == 0)
    assert (info.n_global, info.row_start) == (N, start)
    mine = gidx[dof2vert[start:start + info.n_owned]]  # serial dof of each of my rows
    u, F, J, x = m.create_global_vector().set(0.0), m.create_global_vector(), m.create_matrix(), m.create_global_vector()
    m.residual(0, u, F)
    m.jacobian(0, u, J)
    its_seen = []
    for k in (1, 4, 8):
        want = oracle.pcg(rowptr, colidx, vals, F0, rtol=1e-9, pc=2, cheb_its=k)
        got = ohp.ksp_cg(J, F, x, ohp.ksp_options(rtol=1e-9, pc=ohp.PC_CHEBYSHEV, cheb_its=k))
        assert got["algo"] == ohp.ALGO_GENERAL_PCG and got["reason"] == want["reason"] == 2, (k, got, want["reason"])
        assert abs(got["its"] - want["its"]) <= max(3, want["its"] // 20), (k, got["its"], want["its"])
        # the noise vector is a hash of the GLOBAL row number, but the rows are numbered rank by rank here and vertex by vertex in
        # the serial run: the estimates agree as estimates (both are Ritz values just below lambda_max), not to rounding
        assert abs(got["cheb_emax"] - want["emax"]) <= 0.05 * want["emax"], (got["cheb_emax"], want["emax"])
        assert np.abs(x.to_host() - want["x"][mine]).max(initial=0.0) <= 1e-6 * np.abs(want["x"]).max(), "chebyshev solution k=%d" % k
        its_seen.append(got["its"])
    if rank == 0:
        print("stage 1 ok: chebyshev its (k=1,4,8) %s" % its_seen, flush=True)
    m.destroy()

    # ---- 2. natural boundary condition: nothing constrained, boundary integral, constant null space ----
    gid_n = np.arange(nV, dtype=np.int32)
    rp_n, ci_n = oracle.pattern(cells, gid_n, nV)
    vals_n = oracle.jacobian(0, cells, coords, gid_n, nV, np.zeros(nV), rp_n, ci_n)
    m, part = build(ctx, cells, coords, eo, vo, rank, world, natural=True)
    info = m.info()
    dof2vert, start = dof_maps(m, part, vo, rank, world, np.ones(nV, bool))
    assert (info.n_global, info.row_start) == (nV, start), (info.n_global, info.row_start, nV, start)
    myv = dof2vert[start:start + info.n_owned]
    assert np.array_equal(m.boundary(), bnd[part["gverts"]]), "boundary label on the natural mesh"
    u, F, J, x = m.create_global_vector(), m.create_global_vector(), m.create_matrix(), m.create_global_vector()
    for uh in (np.zeros(nV), np.sin(3 * coords[:, 0]) * np.cos(2 * coords[:, 1])):
        u.from_host(uh[myv])
        want = oracle.bd_residual(cells, coords, gid_n, nV, uh, oracle.residual(0, cells, coords, gid_n, nV, uh))
        got = m.residual(0, u, F).to_host()
        assert np.abs(got - want[myv]).max(initial=0.0) <= 1e-12 * np.abs(want).max(), "natural residual"
    if rank == 0:
        print("stage 2a ok: natural residual", flush=True)
    u.set(0.0)
    m.residual(0, u, F)
    m.jacobian(0, u, J)
    J.set_nullspace(True)
    F0n = oracle.bd_residual(cells, coords, gid_n, nV, np.zeros(nV), oracle.residual(0, cells, coords, gid_n, nV, np.zeros(nV)))
    for pc, k in ((ohp.PC_JACOBI, 0), (ohp.PC_CHEBYSHEV, 4)):
        want = oracle.pcg(rp_n, ci_n, vals_n, F0n, rtol=1e-9, pc=pc, cheb_its=k, nullspace=True)
        got = ohp.ksp_cg(J, F, x, ohp.ksp_options(rtol=1e-9, pc=pc, cheb_its=k))
        assert got["algo"] == ohp.ALGO_GENERAL_PCG and got["reason"] == want["reason"] == 2, (pc, got)
        assert abs(got["its"] - want["its"]) <= max(3, want["its"] // 20), (pc, got["its"], want["its"])
        x.remove_constant()  # global mean (MatNullSpaceRemove)
        xo = want["x"] - want["x"].mean()
        assert np.abs(x.to_host() - xo[myv]).max(initial=0.0) <= 1e-6 * np.abs(xo).max(), "null-space solution pc=%d" % pc
    if rank == 0:
        print("stage 2b ok: null-space solves", flush=True)
    # Newton on top of it
    u.set(0.0)
    r = m.snes_solve(0, J, u, ksp=ohp.ksp_options(rtol=1e-10, pc=ohp.PC_JACOBI))
    assert r["reason"] in (2, 3, 4) and r["algo"] == ohp.ALGO_GENERAL_PCG, r
    u.remove_constant()
    exact = (coords ** 2).sum(axis=1)
    assert np.abs(u.to_host() - (exact - exact.mean())[myv]).max(initial=0.0) < 5e-3
    dist.barrier()
    if rank == 0:
        print("MGPU_NEXT_OK world=%d N=%d chebyshev its (k=1,4,8) %s natural N=%d" % (world, N, its_seen, nV))
    m.destroy()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
