#!/usr/bin/env python
"""bench.py -- P1 Poisson assembly + CG throughput (DOF/s) of the omegahPetsc ex12 hot path.

A step = one `-run_type full` Newton solve (ex12.cpp:1597-1618) on the synthetic workload:
2 residual assemblies + 1 Jacobian assembly + 1 CG solve (pc none, ksp_rtol 1e-9,
CMakeLists.txt:53).  Metric: DOF/s = N_dofs / time of that step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, OpenMP over the
host cores): the reference binary itself needs PETSc + Omega_h + MPI, none of which exist in this
image, so there is no oracle/_ref.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "poisson_p1_assembly_plus_cg_dof_per_s"
UNIT = "DOF/s"
KSP_RTOL = 1e-9
KSP_MAXIT = 1000000

WORKLOADS = {
    # name: (dim, sizes)   -- BASELINE.json configs 4 and 5
    "box2d_2896": (2, (2896, 2896)),   # 16 773 632 triangles, N = 8 381 025
    "box3d_110": (3, (110, 110, 110)),  # 7 986 000 tets, N = 1 295 029
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def make_mesh(workload, n_override, pinned=False):
    from omegahpetsc_b200 import meshgen
    dim, sizes = WORKLOADS[workload]
    if n_override:
        sizes = (n_override,) * dim
    if dim == 2:
        return meshgen.box2d(*sizes, pinned=pinned), sizes
    return meshgen.box3d(*sizes, pinned=pinned), sizes


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "500"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark(self):
        """Start of the timed region: samples written before this point (warm-up) are ignored."""
        try:
            self.offset = os.path.getsize(self.f.name)
        except OSError:
            self.offset = 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(getattr(self, "offset", 0))
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for nm, val in zip(names, c[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def algorithmic_bytes(dim, nC, nV, N, nnz):
    nc = dim + 1
    return {
        "residual": 4 * nc * nC + 8 * dim * nV + 8 * nV + 8 * nV,
        "jacobian": 4 * nc * nC + 8 * dim * nV + 8 * nnz,
        "spmv": 12 * nnz + 4 * (N + 1) + 16 * N,
        "cg_iteration": 12 * nnz + 4 * (N + 1) + 88 * N,
    }


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement (used for cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------
def cpu_step_sample(cells, coords, cg_sample_its, full_its, setup=None):
    """One bounded sample of the step on the host cores: 2 residuals + 1 Jacobian on the FULL mesh and
    `cg_sample_its` CG iterations, extrapolated to `full_its` iterations (same algorithm, same
    matrix => same count).  Returns (seconds for the full step, detail)."""
    import oracle
    t0 = time.perf_counter()
    if setup is None:
        bnd = oracle.mark_boundary(cells, coords.shape[0], omp=True)
        gidx, N = oracle.numbering(bnd)
        rowptr, colidx = oracle.pattern(cells, gidx, N, omp=True)
    else:
        gidx, N, rowptr, colidx = setup
    t_setup = time.perf_counter() - t0
    u = np.zeros(N)
    t0 = time.perf_counter()
    F = oracle.residual(0, cells, coords, gidx, N, u, omp=True)
    t_res = time.perf_counter() - t0
    t0 = time.perf_counter()
    vals = oracle.jacobian(0, cells, coords, gidx, N, u, rowptr, colidx, omp=True)
    t_jac = time.perf_counter() - t0
    t0 = time.perf_counter()
    cg = oracle.cg(rowptr, colidx, vals, F, rtol=KSP_RTOL, maxit=cg_sample_its, omp=True)
    t_cg = time.perf_counter() - t0
    its = max(1, cg["its"])
    t_iter = t_cg / its
    if cg["reason"] == -3 and not full_its:
        raise SystemExit("bench.py: CPU CG sample hit its cap and no full iteration count is known")
    n_its = full_its if cg["reason"] == -3 else cg["its"]
    total = 2 * t_res + t_jac + t_iter * n_its
    return total, dict(N=int(N), t_setup_s=t_setup, t_residual_s=t_res, t_jacobian_s=t_jac, t_cg_iter_s=t_iter,
                       cg_its_timed=int(cg["its"]), cg_its_extrapolated_to=int(n_its), setup=(gidx, N, rowptr, colidx))


def known_iterations(workload, sizes):
    p = os.path.join(ROOT, "profiles", "cg_iterations.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("%s:%s" % (workload, "x".join(map(str, sizes))))
    return None


def run_reference(args):
    import oracle
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    (cells, coords), sizes = make_mesh(args.workload, args.n)
    dim = coords.shape[1]
    full_its = known_iterations(args.workload, sizes)
    its_source = "count recorded in profiles/cg_iterations.json from the GPU arm"
    if full_its is None and min(sizes) >= 64:
        # no recorded count: CG iterations for this operator grow linearly with 1/h, so solve the same
        # problem at 1/8 of the resolution to convergence and scale the count by 8
        from omegahpetsc_b200 import meshgen
        small = tuple(max(8, s // 8) for s in sizes)
        c8, x8 = (meshgen.box2d if dim == 2 else meshgen.box3d)(*small)
        b8 = oracle.mark_boundary(c8, x8.shape[0], omp=True)
        g8, n8 = oracle.numbering(b8)
        r8, k8 = oracle.pattern(c8, g8, n8, omp=True)
        v8 = oracle.jacobian(0, c8, x8, g8, n8, np.zeros(n8), r8, k8, omp=True)
        f8 = oracle.residual(0, c8, x8, g8, n8, np.zeros(n8), omp=True)
        full_its = int(oracle.cg(r8, k8, v8, f8, rtol=KSP_RTOL, maxit=KSP_MAXIT, omp=True)["its"] * (sizes[0] / small[0]))
        its_source = "count ESTIMATED as 8x the converged count at 1/8 resolution"
    cores = oracle.num_threads(True)
    setup = None
    times = []
    detail = None
    for i in range(args.warmup + args.steps):
        t, detail = cpu_step_sample(cells, coords, args.cpu_cg_its, full_its, setup)
        setup = detail.pop("setup")
        if i >= args.warmup:
            times.append(t)
    N = detail["N"]
    sec = float(np.mean(times))
    val = N / sec
    sample = ("full mesh: 2 residual + 1 Jacobian assemblies timed in full; CG timed for %d iterations and "
              "extrapolated to %s iterations (%s)" % (detail["cg_its_timed"], detail["cg_its_extrapolated_to"],
                                                      its_source if detail["cg_its_extrapolated_to"] != detail["cg_its_timed"]
                                                      else "CG converged inside the sample"))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s (%s cells, dim %d) P1 Poisson, CG pc none rtol %g" % (
                args.workload, cells.shape[0], dim, KSP_RTOL), "sizes": list(sizes)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "detail": {k: v for k, v in detail.items()}},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import omegahpetsc_b200 as ohp
    from omegahpetsc_b200 import partition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    ctx = ohp.Context(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ids = [ohp.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(world, rank, ids[0])

    (cells, coords), sizes = make_mesh(args.workload, args.n, pinned=True)
    dim = coords.shape[1]
    n_cells_global, n_verts_global = cells.shape[0], coords.shape[0]
    if world > 1:
        eo, vo = partition.chain_partition(cells, coords, world)
        part = partition.local_part(cells, coords, eo, vo, rank)
        lcells, lcoords = part["cells"], part["coords"]
    else:
        part, lcells, lcoords = None, cells, coords

    def build_mesh():
        m = ohp.Mesh(ctx, lcells, lcoords)
        if part is not None:
            m.set_ghosts(part["ghost_local"], part["ghost_owner"],
                         partition.resolve_ghost_roots(part, vo, rank, world, dist.all_gather_object))
        m.setup()
        if part is not None and os.environ.get("OHP_COMM", "p2p") != "nccl":
            m.p2p_enable(dist.all_gather_object, world)
        return m

    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))
    t0 = time.perf_counter()
    mesh = build_mesh()
    ctx.sync()
    t_setup = time.perf_counter() - t0
    info = mesh.info()
    N = int(info.n_global)
    J, u = mesh.create_matrix(), mesh.create_global_vector()
    ksp = ohp.ksp_options(rtol=KSP_RTOL, max_it=KSP_MAXIT, pc=ohp.PC_NONE, check_every=args.check_every, profile=args.profile)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step():
        mesh.project(0, 0, u)   # u = 0 (ex12.cpp:1598-1602)
        return mesh.snes_solve(0, J, u, ksp=ksp)

    res = None
    # the sampler starts BEFORE the warm-up: the first nvidia-smi query on a fresh box stalls the driver for
    # seconds, which must not land inside the timed region; only samples taken after mark() are reported
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        res = step()
    barrier()
    if sampler:
        sampler.mark()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    phase = {"ms_residual": 0.0, "ms_jacobian": 0.0, "ms_ksp": 0.0, "spmv_ms": [], "ksp_its": 0}
    step_events = []
    for _ in range(args.steps):
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream)
        res = step()
        eb.record(stream)
        step_events.append((ea, eb))
        for k in ("ms_residual", "ms_jacobian", "ms_ksp"):
            phase[k] += res[k]
        phase["ksp_its"] += res["ksp_its"]
        if res["spmv_samples"]:
            phase["spmv_ms"].append(res["spmv_ms"])
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    ms_each = [a.elapsed_time(b) for a, b in step_events]
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = N / (ms_step * 1e-3)
    assert res["reason"] == 3 and res["its"] == 1, res

    # ---- end to end through the C-ABI with HOST buffers: H2D of the mesh, setup, solve, D2H of u ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        m2 = build_mesh()
        J2, u2 = m2.create_matrix(), m2.create_global_vector()
        m2.project(0, 0, u2)
        r2 = m2.snes_solve(0, J2, u2, ksp=ksp)
        uh = u2.to_host()
        assert r2["reason"] == 3 and np.isfinite(uh[:8]).all()
        J2.destroy(); u2.destroy(); m2.destroy()
    barrier()
    t_e2e = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    h2d = int(lcells.nbytes + lcoords.nbytes)
    d2h = int(info.n_owned * 8)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_kind = peaks()
    ab = algorithmic_bytes(dim, info.n_cells, info.n_verts, info.n_owned, info.nnz)
    spmv_ms = float(np.mean(phase["spmv_ms"])) if phase["spmv_ms"] else None
    roof = None
    if spmv_ms:
        ach = ab["spmv"] / (spmv_ms * 1e-3) / 1e9
        roof = {"kernel": "k_spmv_tma<true> (persistent TMA-pipelined CSR SpMV + fused (p,Ap))", "bound": "hbm", "achieved": ach, "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "bytes_per_launch": ab["spmv"], "ms_per_launch": spmv_ms, "launches_timed": int(res["spmv_samples"])}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr) and world == 1 and not args.n:  # the ncu capture is of the full matrix on one GPU
            with open(tr) as f:
                roof["traffic"] = json.load(f).get("%s:spmv" % args.workload)
    its_step = phase["ksp_its"] / args.steps
    kernels = {
        "residual": {"ms": phase["ms_residual"] / (2 * args.steps), "GBps": ab["residual"] / (phase["ms_residual"] / (2 * args.steps) * 1e-3) / 1e9},
        "jacobian": {"ms": phase["ms_jacobian"] / args.steps, "GBps": ab["jacobian"] / (phase["ms_jacobian"] / args.steps * 1e-3) / 1e9},
        "cg_iteration": {"ms": phase["ms_ksp"] / max(1, phase["ksp_its"]), "GBps": ab["cg_iteration"] / (phase["ms_ksp"] / max(1, phase["ksp_its"]) * 1e-3) / 1e9,
                         "iterations_per_step": its_step},
    }
    for k in kernels.values():
        k["frac_of_peak"] = k["GBps"] / peak
    kernels["phase_ms_per_step"] = {k: phase[k] / args.steps for k in ("ms_residual", "ms_jacobian", "ms_ksp")}
    kernels["phase_ms_per_step"]["each_step_ms"] = ms_each

    cpu = None
    if args.gpus == 1 and not args.no_cpu:
        import oracle
        # the oracle's setup is not part of the metric: reuse the (bit-exact, test-checked) device numbering
        gidx = mesh.numbering().astype(np.int32)
        rp, ci = mesh.csr()
        tcpu, d = cpu_step_sample(cells, coords, args.cpu_cg_its, int(round(its_step)), (gidx, info.n_owned, rp, ci.astype(np.int32)))
        d.pop("setup")
        cpu = {"value": N / tcpu, "unit": UNIT, "cores": oracle.num_threads(True), "kind": "port",
               "sample": "oracle (C restatement, OpenMP) on the full mesh: 2 residual + 1 Jacobian timed in full, CG timed for %d "
                         "iterations and extrapolated to the %d iterations the GPU solve took" % (d["cg_its_timed"], d["cg_its_extrapolated_to"]),
               "detail": d}
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "cg_iterations_seen.json"), "a") as f:
            f.write(json.dumps({"%s:%s" % (args.workload, "x".join(map(str, sizes))): int(round(its_step))}) + "\n")

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "%s: %d cells, %d verts, N=%d dofs, nnz=%d, dim %d; one -run_type full Newton step = 2 residual + "
                                   "1 Jacobian assemblies + CG (pc none, rtol %g, %d its)" % (
                                       args.workload, n_cells_global, n_verts_global, N, info.nnz if world == 1 else -1, dim, KSP_RTOL, int(round(its_step))),
                       "sizes": list(sizes), "parallelism": "1 GPU" if world == 1 else "%d-way chain partition; halo push + dot all-reduce %s" % (world, "over NCCL" if os.environ.get("OHP_COMM") == "nccl" else "fused into the CG kernels over NVLink peer memory"),
                       "l2": "inputs larger than L2 (matrix %.0f MB + vectors %.0f MB per rank vs 126 MB)" % (12 * info.nnz / 1e6, 40 * info.n_owned / 1e6),
                       "setup_ms_excluded": t_setup * 1e3},
            "e2e": {"value": N / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e * 1e3,
                    "steps": e2e_steps, "includes": "ohp_mesh_create (H2D of cells+coords from pinned host memory), device numbering + CSR build, "
                                                    "Newton step, D2H of the solution"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": kernels, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="box2d_2896", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="override the box resolution (development only)")
    ap.add_argument("--check-every", type=int, default=0)
    ap.add_argument("--profile", type=int, default=64, help="SpMV launches bracketed by CUDA events per solve")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-cg-its", type=int, default=60)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
