#!/usr/bin/env python
"""bench.py -- P1 Poisson assembly + CG throughput (DOF/s) of the omegahPetsc ex12 hot path.

A step = one `-run_type full` Newton solve (ex12.cpp:1597-1618) on the synthetic workload:
2 residual assemblies + 1 Jacobian assembly + 1 CG solve (pc none, ksp_rtol 1e-9,
CMakeLists.txt:53).  Metric: DOF/s = N_dofs / time of that step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

The JSON line carries, beside the contract keys: `check` (solution against the manufactured solution,
CG iteration count against the oracle's own converged count, and at N=1 the assembled F and J against the
oracle on the FULL mesh), `roofline` (dominant kernel, timed live), `kernels`, `cpu_baseline`, and
`secondary` (BASELINE config 5, the 3-D tet box, measured the same way with its own clock record).

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, OpenMP over ALL host
cores whatever OMP_NUM_THREADS says): the reference binary itself needs PETSc + Omega_h + MPI, none of
which exist in this image, so there is no oracle/_ref.
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
    "box2d_2896": (2, (2896, 2896)),    # 16 773 632 triangles, N = 8 381 025
    "box3d_110": (3, (110, 110, 110)),  # 7 986 000 tets, N = 1 295 029
    # config 4 with the vertices renumbered along a Morton (Z-order) curve: Omega_h::build_box reorders its
    # vertices along a space-filling curve (ex12.cpp:805), so this is the ordering-robustness variant
    "box2d_2896_morton": (2, (2896, 2896)),
    "box2d_2896_hilbert": (2, (2896, 2896)),   # the curve Omega_h::build_box sorts by (ex12.cpp:805); generator only, not measured in round 2
}
# solution tolerance of the `check` object: CG stops at ||r|| <= 1e-9 ||b||; the manufactured solution is nodally
# exact on these boxes, so the nodal error is the algebraic error of the solve
CHECK_NODAL_TOL = 1e-6


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
        cells, coords = meshgen.box2d(*sizes, pinned=pinned)
        if workload.endswith("_morton"):
            cells, coords = meshgen.morton_reorder(cells, coords, sizes, pinned=pinned)
        if workload.endswith("_hilbert"):
            cells, coords = meshgen.hilbert_reorder(cells, coords, sizes, pinned=pinned)
        return (cells, coords), sizes
    return meshgen.box3d(*sizes, pinned=pinned), sizes


def workload_string(workload, sizes, dim, n_cells):
    """Identical in both arms (the driver compares the strings)."""
    return ("%s: %s box, %d cells, dim %d, P1 Poisson; step = one -run_type full Newton solve = 2 residual + 1 Jacobian "
            "assemblies + CG (pc none, rtol %g)" % (workload, "x".join(map(str, sizes)), n_cells, dim, KSP_RTOL))


def exact_solution(coords):
    """quadratic_u_2d / quadratic_u_3d (ex12.cpp:113-117, 408-412)."""
    return (coords ** 2).sum(axis=1) * (1.0 if coords.shape[1] == 2 else 2.0 / 3.0)


def oracle_iterations(workload, sizes):
    """Converged CG iteration count of the ORACLE on this workload (tools/oracle_cg_iterations.py)."""
    p = os.path.join(ROOT, "tests", "golden", "oracle_cg_iterations.json")
    if os.path.exists(p):
        with open(p) as f:
            rec = json.load(f).get("%s:%s" % (workload, "x".join(map(str, sizes))))
        if rec:
            return int(rec["its"])
    return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=200):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", str(period_ms)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark(self):
        """Samples written before this point are not part of the next window."""
        try:
            self.offset = os.path.getsize(self.f.name)
        except OSError:
            self.offset = 0

    def window(self):
        """Summary of the samples since mark() (the sampler keeps running)."""
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvidia-smi"}
        if self.p is None:
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.f.name) as f:
            f.seek(getattr(self, "offset", 0))
            lines = f.read().splitlines()
        for line in lines:
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
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out

    def stop(self):
        if self.p is None:
            return
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.close()
        os.unlink(self.f.name)


class NvmlSampler:
    """The same record taken through NVML from a thread of this process (no nvidia-smi process polling the
    driver): clocks.sm, clocks.max.sm and the clocks-event (throttle) reasons, every `period_ms`."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, period_ms=500):
        import threading
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        self.period = period_ms * 1e-3
        self.rows, self.start, self.halt = [], 0, threading.Event()
        self.p = threading.Thread(target=self._run, daemon=True)
        self.p.start()

    def _run(self):
        nv = self.nv
        while not self.halt.wait(self.period):
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((sm, mask))
            except Exception:
                pass

    def mark(self):
        self.start = len(self.rows)

    def window(self):
        rows = self.rows[self.start:]
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvml"}
        if rows:
            reasons = set()
            for _, mask in rows:
                for bit, nm in self.REASONS.items():
                    if mask & bit:
                        reasons.add(nm)
            out.update(sm_mhz=float(np.median([r[0] for r in rows])), sm_max_mhz=self.mx, reasons=sorted(reasons), samples=len(rows))
        return out

    def stop(self):
        self.halt.set()
        self.p.join(timeout=2)


def make_sampler(kind, index, period_ms):
    if kind == "none":
        return None
    if kind == "nvml":
        try:
            return NvmlSampler(index, period_ms)
        except Exception:
            pass
    return ClockSampler(index, period_ms)


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
    detail = dict(N=int(N), t_setup_s=t_setup, t_residual_s=t_res, t_jacobian_s=t_jac, t_cg_iter_s=t_iter,
                  cg_its_timed=int(cg["its"]), cg_its_extrapolated_to=int(n_its), extrapolated=bool(n_its != cg["its"]),
                  measured_fraction_of_cg=float(cg["its"]) / float(n_its),
                  measured_seconds=2 * t_res + t_jac + t_cg)
    return total, detail, dict(setup=(gidx, N, rowptr, colidx), F=F, vals=vals)


def estimate_iterations(dim, sizes):
    """No recorded oracle count for this size: CG iterations for this operator grow (slightly sub-)linearly
    with 1/h, so solve the same problem at 1/8 of the resolution to convergence and scale the count."""
    import oracle
    from omegahpetsc_b200 import meshgen
    small = tuple(max(8, s // 8) for s in sizes)
    c8, x8 = (meshgen.box2d if dim == 2 else meshgen.box3d)(*small)
    b8 = oracle.mark_boundary(c8, x8.shape[0], omp=True)
    g8, n8 = oracle.numbering(b8)
    r8, k8 = oracle.pattern(c8, g8, n8, omp=True)
    v8 = oracle.jacobian(0, c8, x8, g8, n8, np.zeros(n8), r8, k8, omp=True)
    f8 = oracle.residual(0, c8, x8, g8, n8, np.zeros(n8), omp=True)
    return int(oracle.cg(r8, k8, v8, f8, rtol=KSP_RTOL, maxit=KSP_MAXIT, omp=True)["its"] * (sizes[0] / small[0]))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    cores = oracle.use_all_cores()  # torchrun exports OMP_NUM_THREADS=1: ignore it
    (cells, coords), sizes = make_mesh(args.workload, args.n)
    dim = coords.shape[1]
    full_its = oracle_iterations(args.workload, sizes)
    its_source = "the oracle's own converged count, tests/golden/oracle_cg_iterations.json"
    if full_its is None and min(sizes) >= 64:
        full_its = estimate_iterations(dim, sizes)
        its_source = "count ESTIMATED as 8x the oracle's converged count at 1/8 resolution"
    setup = None
    times, measured, its_timed = [], 0.0, 0
    detail = None
    for i in range(args.warmup + args.steps):
        timed = i >= args.warmup
        t, detail, keep = cpu_step_sample(cells, coords, args.cpu_cg_its if timed else min(20, args.cpu_cg_its), full_its, setup)
        setup = keep["setup"]
        if timed:
            times.append(t)
            measured += detail["measured_seconds"]
            its_timed += detail["cg_its_timed"]
    N = detail["N"]
    sec = float(np.mean(times))
    val = N / sec
    sample = ("full mesh, per step: 2 residual + 1 Jacobian assemblies timed in full; CG timed for %d iterations (%d over the "
              "%d timed steps) and extrapolated to %s iterations (%s)" % (
                  detail["cg_its_timed"], its_timed, args.steps, detail["cg_its_extrapolated_to"],
                  its_source if detail["extrapolated"] else "CG converged inside the sample"))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(args.workload, sizes, dim, cells.shape[0]), "sizes": list(sizes)},
            "extrapolated": bool(detail["extrapolated"]),
            "measured_fraction": detail["measured_fraction_of_cg"],
            "measured_seconds_in_timed_steps": measured,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "extrapolated": bool(detail["extrapolated"]), "detail": detail},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Job:
    """Process-wide state of the GPU arm (one rank = one GPU)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import omegahpetsc_b200 as ohp
        self.torch, self.dist, self.ohp = torch, dist, ohp
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, self.world))
        torch.cuda.set_device(self.local_rank)
        self.ctx = ohp.Context(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            ids = [ohp.Context.comm_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            self.ctx.comm_init(self.world, self.rank, ids[0])
        self.stream = torch.cuda.ExternalStream(self.ctx.stream(), device=torch.device("cuda", self.local_rank))

    def barrier(self):
        self.ctx.sync()
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def measure_workload(job, args, workload, steps, warmup, e2e_steps, sampler, with_cpu):
    """Times `steps` Newton steps of `workload` on the job's GPUs; returns the pieces of the JSON line."""
    torch, dist, ohp = job.torch, job.dist, job.ohp
    from omegahpetsc_b200 import partition
    ctx, world, rank = job.ctx, job.world, job.rank
    (cells, coords), sizes = make_mesh(workload, args.n, pinned=True)
    dim = coords.shape[1]
    n_cells_global, n_verts_global = cells.shape[0], coords.shape[0]
    if world > 1:
        eo, vo = partition.chain_partition(cells, coords, world)
        part = partition.local_part(cells, coords, eo, vo, rank)
        lcells, lcoords = part["cells"], part["coords"]
    else:
        part, lcells, lcoords = None, cells, coords

    breakdown = {} if os.environ.get("OHP_BENCH_E2E_BREAKDOWN") else None

    def lap(name, t_prev):
        """opt-in phase timing of the e2e path (synchronises after every phase: diagnostic only, never the reported e2e)"""
        if breakdown is None:
            return t_prev
        ctx.sync()
        now = time.perf_counter()
        breakdown[name] = breakdown.get(name, 0.0) + (now - t_prev) * 1e3
        return now

    if part is not None:
        # host inputs of the partitioned build, prepared once like the cell list itself (what a picpart file holds: global
        # ids, owners): NOT part of any timed region
        core_gid = np.ascontiguousarray(part["gverts"], dtype=np.int64)
        core_owned = np.ascontiguousarray(vo[part["gverts"]] == rank, dtype=np.uint8)
        ghost_local = np.ascontiguousarray(part["ghost_local"], dtype=np.int32)
        ghost_owner = np.ascontiguousarray(part["ghost_owner"], dtype=np.int32)
        ghost_gid = np.ascontiguousarray(core_gid[ghost_local])

    def build_mesh():
        t = time.perf_counter()
        m = ohp.Mesh(ctx, lcells, lcoords)
        t = lap("mesh_create_h2d", t)
        if part is not None:
            # A6 behind the C-ABI: ghost roots from global vertex ids on the device (ohp_ghost_roots_from_gids)
            roots = ctx.ghost_roots_from_gids(ghost_gid, ghost_owner, core_gid, core_owned)
            m.set_ghosts(ghost_local, ghost_owner, roots)
            t = lap("ghost_roots_from_gids", t)
        m.setup()
        t = lap("mesh_setup", t)
        if part is not None and os.environ.get("OHP_COMM", "p2p") != "nccl":
            m.p2p_enable_native()   # CUDA IPC handles all-gathered over the library's own communicator
            t = lap("p2p_enable", t)
        return m

    t0 = time.perf_counter()
    mesh = build_mesh()
    ctx.sync()
    t_setup = time.perf_counter() - t0
    info = mesh.info()
    N = int(info.n_global)
    J, u = mesh.create_matrix(), mesh.create_global_vector()
    ksp = ohp.ksp_options(rtol=KSP_RTOL, max_it=KSP_MAXIT, pc=ohp.PC_NONE, check_every=args.check_every, profile=args.profile)

    def step():
        mesh.project(0, 0, u)   # u = 0 (ex12.cpp:1598-1602)
        return mesh.snes_solve(0, J, u, ksp=ksp)

    res = None
    for _ in range(warmup):
        res = step()
    job.barrier()
    if sampler:
        sampler.mark()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(job.stream)
    phase = {"ms_residual": 0.0, "ms_jacobian": 0.0, "ms_ksp": 0.0, "spmv_ms": [], "ksp_its": 0}
    step_events = []
    for _ in range(steps):
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(job.stream)
        res = step()
        eb.record(job.stream)
        step_events.append((ea, eb))
        for k in ("ms_residual", "ms_jacobian", "ms_ksp"):
            phase[k] += res[k]
        phase["ksp_its"] += res["ksp_its"]
        if res["spmv_samples"]:
            phase["spmv_ms"].append(res["spmv_ms"])
    e1.record(job.stream)
    job.barrier()
    clocks = sampler.window() if sampler else None
    launches = ctx.launch_count() - l0
    ms_each = [a.elapsed_time(b) for a, b in step_events]
    ms_total = job.max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / steps
    value = N / (ms_step * 1e-3)
    its_step = phase["ksp_its"] / steps

    # ---- check: the result of the LAST timed step against the manufactured solution (every N) ----
    gid = mesh.numbering()
    owned = (gid >= info.row_start) & (gid < info.row_start + info.n_owned)
    uh = u.to_host()
    ex = exact_solution(lcoords[owned])   # owned free vertices in ascending local order = row order
    nodal_err = job.max_over_ranks(float(np.abs(uh - ex).max(initial=0.0)))
    its_oracle = oracle_iterations(workload, sizes)
    check = {"newton_its": int(res["its"]), "snes_reason": int(res["reason"]), "cg_its": int(round(its_step)),
             "cg_its_oracle": its_oracle, "max_nodal_error_vs_exact": nodal_err, "nodal_tol": CHECK_NODAL_TOL,
             "fnorm_reduction": float(res["fnorm"] / res["fnorm0"]) if res["fnorm0"] else None}
    ok = (res["reason"] == 3 and res["its"] == 1 and nodal_err <= CHECK_NODAL_TOL)
    if its_oracle:
        check["cg_its_rel_diff_vs_oracle"] = abs(its_step - its_oracle) / its_oracle
        ok = ok and check["cg_its_rel_diff_vs_oracle"] <= 0.03   # rounding-order band of the recurrence

    # ---- end to end through the C-ABI with HOST buffers: H2D of the mesh, setup, solve, D2H of u ----
    e2e = None
    if e2e_steps > 0:
        job.barrier()
        t0 = time.perf_counter()
        e2e_each = []
        for _ in range(e2e_steps):
            t_one = time.perf_counter()
            m2 = build_mesh()
            t = time.perf_counter()
            J2, u2 = m2.create_matrix(), m2.create_global_vector()
            m2.project(0, 0, u2)
            r2 = m2.snes_solve(0, J2, u2, ksp=ksp)
            t = lap("solve", t)
            uh2 = u2.to_host()
            assert r2["reason"] == 3 and np.isfinite(uh2[:8]).all()
            t = lap("d2h", t)
            J2.destroy(); u2.destroy(); m2.destroy()
            t = lap("destroy", t)
            e2e_each.append((time.perf_counter() - t_one) * 1e3)
        job.barrier()
        t_e2e = job.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        e2e = {"value": N / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(lcells.nbytes + lcoords.nbytes),
               "d2h_bytes_per_step": int(info.n_owned * 8), "ms_per_step": t_e2e * 1e3, "steps": e2e_steps,
               "each_step_ms_rank0": e2e_each,
               **({"breakdown_ms_total_rank0_diagnostic": breakdown} if breakdown else {}),
               "includes": "ohp_mesh_create (H2D of cells+coords from pinned host memory), device numbering + CSR build, "
                           "Newton step, D2H of the solution"}

    peak, peak_kind = peaks()
    ab = algorithmic_bytes(dim, info.n_cells, info.n_verts, info.n_owned, info.nnz)
    spmv_ms = float(np.mean(phase["spmv_ms"])) if phase["spmv_ms"] else None
    roof = None
    algo = int(res.get("algo", 0))
    if spmv_ms:
        # which kernel the solver bracketed with CUDA events (ohp_ksp_result.algo)
        if algo == 4:   # one launch per iteration: SpMV + the whole pipelined-CG vector update of the row
            kname, kbytes = "k_pipecg_fused (persistent TMA-pipelined CSR SpMV with the pipelined-CG vector update fused per row)", ab["spmv"] - 16 * info.n_owned + 96 * info.n_owned
        elif algo == 3:
            kname, kbytes = "k_spmv_tma<true> (persistent TMA-pipelined CSR SpMV; pipelined CG, reducer warp)", ab["spmv"]
        else:
            kname, kbytes = "k_spmv_tma<true> (persistent TMA-pipelined CSR SpMV + fused (p,Ap))", ab["spmv"]
        ach = kbytes / (spmv_ms * 1e-3) / 1e9
        roof = {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "bytes_per_launch": kbytes, "ms_per_launch": spmv_ms, "launches_timed": int(res["spmv_samples"]),
                "cg_recurrence": {0: "classic (3 launches)", 1: "single-reduction (2 launches)", 2: "persistent cooperative kernel",
                                  3: "pipelined (2 launches)", 4: "pipelined (1 launch)"}.get(algo, str(algo)),
                "scope": "rank 0's share of the matrix" if world > 1 else "whole matrix"}
        if world > 1 and 12 * info.nnz + 100 * info.n_owned < 200e6:
            roof["note"] = "rank 0's share (%.0f MB) is of the order of the 126 MB L2: the vectors are L2-resident, achieved may exceed the HBM line" % ((12 * info.nnz + 100 * info.n_owned) / 1e6)
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr) and world == 1 and not args.n and algo != 4:  # ncu capture of the full matrix on one GPU (profiles/)
            with open(tr) as f:
                roof["traffic"] = json.load(f).get("%s:spmv" % workload)
                roof["traffic_source"] = "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel on this workload (profiles/), not re-measured in this run"
    n_res = max(1, 2 * steps)
    kernels = {
        "residual": {"ms": phase["ms_residual"] / n_res, "GBps": ab["residual"] / (phase["ms_residual"] / n_res * 1e-3) / 1e9},
        "jacobian": {"ms": phase["ms_jacobian"] / steps, "GBps": ab["jacobian"] / (phase["ms_jacobian"] / steps * 1e-3) / 1e9},
        "cg_iteration": {"ms": phase["ms_ksp"] / max(1, phase["ksp_its"]), "GBps": ab["cg_iteration"] / (phase["ms_ksp"] / max(1, phase["ksp_its"]) * 1e-3) / 1e9,
                         "iterations_per_step": its_step},
    }
    for k in kernels.values():
        k["frac_of_peak"] = k["GBps"] / peak
    kernels["phase_ms_per_step"] = {k: phase[k] / steps for k in ("ms_residual", "ms_jacobian", "ms_ksp")}
    kernels["phase_ms_per_step"]["each_step_ms"] = ms_each
    kernels["scope"] = "rank 0 (bytes of its share)" if world > 1 else "whole mesh"

    cpu = None
    if with_cpu and world == 1:
        import oracle
        cores = oracle.use_all_cores()
        # the oracle's setup is not part of the metric: reuse the (bit-exact, test-checked) device numbering
        gidx = gid.astype(np.int32)
        rp, ci = mesh.csr()
        full_its = its_oracle or int(round(its_step))
        tcpu, d, keep = cpu_step_sample(cells, coords, args.cpu_baseline_cg_its, full_its, (gidx, info.n_owned, rp, ci.astype(np.int32)))
        cpu = {"value": N / tcpu, "unit": UNIT, "cores": cores, "kind": "port", "extrapolated": bool(d["extrapolated"]),
               "sample": "oracle (C restatement, OpenMP) on the full mesh: 2 residual + 1 Jacobian timed in full, CG timed for %d "
                         "iterations and extrapolated to %d iterations (%s)" % (
                             d["cg_its_timed"], d["cg_its_extrapolated_to"],
                             "the oracle's own converged count" if its_oracle else "the count the GPU solve took"),
               "detail": d}
        # full-size parity of the assembled operator: GPU F(0), J(0) against the oracle's, computed just above
        F0, u0 = mesh.create_global_vector(), mesh.create_global_vector()
        mesh.residual(0, u0.set(0.0), F0)
        mesh.jacobian(0, u0, J)
        Fd, vd = F0.to_host(), J.values()
        Fo, vo_ = keep["F"], keep["vals"]
        check["F_max_abs_diff_over_max_abs"] = float(np.abs(Fd - Fo).max() / np.abs(Fo).max())
        rowmax = np.maximum.reduceat(np.abs(vo_), rp[:-1].astype(np.int64))
        check["J_max_row_relative_diff"] = float((np.maximum.reduceat(np.abs(vd - vo_), rp[:-1].astype(np.int64)) / rowmax).max())
        check["assembly_tol"] = 1e-12
        ok = ok and check["F_max_abs_diff_over_max_abs"] <= 1e-12 and check["J_max_row_relative_diff"] <= 1e-12
        if dim == 2:  # same-diagonal box: the hypotenuse couplings are sums of two exact zeros and must stay exactly 0.0
            check["J_explicit_zeros_match"] = bool(np.array_equal(vd == 0.0, vo_ == 0.0))
            ok = ok and check["J_explicit_zeros_match"]
        F0.destroy(); u0.destroy()
    check["ok"] = bool(ok)

    # ---- time to solution with the polynomial preconditioner, beside the metric: the same linear system
    #      J(0) y = F(0) solved by plain CG (the metric's solver) and by CG preconditioned with the Chebyshev(16)/Jacobi
    #      polynomial (-pc_type ksp -ksp_ksp_type chebyshev -ksp_ksp_max_it 16); device time from the solver's own CUDA
    #      events, second of two runs.  One GPU only; never part of a timed region; a failure here cannot touch the line. ----
    tts = None
    if world == 1 and not os.environ.get("OHP_BENCH_NO_TTS"):
        try:
            Fv, uv, yv = mesh.create_global_vector(), mesh.create_global_vector(), mesh.create_global_vector()
            mesh.residual(0, uv.set(0.0), Fv)
            mesh.jacobian(0, uv, J)
            runs = {}
            for label, kw in (("cg_pc_none", dict(pc=ohp.PC_NONE)), ("cg_pc_chebyshev16_jacobi", dict(pc=ohp.PC_CHEBYSHEV, cheb_its=16))):
                best = None
                for _ in range(2):
                    r_ = ohp.ksp_cg(J, Fv, yv, ohp.ksp_options(rtol=KSP_RTOL, max_it=KSP_MAXIT, **kw))
                    if best is None or r_["total_ms"] < best["total_ms"]:
                        best = r_
                runs[label] = (best, yv.to_host().copy())
            (a, ya), (b, yb) = runs["cg_pc_none"], runs["cg_pc_chebyshev16_jacobi"]
            tts = {"system": "J(0) y = F(0), rtol %g" % KSP_RTOL,
                   "cg_pc_none": {"ms": a["total_ms"], "iterations": a["its"], "operator_applications": a["its"], "reason": a["reason"]},
                   "cg_pc_chebyshev16_jacobi": {"ms": b["total_ms"], "iterations": b["its"], "operator_applications": b["spmvs"],
                                                "reason": b["reason"], "interval": [b["cheb_emin"], b["cheb_emax"]]},
                   "speedup": a["total_ms"] / b["total_ms"] if b["total_ms"] else None,
                   "reductions_ratio": a["its"] / b["its"] if b["its"] else None,
                   "max_rel_diff_of_solutions": float(np.abs(ya - yb).max(initial=0.0) / max(np.abs(ya).max(initial=0.0), 1e-300))}
            Fv.destroy(); uv.destroy(); yv.destroy()
        except Exception as e:  # noqa: BLE001 -- diagnostic extra, must not break the contract line
            tts = {"error": repr(e)}

    out = dict(workload=workload, sizes=sizes, dim=dim, n_cells=n_cells_global, n_verts=n_verts_global, N=N, nnz=int(info.nnz), tts=tts,
               value=value, ms_step=ms_step, its_step=its_step, launches=int(launches), clocks=clocks, roofline=roof,
               kernels=kernels, cpu=cpu, e2e=e2e, check=check, setup_ms=t_setup * 1e3, n_owned=int(info.n_owned))
    J.destroy(); u.destroy(); mesh.destroy()
    return out


def run_ours(args):
    job = Job(args)
    dist, world, rank = job.dist, job.world, job.rank
    # the sampler starts BEFORE the warm-up: the first nvidia-smi query on a fresh box stalls the driver for
    # seconds, which must not land inside a timed region; only samples taken after mark() are reported
    sampler = make_sampler(args.sampler, job.local_rank, args.sampler_ms) if rank == 0 else None
    e2e_steps = max(1, args.e2e_steps)
    main_ = measure_workload(job, args, args.workload, args.steps, args.warmup, e2e_steps, sampler, with_cpu=not args.no_cpu)
    second = None
    if args.secondary and args.secondary != "none" and not args.n:
        second = measure_workload(job, args, args.secondary, args.secondary_steps, max(3, args.warmup), min(3, e2e_steps), sampler, with_cpu=False)
    if sampler:
        sampler.stop()
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        job.ctx.close()
        return
    r = main_
    par = "1 GPU" if world == 1 else "%d-way chain partition; halo push + dot all-reduce %s" % (
        world, "over NCCL" if os.environ.get("OHP_COMM") == "nccl" else "fused into the CG kernels over NVLink peer memory")
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_string(r["workload"], r["sizes"], r["dim"], r["n_cells"]), "sizes": list(r["sizes"]),
                       "n_verts": r["n_verts"], "N": r["N"], "nnz_rank0": r["nnz"], "cg_iterations_per_step": int(round(r["its_step"])),
                       "parallelism": par,
                       "l2": "inputs larger than L2 (matrix %.0f MB + vectors %.0f MB per rank vs 126 MB)" % (12 * r["nnz"] / 1e6, 40 * r["n_owned"] / 1e6),
                       "setup_ms_excluded": r["setup_ms"]},
            "e2e": r["e2e"], "gpu_launches": r["launches"], "clocks": r["clocks"], "roofline": r["roofline"], "kernels": r["kernels"],
            "check": r["check"], "cpu_baseline": r["cpu"]}
    if r.get("tts"):
        line["time_to_solution"] = r["tts"]
    if second:
        s = second
        line["secondary"] = {"metric": METRIC, "value": s["value"], "unit": UNIT, "steps": args.secondary_steps, "ms_per_step": s["ms_step"],
                             "config": {"workload": workload_string(s["workload"], s["sizes"], s["dim"], s["n_cells"]), "sizes": list(s["sizes"]),
                                        "N": s["N"], "nnz_rank0": s["nnz"], "cg_iterations_per_step": int(round(s["its_step"])), "parallelism": par,
                                        "setup_ms_excluded": s["setup_ms"]},
                             "e2e": s["e2e"], "gpu_launches": s["launches"], "clocks": s["clocks"], "roofline": s["roofline"],
                             "kernels": s["kernels"], "check": s["check"]}
        if s.get("tts"):
            line["secondary"]["time_to_solution"] = s["tts"]
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    job.ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="box2d_2896", choices=sorted(WORKLOADS))
    ap.add_argument("--secondary", default="box3d_110", help="second workload measured into the `secondary` object ('none' to skip)")
    ap.add_argument("--secondary-steps", type=int, default=60)  # ~2 s of timed steps: the clock sampler needs a few samples
    ap.add_argument("--n", type=int, default=0, help="override the box resolution (development only)")
    ap.add_argument("--check-every", type=int, default=0)
    ap.add_argument("--profile", type=int, default=64, help="SpMV launches bracketed by CUDA events per solve")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-cg-its", type=int, default=200, help="--impl reference: CG iterations timed per step")
    ap.add_argument("--cpu-baseline-cg-its", type=int, default=100, help="cpu_baseline leg of the GPU arm: CG iterations timed")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--sampler", default="smi", choices=["smi", "nvml", "none"], help="how clocks / throttle reasons are sampled during the timed region")
    ap.add_argument("--sampler-ms", type=int, default=500)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
