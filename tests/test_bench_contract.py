"""CPU test (-m "not gpu") of bench.py's reference arm and JSON contract: `--impl reference` must run
without a GPU (it times the oracle on the host cores) and print ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def run(args, env=None):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, p.stdout + p.stderr
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run(["--impl", "reference", "--n", "96", "--steps", "1", "--warmup", "1"])
    assert d["impl"] == "reference" and d["metric"] == "poisson_p1_assembly_plus_cg_dof_per_s" and d["unit"] == "DOF/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["n_gpus"] == 1 and (d["steps"], d["warmup"]) == (1, 1)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "CG" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # the oracle uses every host core whatever OMP_NUM_THREADS says (torchrun exports 1), and says how much it measured
    import oracle
    assert cb["cores"] == oracle.use_all_cores()
    assert isinstance(d["extrapolated"], bool) and 0.0 < d["measured_fraction"] <= 1.0


def test_reference_arm_ignores_torchrun_omp_setting():
    d = run(["--impl", "reference", "--n", "64", "--steps", "1", "--warmup", "0"], env={"OMP_NUM_THREADS": "1"})
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_both_arms_print_the_same_workload_string():
    import bench
    a = bench.workload_string("box2d_2896", (2896, 2896), 2, 16773632)
    assert "box2d_2896" in a and "16773632 cells" in a and "rtol 1e-09" in a
    # nothing run-dependent (iteration counts, nnz, rank counts) may enter the string the driver compares
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"workload": workload_string(') == 3   # reference arm, GPU arm, secondary


def test_reference_arm_nonzero_ranks_stay_silent():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "64", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_oracle_iteration_counts_are_recorded():
    """The CPU arm extrapolates to the ORACLE's own converged counts (tools/oracle_cg_iterations.py), not to
    anything the GPU arm measured."""
    import bench
    assert bench.oracle_iterations("box2d_2896", (2896, 2896)) == 7643
    assert bench.oracle_iterations("box3d_110", (110, 110, 110)) == 386
    assert bench.oracle_iterations("box2d_2896", (64, 64)) is None


def test_morton_workload_is_the_same_mesh_renumbered():
    import numpy as np
    import oracle
    from omegahpetsc_b200 import meshgen
    c, x = meshgen.box2d(24, 17)
    c2, x2 = meshgen.morton_reorder(c, x, (24, 17))
    assert sorted(map(tuple, np.round(x2, 12))) == sorted(map(tuple, np.round(x, 12)))
    tri = lambda cc, xx: sorted(tuple(sorted(map(tuple, np.round(xx[t], 12)))) for t in cc)
    assert tri(c, x) == tri(c2, x2)
    X = x2[c2]
    area2 = (X[:, 1, 0] - X[:, 0, 0]) * (X[:, 2, 1] - X[:, 0, 1]) - (X[:, 1, 1] - X[:, 0, 1]) * (X[:, 2, 0] - X[:, 0, 0])
    assert (area2 > 0).all()
    b = oracle.mark_boundary(c2, x2.shape[0])
    g, N = oracle.numbering(b)
    rp, ci = oracle.pattern(c2, g, N)
    rows = np.repeat(np.arange(N), np.diff(rp))
    assert np.abs(ci - rows).max() > 3 * 25   # no longer banded by the grid width


def test_hilbert_workload_is_the_same_mesh_renumbered_along_a_continuous_curve():
    import numpy as np
    import oracle
    from omegahpetsc_b200 import meshgen
    # the curve itself: a bijection on the 2^b x 2^b grid whose consecutive points are grid neighbours
    for bits in (1, 2, 3, 5):
        n = 1 << bits
        i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        d = meshgen.hilbert_index(i.reshape(-1), j.reshape(-1), bits)
        assert np.array_equal(np.sort(d), np.arange(n * n))
        order = np.argsort(d)
        step = np.abs(np.diff(i.reshape(-1)[order])) + np.abs(np.diff(j.reshape(-1)[order]))
        assert (step == 1).all()
    # the workload: same vertices, same triangles, positive orientation kept, numbering no longer banded
    c, x = meshgen.box2d(24, 17)
    c2, x2 = meshgen.hilbert_reorder(c, x, (24, 17))
    assert sorted(map(tuple, np.round(x2, 12))) == sorted(map(tuple, np.round(x, 12)))
    tri = lambda cc, xx: sorted(tuple(sorted(map(tuple, np.round(xx[t], 12)))) for t in cc)
    assert tri(c, x) == tri(c2, x2)
    X = x2[c2]
    area2 = (X[:, 1, 0] - X[:, 0, 0]) * (X[:, 2, 1] - X[:, 0, 1]) - (X[:, 1, 1] - X[:, 0, 1]) * (X[:, 2, 0] - X[:, 0, 0])
    assert (area2 > 0).all()
    b = oracle.mark_boundary(c2, x2.shape[0])
    g, N = oracle.numbering(b)
    rp, ci = oracle.pattern(c2, g, N)
    rows = np.repeat(np.arange(N), np.diff(rp))
    assert np.abs(ci - rows).max() > 3 * 25
    # the operator is the same up to the permutation: same CG iteration count, same solution at the same points
    vals = oracle.jacobian(0, c2, x2, g, N, np.zeros(N), rp, ci)
    F = oracle.residual(0, c2, x2, g, N, np.zeros(N))
    b0 = oracle.mark_boundary(c, x.shape[0])
    g0, N0 = oracle.numbering(b0)
    rp0, ci0 = oracle.pattern(c, g0, N0)
    r0 = oracle.cg(rp0, ci0, oracle.jacobian(0, c, x, g0, N0, np.zeros(N0), rp0, ci0), oracle.residual(0, c, x, g0, N0, np.zeros(N0)), rtol=1e-10)
    r2 = oracle.cg(rp, ci, vals, F, rtol=1e-10)
    assert N == N0 and abs(r2["its"] - r0["its"]) <= 1
    key = lambda xx, gg: {tuple(np.round(p, 12)): k for p, k in zip(xx, gg) if k >= 0}
    k0, k2 = key(x, g0), key(x2, g)
    assert max(abs(r0["x"][k0[p]] - r2["x"][k2[p]]) for p in k0) < 1e-8
