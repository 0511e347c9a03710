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


def test_reference_arm_nonzero_ranks_stay_silent():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "64", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_known_iteration_counts_are_recorded():
    with open(os.path.join(ROOT, "profiles", "cg_iterations.json")) as f:
        k = json.load(f)
    assert k["box2d_2896:2896x2896"] == 7643 and k["box3d_110:110x110x110"] == 386
