#!/usr/bin/env python
"""Converged CG iteration counts of the ORACLE (CPU restatement) on the bench workloads, written to
tests/golden/oracle_cg_iterations.json.  bench.py --impl reference extrapolates its bounded CG sample to
these counts, so the CPU arm does not depend on anything the GPU arm measured.

  python tools/oracle_cg_iterations.py [workload ...]      # minutes on 8 cores for box2d_2896
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from bench import KSP_MAXIT, KSP_RTOL, WORKLOADS, make_mesh  # noqa: E402


def main():
    names = sys.argv[1:] or sorted(WORKLOADS)
    out_path = os.path.join(ROOT, "tests", "golden", "oracle_cg_iterations.json")
    out = {}
    if os.path.exists(out_path):
        with open(out_path) as f:
            out = json.load(f)
    cores = oracle.use_all_cores()
    for name in names:
        (cells, coords), sizes = make_mesh(name, 0)
        t0 = time.time()
        bnd = oracle.mark_boundary(cells, coords.shape[0], omp=True)
        gidx, N = oracle.numbering(bnd)
        rowptr, colidx = oracle.pattern(cells, gidx, N, omp=True)
        u = np.zeros(N)
        vals = oracle.jacobian(0, cells, coords, gidx, N, u, rowptr, colidx, omp=True)
        F = oracle.residual(0, cells, coords, gidx, N, u, omp=True)
        t1 = time.time()
        cg = oracle.cg(rowptr, colidx, vals, F, rtol=KSP_RTOL, maxit=KSP_MAXIT, omp=True)
        t2 = time.time()
        assert cg["reason"] == 2, cg["reason"]
        exact = (coords ** 2).sum(axis=1) * (1.0 if coords.shape[1] == 2 else 2.0 / 3.0)
        err = float(np.abs(-cg["x"] - exact[gidx >= 0]).max())  # Newton update from u = 0: u = -y
        out["%s:%s" % (name, "x".join(map(str, sizes)))] = {
            "its": int(cg["its"]), "reason": int(cg["reason"]), "N": int(N), "nnz": int(colidx.size),
            "max_nodal_error_vs_exact": err, "cores": cores, "setup_s": t1 - t0, "cg_s": t2 - t1}
        print(name, out["%s:%s" % (name, "x".join(map(str, sizes)))], flush=True)
        with open(out_path, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
