"""Records what the oracle computes for the SURVEY 8(f) rank-4 slice (natural boundary condition, null space, Chebyshev/Jacobi
polynomial preconditioner) on the fixtures already under tests/golden/meshes and on two synthetic boxes, into
tests/golden/expected_next_rows.json.  Needs no GPU and no /root/reference:  python tests/golden/make_golden_next_rows.py

Like tests/golden/expected.json these are ORACLE outputs ("parity unpinned": the reference ships no result files); the oracle's
boundary integral is cross-checked against an independent numpy integration, its CG variants against scipy, in
tests/test_next_rows_cpu.py.  The file pins them against silent changes and gives the GPU tests fixed numbers to hit."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from oracle import osh_oracle as oo  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def describe(cells, coords):
    nV = coords.shape[0]
    gid = np.arange(nV, dtype=np.int32)
    rowptr, colidx = oracle.pattern(cells, gid, nV)
    F0 = oracle.bd_residual(cells, coords, gid, nV, np.zeros(nV), oracle.residual(0, cells, coords, gid, nV, np.zeros(nV)))
    vals = oracle.jacobian(0, cells, coords, gid, nV, np.zeros(nV), rowptr, colidx)
    out = dict(natural_N=int(nV), natural_nnz=int(colidx.size), natural_sha_rowptr=sha(rowptr.astype(np.int32)),
               natural_sha_colidx=sha(colidx.astype(np.int64)), natural_F0_norm=float(np.linalg.norm(F0)),
               natural_F0_sum=float(F0.sum()), natural_vals_fro=float(np.linalg.norm(vals)),
               natural_max_abs_row_sum=float(np.abs(oracle.spmv(rowptr, colidx, vals, np.ones(nV))).max()))
    if abs(F0.sum()) < 1e-10 * np.abs(F0).sum():  # consistent right-hand side (2-D): the null-space solve is well posed
        ns = oracle.pcg(rowptr, colidx, vals, F0, rtol=1e-9, pc=1, nullspace=True, maxit=100000)
        out.update(nullspace_its_jacobi=int(ns["its"]), nullspace_reason=int(ns["reason"]),
                   nullspace_x_norm_mean_removed=float(np.linalg.norm(ns["x"] - ns["x"].mean())))
    bnd = oracle.mark_boundary(cells, nV)
    gidx, N = oracle.numbering(bnd)
    rp, ci = oracle.pattern(cells, gidx, N)
    v = oracle.jacobian(0, cells, coords, gidx, N, np.zeros(N), rp, ci)
    F = oracle.residual(0, cells, coords, gidx, N, np.zeros(N))
    for k in (1, 4, 8, 16):
        r = oracle.pcg(rp, ci, v, F, rtol=1e-9, pc=2, cheb_its=k, maxit=100000)
        out["chebyshev_k%d" % k] = dict(its=int(r["its"]), reason=int(r["reason"]), emin=float(r["emin"]), emax=float(r["emax"]),
                                        x_norm=float(np.linalg.norm(r["x"])))
    return out


def main():
    expected = {}
    for name, rel in (("T23", "meshes/23elm.osh"), ("T24k", "meshes/24k.osh")):
        m = oo.read_osh(os.path.join(HERE, rel))
        expected[name] = describe(oo.elem_verts(m), oo.coords(m))
    expected["box2d_33x17"] = describe(*oracle.box2d(33, 17))
    expected["box3d_5x4x3"] = describe(*oracle.box3d(5, 4, 3))
    with open(os.path.join(HERE, "expected_next_rows.json"), "w") as f:
        json.dump(expected, f, indent=1, sort_keys=True)
    print("wrote expected_next_rows.json:", {k: (v["natural_N"], v["natural_nnz"], v["chebyshev_k8"]["its"]) for k, v in expected.items()})


if __name__ == "__main__":
    main()
