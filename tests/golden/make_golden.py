"""Regenerates tests/golden/: copies the reference's mesh FIXTURES (binary input data, not
source) out of /root/reference and records what the oracle computes on them.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

The reference itself pins no outputs (its CTests check exit status only,
CMakeLists.txt:66-120) and cannot be built here (no PETSc/Omega_h/MPI), so these vectors are
ORACLE outputs, cross-checked in tests/test_oracle.py against an independent scipy assembly
and the analytic anchors of ex12.cpp:88-96 -- "parity unpinned" in the task's sense.
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from oracle import osh_oracle as oo  # noqa: E402

REF = "/root/reference"
COPIES = {
    "23elmMesh/mesh.osh": "meshes/23elm.osh",
    "23elmMesh/23elm_bfs_bfs_1.ppm/23elm_bfs_bfs_0.osh": "meshes/23elm_1p_0.osh",
    "23elmMesh/23elm_bfs_bfs_2.ppm/23elm_bfs_bfs_0.osh": "meshes/23elm_2p_0.osh",
    "23elmMesh/23elm_bfs_bfs_2.ppm/23elm_bfs_bfs_1.osh": "meshes/23elm_2p_1.osh",
    "24k.osh": "meshes/24k.osh",
    "xgc-picparts/24k/4p/picpart0.osh": "meshes/24k_4p_0.osh",
    "xgc-picparts/24k/4p/picpart1.osh": "meshes/24k_4p_1.osh",
    "xgc-picparts/24k/4p/picpart2.osh": "meshes/24k_4p_2.osh",
    "xgc-picparts/24k/4p/picpart3.osh": "meshes/24k_4p_3.osh",
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def describe(cells, coords, problem=0, rtol=1e-9):
    nV = coords.shape[0]
    bnd = oracle.mark_boundary(cells, nV)
    gidx, N = oracle.numbering(bnd)
    rowptr, colidx = oracle.pattern(cells, gidx, N)
    u0 = np.zeros(N)
    F0 = oracle.residual(problem, cells, coords, gidx, N, u0)
    vals = oracle.jacobian(problem, cells, coords, gidx, N, u0, rowptr, colidx)
    cg = oracle.cg(rowptr, colidx, vals, F0, rtol=rtol, maxit=100000)
    cgj = oracle.cg(rowptr, colidx, vals, F0, rtol=rtol, maxit=100000, pc=1)
    nt = oracle.newton(problem, cells, coords, gidx, N, rowptr, colidx, ksp_rtol=rtol, ksp_maxit=100000)
    exact = (coords ** 2).sum(axis=1) * (1.0 if coords.shape[1] == 2 else 2.0 / 3.0)
    free = gidx >= 0
    return dict(
        n_verts=int(nV), n_cells=int(cells.shape[0]), n_boundary=int(bnd.sum()), N=int(N), nnz=int(colidx.size),
        max_row=int(np.diff(rowptr).max()) if N else 0,
        sha_bnd=sha(bnd.astype(np.uint8)), sha_gidx=sha(gidx.astype(np.int64)),
        sha_rowptr=sha(rowptr.astype(np.int32)), sha_colidx=sha(colidx.astype(np.int64)),
        F0_norm=float(np.linalg.norm(F0)), F0_sum=float(F0.sum()),
        vals_fro=float(np.linalg.norm(vals)), vals_sum=float(vals.sum()), vals_absmax=float(np.abs(vals).max()) if N else 0.0,
        cg_its_none=int(cg["its"]), cg_reason_none=int(cg["reason"]), cg_its_jacobi=int(cgj["its"]),
        newton_its=int(nt["snes_its"]), newton_ksp_its=int(nt["ksp_its"]), newton_reason=int(nt["reason"]),
        newton_fnorms=[float(x) for x in nt["norms"]],
        u_norm=float(np.linalg.norm(nt["u"])), u_max_nodal_err=float(np.abs(nt["u"] - exact[free]).max()) if N else 0.0,
    )


def main():
    expected = {}
    for src, dst in COPIES.items():
        d = os.path.join(HERE, dst)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(REF, src), d)
    for name, rel in (("T23", "meshes/23elm.osh"), ("T24k", "meshes/24k.osh")):
        m = oo.read_osh(os.path.join(HERE, rel))
        assert m["consumed"] == m["total"]
        cells, coords = oo.elem_verts(m), oo.coords(m)
        e = describe(cells, coords)
        e["sha_elem_verts"] = sha(cells.astype(np.int32))
        e["sha_coords"] = sha(coords)
        if name == "T23":
            bnd = oracle.mark_boundary(cells, coords.shape[0])
            gidx, N = oracle.numbering(bnd)
            rowptr, colidx = oracle.pattern(cells, gidx, N)
            e["rowptr"] = rowptr.tolist()
            e["colidx"] = colidx.tolist()
            e["interior_verts"] = np.nonzero(gidx >= 0)[0].tolist()
        expected[name] = e
    # picpart ownership summaries
    parts = {}
    for r in range(4):
        m = oo.read_osh(os.path.join(HERE, "meshes/24k_4p_%d.osh" % r))
        eo = m["tags"][2]["ownership"][1]
        vo = m["tags"][0]["ownership"][1]
        ev = oo.elem_verts(m)
        core = np.zeros(m["nverts"], bool)
        core[ev[eo == r].reshape(-1)] = True
        parts[str(r)] = dict(owned_elems=int((eo == r).sum()), owned_verts=int((vo == r).sum()),
                             core_verts=int(core.sum()), ghosts=int((core & (vo != r)).sum()),
                             sha_gids=sha(m["tags"][0]["gids"][1].astype(np.int64)))
    expected["T24k_4p"] = parts
    # small synthetic boxes
    for name, (cells, coords) in (("box2d_8x8", oracle.box2d(8, 8)), ("box2d_33x17", oracle.box2d(33, 17)),
                                   ("box3d_5x4x3", oracle.box3d(5, 4, 3))):
        expected[name] = describe(cells, coords)
    with open(os.path.join(HERE, "expected.json"), "w") as f:
        json.dump(expected, f, indent=1, sort_keys=True)
    print(json.dumps({k: {kk: vv for kk, vv in v.items() if not isinstance(vv, (list, dict)) and not kk.startswith("sha")}
                      for k, v in expected.items() if k != "T24k_4p"}, indent=1))
    print(expected["T24k_4p"])


if __name__ == "__main__":
    main()
