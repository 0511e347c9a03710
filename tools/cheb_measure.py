"""Time-to-solution of the Chebyshev/Jacobi polynomial preconditioner against plain CG on the bench workloads
(one linear solve J(0) y = F(0), rtol 1e-9, device time from the solver's own CUDA events).  Writes one JSON line per
configuration.  Usage: python tools/cheb_measure.py [out.json]"""
import json
import sys
import time

import numpy as np

import omegahpetsc_b200 as ohp
from omegahpetsc_b200 import meshgen


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02_chebyshev_time_to_solution.json"
    ctx = ohp.Context(0)
    lines = []
    for name, (cells, coords) in (("box2d_2896", meshgen.box2d(2896, 2896)), ("box3d_110", meshgen.box3d(110, 110, 110))):
        m = ohp.Mesh(ctx, cells, coords).setup()
        u, F, J, x = m.create_global_vector().set(0.0), m.create_global_vector(), m.create_matrix(), m.create_global_vector()
        m.residual(0, u, F)
        m.jacobian(0, u, J)
        ref = None
        for label, kw in (("cg pc none (fused recurrences)", dict(pc=ohp.PC_NONE)),
                          ("cg pc jacobi (fused recurrences)", dict(pc=ohp.PC_JACOBI)),
                          ("cg pc chebyshev k=2", dict(pc=ohp.PC_CHEBYSHEV, cheb_its=2)),
                          ("cg pc chebyshev k=4", dict(pc=ohp.PC_CHEBYSHEV, cheb_its=4)),
                          ("cg pc chebyshev k=8", dict(pc=ohp.PC_CHEBYSHEV, cheb_its=8)),
                          ("cg pc chebyshev k=16", dict(pc=ohp.PC_CHEBYSHEV, cheb_its=16))):
            best = None
            for rep in range(2):  # second run: pools and caches warm
                t0 = time.perf_counter()
                r = ohp.ksp_cg(J, F, x, ohp.ksp_options(rtol=1e-9, max_it=1000000, **kw))
                wall = (time.perf_counter() - t0) * 1e3
                if best is None or r["total_ms"] < best[0]["total_ms"]:
                    best = (r, wall)
            r, wall = best
            xh = x.to_host()
            if ref is None:
                ref = xh.copy()
            rec = dict(workload=name, N=int(m.info().n_global), solver=label, its=r["its"], reason=r["reason"], algo=r["algo"],
                       operator_applications=(r["spmvs"] if r["spmvs"] else r["its"]), device_ms=r["total_ms"], wall_ms=wall,
                       cheb_interval=[r["cheb_emin"], r["cheb_emax"]],
                       max_rel_diff_vs_plain_cg=float(np.abs(xh - ref).max() / np.abs(ref).max()))
            print(json.dumps(rec), flush=True)
            lines.append(rec)
        m.destroy()
    with open(out, "w") as f:
        for rec in lines:
            f.write(json.dumps(rec) + "\n")
    ctx.close()


if __name__ == "__main__":
    main()
