"""Assembly-only probe for ncu: B16M mesh, a few residual + Jacobian launches, event-timed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import omegahpetsc_b200 as ohp
from omegahpetsc_b200 import meshgen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2896
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = ohp.Context(0)
cells, coords = meshgen.box2d(n, n) if dim == 2 else meshgen.box3d(n, n, n)
m = ohp.Mesh(ctx, cells, coords).setup()
J, u, F = m.create_matrix(), m.create_global_vector(), m.create_global_vector()
st = torch.cuda.ExternalStream(ctx.stream())
for name, fn in (("residual", lambda: m.residual(0, u, F)), ("jacobian", lambda: m.jacobian(0, u, J))):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10): fn()
    e1.record(st); ctx.sync(); torch.cuda.synchronize()
    print(name, "%.4f ms" % (e0.elapsed_time(e1) / 10))
