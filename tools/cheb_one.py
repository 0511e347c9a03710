"""One application of the Chebyshev/Jacobi polynomial on the 16 M-triangle box and exit (target of an ncu capture of k_spmv_cheb)."""
import omegahpetsc_b200 as ohp
from omegahpetsc_b200 import meshgen

ctx = ohp.Context(0)
cells, coords = meshgen.box2d(2896, 2896)
m = ohp.Mesh(ctx, cells, coords).setup()
u, F, J, x = m.create_global_vector().set(0.0), m.create_global_vector(), m.create_matrix(), m.create_global_vector()
m.residual(0, u, F)
m.jacobian(0, u, J)
# explicit interval: no eigenvalue estimate; max_it 1: two polynomial applications of degree 4 = 6 launches of k_spmv_cheb
r = ohp.ksp_cg(J, F, x, ohp.ksp_options(rtol=1e-9, max_it=1, pc=ohp.PC_CHEBYSHEV, cheb_its=4, cheb_emin=0.02, cheb_emax=2.17))
print("its", r["its"], "reason", r["reason"], "spmvs", r["spmvs"], flush=True)
