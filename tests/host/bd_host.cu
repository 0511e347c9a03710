/*
 * TEST HARNESS (not product code): runs the product's natural-boundary row function (ohp::bd_residual_row in
 * include/ohp_fem.cuh, a __host__ __device__ function free of CUDA intrinsics) on HOST arrays, so that the arithmetic the
 * device kernel k_bd_residual_rows executes per row can be compared with the CPU restatement on a machine without a GPU.
 * Built by tests/test_next_rows_cpu.py with nvcc (host code only is executed) into tests/_build/.
 */
#include <string.h>

#include "ohp_fem.cuh"

namespace {
/* ex12's boundary callbacks (ex12.cpp:179-195), as the library's built-in problems carry them */
struct BdEx12 {
  OHP_HD static void f0_bd(OHP_BD_ARGS, double f0[]) {
    f0[0] = 0.0;
    for (int d = 0; d < dim; ++d) f0[0] += -n[d] * 2.0 * x[d];
  }
  OHP_HD static void f1_bd(OHP_BD_ARGS, double f1[]) { for (int d = 0; d < dim; ++d) f1[d] = 0.0; }
  OHP_HD static int bc(int dim, double, const double x[], int, double *u, void *) {
    *u = dim == 3 ? 2.0 * (x[0] * x[0] + x[1] * x[1] + x[2] * x[2]) / 3.0 : x[0] * x[0] + x[1] * x[1];
    return 0;
  }
};
/* the synthetic probe set of the CPU restatement: reads u, grad u, x, n; non-zero f1 */
struct BdProbe : BdEx12 {
  OHP_HD static void f0_bd(OHP_BD_ARGS, double f0[]) {
    f0[0] = 0.5 * u[0];
    for (int d = 0; d < dim; ++d) f0[0] += n[d] * u_x[d] + (d + 1) * x[d] * n[d];
  }
  OHP_HD static void f1_bd(OHP_BD_ARGS, double f1[]) { for (int d = 0; d < dim; ++d) f1[d] = u[0] * n[d] + 0.25 * x[d]; }
};

template <int DIM, class P> void rows(const ohp_assembly_view &vw, double *out) {
  for (int r = 0; r < vw.n_rows; ++r) {
    const int v = vw.row_vertex_dev[r];
    if (!vw.bnd_dev[v]) continue; /* what k_bd_residual_rows does */
    out[r] += ohp::bd_residual_row<DIM, P>(vw, vw.v2c_ptr_dev[v], vw.v2c_ptr_dev[v + 1]);
  }
}
}  // namespace

extern "C" int bd_host_rows(int bd_set, int dim, int n_cells, int n_verts, const int *cells, const double *coords, const int *lidx,
                            const double *u, const int *v2c_ptr, const int *v2c, int n_rows, const int *row_vertex,
                            const unsigned char *bnd, double *out) {
  ohp_assembly_view vw;
  memset(&vw, 0, sizeof(vw));
  vw.dim = dim; vw.n_cells = n_cells; vw.n_verts = n_verts; vw.n_rows = n_rows;
  vw.cells_dev = cells; vw.coords_dev = coords; vw.lidx_dev = lidx; vw.row_vertex_dev = row_vertex;
  vw.v2c_ptr_dev = v2c_ptr; vw.v2c_dev = v2c; vw.u_dev = u; vw.out_dev = out; vw.bnd_dev = bnd;
  if (dim == 2) { if (bd_set) rows<2, BdProbe>(vw, out); else rows<2, BdEx12>(vw, out); }
  else if (dim == 3) { if (bd_set) rows<3, BdProbe>(vw, out); else rows<3, BdEx12>(vw, out); }
  else return 1;
  return 0;
}
