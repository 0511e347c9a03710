/*
 * ohp_comm.cu -- communicator, halo plans and halo exchange (stub while the serial path is
 * brought up; see the multi-GPU section of DESIGN.md).
 */
#include "ohp_internal.h"

struct ohp_halo { int dummy; };

extern "C" int ohp_comm_unique_id(void *) { OHP_FAIL(OHP_ERR_SUP, "ohp_comm_unique_id: not built yet"); }
extern "C" int ohp_comm_init(ohp_context, int, int, const void *) { OHP_FAIL(OHP_ERR_SUP, "ohp_comm_init: not built yet"); }
extern "C" int ohp_comm_rank_size(ohp_context ctx, int *rank, int *nranks) { *rank = ctx->rank; *nranks = ctx->nranks; return 0; }

int ohp_halo_build(ohp_mesh) { OHP_FAIL(OHP_ERR_SUP, "halo: not built yet"); }
void ohp_halo_free(ohp_halo *h) { delete h; }
int ohp_halo_exchange(ohp_mesh, double *) { OHP_FAIL(OHP_ERR_SUP, "halo: not built yet"); }
int ohp_allreduce_sum(ohp_context, double *, int) { OHP_FAIL(OHP_ERR_SUP, "allreduce: not built yet"); }
int ohp_vertex_halo_i32(ohp_mesh, int32_t *) { OHP_FAIL(OHP_ERR_SUP, "halo: not built yet"); }
int ohp_allgather_i64(ohp_context, int64_t, std::vector<int64_t> &) { OHP_FAIL(OHP_ERR_SUP, "allgather: not built yet"); }
