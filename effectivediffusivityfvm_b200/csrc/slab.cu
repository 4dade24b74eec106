// slab.cu -- multi-GPU row-slab decomposition (placeholder: single-GPU contexts only).
#include "context.h"

namespace deff2d {
int slab_allreduce_q(deff2d_ctx *c) { (void)c; return DEFF2D_OK; }
void slab_destroy(deff2d_ctx *c) { (void)c; }
}  // namespace deff2d

DEFF2D_EXPORT int deff2d_nccl_unique_id(uint8_t id[DEFF2D_NCCL_ID_BYTES]) { (void)id; return DEFF2D_ERR_NCCL; }
DEFF2D_EXPORT int deff2d_nccl_init(deff2d_ctx *c, const uint8_t id[DEFF2D_NCCL_ID_BYTES], int rank, int nranks)
{
    (void)id; (void)rank; (void)nranks;
    deff2d::set_error(c, "NCCL slab mode not built yet");
    return DEFF2D_ERR_NCCL;
}
DEFF2D_EXPORT int deff2d_slab_sweeps(deff2d_ctx *c, int64_t n) { (void)n; deff2d::set_error(c, "NCCL slab mode not built yet"); return DEFF2D_ERR_NCCL; }
DEFF2D_EXPORT int deff2d_slab_flux(deff2d_ctx *c, double *d) { (void)d; deff2d::set_error(c, "NCCL slab mode not built yet"); return DEFF2D_ERR_NCCL; }
