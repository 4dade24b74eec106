// batch.cu -- K5: resident small-image batch solve (placeholder: falls back to per-image solves).
#include "context.h"

namespace deff2d {
int batch_resident_solve(deff2d_ctx *c, const uint8_t *gray, int count, int W, int H, const deff2d_params *p,
                         deff2d_result *results, double *fields)
{
    (void)c; (void)gray; (void)count; (void)W; (void)H; (void)p; (void)results; (void)fields;
    return 1;
}
}  // namespace deff2d
