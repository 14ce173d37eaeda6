// pg_certified.cu -- certified-margin fast path (mode 1).  Placeholder: the
// quantised table is not derived yet and mode 1 is refused.
#include "pg_internal.cuh"

struct Bucket { int nmax; int lpr; int block; };

int pg_model_derive_quantised(pg_model *md)
{
    (void)md;
    return PG_OK;
}

int pg_classify_certified_launch(pg_ctx *ctx, const pg_model *, const Bucket &, unsigned, int,
                                 const uint16_t *, const int64_t *, const int32_t *, const int32_t *, int64_t, int,
                                 unsigned long long *)
{
    return pg_fail(ctx, PG_EINVAL, "classify mode 1 (certified) is not built yet");
}
