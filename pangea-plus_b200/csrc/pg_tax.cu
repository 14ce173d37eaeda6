// pg_tax.cu -- Stage B (placeholder until the lineage kernels land).
#include "pg_internal.cuh"
extern "C" int pg_tax_build(const char *) { return pg_fail(NULL, PG_EINVAL, "pg_tax_build: not built yet"); }
extern "C" int pg_tax_load(pg_ctx *ctx, const char *, pg_tax **) { return pg_fail(ctx, PG_EINVAL, "pg_tax_load: not built yet"); }
extern "C" void pg_tax_free(pg_tax *) {}
extern "C" int pg_tax_leaf(pg_ctx *ctx, const pg_tax *, const int32_t *, int64_t, int32_t *) { return pg_fail(ctx, PG_EINVAL, "not built yet"); }
extern "C" int pg_tax_lineage(pg_ctx *ctx, const pg_tax *, const int32_t *, int64_t, char *, int64_t, int64_t *) { return pg_fail(ctx, PG_EINVAL, "not built yet"); }
