// pg_consensus.cu -- Stage C (placeholder until the vote kernel lands).
#include "pg_internal.cuh"
extern "C" int pg_consensus(pg_ctx *ctx, const pg_consensus_in *, int64_t *, int32_t *) { return pg_fail(ctx, PG_EINVAL, "pg_consensus: not built yet"); }
