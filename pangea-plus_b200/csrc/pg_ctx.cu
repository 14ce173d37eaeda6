// pg_ctx.cu -- context, error reporting, scratch memory.
#include "pg_internal.cuh"

static char g_init_err[512] = "";

int pg_fail(const pg_ctx *ctx, int code, const char *fmt, ...)
{
    char *dst = ctx ? ctx->err : g_init_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *pg_last_error(const pg_ctx *ctx) { return ctx ? ctx->err : g_init_err; }

extern "C" pg_ctx *pg_init(int device)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        pg_fail(NULL, PG_ENODEV,
                "pg_init: no CUDA device (%s); libpangea_b200 has no CPU path",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return NULL;
    }
    if (device < 0 || device >= ndev) {
        pg_fail(NULL, PG_EINVAL, "pg_init: device %d out of range (0..%d)", device, ndev - 1);
        return NULL;
    }
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        pg_fail(NULL, PG_ECUDA, "pg_init: %s", cudaGetErrorString(e));
        return NULL;
    }
    if (prop.major != 10) {
        pg_fail(NULL, PG_ENODEV,
                "pg_init: device %d is sm_%d%d; this library carries sm_100a code only",
                device, prop.major, prop.minor);
        return NULL;
    }
    pg_ctx *ctx = new pg_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    ctx->err[0] = 0;
    ctx->launches = 0;
    ctx->d_boot_pool = NULL;
    ctx->boot_cap = ctx->boot_used = 0;
    ctx->d_boot_off = NULL;
    ctx->boot_min_words = -1;
    ctx->classify_ms = 0.0;
    ctx->classify_launches = 0;
    ctx->st_certified = ctx->st_strict = ctx->st_handed_back = 0;
    ctx->st_heavy = ctx->st_items = 0;
    ctx->st_mma = 0;
    memset(&ctx->s_words, 0, sizeof(pg_ctx::Scratch) * pg_ctx::kNumScratch);
    ctx->h_pin = NULL;
    ctx->h_pin_cap = 0;
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        pg_fail(NULL, PG_ECUDA, "pg_init: cudaStreamCreate: %s", cudaGetErrorString(e));
        delete ctx;
        return NULL;
    }
    ctx->stream = ctx->own_stream;
    ctx->copy_stream = ctx->down_stream = NULL;
    {
        // keep the stream-ordered allocator's pool warm: transient buffers (pg_dev_alloc) come back without a trip
        // to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        (void)cudaGetLastError();
    }
    return ctx;
}

extern "C" void pg_shutdown(pg_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    pg_ctx::Scratch *s = &ctx->s_words;
    for (int i = 0; i < pg_ctx::kNumScratch; i++) cudaFree(s[i].p);
    cudaFree(ctx->d_boot_pool);
    cudaFree(ctx->d_boot_off);
    cudaFree(ctx->d_cnt_img);
    if (ctx->aux_stream) { cudaStreamDestroy(ctx->aux_stream); for (int i = 0; i < 4; i++) cudaEventDestroy(ctx->ev_pipe[i]); }
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    for (auto &p : ctx->ev_pending) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    for (auto &ev : ctx->ev_free) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->down_stream) cudaStreamDestroy(ctx->down_stream);
    delete ctx;
}

extern "C" int pg_set_stream(pg_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return PG_EINVAL;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return PG_OK;
}

extern "C" int pg_sync(pg_ctx *ctx)
{
    if (!ctx) return PG_EINVAL;
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PG_OK;
}

extern "C" int pg_classify_stats2(const pg_ctx *ctx, int64_t *heavy_reads, int64_t *items)
{
    if (!ctx) return PG_EINVAL;
    if (heavy_reads) *heavy_reads = ctx->st_heavy;
    if (items) *items = ctx->st_items;
    return PG_OK;
}

extern "C" int pg_classify_stats3(const pg_ctx *ctx, int64_t *tensor_core_reads)
{
    if (!ctx) return PG_EINVAL;
    if (tensor_core_reads) *tensor_core_reads = ctx->st_mma;
    return PG_OK;
}

extern "C" int pg_classify_stats(const pg_ctx *ctx, int64_t *certified_reads, int64_t *strict_reads, int64_t *handed_back)
{
    if (!ctx) return PG_EINVAL;
    if (certified_reads) *certified_reads = ctx->st_certified;
    if (strict_reads) *strict_reads = ctx->st_strict;
    if (handed_back) *handed_back = ctx->st_handed_back;
    return PG_OK;
}

extern "C" int64_t pg_launch_count(const pg_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int pg_kernel_time(pg_ctx *ctx, double *classify_ms, int64_t *classify_launches)
{
    if (!ctx) return PG_EINVAL;
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &p : ctx->ev_pending) {
        float ms = 0.f;
        PG_CUDA(ctx, cudaEventElapsedTime(&ms, p.first, p.second));
        ctx->classify_ms += (double)ms;
        ctx->classify_launches++;
        ctx->ev_free.push_back(p.first);
        ctx->ev_free.push_back(p.second);
    }
    ctx->ev_pending.clear();
    if (classify_ms) *classify_ms = ctx->classify_ms;
    if (classify_launches) *classify_launches = ctx->classify_launches;
    return PG_OK;
}

extern "C" int pg_kernel_time_reset(pg_ctx *ctx)
{
    if (!ctx) return PG_EINVAL;
    PG_TRY(pg_kernel_time(ctx, NULL, NULL));
    ctx->classify_ms = 0.0;
    ctx->classify_launches = 0;
    return PG_OK;
}

int pg_scratch(pg_ctx *ctx, pg_ctx::Scratch *s, size_t bytes)
{
    if (bytes <= s->cap) return PG_OK;
    // Grow-only.  The old block may still be in use by work queued on the
    // stream, so drain it before freeing.
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (s->p) PG_CUDA(ctx, cudaFree(s->p));
    s->p = NULL;
    s->cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&s->p, want);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();   // clear the sticky last-error
        return pg_fail(ctx, PG_ENOMEM, "device allocation of %zu bytes failed: %s", want,
                       cudaGetErrorString(e));
    }
    s->cap = want;
    return PG_OK;
}

int pg_pinned(pg_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->h_pin_cap) return PG_OK;
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_pin) PG_CUDA(ctx, cudaFreeHost(ctx->h_pin));
    ctx->h_pin = NULL;
    ctx->h_pin_cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMallocHost(&ctx->h_pin, want);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return pg_fail(ctx, PG_ENOMEM, "pinned allocation of %zu bytes failed: %s", want,
                       cudaGetErrorString(e));
    }
    ctx->h_pin_cap = want;
    return PG_OK;
}

extern "C" void pg_free(void *p) { free(p); }
