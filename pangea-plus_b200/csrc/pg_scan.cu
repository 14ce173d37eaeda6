// pg_scan.cu -- exclusive prefix sum over int64 on the device (lengths -> offsets), shared by the
// lineage and trim paths.
#include "pg_internal.cuh"

// ------------------------------------------------------------------ device: exclusive scan (int64)

__global__ void k_scan_block(const int64_t *__restrict__ in, int64_t n, int64_t *__restrict__ out,
                             int64_t *__restrict__ block_sums)
{
    __shared__ int64_t s[1024];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    int64_t v = i < n ? in[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int64_t add = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
        __syncthreads();
        s[threadIdx.x] += add;
        __syncthreads();
    }
    if (i < n) out[i] = s[threadIdx.x] - v;                  // exclusive
    if (threadIdx.x == 1023) block_sums[blockIdx.x] = s[1023];
}
__global__ void k_scan_sums(int64_t *block_sums, int64_t nb, int64_t *total)
{
    // few thousand blocks at most: one thread is enough
    int64_t acc = 0;
    for (int64_t b = 0; b < nb; b++) { int64_t v = block_sums[b]; block_sums[b] = acc; acc += v; }
    *total = acc;
}
__global__ void k_scan_add(int64_t *__restrict__ out, int64_t n, const int64_t *__restrict__ block_sums)
{
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    if (i < n) out[i] += block_sums[blockIdx.x];
}

// out[0..n) = exclusive scan of in, out[n] = total (device arrays; out has n+1 entries)
int pg_device_scan(pg_ctx *ctx, const int64_t *d_in, int64_t n, int64_t *d_out)
{
    const int64_t nb = (n + 1023) / 1024;
    int64_t *d_bs = NULL;
    PG_CUDA(ctx, pg_dev_alloc(ctx, (void **)&d_bs, (size_t)(nb + 1) * 8));
    if (nb > 0) {
        k_scan_block<<<(unsigned)nb, 1024, 0, ctx->stream>>>(d_in, n, d_out, d_bs);
        PG_LAUNCHED(ctx);
    }
    k_scan_sums<<<1, 1, 0, ctx->stream>>>(d_bs, nb, d_out + n);
    PG_LAUNCHED(ctx);
    if (nb > 0) {
        k_scan_add<<<(unsigned)nb, 1024, 0, ctx->stream>>>(d_out, n, d_bs);
        PG_LAUNCHED(ctx);
    }
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    pg_dev_free(ctx, d_bs);
    return PG_OK;
}

