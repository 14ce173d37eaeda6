// pg_synth.cu -- seeded synthetic training members and reads, generated on the device (include/pangea_b200_synth.h).
// Bench / test support for BASELINE configs[3] (3 M sequences, 100 M reads): every byte is a pure function of
// (seed, record, position); pangea_b200/synth.py restates the same functions in numpy for the CPU oracle.
#include "pg_internal.cuh"
#include "pangea_b200_synth.h"

__host__ __device__ static inline uint64_t pg_mix64(uint64_t x)          // splitmix64 finaliser
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__host__ __device__ static inline uint64_t pg_h3(uint64_t seed, uint64_t stream, uint64_t a, uint64_t b)
{
    return pg_mix64(pg_mix64(pg_mix64(seed ^ (stream * 0xD6E8FEB86659FD93ULL)) + a) + b);
}

#define SY_SUB_MEMBER 42949673u       // 1 %   of 2^32
#define SY_N_MEMBER   1073742u        // 0.1 % of 2^30
#define SY_SUB_READ   21474836u       // 0.5 % of 2^32

// one thread per base of the slice: binary search of the record, then the three decisions from one 64-bit hash
__global__ void k_synth_members(uint64_t seed, const uint8_t *__restrict__ cent, int length, const int32_t *__restrict__ genus,
                                const int64_t *__restrict__ off, int64_t first, int64_t count, char *__restrict__ out)
{
    const int64_t total = off[count];
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = count - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (off[mid] <= idx) lo = mid; else hi = mid - 1;
        }
        const int64_t p = idx - off[lo];
        const uint64_t h = pg_h3(seed, 2, (uint64_t)(first + lo), (uint64_t)p);
        char c = (char)cent[(size_t)genus[lo] * length + p];
        if ((uint32_t)h < SY_SUB_MEMBER) c = "acgt"[(h >> 32) & 3];
        if ((uint32_t)(h >> 34) < SY_N_MEMBER) c = 'n';
        out[idx] = c;
    }
}

__device__ __forceinline__ char sy_comp(char c)
{
    switch (c) {
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    default: return c;
    }
}

// one warp per read
__global__ void __launch_bounds__(256)
k_synth_reads(uint64_t seed, const char *__restrict__ members, const int64_t *__restrict__ moff,
              const int32_t *__restrict__ mgenus, int64_t nmembers, int64_t first, int64_t count, int read_len, int gap,
              int paired, char *__restrict__ out, int32_t *__restrict__ src)
{
    const int lane = threadIdx.x & 31;
    const int span = paired ? 2 * read_len + gap : read_len;
    for (int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < count;
         j += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const uint64_t r = (uint64_t)(first + j);
        int64_t m = 0, len = 0;
        for (uint64_t attempt = 0; attempt < 64; attempt++) {          // members shorter than the span: draw again
            m = (int64_t)(pg_h3(seed, 1, r, attempt) % (uint64_t)nmembers);
            len = moff[m + 1] - moff[m];
            if (len >= span) break;
        }
        char *o = out + (size_t)j * span;
        if (len < span) {                                              // no member is long enough: an all-N record
            for (int p = lane; p < span; p += 32) o[p] = 'N';
            if (src && lane == 0) src[j] = -1;
            continue;
        }
        const int64_t start = (int64_t)(pg_h3(seed, 2, r, 0) % (uint64_t)(len - span + 1));
        const bool flip = (pg_h3(seed, 4, r, 0) & 1ULL) != 0ULL;
        const char *s = members + moff[m] + start;
        for (int p = lane; p < span; p += 32) {
            char c = s[p];
            const uint64_t h = pg_h3(seed, 3, r, (uint64_t)p);
            if ((uint32_t)h < SY_SUB_READ) c = "acgt"[(h >> 32) & 3];
            if (paired && p >= read_len && p < read_len + gap) c = 'N';
            if (flip) o[span - 1 - p] = sy_comp(c); else o[p] = c;
        }
        if (src && lane == 0) src[j] = mgenus[m];
    }
}

extern "C" int pg_synth_members(pg_ctx *ctx, uint64_t seed, const uint8_t *centroids_dev, int length, const int32_t *genus_dev,
                                const int64_t *off_dev, int64_t first, int64_t count, char *bytes_dev)
{
    if (!ctx || !centroids_dev || !genus_dev || !off_dev || !bytes_dev || count < 0 || length <= 0)
        return pg_fail(ctx, PG_EINVAL, "pg_synth_members: bad arguments");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return PG_OK;
    k_synth_members<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(seed, centroids_dev, length, genus_dev, off_dev, first, count, bytes_dev);
    PG_LAUNCHED(ctx);
    return PG_OK;
}

extern "C" int pg_synth_reads(pg_ctx *ctx, uint64_t seed, const char *members_dev, const int64_t *member_off_dev,
                              const int32_t *member_genus_dev, int64_t nmembers, int64_t first, int64_t count, int read_len,
                              int gap, int paired, char *out_dev, int32_t *src_genus_dev)
{
    if (!ctx || !members_dev || !member_off_dev || !member_genus_dev || !out_dev || nmembers <= 0 || count < 0 || read_len <= 0 || gap < 0)
        return pg_fail(ctx, PG_EINVAL, "pg_synth_reads: bad arguments");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return PG_OK;
    k_synth_reads<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(seed, members_dev, member_off_dev, member_genus_dev, nmembers, first,
                                                             count, read_len, gap, paired, out_dev, src_genus_dev);
    PG_LAUNCHED(ctx);
    return PG_OK;
}
