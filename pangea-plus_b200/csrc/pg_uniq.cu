// pg_uniq.cu -- first hit per read on the GPU (first half).
//
// Replaces Scripts/get_uniq.pl :34-40: a line is kept iff no earlier line has the same first column
// (`print OUT "$line" if ! $seen{$columns[0]}++` with `split('\t', $line)`; the `chomp` there works on $_, not on
// $line, so a line without a TAB keys on its text INCLUDING the newline).  Output keeps the input order.
//
//   k_uq_insert   thread per line: FNV-1a of the key, open-addressing table of line indices, atomicMin keeps
//                 the earliest line of every key (byte compare on a hash match)
//   k_uq_mark     thread per line: kept iff it is the line its key's slot holds; kept lengths for the scan
//   k_uq_copy     thread per kept line: bytes to their place in the output
#include "pg_internal.cuh"

int pg_index_lines(pg_ctx *ctx, const char *d_text, int64_t n, int64_t **d_start_out, int64_t *nlines_out);   // pg_trim.cu

__device__ __forceinline__ int uq_key_len(const char *s, int n)
{
    for (int i = 0; i < n; i++)
        if (s[i] == '\t') return i;
    return n;                                       // no TAB: the whole line, newline included
}
__device__ __forceinline__ uint64_t uq_hash(const char *s, int n)
{
    uint64_t h = 0xCBF29CE484222325ULL;
    for (int i = 0; i < n; i++) { h ^= (unsigned char)s[i]; h *= 0x100000001B3ULL; }
    h ^= h >> 33; h *= 0xFF51AFD7ED558CCDULL; h ^= h >> 33;
    return h;
}

__global__ void k_uq_insert(const char *__restrict__ t, const int64_t *__restrict__ start, int64_t nlines,
                            unsigned long long *__restrict__ table, uint64_t mask, int32_t *__restrict__ klen,
                            unsigned long long *__restrict__ slot_of)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlines) return;
    const int64_t a = start[i];
    const int n = (int)(start[i + 1] - a);
    const int kl = uq_key_len(t + a, n);
    klen[i] = kl;
    uint64_t slot = uq_hash(t + a, kl) & mask;
    for (;;) {
        unsigned long long cur = table[slot];
        if (cur == ~0ULL) {
            cur = atomicCAS(&table[slot], ~0ULL, (unsigned long long)i);
            if (cur == ~0ULL) break;                                  // claimed an empty slot
        }
        // cur = a line whose key owns this slot (any of them will do for the comparison)
        const int64_t b = start[cur];
        const int nb = (int)(start[cur + 1] - b);
        const int kb = uq_key_len(t + b, nb);
        bool same = kb == kl;
        for (int j = 0; same && j < kl; j++) same = t[a + j] == t[b + j];
        if (same) { atomicMin(&table[slot], (unsigned long long)i); break; }
        slot = (slot + 1) & mask;
    }
    slot_of[i] = slot;
}

__global__ void k_uq_mark(const int64_t *__restrict__ start, int64_t nlines, const unsigned long long *__restrict__ table,
                          const unsigned long long *__restrict__ slot_of, int64_t *__restrict__ keep_len,
                          int64_t *__restrict__ keep_flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlines) return;
    const bool keep = table[slot_of[i]] == (unsigned long long)i;
    keep_len[i] = keep ? start[i + 1] - start[i] : 0;
    keep_flag[i] = keep ? 1 : 0;
}

__global__ void k_uq_copy(const char *__restrict__ t, const int64_t *__restrict__ start, int64_t nlines,
                          const int64_t *__restrict__ keep_len, const int64_t *__restrict__ out_off,
                          const int64_t *__restrict__ rank, char *__restrict__ out, int64_t *__restrict__ kept_lines)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlines || keep_len[i] == 0) return;
    const int64_t a = start[i], n = keep_len[i];
    char *dst = out + out_off[i];
    for (int64_t j = 0; j < n; j++) dst[j] = t[a + j];
    if (kept_lines) kept_lines[rank[i]] = i;
}

extern "C" int pg_first_hits(pg_ctx *ctx, const char *text_host, int64_t len, char *out_host, int64_t out_cap,
                             int64_t *out_len, int64_t *kept_lines_host, int64_t lines_cap, int64_t *n_kept)
{
    if (!ctx || (!text_host && len > 0) || len < 0 || !out_len) return pg_fail(ctx, PG_EINVAL, "pg_first_hits: bad arguments");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    *out_len = 0;
    if (n_kept) *n_kept = 0;
    if (len == 0) return PG_OK;
    PG_TRY(pg_scratch(ctx, &ctx->s_bytes, (size_t)len + 16));
    char *d_text = (char *)ctx->s_bytes.p;
    PG_CUDA(ctx, cudaMemcpyAsync(d_text, text_host, (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    int64_t *d_start = NULL, nlines = 0;
    PG_TRY(pg_index_lines(ctx, d_text, len, &d_start, &nlines));
    int rc = PG_OK;
    uint64_t size = 16;
    while (size < (uint64_t)nlines * 2) size <<= 1;
    int64_t total = 0, kept = 0;
#define UQ_TRY(call) do { if ((rc = (call)) != PG_OK) goto done; } while (0)
#define UQ_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = pg_fail(ctx, PG_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); goto done; } } while (0)
    {
        const size_t nl1 = (size_t)nlines + 2;
        UQ_TRY(pg_scratch(ctx, &ctx->s_candl, size * 8));
        UQ_TRY(pg_scratch(ctx, &ctx->s_champ, nl1 * 8 * 6 + nl1 * 4));
        UQ_TRY(pg_scratch(ctx, &ctx->s_cand, (size_t)len + 16));
        unsigned long long *d_table = (unsigned long long *)ctx->s_candl.p;
        unsigned long long *d_slot = (unsigned long long *)ctx->s_champ.p;
        int64_t *d_klen64 = (int64_t *)(d_slot + nl1), *d_flag = d_klen64 + nl1, *d_off = d_flag + nl1, *d_rank = d_off + nl1;
        int64_t *d_kept = d_rank + nl1;
        int32_t *d_klen = (int32_t *)(d_kept + nl1);
        char *d_out = (char *)ctx->s_cand.p;
        const unsigned nb = (unsigned)((nlines + 255) / 256);
        UQ_CUDA(cudaMemsetAsync(d_table, 0xFF, size * 8, ctx->stream));
        k_uq_insert<<<nb, 256, 0, ctx->stream>>>(d_text, d_start, nlines, d_table, size - 1, d_klen, d_slot);
        k_uq_mark<<<nb, 256, 0, ctx->stream>>>(d_start, nlines, d_table, d_slot, d_klen64, d_flag);
        ctx->launches += 2;
        UQ_CUDA(cudaGetLastError());
        UQ_TRY(pg_device_scan(ctx, d_klen64, nlines, d_off));
        UQ_TRY(pg_device_scan(ctx, d_flag, nlines, d_rank));
        UQ_CUDA(pg_copy_sync(ctx, &total, d_off + nlines, 8, cudaMemcpyDeviceToHost));
        UQ_CUDA(pg_copy_sync(ctx, &kept, d_rank + nlines, 8, cudaMemcpyDeviceToHost));
        *out_len = total;
        if (n_kept) *n_kept = kept;
        if (total > out_cap || (kept_lines_host && kept > lines_cap)) {
            rc = pg_fail(ctx, PG_ERANGE, "pg_first_hits: %lld bytes / %lld lines kept, room for %lld / %lld", (long long)total,
                         (long long)kept, (long long)out_cap, (long long)lines_cap);
            goto done;
        }
        k_uq_copy<<<nb, 256, 0, ctx->stream>>>(d_text, d_start, nlines, d_klen64, d_off, d_rank, d_out,
                                               kept_lines_host ? d_kept : NULL);
        ctx->launches++;
        UQ_CUDA(cudaGetLastError());
        if (out_host && total) UQ_CUDA(cudaMemcpyAsync(out_host, d_out, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
        if (kept_lines_host && kept) UQ_CUDA(cudaMemcpyAsync(kept_lines_host, d_kept, (size_t)kept * 8, cudaMemcpyDeviceToHost, ctx->stream));
        UQ_CUDA(cudaStreamSynchronize(ctx->stream));
    }
done:
#undef UQ_TRY
#undef UQ_CUDA
    pg_dev_free(ctx, d_start);
    return rc;
}
