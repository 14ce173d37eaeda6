// pg_fasta.cu -- FASTA ingest on the GPU (the other half of the Trim join):
// the query file of `rdp_classifier -q <in.fa>` (README.md:119) goes to the device as text and comes out
// as the packed read store of pg_classify_packed plus, per record, where its header sits in the text.
//
// Same reading of the file as the host parser (host/pg_host_common.c pg_fasta_read), which the tests
// use as the yardstick: a line starting with '>' opens a record; every other line after the first header
// is sequence, with '\r' before the newline and any blanks inside the line dropped; lines before the
// first header are ignored; the id is the header up to its first blank.
//
//   k_fa_lines   thread per line: header flag, number of sequence bytes the line contributes
//   (two scans)  sequence bytes -> positions, header flags -> record numbers
//   k_fa_emit    thread per line: a header writes its record's offsets, a sequence line copies its bytes
//   then k_pack (pg_reads.cu) turns the compacted bytes into the 3-plane store.
#include "pg_internal.cuh"

int pg_index_lines(pg_ctx *ctx, const char *d_text, int64_t n, int64_t **d_start_out, int64_t *nlines_out);   // pg_trim.cu

__device__ __forceinline__ bool fa_space(int c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

// line j = [start[j], start[j+1]) incl. its newline; returns the end without "\n" and trailing '\r's
__device__ __forceinline__ void fa_line(const char *t, const int64_t *start, int64_t j, int64_t &a, int64_t &e)
{
    a = start[j];
    e = start[j + 1];
    if (e > a && t[e - 1] == '\n') e--;
    while (e > a && t[e - 1] == '\r') e--;
}

__global__ void k_fa_lines(const char *__restrict__ t, const int64_t *__restrict__ start, int64_t nlines,
                           int64_t *__restrict__ is_header, int64_t *__restrict__ seq_bytes)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nlines) return;
    int64_t a, e;
    fa_line(t, start, j, a, e);
    const bool hdr = start[j + 1] > start[j] && t[a] == '>';
    int64_t n = 0;
    if (!hdr)
        for (int64_t p = a; p < e; p++) n += fa_space((unsigned char)t[p]) ? 0 : 1;
    is_header[j] = hdr ? 1 : 0;
    seq_bytes[j] = n;
}

// lines before the first header contribute nothing: their byte counts are cleared before the scan
__global__ void k_fa_drop_preamble(const int64_t *__restrict__ rec_before, int64_t nlines, int64_t *__restrict__ seq_bytes)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nlines) return;
    if (rec_before[j] == 0) seq_bytes[j] = 0;       // no header above this line (a header line itself holds 0 already)
}

__global__ void k_fa_emit(const char *__restrict__ t, const int64_t *__restrict__ start, int64_t nlines,
                          const int64_t *__restrict__ is_header, const int64_t *__restrict__ rec_before,
                          const int64_t *__restrict__ pos, char *__restrict__ bytes, int64_t *__restrict__ off,
                          int64_t *__restrict__ hdr_off, int32_t *__restrict__ id_len, int32_t *__restrict__ hdr_len)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nlines) return;
    int64_t a, e;
    fa_line(t, start, j, a, e);
    if (is_header[j]) {
        const int64_t r = rec_before[j];
        off[r] = pos[j];
        hdr_off[r] = a + 1;
        hdr_len[r] = (int32_t)(e - a - 1);
        int64_t q = a + 1;
        while (q < e && !fa_space((unsigned char)t[q])) q++;
        id_len[r] = (int32_t)(q - a - 1);
    } else if (rec_before[j] > 0) {
        char *dst = bytes + pos[j];
        for (int64_t p = a; p < e; p++) {
            const char c = t[p];
            if (!fa_space((unsigned char)c)) *dst++ = c;
        }
    }
}

extern "C" int pg_fasta_ingest(pg_ctx *ctx, const char *text_host, int64_t len, int64_t cap, int64_t *nrec_out,
                               int64_t *hdr_off, int32_t *id_len, int32_t *hdr_len, pg_reads **reads_out)
{
    if (!ctx || (!text_host && len > 0) || len < 0 || !nrec_out || cap < 0)
        return pg_fail(ctx, PG_EINVAL, "pg_fasta_ingest: bad arguments");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    *nrec_out = 0;
    if (reads_out) *reads_out = NULL;
    PG_TRY(pg_scratch(ctx, &ctx->s_candl, (size_t)len + 16));            // the text
    char *d_text = (char *)ctx->s_candl.p;
    if (len) PG_CUDA(ctx, cudaMemcpyAsync(d_text, text_host, (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    int64_t *d_start = NULL, nlines = 0;
    PG_TRY(pg_index_lines(ctx, d_text, len, &d_start, &nlines));
    int rc = PG_OK;
    int64_t nrec = 0, total = 0;
    int64_t *d_hdr = NULL, *d_cnt = NULL, *d_rec = NULL, *d_pos = NULL, *d_hoff = NULL;
    int32_t *d_idl = NULL, *d_hl = NULL;
    pg_reads *rd = NULL;
#define FA_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = pg_fail(ctx, PG_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); goto done; } } while (0)
#define FA_TRY(call) do { if ((rc = (call)) != PG_OK) goto done; } while (0)
    {
        const size_t nl1 = (size_t)nlines + 2;
        FA_TRY(pg_scratch(ctx, &ctx->s_champ, nl1 * 8 * 4));
        d_hdr = (int64_t *)ctx->s_champ.p;
        d_cnt = d_hdr + nl1;
        d_rec = d_cnt + nl1;
        d_pos = d_rec + nl1;
        const unsigned nb = (unsigned)((nlines + 255) / 256);
        if (nlines) {
            k_fa_lines<<<nb, 256, 0, ctx->stream>>>(d_text, d_start, nlines, d_hdr, d_cnt);
            ctx->launches++;
        }
        FA_TRY(pg_device_scan(ctx, d_hdr, nlines, d_rec));              // records opened above each line
        if (nlines) {
            k_fa_drop_preamble<<<nb, 256, 0, ctx->stream>>>(d_rec, nlines, d_cnt);
            ctx->launches++;
        }
        FA_TRY(pg_device_scan(ctx, d_cnt, nlines, d_pos));              // sequence bytes above each line
        FA_CUDA(pg_copy_sync(ctx, &nrec, d_rec + nlines, 8, cudaMemcpyDeviceToHost));
        FA_CUDA(pg_copy_sync(ctx, &total, d_pos + nlines, 8, cudaMemcpyDeviceToHost));
        *nrec_out = nrec;
        if (nrec > cap) { rc = pg_fail(ctx, PG_ERANGE, "pg_fasta_ingest: %lld records, room for %lld", (long long)nrec, (long long)cap); goto done; }
        FA_TRY(pg_scratch(ctx, &ctx->s_bytes, (size_t)total + 64));    // compacted sequence bytes
        FA_TRY(pg_scratch(ctx, &ctx->s_off, (size_t)(nrec + 1) * 8));
        FA_TRY(pg_scratch(ctx, &ctx->s_results, (size_t)(nrec + 1) * 16 + 64));
        d_hoff = (int64_t *)ctx->s_results.p;
        d_idl = (int32_t *)(d_hoff + nrec + 1);
        d_hl = d_idl + nrec + 1;
        int64_t *d_off = (int64_t *)ctx->s_off.p;
        if (nlines) {
            k_fa_emit<<<nb, 256, 0, ctx->stream>>>(d_text, d_start, nlines, d_hdr, d_rec, d_pos, (char *)ctx->s_bytes.p, d_off,
                                                   d_hoff, d_idl, d_hl);
            ctx->launches++;
        }
        FA_CUDA(cudaMemcpyAsync(d_off + nrec, &total, 8, cudaMemcpyHostToDevice, ctx->stream));
        FA_CUDA(cudaGetLastError());
        if (nrec) {
            if (hdr_off) FA_CUDA(cudaMemcpyAsync(hdr_off, d_hoff, (size_t)nrec * 8, cudaMemcpyDeviceToHost, ctx->stream));
            if (id_len) FA_CUDA(cudaMemcpyAsync(id_len, d_idl, (size_t)nrec * 4, cudaMemcpyDeviceToHost, ctx->stream));
            if (hdr_len) FA_CUDA(cudaMemcpyAsync(hdr_len, d_hl, (size_t)nrec * 4, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (reads_out) {
            pg_seqbatch sb;
            sb.bytes = (const char *)ctx->s_bytes.p;
            sb.off = d_off;
            sb.count = nrec;
            FA_TRY(pg_reads_pack_dev(ctx, &sb, total, &rd));
            *reads_out = rd;
        }
        FA_CUDA(cudaStreamSynchronize(ctx->stream));
    }
done:
#undef FA_CUDA
#undef FA_TRY
    pg_dev_free(ctx, d_start);
    if (rc != PG_OK && rd) { pg_reads_free(rd); if (reads_out) *reads_out = NULL; }
    return rc;
}

// parity hook: the compacted sequence bytes and offsets of the last pg_fasta_ingest() of this context
extern "C" int pg_fasta_last_bytes(pg_ctx *ctx, int64_t nrec, char *bytes_host, int64_t cap, int64_t *off_host)
{
    if (!ctx || nrec < 0 || !off_host) return pg_fail(ctx, PG_EINVAL, "pg_fasta_last_bytes: bad arguments");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    PG_CUDA(ctx, pg_copy_sync(ctx, off_host, ctx->s_off.p, (size_t)(nrec + 1) * 8, cudaMemcpyDeviceToHost));
    const int64_t total = off_host[nrec];
    if (total > cap) return pg_fail(ctx, PG_ERANGE, "pg_fasta_last_bytes: %lld bytes, room for %lld", (long long)total, (long long)cap);
    if (total && bytes_host) PG_CUDA(ctx, pg_copy_sync(ctx, bytes_host, ctx->s_bytes.p, (size_t)total, cudaMemcpyDeviceToHost));
    return PG_OK;
}
