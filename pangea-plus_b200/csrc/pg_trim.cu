// pg_trim.cu -- the Trim join on the GPU: quality trim of Illumina
// QSEQ pairs / FASTQ records and the A + N x gap + B join that defines the reads entering Stage A.
//
// Replaces Trim/trim2.4.pl (trim2.3.pl is identical on these paths): parse_qseq :169-242,
// trim_qseq :244-298, parse_fastq :467-521, trim_fastq :527-578.  The output is the text of
// <prefix>_runblast.fasta byte for byte -- including the script's accidents (listed in
// oracle/trim_ref.c) -- and, on request, the joined sequences already packed into the device
// read store, so trimmed reads can go straight into pg_classify_packed without a text round trip.
//
//   k_count_newlines / k_fill_line_starts   line index of a text buffer (two passes around a scan)
//   k_trim_qseq / k_trim_fastq              thread per output record; pass 1 lengths, pass 2 bytes
#include "pg_internal.cuh"

#define PG_TRIM_QUALITY_CUTOFF 20     // -qc cannot be parsed by the script's getopts string: always 20
#define PG_TRIM_LENGTH_CUTOFF  70     // -lc likewise: always 70
#define PG_SEG 512

__global__ void k_count_newlines(const char *__restrict__ t, int64_t n, int64_t *__restrict__ seg_count)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t a = s * PG_SEG;
    if (a >= n) return;
    const int64_t b = a + PG_SEG < n ? a + PG_SEG : n;
    int64_t c = 0;
    for (int64_t p = a; p < b; p++) c += t[p] == '\n';
    seg_count[s] = c;
}

// line_start[j] = offset of line j; line_start[nlines] = n (+1 past a final newline-less line)
__global__ void k_fill_line_starts(const char *__restrict__ t, int64_t n, const int64_t *__restrict__ seg_off,
                                   int64_t *__restrict__ line_start)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t a = s * PG_SEG;
    if (a >= n) return;
    const int64_t b = a + PG_SEG < n ? a + PG_SEG : n;
    int64_t j = seg_off[s] + 1;
    for (int64_t p = a; p < b; p++)
        if (t[p] == '\n') line_start[j++] = p + 1;
}

struct TextLines {
    const char *t;
    const int64_t *start;     // nlines + 1 entries; line j = [start[j], start[j+1]) incl. its newline if any
    int64_t nlines;
    int64_t nbytes;
};

__device__ __forceinline__ void pg_line(const TextLines &L, int64_t j, const char *&s, int &len, bool &has_nl)
{
    if (j >= L.nlines) { s = L.t; len = 0; has_nl = false; return; }
    const int64_t a = L.start[j], b = L.start[j + 1];
    s = L.t + a;
    len = (int)(b - a);
    has_nl = len > 0 && s[len - 1] == '\n';
    if (has_nl) len--;
}

// running-sum maximum: index of the last new maximum of sum(q - base - cutoff), sum restarting at
// 0 whenever it turns negative (trim_qseq :262-277, trim_fastq :541-559)
__device__ int pg_best_end(const char *q, int n, int base, int extra)
{
    int sum = 0, mx = 0, end = 0;
    const int total = n + (extra >= 0 ? 1 : 0);
    for (int a = 0; a < total; a++) {
        const int c = a < n ? (int)(unsigned char)q[a] : extra;
        sum += c - base - PG_TRIM_QUALITY_CUTOFF;
        if (sum > mx) { mx = sum; end = a; }
        if (sum < 0) sum = 0;
    }
    return end;
}

// tab-separated field k of a QSEQ line
__device__ void pg_field(const char *s, int n, int k, const char *&f, int &fl)
{
    int idx = 0, start = 0;
    for (int p = 0; p <= n; p++)
        if (p == n || s[p] == '\t') {
            if (idx == k) { f = s + start; fl = p - start; return; }
            idx++;
            start = p + 1;
        }
    f = s;
    fl = 0;
}

__device__ int pg_trim_qseq_one(const char *seq, int sl, const char *qual, int ql, int truncate, const char *&out)
{
    int s0 = truncate <= sl ? truncate : sl;
    const int q0 = truncate <= ql ? truncate : ql;
    int s_len = sl - s0;
    const int q_len = ql - q0;
    int cut = truncate - 1;
    if (cut < 0) cut = 0;
    if (cut > s_len) cut = s_len;
    s0 += cut;
    s_len -= cut;
    const int end = pg_best_end(qual + q0, q_len, 64, -1);
    const int keep = end <= s_len ? end : s_len;
    out = seq + s0;
    return keep < PG_TRIM_LENGTH_CUTOFF ? -1 : keep;
}

// pass 1 (text_out == NULL): text_len[i], seq_len[i];  pass 2: bytes at text_off[i] / seq_off[i]
__global__ void k_trim_qseq(TextLines A, TextLines B, int64_t nrec, int gap, int truncate,
                            int64_t *__restrict__ text_len, int64_t *__restrict__ seq_len,
                            const int64_t *__restrict__ text_off, const int64_t *__restrict__ seq_off,
                            char *__restrict__ text_out, char *__restrict__ seq_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrec) return;
    const char *la, *lb;
    int na, nb;
    bool nl;
    pg_line(A, i, la, na, nl);
    pg_line(B, i, lb, nb, nl);
    const char *f7, *s1, *q1, *s2, *q2;
    int l7, k1, lq1, k2, lq2;
    pg_field(la, na, 7, f7, l7);
    pg_field(la, na, 8, s1, k1);
    pg_field(la, na, 9, q1, lq1);
    pg_field(lb, nb, 8, s2, k2);
    pg_field(lb, nb, 9, q2, lq2);
    const bool trimmed = (l7 == 1 && f7[0] == '1');
    if (trimmed) {
        k1 = pg_trim_qseq_one(s1, k1, q1, lq1, truncate, s1);
        k2 = pg_trim_qseq_one(s2, k2, q2, lq2, truncate, s2);
    } else {
        if (k1 == 1 && s1[0] == '0') k1 = -1;
        if (k2 == 1 && s2[0] == '0') k2 = -1;
    }
    const bool keep = k1 >= 0 && k2 >= 0;
    // header = join(':', fields 0..7) = the line up to the 8th TAB with TABs as ':' (missing fields are empty)
    int hlen = 0, tabs = 0;
    for (int p = 0; p < na && tabs < 8; p++) { if (la[p] == '\t') tabs++; if (tabs < 8) hlen++; }
    const int missing = tabs < 8 ? 7 - tabs : 0;            // short line: the absent fields still get their ':'
    if (!text_out) {
        text_len[i] = keep ? (int64_t)1 + hlen + missing + 4 + k1 + gap + k2 + 1 : 0;
        seq_len[i] = keep ? (int64_t)k1 + gap + k2 : 0;
        return;
    }
    if (!keep) return;
    char *o = text_out + text_off[i];
    *o++ = '>';
    for (int p = 0; p < hlen; p++) *o++ = la[p] == '\t' ? ':' : la[p];
    for (int p = 0; p < missing; p++) *o++ = ':';
    *o++ = ':'; *o++ = 'A'; *o++ = 'B'; *o++ = '\n';
    char *q = seq_out ? seq_out + seq_off[i] : NULL;
    for (int p = 0; p < k1; p++) { const char c = (trimmed && s1[p] == '.') ? 'N' : s1[p]; *o++ = c; if (q) *q++ = c; }
    for (int p = 0; p < gap; p++) { *o++ = 'N'; if (q) *q++ = 'N'; }
    for (int p = 0; p < k2; p++) { const char c = (trimmed && s2[p] == '.') ? 'N' : s2[p]; *o++ = c; if (q) *q++ = c; }
    *o++ = '\n';
}

// one FASTQ mate: kept prefix length of the sequence line, -1 for "0"; extra_nl = the kept prefix
// reaches into the sequence line's own newline (only when the quality line is longer than the bases)
__device__ int pg_trim_fastq_one(const TextLines &A, int64_t first_line, const char *&seq, bool &extra_nl)
{
    const char *s, *q, *d;
    int sl, ql, dl;
    bool snl, qnl, dnl;
    pg_line(A, first_line + 1, s, sl, snl);
    pg_line(A, first_line + 3, q, ql, qnl);
    (void)d; (void)dl; (void)dnl;
    const int end = pg_best_end(q, ql, 33, qnl ? '\n' : -1);
    const int avail = sl + 1;
    const int keep = end <= avail ? end : avail;
    seq = s;
    extra_nl = keep > sl;
    if (keep < PG_TRIM_LENGTH_CUTOFF) return -1;
    return keep <= sl ? keep : sl;
}

__global__ void k_trim_fastq(TextLines A, int64_t nrec, int paired, int gap, int64_t *__restrict__ text_len,
                             int64_t *__restrict__ seq_len, const int64_t *__restrict__ text_off,
                             const int64_t *__restrict__ seq_off, char *__restrict__ text_out,
                             char *__restrict__ seq_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrec) return;
    const int64_t l0 = i * (paired ? 8 : 4);
    const char *h, *s1, *s2 = NULL;
    int hl;
    bool hnl, x1, x2 = false;
    pg_line(A, l0, h, hl, hnl);
    int hkeep = 0;
    for (int p = 0; p < hl; p++) hkeep += h[p] != '@';          // s/@//g
    const int k1 = pg_trim_fastq_one(A, l0, s1, x1);
    int k2 = -1;
    const bool have2 = paired && (l0 + 4 < A.nlines);
    if (have2) k2 = pg_trim_fastq_one(A, l0 + 4, s2, x2);
    const int n1 = k1 < 0 ? 1 : k1;                             // "0" for a mate that is too short
    // mate 2 keeps the TAB of "SEQ\t" (only mate 1 goes through s/\s//g); "0" has none
    const int n2 = !paired ? 0 : (k2 < 0 ? 1 : k2 + (x2 ? 1 : 0) + 1);
    if (!text_out) {
        text_len[i] = (int64_t)1 + hkeep + 4 + n1 + (paired ? gap + n2 : 0) + 1;
        seq_len[i] = (int64_t)n1 + (paired ? gap + (k2 < 0 ? 1 : k2) : 0);
        return;
    }
    char *o = text_out + text_off[i];
    char *q = seq_out ? seq_out + seq_off[i] : NULL;
    *o++ = '>';
    for (int p = 0; p < hl; p++) if (h[p] != '@') *o++ = h[p];
    *o++ = ':'; *o++ = 'A'; *o++ = 'B'; *o++ = '\n';
    if (k1 < 0) { *o++ = '0'; if (q) *q++ = '0'; }
    else for (int p = 0; p < k1; p++) { *o++ = s1[p]; if (q) *q++ = s1[p]; }
    if (paired) {
        for (int p = 0; p < gap; p++) { *o++ = 'N'; if (q) *q++ = 'N'; }
        if (k2 < 0) { *o++ = '0'; if (q) *q++ = '0'; }
        else {
            for (int p = 0; p < k2; p++) { *o++ = s2[p]; if (q) *q++ = s2[p]; }
            if (x2) *o++ = '\n';
            *o++ = '\t';
        }
    }
    *o++ = '\n';
}

// ------------------------------------------------------------------ host

int pg_index_lines(pg_ctx *ctx, const char *d_text, int64_t n, int64_t **d_start_out, int64_t *nlines_out)
{
    const int64_t nseg = (n + PG_SEG - 1) / PG_SEG;
    int64_t *d_cnt = NULL, *d_off = NULL, *d_start = NULL;
    PG_CUDA(ctx, pg_dev_alloc(ctx, (void **)&d_cnt, (size_t)(nseg + 1) * 8));
    PG_CUDA(ctx, pg_dev_alloc(ctx, (void **)&d_off, (size_t)(nseg + 2) * 8));
    if (nseg) {
        k_count_newlines<<<(unsigned)((nseg + 255) / 256), 256, 0, ctx->stream>>>(d_text, n, d_cnt);
        PG_LAUNCHED(ctx);
    }
    PG_TRY(pg_device_scan(ctx, d_cnt, nseg, d_off));
    int64_t nnl = 0;
    PG_CUDA(ctx, pg_copy_sync(ctx, &nnl, d_off + nseg, 8, cudaMemcpyDeviceToHost));
    char last = '\n';
    if (n) PG_CUDA(ctx, pg_copy_sync(ctx, &last, d_text + n - 1, 1, cudaMemcpyDeviceToHost));
    const int64_t nlines = nnl + (n > 0 && last != '\n' ? 1 : 0);
    PG_CUDA(ctx, pg_dev_alloc(ctx, (void **)&d_start, (size_t)(nlines + 2) * 8));
    const int64_t zero = 0;
    PG_CUDA(ctx, pg_copy_sync(ctx, d_start, &zero, 8, cudaMemcpyHostToDevice));
    if (nseg) {
        k_fill_line_starts<<<(unsigned)((nseg + 255) / 256), 256, 0, ctx->stream>>>(d_text, n, d_off, d_start);
        PG_LAUNCHED(ctx);
    }
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PG_CUDA(ctx, pg_copy_sync(ctx, d_start + nlines, &n, 8, cudaMemcpyHostToDevice));     // end of the last line
    pg_dev_free(ctx, d_cnt);
    pg_dev_free(ctx, d_off);
    *d_start_out = d_start;                                // the caller releases it with pg_dev_free()
    *nlines_out = nlines;
    return PG_OK;
}

extern "C" int pg_trim_join(pg_ctx *ctx, const char *a_host, int64_t a_len, const char *b_host, int64_t b_len,
                            int paired, const pg_trim_opts *opts, char *out_host, int64_t out_cap, int64_t *out_len,
                            pg_reads **reads_out)
{
    if (!ctx || !a_host || a_len < 0 || b_len < 0 || !out_len || (out_cap > 0 && !out_host))
        return pg_fail(ctx, PG_EINVAL, "pg_trim_join: bad arguments");
    const int gap = opts ? opts->gap : 189, truncate = opts ? opts->truncate : 11;
    if (gap < 0 || gap > 100000 || truncate < 0) return pg_fail(ctx, PG_EINVAL, "pg_trim_join: bad gap / truncate");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    *out_len = 0;
    if (reads_out) *reads_out = NULL;
    if (a_len == 0) return PG_OK;
    const bool fastq = a_host[0] == '@';
    if (!fastq && (!b_host || a_host[0] == '>'))
        return pg_fail(ctx, PG_EINVAL, "pg_trim_join: QSEQ input needs both files (FASTA + quality-file trimming is not built)");
    char *d_a = NULL, *d_b = NULL;
    int64_t *d_sa = NULL, *d_sb = NULL, na = 0, nb = 0;
    PG_CUDA(ctx, cudaMalloc(&d_a, (size_t)a_len + 16));
    PG_CUDA(ctx, cudaMemcpyAsync(d_a, a_host, (size_t)a_len, cudaMemcpyHostToDevice, ctx->stream));
    PG_TRY(pg_index_lines(ctx, d_a, a_len, &d_sa, &na));
    if (!fastq) {
        PG_CUDA(ctx, cudaMalloc(&d_b, (size_t)b_len + 16));
        PG_CUDA(ctx, cudaMemcpyAsync(d_b, b_host, (size_t)b_len, cudaMemcpyHostToDevice, ctx->stream));
        PG_TRY(pg_index_lines(ctx, d_b, b_len, &d_sb, &nb));
    }
    TextLines A = {d_a, d_sa, na, a_len}, B = {d_b, d_sb, nb, b_len};
    const int64_t per = fastq ? (paired ? 8 : 4) : 1;
    const int64_t nrec = fastq ? (na + per - 1) / per : na;
    int64_t *d_len = NULL, *d_off = NULL;
    PG_CUDA(ctx, cudaMalloc(&d_len, (size_t)(nrec + 1) * 16));
    PG_CUDA(ctx, cudaMalloc(&d_off, (size_t)(nrec + 2) * 16));
    int64_t *d_tlen = d_len, *d_slen = d_len + (nrec + 1), *d_toff = d_off, *d_soff = d_off + (nrec + 2);
    const unsigned blocks = (unsigned)((nrec + 127) / 128);
    if (nrec) {
        if (fastq) k_trim_fastq<<<blocks, 128, 0, ctx->stream>>>(A, nrec, paired ? 1 : 0, gap, d_tlen, d_slen, NULL, NULL, NULL, NULL);
        else k_trim_qseq<<<blocks, 128, 0, ctx->stream>>>(A, B, nrec, gap, truncate, d_tlen, d_slen, NULL, NULL, NULL, NULL);
        PG_LAUNCHED(ctx);
    }
    PG_TRY(pg_device_scan(ctx, d_tlen, nrec, d_toff));
    PG_TRY(pg_device_scan(ctx, d_slen, nrec, d_soff));
    int64_t ttot = 0, stot = 0;
    PG_CUDA(ctx, pg_copy_sync(ctx, &ttot, d_toff + nrec, 8, cudaMemcpyDeviceToHost));
    PG_CUDA(ctx, pg_copy_sync(ctx, &stot, d_soff + nrec, 8, cudaMemcpyDeviceToHost));
    *out_len = ttot;
    int rc = PG_OK;
    if (ttot > out_cap) rc = pg_fail(ctx, PG_ERANGE, "pg_trim_join: output needs %lld bytes", (long long)ttot);
    char *d_text = NULL, *d_seq = NULL;
    if (rc == PG_OK && nrec) {
        PG_CUDA(ctx, cudaMalloc(&d_text, (size_t)ttot + 16));
        if (reads_out) PG_CUDA(ctx, cudaMalloc(&d_seq, (size_t)stot + 16));
        if (fastq) k_trim_fastq<<<blocks, 128, 0, ctx->stream>>>(A, nrec, paired ? 1 : 0, gap, NULL, NULL, d_toff, d_soff, d_text, d_seq);
        else k_trim_qseq<<<blocks, 128, 0, ctx->stream>>>(A, B, nrec, gap, truncate, NULL, NULL, d_toff, d_soff, d_text, d_seq);
        PG_LAUNCHED(ctx);
        PG_CUDA(ctx, cudaMemcpyAsync(out_host, d_text, (size_t)ttot, cudaMemcpyDeviceToHost, ctx->stream));
        PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (reads_out) {
            // joined sequences -> packed device read store (records dropped by the trim have length 0)
            pg_seqbatch sb = {d_seq, d_soff, nrec};
            rc = pg_reads_pack_dev(ctx, &sb, stot, reads_out);
        }
    }
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_a); cudaFree(d_b); pg_dev_free(ctx, d_sa); pg_dev_free(ctx, d_sb); cudaFree(d_len); cudaFree(d_off);
    cudaFree(d_text); cudaFree(d_seq);
    return rc;
}
