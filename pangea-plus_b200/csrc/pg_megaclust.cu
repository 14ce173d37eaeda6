// pg_megaclust.cu -- Megaclust thresholding + OTU counting on the GPU.
//
// Replaces the body of Megaclust/megaclust2.pl :80-139 (README.md:176): every non-comment line of a
// BLAST-tabular / consensus file is split the script's way (`split /\t\t|\t\s|\s\t|\t/` :98), tested
// against the three thresholds with Perl's numeric reading of the fields (:126-130), and counted per
// subject -- once per distinct (subject, query) pair unless -c is given (:133-143).
//
//   k_mc_parse     thread per line: fields, thresholds, 64-bit hashes of subject and (subject, query)
//   k_mc_pairs     thread per passing line: insert-if-absent into an open-addressing table keyed by the
//                  (subject, query) strings (atomicCAS on the slot, byte compare on a hash match)
//   k_mc_subjects  thread per passing line: subject table, atomicAdd on the count (counted lines), atomicMin on the
//                  first line (the output lists OTUs in order of first appearance; the script's own
//                  order is Perl's hash order, i.e. unspecified)
//   k_mc_collect   occupied subject slots -> compact (first line, count) records
#include "pg_internal.cuh"
#include "pg_perlnum.h"
#include <algorithm>

int pg_index_lines(pg_ctx *ctx, const char *d_text, int64_t n, int64_t **d_start_out, int64_t *nlines_out);   // pg_trim.cu

struct McLine {
    int64_t  subj_off, query_off;     // byte offsets into the text
    int32_t  subj_len, query_len;
    uint64_t h_subj, h_pair;
    int32_t  state;                   // 0 comment, 1 beyond thresholds, 2 passes
    int32_t  counted;                 // set by k_mc_pairs: this line adds one to its subject
};

__device__ __forceinline__ bool mc_space(int c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

__device__ __forceinline__ uint64_t mc_hash(const char *s, int n, uint64_t h)
{
    for (int i = 0; i < n; i++) {                     // FNV-1a, then a finaliser below
        h ^= (unsigned char)s[i];
        h *= 0x100000001B3ULL;
    }
    return h;
}
__device__ __forceinline__ uint64_t mc_mix(uint64_t h)
{
    h ^= h >> 33; h *= 0xFF51AFD7ED558CCDULL; h ^= h >> 33; h *= 0xC4CEB9FE1A85EC53ULL; h ^= h >> 33;
    return h;
}

__global__ void k_mc_parse(const char *__restrict__ t, const int64_t *__restrict__ start, int64_t nlines,
                           double sim_thr, double eval_thr, double bit_thr, McLine *__restrict__ lines,
                           unsigned long long *__restrict__ stats /* [0] examined, [1] beyond */)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlines) return;
    const int64_t a = start[i];
    int n = (int)(start[i + 1] - a);
    const char *s = t + a;
    McLine L;
    L.subj_off = L.query_off = a;
    L.subj_len = L.query_len = 0;
    L.h_subj = L.h_pair = 0;
    L.counted = 0;
    if (n > 0 && s[0] == '#') {                       // next if (/^\#/)
        L.state = 0;
        lines[i] = L;
        return;
    }
    if (n > 0 && s[n - 1] == '\n') n--;               // chomp
    // split /\t\t|\t\s|\s\t|\t/: leftmost match, alternatives in order
    int fstart[13], flen[13], nf = 0, p = 0, f0 = 0;
    while (nf < 13) {
        int dl = 0;
        if (p < n) {
            const int c = (unsigned char)s[p];
            if (c == '\t') dl = (p + 1 < n && mc_space((unsigned char)s[p + 1])) ? 2 : 1;
            else if (mc_space(c) && p + 1 < n && s[p + 1] == '\t') dl = 2;
        }
        if (p >= n || dl) {
            fstart[nf] = f0;
            flen[nf] = p - f0;
            nf++;
            if (p >= n) break;
            p += dl;
            f0 = p;
        } else {
            p++;
        }
    }
    for (int k = nf; k < 13; k++) { fstart[k] = 0; flen[k] = 0; }     // undef -> "" / 0
    const double pid = pg_perl_number(s + fstart[2], flen[2]);
    const double ev = pg_perl_number(s + fstart[10], flen[10]);
    const double bs = pg_perl_number(s + fstart[11], flen[11]);
    const bool beyond = (pid < sim_thr) || (ev > eval_thr) || (bs < bit_thr);
    atomicAdd(&stats[0], 1ULL);
    if (beyond) atomicAdd(&stats[1], 1ULL);
    L.state = beyond ? 1 : 2;
    L.query_off = a + fstart[0];
    L.query_len = flen[0];
    L.subj_off = a + fstart[1];
    L.subj_len = flen[1];
    const uint64_t hs = mc_hash(s + fstart[1], flen[1], 0xCBF29CE484222325ULL);
    L.h_subj = mc_mix(hs);
    L.h_pair = mc_mix(mc_hash(s + fstart[0], flen[0], hs ^ 0x9E3779B97F4A7C15ULL));
    lines[i] = L;
}

__device__ __forceinline__ bool mc_same(const char *t, int64_t a, int la, int64_t b, int lb)
{
    if (la != lb) return false;
    for (int i = 0; i < la; i++)
        if (t[a + i] != t[b + i]) return false;
    return true;
}

// slot value: 0 = empty, else (line index + 1) of the line that claimed it
__global__ void k_mc_pairs(const char *__restrict__ t, McLine *__restrict__ lines, int64_t nlines,
                           unsigned long long *__restrict__ table, uint64_t mask, int count_every)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlines) return;
    McLine &L = lines[i];
    if (L.state != 2) return;
    if (count_every) { L.counted = 1; return; }
    uint64_t slot = L.h_pair & mask;
    for (;;) {
        unsigned long long cur = table[slot];
        if (cur == 0ULL) {
            cur = atomicCAS(&table[slot], 0ULL, (unsigned long long)(i + 1));
            if (cur == 0ULL) { L.counted = 1; return; }           // first line of this (subject, query) pair
        }
        const McLine &O = lines[cur - 1];
        if (O.h_pair == L.h_pair && mc_same(t, O.subj_off, O.subj_len, L.subj_off, L.subj_len) &&
            mc_same(t, O.query_off, O.query_len, L.query_off, L.query_len))
            return;                                               // pair already counted
        slot = (slot + 1) & mask;
    }
}

struct McSlot { unsigned long long owner; unsigned long long first; unsigned long long count; };

__global__ void k_mc_subjects(const char *__restrict__ t, const McLine *__restrict__ lines, int64_t nlines,
                              McSlot *__restrict__ table, uint64_t mask)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlines) return;
    const McLine &L = lines[i];
    if (L.state != 2) return;
    // every passing line votes for the subject's first line (which of two lines of the same (subject, query)
    // pair won the race in k_mc_pairs must not show in the output order); only counted lines add to the count
    uint64_t slot = L.h_subj & mask;
    for (;;) {
        unsigned long long cur = table[slot].owner;
        if (cur == 0ULL) {
            cur = atomicCAS(&table[slot].owner, 0ULL, (unsigned long long)(i + 1));
            if (cur == 0ULL) cur = (unsigned long long)(i + 1);
        }
        const McLine &O = lines[cur - 1];
        if (O.h_subj == L.h_subj && mc_same(t, O.subj_off, O.subj_len, L.subj_off, L.subj_len)) {
            if (L.counted) atomicAdd(&table[slot].count, 1ULL);
            atomicMin(&table[slot].first, (unsigned long long)i);
            return;
        }
        slot = (slot + 1) & mask;
    }
}

__global__ void k_mc_init_slots(McSlot *__restrict__ table, uint64_t size)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    table[i].owner = 0ULL;
    table[i].first = ~0ULL;
    table[i].count = 0ULL;
}

__global__ void k_mc_collect(const McSlot *__restrict__ table, uint64_t size, unsigned long long *__restrict__ nout,
                             unsigned long long *__restrict__ out /* pairs (first line, count) */)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size || table[i].owner == 0ULL) return;
    const unsigned long long k = atomicAdd(nout, 1ULL);
    out[2 * k] = table[i].first;
    out[2 * k + 1] = table[i].count;
}

__global__ void k_mc_fields(const McLine *__restrict__ lines, const unsigned long long *__restrict__ recs, int64_t n,
                            int64_t *__restrict__ off, int32_t *__restrict__ len)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const McLine &L = lines[recs[2 * i]];
    off[i] = L.subj_off;
    len[i] = L.subj_len;
}

extern "C" int pg_megaclust(pg_ctx *ctx, const char *text_host, int64_t len, const pg_megaclust_opts *opts,
                            int64_t cap, int64_t *n_otus, int64_t *otu_off, int32_t *otu_len, int64_t *otu_count,
                            int64_t *lines_examined, int64_t *lines_beyond)
{
    if (!ctx || (!text_host && len > 0) || len < 0 || !n_otus || cap < 0)
        return pg_fail(ctx, PG_EINVAL, "pg_megaclust: bad arguments");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    const double sim = opts ? opts->sim_threshold : 95.0;
    const double ev = opts ? opts->eval_threshold : 1e-20;
    const double bs = opts ? opts->bitscore_threshold : 200.0;
    const int every = opts ? opts->count_every_hit : 0;
    *n_otus = 0;
    if (lines_examined) *lines_examined = 0;
    if (lines_beyond) *lines_beyond = 0;
    if (len == 0) return PG_OK;

    char *d_text = NULL;
    int64_t *d_start = NULL, nlines = 0;
    McLine *d_lines = NULL;
    unsigned long long *d_pairs = NULL, *d_stats = NULL, *d_recs = NULL;
    McSlot *d_subj = NULL;
    int64_t *d_off = NULL;
    int32_t *d_len = NULL;
    int rc = PG_OK;
    std::vector<unsigned long long> recs;
    std::vector<size_t> order;
    unsigned long long st[3] = {0, 0, 0};
    uint64_t size = 1;
#define MC_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = pg_fail(ctx, PG_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); goto done; } } while (0)
    // grow-only context scratch (shared with the other stages; one call at a time per context)
    if ((rc = pg_scratch(ctx, &ctx->s_bytes, (size_t)len + 16)) != PG_OK) goto done;
    d_text = (char *)ctx->s_bytes.p;
    MC_CUDA(cudaMemcpyAsync(d_text, text_host, (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = pg_index_lines(ctx, d_text, len, &d_start, &nlines)) != PG_OK) goto done;
    if (nlines == 0) goto done;
    while (size < (uint64_t)nlines * 2) size <<= 1;                    // load factor <= 0.5
    if ((rc = pg_scratch(ctx, &ctx->s_cand, (size_t)nlines * sizeof(McLine))) != PG_OK ||
        (rc = pg_scratch(ctx, &ctx->s_candl, (size_t)size * 8)) != PG_OK ||
        (rc = pg_scratch(ctx, &ctx->s_champ, (size_t)size * sizeof(McSlot))) != PG_OK ||
        (rc = pg_scratch(ctx, &ctx->s_results, (size_t)nlines * 16 + 96)) != PG_OK)
        goto done;
    d_lines = (McLine *)ctx->s_cand.p;
    d_pairs = (unsigned long long *)ctx->s_candl.p;
    d_subj = (McSlot *)ctx->s_champ.p;
    d_recs = (unsigned long long *)ctx->s_results.p + 4;
    d_stats = (unsigned long long *)ctx->s_results.p;
    MC_CUDA(cudaMemsetAsync(d_pairs, 0, (size_t)size * 8, ctx->stream));
    MC_CUDA(cudaMemsetAsync(d_stats, 0, 32, ctx->stream));
    {
        const unsigned nb = (unsigned)((nlines + 255) / 256);
        k_mc_init_slots<<<(unsigned)((size + 255) / 256), 256, 0, ctx->stream>>>(d_subj, size);
        k_mc_parse<<<nb, 256, 0, ctx->stream>>>(d_text, d_start, nlines, sim, ev, bs, d_lines, d_stats);
        k_mc_pairs<<<nb, 256, 0, ctx->stream>>>(d_text, d_lines, nlines, d_pairs, size - 1, every);
        k_mc_subjects<<<nb, 256, 0, ctx->stream>>>(d_text, d_lines, nlines, d_subj, size - 1);
        k_mc_collect<<<(unsigned)((size + 255) / 256), 256, 0, ctx->stream>>>(d_subj, size, d_stats + 2, d_recs);
        ctx->launches += 5;
        MC_CUDA(cudaGetLastError());
    }
    MC_CUDA(cudaMemcpyAsync(st, d_stats, 24, cudaMemcpyDeviceToHost, ctx->stream));
    MC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (lines_examined) *lines_examined = (int64_t)st[0];
    if (lines_beyond) *lines_beyond = (int64_t)st[1];
    *n_otus = (int64_t)st[2];
    if ((int64_t)st[2] > cap) { rc = pg_fail(ctx, PG_ERANGE, "pg_megaclust: %lld OTUs, room for %lld", (long long)st[2], (long long)cap); goto done; }
    if (st[2] == 0) goto done;
    // order of first appearance (the table's own order depends on the hash)
    recs.resize((size_t)st[2] * 2);
    MC_CUDA(pg_copy_sync(ctx, recs.data(), d_recs, recs.size() * 8, cudaMemcpyDeviceToHost));
    order.resize((size_t)st[2]);
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return recs[2 * a] < recs[2 * b]; });
    {
        std::vector<unsigned long long> sorted(recs.size());
        for (size_t i = 0; i < order.size(); i++) { sorted[2 * i] = recs[2 * order[i]]; sorted[2 * i + 1] = recs[2 * order[i] + 1]; }
        MC_CUDA(pg_copy_sync(ctx, d_recs, sorted.data(), sorted.size() * 8, cudaMemcpyHostToDevice));
        MC_CUDA(cudaMalloc(&d_off, (size_t)st[2] * 8));
        MC_CUDA(cudaMalloc(&d_len, (size_t)st[2] * 4));
        k_mc_fields<<<(unsigned)((st[2] + 255) / 256), 256, 0, ctx->stream>>>(d_lines, d_recs, (int64_t)st[2], d_off, d_len);
        ctx->launches++;
        MC_CUDA(cudaGetLastError());
        if (otu_off) MC_CUDA(cudaMemcpyAsync(otu_off, d_off, (size_t)st[2] * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (otu_len) MC_CUDA(cudaMemcpyAsync(otu_len, d_len, (size_t)st[2] * 4, cudaMemcpyDeviceToHost, ctx->stream));
        MC_CUDA(cudaStreamSynchronize(ctx->stream));
        if (otu_count)
            for (size_t i = 0; i < order.size(); i++) otu_count[i] = (int64_t)sorted[2 * i + 1];
    }
done:
#undef MC_CUDA
    pg_dev_free(ctx, d_start); cudaFree(d_off); cudaFree(d_len);
    return rc;
}
