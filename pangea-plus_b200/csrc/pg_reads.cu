// pg_reads.cu -- K3: 2-bit packing, 8-mer extraction and orientation
// (SURVEY.md 8(a) rows A1, A2, A3).
//
// Replaces upstream GoodWordIterator (A1), the ShortSequenceException gate (A2)
// and TrainingInfo.isSeqReversed + ClassifierSequence.getReversedSeq (A3) of RDP
// Classifier 2.5 (the jar invoked at README.md:119).
//
// Device read store: three bit planes per 32-base chunk, {lo, hi, valid}:
//   code = A0 T/U1 G2 C3,  lo = code&1, hi = code>>1, valid = base is ACGTU.
// A warp packs 32 bases with three ballots (fully coalesced byte loads), and a
// lane recovers the 8-mer ending at its base with two funnel shifts.
#include "pg_internal.cuh"

__device__ __forceinline__ int pg_base_code2(unsigned char c)
{
    c &= 0xDF;
    int code = -1;
    if (c == 'A') code = 0;
    else if (c == 'T' || c == 'U') code = 1;
    else if (c == 'G') code = 2;
    else if (c == 'C') code = 3;
    return code;
}

// K3a: one warp per read.
__global__ void __launch_bounds__(256)
k_pack(const char *__restrict__ bytes, const int64_t *__restrict__ off, int64_t nreads,
       uint32_t *__restrict__ planes)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= nreads) return;
    const int64_t o = off[i], len = off[i + 1] - o;
    const char *seq = bytes + o;
    uint32_t *dst = planes + 3 * pg_chunk_start(o, i);
    const int64_t nchunks = (len + 31) >> 5;
    for (int64_t c = 0; c < nchunks; c++) {
        int64_t p = (c << 5) + lane;
        int code = p < len ? pg_base_code2((unsigned char)seq[p]) : -1;
        uint32_t va = __ballot_sync(0xffffffffu, code >= 0);
        uint32_t lo = __ballot_sync(0xffffffffu, code >= 0 && (code & 1));
        uint32_t hi = __ballot_sync(0xffffffffu, code >= 0 && (code & 2));
        if (lane == 0) {
            dst[3 * c + 0] = lo;
            dst[3 * c + 1] = hi;
            dst[3 * c + 2] = va;
        }
    }
}

// reverse complement of one 8-mer id: reverse the 2-bit groups, complement = XOR 1.
__device__ __forceinline__ uint32_t pg_revcomp_word(uint32_t w)
{
    uint32_t r = __brev(w) >> 16;                             // reverses all 16 bits
    r = ((r & 0xAAAAu) >> 1) | ((r & 0x5555u) << 1);          // restore bit order inside each base
    return r ^ 0x5555u;
}

// spread the low 8 bits of x to the even bit positions of a 16-bit value
__device__ __forceinline__ uint32_t pg_spread8(uint32_t x)
{
    x = (x | (x << 4)) & 0x0F0Fu;
    x = (x | (x << 2)) & 0x3333u;
    x = (x | (x << 1)) & 0x5555u;
    return x;
}

// K3b: one warp per read: word list (A1), orientation (A3), in-place reverse
// complement of the list when the reverse strand has the larger prior sum.
template <bool ASCII>
__global__ void __launch_bounds__(256)
k_extract(const uint32_t *__restrict__ planes, const char *__restrict__ bytes, const int64_t *__restrict__ off, int64_t nreads,
          const float *__restrict__ priorDiff, uint16_t *__restrict__ words,
          int32_t *__restrict__ nwords, uint8_t *__restrict__ flags /* [2*nreads]: reversed, status */)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= nreads) return;
    const int64_t o = off[i], len = off[i + 1] - o;
    if (len < PG_MIN_SEQ_LEN) {                 // A2: ShortSequenceException
        if (lane == 0) { nwords[i] = 0; flags[2 * i] = 0; flags[2 * i + 1] = 1; }
        return;
    }
    const uint32_t *src = ASCII ? NULL : planes + 3 * pg_chunk_start(o, i);
    uint16_t *w = words + o;
    const int64_t nchunks = (len + 31) >> 5;

    // ---- A1: good words in sequence order
    int n = 0;
    uint32_t plo = 0, phi = 0, pva = 0;
    for (int64_t c = 0; c < nchunks; c++) {
        uint32_t lo, hi, va;
        if (ASCII) {                                // the three bit planes of k_pack, straight from the text
            const int64_t p = (c << 5) + lane;
            const int code = p < len ? pg_base_code2((unsigned char)bytes[o + p]) : -1;
            va = __ballot_sync(0xffffffffu, code >= 0);
            lo = __ballot_sync(0xffffffffu, code >= 0 && (code & 1));
            hi = __ballot_sync(0xffffffffu, code >= 0 && (code & 2));
        } else {
            lo = src[3 * c + 0]; hi = src[3 * c + 1]; va = src[3 * c + 2];
        }
        // bits [lane-7 .. lane] of the 64-bit stream (prev:cur), oldest base lowest
        // = bits [lane+25 .. lane+32] of (cur:prev); __funnelshift_r wraps its shift
        // mod 32, so the part that lies wholly inside `cur` is shifted directly.
        const uint32_t sh = (uint32_t)(lane + 32 - 7);
        uint32_t v8, l8, h8;
        if (sh >= 32) {
            v8 = (va >> (sh - 32)) & 0xFFu;
            l8 = (lo >> (sh - 32)) & 0xFFu;
            h8 = (hi >> (sh - 32)) & 0xFFu;
        } else {
            v8 = __funnelshift_r(pva, va, sh) & 0xFFu;
            l8 = __funnelshift_r(plo, lo, sh) & 0xFFu;
            h8 = __funnelshift_r(phi, hi, sh) & 0xFFu;
        }
        bool ok = (v8 == 0xFFu);
        // oldest base is the most significant pair of the word id
        uint32_t word = pg_spread8(__brev(l8) >> 24) | (pg_spread8(__brev(h8) >> 24) << 1);
        uint32_t bal = __ballot_sync(0xffffffffu, ok);
        if (ok) w[n + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)word;
        n += __popc(bal);
        plo = lo; phi = hi; pva = va;
    }
    __syncwarp();

    // ---- A3 (upstream TrainingInfo.isSeqReversed): prior = sum over the words, in word order, of
    // priorDiff[w] = logPrior[w] - logPrior[rc(w)] -- ONE fp32 accumulator; the read is reverse-complemented iff
    // prior < 0.  priorDiff is a per-model table (k_prior_diff).  Lane l gathers the 16 consecutive terms
    // 16l .. 16l+15 of a 512-word pass; the running sum hops from lane to lane, each adding its own 16 terms in
    // order: the adds stay one chain in word order for one shuffle per 16 words.
    float acc = 0.0f;
    for (int base = 0; base < n; base += 512) {
        float v[16];
        const int j0 = base + lane * 16;
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const int j = j0 + u;
            v[u] = 0.0f;
            if (j < n) v[u] = __ldg(priorDiff + w[j]);
        }
        const int cnt = n - j0;
        const int hops = (n - base + 15) / 16 < 32 ? (n - base + 15) / 16 : 32;
        for (int h = 0; h < hops; h++) {
            const float in = __shfl_sync(0xffffffffu, acc, h == 0 ? 31 : h - 1);
            if (lane == h) {
                float x = (h == 0 && base == 0) ? 0.0f : in;
#pragma unroll
                for (int u = 0; u < 16; u++)
                    if (u < cnt) x = __fadd_rn(x, v[u]);
                acc = x;
            }
        }
        acc = __shfl_sync(0xffffffffu, acc, hops - 1);              // every lane holds the pass result (lane 31 for the next pass)
    }
    const bool reversed = n > 0 && acc < 0.0f;
    if (reversed) {
        // word list of the reverse-complemented read = reversed list of rc words
        for (int j = lane; j < n / 2; j += 32) {
            uint32_t a = w[j], b = w[n - 1 - j];
            w[j] = (uint16_t)pg_revcomp_word(b);
            w[n - 1 - j] = (uint16_t)pg_revcomp_word(a);
        }
        if ((n & 1) && lane == 0) w[n / 2] = (uint16_t)pg_revcomp_word(w[n / 2]);
    }
    if (lane == 0) {
        nwords[i] = n;
        flags[2 * i] = reversed ? 1 : 0;
        flags[2 * i + 1] = 0;
    }
}

// A3: priorDiff[w] = fp32(logPrior[w] - logPrior[rc(w)]) (upstream wordPairPriorDiffArr), once per model
__global__ void k_prior_diff(const float *__restrict__ logPrior, float *__restrict__ priorDiff)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < PG_NWORDS) priorDiff[w] = __fsub_rn(logPrior[w], logPrior[pg_revcomp_word(w)]);
}

int pg_prior_diff_launch(pg_ctx *ctx, pg_model *md)
{
    k_prior_diff<<<PG_NWORDS / 256, 256, 0, ctx->stream>>>(md->d_logPrior, md->d_pdiff);
    PG_LAUNCHED(ctx);
    return PG_OK;
}

// ------------------------------------------------------------------ host side

static int reads_alloc(pg_ctx *ctx, int64_t count, int64_t total_bytes, pg_reads **out)
{
    pg_reads *r = new pg_reads();
    memset(r, 0, sizeof *r);
    r->ctx = ctx;
    r->count = count;
    r->total_bytes = total_bytes;
    r->nchunks_cap = (total_bytes >> 5) + count + 2;
    cudaError_t e;
    if ((e = pg_dev_alloc(ctx, (void **)&r->d_off, (size_t)(count + 1) * 8)) != cudaSuccess ||
        (e = pg_dev_alloc(ctx, (void **)&r->d_planes, (size_t)r->nchunks_cap * 12)) != cudaSuccess) {
        (void)cudaGetLastError();
        pg_reads_free(r);
        return pg_fail(ctx, PG_ENOMEM, "pg_reads: device allocation failed: %s", cudaGetErrorString(e));
    }
    *out = r;
    return PG_OK;
}

extern "C" void pg_reads_free(pg_reads *r)
{
    if (!r) return;
    cudaSetDevice(r->ctx->device);
    pg_dev_free(r->ctx, r->d_off);                       // stream-ordered: after everything queued on the context's stream
    pg_dev_free(r->ctx, r->d_planes);
    delete r;
}

extern "C" int64_t pg_reads_count(const pg_reads *r) { return r ? r->count : 0; }

int pg_pack_launch(pg_ctx *ctx, const char *d_bytes, const int64_t *d_off, int64_t count, uint32_t *d_planes)
{
    if (count == 0) return PG_OK;
    const int wpb = 8;
    k_pack<<<(unsigned)((count + wpb - 1) / wpb), wpb * 32, 0, ctx->stream>>>(d_bytes, d_off, count, d_planes);
    PG_LAUNCHED(ctx);
    return PG_OK;
}

int pg_extract_launch(pg_ctx *ctx, const pg_model *md, const uint32_t *d_planes, const int64_t *d_off,
                      int64_t count, uint16_t *d_words, int32_t *d_nwords, uint8_t *d_flags)
{
    if (count == 0) return PG_OK;
    const int wpb = 8;
    k_extract<false><<<(unsigned)((count + wpb - 1) / wpb), wpb * 32, 0, ctx->stream>>>(
        d_planes, NULL, d_off, count, md->d_pdiff, d_words, d_nwords, d_flags);
    PG_LAUNCHED(ctx);
    return PG_OK;
}

// the same from ASCII text on the device: packing and extraction in one pass (no plane store)
int pg_extract_ascii_launch(pg_ctx *ctx, const pg_model *md, const char *d_bytes, const int64_t *d_off,
                            int64_t count, uint16_t *d_words, int32_t *d_nwords, uint8_t *d_flags)
{
    if (count == 0) return PG_OK;
    const int wpb = 8;
    k_extract<true><<<(unsigned)((count + wpb - 1) / wpb), wpb * 32, 0, ctx->stream>>>(
        NULL, d_bytes, d_off, count, md->d_pdiff, d_words, d_nwords, d_flags);
    PG_LAUNCHED(ctx);
    return PG_OK;
}

extern "C" int pg_reads_pack_dev(pg_ctx *ctx, const pg_seqbatch *seqs, int64_t total_bytes, pg_reads **out)
{
    if (!ctx || !seqs || !out || seqs->count < 0 || total_bytes < 0)
        return pg_fail(ctx, PG_EINVAL, "pg_reads_pack_dev: bad arguments");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    pg_reads *r = NULL;
    PG_TRY(reads_alloc(ctx, seqs->count, total_bytes, &r));
    cudaError_t e = cudaMemcpyAsync(r->d_off, seqs->off, (size_t)(seqs->count + 1) * 8,
                                    cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) { pg_reads_free(r); return pg_fail(ctx, PG_ECUDA, "pg_reads_pack_dev: %s", cudaGetErrorString(e)); }
    int rc = pg_pack_launch(ctx, seqs->bytes, r->d_off, r->count, r->d_planes);
    if (rc != PG_OK) { pg_reads_free(r); return rc; }
    *out = r;
    return PG_OK;
}

extern "C" int pg_reads_pack(pg_ctx *ctx, const pg_seqbatch *seqs, pg_reads **out)
{
    if (!ctx || !seqs || !out || seqs->count < 0)
        return pg_fail(ctx, PG_EINVAL, "pg_reads_pack: bad arguments");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    int64_t n = seqs->count, total = n ? seqs->off[n] : 0;
    pg_reads *r = NULL;
    PG_TRY(reads_alloc(ctx, n, total, &r));
    int rc = pg_scratch(ctx, &ctx->s_bytes, (size_t)total + 64);
    if (rc != PG_OK) { pg_reads_free(r); return rc; }
    cudaError_t e;
    if ((e = cudaMemcpyAsync(ctx->s_bytes.p, seqs->bytes, (size_t)total, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(r->d_off, seqs->off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) {
        pg_reads_free(r);
        return pg_fail(ctx, PG_ECUDA, "pg_reads_pack: upload failed: %s", cudaGetErrorString(e));
    }
    rc = pg_pack_launch(ctx, (const char *)ctx->s_bytes.p, r->d_off, n, r->d_planes);
    if (rc != PG_OK) { pg_reads_free(r); return rc; }
    *out = r;
    return PG_OK;
}

// parity hook: planes of read i as interleaved 2-bit codes + mask
extern "C" int64_t pg_reads_unpack(const pg_reads *r, int64_t i, uint32_t *codes, uint32_t *mask, int64_t cap_words)
{
    if (!r || i < 0 || i >= r->count) return PG_EINVAL;
    pg_ctx *ctx = r->ctx;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return PG_ECUDA;
    int64_t o[2];
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
        cudaMemcpy(o, r->d_off + i, 16, cudaMemcpyDeviceToHost) != cudaSuccess)
        return pg_fail(ctx, PG_ECUDA, "pg_reads_unpack: copy failed");
    int64_t len = o[1] - o[0], nch = (len + 31) >> 5;
    if (cap_words < 2 * nch) return pg_fail(ctx, PG_ERANGE, "pg_reads_unpack: need %lld words", (long long)(2 * nch));
    std::vector<uint32_t> pl((size_t)nch * 3);
    if (nch && cudaMemcpy(pl.data(), r->d_planes + 3 * pg_chunk_start(o[0], i), (size_t)nch * 12,
                          cudaMemcpyDeviceToHost) != cudaSuccess)
        return pg_fail(ctx, PG_ECUDA, "pg_reads_unpack: copy failed");
    for (int64_t c = 0; c < nch; c++) {
        uint32_t lo = pl[3 * c], hi = pl[3 * c + 1];
        uint32_t c0 = 0, c1 = 0;
        for (int b = 0; b < 16; b++) {
            c0 |= (((lo >> b) & 1u) | (((hi >> b) & 1u) << 1)) << (2 * b);
            c1 |= (((lo >> (b + 16)) & 1u) | (((hi >> (b + 16)) & 1u) << 1)) << (2 * b);
        }
        codes[2 * c] = c0;
        codes[2 * c + 1] = c1;
        mask[c] = pl[3 * c + 2];
    }
    return len;
}
