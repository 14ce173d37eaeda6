// pg_train.cu -- Stage A training.
//
// Replaces upstream RawHierarchyTree.initWordOccurrence / TreeFactory.addSequence
// (word x genus occurrence counts) and TreeFactory.createGenusWordConditionalProb /
// TrainingInfo.createLogLeaveCount (log tables) of RDP Classifier 2.5, the jar
// invoked at README.md:119.  Integer atomics for the counts, then one
// elementwise pass that writes the dense fp32 table genus-tiled:
//     table[tile][word][32]   (tile = genus / 32)
// so that a word's row segment for one tile is one aligned 128-byte line.
#include "pg_internal.cuh"
#include <algorithm>

// A1 base code: A=0, T/U=1, G=2, C=3, anything else -1.
__device__ __forceinline__ int pg_base_code(unsigned char c)
{
    c &= 0xDF;   // fold case: 'a'..'z' -> 'A'..'Z' (other bytes never map onto ACGTU)
    int code = -1;
    if (c == 'A') code = 0;
    else if (c == 'T' || c == 'U') code = 1;
    else if (c == 'G') code = 2;
    else if (c == 'C') code = 3;
    return code;
}

// K1: one CTA per training sequence.  A 65536-bit shared bitmap collects the
// DISTINCT 8-mers of the sequence (forward strand only); each set bit then
// becomes one atomicAdd on m[w][genus] and one on n[w].
__global__ void __launch_bounds__(256)
k_train_count(const char *__restrict__ bytes, const int64_t *__restrict__ off,
              const int32_t *__restrict__ genus_of_seq, int64_t nseq, int G,
              int32_t *__restrict__ m, int32_t *__restrict__ nw, int32_t *__restrict__ M,
              unsigned long long *__restrict__ N, int *__restrict__ bad)
{
    __shared__ uint32_t bitmap[PG_NWORDS / 32];
    const int64_t s = blockIdx.x;
    if (s >= nseq) return;
    const int g = genus_of_seq[s];
    if (g < 0 || g >= G) {
        if (threadIdx.x == 0) atomicExch(bad, 1);
        return;
    }
    for (int i = threadIdx.x; i < PG_NWORDS / 32; i += blockDim.x) bitmap[i] = 0u;
    __syncthreads();

    const char   *seq = bytes + off[s];
    const int64_t len = off[s + 1] - off[s];
    for (int64_t i = threadIdx.x; i + PG_WORDSIZE <= len; i += blockDim.x) {
        uint32_t w = 0;
        bool ok = true;
#pragma unroll
        for (int j = 0; j < PG_WORDSIZE; j++) {
            int c = pg_base_code((unsigned char)seq[i + j]);
            ok = ok && (c >= 0);
            w = (w << 2) | (uint32_t)(c & 3);
        }
        if (ok) atomicOr(&bitmap[w >> 5], 1u << (w & 31));
    }
    __syncthreads();

    int32_t *mg = m + ((size_t)(g >> 5) * PG_NWORDS) * PG_GENUS_TILE + (g & 31);
    for (int q = threadIdx.x; q < PG_NWORDS / 32; q += blockDim.x) {
        uint32_t bits = bitmap[q];
        while (bits) {
            int b = __ffs(bits) - 1;
            bits &= bits - 1;
            int w = q * 32 + b;
            atomicAdd(mg + (size_t)w * PG_GENUS_TILE, 1);
            atomicAdd(nw + w, 1);
        }
    }
    if (threadIdx.x == 0) {
        atomicAdd(M + g, 1);
        atomicAdd(N, 1ULL);
    }
}

// K2a: word prior P_w, logPrior[w], logLeave[g].  Java keeps these quotients in
// `float`, widens for Math.log and narrows the result (A6).
__global__ void k_derive_prior(const int32_t *__restrict__ nw, const int32_t *__restrict__ M,
                               const unsigned long long *__restrict__ N, int Gpad,
                               float *__restrict__ Pw, float *__restrict__ logPrior,
                               float *__restrict__ logLeave)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float Nf1 = __fadd_rn((float)(long long)(*N), 1.0f);
    if (i < PG_NWORDS) {
        float p = __fdiv_rn(__fadd_rn((float)nw[i], 0.5f), Nf1);
        Pw[i] = p;
        logPrior[i] = (float)log((double)p);
    }
    if (i < Gpad) logLeave[i] = (float)log((double)__fadd_rn((float)M[i], 1.0f));
}

// K2b: dense table, one thread per (tile, word, lane).  Absent entries (m == 0)
// are DEFINED as the fp32 result of logPrior[w] - logLeave[g] (A4); lanes past G
// are -inf so a padded genus can never win an argmax.
__global__ void k_derive_table(const int32_t *__restrict__ m, const int32_t *__restrict__ M,
                               const float *__restrict__ Pw, const float *__restrict__ logPrior,
                               const float *__restrict__ logLeave, int G, size_t total,
                               float *__restrict__ table)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int lane = (int)(idx & 31);
    int w    = (int)((idx >> 5) & (PG_NWORDS - 1));
    int tile = (int)(idx >> 21);
    int g    = tile * 32 + lane;
    float v;
    if (g >= G) {
        v = __int_as_float(0xff800000);
    } else {
        int c = m[idx];
        if (c > 0) {
            float q = __fdiv_rn(__fadd_rn((float)c, Pw[w]), __fadd_rn((float)M[g], 1.0f));
            v = (float)log((double)q);
        } else {
            v = __fsub_rn(logPrior[w], logLeave[g]);
        }
    }
    table[idx] = v;
}

// plain [w][g] views of the tiled arrays for the parity hooks
template <typename T>
__global__ void k_untile(const T *__restrict__ tiled, int G, T *__restrict__ dense)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)PG_NWORDS * G;
    if (idx >= total) return;
    int w = (int)(idx / G), g = (int)(idx % G);
    dense[idx] = tiled[((size_t)(g >> 5) * PG_NWORDS + w) * PG_GENUS_TILE + (g & 31)];
}

static int model_alloc(pg_ctx *ctx, int G, pg_model **out)
{
    if (!ctx || !out || G <= 0) return pg_fail(ctx, PG_EINVAL, "model: bad arguments (G=%d)", G);
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    pg_model *md = new pg_model();
    memset(md, 0, sizeof *md);
    md->ctx = ctx;
    md->G = G;
    md->ntile = (G + PG_GENUS_TILE - 1) / PG_GENUS_TILE;
    md->depth = 0;
    size_t cells = (size_t)md->ntile * PG_NWORDS * PG_GENUS_TILE;
    cudaError_t e;
    if ((e = cudaMalloc(&md->d_m, cells * 4)) != cudaSuccess ||
        (e = cudaMalloc(&md->d_table, cells * 4)) != cudaSuccess ||
        (e = cudaMalloc(&md->d_nw, PG_NWORDS * 4)) != cudaSuccess ||
        (e = cudaMalloc(&md->d_M, (size_t)md->ntile * 32 * 4)) != cudaSuccess ||
        (e = cudaMalloc(&md->d_N, 8)) != cudaSuccess ||
        (e = cudaMalloc(&md->d_logPrior, PG_NWORDS * 4)) != cudaSuccess ||
        (e = cudaMalloc(&md->d_pdiff, PG_NWORDS * 4)) != cudaSuccess ||
        (e = cudaMalloc(&md->d_Pw, PG_NWORDS * 4)) != cudaSuccess ||
        (e = cudaMalloc(&md->d_logLeave, (size_t)md->ntile * 32 * 4)) != cudaSuccess) {
        (void)cudaGetLastError();
        pg_model_free(md);
        return pg_fail(ctx, PG_ENOMEM, "model allocation failed (G=%d): %s", G, cudaGetErrorString(e));
    }
    PG_CUDA(ctx, cudaMemsetAsync(md->d_m, 0, cells * 4, ctx->stream));
    PG_CUDA(ctx, cudaMemsetAsync(md->d_nw, 0, PG_NWORDS * 4, ctx->stream));
    PG_CUDA(ctx, cudaMemsetAsync(md->d_M, 0, (size_t)md->ntile * 32 * 4, ctx->stream));
    PG_CUDA(ctx, cudaMemsetAsync(md->d_N, 0, 8, ctx->stream));
    // callers may follow up with blocking copies on the legacy stream
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = md;
    return PG_OK;
}

extern "C" void pg_model_free(pg_model *md)
{
    if (!md) return;
    cudaSetDevice(md->ctx->device);
    cudaStreamSynchronize(md->ctx->stream);
    cudaFree(md->d_m); cudaFree(md->d_table); cudaFree(md->d_nw); cudaFree(md->d_M);
    cudaFree(md->d_N); cudaFree(md->d_logPrior); cudaFree(md->d_pdiff); cudaFree(md->d_Pw); cudaFree(md->d_logLeave);
    cudaFree(md->d_anc); cudaFree(md->d_qtable); cudaFree(md->d_rowmax);
    cudaFree(md->d_perm); cudaFree(md->d_bmtable); cudaFree(md->d_blockmask); cudaFree(md->d_hmtable); cudaFree(md->d_bm8);
    cudaFree(md->d_qx); cudaFree(md->d_bm8x); cudaFree(md->d_hm8x);
    delete md;
}

extern "C" int pg_model_create(pg_ctx *ctx, int G, pg_model **out) { return model_alloc(ctx, G, out); }

int pg_model_derive_quantised(pg_model *md);   // pg_certified.cu
int pg_model_layout_from_lineage(pg_model *md, const int32_t *anc, int depth);

extern "C" int pg_model_commit(pg_model *md)
{
    if (!md) return PG_EINVAL;
    pg_ctx *ctx = md->ctx;
    if (md->tables_only) return pg_fail(ctx, PG_EINVAL, "pg_model_commit: this model was built from tables and has no counts to derive them from");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    int Gpad = md->ntile * 32;
    int n1 = PG_NWORDS > Gpad ? PG_NWORDS : Gpad;
    k_derive_prior<<<(n1 + 255) / 256, 256, 0, ctx->stream>>>(md->d_nw, md->d_M, md->d_N, Gpad, md->d_Pw,
                                                            md->d_logPrior, md->d_logLeave);
    PG_LAUNCHED(ctx);
    PG_TRY(pg_prior_diff_launch(ctx, md));
    size_t cells = (size_t)md->ntile * PG_NWORDS * PG_GENUS_TILE;
    k_derive_table<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->stream>>>(
        md->d_m, md->d_M, md->d_Pw, md->d_logPrior, md->d_logLeave, md->G, cells, md->d_table);
    PG_LAUNCHED(ctx);
    unsigned long long N = 0;
    PG_CUDA(ctx, cudaMemcpyAsync(&N, md->d_N, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    md->N = (int64_t)N;
    md->committed = true;
    PG_TRY(pg_model_derive_quantised(md));
    return PG_OK;
}

// counts of one batch of sequences added to the model's integer arrays (no tables yet)
static int count_on_device(pg_ctx *ctx, pg_model *md, const char *d_bytes, const int64_t *d_off, int64_t nseq,
                           const int32_t *d_genus)
{
    int *d_bad = NULL;
    PG_CUDA(ctx, cudaMalloc(&d_bad, 4));
    PG_CUDA(ctx, cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
    // gridDim.x is 2^31-1; chunk anyway so the int64 sequence index is explicit
    const int64_t step = 1 << 30;
    for (int64_t s0 = 0; s0 < nseq; s0 += step) {
        int64_t cnt = nseq - s0 < step ? nseq - s0 : step;
        k_train_count<<<(unsigned)cnt, 256, 0, ctx->stream>>>(d_bytes, d_off + s0, d_genus + s0, cnt, md->G,
                                                             md->d_m, md->d_nw, md->d_M, md->d_N, d_bad);
        PG_LAUNCHED(ctx);
    }
    int bad = 0;
    PG_CUDA(ctx, cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_bad);
    if (bad) return pg_fail(ctx, PG_EINVAL, "pg_train: genus_of_seq holds an index outside [0,%d)", md->G);
    md->committed = false;
    return PG_OK;
}

static int train_on_device(pg_ctx *ctx, const char *d_bytes, const int64_t *d_off, int64_t nseq,
                           const int32_t *d_genus, int G, pg_model **out)
{
    pg_model *md = NULL;
    PG_TRY(model_alloc(ctx, G, &md));
    int r = count_on_device(ctx, md, d_bytes, d_off, nseq, d_genus);
    if (r == PG_OK) r = pg_model_commit(md);
    if (r != PG_OK) { pg_model_free(md); return r; }
    *out = md;
    return PG_OK;
}

// Sharded / streamed training: the counts of one more batch of sequences, added to a model made by
// pg_model_create() (or trained before).  Integer atomics, so the order of batches and the way the training set is
// cut across GPUs do not matter: after an integer all-reduce of pg_model_buffers() every rank holds the counts of
// the whole set, bit for bit those of a single pg_train() call.  Tables are derived by pg_model_commit().
extern "C" int pg_train_accumulate_dev(pg_ctx *ctx, pg_model *md, const pg_seqbatch *seqs, const int32_t *genus_dev)
{
    if (!ctx || !md || !seqs || !genus_dev || seqs->count < 0 || md->ctx != ctx)
        return pg_fail(ctx, PG_EINVAL, "pg_train_accumulate_dev: bad arguments");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    return count_on_device(ctx, md, seqs->bytes, seqs->off, seqs->count, genus_dev);
}

extern "C" int pg_train_dev(pg_ctx *ctx, const pg_seqbatch *seqs, const int32_t *genus_dev, int G,
                            pg_model **out)
{
    if (!ctx || !seqs || !genus_dev || !out || seqs->count < 0)
        return pg_fail(ctx, PG_EINVAL, "pg_train_dev: bad arguments");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    return train_on_device(ctx, seqs->bytes, seqs->off, seqs->count, genus_dev, G, out);
}

extern "C" int pg_train(pg_ctx *ctx, const pg_seqbatch *seqs, const int32_t *genus_host, int G,
                        pg_model **out)
{
    if (!ctx || !seqs || !genus_host || !out || seqs->count < 0)
        return pg_fail(ctx, PG_EINVAL, "pg_train: bad arguments");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    int64_t n = seqs->count;
    int64_t total = n ? seqs->off[n] : 0;
    char *d_bytes = NULL; int64_t *d_off = NULL; int32_t *d_genus = NULL;
    cudaError_t e;
    if ((e = cudaMalloc(&d_bytes, (size_t)total + 16)) != cudaSuccess ||
        (e = cudaMalloc(&d_off, (size_t)(n + 1) * 8)) != cudaSuccess ||
        (e = cudaMalloc(&d_genus, (size_t)(n + 1) * 4)) != cudaSuccess) {
        (void)cudaGetLastError();
        cudaFree(d_bytes); cudaFree(d_off); cudaFree(d_genus);
        return pg_fail(ctx, PG_ENOMEM, "pg_train: device allocation failed: %s", cudaGetErrorString(e));
    }
    int r = PG_OK;
    if ((e = cudaMemcpyAsync(d_bytes, seqs->bytes, (size_t)total, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_off, seqs->off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_genus, genus_host, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
        r = pg_fail(ctx, PG_ECUDA, "pg_train: upload failed: %s", cudaGetErrorString(e));
    if (r == PG_OK) r = train_on_device(ctx, d_bytes, d_off, n, d_genus, G, out);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_bytes); cudaFree(d_off); cudaFree(d_genus);
    return r;
}

extern "C" int pg_train_accumulate(pg_ctx *ctx, pg_model *md, const pg_seqbatch *seqs, const int32_t *genus_host)
{
    if (!ctx || !md || !seqs || !genus_host || seqs->count < 0 || md->ctx != ctx)
        return pg_fail(ctx, PG_EINVAL, "pg_train_accumulate: bad arguments");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n = seqs->count;
    if (n == 0) return PG_OK;
    const int64_t base = seqs->off[0], total = seqs->off[n] - base;
    char *d_bytes = NULL; int64_t *d_off = NULL; int32_t *d_genus = NULL;
    cudaError_t e;
    if ((e = cudaMalloc(&d_bytes, (size_t)total + 16)) != cudaSuccess ||
        (e = cudaMalloc(&d_off, (size_t)(n + 1) * 8)) != cudaSuccess ||
        (e = cudaMalloc(&d_genus, (size_t)(n + 1) * 4)) != cudaSuccess) {
        (void)cudaGetLastError();
        cudaFree(d_bytes); cudaFree(d_off); cudaFree(d_genus);
        return pg_fail(ctx, PG_ENOMEM, "pg_train_accumulate: device allocation failed: %s", cudaGetErrorString(e));
    }
    std::vector<int64_t> off((size_t)n + 1);
    for (int64_t i = 0; i <= n; i++) off[(size_t)i] = seqs->off[i] - base;     // a slice of a larger batch: rebase
    int r = PG_OK;
    if ((e = cudaMemcpyAsync(d_bytes, seqs->bytes + base, (size_t)total, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_off, off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_genus, genus_host, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
        r = pg_fail(ctx, PG_ECUDA, "pg_train_accumulate: upload failed: %s", cudaGetErrorString(e));
    if (r == PG_OK) r = count_on_device(ctx, md, d_bytes, d_off, n, d_genus);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_bytes); cudaFree(d_off); cudaFree(d_genus);
    return r;
}

extern "C" int pg_model_set_lineage(pg_model *md, const int32_t *anc_host, int depth)
{
    if (!md || !anc_host || depth <= 0 || depth > PG_MAX_DEPTH)
        return pg_fail(md ? md->ctx : NULL, PG_EINVAL, "pg_model_set_lineage: depth must be 1..%d", PG_MAX_DEPTH);
    pg_ctx *ctx = md->ctx;
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (md->d_anc) { cudaFree(md->d_anc); md->d_anc = NULL; }
    size_t bytes = (size_t)md->G * depth * 4;
    PG_CUDA(ctx, cudaMalloc(&md->d_anc, bytes));
    PG_CUDA(ctx, pg_copy_sync(ctx, md->d_anc, anc_host, bytes, cudaMemcpyHostToDevice));
    md->depth = depth;
    // Certified mode lays the genera out in lineage order, so that the relatives of a read's genus
    // share its 64-genus block and every other block is dismissed by its lower bound
    // (pg_certified.cu).  Results do not depend on the layout; ties resolve on the genus index.
    PG_TRY(pg_model_layout_from_lineage(md, anc_host, depth));
    if (md->committed) PG_TRY(pg_model_derive_quantised(md));
    return PG_OK;
}

extern "C" int pg_model_genera(const pg_model *md) { return md ? md->G : 0; }
extern "C" int pg_model_certifiable(const pg_model *md) { return md && md->q_ok ? 1 : 0; }
extern "C" int pg_model_bound_columns(const pg_model *md) { return !md || !md->bounds_tuned ? -1 : (md->part_bounds ? 1 : 0); }
extern "C" int64_t pg_model_sequences(const pg_model *md) { return md ? md->N : 0; }

extern "C" int pg_model_buffers(pg_model *md, void **dev_ptrs, size_t *nbytes, int max, int *n)
{
    if (!md || !dev_ptrs || !nbytes || max < 4) return PG_EINVAL;
    size_t cells = (size_t)md->ntile * PG_NWORDS * PG_GENUS_TILE;
    dev_ptrs[0] = md->d_m;  nbytes[0] = cells * 4;
    dev_ptrs[1] = md->d_nw; nbytes[1] = PG_NWORDS * 4;
    dev_ptrs[2] = md->d_M;  nbytes[2] = (size_t)md->ntile * 32 * 4;
    dev_ptrs[3] = md->d_N;  nbytes[3] = 8;
    if (n) *n = 4;
    return PG_OK;
}

extern "C" int pg_model_counts(const pg_model *md, int32_t *m_wg, int32_t *n_w, int32_t *M_g, int64_t *N)
{
    if (!md) return PG_EINVAL;
    pg_ctx *ctx = md->ctx;
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (m_wg) {
        size_t total = (size_t)PG_NWORDS * md->G;
        int32_t *d = NULL;
        PG_CUDA(ctx, cudaMalloc(&d, total * 4));
        k_untile<int32_t><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(md->d_m, md->G, d);
        PG_LAUNCHED(ctx);
        PG_CUDA(ctx, cudaMemcpyAsync(m_wg, d, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(d);
    }
    if (n_w) PG_CUDA(ctx, cudaMemcpyAsync(n_w, md->d_nw, PG_NWORDS * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (M_g) PG_CUDA(ctx, cudaMemcpyAsync(M_g, md->d_M, (size_t)md->G * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (N) *N = md->N;
    return PG_OK;
}

extern "C" int pg_model_tables(const pg_model *md, float *logPrior, float *logLeave, float *logP)
{
    if (!md || !md->committed) return PG_EINVAL;
    pg_ctx *ctx = md->ctx;
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (logP) {
        size_t total = (size_t)PG_NWORDS * md->G;
        float *d = NULL;
        PG_CUDA(ctx, cudaMalloc(&d, total * 4));
        k_untile<float><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(md->d_table, md->G, d);
        PG_LAUNCHED(ctx);
        PG_CUDA(ctx, cudaMemcpyAsync(logP, d, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(d);
    }
    if (logPrior) PG_CUDA(ctx, cudaMemcpyAsync(logPrior, md->d_logPrior, PG_NWORDS * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (logLeave) PG_CUDA(ctx, cudaMemcpyAsync(logLeave, md->d_logLeave, (size_t)md->G * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PG_OK;
}
