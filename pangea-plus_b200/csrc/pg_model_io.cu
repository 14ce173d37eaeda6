// pg_model_io.cu -- model persistence.
//
// The reference's RDP stage carries its trained model inside the jar (four
// files); a drop-in needs an on-disk model of its own.  The file
// holds the INTEGER training counts (sparse, word-major) plus the genus
// lineages; the fp32 tables are re-derived on the GPU at load time, so a saved
// model reproduces the trained one bit for bit.
//
//   "PGMODEL1" | int32 G | int32 depth | int64 N | int64 npairs | int64 blob_len
//   int32 M[G] | int32 nw[65536] | int32 anc[G*depth]
//   int64 idx[65537] | int32 genus[npairs] | int32 count[npairs] | blob
#include "pg_internal.cuh"

// one warp per word: its (genus, count) pairs go to m[genus/32][word][genus%32]
__global__ void k_scatter_counts(const int64_t *__restrict__ idx, const int32_t *__restrict__ genus,
                                 const int32_t *__restrict__ count, int32_t *__restrict__ m)
{
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= PG_NWORDS) return;
    for (int64_t p = idx[w] + (threadIdx.x & 31); p < idx[w + 1]; p += 32) {
        const int g = genus[p];
        m[((size_t)(g >> 5) * PG_NWORDS + w) * PG_GENUS_TILE + (g & 31)] = count[p];
    }
}

static const char kMagic[8] = {'P', 'G', 'M', 'O', 'D', 'E', 'L', '1'};

extern "C" int pg_model_save(const pg_model *md, const char *path, const void *blob, int64_t blob_len)
{
    if (!md || !path || blob_len < 0 || (blob_len > 0 && !blob)) return pg_fail(md ? md->ctx : NULL, PG_EINVAL, "pg_model_save: bad arguments");
    pg_ctx *ctx = md->ctx;
    if (md->tables_only) return pg_fail(ctx, PG_EINVAL, "pg_model_save: this model was built from tables (pg_model_from_tables) and has no counts");
    const int G = md->G;
    std::vector<int32_t> m((size_t)PG_NWORDS * G), nw(PG_NWORDS), M(G);
    int64_t N = 0;
    PG_TRY(pg_model_counts(md, m.data(), nw.data(), M.data(), &N));
    std::vector<int32_t> anc((size_t)G * md->depth);
    if (md->depth)
        PG_CUDA(ctx, pg_copy_sync(ctx, anc.data(), md->d_anc, anc.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<int64_t> idx(PG_NWORDS + 1);
    std::vector<int32_t> pg, pc;
    for (int w = 0; w < PG_NWORDS; w++) {
        idx[w] = (int64_t)pg.size();
        const int32_t *row = m.data() + (size_t)w * G;
        for (int g = 0; g < G; g++)
            if (row[g]) { pg.push_back(g); pc.push_back(row[g]); }
    }
    idx[PG_NWORDS] = (int64_t)pg.size();
    FILE *f = fopen(path, "wb");
    if (!f) return pg_fail(ctx, PG_EIO, "pg_model_save: cannot open %s for writing", path);
    int32_t hdr32[2] = {G, md->depth};
    int64_t hdr64[3] = {N, (int64_t)pg.size(), blob_len};
    bool ok = fwrite(kMagic, 1, 8, f) == 8 && fwrite(hdr32, 4, 2, f) == 2 && fwrite(hdr64, 8, 3, f) == 3 &&
              fwrite(M.data(), 4, G, f) == (size_t)G && fwrite(nw.data(), 4, PG_NWORDS, f) == PG_NWORDS &&
              fwrite(anc.data(), 4, anc.size(), f) == anc.size() &&
              fwrite(idx.data(), 8, idx.size(), f) == idx.size() &&
              fwrite(pg.data(), 4, pg.size(), f) == pg.size() && fwrite(pc.data(), 4, pc.size(), f) == pc.size() &&
              (blob_len == 0 || fwrite(blob, 1, (size_t)blob_len, f) == (size_t)blob_len);
    ok = (fclose(f) == 0) && ok;
    if (!ok) return pg_fail(ctx, PG_EIO, "pg_model_save: short write to %s", path);
    return PG_OK;
}

extern "C" int pg_model_load(pg_ctx *ctx, const char *path, pg_model **out, void **blob, int64_t *blob_len)
{
    if (!ctx || !path || !out) return pg_fail(ctx, PG_EINVAL, "pg_model_load: bad arguments");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    FILE *f = fopen(path, "rb");
    if (!f) return pg_fail(ctx, PG_EIO, "pg_model_load: cannot open %s", path);
    char magic[8];
    int32_t hdr32[2];
    int64_t hdr64[3];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kMagic, 8) != 0 || fread(hdr32, 4, 2, f) != 2 ||
        fread(hdr64, 8, 3, f) != 3) {
        fclose(f);
        return pg_fail(ctx, PG_EFORMAT, "pg_model_load: %s is not a PGMODEL1 file", path);
    }
    const int G = hdr32[0], depth = hdr32[1];
    const int64_t N = hdr64[0], npairs = hdr64[1], blen = hdr64[2];
    if (G <= 0 || depth < 0 || depth > PG_MAX_DEPTH || N < 0 || npairs < 0 || blen < 0 ||
        npairs > (int64_t)PG_NWORDS * G) {
        fclose(f);
        return pg_fail(ctx, PG_EFORMAT, "pg_model_load: %s has an implausible header", path);
    }
    std::vector<int32_t> M(G), nw(PG_NWORDS), anc((size_t)G * depth), pgx((size_t)npairs), pcx((size_t)npairs);
    std::vector<int64_t> idx(PG_NWORDS + 1);
    void *b = NULL;
    bool ok = fread(M.data(), 4, G, f) == (size_t)G && fread(nw.data(), 4, PG_NWORDS, f) == PG_NWORDS &&
              fread(anc.data(), 4, anc.size(), f) == anc.size() && fread(idx.data(), 8, idx.size(), f) == idx.size() &&
              fread(pgx.data(), 4, pgx.size(), f) == pgx.size() && fread(pcx.data(), 4, pcx.size(), f) == pcx.size();
    if (ok && blen > 0 && blob) {
        b = malloc((size_t)blen);
        ok = b && fread(b, 1, (size_t)blen, f) == (size_t)blen;
    }
    fclose(f);
    if (ok) {
        if (idx[0] != 0 || idx[PG_NWORDS] != npairs) ok = false;
        for (int w = 0; ok && w < PG_NWORDS; w++) ok = idx[w] <= idx[w + 1];
        for (int64_t p = 0; ok && p < npairs; p++) ok = pgx[p] >= 0 && pgx[p] < G && pcx[p] > 0;
    }
    if (!ok) {
        free(b);
        return pg_fail(ctx, PG_EFORMAT, "pg_model_load: %s is truncated or corrupt", path);
    }
    pg_model *md = NULL;
    int rc = pg_model_create(ctx, G, &md);
    if (rc != PG_OK) { free(b); return rc; }
    // upload the sparse pairs and scatter them into the genus-tiled layout on the device
    unsigned long long N64 = (unsigned long long)N;
    int64_t *d_idx = NULL;
    int32_t *d_pg = NULL, *d_pc = NULL;
    cudaError_t e;
    if ((e = cudaMalloc(&d_idx, (PG_NWORDS + 1) * 8)) != cudaSuccess ||
        (e = cudaMalloc(&d_pg, (size_t)(npairs + 1) * 4)) != cudaSuccess ||
        (e = cudaMalloc(&d_pc, (size_t)(npairs + 1) * 4)) != cudaSuccess ||
        (e = pg_copy_sync(ctx, d_idx, idx.data(), (PG_NWORDS + 1) * 8, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = pg_copy_sync(ctx, d_pg, pgx.data(), (size_t)npairs * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = pg_copy_sync(ctx, d_pc, pcx.data(), (size_t)npairs * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = pg_copy_sync(ctx, md->d_nw, nw.data(), PG_NWORDS * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = pg_copy_sync(ctx, md->d_M, M.data(), (size_t)G * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = pg_copy_sync(ctx, md->d_N, &N64, 8, cudaMemcpyHostToDevice)) != cudaSuccess) {
        cudaFree(d_idx); cudaFree(d_pg); cudaFree(d_pc);
        pg_model_free(md);
        free(b);
        return pg_fail(ctx, PG_ECUDA, "pg_model_load: upload failed: %s", cudaGetErrorString(e));
    }
    k_scatter_counts<<<PG_NWORDS / 8, 256, 0, ctx->stream>>>(d_idx, d_pg, d_pc, md->d_m);
    ctx->launches++;
    e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_idx); cudaFree(d_pg); cudaFree(d_pc);
    if (e != cudaSuccess) {
        pg_model_free(md);
        free(b);
        return pg_fail(ctx, PG_ECUDA, "pg_model_load: scatter failed: %s", cudaGetErrorString(e));
    }
    rc = pg_model_commit(md);
    if (rc == PG_OK && depth > 0) rc = pg_model_set_lineage(md, anc.data(), depth);
    if (rc != PG_OK) { pg_model_free(md); free(b); return rc; }
    *out = md;
    if (blob) *blob = b;
    if (blob_len) *blob_len = blen;
    return PG_OK;
}


// ------------------------------------------------------------------ models given as TABLES (stock RDP trainset files)
//
// RDP ships its model as log tables, not counts -- logWordPrior (65 536 floats), the genus
// leave counts in the taxonomy tree, and a sparse word-major list of (genus, log conditional probability).  The dense
// table the kernels use is the same thing with the absent cells spelled out: V[w][g] = fp32(logPrior[w] - logLeave[g])
// (row A4), overwritten by the listed cells.  A model built this way has no counts: it classifies, it cannot be saved
// as a .pgm or re-trained.

// one thread per table cell: the absent-cell rule
__global__ void k_table_default(const float *__restrict__ logPrior, const float *__restrict__ logLeave, int G, size_t total,
                                float *__restrict__ table)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int lane = (int)(idx & 31), w = (int)((idx >> 5) & (PG_NWORDS - 1)), g = (int)(idx >> 21) * 32 + lane;
    table[idx] = g < G ? __fsub_rn(logPrior[w], logLeave[g]) : __int_as_float(0xff800000);
}

// one warp per word: the listed cells
__global__ void k_table_scatter(const int64_t *__restrict__ idx, const int32_t *__restrict__ genus, const float *__restrict__ logp,
                                float *__restrict__ table)
{
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= PG_NWORDS) return;
    for (int64_t p = idx[w] + (threadIdx.x & 31); p < idx[w + 1]; p += 32) {
        const int g = genus[p];
        table[((size_t)(g >> 5) * PG_NWORDS + w) * PG_GENUS_TILE + (g & 31)] = logp[p];
    }
}

int pg_model_derive_quantised(pg_model *md);   // pg_certified.cu

// logLeave[g] = (float)ln((float)leaveCount + 1): the expression of k_derive_prior (pg_train.cu), so that a model exported
// from counts and imported again holds the same bits
__global__ void k_log_leave(const int32_t *__restrict__ M, int Gpad, float *__restrict__ logLeave)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Gpad) logLeave[i] = (float)log((double)__fadd_rn((float)M[i], 1.0f));
}

extern "C" int pg_model_from_tables(pg_ctx *ctx, int G, const float *logPrior, const int32_t *leave_count, const int64_t *idx,
                                    const int32_t *genus_of_entry, const float *logp_of_entry, pg_model **out)
{
    if (!ctx || G <= 0 || !logPrior || !leave_count || !idx || !out || (idx[PG_NWORDS] > 0 && (!genus_of_entry || !logp_of_entry)))
        return pg_fail(ctx, PG_EINVAL, "pg_model_from_tables: bad arguments");
    const int64_t nnz = idx[PG_NWORDS];
    if (idx[0] != 0 || nnz < 0 || nnz > (int64_t)PG_NWORDS * G) return pg_fail(ctx, PG_EFORMAT, "pg_model_from_tables: implausible index");
    for (int w = 0; w < PG_NWORDS; w++)
        if (idx[w] > idx[w + 1]) return pg_fail(ctx, PG_EFORMAT, "pg_model_from_tables: index not monotone at word %d", w);
    for (int64_t p = 0; p < nnz; p++)
        if (genus_of_entry[p] < 0 || genus_of_entry[p] >= G) return pg_fail(ctx, PG_EFORMAT, "pg_model_from_tables: genus index %d out of range", genus_of_entry[p]);
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    pg_model *md = NULL;
    PG_TRY(pg_model_create(ctx, G, &md));
    int64_t *d_idx = NULL;
    int32_t *d_g = NULL;
    float *d_p = NULL;
    int rc = PG_OK;
    cudaError_t e;
    if ((e = pg_dev_alloc(ctx, (void **)&d_idx, (PG_NWORDS + 1) * 8)) != cudaSuccess ||
        (e = pg_dev_alloc(ctx, (void **)&d_g, (size_t)(nnz + 1) * 4)) != cudaSuccess ||
        (e = pg_dev_alloc(ctx, (void **)&d_p, (size_t)(nnz + 1) * 4)) != cudaSuccess ||
        (e = pg_copy_sync(ctx, d_idx, idx, (PG_NWORDS + 1) * 8, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (nnz && (e = pg_copy_sync(ctx, d_g, genus_of_entry, (size_t)nnz * 4, cudaMemcpyHostToDevice)) != cudaSuccess) ||
        (nnz && (e = pg_copy_sync(ctx, d_p, logp_of_entry, (size_t)nnz * 4, cudaMemcpyHostToDevice)) != cudaSuccess) ||
        (e = pg_copy_sync(ctx, md->d_logPrior, logPrior, PG_NWORDS * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = pg_copy_sync(ctx, md->d_M, leave_count, (size_t)G * 4, cudaMemcpyHostToDevice)) != cudaSuccess)
        rc = pg_fail(ctx, PG_ECUDA, "pg_model_from_tables: upload failed: %s", cudaGetErrorString(e));
    if (rc == PG_OK) {
        const size_t cells = (size_t)md->ntile * PG_NWORDS * PG_GENUS_TILE;
        k_log_leave<<<(md->ntile * 32 + 255) / 256, 256, 0, ctx->stream>>>(md->d_M, md->ntile * 32, md->d_logLeave);
        ctx->launches++;
        pg_prior_diff_launch(ctx, md);
        k_table_default<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->stream>>>(md->d_logPrior, md->d_logLeave, G, cells, md->d_table);
        ctx->launches++;
        k_table_scatter<<<PG_NWORDS / 8, 256, 0, ctx->stream>>>(d_idx, d_g, d_p, md->d_table);
        ctx->launches++;
        if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) rc = pg_fail(ctx, PG_ECUDA, "pg_model_from_tables: %s", cudaGetErrorString(e));
    }
    pg_dev_free(ctx, d_idx); pg_dev_free(ctx, d_g); pg_dev_free(ctx, d_p);
    if (rc == PG_OK) {
        md->N = 0;
        md->committed = true;
        md->tables_only = true;
        rc = pg_model_derive_quantised(md);
    }
    if (rc != PG_OK) { pg_model_free(md); return rc; }
    *out = md;
    return PG_OK;
}
