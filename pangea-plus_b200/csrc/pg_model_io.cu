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
