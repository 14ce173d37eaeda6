// pg_tax.cu -- Stage B: gi -> taxid -> lineage.
//
// Replaces Tax_class/ncbitc.c (`tax_class -c|-s|-g|-t|-n`, main at :841-1004) and the
// per-hit walk of Tax_class/NCBI-taxcollector-0.01.pl:58-300.  The reference forks
// tax_class 10-40 times per BLAST hit and re-opens a .bin file in each; here the three
// .bin files are loaded once, the lineage STRING of every taxid is built once by a kernel
// (thread per taxid: parent walk, rank filter, scientific names, and the script's
// print-time rewrite rules), and a batch of hits is then two direct-index lookups
// (gi -> taxid -> string) plus a byte gather, all on the device.
//
// The .bin layouts are the reference's own (ncbitc.c:98-140), so files written by either
// tool can be read by the other.
#include "pg_internal.cuh"
#include <string>

#define NODE_REC 28
#define NAME_REC 196
#define PG_LIN_CAP 1024          // longest lineage string / raw piece list handled
#define PG_WALK_CAP 128          // longest parent chain followed

struct pg_tax {
    pg_ctx *ctx;
    // host copies (tax_class -t / -n / -s printing)
    std::vector<unsigned char> h_nodes, h_names;
    std::vector<int32_t> h_gi;
    int64_t ngi, nnodes;
    int32_t nnames;
    // device
    int32_t *d_gi2tax;           // [ngi]
    int32_t *d_parent;           // [nnodes]
    int8_t  *d_rank;             // [nnodes]  enum ncbitc_rank, -1 invalid
    char    *d_namepool;         // scientific names, concatenated
    uint32_t *d_nameoff;         // [nnodes+1] (empty range = no scientific name)
    char    *d_linpool;          // lineage strings of every taxid
    int64_t *d_linoff;           // [nnodes+2]; entry nnodes = "[0]Unclassified;"-less empty string
};

static const char *RANK_STR[29] = {
    "class", "family", "forma", "genus", "infraclass", "infraorder", "kingdom", "no rank", "order",
    "parvorder", "phylum", "species", "species group", "species subgroup", "subclass", "subfamily",
    "subgenus", "subkingdom", "suborder", "subphylum", "subspecies", "subtribe", "superclass",
    "superfamily", "superkingdom", "superorder", "superphylum", "tribe", "varietas"};
// index in the taxcollector's @ranklist (taxcollector:228-237), -1 = not a printed rank
//                                   class fam forma genus infc info king norank order parv phylum species
__constant__ int8_t c_rank_idx[29] = {2, 4, -1, 5, -1, -1, 7, -2, 3, -1, 1, 6,
                                      -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, -1, -1, -1, -1};
#define PG_RANK_NORANK (-2)

// ------------------------------------------------------------------ tax_class -c

static std::string cleanup_field(const std::string &f, size_t maxlen)
{
    // ncbitc_cleanup_str (:495-509): "\tvalue\t" -> " value"; two bytes or fewer -> ""
    size_t len = f.size() < maxlen ? f.size() : maxlen;
    if (len <= 2) return std::string();
    std::string r = f.substr(0, len - 1);
    r[0] = ' ';
    return r;
}

static int rank_of(const std::string &s)
{
    for (int i = 0; i < 29; i++)
        if (s == RANK_STR[i]) return i;
    return -1;
}

static std::vector<std::string> split_bar(const char *line)
{
    std::vector<std::string> f;
    const char *p = line;
    for (;;) {
        const char *b = strchr(p, '|');
        if (!b) { f.push_back(std::string(p)); break; }
        f.push_back(std::string(p, (size_t)(b - p)));
        p = b + 1;
    }
    return f;
}

extern "C" int pg_tax_build(const char *dir)
{
    if (!dir) return pg_fail(NULL, PG_EINVAL, "pg_tax_build: no directory");
    std::string d(dir);
    char line[4096];
    // ---- gi_taxid_nucl.dmp -> dense int32, gaps zero (ncbitc_create_gi_taxid :701-748)
    {
        FILE *fi = fopen((d + "/gi_taxid_nucl.dmp").c_str(), "r");
        if (!fi) return pg_fail(NULL, PG_EIO, "pg_tax_build: cannot open %s/gi_taxid_nucl.dmp", dir);
        FILE *fo = fopen((d + "/gi_taxid_nucl.dmp.bin").c_str(), "wb");
        if (!fo) { fclose(fi); return pg_fail(NULL, PG_EIO, "pg_tax_build: cannot write %s/gi_taxid_nucl.dmp.bin", dir); }
        int last = 0, zero = 0;
        while (fgets(line, sizeof line, fi)) {
            int gi = 0, tax = 0;
            if (sscanf(line, "%d\t%d", &gi, &tax) < 2) continue;
            for (int i = 1; i < gi - last; i++) fwrite(&zero, 4, 1, fo);
            fwrite(&tax, 4, 1, fo);
            last = gi;
        }
        fclose(fi);
        if (fclose(fo)) return pg_fail(NULL, PG_EIO, "pg_tax_build: short write");
    }
    // ---- nodes.dmp -> 28-byte records at (taxid-1)*28 (ncbitc_create_nodes :750-794)
    {
        FILE *fi = fopen((d + "/nodes.dmp").c_str(), "r");
        if (!fi) return pg_fail(NULL, PG_EIO, "pg_tax_build: cannot open %s/nodes.dmp", dir);
        FILE *fo = fopen((d + "/nodes.dmp.bin").c_str(), "wb");
        if (!fo) { fclose(fi); return pg_fail(NULL, PG_EIO, "pg_tax_build: cannot write %s/nodes.dmp.bin", dir); }
        unsigned char zero[NODE_REC] = {0}, rec[NODE_REC];
        int last = 0;
        while (fgets(line, sizeof line, fi)) {
            std::vector<std::string> f = split_bar(line);
            if (f.size() < 12) continue;
            memset(rec, 0, sizeof rec);
            int32_t tax = atoi(f[0].c_str()), parent = atoi(f[1].c_str());
            memcpy(rec + 0, &tax, 4);
            memcpy(rec + 4, &parent, 4);
            std::string rk = cleanup_field(f[2], 32);
            rec[8] = (unsigned char)(signed char)rank_of(rk.empty() ? rk : rk.substr(1));
            std::string em = cleanup_field(f[3], 32);
            if (!em.empty()) { rec[9] = em.size() > 1 ? (unsigned char)em[1] : 0; rec[10] = em.size() > 2 ? (unsigned char)em[2] : 0; }
            int16_t div = (int16_t)atoi(f[4].c_str()), gc = (int16_t)atoi(f[6].c_str());
            int32_t mgc = atoi(f[8].c_str());
            memcpy(rec + 12, &div, 2);
            rec[14] = (unsigned char)atoi(f[5].c_str());
            memcpy(rec + 16, &gc, 2);
            rec[18] = (unsigned char)atoi(f[7].c_str());
            memcpy(rec + 20, &mgc, 4);
            rec[24] = (unsigned char)atoi(f[9].c_str());
            rec[25] = (unsigned char)atoi(f[10].c_str());
            rec[26] = (unsigned char)atoi(f[11].c_str());
            for (int i = 1; i < tax - last; i++) fwrite(zero, NODE_REC, 1, fo);
            fwrite(rec, NODE_REC, 1, fo);
            last = tax;
        }
        fclose(fi);
        if (fclose(fo)) return pg_fail(NULL, PG_EIO, "pg_tax_build: short write");
    }
    // ---- names.dmp -> int32 count + 196-byte records in file order (ncbitc_create_names :796-839)
    {
        FILE *fi = fopen((d + "/names.dmp").c_str(), "r");
        if (!fi) return pg_fail(NULL, PG_EIO, "pg_tax_build: cannot open %s/names.dmp", dir);
        FILE *fo = fopen((d + "/names.dmp.bin").c_str(), "wb");
        if (!fo) { fclose(fi); return pg_fail(NULL, PG_EIO, "pg_tax_build: cannot write %s/names.dmp.bin", dir); }
        int32_t num = 0;
        fwrite(&num, 4, 1, fo);
        unsigned char rec[NAME_REC];
        while (fgets(line, sizeof line, fi)) {
            std::vector<std::string> f = split_bar(line);
            if (f.size() < 4) continue;
            memset(rec, 0, sizeof rec);
            int32_t tax = atoi(f[0].c_str());
            memcpy(rec, &tax, 4);
            std::string a = cleanup_field(f[1], 64), b = cleanup_field(f[2], 64), c = cleanup_field(f[3], 32);
            memcpy(rec + 4, a.data(), a.size() < 63 ? a.size() : 63);
            memcpy(rec + 68, b.data(), b.size() < 63 ? b.size() : 63);
            memcpy(rec + 132, c.data(), c.size() < 63 ? c.size() : 63);
            fwrite(rec, NAME_REC, 1, fo);
            num++;
        }
        fseek(fo, 0, SEEK_SET);
        fwrite(&num, 4, 1, fo);
        fclose(fi);
        if (fclose(fo)) return pg_fail(NULL, PG_EIO, "pg_tax_build: short write");
    }
    return PG_OK;
}

// ------------------------------------------------------------------ device: lineage strings

struct TaxDev {
    const int32_t *parent;
    const int8_t *rank;
    const char *namepool;
    const uint32_t *nameoff;
    int64_t nnodes;
};

__device__ __forceinline__ bool pg_is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

// taxcollector:92-155 on the '|'-separated pieces in raw[0..rn): elements are printed last to
// first; an element containing '6' gets blanks -> '_' and, if NO element holds a '5', is
// printed twice (first '6' as '5', then restored); otherwise a first '7' becomes '9'.
// Returns the output length; writes when out != NULL.
__device__ int pg_format_pieces(char *raw, int rn, char *out)
{
    // element boundaries
    short start[176], len[176];
    int ne = 0, p = 0;
    while (p <= rn && ne < 176) {            // an element with its '|' is at least 6 bytes: 1024/6 < 176
        int q = p;
        while (q < rn && raw[q] != '|') q++;
        start[ne] = p;
        len[ne] = q - p;
        ne++;
        if (q >= rn) break;
        p = q + 1;
    }
    while (ne > 0 && len[ne - 1] == 0) ne--;
    int o = 0;
    for (int i = ne - 1; i >= 0; i--) {
        char *e = raw + start[i];
        const int n = len[i];
        int six = -1, seven = -1;
        for (int k = 0; k < n; k++) {
            if (e[k] == '6' && six < 0) six = k;
            if (e[k] == '7' && seven < 0) seven = k;
        }
        if (six >= 0) {
            for (int k = 0; k < n; k++) if (pg_is_space(e[k])) e[k] = '_';
            bool any5 = false;
            for (int j = 0; j < ne && !any5; j++)
                for (int k = 0; k < len[j]; k++) if (raw[start[j] + k] == '5') { any5 = true; break; }
            if (!any5) {
                if (out) for (int k = 0; k < n; k++) out[o + k] = (k == six) ? '5' : e[k];
                o += n;
            }
            if (out) for (int k = 0; k < n; k++) out[o + k] = e[k];
            o += n;
        } else {
            if (out) for (int k = 0; k < n; k++) out[o + k] = (k == seven) ? '9' : e[k];
            o += n;
        }
    }
    return o;
}

// get_uptaxa (taxcollector:226-300) for one leaf taxid: the raw piece list.
// Returns its length, or -1 if it does not fit PG_LIN_CAP.
__device__ int pg_walk_pieces(const TaxDev &T, int taxid, char *raw)
{
    int rn = 0;
    for (int depth = 0; depth < PG_WALK_CAP; depth++) {
        if (taxid <= 0 || taxid > T.nnodes) break;
        const int parent = T.parent[taxid - 1];
        const int rk = T.rank[taxid - 1];
        const int idx = (rk >= 0 && rk < 29) ? c_rank_idx[rk] : -1;
        if (idx >= 0) {
            const uint32_t a = T.nameoff[taxid - 1], b = T.nameoff[taxid];
            if (rn + 3 + (int)(b - a) + 2 > PG_LIN_CAP) return -1;
            raw[rn++] = '['; raw[rn++] = (char)('0' + idx); raw[rn++] = ']';
            if (b > a) {
                for (uint32_t k = a; k < b; k++) raw[rn++] = T.namepool[k];
                raw[rn++] = ';'; raw[rn++] = '|';
            }
            if (idx == 0) break;                              // superkingdom: stop
            taxid = parent;
        } else if (idx == PG_RANK_NORANK) {
            if (parent == 1) {
                const char u[] = "[0]Unclassified;|";
                if (rn + 17 > PG_LIN_CAP) return -1;
                for (int k = 0; k < 17; k++) raw[rn++] = u[k];
                break;
            }
            taxid = parent;
        } else {
            taxid = parent;
        }
    }
    return rn;
}

// pass 1 (out == NULL): lengths; pass 2: strings.  One thread per taxid (1-based id = t+1).
__global__ void k_lineage_table(TaxDev T, const int64_t *__restrict__ linoff, int64_t *__restrict__ linlen,
                                char *__restrict__ linpool, int *__restrict__ too_long)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T.nnodes) return;
    char raw[PG_LIN_CAP];
    const int rn = pg_walk_pieces(T, (int)(t + 1), raw);
    if (rn < 0) {
        atomicExch(too_long, 1);
        if (linlen) linlen[t] = 0;
        return;
    }
    if (linlen) linlen[t] = pg_format_pieces(raw, rn, NULL);
    else pg_format_pieces(raw, rn, linpool + linoff[t]);
}

// "Unidentified(GI:<gi>);" goes through the same print rules as any other element
__device__ int pg_unidentified(int gi, char *out)
{
    char raw[48];
    const char head[] = "Unidentified(GI:";
    int rn = 0;
    for (int k = 0; k < 16; k++) raw[rn++] = head[k];
    char dig[12];
    int nd = 0;
    long long v = gi;
    if (v < 0) { raw[rn++] = '-'; v = -v; }
    do { dig[nd++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (nd) raw[rn++] = dig[--nd];
    raw[rn++] = ')'; raw[rn++] = ';'; raw[rn++] = '|';
    return pg_format_pieces(raw, rn, out);
}

__global__ void k_tax_leaf(const int32_t *__restrict__ gi2tax, int64_t ngi, const int32_t *__restrict__ gi, int64_t n,
                           int32_t *__restrict__ leaf)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = gi[i];
    leaf[i] = (g >= 1 && g <= ngi) ? gi2tax[g - 1] : 0;
}

// per hit: length of its lineage string (pass 1) / copy of the string (pass 2)
__global__ void k_hit_lineage(const int32_t *__restrict__ leaf, const int32_t *__restrict__ gi, int64_t n, int64_t nnodes,
                              const int64_t *__restrict__ linoff, const char *__restrict__ linpool,
                              const int64_t *__restrict__ outoff, int64_t *__restrict__ outlen, char *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = leaf[i];
    if (t >= 1 && t <= nnodes) {
        const int64_t a = linoff[t - 1], b = linoff[t];
        if (outlen) outlen[i] = b - a;
        else for (int64_t k = a; k < b; k++) out[outoff[i] + (k - a)] = linpool[k];
    } else {
        if (outlen) outlen[i] = pg_unidentified(gi[i], NULL);
        else pg_unidentified(gi[i], out + outoff[i]);
    }
}

// pass 2 with one WARP per hit: the lineage string of the leaf is copied with coalesced byte loads and stores (the
// thread-per-hit loop above moves every string through 110 scattered single-byte transactions per lane)
__global__ void __launch_bounds__(256)
k_hit_lineage_copy(const int32_t *__restrict__ leaf, const int32_t *__restrict__ gi, int64_t n, int64_t nnodes,
                   const int64_t *__restrict__ linoff, const char *__restrict__ linpool,
                   const int64_t *__restrict__ outoff, char *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int t = leaf[i];
    char *dst = out + outoff[i];
    if (t >= 1 && t <= nnodes) {
        const int64_t a = linoff[t - 1];
        const int len = (int)(linoff[t] - a);
        const char *src = linpool + a;
        for (int k = lane; k < len; k += 32) dst[k] = src[k];
    } else if (lane == 0) {
        pg_unidentified(gi[i], dst);
    }
}

// ------------------------------------------------------------------ load

static int slurp(const std::string &path, std::vector<unsigned char> &buf)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)n);
    size_t got = n ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n ? 0 : -1;
}

extern "C" void pg_tax_free(pg_tax *t)
{
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    cudaFree(t->d_gi2tax); cudaFree(t->d_parent); cudaFree(t->d_rank); cudaFree(t->d_namepool);
    cudaFree(t->d_nameoff); cudaFree(t->d_linpool); cudaFree(t->d_linoff);
    delete t;
}

extern "C" int pg_tax_load(pg_ctx *ctx, const char *dir, pg_tax **out)
{
    if (!ctx || !dir || !out) return pg_fail(ctx, PG_EINVAL, "pg_tax_load: bad arguments");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    pg_tax *t = new pg_tax();
    t->ctx = ctx;
    t->d_gi2tax = NULL; t->d_parent = NULL; t->d_rank = NULL; t->d_namepool = NULL;
    t->d_nameoff = NULL; t->d_linpool = NULL; t->d_linoff = NULL;
    std::string d(dir);
    std::vector<unsigned char> gi;
    if (slurp(d + "/gi_taxid_nucl.dmp.bin", gi) || slurp(d + "/nodes.dmp.bin", t->h_nodes) ||
        slurp(d + "/names.dmp.bin", t->h_names) || t->h_names.size() < 4) {
        delete t;
        return pg_fail(ctx, PG_EIO, "pg_tax_load: cannot read the three .bin files in %s (run tax_class -c)", dir);
    }
    t->ngi = (int64_t)(gi.size() / 4);
    t->h_gi.resize((size_t)t->ngi);
    if (t->ngi) memcpy(t->h_gi.data(), gi.data(), (size_t)t->ngi * 4);
    t->nnodes = (int64_t)(t->h_nodes.size() / NODE_REC);
    memcpy(&t->nnames, t->h_names.data(), 4);
    if (t->nnames < 0 || (size_t)t->nnames * NAME_REC + 4 > t->h_names.size()) {
        delete t;
        return pg_fail(ctx, PG_EFORMAT, "pg_tax_load: names.dmp.bin is truncated");
    }
    // parent / rank columns
    std::vector<int32_t> parent((size_t)t->nnodes);
    std::vector<int8_t> rank((size_t)t->nnodes);
    for (int64_t i = 0; i < t->nnodes; i++) {
        memcpy(&parent[(size_t)i], t->h_nodes.data() + (size_t)i * NODE_REC + 4, 4);
        rank[(size_t)i] = (int8_t)t->h_nodes[(size_t)i * NODE_REC + 8];
    }
    // scientific name per taxid: first record of the taxid's run whose class matches
    // /scientific name/, trimmed (taxcollector:188-224).  ncbitc_search_name (:647-699)
    // never sees the LAST record of the file, so neither does this index.
    std::vector<uint32_t> nameoff((size_t)t->nnodes + 1, 0);
    std::string pool;
    {
        std::vector<std::string> nm((size_t)t->nnodes);
        std::vector<char> has((size_t)t->nnodes, 0);
        for (int32_t r = 0; r + 1 < t->nnames; r++) {
            const unsigned char *rec = t->h_names.data() + 4 + (size_t)r * NAME_REC;
            int32_t tax;
            memcpy(&tax, rec, 4);
            if (tax < 1 || tax > t->nnodes || has[(size_t)tax - 1]) continue;
            const char *cls = (const char *)rec + 132;
            if (!memmem(cls, strnlen(cls, 64), "scientific name", 15)) continue;
            std::string s;
            for (const char *p = (const char *)rec + 4; *p && s.size() < 63; p++)
                if (*p != '\t') s.push_back(*p);
            size_t a = 0, b = s.size();
            while (a < b && strchr(" \n\r\f\v", s[a])) a++;
            while (b > a && strchr(" \n\r\f\v", s[b - 1])) b--;
            nm[(size_t)tax - 1] = s.substr(a, b - a);
            has[(size_t)tax - 1] = 1;
        }
        for (int64_t i = 0; i < t->nnodes; i++) {
            nameoff[(size_t)i] = (uint32_t)pool.size();
            if (has[(size_t)i]) pool += nm[(size_t)i];
        }
        nameoff[(size_t)t->nnodes] = (uint32_t)pool.size();
    }
    cudaError_t e;
    if ((e = cudaMalloc(&t->d_gi2tax, (size_t)(t->ngi + 1) * 4)) != cudaSuccess ||
        (e = cudaMalloc(&t->d_parent, (size_t)(t->nnodes + 1) * 4)) != cudaSuccess ||
        (e = cudaMalloc(&t->d_rank, (size_t)t->nnodes + 1)) != cudaSuccess ||
        (e = cudaMalloc(&t->d_namepool, pool.size() + 1)) != cudaSuccess ||
        (e = cudaMalloc(&t->d_nameoff, (size_t)(t->nnodes + 1) * 4)) != cudaSuccess ||
        (e = cudaMalloc(&t->d_linoff, (size_t)(t->nnodes + 2) * 8)) != cudaSuccess) {
        (void)cudaGetLastError();
        pg_tax_free(t);
        return pg_fail(ctx, PG_ENOMEM, "pg_tax_load: device allocation failed: %s", cudaGetErrorString(e));
    }
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PG_CUDA(ctx, pg_copy_sync(ctx, t->d_gi2tax, t->h_gi.data(), (size_t)t->ngi * 4, cudaMemcpyHostToDevice));
    PG_CUDA(ctx, pg_copy_sync(ctx, t->d_parent, parent.data(), (size_t)t->nnodes * 4, cudaMemcpyHostToDevice));
    PG_CUDA(ctx, pg_copy_sync(ctx, t->d_rank, rank.data(), (size_t)t->nnodes, cudaMemcpyHostToDevice));
    PG_CUDA(ctx, pg_copy_sync(ctx, t->d_namepool, pool.data(), pool.size(), cudaMemcpyHostToDevice));
    PG_CUDA(ctx, pg_copy_sync(ctx, t->d_nameoff, nameoff.data(), (size_t)(t->nnodes + 1) * 4, cudaMemcpyHostToDevice));

    // lineage string of every taxid: lengths, scan, strings
    if (t->nnodes > 0) {
        TaxDev T = {t->d_parent, t->d_rank, t->d_namepool, t->d_nameoff, t->nnodes};
        int64_t *d_len = NULL;
        int *d_flag = NULL;
        PG_CUDA(ctx, cudaMalloc(&d_len, (size_t)t->nnodes * 8));
        PG_CUDA(ctx, cudaMalloc(&d_flag, 4));
        PG_CUDA(ctx, cudaMemsetAsync(d_flag, 0, 4, ctx->stream));
        const unsigned nb = (unsigned)((t->nnodes + 127) / 128);
        k_lineage_table<<<nb, 128, 0, ctx->stream>>>(T, NULL, d_len, NULL, d_flag);
        PG_LAUNCHED(ctx);
        PG_TRY(pg_device_scan(ctx, d_len, t->nnodes, t->d_linoff));
        int64_t total = 0;
        int flag = 0;
        PG_CUDA(ctx, pg_copy_sync(ctx, &total, t->d_linoff + t->nnodes, 8, cudaMemcpyDeviceToHost));
        PG_CUDA(ctx, pg_copy_sync(ctx, &flag, d_flag, 4, cudaMemcpyDeviceToHost));
        cudaFree(d_len);
        cudaFree(d_flag);
        if (flag) {
            pg_tax_free(t);
            return pg_fail(ctx, PG_ERANGE, "pg_tax_load: a lineage needs more than %d bytes", PG_LIN_CAP);
        }
        PG_CUDA(ctx, cudaMalloc(&t->d_linpool, (size_t)total + 16));
        k_lineage_table<<<nb, 128, 0, ctx->stream>>>(T, t->d_linoff, NULL, t->d_linpool, NULL);
        PG_LAUNCHED(ctx);
        PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        PG_CUDA(ctx, cudaMemset(t->d_linoff, 0, 16));
    }
    *out = t;
    return PG_OK;
}

// ------------------------------------------------------------------ batched lookups

extern "C" int pg_tax_leaf(pg_ctx *ctx, const pg_tax *t, const int32_t *gi_host, int64_t n, int32_t *taxid_host)
{
    if (!ctx || !t || !gi_host || !taxid_host || n < 0) return pg_fail(ctx, PG_EINVAL, "pg_tax_leaf: bad arguments");
    if (n == 0) return PG_OK;
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    PG_TRY(pg_scratch(ctx, &ctx->s_bytes, (size_t)n * 8));
    int32_t *d_gi = (int32_t *)ctx->s_bytes.p, *d_leaf = d_gi + n;
    PG_CUDA(ctx, cudaMemcpyAsync(d_gi, gi_host, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    k_tax_leaf<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(t->d_gi2tax, t->ngi, d_gi, n, d_leaf);
    PG_LAUNCHED(ctx);
    PG_CUDA(ctx, cudaMemcpyAsync(taxid_host, d_leaf, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PG_OK;
}

extern "C" int pg_tax_lineage(pg_ctx *ctx, const pg_tax *t, const int32_t *gi_host, int64_t n, char *out_bytes,
                              int64_t out_cap, int64_t *out_off)
{
    if (!ctx || !t || !gi_host || !out_off || n < 0 || out_cap < 0 || (out_cap > 0 && !out_bytes))
        return pg_fail(ctx, PG_EINVAL, "pg_tax_lineage: bad arguments");
    out_off[0] = 0;
    if (n == 0) return PG_OK;
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    PG_TRY(pg_scratch(ctx, &ctx->s_bytes, (size_t)n * 8));
    PG_TRY(pg_scratch(ctx, &ctx->s_off, (size_t)(n + 1) * 16));
    int32_t *d_gi = (int32_t *)ctx->s_bytes.p, *d_leaf = d_gi + n;
    int64_t *d_len = (int64_t *)ctx->s_off.p, *d_off = d_len + (n + 1);
    PG_CUDA(ctx, cudaMemcpyAsync(d_gi, gi_host, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    const unsigned nb = (unsigned)((n + 255) / 256);
    k_tax_leaf<<<nb, 256, 0, ctx->stream>>>(t->d_gi2tax, t->ngi, d_gi, n, d_leaf);
    PG_LAUNCHED(ctx);
    k_hit_lineage<<<nb, 256, 0, ctx->stream>>>(d_leaf, d_gi, n, t->nnodes, t->d_linoff, t->d_linpool, NULL, d_len, NULL);
    PG_LAUNCHED(ctx);
    PG_TRY(pg_device_scan(ctx, d_len, n, d_off));
    PG_CUDA(ctx, cudaMemcpyAsync(out_off, d_off, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int64_t total = out_off[n];
    if (total > out_cap) return pg_fail(ctx, PG_ERANGE, "pg_tax_lineage: output needs %lld bytes", (long long)total);
    if (total == 0) return PG_OK;
    PG_TRY(pg_scratch(ctx, &ctx->s_results, (size_t)total + 16));
    char *d_out = (char *)ctx->s_results.p;
    k_hit_lineage_copy<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(d_leaf, d_gi, n, t->nnodes, t->d_linoff, t->d_linpool, d_off, d_out);
    PG_LAUNCHED(ctx);
    PG_CUDA(ctx, cudaMemcpyAsync(out_bytes, d_out, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PG_OK;
}

// ------------------------------------------------------------------ tax_class -s / -g / -t / -n views
// Record access for the drop-in tax_class executable: the chain is resolved on the device
// (pg_tax_chain), the 28-/196-byte records it prints come from the loaded files.

__global__ void k_tax_chain(const int32_t *__restrict__ parent, int64_t nnodes, const int32_t *__restrict__ start,
                            int64_t n, int maxlen, int32_t *__restrict__ chain, int32_t *__restrict__ chain_len)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int t = start[i], k = 0;
    // ncbitc.c:941-952: while (tax_id != 1) { node = lookup(tax_id); tax_id = node.parent; if (tax_id != 1) print node }
    while (t != 1 && k < maxlen) {
        if (t < 1 || t > nnodes) { k = -1 - k; break; }          // past the table: the reference reads garbage
        const int p = parent[t - 1];
        if (p != 1) chain[i * maxlen + k++] = t;
        t = p;
    }
    chain_len[i] = k;
}

extern "C" int pg_tax_chain(pg_ctx *ctx, const pg_tax *t, const int32_t *taxid_host, int64_t n, int32_t maxlen,
                            int32_t *chain_host, int32_t *chain_len_host)
{
    if (!ctx || !t || !taxid_host || !chain_host || !chain_len_host || n < 0 || maxlen <= 0)
        return pg_fail(ctx, PG_EINVAL, "pg_tax_chain: bad arguments");
    if (n == 0) return PG_OK;
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    PG_TRY(pg_scratch(ctx, &ctx->s_bytes, (size_t)n * 4 * (2 + (size_t)maxlen)));
    int32_t *d_start = (int32_t *)ctx->s_bytes.p, *d_len = d_start + n, *d_chain = d_len + n;
    PG_CUDA(ctx, cudaMemcpyAsync(d_start, taxid_host, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    k_tax_chain<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(t->d_parent, t->nnodes, d_start, n, maxlen, d_chain, d_len);
    PG_LAUNCHED(ctx);
    PG_CUDA(ctx, cudaMemcpyAsync(chain_host, d_chain, (size_t)n * maxlen * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(ctx, cudaMemcpyAsync(chain_len_host, d_len, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PG_OK;
}

extern "C" int pg_tax_node_record(const pg_tax *t, int32_t taxid, void *rec28)
{
    if (!t || !rec28) return PG_EINVAL;
    if (taxid < 1 || taxid > t->nnodes) return PG_ERANGE;
    memcpy(rec28, t->h_nodes.data() + (size_t)(taxid - 1) * NODE_REC, NODE_REC);
    return PG_OK;
}

// records of `taxid` as tax_class -n sees them (run inside records 0..num-2); returns the count
extern "C" int pg_tax_name_records(const pg_tax *t, int32_t taxid, void *rec196, int32_t max_records)
{
    if (!t) return PG_EINVAL;
    int lo = 0, hi = t->nnames - 2, found = -1;
    const unsigned char *base = t->h_names.data() + 4;
    while (lo <= hi) {
        int mid = (lo + hi) / 2, v;
        memcpy(&v, base + (size_t)mid * NAME_REC, 4);
        if (v == taxid) { found = mid; break; }
        if (v < taxid) lo = mid + 1; else hi = mid - 1;
    }
    if (found < 0) return 0;
    int first = found;
    while (first > 0) {
        int v;
        memcpy(&v, base + (size_t)(first - 1) * NAME_REC, 4);
        if (v != taxid) break;
        first--;
    }
    int cnt = 0;
    for (int r = first; r <= t->nnames - 2; r++) {
        int v;
        memcpy(&v, base + (size_t)r * NAME_REC, 4);
        if (v != taxid) break;
        if (rec196 && cnt < max_records) memcpy((unsigned char *)rec196 + (size_t)cnt * NAME_REC, base + (size_t)r * NAME_REC, NAME_REC);
        cnt++;
    }
    return cnt;
}

extern "C" int64_t pg_tax_max_gi(const pg_tax *t) { return t ? t->ngi : 0; }
extern "C" int64_t pg_tax_max_taxid(const pg_tax *t) { return t ? t->nnodes : 0; }
