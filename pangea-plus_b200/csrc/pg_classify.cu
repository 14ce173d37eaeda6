// pg_classify.cu -- K4/K5: gather-sum, 100 bootstraps, argmax, vote
// (SURVEY.md 8(a) rows A4, A7, A8, A9).
//
// Replaces upstream Classifier.classify of RDP Classifier 2.5 (the jar invoked
// at README.md:119): phase 1 per-word rows, phase 2 full-sum assignment,
// phase 3 bootstrap with java.util.Random re-seeded to 1 per read, and the
// per-read ancestor vote.
//
// Strict mode keeps the reference's arithmetic exactly: every sum is a chain of
// single fp32 adds in the reference's order (word order for the assignment,
// draw order for each replicate), and every argmax is "first strict maximum in
// ascending genus index".  Parallelism is over (read, genus block, replicate),
// never over the adds of one sum.
//
// Layout: one CTA = one read x one block of TG = 4*LPR genera.  The read's n
// table rows (TG floats each) are staged once in shared memory with 16-byte
// cp.async copies from the genus-tiled table; each group of LPR lanes then
// owns one replicate at a time and walks its sample list with one LDS.128 and
// four dependent FADDs per draw.  The grid is (reads, genus blocks) with reads
// fastest, so all CTAs in flight gather from the same 65536 x 128 B table slab,
// which stays resident in L2.
#include "pg_classify_common.cuh"
#include <algorithm>

// ------------------------------------------------------------------ K4 strict

#define PG_ADD4(v)                                                              \
    a0 = __fadd_rn(a0, (v).x); a1 = __fadd_rn(a1, (v).y);                       \
    a2 = __fadd_rn(a2, (v).z); a3 = __fadd_rn(a3, (v).w);

// Warp 0 computes the A7 full sum with one genus per lane (the longest chain of
// the CTA: n dependent adds); the other warps split into groups of LPR lanes,
// each group owning one A8 replicate at a time.  The replicate loop is software
// pipelined two 4-draw batches deep (offsets one batch further), so the LDS and
// list-load latencies overlap the FADD chains of the previous batch.
template <int LPR, int BLOCK>
__global__ void __launch_bounds__(BLOCK, (BLOCK > 512) ? 1 : ((BLOCK > 256) ? 2 : 5))
k_classify_strict(const float *__restrict__ table, const uint16_t *__restrict__ words,
                  const int64_t *__restrict__ off, const int32_t *__restrict__ nwords,
                  const int32_t *__restrict__ order, int64_t slot0,
                  const uint32_t *__restrict__ boot_pool, const int32_t *__restrict__ boot_off,
                  int min_boot, unsigned long long *__restrict__ best)
{
    constexpr int TG = 4 * LPR;                    // genera per CTA
    constexpr int NGR = (BLOCK - 32) / LPR;        // replicate groups
    constexpr int SH = (LPR == 8) ? 0 : ((LPR == 4) ? 1 : 2);   // list offsets are row*128
    constexpr int IL = 32 / LPR;                   // replicates interleaved in the list
    extern __shared__ float4 sV[];                 // (n+1) rows x LPR float4; row n is all zero

    const int tid = threadIdx.x;
    const int64_t read = order[blockIdx.x];
    const int n = nwords[read];
    const int gbase = blockIdx.y * TG;
    const float *tbase = table + ((size_t)(gbase >> 5) * PG_NWORDS) * PG_GENUS_TILE + (gbase & 31);
    const uint16_t *w = words + off[read];
    unsigned long long *mybest = best + ((size_t)slot0 + blockIdx.x) * (PG_NUM_BOOT + 1);   // slot = position in the chunk's order array

    // ---- A4: stage the read's rows for this genus block
    for (int c = tid; c < n * LPR; c += BLOCK) {
        const int r = c / LPR, l = c % LPR;
        pg_cp_async16(&sV[c], tbase + (size_t)w[r] * PG_GENUS_TILE + l * 4);
    }
    if (tid < LPR) sV[n * LPR + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
    pg_cp_async_wait_all();
    __syncthreads();

    if (tid < 32) {
        // ---- A7: score[g] = ((0 + v[0][g]) + v[1][g]) + ... in word order; lane = genus
        if (tid >= TG) return;
        constexpr unsigned fmask = (TG == 32) ? 0xffffffffu : ((1u << TG) - 1u);
        const float *col = reinterpret_cast<const float *>(sV) + tid;
        float a = 0.f;
        int j = 0;
        for (; j + 16 <= n; j += 16) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; u++) v[u] = col[(j + u) * TG];
#pragma unroll
            for (int u = 0; u < 16; u++) a = __fadd_rn(a, v[u]);
        }
        for (; j < n; j++) a = __fadd_rn(a, col[j * TG]);
        const uint32_t ob = pg_ord(a);
        const uint32_t gm = __reduce_max_sync(fmask, ob);
        const uint32_t gi = __reduce_min_sync(fmask, ob == gm ? (uint32_t)(gbase + tid) : 0xFFFFFFFFu);
        if (tid == 0) atomicMax(mybest, ((unsigned long long)gm << 32) | (unsigned long long)(0xFFFFFFFFu - gi));
        return;
    }

    // ---- A8: bootstrap replicates
    const int t2 = tid - 32;
    const int group = t2 / LPR, l = t2 % LPR;
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (((tid & 31) / LPR) * LPR));
    int k = n >> 3;
    if (k < min_boot) k = min_boot;
    const int nb = (k + 3) >> 2;                   // 4-draw batches per replicate
    const uint4 *lists = reinterpret_cast<const uint4 *>(boot_pool + boot_off[n]);
    const char *lane_base = reinterpret_cast<const char *>(sV) + l * 16;
    const uint32_t g0 = (uint32_t)(gbase + l * 4);
#define PG_ROW(o) (*reinterpret_cast<const float4 *>(lane_base + ((o) >> SH)))

    for (int task = group; task < PG_NUM_BOOT; task += NGR) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (nb > 0) {
            const uint4 *lp = lists + (size_t)(task / IL) * nb * IL + (task % IL);
            uint4 qa = __ldg(lp), qb = __ldg(lp + IL);
            float4 x0 = PG_ROW(qa.x), x1 = PG_ROW(qa.y), x2 = PG_ROW(qa.z), x3 = PG_ROW(qa.w);
            float4 y0, y1, y2, y3;
            int b = 0;
            while (b + 1 < nb) {
                y0 = PG_ROW(qb.x); y1 = PG_ROW(qb.y); y2 = PG_ROW(qb.z); y3 = PG_ROW(qb.w);
                qa = __ldg(lp + (b + 2) * IL);
                PG_ADD4(x0) PG_ADD4(x1) PG_ADD4(x2) PG_ADD4(x3)
                x0 = PG_ROW(qa.x); x1 = PG_ROW(qa.y); x2 = PG_ROW(qa.z); x3 = PG_ROW(qa.w);
                qb = __ldg(lp + (b + 3) * IL);
                PG_ADD4(y0) PG_ADD4(y1) PG_ADD4(y2) PG_ADD4(y3)
                b += 2;
            }
            if (b < nb) { PG_ADD4(x0) PG_ADD4(x1) PG_ADD4(x2) PG_ADD4(x3) }
        }
        // first strict max in ascending genus index: inside the lane, then across the group
        float bv = a0;
        uint32_t bi = g0;
        if (a1 > bv) { bv = a1; bi = g0 + 1; }
        if (a2 > bv) { bv = a2; bi = g0 + 2; }
        if (a3 > bv) { bv = a3; bi = g0 + 3; }
        const uint32_t ob = pg_ord(bv);
        const uint32_t gm = __reduce_max_sync(gmask, ob);
        const uint32_t gi = __reduce_min_sync(gmask, ob == gm ? bi : 0xFFFFFFFFu);
        if (l == 0)
            atomicMax(mybest + 1 + task, ((unsigned long long)gm << 32) | (unsigned long long)(0xFFFFFFFFu - gi));
    }
#undef PG_ROW
}

// ------------------------------------------------------------------ K5 vote (A9)

// One warp per read: decode the 101 winners, count for every lineage level of
// the determined genus the replicates whose winner shares that ancestor.
__global__ void __launch_bounds__(256)
k_vote(const unsigned long long *__restrict__ best, int64_t nreads, int64_t slot0,
       const int32_t *__restrict__ order, const int32_t *__restrict__ nwords, const uint8_t *__restrict__ flags,
       const int32_t *__restrict__ anc, int depth, pg_result *__restrict__ results,
       int32_t *__restrict__ boot_winners)
{
    const int lane = threadIdx.x & 31;
    const int64_t ic = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ic >= nreads) return;
    const int64_t i = (int64_t)order[ic];
    pg_result *res = results + i;
    uint32_t *raw = reinterpret_cast<uint32_t *>(res);          // 16 x uint32
    if (flags[2 * i + 1]) {                                     // A2: short read
        if (lane < 16) raw[lane] = (lane == 0) ? 0xFFFFFFFFu : (lane == 3 ? (1u << 8) : 0u);
        if (boot_winners)
            for (int r = lane; r < PG_NUM_BOOT; r += 32) boot_winners[i * PG_NUM_BOOT + r] = -1;
        return;
    }
    const unsigned long long *b = best + (size_t)(slot0 + ic) * (PG_NUM_BOOT + 1);
    const unsigned long long key0 = b[0];
    const int genus = (int)(0xFFFFFFFFu - (uint32_t)key0);
    const float score = pg_unord((uint32_t)(key0 >> 32));

    int gb[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int r = q * 32 + lane;
        gb[q] = r < PG_NUM_BOOT ? (int)(0xFFFFFFFFu - (uint32_t)b[1 + r]) : -1;
        if (boot_winners && r < PG_NUM_BOOT) boot_winners[i * PG_NUM_BOOT + r] = gb[q];
    }
    int myvote = 0, levels = 0;
    const int nd = anc ? depth : 1;
    for (int d = 0; d < nd; d++) {
        const int a = anc ? anc[(size_t)genus * depth + d] : genus;
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            bool match = false;
            if (gb[q] >= 0 && a >= 0) match = (anc ? anc[(size_t)gb[q] * depth + d] : gb[q]) == a;
            cnt += __popc(__ballot_sync(0xffffffffu, match));
        }
        if (a >= 0) levels = d + 1;
        if (lane == d) myvote = cnt;
    }
    // votes[32] live in raw[4..11]; pack four lanes per word
    uint32_t packed = (uint32_t)myvote & 0xFFu;
    uint32_t w0 = __shfl_sync(0xffffffffu, packed, (lane & 7) * 4 + 0);
    uint32_t w1 = __shfl_sync(0xffffffffu, packed, (lane & 7) * 4 + 1);
    uint32_t w2 = __shfl_sync(0xffffffffu, packed, (lane & 7) * 4 + 2);
    uint32_t w3 = __shfl_sync(0xffffffffu, packed, (lane & 7) * 4 + 3);
    const uint32_t vw = w0 | (w1 << 8) | (w2 << 16) | (w3 << 24);
    if (lane == 0) {
        raw[0] = (uint32_t)genus;
        raw[1] = (uint32_t)nwords[i];
        raw[2] = __float_as_uint(score);
        raw[3] = (uint32_t)flags[2 * i] | (0u << 8) | ((uint32_t)levels << 16);
    }
    if (lane < 8) raw[4 + lane] = vw;
    if (lane >= 8 && lane < 12) raw[4 + lane] = 0u;             // _pad1
}

// ------------------------------------------------------------------ reads of a range in order of word count
//
// The launches of a range are sized per bucket of word counts, and reads of equal length should sit next to each other
// (they share their sample lists).  The order array is a counting sort by n, done on the device: a histogram of the
// range's word counts goes to the host (28 KB, fetched one range ahead -- the host needs it anyway to size the
// launches and to know which sample lists to build), the start of every n comes back, and one pass scatters the read
// indices.  The host never touches per-read data in the steady state.
#define PG_HB (PG_MAX_WORDS + 2)                     // bins 0 .. PG_MAX_WORDS, and one for longer reads
static __global__ void k_words_hist(const int32_t *__restrict__ nwords, int64_t cnt, int32_t *__restrict__ hist)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const int n = nwords[i];
    atomicAdd(hist + (n > PG_MAX_WORDS ? PG_MAX_WORDS + 1 : n), 1);
}
static __global__ void k_words_scatter(const int32_t *__restrict__ nwords, int64_t cnt, int64_t first, const int32_t *__restrict__ start,
                                       int32_t *__restrict__ cursor, int32_t *__restrict__ order)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const int n = nwords[i];
    order[start[n] + atomicAdd(cursor + n, 1)] = (int32_t)(first + i);
}

// ------------------------------------------------------------------ host orchestration

// block = 32 (full-sum warp) + groups*LPR: 20 groups x 5 replicates, 52 x 2, 100 x 1.
static const Bucket kBuckets[] = {
    {215, 8, 192},  {250, 8, 192},  {290, 8, 192},  {350, 8, 192},  {440, 8, 192},
    {590, 8, 192},  {640, 8, 448},  {900, 8, 448},  {1800, 8, 832}, {3600, 4, 832}, {PG_MAX_WORDS, 2, 832},   // 640: the longest read of plan 4
};
static const int kNumBuckets = (int)(sizeof(kBuckets) / sizeof(kBuckets[0]));
static const Bucket &bucket_of(int n)
{
    int b = 0;
    while (kBuckets[b].nmax < n) b++;
    return kBuckets[b];
}

template <int LPR, int BLOCK>
static int launch_strict(pg_ctx *ctx, const pg_model *md, unsigned nreads_b, unsigned nblocks_g, size_t smem,
                         const uint16_t *d_words, const int64_t *d_off, const int32_t *d_nwords,
                         const int32_t *d_order, int64_t slot0, int min_boot, unsigned long long *d_best)
{
    // per device, so set it on every launch (a few microseconds)
    PG_CUDA(ctx, pg_smem_unlock(ctx, k_classify_strict<LPR, BLOCK>));
    dim3 grid(nreads_b, nblocks_g);
    k_classify_strict<LPR, BLOCK><<<grid, BLOCK, smem, ctx->stream>>>(
        md->d_table, d_words, d_off, d_nwords, d_order, slot0, ctx->d_boot_pool, ctx->d_boot_off, min_boot, d_best);
    PG_LAUNCHED(ctx);
    return PG_OK;
}

int pg_certified_phase1(pg_ctx *ctx, const pg_model *md, const Bucket &bk, unsigned nreads_b, int nmax,
                        const uint16_t *d_words, const int64_t *d_off, const int32_t *d_nwords,
                        const uint8_t *d_flags, const int32_t *d_order, int64_t slot0, int min_boot,
                        const PgCertBufs &cb, int version);
int pg_certified_phase2(pg_ctx *ctx, const pg_model *md, unsigned nreads_b, int nmax, const uint16_t *d_words,
                        const int64_t *d_off, const int32_t *d_nwords, const uint8_t *d_flags,
                        const int32_t *d_order, int64_t slot0, int min_boot, const PgCertBufs &cb, bool use_heavy,
                        pg_result *d_results, int32_t *d_boot_winners);
int pg_mma_ensure_images(pg_ctx *ctx, const std::vector<int> &need_n, int min_boot);          // pg_mma.cu
bool pg_mma_usable(const pg_model *md, int nmax);
#define PG_PIPE_SLICE 16384                         // reads per slice of a pipelined plan-4 bucket
int pg_certified_reset(pg_ctx *ctx, const int32_t *d_list, int cnt, int64_t slot0, const PgCertBufs &cb);   // pg_certified.cu
#define PG_CANDCAP 128

static int ensure_boot_lists(pg_ctx *ctx, const std::vector<int> &need_n, int min_boot)
{
    if (ctx->boot_min_words != min_boot) {          // different k rule: drop the cache
        ctx->boot_used = 0;
        ctx->h_boot_off.assign(PG_MAX_WORDS + 1, -1);
        ctx->boot_min_words = min_boot;
    }
    if (!ctx->d_boot_off) {
        PG_CUDA(ctx, cudaMalloc(&ctx->d_boot_off, (PG_MAX_WORDS + 1) * sizeof(int32_t)));
        ctx->h_boot_off.assign(PG_MAX_WORDS + 1, -1);
    }
    std::vector<int32_t> ns, offs, ils;
    size_t used = ctx->boot_used;
    for (int n : need_n) {
        if (ctx->h_boot_off[n] >= 0) continue;
        int k = n >> 3;
        if (k < min_boot) k = min_boot;
        ns.push_back(n);
        offs.push_back((int32_t)used);
        ils.push_back(32 / bucket_of(n).lpr);
        ctx->h_boot_off[n] = (int32_t)used;
        used += (pg_boot_list_entries(k, ils.back()) + 31) & ~(size_t)31;   // keep every list 128-byte aligned
    }
    if (ns.empty()) return PG_OK;
    if (used > ctx->boot_cap) {
        size_t cap = used * 2 + 4096;
        uint32_t *np = NULL;
        PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PG_CUDA(ctx, cudaMalloc(&np, cap * sizeof(uint32_t)));
        if (ctx->d_boot_pool) {
            PG_CUDA(ctx, pg_copy_sync(ctx, np, ctx->d_boot_pool, ctx->boot_used * sizeof(uint32_t), cudaMemcpyDeviceToDevice));
            PG_CUDA(ctx, cudaFree(ctx->d_boot_pool));
        }
        ctx->d_boot_pool = np;
        ctx->boot_cap = cap;
    }
    ctx->boot_used = used;
    int cnt = (int)ns.size();
    PG_TRY(pg_scratch(ctx, &ctx->s_boot, (size_t)cnt * 12));
    int32_t *d_ns = (int32_t *)ctx->s_boot.p, *d_offs = d_ns + cnt, *d_ils = d_offs + cnt;
    PG_CUDA(ctx, cudaMemcpyAsync(d_ils, ils.data(), (size_t)cnt * 4, cudaMemcpyHostToDevice, ctx->stream));
    // pageable source: cudaMemcpyAsync stages it before returning
    PG_CUDA(ctx, cudaMemcpyAsync(d_ns, ns.data(), (size_t)cnt * 4, cudaMemcpyHostToDevice, ctx->stream));
    PG_CUDA(ctx, cudaMemcpyAsync(d_offs, offs.data(), (size_t)cnt * 4, cudaMemcpyHostToDevice, ctx->stream));
    PG_CUDA(ctx, cudaMemcpyAsync(ctx->d_boot_off, ctx->h_boot_off.data(), (PG_MAX_WORDS + 1) * sizeof(int32_t),
                                 cudaMemcpyHostToDevice, ctx->stream));
    k_boot_indices<<<(cnt + 31) / 32, 32, 0, ctx->stream>>>(d_ns, d_offs, d_ils, cnt, min_boot, ctx->d_boot_pool);
    PG_LAUNCHED(ctx);
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // ns/offs are stack vectors
    return PG_OK;
}

int pg_extract_launch(pg_ctx *ctx, const pg_model *md, const uint32_t *d_planes, const int64_t *d_off,
                      int64_t count, uint16_t *d_words, int32_t *d_nwords, uint8_t *d_flags);
int pg_pack_launch(pg_ctx *ctx, const char *d_bytes, const int64_t *d_off, int64_t count, uint32_t *d_planes);
int pg_extract_ascii_launch(pg_ctx *ctx, const pg_model *md, const char *d_bytes, const int64_t *d_off,
                            int64_t count, uint16_t *d_words, int32_t *d_nwords, uint8_t *d_flags);

#include <time.h>
static double pg_now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

static cudaEvent_t take_event(pg_ctx *ctx)
{
    cudaEvent_t ev;
    if (!ctx->ev_free.empty()) { ev = ctx->ev_free.back(); ctx->ev_free.pop_back(); return ev; }
    cudaEventCreate(&ev);
    return ev;
}

// ------------------------------------------------------------------ host orchestration
//
// A batch is classified range by range (a range = one chunk of reads).  Nothing in the steady state
// waits for the device except the 4-byte-per-read word counts the host needs to bucket a range, and
// those are fetched one range ahead, so the host enqueues range i while the device still works on
// range i-1.  Fallback lists (heavy reads, overflowing near-tie lists) grow on the device across the
// whole batch and are dealt with once, in finish().
struct ClassifyJob {
    pg_ctx *ctx;
    const pg_model *md;
    const int64_t *d_off;
    int64_t count;
    const uint16_t *d_words;
    const int32_t *d_nwords;
    const uint8_t *d_flags;
    pg_result *d_results;
    int32_t *d_boot_winners;
    int min_boot, mode, cert_version, nkeys;
    bool certified;
    int64_t CHUNK, cmax;
    unsigned long long *d_best;
    int32_t *d_order;
    PgCertBufs cb;
    int32_t *h_n, *h_order;
    int64_t *h_off;
    int32_t *h_hist;                             // [2][PG_HB] pinned: word-count histograms of the range in flight and the next
    int32_t *d_hist, *d_start, *d_cursor;        // [2][PG_HB], [PG_HB], [PG_HB]
    int64_t nfetch, nrange;
    std::vector<int32_t> hstart;
    std::vector<char> seen;
    std::vector<int32_t> redone;                 // reads whose records were rewritten by finish()
    int64_t bcount[16], bstart[16], bmaxn[16];

    int begin(pg_ctx *c, const pg_model *m, const int64_t *off, int64_t cnt, const uint16_t *words, const int32_t *nwords,
              const uint8_t *flags, const pg_classify_opts *opts, pg_result *results, int32_t *boot_winners)
    {
        ctx = c; md = m; d_off = off; count = cnt; d_words = words; d_nwords = nwords; d_flags = flags;
        d_results = results; d_boot_winners = boot_winners;
        min_boot = opts ? opts->min_boot_words : 0;
        mode = opts ? opts->mode : 0;
        nkeys = PG_NUM_BOOT + 1;
        if (min_boot < 0 || min_boot > 64) return pg_fail(ctx, PG_EINVAL, "min_boot_words out of range");
        if (mode != 0 && mode != 1) return pg_fail(ctx, PG_EINVAL, "unknown classify mode %d", mode);
        if (opts && (opts->cert_plan < 0 || opts->cert_plan > 3)) return pg_fail(ctx, PG_EINVAL, "unknown cert_plan %d", opts->cert_plan);
        if (count > 0x7fffffffLL) return pg_fail(ctx, PG_ERANGE, "more than 2^31-1 reads in one batch");
        certified = (mode == 1) && md->q_ok;
        static int env_v1 = -1;                         // PG_CERT_V1=1: the all-block kernel for every read (A/B switch)
        if (env_v1 < 0) { const char *e = getenv("PG_CERT_V1"); env_v1 = (e && atoi(e)) ? 1 : 0; }
        // 4 = plan 4, the default: plan 3's best part and bounds on the tensor cores (pg_mma.cu); an explicit
        // bound_level or cert_plan 3 asks for plan 3's own kernels
        { static int env_pipe = -2; if (env_pipe == -2) { const char *e = getenv("PG_PIPE"); env_pipe = e ? atoi(e) : 0; } pipe_ok = env_pipe != 0; }
        cert_version = (env_v1 || (opts && opts->cert_plan == 1)) ? 1 : ((opts && opts->cert_plan == 2) ? 2 :
                       ((opts && (opts->cert_plan == 3 || opts->bound_level != 0)) ? 3 : 4));
        ctx->st_certified = ctx->st_strict = ctx->st_handed_back = 0;
        ctx->st_heavy = ctx->st_items = 0;
        ctx->st_mma = 0;
        // Chunk of reads per pass.  Plan 1 walks every genus block of a chunk (tile-major grid) and wants the
        // chunk's word ids, champion slots and near-tie lists L2-resident across those passes (2^14 measured
        // best); plan 2 touches a read's data once per kernel and prefers fewer, larger launches (2^16); plan 4's
        // persistent kernel has a ramp and a tail per launch (2^16 / 2^17 / 2^18: 26.3 / 27.5 / 28.3 M reads/s).
        static int64_t chunk_override = -1;
        if (chunk_override < 0) { const char *e = getenv("PG_CHUNK_LOG2"); chunk_override = e ? atoi(e) : 0; }
        CHUNK = !certified ? ((int64_t)1 << 20) : ((int64_t)1 << (chunk_override ? chunk_override : (cert_version == 1 ? 14 : (cert_version == 4 ? 18 : 16))));
        cmax = count < CHUNK ? count : CHUNK;
        if (cmax < 1) cmax = 1;
        PG_TRY(pg_pinned(ctx, (size_t)count * 16 + 128 + 2 * PG_HB * 4));
        h_n = (int32_t *)ctx->h_pin;
        h_order = h_n + count;
        h_off = (int64_t *)(h_order + count + 2);        // rebased read offsets of pg_classify()'s uploads
        h_hist = (int32_t *)(h_off + count + 2);
        PG_TRY(pg_scratch(ctx, &ctx->s_hist, (size_t)4 * PG_HB * 4));
        d_hist = (int32_t *)ctx->s_hist.p;
        d_start = d_hist + 2 * PG_HB;
        d_cursor = d_start + PG_HB;
        nfetch = nrange = 0;
        hstart.assign(PG_HB, 0);
        seen.assign(PG_MAX_WORDS + 1, 0);
        redone.clear();
        PG_TRY(pg_scratch(ctx, &ctx->s_best, (size_t)cmax * nkeys * 8));
        PG_TRY(pg_scratch(ctx, &ctx->s_order, (size_t)cmax * 4));
        d_best = (unsigned long long *)ctx->s_best.p;
        d_order = (int32_t *)ctx->s_order.p;
        memset(&cb, 0, sizeof cb);
        cb.light_max = opts ? opts->light_max : 0;
        cb.bound_level = opts ? opts->bound_level : 0;
        cb.force_part = -1;
        if (cb.bound_level < 0 || cb.bound_level > 3) return pg_fail(ctx, PG_EINVAL, "unknown bound_level %d", cb.bound_level);
        if (certified) {
            PG_TRY(pg_scratch(ctx, &ctx->s_champ, (size_t)cmax * nkeys * 8));
            PG_TRY(pg_scratch(ctx, &ctx->s_ncand, (size_t)cmax * 4));
            PG_TRY(pg_scratch(ctx, &ctx->s_candl, (size_t)cmax * PG_CANDCAP * 8));
            PG_TRY(pg_scratch(ctx, &ctx->s_fb, (size_t)count * 4 + 32));
            PG_TRY(pg_scratch(ctx, &ctx->s_guess, (size_t)cmax * 4));
            // room for 512 open (task, block) pairs per read of a chunk on average (268 MB for 2^16 reads): a full buffer
            // sends reads to the all-block kernel, which costs far more than the items would (models with hundreds of
            // members per genus leave ~200 pairs per read open)
            cb.item_cap = (unsigned int)(cmax * 512 < 4096 ? 4096 : cmax * 512);
            PG_TRY(pg_scratch(ctx, &ctx->s_items, (size_t)cb.item_cap * 8));
            PG_TRY(pg_scratch(ctx, &ctx->s_heavy, (size_t)count * 4 + (size_t)cmax + 64));
            cb.champ = (unsigned long long *)ctx->s_champ.p;
            cb.ncand = (unsigned int *)ctx->s_ncand.p;
            cb.cand = (unsigned long long *)ctx->s_candl.p;
            cb.counters = (unsigned int *)ctx->s_fb.p;           // [4], [5]: item cursor of the second half of a pipelined bucket, spare
            cb.item_count = cb.counters + 2;
            cb.fb_list = (int32_t *)ctx->s_fb.p + 8;
            cb.guess = (int32_t *)ctx->s_guess.p;
            cb.items = (unsigned long long *)ctx->s_items.p;
            cb.hv_list = (int32_t *)ctx->s_heavy.p;
            cb.heavy = (uint8_t *)((int32_t *)ctx->s_heavy.p + count);
            PG_CUDA(ctx, cudaMemsetAsync(cb.counters, 0, 32, ctx->stream));
        }
        return PG_OK;
    }

    // word counts of reads [r0, r1) to the host, asynchronously; *ev fires when they have landed
    int fetch_counts(int64_t r0, int64_t r1, cudaEvent_t *ev)
    {
        int32_t *dh = d_hist + (nfetch & 1) * PG_HB, *hh = h_hist + (nfetch & 1) * PG_HB;
        nfetch++;
        PG_CUDA(ctx, cudaMemsetAsync(dh, 0, PG_HB * 4, ctx->stream));
        k_words_hist<<<(unsigned)((r1 - r0 + 255) / 256), 256, 0, ctx->stream>>>(d_nwords + r0, r1 - r0, dh);
        PG_LAUNCHED(ctx);
        PG_CUDA(ctx, cudaMemcpyAsync(hh, dh, PG_HB * 4, cudaMemcpyDeviceToHost, ctx->stream));
        // the per-read counts follow for the fallback passes of finish(); nobody waits for them in the steady state
        PG_CUDA(ctx, cudaMemcpyAsync(h_n + r0, d_nwords + r0, (size_t)(r1 - r0) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        *ev = take_event(ctx);
        PG_CUDA(ctx, cudaEventRecord(*ev, ctx->stream));
        return PG_OK;
    }

    // stable counting sort of a list of reads by word count -> h_dst; buckets are ranges of n, so they come out
    // contiguous (extents in bcount/bstart/bmaxn), and reads of equal length end up adjacent -- plan 3 pairs
    // neighbours in one CTA and equal lengths share their sample lists
    std::vector<int64_t> ncount;
    void bucket_sort(const int32_t *src, int64_t base, int64_t cn, int32_t *h_dst)
    {
        ncount.assign(PG_MAX_WORDS + 2, 0);
        for (int64_t i = 0; i < cn; i++) ncount[(size_t)h_n[src ? src[i] : (int32_t)(base + i)] + 1]++;
        for (int b = 0; b < 16; b++) bcount[b] = bstart[b] = bmaxn[b] = 0;
        {
            int b = 0;
            int64_t acc = 0;
            for (int n = 0; n <= PG_MAX_WORDS; n++) {
                const int64_t c = ncount[(size_t)n + 1];
                ncount[(size_t)n + 1] = 0;
                if (c) {
                    while (kBuckets[b].nmax < n) b++;
                    if (!bcount[b]) bstart[b] = acc;
                    bcount[b] += c;
                    bmaxn[b] = n;
                }
                ncount[(size_t)n] = acc;                // start of the reads with n words
                acc += c;
            }
        }
        for (int64_t i = 0; i < cn; i++) {
            const int32_t r = src ? src[i] : (int32_t)(base + i);
            h_dst[ncount[(size_t)h_n[r]]++] = r;
        }
    }

    // one bucket of reads through the strict kernels (also the certified path's last resort)
    int run_strict(const Bucket &bk, const int32_t *ord, int64_t slot0, unsigned cnt, int nmax, bool timed)
    {
        const int wpb = 8;
        cudaEvent_t e0 = NULL, e1 = NULL;
        if (timed) {
            e0 = take_event(ctx); e1 = take_event(ctx);
            PG_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        }
        const int TG = 4 * bk.lpr;
        const unsigned ngb = (unsigned)((md->G + TG - 1) / TG);
        const size_t smem = (size_t)(nmax + 1) * bk.lpr * 16;
        int rc;
        if (bk.lpr == 8 && bk.block == 192)
            rc = launch_strict<8, 192>(ctx, md, cnt, ngb, smem, d_words, d_off, d_nwords, ord, slot0, min_boot, d_best);
        else if (bk.lpr == 8 && bk.block == 448)
            rc = launch_strict<8, 448>(ctx, md, cnt, ngb, smem, d_words, d_off, d_nwords, ord, slot0, min_boot, d_best);
        else if (bk.lpr == 8)
            rc = launch_strict<8, 832>(ctx, md, cnt, ngb, smem, d_words, d_off, d_nwords, ord, slot0, min_boot, d_best);
        else if (bk.lpr == 4)
            rc = launch_strict<4, 832>(ctx, md, cnt, ngb, smem, d_words, d_off, d_nwords, ord, slot0, min_boot, d_best);
        else
            rc = launch_strict<2, 832>(ctx, md, cnt, ngb, smem, d_words, d_off, d_nwords, ord, slot0, min_boot, d_best);
        PG_TRY(rc);
        if (timed) {
            PG_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
            ctx->ev_pending.push_back(std::make_pair(e0, e1));
        }
        k_vote<<<(cnt + wpb - 1) / wpb, wpb * 32, 0, ctx->stream>>>(d_best, cnt, slot0, ord, d_nwords, d_flags, md->d_anc,
                                                                  md->depth, d_results, d_boot_winners);
        PG_LAUNCHED(ctx);
        return PG_OK;
    }

    // second stream and events of the pipelined buckets (run_pass)
    bool pipe_ok = false;                            // PG_PIPE=1: two streams (A/B switch; measured slower, see run_pass)
    int pipe_setup()
    {
        if (ctx->aux_stream) return PG_OK;
        PG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 4; i++) PG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_pipe[i], cudaEventDisableTiming));
        return PG_OK;
    }

    // one pass over a set of reads (a chunk of the batch, or a list): plan 2 / plan 1 / strict (plan 0).
    // A "slot" is a position in the order array of the pass; the per-read scratch (champion slots,
    // near-tie lists, strict keys, guesses) is indexed by slot.
    int run_pass(const int32_t *h_list, int64_t cn, int plan, bool timed)
    {
        if (h_list)                                         // NULL: the order array was built on the device (k_words_scatter)
            PG_CUDA(ctx, cudaMemcpyAsync(d_order, h_list, (size_t)cn * 4, cudaMemcpyHostToDevice, ctx->stream));
        bool need_best = plan == 0 || !certified;           // strict keys: also for buckets the certified kernels do not take
        for (int b = 0; b < kNumBuckets; b++)
            if (bcount[b] && kBuckets[b].lpr != 8) need_best = true;
        if (need_best) PG_CUDA(ctx, cudaMemsetAsync(d_best, 0, (size_t)cn * nkeys * 8, ctx->stream));
        if (plan != 0 && certified) {
            PG_CUDA(ctx, cudaMemsetAsync(cb.champ, 0xFF, (size_t)cn * nkeys * 8, ctx->stream));
            PG_CUDA(ctx, cudaMemsetAsync(cb.ncand, 0, (size_t)cn * 4, ctx->stream));
            PG_CUDA(ctx, cudaMemsetAsync(cb.heavy, 0, (size_t)cn, ctx->stream));
        }
        for (int b = 0; b < kNumBuckets; b++) {
            if (!bcount[b]) continue;
            const Bucket &bk = kBuckets[b];
            const int nmax = (int)bmaxn[b];
            const int32_t *ord = d_order + bstart[b];
            if (certified && plan != 0 && bk.lpr == 8) {
                if (timed) ctx->st_certified += bcount[b];
                cb.count_mma = timed;
                cudaEvent_t e0 = NULL, e1 = NULL;
                if (timed) {
                    e0 = take_event(ctx); e1 = take_event(ctx);
                    PG_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
                }
                if (plan == 4 && cb.bound_level == 0 && cb.force_part < 0 && bcount[b] >= 2 * PG_PIPE_SLICE && pipe_ok && pg_mma_usable(md, nmax)) {
                    // Plan 4, a large bucket: slices of PG_PIPE_SLICE reads; the item kernel and phase 2 of slice j run on
                    // a second stream UNDER the tensor-core kernel of slice j + 1 -- that kernel is one latency-bound CTA
                    // per SM (issue slots 40 % busy), the others fit beside it.  Two halves of the item list alternate.
                    // MEASURED (PG_PIPE=1): same records, 24.0 M reads/s against 26.1 M on one stream -- the kernels
                    // compete for the same thing, L2 requests in flight per SM, and 16 384-read launches add tails.  Off by default.
                    PG_TRY(pipe_setup());
                    cudaStream_t main_stream = ctx->stream;
                    const unsigned int half = cb.item_cap / 2;
                    int64_t j = 0;
                    for (int64_t o = 0; o < bcount[b]; o += PG_PIPE_SLICE, j++) {
                        const unsigned c = (unsigned)(bcount[b] - o < PG_PIPE_SLICE ? bcount[b] - o : PG_PIPE_SLICE);
                        const int p = (int)(j & 1);
                        PgCertBufs cs = cb;
                        cs.items = cb.items + (size_t)p * half;
                        cs.item_cap = half;
                        cs.item_count = cb.counters + (p ? 4 : 2);
                        if (j >= 2) PG_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_pipe[2 + p], 0));   // that half of the list is consumed
                        cs.stage = 1;
                        PG_TRY(pg_certified_phase1(ctx, md, bk, c, nmax, d_words, d_off, d_nwords, d_flags, ord + o, bstart[b] + o, min_boot, cs, plan));
                        PG_CUDA(ctx, cudaEventRecord(ctx->ev_pipe[p], main_stream));
                        PG_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_pipe[p], 0));
                        ctx->stream = ctx->aux_stream;
                        cs.stage = 2;
                        int rc2 = pg_certified_phase1(ctx, md, bk, c, nmax, d_words, d_off, d_nwords, d_flags, ord + o, bstart[b] + o, min_boot, cs, plan);
                        if (rc2 == PG_OK)
                            rc2 = pg_certified_phase2(ctx, md, c, nmax, d_words, d_off, d_nwords, d_flags, ord + o, bstart[b] + o, min_boot, cs, true,
                                                      d_results, d_boot_winners);
                        cudaError_t ee = cudaEventRecord(ctx->ev_pipe[2 + p], ctx->aux_stream);
                        ctx->stream = main_stream;
                        PG_TRY(rc2);
                        PG_CUDA(ctx, ee);
                    }
                    PG_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_pipe[2], 0));
                    if (j >= 2) PG_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_pipe[3], 0));
                    if (timed) {
                        PG_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
                        ctx->ev_pending.push_back(std::make_pair(e0, e1));
                    }
                    continue;
                }
                PG_TRY(pg_certified_phase1(ctx, md, bk, (unsigned)bcount[b], nmax, d_words, d_off, d_nwords, d_flags, ord,
                                           bstart[b], min_boot, cb, plan));
                if (timed) {
                    PG_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
                    ctx->ev_pending.push_back(std::make_pair(e0, e1));
                }
                PG_TRY(pg_certified_phase2(ctx, md, (unsigned)bcount[b], nmax, d_words, d_off, d_nwords, d_flags, ord,
                                           bstart[b], min_boot, cb, plan >= 2, d_results, d_boot_winners));
            } else {
                if (timed) ctx->st_strict += bcount[b];
                PG_TRY(run_strict(bk, ord, bstart[b], (unsigned)bcount[b], nmax, timed));
            }
        }
        return PG_OK;
    }

    // Block or part columns for the lower bounds?  Which is faster depends on the MODEL (with a few training members
    // per genus a 64-genus block minimum dismisses far blocks; with hundreds, once-seen words fill every genus and only
    // the 16-position part minima do), so a model with more than one group of blocks is tried both ways, once, on the
    // head of the first batch it classifies.  Results never depend on the choice; the trial's records are rewritten
    // by the real pass.
    int tune_bounds(int64_t r0, int64_t r1)
    {
        if (md->bounds_tuned || !certified || cert_version != 3 || cb.bound_level != 0 || md->ngroup < 2) return PG_OK;
        if (r0 != 0 || r1 - r0 < 2048) return PG_OK;                 // counters must still be zero; too few reads say nothing
        // the range's order array and buckets are in place (range()): the trial is the first range of the batch
        const int64_t cn = r1 - r0;
        cudaEvent_t ev[3];
        for (int i = 0; i < 3; i++) ev[i] = take_event(ctx);
        float ms[2] = {0.f, 0.f};
        unsigned int hv[2] = {0u, 0u};
        for (int rep = 0; rep < 2; rep++) {                          // the first round warms the sample lists and the caches
            PG_CUDA(ctx, cudaMemsetAsync(cb.counters, 0, 16, ctx->stream));
            for (int mode = 0; mode < 2; mode++) {
                cb.force_part = mode;
                PG_CUDA(ctx, cudaEventRecord(ev[mode], ctx->stream));
                PG_TRY(run_pass(NULL, cn, 3, false));
                PG_CUDA(ctx, cudaEventRecord(ev[mode + 1], ctx->stream));
                PG_CUDA(ctx, cudaMemcpyAsync(&hv[mode], cb.counters + 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
            }
            PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            PG_CUDA(ctx, cudaEventElapsedTime(&ms[0], ev[0], ev[1]));
            PG_CUDA(ctx, cudaEventElapsedTime(&ms[1], ev[1], ev[2]));
        }
        // reads the bounds leave too many open pairs for are redone later by the all-block kernel (finish()); its cost
        // is not in the timed passes: about 22 ns per read and block (2.05 M reads/s on 20 blocks, 0.28 M on 184)
        const double redo_ms = 22e-6 * (double)md->ntile64;
        const double cost0 = ms[0] + redo_ms * (double)hv[0], cost1 = ms[1] + redo_ms * (double)(hv[1] - hv[0]);
        for (int i = 0; i < 3; i++) ctx->ev_free.push_back(ev[i]);
        cb.force_part = -1;
        md->part_bounds = cost1 < cost0;
        if (getenv("PG_TIMING"))
            fprintf(stderr, "[pg_classify] bound columns of this model: blocks %.3f ms + %u reads to redo, parts %.3f ms + %u on %lld reads -> %s\n",
                    ms[0], hv[0], ms[1], hv[1] - hv[0], (long long)cn, md->part_bounds ? "parts" : "blocks");
        md->bounds_tuned = true;
        PG_CUDA(ctx, cudaMemsetAsync(cb.counters, 0, 16, ctx->stream));      // the trial's fallback lists are void
        return PG_OK;
    }

    // enqueue the classification of reads [r0, r1); their word counts must have landed (fetch_counts)
    int range(int64_t r0, int64_t r1, cudaEvent_t counts_ready)
    {
        PG_CUDA(ctx, cudaEventSynchronize(counts_ready));
        ctx->ev_free.push_back(counts_ready);
        const int32_t *hist = h_hist + (nrange & 1) * PG_HB;
        nrange++;
        const int64_t cn = r1 - r0;
        if (cn > CHUNK) return pg_fail(ctx, PG_EINVAL, "internal: a range of %lld reads exceeds the chunk", (long long)cn);
        if (hist[PG_MAX_WORDS + 1] > 0)
            return pg_fail(ctx, PG_ERANGE, "%d read(s) of the batch have more than %d good words, the limit of this build", hist[PG_MAX_WORDS + 1],
                           PG_MAX_WORDS);
        // from the histogram: the sample lists still to build, the start of every n in the order array, the buckets
        std::vector<int> need;
        for (int b = 0; b < 16; b++) bcount[b] = bstart[b] = bmaxn[b] = 0;
        {
            int b = 0;
            int32_t acc = 0;
            for (int n = 0; n <= PG_MAX_WORDS; n++) {
                const int32_t c = hist[n];
                hstart[(size_t)n] = acc;
                if (!c) continue;
                if (!seen[n]) { seen[n] = 1; need.push_back(n); }
                while (kBuckets[b].nmax < n) b++;
                if (!bcount[b]) bstart[b] = acc;
                bcount[b] += c;
                bmaxn[b] = n;
                acc += c;
            }
        }
        if (!need.empty()) PG_TRY(ensure_boot_lists(ctx, need, min_boot));
        if (!need.empty() && certified && cert_version == 4) PG_TRY(pg_mma_ensure_images(ctx, need, min_boot));
        // pageable source: staged before the call returns, so hstart may be rewritten for the next range
        PG_CUDA(ctx, cudaMemcpyAsync(d_start, hstart.data(), PG_HB * 4, cudaMemcpyHostToDevice, ctx->stream));
        PG_CUDA(ctx, cudaMemsetAsync(d_cursor, 0, PG_HB * 4, ctx->stream));
        k_words_scatter<<<(unsigned)((cn + 255) / 256), 256, 0, ctx->stream>>>(d_nwords + r0, cn, r0, d_start, d_cursor, d_order);
        PG_LAUNCHED(ctx);
        PG_TRY(tune_bounds(r0, r1));
        return run_pass(NULL, cn, certified ? cert_version : 0, true);
    }

    // Deferred work, read once for the whole batch:
    //   heavy reads (plan 2 left too many (task, block) pairs open) -> plan 1, the all-block kernel;
    //   reads whose near-tie list overflowed (in either plan)        -> the strict kernels.
    // Leaves the stream synchronised; `redone` lists the reads whose records were rewritten here.
    int finish()
    {
        if (!certified) {
            PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            return PG_OK;
        }
        unsigned int cnts[4] = {0, 0, 0, 0};
        PG_CUDA(ctx, cudaMemcpyAsync(cnts, cb.counters, 16, cudaMemcpyDeviceToHost, ctx->stream));
        PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->st_items = cnts[3];
        ctx->st_heavy = cnts[1];
        std::vector<int32_t> lst, sorted;
        // Heavy reads of plan 4 get a second try first: nearly all of them are heavy because the guess (32 sampled words)
        // picked the wrong part, so their champions are poor and every block stays open.  A more careful guess (64 words,
        // exact minima of every part) and the same kernels again cost a few dozen ns per such read; the all-block kernel
        // costs ~22 ns per read AND BLOCK (4 us on a 10 000-genus model).
        unsigned int first_heavy = 0;
        // (when more than a twentieth of the batch is heavy the guess is not the problem -- random or low-complexity
        // reads -- and the second try would only add its own cost)
        if (cnts[1] > 0 && cert_version == 4 && (int64_t)cnts[1] * 20 <= count) {
            first_heavy = cnts[1];
            lst.resize(cnts[1]);
            sorted.resize(cnts[1]);
            PG_CUDA(ctx, cudaMemcpyAsync(lst.data(), cb.hv_list, lst.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
            PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            std::sort(lst.begin(), lst.end());
            redone.insert(redone.end(), lst.begin(), lst.end());
            cb.retry = true;
            for (size_t p0 = 0; p0 < lst.size(); p0 += (size_t)CHUNK) {
                const int64_t cn = (int64_t)(lst.size() - p0 < (size_t)CHUNK ? lst.size() - p0 : (size_t)CHUNK);
                bucket_sort(lst.data() + p0, 0, cn, sorted.data() + p0);
                PG_TRY(run_pass(sorted.data() + p0, cn, 4, false));
            }
            cb.retry = false;
            PG_CUDA(ctx, cudaMemcpyAsync(cnts, cb.counters, 16, cudaMemcpyDeviceToHost, ctx->stream));
            PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // `sorted` is read by the copies above
            ctx->st_heavy = cnts[1] - first_heavy;                 // what the second try left heavy as well
        }
        if (cnts[1] > first_heavy) {
            lst.resize(cnts[1] - first_heavy);
            sorted.resize(lst.size());
            PG_CUDA(ctx, cudaMemcpyAsync(lst.data(), cb.hv_list + first_heavy, lst.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
            PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            std::sort(lst.begin(), lst.end());                   // device order is arbitrary; keep runs reproducible
            if (!first_heavy) redone.insert(redone.end(), lst.begin(), lst.end());
            const int64_t step = cert_version == 1 ? CHUNK : (int64_t)1 << 14;
            for (size_t p0 = 0; p0 < lst.size(); p0 += (size_t)step) {
                const int64_t cn = (int64_t)(lst.size() - p0 < (size_t)step ? lst.size() - p0 : (size_t)step);
                bucket_sort(lst.data() + p0, 0, cn, sorted.data() + p0);
                PG_TRY(run_pass(sorted.data() + p0, cn, 1, false));
            }
            PG_CUDA(ctx, cudaMemcpyAsync(cnts, cb.counters, 4, cudaMemcpyDeviceToHost, ctx->stream));
            PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // `sorted` is read by the copies above
        }
        if (cnts[0] > 0) {
            ctx->st_handed_back = cnts[0];
            lst.resize(cnts[0]);
            sorted.resize(cnts[0]);
            PG_CUDA(ctx, cudaMemcpyAsync(lst.data(), cb.fb_list, lst.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
            PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            std::sort(lst.begin(), lst.end());
            redone.insert(redone.end(), lst.begin(), lst.end());
            for (size_t p0 = 0; p0 < lst.size(); p0 += (size_t)CHUNK) {
                const int64_t cn = (int64_t)(lst.size() - p0 < (size_t)CHUNK ? lst.size() - p0 : (size_t)CHUNK);
                bucket_sort(lst.data() + p0, 0, cn, sorted.data() + p0);
                PG_TRY(run_pass(sorted.data() + p0, cn, 0, false));
            }
            PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // `sorted` is a stack vector
        }
        return PG_OK;
    }

    // ranges of the batch: a short first one (its preparation cannot overlap anything), then chunks.  With
    // uploads in the picture (pg_classify) the ranges double from 8 192 reads up to the chunk size: the upload of
    // range i+1 (twice the bytes) then fits under the kernels of range i, and the first kernels wait for 17 MB
    // of text instead of 56 MB.
    int64_t next_range(int64_t r0, bool uploads = false) const
    {
        int64_t r1;
        if (uploads) {
            int64_t len = 8192, a = 0;                  // 8 192, 16 384, ... , CHUNK, CHUNK, ...
            while (a + len <= r0 && len < CHUNK) { a += len; len *= 2; }
            if (len > CHUNK) len = CHUNK;
            r1 = r0 + len;
        } else {
            const int64_t first = CHUNK < 16384 ? CHUNK : 16384;
            r1 = r0 == 0 ? first : r0 + CHUNK;
        }
        return r1 < count ? r1 : count;
    }
};

// packed reads on the device -> records on the device.  Word extraction of range i+1 is queued ahead of
// the classification of range i.
static int classify_planes(pg_ctx *ctx, const pg_model *md, const uint32_t *d_planes, const int64_t *d_off,
                           int64_t count, int64_t total_bytes, const pg_classify_opts *opts,
                           pg_result *d_results, int32_t *d_boot_winners)
{
    PG_TRY(pg_scratch(ctx, &ctx->s_words, (size_t)total_bytes * 2 + 64));
    PG_TRY(pg_scratch(ctx, &ctx->s_nwords, (size_t)count * 4 + 64));
    PG_TRY(pg_scratch(ctx, &ctx->s_flags, (size_t)count * 2 + 64));
    uint16_t *d_words = (uint16_t *)ctx->s_words.p;
    int32_t *d_nwords = (int32_t *)ctx->s_nwords.p;
    uint8_t *d_flags = (uint8_t *)ctx->s_flags.p;
    if (count == 0) return PG_OK;
    ClassifyJob job;
    PG_TRY(job.begin(ctx, md, d_off, count, d_words, d_nwords, d_flags, opts, d_results, d_boot_winners));
    auto prepare = [&](int64_t r0, int64_t r1, cudaEvent_t *ev) -> int {
        PG_TRY(pg_extract_launch(ctx, md, d_planes + 3 * r0, d_off + r0, r1 - r0, d_words, d_nwords + r0, d_flags + 2 * r0));
        return job.fetch_counts(r0, r1, ev);
    };
    cudaEvent_t ev_cur = NULL, ev_next = NULL;
    int64_t r0 = 0, r1 = job.next_range(0);
    PG_TRY(prepare(r0, r1, &ev_cur));
    while (r0 < count) {
        const int64_t r2 = job.next_range(r1);
        if (r1 < count) PG_TRY(prepare(r1, r2, &ev_next));
        PG_TRY(job.range(r0, r1, ev_cur));
        r0 = r1; r1 = r2; ev_cur = ev_next; ev_next = NULL;
    }
    return job.finish();
}

extern "C" int pg_classify_packed(pg_ctx *ctx, const pg_model *md, const pg_reads *reads,
                                  const pg_classify_opts *opts, pg_result *results_dev, int32_t *boot_winners_dev)
{
    if (!ctx || !md || !reads || !results_dev) return pg_fail(ctx, PG_EINVAL, "pg_classify_packed: bad arguments");
    if (!md->committed) return pg_fail(ctx, PG_EINVAL, "pg_classify_packed: model has no tables (commit it first)");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    return classify_planes(ctx, md, reads->d_planes, reads->d_off, reads->count, reads->total_bytes, opts,
                           results_dev, boot_winners_dev);
}

extern "C" int pg_classify_packed_host(pg_ctx *ctx, const pg_model *md, const pg_reads *reads, const pg_classify_opts *opts,
                                       pg_result *results_host, int32_t *boot_winners_host)
{
    if (!ctx || !md || !reads || !results_host) return pg_fail(ctx, PG_EINVAL, "pg_classify_packed_host: bad arguments");
    if (!md->committed) return pg_fail(ctx, PG_EINVAL, "pg_classify_packed_host: model has no tables (commit it first)");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n = reads->count;
    if (n == 0) return PG_OK;
    PG_TRY(pg_scratch(ctx, &ctx->s_results, (size_t)n * sizeof(pg_result) + (boot_winners_host ? (size_t)n * 400 : 0)));
    pg_result *d_res = (pg_result *)ctx->s_results.p;
    int32_t *d_bw = boot_winners_host ? (int32_t *)(d_res + n) : NULL;
    PG_TRY(classify_planes(ctx, md, reads->d_planes, reads->d_off, n, reads->total_bytes, opts, d_res, d_bw));
    PG_CUDA(ctx, cudaMemcpyAsync(results_host, d_res, (size_t)n * sizeof(pg_result), cudaMemcpyDeviceToHost, ctx->stream));
    if (d_bw) PG_CUDA(ctx, cudaMemcpyAsync(boot_winners_host, d_bw, (size_t)n * 400, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PG_OK;
}

// records of the listed reads, gathered for one small copy to the host
static __global__ void k_gather_words(const uint32_t *__restrict__ src, int words_per_rec, const int32_t *__restrict__ list,
                                      int cnt, uint32_t *__restrict__ out)
{
    const int i = blockIdx.x;
    if (i >= cnt) return;
    const uint32_t *s = src + (size_t)list[i] * words_per_rec;
    for (int w = threadIdx.x; w < words_per_rec; w += blockDim.x) out[(size_t)i * words_per_rec + w] = s[w];
}

// Host ASCII in, host records out.  Three things overlap: the upload of range i+1 (copy stream), the
// kernels of range i (compute stream) and the download of the records of range i-1 (copy stream).
static int classify_host_batch(pg_ctx *ctx, const pg_model *md, const pg_seqbatch *reads, const pg_classify_opts *opts,
                               pg_result *results_host, int32_t *boot_winners_host)
{
    // reads->off need not start at 0 (a slice of a larger batch): device offsets are rebased to off[0]
    const int64_t n = reads->count;
    if (n == 0) return PG_OK;
    const int64_t base0 = reads->off[0];
    const int64_t total = reads->off[n] - base0;
    PG_TRY(pg_scratch(ctx, &ctx->s_bytes, (size_t)total + 64));
    PG_TRY(pg_scratch(ctx, &ctx->s_off, (size_t)(n + 1) * 8));
    PG_TRY(pg_scratch(ctx, &ctx->s_results, (size_t)n * sizeof(pg_result) + (boot_winners_host ? (size_t)n * 400 : 0)));
    PG_TRY(pg_scratch(ctx, &ctx->s_words, (size_t)total * 2 + 64));
    PG_TRY(pg_scratch(ctx, &ctx->s_nwords, (size_t)n * 4 + 64));
    PG_TRY(pg_scratch(ctx, &ctx->s_flags, (size_t)n * 2 + 64));
    char *d_bytes = (char *)ctx->s_bytes.p;
    int64_t *d_off = (int64_t *)ctx->s_off.p;
    pg_result *d_res = (pg_result *)ctx->s_results.p;
    int32_t *d_bw = boot_winners_host ? (int32_t *)(d_res + n) : NULL;
    uint16_t *d_words = (uint16_t *)ctx->s_words.p;
    int32_t *d_nwords = (int32_t *)ctx->s_nwords.p;
    uint8_t *d_flags = (uint8_t *)ctx->s_flags.p;
    // two copy streams: a download waits for its range's kernels and must not hold back the next upload
    if (!ctx->copy_stream) PG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->down_stream) PG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking));
    cudaStream_t cs = ctx->copy_stream, ds = ctx->down_stream;

    static const bool timing = getenv("PG_TIMING") != NULL;          // host-side laps of one call, on stderr
    const double t_start = timing ? pg_now_ms() : 0.0;
#define PG_LAPMSG(what) do { if (timing) fprintf(stderr, "[pg_classify] %-28s %8.3f ms\n", what, pg_now_ms() - t_start); } while (0)
    ClassifyJob job;
    PG_TRY(job.begin(ctx, md, d_off, n, d_words, d_nwords, d_flags, opts, d_res, d_bw));
    PG_LAPMSG("begin");
    // everything queued so far on the compute stream (scratch growth, counters) precedes the first upload
    {
        cudaEvent_t e = take_event(ctx);
        PG_CUDA(ctx, cudaEventRecord(e, ctx->stream));
        PG_CUDA(ctx, cudaStreamWaitEvent(cs, e, 0));
        ctx->ev_free.push_back(e);
    }
    auto prepare = [&](int64_t r0, int64_t r1, cudaEvent_t *ev) -> int {
        // upload on the copy stream, then pack + extract + word counts on the compute stream
        const int64_t b0 = reads->off[r0], b1 = reads->off[r1];
        for (int64_t r = r0; r <= r1; r++) job.h_off[r] = reads->off[r] - base0;
        PG_CUDA(ctx, cudaMemcpyAsync(d_bytes + (b0 - base0), reads->bytes + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, cs));
        PG_CUDA(ctx, cudaMemcpyAsync(d_off + r0, job.h_off + r0, (size_t)(r1 - r0 + 1) * 8, cudaMemcpyHostToDevice, cs));
        cudaEvent_t up = take_event(ctx);
        PG_CUDA(ctx, cudaEventRecord(up, cs));
        PG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, up, 0));
        ctx->ev_free.push_back(up);
        // words straight from the text: the 2-bit plane store is only built for reads that stay on the device
        PG_TRY(pg_extract_ascii_launch(ctx, md, d_bytes, d_off + r0, r1 - r0, d_words, d_nwords + r0, d_flags + 2 * r0));
        return job.fetch_counts(r0, r1, ev);
    };
    auto download = [&](int64_t r0, int64_t r1) -> int {
        cudaEvent_t done = take_event(ctx);
        PG_CUDA(ctx, cudaEventRecord(done, ctx->stream));
        PG_CUDA(ctx, cudaStreamWaitEvent(ds, done, 0));
        ctx->ev_free.push_back(done);
        PG_CUDA(ctx, cudaMemcpyAsync(results_host + r0, d_res + r0, (size_t)(r1 - r0) * sizeof(pg_result), cudaMemcpyDeviceToHost, ds));
        if (d_bw)
            PG_CUDA(ctx, cudaMemcpyAsync(boot_winners_host + r0 * PG_NUM_BOOT, d_bw + r0 * PG_NUM_BOOT, (size_t)(r1 - r0) * 400,
                                         cudaMemcpyDeviceToHost, ds));
        return PG_OK;
    };
    cudaEvent_t ev_cur = NULL, ev_next = NULL;
    int64_t r0 = 0, r1 = job.next_range(0, true);
    PG_TRY(prepare(r0, r1, &ev_cur));
    while (r0 < n) {
        const int64_t r2 = job.next_range(r1, true);
        if (r1 < n) PG_TRY(prepare(r1, r2, &ev_next));
        PG_TRY(job.range(r0, r1, ev_cur));
        PG_TRY(download(r0, r1));
        PG_LAPMSG("range enqueued");
        r0 = r1; r1 = r2; ev_cur = ev_next; ev_next = NULL;
    }
    PG_LAPMSG("all ranges enqueued");
    PG_TRY(job.finish());
    PG_LAPMSG("finish (fallback passes)");
    PG_CUDA(ctx, cudaStreamSynchronize(ds));
    PG_LAPMSG("downloads done");
    if (!job.redone.empty()) {
        // records rewritten by the fallback passes: gather them, one small copy, scatter on the host
        const int cnt = (int)job.redone.size();
        const int wpr = (int)(sizeof(pg_result) / 4);
        PG_TRY(pg_scratch(ctx, &ctx->s_boot, (size_t)cnt * (4 + sizeof(pg_result) + (d_bw ? 400 : 0))));
        int32_t *d_list = (int32_t *)ctx->s_boot.p;
        uint32_t *d_rec = (uint32_t *)(d_list + cnt);
        uint32_t *d_bwc = d_rec + (size_t)cnt * wpr;
        std::vector<uint32_t> rec((size_t)cnt * wpr), bwc(d_bw ? (size_t)cnt * PG_NUM_BOOT : 0);
        PG_CUDA(ctx, cudaMemcpyAsync(d_list, job.redone.data(), (size_t)cnt * 4, cudaMemcpyHostToDevice, ctx->stream));
        k_gather_words<<<cnt, 32, 0, ctx->stream>>>((const uint32_t *)d_res, wpr, d_list, cnt, d_rec);
        PG_LAUNCHED(ctx);
        PG_CUDA(ctx, cudaMemcpyAsync(rec.data(), d_rec, rec.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (d_bw) {
            k_gather_words<<<cnt, 128, 0, ctx->stream>>>((const uint32_t *)d_bw, PG_NUM_BOOT, d_list, cnt, d_bwc);
            PG_LAUNCHED(ctx);
            PG_CUDA(ctx, cudaMemcpyAsync(bwc.data(), d_bwc, bwc.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
        }
        PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < cnt; i++) {
            memcpy(results_host + job.redone[(size_t)i], rec.data() + (size_t)i * wpr, sizeof(pg_result));
            if (d_bw) memcpy(boot_winners_host + (size_t)job.redone[(size_t)i] * PG_NUM_BOOT, bwc.data() + (size_t)i * PG_NUM_BOOT, 400);
        }
    }
    PG_LAPMSG("fallback records scattered");
#undef PG_LAPMSG
    return PG_OK;
}

// Host ASCII in, host records out.  Large batches are cut into slices of at most PG_SLICE_READS reads /
// PG_SLICE_BYTES bases so that the device scratch (text, planes, 2 bytes of word id per base, records) stays
// bounded whatever the caller passes -- 100 M reads in one call included.
#define PG_SLICE_READS ((int64_t)1 << 22)
#define PG_SLICE_BYTES ((int64_t)3 << 30)
extern "C" int pg_classify(pg_ctx *ctx, const pg_model *md, const pg_seqbatch *reads, const pg_classify_opts *opts,
                           pg_result *results_host, int32_t *boot_winners_host)
{
    if (!ctx || !md || !reads || !results_host || reads->count < 0)
        return pg_fail(ctx, PG_EINVAL, "pg_classify: bad arguments");
    if (!md->committed) return pg_fail(ctx, PG_EINVAL, "pg_classify: model has no tables (commit it first)");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n = reads->count;
    int64_t st[6] = {0, 0, 0, 0, 0, 0};
    int64_t slice_reads = PG_SLICE_READS;
    if (const char *e = getenv("PG_SLICE_READS")) {         // tests: exercise the slicing on small batches
        const long long v = atoll(e);
        if (v > 0) slice_reads = v;
    }
    for (int64_t r0 = 0; r0 < n;) {
        int64_t r1 = r0 + slice_reads < n ? r0 + slice_reads : n;
        while (r1 > r0 + 1 && reads->off[r1] - reads->off[r0] > PG_SLICE_BYTES) r1 = r0 + (r1 - r0) / 2;
        pg_seqbatch slice;
        slice.bytes = reads->bytes;
        slice.off = reads->off + r0;
        slice.count = r1 - r0;
        PG_TRY(classify_host_batch(ctx, md, &slice, opts, results_host + r0,
                                   boot_winners_host ? boot_winners_host + r0 * PG_NUM_BOOT : NULL));
        st[0] += ctx->st_certified; st[1] += ctx->st_strict; st[2] += ctx->st_handed_back;
        st[3] += ctx->st_heavy; st[4] += ctx->st_items; st[5] += ctx->st_mma;
        r0 = r1;
    }
    ctx->st_certified = st[0]; ctx->st_strict = st[1]; ctx->st_handed_back = st[2];
    ctx->st_heavy = st[3]; ctx->st_items = st[4]; ctx->st_mma = st[5];
    return PG_OK;
}

extern "C" int pg_extract_words(pg_ctx *ctx, const pg_model *md, const pg_seqbatch *reads,
                                uint16_t *words_host, int32_t *n_words_host, uint8_t *reversed_host)
{
    if (!ctx || !md || !reads || reads->count < 0) return pg_fail(ctx, PG_EINVAL, "pg_extract_words: bad arguments");
    if (!md->committed) return pg_fail(ctx, PG_EINVAL, "pg_extract_words: model has no tables");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n = reads->count;
    if (n == 0) return PG_OK;
    const int64_t total = reads->off[n];
    PG_TRY(pg_scratch(ctx, &ctx->s_bytes, (size_t)total + 64));
    PG_TRY(pg_scratch(ctx, &ctx->s_off, (size_t)(n + 1) * 8));
    PG_TRY(pg_scratch(ctx, &ctx->s_cand, ((size_t)(total >> 5) + n + 2) * 12));
    PG_TRY(pg_scratch(ctx, &ctx->s_words, (size_t)total * 2 + 64));
    PG_TRY(pg_scratch(ctx, &ctx->s_nwords, (size_t)n * 4 + 64));
    PG_TRY(pg_scratch(ctx, &ctx->s_flags, (size_t)n * 2 + 64));
    PG_CUDA(ctx, cudaMemcpyAsync(ctx->s_bytes.p, reads->bytes, (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
    PG_CUDA(ctx, cudaMemcpyAsync(ctx->s_off.p, reads->off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    PG_TRY(pg_pack_launch(ctx, (const char *)ctx->s_bytes.p, (const int64_t *)ctx->s_off.p, n, (uint32_t *)ctx->s_cand.p));
    PG_TRY(pg_extract_launch(ctx, md, (const uint32_t *)ctx->s_cand.p, (const int64_t *)ctx->s_off.p, n,
                             (uint16_t *)ctx->s_words.p, (int32_t *)ctx->s_nwords.p, (uint8_t *)ctx->s_flags.p));
    if (words_host)
        PG_CUDA(ctx, cudaMemcpyAsync(words_host, ctx->s_words.p, (size_t)total * 2, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_words_host)
        PG_CUDA(ctx, cudaMemcpyAsync(n_words_host, ctx->s_nwords.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<uint8_t> fl;
    if (reversed_host) {
        fl.resize((size_t)n * 2);
        PG_CUDA(ctx, cudaMemcpyAsync(fl.data(), ctx->s_flags.p, (size_t)n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (reversed_host)
        for (int64_t i = 0; i < n; i++) reversed_host[i] = fl[2 * i];
    return PG_OK;
}

extern "C" int pg_boot_indices(pg_ctx *ctx, int32_t n, int32_t min_boot_words, uint16_t *out_host)
{
    if (!ctx || !out_host || n < 0 || n > PG_MAX_WORDS) return pg_fail(ctx, PG_EINVAL, "pg_boot_indices: bad arguments");
    PG_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<int> need(1, n);
    PG_TRY(ensure_boot_lists(ctx, need, min_boot_words));
    int k = n >> 3;
    if (k < min_boot_words) k = min_boot_words;
    const int nb = (k + 3) >> 2, il = 32 / bucket_of(n).lpr;
    std::vector<uint32_t> tmp(pg_boot_list_entries(k, il));
    PG_CUDA(ctx, pg_copy_sync(ctx, tmp.data(), ctx->d_boot_pool + ctx->h_boot_off[n], tmp.size() * 4, cudaMemcpyDeviceToHost));
    for (int run = 0; run < PG_NUM_BOOT; run++)
        for (int j = 0; j < k; j++) {
            size_t e = ((size_t)(run / il) * nb * il + (run % il)) * 4 + (size_t)(j >> 2) * il * 4 + (j & 3);
            out_host[run * k + j] = (uint16_t)(tmp[e] / PG_ROW_PITCH);
        }
    return PG_OK;
}
