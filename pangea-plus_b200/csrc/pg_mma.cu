// pg_mma.cu -- certified mode, plan 4: the best part AND the block bounds of a read as ONE integer matrix product on
// the tensor cores (tcgen05, kind::i8, accumulators in tensor memory).
//
// What plan 3 computes for a read with n words, task by task and draw by draw (k_classify_h, k_bound):
//     Sq[t][g]  = sum over the draws of task t of q[w_draw][g]          the 16 positions g of the read's best part, exact
//     LB[t][b]  = sum over the draws of task t of bm[w_draw][b]         every block b (and the 3 sibling parts): a bound
// is a product  D = C x B  of
//     C [101 x n]   C[t][j] = how often task t draws word j (task 0, the full sum: 1 everywhere).  java.util.Random is
//                   re-seeded per read, so C depends on n alone: one byte image per n, built once (k_cnt_image);
//     B [n x N]     row j = the table bytes of word w_j: gathered from L2 by cp.async, 16 bytes per (word, 16 columns).
// Columns of B (N = 48 + 16 * ceil(blocks / 16); 80 for 1 219 genera, 240 for 10 000):
//     0..15   low bytes of the 16 exact deficits of the best part      (qx, 12-bit values split in two bytes)
//     16..31  their high bytes:  Sq = 256 * D[16 + i] + D[i], an exact integer -- the same value plan 3 adds up
//     32..    one byte per block, c = min(bm >> 2, 255) (bm8x): rounded DOWN and capped, so 4 * D[col] <= LB <= every
//             genus sum of the block; a block is dismissed when D[col] > (champion + margin) >> 2
//     last 16 the coarse part minima (hm8x) of the four blocks around the best one: its 3 sibling parts compete like blocks
// One persistent CTA per SM, three roles over mbarriers:
//     8 producer warps  gather the rows of 128 words per stage into a 4-stage ring (MN-major, no swizzle: 16 columns of
//                       8 consecutive words = one 128-byte core matrix), and the count image when n changes;
//     1 issuing thread  tcgen05.mma M = 128 (tasks), N, K = 32 words per instruction, A = count image (K-major) and
//                       B = ring stage straight from shared memory, accumulating in one of two TMEM buffers;
//     4 epilogue warps  thread = task: tcgen05.ld its row, champion of the part -> champion slot and near-ties (what
//                       k_classify_h's epilogue does), then the bound columns -> (task, block) items (what k_bound does).
// Results are those of plan 3 (and so of the strict kernels): the exact columns are the same integers, the bound columns
// only decide which pairs k_light evaluates exactly.  tests: test_certified_plans_agree and every parity test (default plan).
#include "pg_certified.cuh"

#define PG_MMA_MAXN   640                   // reads with more good words go through plan 3
#define PG_MMA_KC     128                   // words per ring stage
#define PG_MMA_STAGES 4
#define PG_MMA_IMG    (128 * PG_MMA_MAXN)   // bytes of one count image slot
#define PG_MMA_NPROD  256                   // producer threads
#define PG_MMA_THREADS (128 + 32 + PG_MMA_NPROD)
#define PG_MMA_LIST   2048                  // open pairs per read kept in shared memory (more: the read is "heavy")
#define PG_X8_SHIFT   2

// ------------------------------------------------------------------ tables

// qx[(blk * 4 + part) * 65536 + w][32] = low bytes | high bytes of the part's 16 deficits
__global__ void k_qx(const uint16_t *__restrict__ q, int ntile64, uint8_t *__restrict__ qx)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;        // (blk, part, w)
    const int w = (int)(idx & (PG_NWORDS - 1));
    const int bp = (int)(idx >> 16);
    if (bp >= ntile64 * PG_PARTS) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(q + ((size_t)(bp / PG_PARTS) * PG_NWORDS + w) * 64 + (bp % PG_PARTS) * PG_PART_POS);
    const uint4 a = src[0], b = src[1];
    const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};           // two 16-bit deficits each
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        lo[i] = __byte_perm(v[2 * i], v[2 * i + 1], 0x6420);
        hi[i] = __byte_perm(v[2 * i], v[2 * i + 1], 0x7531);
    }
    uint4 *dst = reinterpret_cast<uint4 *>(qx + idx * 32);
    dst[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    dst[1] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
}

// bm8x[w][pitch8]: one byte per block; hm8x[w][hpitch]: one byte per part (4 * block + part); padding = 255
__global__ void k_x8(const uint16_t *__restrict__ bm, const uint16_t *__restrict__ hm, int ntile64, int pitch8, int hpitch,
                     uint8_t *__restrict__ bm8x, uint8_t *__restrict__ hm8x)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = pitch8 + hpitch;
    const size_t w = idx / (size_t)per;
    const int col = (int)(idx - w * (size_t)per);
    if (w >= PG_NWORDS) return;
    uint32_t c = 255u;
    if (col < pitch8) {
        if (col < ntile64) c = min((uint32_t)bm[((size_t)(col / PG_GB) * PG_NWORDS + w) * 32 + (col % PG_GB)] >> PG_X8_SHIFT, 255u);
        bm8x[w * (size_t)pitch8 + col] = (uint8_t)c;
    } else {
        const int pb = col - pitch8;
        if (pb < PG_PARTS * ntile64) c = min((uint32_t)hm[((size_t)(pb >> 5) * PG_NWORDS + w) * 32 + (pb & 31)] >> PG_X8_SHIFT, 255u);
        hm8x[w * (size_t)hpitch + pb] = (uint8_t)c;
    }
}

int pg_mma_build_tables(pg_model *md)
{
    pg_ctx *ctx = md->ctx;
    const int pitch8 = (md->ntile64 + 15) & ~15, hpitch = (PG_PARTS * md->ntile64 + 15) & ~15;
    if (md->x8_blocks != md->ntile64) {
        cudaFree(md->d_qx); cudaFree(md->d_bm8x); cudaFree(md->d_hm8x);
        md->d_qx = md->d_bm8x = md->d_hm8x = NULL;
        md->x8_blocks = 0;
    }
    md->pitch8 = pitch8;
    md->hpitch = hpitch;
    if (32 + pitch8 + 16 > 256) return PG_OK;            // more than 208 blocks: one instruction cannot hold the columns; plan 3
    const size_t qxb = (size_t)md->ntile64 * PG_PARTS * PG_NWORDS * 32;
    if (!md->d_qx) {
        cudaError_t e;
        if ((e = cudaMalloc(&md->d_qx, qxb)) != cudaSuccess || (e = cudaMalloc(&md->d_bm8x, (size_t)PG_NWORDS * pitch8)) != cudaSuccess ||
            (e = cudaMalloc(&md->d_hm8x, (size_t)PG_NWORDS * hpitch)) != cudaSuccess) {
            (void)cudaGetLastError();
            cudaFree(md->d_qx); cudaFree(md->d_bm8x); cudaFree(md->d_hm8x);
            md->d_qx = md->d_bm8x = md->d_hm8x = NULL;
            return PG_OK;                                // no room for the byte tables: plan 3 needs none of them
        }
    }
    k_qx<<<(unsigned)(((size_t)md->ntile64 * PG_PARTS * PG_NWORDS + 255) / 256), 256, 0, ctx->stream>>>(md->d_qtable, md->ntile64, md->d_qx);
    PG_LAUNCHED(ctx);
    const size_t cells = (size_t)PG_NWORDS * (pitch8 + hpitch);
    k_x8<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->stream>>>(md->d_bmtable, md->d_hmtable, md->ntile64, pitch8, hpitch, md->d_bm8x,
                                                                  md->d_hm8x);
    PG_LAUNCHED(ctx);
    md->x8_blocks = md->ntile64;
    return PG_OK;
}

// ------------------------------------------------------------------ count images

// One CTA per distinct n: image[kc][m][16] (kc = 16-word chunk, m = task row 0..127), the K-major no-swizzle operand
// layout: 8 rows x 16 bytes = one core matrix, row groups 128 bytes apart, 16-word chunks 2 048 bytes apart.
// Row 0 = the full sum (every word once), rows 1..100 = the draw counts of the replicates (from the sample lists of
// k_boot_indices, interleave 4), rows 101..127 = 0.
__global__ void __launch_bounds__(128)
k_cnt_image(const uint32_t *__restrict__ boot_pool, const int32_t *__restrict__ boot_off, const int32_t *__restrict__ ns,
            int min_boot, uint8_t *__restrict__ images)
{
    const int n = ns[blockIdx.x], t = threadIdx.x;
    uint8_t *img = images + (size_t)n * PG_MMA_IMG;
    const int kpad = (n + 31) & ~31;
    uint4 *z = reinterpret_cast<uint4 *>(img);
    for (int i = t; i < 128 * kpad / 16; i += 128) z[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    int k = n >> 3;
    if (k < min_boot) k = min_boot;
    const int nb = (k + 3) >> 2;
    if (t == 0) {
        for (int j = 0; j < n; j++) img[(size_t)(j >> 4) * 2048 + (j & 15)] = 1;
    } else if (t <= PG_NUM_BOOT && n > 0) {
        const int r = t - 1;
        const uint32_t *lp = boot_pool + boot_off[n] + ((size_t)(r >> 2) * nb * 4 + (r & 3)) * 4;
        for (int d = 0; d < k; d++) {
            const uint32_t j = lp[(size_t)(d >> 2) * 16 + (d & 3)] / PG_ROW_PITCH;
            img[(size_t)(j >> 4) * 2048 + t * 16 + (j & 15)]++;           // one thread per row: no race; k <= 80 < 256
        }
    }
}

int pg_mma_ensure_images(pg_ctx *ctx, const std::vector<int> &need_n, int min_boot)
{
    if (ctx->cnt_min_words != min_boot || ctx->cnt_built.empty()) {
        ctx->cnt_built.assign(PG_MMA_MAXN + 1, 0);
        ctx->cnt_min_words = min_boot;
    }
    std::vector<int32_t> ns;
    for (int n : need_n)
        if (n >= 1 && n <= PG_MMA_MAXN && !ctx->cnt_built[(size_t)n]) { ns.push_back(n); ctx->cnt_built[(size_t)n] = 1; }
    if (ns.empty()) return PG_OK;
    if (!ctx->d_cnt_img) PG_CUDA(ctx, cudaMalloc(&ctx->d_cnt_img, (size_t)(PG_MMA_MAXN + 1) * PG_MMA_IMG));
    int32_t *d_ns = NULL;
    PG_CUDA(ctx, pg_dev_alloc(ctx, (void **)&d_ns, ns.size() * 4));
    PG_CUDA(ctx, cudaMemcpyAsync(d_ns, ns.data(), ns.size() * 4, cudaMemcpyHostToDevice, ctx->stream));   // pageable: staged before return
    k_cnt_image<<<(unsigned)ns.size(), 128, 0, ctx->stream>>>(ctx->d_boot_pool, ctx->d_boot_off, d_ns, min_boot, ctx->d_cnt_img);
    PG_LAUNCHED(ctx);
    pg_dev_free(ctx, d_ns);
    return PG_OK;
}

// ------------------------------------------------------------------ PTX pieces

__device__ __forceinline__ uint32_t pgm_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pgm_cp16(uint32_t s, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void pgm_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pgm_mbar_arrive(uint32_t bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(bar) : "memory");
}
// bounded wait: a protocol error must end the kernel with an error, never hang the device
__device__ __forceinline__ void pgm_mbar_wait(uint32_t bar, uint32_t parity)
{
    for (unsigned spin = 0; spin < (1u << 28); spin++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}
// shared-memory matrix descriptor, no swizzle: start address, leading / stride byte offsets (16-byte units), version 1
__device__ __forceinline__ uint64_t pgm_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ULL << 46);
}
__device__ __forceinline__ void pgm_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void pgm_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
#define PGM_LD16(taddr, v)                                                                                            \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),         \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])    \
                 : "r"(taddr))
#define PGM_LD_WAIT() asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory")

// ------------------------------------------------------------------ the kernel

struct PgMmaArgs {
    const uint8_t *qx, *bm8x, *hm8x, *images;
    int pitch8, hpitch, ntile64, kmax;               // kmax: words of the longest read of the launch, rounded up to 32
    const uint16_t *words;
    const int64_t *off;
    const int32_t *nwords;
    const uint8_t *flags;
    const int32_t *order;
    int nreads_b;
    int64_t slot0;
    int min_boot;
    const unsigned long long *blockmask;
    double vmax;
    unsigned long long *champ;
    unsigned int *ncand;
    unsigned long long *cand;
    const int32_t *guess;
    unsigned long long *items;
    unsigned int *item_count;
    unsigned int item_cap;
    uint8_t *heavy;
    unsigned int light_max;
};

// the reads of one CTA, in the order every role walks them
struct PgMmaRead {
    int slot, n, gs;
    int64_t read;
};
__device__ __forceinline__ bool pgm_next_read(const PgMmaArgs &a, int &slot, PgMmaRead &r)
{
    for (; slot < a.nreads_b; slot += (int)gridDim.x) {
        const int64_t read = a.order[slot];
        if (a.flags[2 * read + 1]) continue;            // short read (A2)
        const int n = a.nwords[read];
        if (n == 0) continue;                           // no word: phase 2 writes genus 0 directly
        r.slot = slot; r.n = n; r.read = read; r.gs = a.guess[(size_t)a.slot0 + slot];
        slot += (int)gridDim.x;
        return true;
    }
    return false;
}

__global__ void __launch_bounds__(PG_MMA_THREADS, 1)
k_mma_bound(const __grid_constant__ PgMmaArgs a)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_full[PG_MMA_STAGES], s_empty[PG_MMA_STAGES], s_dfull[2], s_dempty[2];
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_list[2][PG_MMA_LIST];
    __shared__ unsigned int s_cnt[2], s_base;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nch = 3 + a.pitch8 / 16;                   // 16-column chunks of B
    const int N = nch * 16;
    uint8_t *sA = smem;                                  // count image of the current n: [kmax / 16][128][16]
    uint8_t *ring = smem + (size_t)128 * a.kmax;         // [stage][chunk][PG_MMA_KC words][16]
    const uint32_t stage_bytes = (uint32_t)nch * PG_MMA_KC * 16u;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(pgm_smem(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        for (int s = 0; s < PG_MMA_STAGES; s++) { pgm_mbar_init(pgm_smem(&s_full[s]), PG_MMA_NPROD); pgm_mbar_init(pgm_smem(&s_empty[s]), 1); }
        for (int i = 0; i < 2; i++) { pgm_mbar_init(pgm_smem(&s_dfull[i]), 1); pgm_mbar_init(pgm_smem(&s_dempty[i]), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        s_cnt[0] = s_cnt[1] = 0u;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp >= 5) {
        // ================================================================ producers
        const int p = tid - 160, jr = p & (PG_MMA_KC - 1), c0 = p >> 7;      // this thread's row of a stage; its chunks: c0, c0 + 2, ...
        int slot = (int)blockIdx.x, cur_n = -1;
        unsigned it = 0;
        bool owe = false;                                // the previous stage is issued but not yet signalled
        PgMmaRead r;
        bool have = pgm_next_read(a, slot, r);
        int kc = 0;
        uint32_t wid = 0u;
        if (have && jr < r.n) wid = a.words[a.off[r.read] + jr];
        while (have) {
            // the stage after this one: its word id is fetched now and used a whole stage later
            PgMmaRead rn = r;
            int kcn = kc + 1;
            bool have_n = true;
            if (kcn * PG_MMA_KC >= r.n) { kcn = 0; have_n = pgm_next_read(a, slot, rn); }
            uint32_t wid_n = 0u;
            if (have_n && kcn * PG_MMA_KC + jr < rn.n) wid_n = a.words[a.off[rn.read] + kcn * PG_MMA_KC + jr];

            const unsigned s = it % PG_MMA_STAGES;
            if (kc == 0 && r.n != cur_n) {
                // a new count image: every product that reads the old one must be done.  Signal the stage we owe first
                // (its products cannot start before), then wait for the products of the last stage issued.
                if (owe) {
                    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                    pgm_mbar_arrive(pgm_smem(&s_full[(it - 1) % PG_MMA_STAGES]));
                    owe = false;
                }
                if (it > 0) pgm_mbar_wait(pgm_smem(&s_empty[(it - 1) % PG_MMA_STAGES]), ((it - 1) / PG_MMA_STAGES) & 1u);
                const uint8_t *img = a.images + (size_t)r.n * PG_MMA_IMG;
                const int pieces = 128 * ((r.n + 31) & ~31) / 16;
                for (int i = p; i < pieces; i += PG_MMA_NPROD) pgm_cp16(pgm_smem(sA) + (uint32_t)i * 16u, img + (size_t)i * 16);
                cur_n = r.n;
            }
            pgm_mbar_wait(pgm_smem(&s_empty[s]), ((it / PG_MMA_STAGES) & 1u) ^ 1u);   // the ring slot is free
            if (kc * PG_MMA_KC + jr < r.n) {
                const int blk = r.gs / PG_PARTS;
                const uint32_t dst0 = pgm_smem(ring) + s * stage_bytes + (uint32_t)jr * 16u;
                for (int c = c0; c < nch; c += PG_MMA_NPROD / PG_MMA_KC) {
                    const uint8_t *src;
                    if (c < 2) src = a.qx + ((size_t)r.gs * PG_NWORDS + wid) * 32 + c * 16;
                    else if (c < nch - 1) src = a.bm8x + (size_t)wid * a.pitch8 + (c - 2) * 16;
                    else src = a.hm8x + (size_t)wid * a.hpitch + (blk >> 2) * 16;
                    pgm_cp16(dst0 + (uint32_t)c * (PG_MMA_KC * 16u), src);
                }
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            if (owe) {
                asm volatile("cp.async.wait_group 1;\n" ::: "memory");               // everything but the stage just issued has landed
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");       // generic-proxy writes -> the tensor core's reads
                pgm_mbar_arrive(pgm_smem(&s_full[(it - 1) % PG_MMA_STAGES]));
            }
            owe = true;
            it++;
            r = rn; kc = kcn; have = have_n; wid = wid_n;
        }
        if (owe) {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            pgm_mbar_arrive(pgm_smem(&s_full[(it - 1) % PG_MMA_STAGES]));
        }
    } else if (warp == 4) {
        // ================================================================ the issuing thread
        if (lane == 0) {
            const uint32_t idesc = (2u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);   // u8 x u8 -> s32, B MN-major, M = 128
            const uint64_t adesc0 = pgm_desc(pgm_smem(sA), 2048u, 128u);             // K chunks 2 048 B apart, row groups 128 B
            int slot = (int)blockIdx.x;
            unsigned it = 0, rd = 0;
            PgMmaRead r;
            while (pgm_next_read(a, slot, r)) {
                const unsigned acc = rd & 1u;
                pgm_mbar_wait(pgm_smem(&s_dempty[acc]), ((rd >> 1) & 1u) ^ 1u);       // the epilogue has drained this buffer
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t dcol = tmem + acc * 256u;
                const int nks = (r.n + 31) >> 5;
                for (int ks0 = 0; ks0 < nks; ks0 += PG_MMA_KC / 32) {
                    const unsigned s = it % PG_MMA_STAGES;
                    pgm_mbar_wait(pgm_smem(&s_full[s]), (it / PG_MMA_STAGES) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    // words 8 apart 128 B apart (leading), 16-column chunks PG_MMA_KC * 16 B apart (stride)
                    const uint64_t bdesc0 = pgm_desc(pgm_smem(ring) + s * stage_bytes, 128u, PG_MMA_KC * 16u);
#pragma unroll
                    for (int u = 0; u < PG_MMA_KC / 32; u++) {
                        const int ks = ks0 + u;
                        if (ks < nks)
                            pgm_mma_i8(dcol, adesc0 + (uint64_t)(ks * (4096 >> 4)), bdesc0 + (uint64_t)(u * (512 >> 4)), idesc, ks > 0 ? 1u : 0u);
                    }
                    pgm_commit(pgm_smem(&s_empty[s]));                                // the slot is free once these products are done
                    it++;
                }
                pgm_commit(pgm_smem(&s_dfull[acc]));
                rd++;
            }
        }
    } else {
        // ================================================================ epilogue: thread = task
        const int t = tid;
        int slot = (int)blockIdx.x;
        unsigned rd = 0;
        PgMmaRead r;
        while (pgm_next_read(a, slot, r)) {
            const unsigned acc = rd & 1u;
            const size_t rc = (size_t)a.slot0 + r.slot;
            const int n = r.n;
            int k = n >> 3;
            if (k < a.min_boot) k = a.min_boot;
            const bool live = t == 0 || (t <= PG_NUM_BOOT && k > 0);
            const int best = r.gs / PG_PARTS, own = r.gs % PG_PARTS;
            const uint32_t gbase = (uint32_t)best * 64u, poff = (uint32_t)own * PG_PART_POS;
            const uint32_t vb16 = (uint32_t)(a.blockmask[best] >> poff) & 0xFFFFu;
            const uint32_t margin = pg_margin(t == 0 ? n : k, a.vmax);
            pgm_mbar_wait(pgm_smem(&s_dfull[acc]), (rd >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t trow = tmem + acc * 256u + ((uint32_t)(warp * 32) << 16);
            uint32_t lo[16], hi[16];
            PGM_LD16(trow, lo);
            PGM_LD16(trow + 16u, hi);
            PGM_LD_WAIT();
            // ---- the part's champion and near-ties (k_classify_h's epilogue; the slot is still empty)
            uint32_t sum[16], bkey = 0xFFFFFFFFu;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                sum[i] = (hi[i] << 8) + lo[i];
                const uint32_t key = ((vb16 >> i) & 1u) ? (sum[i] << 6) + poff + (uint32_t)i : 0xFFFFFFFFu;
                bkey = min(bkey, key);
            }
            const uint32_t csum = bkey >> 6, cpos = bkey & 63u;
            const uint32_t thr = csum + margin;
            unsigned int *lcnt = &s_cnt[acc];
            uint32_t *list = s_list[acc];
            if (live) {
                a.champ[rc * (PG_NUM_BOOT + 1) + t] = ((unsigned long long)csum << 32) | (gbase + cpos);
#pragma unroll
                for (int i = 0; i < 16; i++)
                    if (((vb16 >> i) & 1u) && sum[i] <= thr && poff + (uint32_t)i != cpos)
                        pg_emit(a.ncand + rc, a.cand + rc * PG_CANDCAP, t, gbase + poff + (uint32_t)i, sum[i]);
            }
            // ---- bound columns: 4 * D <= the smallest genus sum of the block
            const uint32_t thr8 = thr >> PG_X8_SHIFT;
            for (int cc = 0; cc < a.pitch8 / 16; cc++) {
                uint32_t v[16];
                PGM_LD16(trow + 32u + (uint32_t)cc * 16u, v);
                PGM_LD_WAIT();
                if (!live) continue;
                uint32_t open = 0u;
#pragma unroll
                for (int i = 0; i < 16; i++) open |= (v[i] <= thr8 ? 1u : 0u) << i;
                while (open) {
                    const int i = __ffs(open) - 1;
                    open &= open - 1u;
                    const int b = cc * 16 + i;
                    if (b < a.ntile64 && b != best) {
                        const unsigned int pos = atomicAdd(lcnt, 1u);
                        if (pos < PG_MMA_LIST) list[pos] = ((uint32_t)t << 16) | (uint32_t)b;
                    }
                }
            }
            {
                uint32_t v[16];
                PGM_LD16(trow + 32u + (uint32_t)a.pitch8, v);
                PGM_LD_WAIT();
                if (live) {
                    const int q4 = (best & 3) * 4;
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int part = i - q4;
                        if (part >= 0 && part < PG_PARTS && part != own && v[i] <= thr8) {
                            const unsigned int pos = atomicAdd(lcnt, 1u);
                            if (pos < PG_MMA_LIST) list[pos] = ((uint32_t)t << 16) | 0x8000u | ((uint32_t)part << 13) | (uint32_t)best;
                        }
                    }
                }
            }
            // the accumulator is read: hand the buffer back to the issuing thread
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            pgm_mbar_arrive(pgm_smem(&s_dempty[acc]));
            // ---- the read's open pairs -> the global item list (one atomic per read), or the read is heavy
            asm volatile("bar.sync 1, 128;\n" ::: "memory");
            const unsigned int cnt = *lcnt;
            if (t == 0) {
                unsigned int b = 0xFFFFFFFFu;
                if (cnt <= a.light_max && cnt <= PG_MMA_LIST) {
                    b = 0u;
                    if (cnt > 0u) {
                        b = atomicAdd(a.item_count, cnt);
                        if (b > a.item_cap || cnt > a.item_cap - b) {      // buffer full: blank what fits, redo the read
                            for (unsigned int i = b; i < a.item_cap && i < b + cnt; i++) a.items[i] = (unsigned long long)PG_ITEM_NULL << 16;
                            b = 0xFFFFFFFFu;
                        }
                    }
                }
                if (b == 0xFFFFFFFFu) a.heavy[rc] = 1;
                s_base = b;
                s_cnt[acc ^ 1u] = 0u;                                      // the other list is idle until the next read
            }
            asm volatile("bar.sync 1, 128;\n" ::: "memory");
            const unsigned int gb = s_base;
            if (gb != 0xFFFFFFFFu)
                for (unsigned int i = t; i < cnt; i += 128) a.items[gb + i] = ((unsigned long long)rc << 32) | list[i];
            rd++;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512) : "memory");
}

// ------------------------------------------------------------------ host

bool pg_mma_usable(const pg_model *md, int nmax)
{
    static int env = -2;                                 // PG_MMA=0: plan 3 kernels (A/B switch)
    if (env == -2) { const char *e = getenv("PG_MMA"); env = e ? atoi(e) : -1; }
    return env != 0 && md->d_qx && md->x8_blocks == md->ntile64 && nmax >= 1 && nmax <= PG_MMA_MAXN && md->ctx->d_cnt_img;
}

// best part + bounds of one bucket (the reads carry n <= PG_MMA_MAXN words); champion slots must be empty
int pg_mma_launch(pg_ctx *ctx, const pg_model *md, unsigned nreads_b, int nmax, const uint16_t *d_words, const int64_t *d_off,
                  const int32_t *d_nwords, const uint8_t *d_flags, const int32_t *d_order, int64_t slot0, int min_boot,
                  const PgCertBufs &cb, const int32_t *d_guess, unsigned int light_max)
{
    PgMmaArgs a;
    a.qx = md->d_qx; a.bm8x = md->d_bm8x; a.hm8x = md->d_hm8x; a.images = ctx->d_cnt_img;
    a.pitch8 = md->pitch8; a.hpitch = md->hpitch; a.ntile64 = md->ntile64;
    a.kmax = (nmax + 31) & ~31;
    a.words = d_words; a.off = d_off; a.nwords = d_nwords; a.flags = d_flags; a.order = d_order;
    a.nreads_b = (int)nreads_b; a.slot0 = slot0; a.min_boot = min_boot; a.blockmask = md->d_blockmask; a.vmax = md->vmax;
    a.champ = cb.champ; a.ncand = cb.ncand; a.cand = cb.cand; a.guess = d_guess;
    a.items = cb.items; a.item_count = cb.counters + 2; a.item_cap = cb.item_cap; a.heavy = cb.heavy; a.light_max = light_max;
    const int nch = 3 + md->pitch8 / 16;
    size_t smem = (size_t)128 * a.kmax + (size_t)PG_MMA_STAGES * nch * PG_MMA_KC * 16;
    // the kernel owns all 512 columns of tensor memory: never two CTAs on one SM
    if (smem < (size_t)120 * 1024) smem = (size_t)120 * 1024;
    PG_CUDA(ctx, pg_smem_unlock(ctx, k_mma_bound));
    const unsigned grid = nreads_b < (unsigned)ctx->sm_count ? nreads_b : (unsigned)ctx->sm_count;
    k_mma_bound<<<grid, PG_MMA_THREADS, smem, ctx->stream>>>(a);
    PG_LAUNCHED(ctx);
    return PG_OK;
}
