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
// One persistent CTA per SM (it owns all 512 columns of tensor memory), three roles over mbarriers:
//     8 producer warps  gather the rows of 128 words per stage into a ring of up to 8 stages (MN-major, no swizzle: 16
//                       columns of 8 consecutive words = one 128-byte core matrix) and the count image when n changes;
//                       word ids two reads ahead, the exact rows prefetched into L2 one read ahead; a thread's arrival on
//                       the stage's barrier fires when its copies have landed (cp.async.mbarrier.arrive.noinc);
//     1 issuing warp    tcgen05.mma M = 128 (tasks), N, K = 32 words per instruction, A = count image (K-major) and
//                       B = ring stage straight from shared memory, accumulating in one of two TMEM buffers; the whole
//                       warp walks the loop (uniform operands), one elected lane issues; tcgen05.commit frees the stage;
//     4 epilogue warps  thread = task: tcgen05.ld its row, champion of the part -> champion slot and near-ties (what
//                       k_classify_h's epilogue does), then the bound columns -> (task, block) items (what k_bound does).
// Every barrier wait is bounded: a protocol error ends the kernel with a launch failure, it cannot hang the device.
// Results are those of plan 3 (and so of the strict kernels): the exact columns are the same integers, the bound columns
// only decide which pairs k_light evaluates exactly.  tests: test_certified_plans_agree and every parity test (default plan).
#include "pg_certified.cuh"

#define PG_MMA_MAXN   640                   // reads with more good words go through plan 3
#define PG_MMA_KC     128                   // words per ring stage
#define PG_MMA_MAXSTAGES 8                   // ring depth: as many stages as shared memory holds, at least 3
#define PG_MMA_IMG    (128 * PG_MMA_MAXN)   // bytes of one count image slot
#define PG_MMA_NPROD  256                   // producer threads
#define PG_MMA_NEPI   128                   // epilogue threads.  256 = two groups of four warps taking the reads in turn (one per
                                            // TMEM buffer) was measured: 96 registers and spills, more warps per scheduler: slower;
                                            // with setmaxnreg (64 / 160) it compiled without spills and did not come back from the
                                            // device (the 17th warp is a warpgroup of its own); the bounded barrier waits ended it
#define PG_MMA_THREADS (PG_MMA_NEPI + 32 + PG_MMA_NPROD)
// units of the byte bounds: 2^shift / 128 nat, capped at 255 of them.  A larger shift reaches further (shift 2: 8 nat,
// 3: 16 nat) and rounds more away per draw; PG_X8_SHIFT in the environment overrides the default for A/B runs.
static int pg_x8_shift()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("PG_X8_SHIFT"); v = e ? atoi(e) : 2; if (v < 0 || v > 4) v = 2; }
    return v;
}
#define PG_MMA_WLIST  512                   // open pairs per (read, epilogue warp) above which the read is "heavy"
#define PG_MMA_RESERVE 128                  // entries of the global item list an epilogue warp reserves at a time

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
                     uint8_t *__restrict__ bm8x, uint8_t *__restrict__ hm8x, int shift)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = pitch8 + hpitch;
    const size_t w = idx / (size_t)per;
    const int col = (int)(idx - w * (size_t)per);
    if (w >= PG_NWORDS) return;
    uint32_t c = 255u;
    if (col < pitch8) {
        if (col < ntile64) c = min((uint32_t)bm[((size_t)(col / PG_GB) * PG_NWORDS + w) * 32 + (col % PG_GB)] >> shift, 255u);
        bm8x[w * (size_t)pitch8 + col] = (uint8_t)c;
    } else {
        const int pb = col - pitch8;
        if (pb < PG_PARTS * ntile64) c = min((uint32_t)hm[((size_t)(pb >> 5) * PG_NWORDS + w) * 32 + (pb & 31)] >> shift, 255u);
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
        if (cudaMalloc(&md->d_qx, qxb) != cudaSuccess || cudaMalloc(&md->d_bm8x, (size_t)PG_NWORDS * pitch8) != cudaSuccess ||
            cudaMalloc(&md->d_hm8x, (size_t)PG_NWORDS * hpitch) != cudaSuccess) {
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
                                                                  md->d_hm8x, pg_x8_shift());
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
    if (!ctx->d_cnt_img && cudaMalloc(&ctx->d_cnt_img, (size_t)(PG_MMA_MAXN + 1) * PG_MMA_IMG) != cudaSuccess) {
        (void)cudaGetLastError();                        // no room for the images: pg_mma_usable() says no, plan 3 runs
        ctx->d_cnt_img = NULL;
        ctx->cnt_built.clear();
        return PG_OK;
    }
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
#ifdef PG_MMA_CA
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g) : "memory");
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g) : "memory");
#endif
}
#define PGM_CP16 pgm_cp16
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
    for (unsigned spin = 0; spin < (1u << 26); spin++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        __nanosleep(40);                                 // a waiting warp must not take issue slots from the working ones
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
// one lane of the (converged) warp
__device__ __forceinline__ bool pgm_elect()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0u;
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
    unsigned nstage;                                 // ring depth of this launch
    int shift;                                       // log2 of the byte bounds' unit (pg_x8_shift)
    const uint16_t *words;
    int nreads_b;
    int64_t slot0;
    int min_boot;
    unsigned long long *champ;
    unsigned int *ncand;
    unsigned long long *cand;
    unsigned long long *items;
    unsigned int *item_count, *items_total;        // list cursor (reservations incl. blanks); pairs written, for the statistics
    unsigned int item_cap;
    uint8_t *heavy;
    unsigned int light_max;
    const int4 *meta;                                 // [nreads_b + (gridDim.x + 1) * PG_MMA_RUN][2], k_mma_meta
    long long *prof;                                  // PG_MMA_PROF=1: per CTA 8 cycle counters, else NULL
};

// What the roles need to know about the read in slot s, gathered once by k_mma_meta so that the persistent kernel finds
// it with ONE 16-byte load, issued a whole read ahead (order -> flags / nwords / off -> guess -> blockmask is a chain of
// four dependent loads: ~2 000 cycles per read when every role walks it on its own).
//   x = n (0: the read is skipped -- short, or no word), y = part id | validity bits of the part's 16 positions << 16,
//   z, w = offset of the read's word ids; second entry: x, y = certified margins of the full sum and of a replicate
__global__ void k_mma_meta(const int64_t *__restrict__ off, const int32_t *__restrict__ nwords, const uint8_t *__restrict__ flags,
                           const int32_t *__restrict__ order, int nreads_b, int64_t slot0, const int32_t *__restrict__ guess,
                           const unsigned long long *__restrict__ blockmask, int min_boot, double vmax, int4 *__restrict__ meta,
                           int nmeta)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nmeta) return;
    int4 m = make_int4(0, 0, 0, 0), m2 = make_int4(0, 0, 0, 0);
    if (s < nreads_b) {
        const int64_t read = order[s];
        if (!flags[2 * read + 1]) {
            const int gs = guess[(size_t)slot0 + s];
            const uint32_t vb16 = (uint32_t)(blockmask[gs / PG_PARTS] >> ((gs % PG_PARTS) * PG_PART_POS)) & 0xFFFFu;
            const int64_t wo = off[read];
            const int n = nwords[read];
            int k = n >> 3;
            if (k < min_boot) k = min_boot;
            m = make_int4(n, (int)((uint32_t)gs | (vb16 << 16)), (int)(uint32_t)wo, (int)(uint32_t)((unsigned long long)wo >> 32));
            m2 = make_int4((int)pg_margin(n, vmax), (int)pg_margin(k, vmax), 0, 0);     // double-precision work kept out of the epilogue
        }
    }
    meta[2 * s] = m;
    meta[2 * s + 1] = m2;
}

struct PgMmaRead {
    int slot, n, gs;
    uint32_t vb16, margin_full, margin_rep;
    int64_t woff;
};
// A role's cursor over the slots of its CTA (blockIdx.x, + gridDim.x, ...).  The metadata of the NEXT slot is always in
// flight; meta[] is padded with empty slots past the end, so the look-ahead never needs a bound check.
// A CTA takes the slots in blocks of PG_MMA_RUN consecutive ones (block b = blockIdx.x, + gridDim.x, ...): the order array
// is sorted by word count, so a run shares its count image (one slot at a time, round robin, changed the image at
// nearly every read of a batch with many different lengths: 64 KB of copies per read beside 39 KB of rows).
#define PG_MMA_RUN 16
struct PgMmaCursor {
    int slot;
    int a_n, a_gv, a_wlo, a_whi, a_mf, a_mr;          // metadata of `slot`, in flight
    // Six scalar loads, not two 16-byte ones: a vector load lands in an aligned register quad and the compiler then MOVES
    // it into the loop-carried registers right behind the load -- a full L2 round trip per read in every role (the
    // hottest instruction of the first profile: 6.6 % of all stall samples on one MOV).
    __device__ __forceinline__ static int ld(const int *p)
    {
        int v;
        asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
        return v;
    }
    __device__ __forceinline__ void fetch(const PgMmaArgs &a)
    {
        const int *m = reinterpret_cast<const int *>(a.meta) + (size_t)slot * 8;
        a_n = ld(m); a_gv = ld(m + 1); a_wlo = ld(m + 2); a_whi = ld(m + 3); a_mf = ld(m + 4); a_mr = ld(m + 5);
    }
    __device__ __forceinline__ void start(const PgMmaArgs &a)
    {
        slot = (int)blockIdx.x * PG_MMA_RUN;
        fetch(a);
    }
    // the next slot of this CTA, skipped reads included (n = 0: every role passes over them by itself -- a loop HERE
    // made the compiler copy the freshly loaded registers at once, i.e. wait for the load it had just issued)
    __device__ __forceinline__ bool next(const PgMmaArgs &a, PgMmaRead &r)
    {
        if (slot >= a.nreads_b) return false;
        r.slot = slot; r.n = a_n; r.gs = a_gv & 0xFFFF; r.vb16 = (uint32_t)a_gv >> 16;
        r.margin_full = (uint32_t)a_mf; r.margin_rep = (uint32_t)a_mr;
        r.woff = (int64_t)(((unsigned long long)(uint32_t)a_whi << 32) | (uint32_t)a_wlo);
        slot = ((slot + 1) & (PG_MMA_RUN - 1)) ? slot + 1 : slot + 1 + ((int)gridDim.x - 1) * PG_MMA_RUN;
        fetch(a);
        return true;
    }
};

__global__ void __launch_bounds__(PG_MMA_THREADS, 1)
k_mma_bound(const __grid_constant__ PgMmaArgs a)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_full[PG_MMA_MAXSTAGES], s_empty[PG_MMA_MAXSTAGES], s_dfull[2], s_dempty[2];
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_wlist[PG_MMA_NEPI / 32][PG_MMA_WLIST];    // per epilogue warp: the open pairs of the read in hand
    __shared__ unsigned int s_ntie[2], s_wdone[2];   // per accumulator buffer: near-ties so far, epilogue warps done

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
        for (int s = 0; s < PG_MMA_MAXSTAGES; s++) { pgm_mbar_init(pgm_smem(&s_full[s]), PG_MMA_NPROD); pgm_mbar_init(pgm_smem(&s_empty[s]), 1); }
        for (int i = 0; i < 2; i++) { pgm_mbar_init(pgm_smem(&s_dfull[i]), 1); pgm_mbar_init(pgm_smem(&s_dempty[i]), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        s_ntie[0] = s_ntie[1] = 0u;
        s_wdone[0] = s_wdone[1] = 0u;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;
    // shared-window addresses of the barrier arrays, once (a generic-to-shared conversion per use showed up in the profile)
    const uint32_t full0 = pgm_smem(&s_full[0]), empty0 = pgm_smem(&s_empty[0]), dfull0 = pgm_smem(&s_dfull[0]), dempty0 = pgm_smem(&s_dempty[0]);
    const uint32_t sA0 = pgm_smem(sA), ring0 = pgm_smem(ring);

    if (warp > PG_MMA_NEPI / 32) {
        // ================================================================ producers
        // Thread p gathers row jr = p / 2 of every stage; lanes 2i and 2i + 1 fetch the two 16-byte halves of the same
        // 32-byte sector in one instruction (qx: low | high bytes; bm8x: two chunks of 16 blocks).
        const int p = tid - (PG_MMA_NEPI + 32), jr = p >> 1, sub = p & 1;
        constexpr int MAXST = PG_MMA_MAXN / PG_MMA_KC;                       // stages of the longest read
        int cur_n = -1;
        unsigned s = 0, ph = 1u, sprev = 0, phprev = 0;  // ring slot of the next stage and the parity of its "free" phase
        bool first = true;
        PgMmaCursor cur;
        cur.start(a);
        PgMmaRead r, rn, rnn;
        bool have = cur.next(a, r);
        bool have_n = have && cur.next(a, rn);
        // This thread's word id in every stage of a read is fetched TWO reads ahead (the ids stream from DRAM), so that
        // one read ahead the rows of the exact table -- 235 MB and more, mostly not in L2 -- can be prefetched into L2:
        // the gather is bound by requests in flight x their latency, and a DRAM round trip is twice an L2 hit.
        uint32_t w[MAXST], wn[MAXST], wnn[MAXST];
#pragma unroll
        for (int i = 0; i < MAXST; i++) {
            w[i] = (have && i * PG_MMA_KC + jr < r.n) ? (uint32_t)a.words[r.woff + i * PG_MMA_KC + jr] : 0u;
            wn[i] = (have_n && i * PG_MMA_KC + jr < rn.n) ? (uint32_t)a.words[rn.woff + i * PG_MMA_KC + jr] : 0u;
        }
        long long t_wait = 0, t_all = clock64();
        unsigned nst_total = 0;
        const int nb8 = a.pitch8 / 16;                   // chunks of block columns: chunk 2 + i
        while (have) {
            const bool have_nn = have_n && cur.next(a, rnn);
#pragma unroll
            for (int i = 0; i < MAXST; i++) wnn[i] = (have_nn && i * PG_MMA_KC + jr < rnn.n) ? (uint32_t)a.words[rnn.woff + i * PG_MMA_KC + jr] : 0u;
            if (have_n && sub == 0) {
                const uint8_t *qn = a.qx + (size_t)rn.gs * PG_NWORDS * 32;
#pragma unroll
                for (int i = 0; i < MAXST; i++)
                    if (i * PG_MMA_KC + jr < rn.n) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(qn + (size_t)wn[i] * 32));
            }
            if (r.n > 0 && r.n != cur_n) {
                // a new count image: every product that reads the old one must be done, i.e. those of the last stage issued
                if (!first && lane == 0) pgm_mbar_wait((empty0 + 8u * (sprev)), phprev);
                __syncwarp();
                const uint8_t *img = a.images + (size_t)r.n * PG_MMA_IMG;
                const int pieces = 128 * ((r.n + 31) & ~31) / 16;
                for (int i = p; i < pieces; i += PG_MMA_NPROD) PGM_CP16(sA0 + (uint32_t)i * 16u, img + (size_t)i * 16);
                cur_n = r.n;
            }
            const uint8_t *qrow = a.qx + (size_t)r.gs * PG_NWORDS * 32 + sub * 16;
            const int sibc = ((r.gs / PG_PARTS) >> 2) * 16;
#pragma unroll
            for (int kc = 0; kc < MAXST; kc++) {
                if (kc * PG_MMA_KC >= r.n) break;
                const long long tw0 = a.prof ? clock64() : 0;
                if (lane == 0) pgm_mbar_wait((empty0 + 8u * (s)), ph);             // the ring slot is free
                __syncwarp();
                if (a.prof) t_wait += clock64() - tw0;
                if (kc * PG_MMA_KC + jr < r.n) {
                    const uint32_t wid = w[kc];
                    const uint32_t dst0 = ring0 + s * stage_bytes + (uint32_t)jr * 16u;
                    PGM_CP16(dst0 + (uint32_t)sub * (PG_MMA_KC * 16u), qrow + (size_t)wid * 32);
                    const uint8_t *brow = a.bm8x + (size_t)wid * a.pitch8;
                    for (int c = sub; c < nb8; c += 2) PGM_CP16(dst0 + (uint32_t)(2 + c) * (PG_MMA_KC * 16u), brow + c * 16);
                    if (sub == (nb8 & 1)) PGM_CP16(dst0 + (uint32_t)(2 + nb8) * (PG_MMA_KC * 16u), a.hm8x + (size_t)wid * a.hpitch + sibc);
                }
                // This thread's arrival fires when its copies have landed; the thread itself never waits for data, so a
                // whole ring of stages is in flight.  (Waiting per stage -- cp.async.wait_group, then fence.proxy.async --
                // cost 1 700 cycles a stage: the fence waits for every copy in flight, not just the stage being signalled.)
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((full0 + 8u * (s))) : "memory");
                sprev = s; phprev = ph ^ 1u; first = false;
                if (++s == a.nstage) { s = 0; ph ^= 1u; }
                nst_total++;
            }
            r = rn; rn = rnn; have = have_n; have_n = have_nn;
#pragma unroll
            for (int i = 0; i < MAXST; i++) { w[i] = wn[i]; wn[i] = wnn[i]; }
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        if (a.prof && p == 0) { a.prof[blockIdx.x * 16 + 0] = clock64() - t_all; a.prof[blockIdx.x * 16 + 1] = t_wait; a.prof[blockIdx.x * 16 + 7] = nst_total; }
    } else if (warp == PG_MMA_NEPI / 32) {
        // ================================================================ the issuing warp
        // The whole warp walks the loop (so every operand of tcgen05.mma is warp-uniform and lives in uniform
        // registers); one elected lane issues.  With a single lane inside `if (lane == 0)` the compiler wrapped every
        // product in a divergence loop and ~12 register-to-uniform moves: 150 cycles per instruction.
        const uint32_t idesc = (2u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);   // u8 x u8 -> s32, B MN-major, M = 128
        const uint32_t a_lo = ((sA0 >> 4) & 0x3FFFu) | ((2048u >> 4) << 16);               // K chunks 2 048 B apart (leading)
        const uint32_t a_hi = (128u >> 4) | (1u << 14);                                            // row groups 128 B apart (stride); version 1
        const uint32_t b_hi = ((PG_MMA_KC * 16u) >> 4) | (1u << 14);                               // 16-column chunks PG_MMA_KC * 16 B apart
        const uint32_t b_lo0 = ((ring0 >> 4) & 0x3FFFu) | ((128u >> 4) << 16);             // words 8 apart 128 B apart
        unsigned s = 0, ph = 0u, rd = 0;
        PgMmaCursor cur;
        cur.start(a);
        PgMmaRead r;
        long long t_wd = 0, t_wf = 0, t_fence = 0, t_all = clock64();
        while (cur.next(a, r)) {
            if (r.n == 0) continue;                      // skipped read (short, or no word)
            const unsigned acc = rd & 1u;
            long long tw0 = a.prof ? clock64() : 0;
            if (lane == 0) pgm_mbar_wait((dempty0 + 8u * (acc)), ((rd >> 1) & 1u) ^ 1u);          // the epilogue has drained this buffer
            __syncwarp();
            if (a.prof) t_wd += clock64() - tw0;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t dcol = tmem + acc * 256u;
            const int nks = (r.n + 31) >> 5;
            uint32_t alo = a_lo;
            for (int ks0 = 0; ks0 < nks; ks0 += PG_MMA_KC / 32) {
                tw0 = a.prof ? clock64() : 0;
                if (lane == 0) pgm_mbar_wait((full0 + 8u * (s)), ph);
                __syncwarp();
                if (a.prof) t_wf += clock64() - tw0;
                // the rows were written through the generic proxy (cp.async) and are read through the async proxy
                tw0 = a.prof ? clock64() : 0;
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                if (a.prof) t_fence += clock64() - tw0;
                const uint32_t blo = b_lo0 + s * (stage_bytes >> 4);
                const int left = nks - ks0;
                if (pgm_elect()) {
#pragma unroll
                    for (int u = 0; u < PG_MMA_KC / 32; u++)
                        if (u < left)
                            pgm_mma_i8(dcol, ((uint64_t)a_hi << 32) | (alo + (uint32_t)u * (4096u >> 4)),
                                       ((uint64_t)b_hi << 32) | (blo + (uint32_t)u * (512u >> 4)), idesc, (ks0 + u) > 0 ? 1u : 0u);
                    pgm_commit((empty0 + 8u * (s)));                                // the slot is free once these products are done
                    if (ks0 + PG_MMA_KC / 32 >= nks) pgm_commit((dfull0 + 8u * (acc)));
                }
                __syncwarp();
                alo += (PG_MMA_KC / 32) * (4096u >> 4);
                if (++s == a.nstage) { s = 0; ph ^= 1u; }
            }
            rd++;
        }
        if (a.prof && lane == 0) {
            a.prof[blockIdx.x * 16 + 2] = clock64() - t_all; a.prof[blockIdx.x * 16 + 3] = t_wd; a.prof[blockIdx.x * 16 + 4] = t_wf;
            a.prof[blockIdx.x * 16 + 8] = t_fence;
        }
    } else {
        // ================================================================ epilogue: thread = task
        // No CTA-wide barrier: a warp counts its open pairs, reserves room in the global item list from its own
        // reservation (one returning global atomic per ~35 reads) and writes them itself; the four warps of a read meet
        // only in three shared-memory counters (items so far, near-ties so far, warps done), all updated BEFORE the
        // warp hands the accumulator back, so the next read on the same buffer finds them reset.
        // group g = warp / 4 takes the reads whose products go to TMEM buffer g (every other read of the CTA)
        const int t = tid & 127;
        const unsigned grp = (unsigned)(warp >> 2);
        unsigned rd = 0;
        PgMmaCursor cur;
        cur.start(a);
        PgMmaRead r;
        long long t_wd = 0, t_ld = 0, t_cmp = 0, t_ld2 = 0, t_all = clock64();
        unsigned int chunk_pos = 0u, chunk_end = 0u;     // lane 0: this warp's reservation in the global item list
        unsigned int nitems = 0u;                        // lane 0: pairs this warp wrote
        while (cur.next(a, r)) {
            if (r.n == 0) continue;                      // skipped read (short, or no word)
            const unsigned acc = rd & 1u;
            if (PG_MMA_NEPI == 256 && acc != grp) { rd++; continue; }
            const size_t rc = (size_t)a.slot0 + r.slot;
            const int n = r.n;
            int k = n >> 3;
            if (k < a.min_boot) k = a.min_boot;
            const bool live = t == 0 || (t <= PG_NUM_BOOT && k > 0);
            const int best = r.gs / PG_PARTS, own = r.gs % PG_PARTS;
            const uint32_t gbase = (uint32_t)best * 64u, poff = (uint32_t)own * PG_PART_POS;
            const uint32_t vb16 = r.vb16;
            const uint32_t margin = t == 0 ? r.margin_full : r.margin_rep;
            const long long tw0 = a.prof ? clock64() : 0;
            if (lane == 0) pgm_mbar_wait((dfull0 + 8u * (acc)), (rd >> 1) & 1u);     // one lane polls: the polls share the
            __syncwarp();                                                             // shared-memory port with the operands
            if (a.prof) t_wd += clock64() - tw0;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t trow = tmem + acc * 256u + ((uint32_t)((warp & 3) * 32) << 16);
            const int nbc = a.pitch8 / 16;               // chunks of 16 block columns; one more chunk holds the sibling parts
            uint32_t lo[16], hi[16], v0[16], v1[16], v2[16];
            PGM_LD16(trow, lo);
            PGM_LD16(trow + 16u, hi);
            PGM_LD16(trow + 32u, v0);
            PGM_LD16(trow + 48u, v1);
            if (nbc >= 2) PGM_LD16(trow + 64u, v2);      // (warp-uniform)
            PGM_LD_WAIT();
            long long tp = a.prof ? clock64() : 0;
            if (a.prof) t_ld += tp - tw0;
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
            uint32_t tie = 0u;
            if (live) {
                a.champ[rc * (PG_NUM_BOOT + 1) + t] = ((unsigned long long)csum << 32) | (gbase + cpos);
#pragma unroll
                for (int i = 0; i < 16; i++) tie |= (sum[i] <= thr ? 1u : 0u) << i;
                tie &= vb16 & ~(1u << ((cpos - poff) & 31u));
            }
            if (__any_sync(0xffffffffu, tie != 0u)) {
                // near-ties: dense entries of the read's list, positions from a shared counter (this kernel is the
                // read's first writer: the list starts empty)
                const uint32_t mine = __popc(tie);
                uint32_t incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
                uint32_t pos = 0u;
                if (lane == 31) pos = atomicAdd(&s_ntie[acc], incl);
                pos = __shfl_sync(0xffffffffu, pos, 31) + incl - mine;
                while (tie) {
                    const int i = __ffs(tie) - 1;
                    tie &= tie - 1u;
                    uint32_t sv = 0u;
#pragma unroll
                    for (int u = 0; u < 16; u++)
                        if (u == i) sv = sum[u];
                    if (pos < PG_CANDCAP)
                        a.cand[rc * PG_CANDCAP + pos] = ((unsigned long long)t << 56) | ((unsigned long long)(gbase + poff + (uint32_t)i) << 32) | sv;
                    pos++;
                }
            }
            // ---- bound columns: 4 * D <= the smallest genus sum of the block.  Open pairs of up to three chunks at a time:
            // bit i of m[j] = column i of chunk cc + j is within the threshold and names a competitor
            const uint32_t thr8 = thr >> a.shift;
            auto mask16 = [&](const uint32_t (&v)[16], int cc) -> uint32_t {
                // nearly every chunk has no column within the threshold (~1 % of the (task, chunk) pairs do): one
                // minimum and one compare settle those
                const uint32_t m01 = min(min(v[0], v[1]), min(v[2], v[3])), m23 = min(min(v[4], v[5]), min(v[6], v[7]));
                const uint32_t m45 = min(min(v[8], v[9]), min(v[10], v[11])), m67 = min(min(v[12], v[13]), min(v[14], v[15]));
                if (!live || cc > nbc || min(min(m01, m23), min(m45, m67)) > thr8) return 0u;
                uint32_t open = 0u;
#pragma unroll
                for (int i = 0; i < 16; i++) open |= (v[i] <= thr8 ? 1u : 0u) << i;
                if (cc == nbc) {                          // the part minima around the best block: its 3 sibling parts
                    const int q4 = (best & 3) * 4;
                    return open & (0xFu << q4) & ~(1u << (q4 + own));
                }
                const int left = a.ntile64 - cc * 16;     // real blocks in this chunk
                if (left < 16) open &= (1u << (left > 0 ? left : 0)) - 1u;
                if ((best >> 4) == cc) open &= ~(1u << (best & 15));
                return open;
            };
            // Open pairs go to the warp's own list in shared memory first (positions from ballots, no atomics), and from
            // there to the global item list once per read: one reservation, coalesced stores.
            unsigned int wcount = 0u;                    // warp-uniform: pairs of this read so far
            uint32_t *wl = s_wlist[warp];
            auto push = [&](uint32_t m0, uint32_t m1, uint32_t m2, int cc) {
                const uint32_t mine = __popc(m0) + __popc(m1) + __popc(m2);
                const uint32_t tot = __reduce_add_sync(0xffffffffu, mine);
                if (tot == 0u) return;
                // exclusive prefix of `mine` over the lanes: counts are small (mostly 0 or 1): a ballot per level
                uint32_t pos = wcount;
                const uint32_t below = (1u << lane) - 1u;
                for (uint32_t lvl = 1u;; lvl++) {
                    const uint32_t bal = __ballot_sync(0xffffffffu, mine >= lvl);
                    if (bal == 0u) break;
                    pos += __popc(bal & below);
                }
                wcount += tot;
                if (mine == 0u) return;
                const uint32_t mm[3] = {m0, m1, m2};
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    uint32_t m = mm[j];
                    while (m) {
                        const int i = __ffs(m) - 1;
                        m &= m - 1u;
                        const uint32_t field = (cc + j == nbc) ? (0x8000u | ((uint32_t)(i - (best & 3) * 4) << 13) | (uint32_t)best)
                                                               : (uint32_t)((cc + j) * 16 + i);
                        if (pos < PG_MMA_WLIST) wl[pos] = ((uint32_t)t << 16) | field;
                        pos++;
                    }
                }
            };
            push(mask16(v0, 0), mask16(v1, 1), nbc >= 2 ? mask16(v2, 2) : 0u, 0);
            for (int cc = 3; cc <= nbc; cc += 3) {       // large models: three more chunks per round trip
                const long long tl0 = a.prof ? clock64() : 0;
                PGM_LD16(trow + 32u + (uint32_t)cc * 16u, v0);
                if (cc + 1 <= nbc) PGM_LD16(trow + 32u + (uint32_t)(cc + 1) * 16u, v1);
                if (cc + 2 <= nbc) PGM_LD16(trow + 32u + (uint32_t)(cc + 2) * 16u, v2);
                PGM_LD_WAIT();
                if (a.prof) {                             // a consumer of the loaded registers, so the clock is read after they arrive
                    uint32_t sink;
                    asm volatile("add.u32 %0, %1, %2;" : "=r"(sink) : "r"(v0[0] + v0[15]), "r"(v1[0] + v1[15]));
                    t_ld2 += (clock64() - tl0) + (sink == 0xFFFFFFFFu ? 1 : 0);
                }
                push(mask16(v0, cc), mask16(v1, cc + 1), mask16(v2, cc + 2), cc);
            }
            if (wcount) {
                uint32_t base = 0xFFFFFFFFu;
                if (lane == 0) {
                    if (wcount > a.light_max || wcount > PG_MMA_WLIST) a.heavy[rc] = 1;      // too many open pairs: plan 1 redoes the read
                    else {
                        if (wcount > chunk_end - chunk_pos) {
                            // what the old reservation leaves unused is blanked: k_light skips blank entries
                            for (unsigned int i = chunk_pos; i < chunk_end; i++) a.items[i] = (unsigned long long)PG_ITEM_NULL << 16;
                            const unsigned int want = wcount > PG_MMA_RESERVE ? wcount : PG_MMA_RESERVE;
                            chunk_pos = atomicAdd(a.item_count, want);
                            chunk_end = chunk_pos + want;
                            if (chunk_pos > a.item_cap || want > a.item_cap - chunk_pos) {  // the list is full: blank what fits, redo the read
                                for (unsigned int i = chunk_pos; i < a.item_cap && i < chunk_end; i++) a.items[i] = (unsigned long long)PG_ITEM_NULL << 16;
                                chunk_pos = chunk_end = 0u;
                            }
                        }
                        if (wcount <= chunk_end - chunk_pos) { base = chunk_pos; chunk_pos += wcount; nitems += wcount; }
                        else a.heavy[rc] = 1;
                    }
                }
                base = __shfl_sync(0xffffffffu, base, 0);           // (also orders the list's writes before its reads)
                if (base != 0xFFFFFFFFu)
                    for (unsigned int i = lane; i < wcount; i += 32) a.items[base + i] = ((unsigned long long)rc << 32) | wl[i];
                __syncwarp();
            }
            // ---- the last warp of the read publishes the near-tie count and resets the shared counters
            if (lane == 0) {
                __threadfence_block();
                if (atomicAdd(&s_wdone[acc], 1u) == 3u) {
                    a.ncand[rc] = atomicExch(&s_ntie[acc], 0u);      // more than PG_CANDCAP: phase 2 hands the read to the strict kernels
                    s_wdone[acc] = 0u;
                    __threadfence_block();
                }
            }
            __syncwarp();
            if (a.prof) t_cmp += clock64() - tp;
            // the accumulator is read and the counters are settled: hand the buffer back to the issuing thread
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            pgm_mbar_arrive((dempty0 + 8u * (acc)));
            rd++;
        }
        if (lane == 0) {
            for (unsigned int i = chunk_pos; i < chunk_end; i++) a.items[i] = (unsigned long long)PG_ITEM_NULL << 16;
            if (nitems) atomicAdd(a.items_total, nitems);
        }
        if (tid == 0 && a.prof) {
            a.prof[blockIdx.x * 16 + 5] = clock64() - t_all; a.prof[blockIdx.x * 16 + 6] = t_wd;
            a.prof[blockIdx.x * 16 + 9] = t_ld; a.prof[blockIdx.x * 16 + 10] = t_cmp; a.prof[blockIdx.x * 16 + 11] = t_ld2;
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
    a.shift = pg_x8_shift();
    a.words = d_words;                                   // (read offsets, lengths, flags, order, guesses, margins: per slot, k_mma_meta)
    a.nreads_b = (int)nreads_b; a.slot0 = slot0; a.min_boot = min_boot;
    a.champ = cb.champ; a.ncand = cb.ncand; a.cand = cb.cand;
    a.items = cb.items; a.item_count = cb.item_count; a.items_total = cb.counters + 3; a.item_cap = cb.item_cap; a.heavy = cb.heavy; a.light_max = light_max;
    const int nch = 3 + md->pitch8 / 16;
    PG_CUDA(ctx, pg_smem_unlock(ctx, k_mma_bound));
    cudaFuncAttributes fa;
    PG_CUDA(ctx, cudaFuncGetAttributes(&fa, k_mma_bound));
    const size_t room = (size_t)ctx->smem_optin - fa.sharedSizeBytes - 1024, stage = (size_t)nch * PG_MMA_KC * 16;
    size_t ns = (room - (size_t)128 * a.kmax) / stage;
    if (ns > PG_MMA_MAXSTAGES) ns = PG_MMA_MAXSTAGES;
    if (ns < 3) return pg_fail(ctx, PG_EINVAL, "internal: no room for the ring of the tensor-core kernel");
    a.nstage = (unsigned)ns;
    size_t smem = (size_t)128 * a.kmax + ns * stage;
    // the kernel owns all 512 columns of tensor memory: never two CTAs on one SM
    if (smem < (size_t)120 * 1024) smem = (size_t)120 * 1024;
    const unsigned nrun = (nreads_b + PG_MMA_RUN - 1) / PG_MMA_RUN;
    const unsigned grid = nrun < (unsigned)ctx->sm_count ? nrun : (unsigned)ctx->sm_count;
    const int nmeta = (int)nreads_b + (int)(grid + 1) * PG_MMA_RUN;          // the cursors look one slot past the end
    PG_TRY(pg_scratch(ctx, &ctx->s_meta, (size_t)nmeta * 2 * sizeof(int4) + (size_t)grid * 128));
    int4 *d_meta = (int4 *)ctx->s_meta.p;
    k_mma_meta<<<(nmeta + 255) / 256, 256, 0, ctx->stream>>>(d_off, d_nwords, d_flags, d_order, (int)nreads_b, slot0, d_guess, md->d_blockmask,
                                                            min_boot, md->vmax, d_meta, nmeta);
    PG_LAUNCHED(ctx);
    a.meta = d_meta;
    static int env_prof = -1;                            // PG_MMA_PROF=1: per-role cycle counters of every launch on stderr
    if (env_prof < 0) { const char *e = getenv("PG_MMA_PROF"); env_prof = (e && atoi(e)) ? 1 : 0; }
    a.prof = env_prof ? (long long *)(d_meta + 2 * nmeta) : NULL;
    k_mma_bound<<<grid, PG_MMA_THREADS, smem, ctx->stream>>>(a);
    PG_LAUNCHED(ctx);
    if (env_prof) {
        std::vector<long long> h((size_t)grid * 16);
        PG_CUDA(ctx, pg_copy_sync(ctx, h.data(), a.prof, h.size() * 8, cudaMemcpyDeviceToHost));
        double v[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (unsigned c = 0; c < grid; c++)
            for (int i = 0; i < 16; i++) v[i] += (double)h[(size_t)c * 16 + i] / grid;
        fprintf(stderr, "[k_mma_bound] %u reads, %u CTAs, N=%d: cycles per CTA: producer %.0f (waiting for a free slot %.0f, %.0f stages), "
                        "issuer %.0f (waiting for the epilogue %.0f, for the producers %.0f, fences %.0f), epilogue %.0f (waiting for products %.0f incl. in "
                        "loads %.0f, compare and push %.0f of which later loads %.0f)\n",
                nreads_b, grid, nch * 16, v[0], v[1], v[7], v[2], v[3], v[4], v[8], v[5], v[6], v[9], v[10], v[11]);
    }
    return PG_OK;
}
