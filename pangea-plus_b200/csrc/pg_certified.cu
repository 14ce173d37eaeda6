// pg_certified.cu -- certified-margin fast path for K4/K5 (classify mode 1).
//
// Same results as strict mode (and so as the reference order of operations,
// SURVEY.md rows A7-A9), obtained from a fraction of its shared-memory traffic:
//
//  * At training time every table entry is restated as its DEFICIT below the best
//    genus of the same word, D[w][g] = max_g' V[w][g'] - V[w][g] >= 0, quantised
//    downwards to q = floor(D * 128) (12 bits).  For a multiset of draws the row
//    maxima are common to all genera, so  argmax_g sum V  ==  argmin_g sum D.
//  * Integer sums are exact and order-free: with Sq = sum q,
//        Sq/128  <=  sum D  <  (Sq + terms)/128 .
//  * The reference's fp32 running sum F differs from the real sum S by at most
//        E = gamma(terms-1) * terms * max|V|        (gamma(m) = m u / (1 - m u), u = 2^-24).
//    If g* is the reference's winner and b the genus with the smallest Sq, then
//    S(g*) >= F(g*) - E >= F(b) - E >= S(b) - 2E, hence
//        Sq(g*)  <  Sq(b) + terms + 256 E .
//    Every genus outside that margin is PROVABLY not the reference's winner.
//  * Survivors (almost always one) are re-evaluated in strict fp32 order straight
//    from the fp32 table; ties resolve to the lowest genus index as in the reference.
//
//  * Deficits are non-negative, so any UNDER-estimate of a genus's sum that already exceeds
//    champion + margin dismisses the genus: per-block and per-part minima (k_blockmin, k_partmin)
//    bound whole blocks of the lineage-ordered table without touching their rows (k_bound).
//
// Phase 1, three work plans with identical results (pg_classify_opts.cert_plan):
//   plan 3 (default)  k_guess_bm -> k_classify_h (best 16-position part, four reads per CTA) -> k_bound -> k_light
//   plan 2            k_guess_bm -> k_classify_q<.., false> (whole best block)               -> k_bound -> k_light
//   plan 1            k_guess_block -> k_classify_q<.., true> (every block, partial-sum pruning); also the
//                     fallback for "heavy" reads the bounds leave too many open pairs for
//   packed 2 x 16-bit adds, four per LDS.128; a per-(read, task) champion slot is maintained with 64-bit
//   atomicMin and near-ties are appended to a short per-read list.
// Phase 2 (k_resolve): one warp per read; strict re-check of survivors, then the A9 vote.
// Reads whose list overflows are handed back to the strict kernels.
#include "pg_certified.cuh"
#include <algorithm>


// ------------------------------------------------------------------ derive

// one warp per word: maximum over the real genera, and global max |V|
__global__ void k_rowmax(const float *__restrict__ table, int G, int ntile, float *__restrict__ rowmax,
                         unsigned int *__restrict__ vmax_bits)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= PG_NWORDS) return;
    float mx = __int_as_float(0xff800000), amax = 0.f;
    for (int t = 0; t < ntile; t++) {
        int g = t * 32 + lane;
        if (g < G) {
            float v = table[((size_t)t * PG_NWORDS + w) * PG_GENUS_TILE + lane];
            mx = fmaxf(mx, v);
            amax = fmaxf(amax, fabsf(v));
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    }
    if (lane == 0) {
        rowmax[w] = mx;
        atomicMax(vmax_bits, __float_as_uint(amax));      // non-negative floats order as uints
    }
}

// one thread per (tile64, word, position-in-tile): q = floor((rowmax - x) * 128), EXACTLY.  Both operands are fp32
// values of magnitude 0 or in [2^-24, 2^16): such a value is an integer multiple of 2^-47 below 2^63 * 2^-47, so
// v * 2^47 is an exact int64, the difference of two of them is exact, and a right shift by 40 is the floor of the
// difference in units of 1/128.  (The double subtraction this replaces is exact only while the binary exponents of the
// two values differ by less than 29: a word held by nearly every one of ~3 M training sequences has a row maximum of
// -2^-24 next to cells of -16.)  A table value outside that range (never produced by the training formulas) sets the
// out-of-range flag, which switches certified mode off for the model.
__device__ __forceinline__ bool pg_fixed47(float v, long long *out)
{
    const float a = fabsf(v);
    if (!(a == 0.f || (a >= 5.9604644775390625e-08f && a < 65536.f))) return false;
    *out = (long long)((double)v * 140737488355328.0);         // 2^47: exact, |result| < 2^63
    return true;
}
__global__ void k_quantise(const float *__restrict__ table, const float *__restrict__ rowmax,
                           const int32_t *__restrict__ perm, int G, size_t total, uint16_t *__restrict__ q,
                           unsigned int *__restrict__ qmax)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int l = (int)(idx & 63);
    const int w = (int)((idx >> 6) & (PG_NWORDS - 1));
    const int t64 = (int)(idx >> 22);
    const int g = perm[t64 * 64 + l];
    unsigned int v = PG_Q_MAX;
    if (g < G) {
        const float x = table[((size_t)(g >> 5) * PG_NWORDS + w) * PG_GENUS_TILE + (g & 31)];
        long long fr, fx;
        unsigned int u = 0xFFFFFFFFu;                            // out of range: reported through qmax
        if (pg_fixed47(rowmax[w], &fr) && pg_fixed47(x, &fx)) {
            const long long d = (fr - fx) >> 40;                 // >= 0: rowmax is the row's maximum
            u = d > 0xFFFFFFFELL ? 0xFFFFFFFEu : (unsigned int)d;
        }
        atomicMax(qmax, u);
        v = u > PG_Q_MAX ? PG_Q_MAX : u;
    }
    q[idx] = (uint16_t)v;
}

// Lower-bound tables.  bm: per (word, block) the minimum deficit over the block's 64 positions (padding holds
// PG_Q_MAX and never wins).  A group is PG_GB = 28 blocks: bm[group][w][slot], the last PG_PARTS slots of
// every 64-byte row are spare -- k_bound puts the part minima of the best block there with one 8-byte store
// per row (plan 3).
__global__ void k_blockmin(const uint16_t *__restrict__ q, int ntile64, int ngroup, uint16_t *__restrict__ bm)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = (int)(idx & (PG_NWORDS - 1));
    const int gs = (int)(idx >> 16);
    if (gs >= ngroup * 32) return;
    const int grp = gs >> 5, slot = gs & 31;
    const int blk = grp * PG_GB + slot;
    uint32_t m = PG_Q_MAX;
    if (slot < PG_GB && blk < ntile64) {
        m = 0xFFFFu;
        const uint4 *row = reinterpret_cast<const uint4 *>(q + ((size_t)blk * PG_NWORDS + w) * 64);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint4 v = row[i];
            const uint32_t a = __vminu2(__vminu2(v.x, v.y), __vminu2(v.z, v.w));
            m = min(m, min(a & 0xFFFFu, a >> 16));
        }
    }
    bm[((size_t)grp * PG_NWORDS + w) * 32 + slot] = (uint16_t)m;
}

// hm: per (word, part) the minimum over the part's positions; hm[pb / 32][w][pb % 32], pb = PG_PARTS*block + part
// (the parts of one block are adjacent slots of one row)
__global__ void k_partmin(const uint16_t *__restrict__ q, int ntile64, uint16_t *__restrict__ hm)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = (int)(idx & (PG_NWORDS - 1));
    const int blk = (int)(idx >> 16);
    if (blk >= ntile64) return;
    const uint4 *row = reinterpret_cast<const uint4 *>(q + ((size_t)blk * PG_NWORDS + w) * 64);
#pragma unroll
    for (int h = 0; h < PG_PARTS; h++) {
        uint32_t m = 0xFFFFu;
#pragma unroll
        for (int i = 0; i < 8 / PG_PARTS; i++) {
            const uint4 v = row[h * (8 / PG_PARTS) + i];
            const uint32_t a = __vminu2(__vminu2(v.x, v.y), __vminu2(v.z, v.w));
            m = min(m, min(a & 0xFFFFu, a >> 16));
        }
        const int pb = PG_PARTS * blk + h;
        hm[((size_t)(pb >> 5) * PG_NWORDS + w) * 32 + (pb & 31)] = (uint16_t)m;
    }
}

// Coarse first-level bound table of models with more than one group of blocks (k_bound8): one byte per (word, block),
// c = min(bm >> PG_C8_SHIFT, PG_C8_MAX).  Rounded DOWN and capped, so (sum of c) << PG_C8_SHIFT is still a lower bound
// of every genus of the block; 16 draws x PG_C8_MAX < 256, so sixteen rows add up in packed 8-bit fields without a carry.
#define PG_C8_SHIFT 6                    // units of 64 / 128 = 0.5 nat
#define PG_C8_MAX   15u
__global__ void k_bm8(const uint16_t *__restrict__ bm, int ntile64, int pitch, uint8_t *__restrict__ bm8)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int col = (int)(idx % (size_t)pitch);
    const size_t w = idx / (size_t)pitch;
    if (w >= PG_NWORDS) return;
    uint32_t c = PG_C8_MAX;                            // padding columns close at once; sibling-part columns are rewritten per read
    if (col < ntile64) {
        const uint32_t v = bm[((size_t)(col / PG_GB) * PG_NWORDS + w) * 32 + (col % PG_GB)];
        c = min(v >> PG_C8_SHIFT, PG_C8_MAX);
    }
    bm8[idx] = (uint8_t)c;
}

// Table layout of certified mode.  pos_host[p] = genus stored at table position p, or -1 for padding;
// npos is a multiple of 64 (one 64-position block per 128-byte row segment).  NULL = the genus order of
// the training file.  The quantised tables are rebuilt by pg_model_derive_quantised().
int pg_model_set_layout(pg_model *md, const int32_t *pos_host, int npos)
{
    pg_ctx *ctx = md->ctx;
    if (!pos_host) npos = ((md->G + 63) / 64) * 64;
    const int nblk = npos / 64;
    std::vector<int32_t> full((size_t)npos);
    std::vector<unsigned long long> mask((size_t)nblk, 0ULL);
    for (int p = 0; p < npos; p++) {
        const int32_t g = pos_host ? pos_host[p] : (p < md->G ? p : -1);
        full[(size_t)p] = g >= 0 ? g : 0x7FFFFFFF;
        if (g >= 0) mask[(size_t)(p >> 6)] |= 1ULL << (p & 63);
    }
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (nblk != md->ntile64) {                           // a different block count: the tables are re-allocated
        cudaFree(md->d_perm); cudaFree(md->d_blockmask); cudaFree(md->d_qtable); cudaFree(md->d_bmtable); cudaFree(md->d_hmtable);
        cudaFree(md->d_bm8);
        md->d_perm = NULL; md->d_blockmask = NULL; md->d_qtable = NULL; md->d_bmtable = NULL; md->d_hmtable = NULL; md->d_bm8 = NULL;
    }
    md->ntile64 = nblk;
    md->ngroup = (nblk + PG_GB - 1) / PG_GB;
    md->ngroup_h = (PG_PARTS * nblk + 31) / 32;
    md->sib0 = (nblk + 3) & ~3;                                  // the four sibling-part columns start 4-byte aligned
    md->bm8_pitch = (md->sib0 + PG_PARTS + 15) & ~15;
    if (!md->d_perm) {
        PG_CUDA(ctx, cudaMalloc(&md->d_perm, full.size() * 4));
        PG_CUDA(ctx, cudaMalloc(&md->d_blockmask, mask.size() * 8));
    }
    // on the context's (non-blocking) stream, like every consumer; `full` and `mask` are locals: wait before returning
    PG_CUDA(ctx, cudaMemcpyAsync(md->d_perm, full.data(), full.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    PG_CUDA(ctx, cudaMemcpyAsync(md->d_blockmask, mask.data(), mask.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PG_OK;
}

// Lineage-aligned layout: genera in depth-first lineage order, cut into blocks of at most 64 so that a
// clade is split only when it does not fit one block.  The relatives of a read's genus then share its
// block, and the blocks next to it hold other clades whose lower bound (k_bound) dismisses them.
int pg_model_layout_from_lineage(pg_model *md, const int32_t *anc, int depth)
{
    const int G = md->G;
    std::vector<int32_t> order((size_t)G);
    for (int g = 0; g < G; g++) order[(size_t)g] = g;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
        const int32_t *ra = anc + (size_t)a * depth, *rb = anc + (size_t)b * depth;
        for (int d = 0; d < depth; d++)
            if (ra[d] != rb[d]) return ra[d] < rb[d];
        return false;
    });
    // units: maximal clades of at most 64 genera (ranges of `order`), found top-down
    std::vector<std::pair<int, int>> units, stack;
    std::vector<int> level;
    stack.push_back(std::make_pair(0, G));
    level.push_back(0);
    while (!stack.empty()) {
        const std::pair<int, int> r = stack.back();
        const int lv = level.back();
        stack.pop_back();
        level.pop_back();
        if (r.second - r.first <= 64 || lv >= depth) {
            for (int s0 = r.first; s0 < r.second; s0 += 64) units.push_back(std::make_pair(s0, s0 + 64 < r.second ? s0 + 64 : r.second));
            continue;
        }
        // children at this level, pushed in reverse so they pop in order
        std::vector<std::pair<int, int>> kids;
        int i = r.first;
        while (i < r.second) {
            int j = i + 1;
            while (j < r.second && anc[(size_t)order[(size_t)j] * depth + lv] == anc[(size_t)order[(size_t)i] * depth + lv]) j++;
            kids.push_back(std::make_pair(i, j));
            i = j;
        }
        for (size_t c = kids.size(); c-- > 0;) { stack.push_back(kids[c]); level.push_back(lv + 1); }
    }
    std::sort(units.begin(), units.end());
    // greedy packing of consecutive units into blocks of <= 64
    std::vector<int32_t> pos;
    int fill = 0;
    for (const std::pair<int, int> &u : units) {
        const int len = u.second - u.first;
        if (fill + len > 64) {
            pos.resize(((pos.size() + 63) / 64) * 64, -1);
            fill = 0;
        }
        for (int i = u.first; i < u.second; i++) pos.push_back(order[(size_t)i]);
        fill += len;
    }
    pos.resize(((pos.size() + 63) / 64) * 64, -1);
    return pg_model_set_layout(md, pos.data(), (int)pos.size());
}

int pg_mma_build_tables(pg_model *md);               // pg_mma.cu

int pg_model_derive_quantised(pg_model *md)
{
    pg_ctx *ctx = md->ctx;
    md->q_ok = false;
    md->bounds_tuned = false;
    md->part_bounds = false;
    if (!md->d_perm) PG_TRY(pg_model_set_layout(md, NULL, 0));
    const size_t cells = (size_t)md->ntile64 * PG_NWORDS * 64;
    const size_t bmcells = (size_t)md->ngroup * PG_NWORDS * 32;
    if (!md->d_rowmax) PG_CUDA(ctx, cudaMalloc(&md->d_rowmax, PG_NWORDS * 4));
    if (!md->d_qtable) {
        cudaError_t e;
        if ((e = cudaMalloc(&md->d_qtable, cells * 2)) != cudaSuccess ||
            (e = cudaMalloc(&md->d_bmtable, bmcells * 2)) != cudaSuccess ||
            (e = cudaMalloc(&md->d_hmtable, (size_t)md->ngroup_h * PG_NWORDS * 32 * 2)) != cudaSuccess) {
            (void)cudaGetLastError();
            return pg_fail(ctx, PG_ENOMEM, "quantised table allocation failed: %s", cudaGetErrorString(e));
        }
    }
    unsigned int *d_stat = NULL;
    PG_CUDA(ctx, cudaMalloc(&d_stat, 8));
    PG_CUDA(ctx, cudaMemsetAsync(d_stat, 0, 8, ctx->stream));
    k_rowmax<<<PG_NWORDS / 8, 256, 0, ctx->stream>>>(md->d_table, md->G, md->ntile, md->d_rowmax, d_stat);
    PG_LAUNCHED(ctx);
    k_quantise<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->stream>>>(md->d_table, md->d_rowmax, md->d_perm, md->G,
                                                                         cells, md->d_qtable, d_stat + 1);
    PG_LAUNCHED(ctx);
    k_blockmin<<<(unsigned)((bmcells + 255) / 256), 256, 0, ctx->stream>>>(md->d_qtable, md->ntile64, md->ngroup, md->d_bmtable);
    PG_LAUNCHED(ctx);
    PG_CUDA(ctx, cudaMemsetAsync(md->d_hmtable, 0xFF, (size_t)md->ngroup_h * PG_NWORDS * 32 * 2, ctx->stream));
    k_partmin<<<(unsigned)(((size_t)md->ntile64 * PG_NWORDS + 255) / 256), 256, 0, ctx->stream>>>(md->d_qtable, md->ntile64,
                                                                                                  md->d_hmtable);
    PG_LAUNCHED(ctx);
    if (!md->d_bm8) {
        cudaError_t e = cudaMalloc(&md->d_bm8, (size_t)PG_NWORDS * md->bm8_pitch);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return pg_fail(ctx, PG_ENOMEM, "coarse bound table allocation failed: %s", cudaGetErrorString(e)); }
    }
    {
        const size_t c8 = (size_t)PG_NWORDS * md->bm8_pitch;
        k_bm8<<<(unsigned)((c8 + 255) / 256), 256, 0, ctx->stream>>>(md->d_bmtable, md->ntile64, md->bm8_pitch, md->d_bm8);
        PG_LAUNCHED(ctx);
    }
    unsigned int stat[2];
    PG_CUDA(ctx, cudaMemcpyAsync(stat, d_stat, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_stat);
    float vm;
    memcpy(&vm, &stat[0], 4);
    md->vmax = (double)vm;
    md->q_ok = stat[1] <= PG_Q_MAX;      // no entry was clamped: upper bounds hold for every genus
    if (md->q_ok) PG_TRY(pg_mma_build_tables(md));      // byte planes of plan 4 (pg_mma.cu)
    // field widths of the certified bookkeeping: part ids in 12 bits (k_guess_bm), block ids in 13 (items)
    if (PG_PARTS * md->ntile64 >= 4096) md->q_ok = false;       // > ~50 000 genera: mode 1 runs the strict kernels
    return PG_OK;
}

// ------------------------------------------------------------------ shared device pieces


// Task epilogue, split in two so the returning atomic's latency hides behind the
// next replicate's main loop.
//   begin : block minimum as one REDUX over (sum << 6 | genus-in-block), then the group
//           leader posts it to the (read, task) champion slot with a 64-bit atomicMin.
//   finish: with the slot's previous value, append near-ties of the running minimum to
//           the read's list; the displaced champion stays listed if it is still close.
// sums[i] belongs to table position genus0+i, bit i of vbits says whether a genus lives there
// (padding positions never compete); `mask` = the lanes sharing this task.
struct PgPending {
    unsigned long long old;      // leader only: value the slot held before our atomicMin
    uint32_t bkey;               // block minimum: sum << 6 | genus - gbase
    uint32_t lmin;               // this lane's smallest sum
};

template <int NV>
__device__ __forceinline__ void pg_epilogue_begin(unsigned mask, bool leader, const uint32_t *sums, uint32_t genus0,
                                                  uint32_t gbase, uint32_t vbits, unsigned long long *champ_slot,
                                                  PgPending &p)
{
    const uint32_t loff = genus0 - gbase;
    uint32_t k[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) k[i] = (sums[i] << 6) + loff + i;
    if (vbits != (1u << NV) - 1u) {                            // padding positions never compete
#pragma unroll
        for (int i = 0; i < NV; i++)
            if (!((vbits >> i) & 1u)) k[i] = 0xFFFFFFFFu;
    }
    uint32_t lkey = k[0];
#pragma unroll
    for (int i = 1; i < NV; i++) lkey = min(lkey, k[i]);
    p.lmin = lkey >> 6;
    p.bkey = __reduce_min_sync(mask, lkey);
    p.old = 0;
    if (leader)
        p.old = atomicMin(champ_slot, ((unsigned long long)(p.bkey >> 6) << 32) | (gbase + (p.bkey & 63u)));
}

template <int NV>
__device__ __forceinline__ void pg_epilogue_finish(unsigned mask, bool leader, int leader_lane, const uint32_t *sums,
                                                   uint32_t genus0, uint32_t gbase, uint32_t vbits, int task,
                                                   uint32_t margin, const PgPending &p, unsigned int *ncand,
                                                   unsigned long long *cand)
{
    // sums stay below 2^26 (7 000 terms x 4 095), so everything but the slot itself is 32-bit arithmetic
    const unsigned long long old = __shfl_sync(mask, p.old, leader_lane);
    const uint32_t bm = p.bkey >> 6, bg = gbase + (p.bkey & 63u);
    const uint32_t osum = (uint32_t)(old >> 32), opos = (uint32_t)old;       // empty slot: 0xFFFFFFFF / 0xFFFFFFFF
    const bool took = bm < osum || (bm == osum && bg < opos);                 // this block holds the new champion
    const uint32_t thr = (took ? bm : osum) + margin;
    if (p.lmin <= thr) {
        // usually only the champion's own lane gets here, and finds nothing but the champion itself
        uint32_t hit = 0u;
#pragma unroll
        for (int i = 0; i < NV; i++) hit |= (sums[i] <= thr ? 1u : 0u) << i;
        hit &= vbits;
        if (took && bg - genus0 < (uint32_t)NV) hit &= ~(1u << (bg - genus0));
        // a block evaluated a second time meets the
        // standing champion again: it is not its own near-tie
        if (!took && opos - genus0 < (uint32_t)NV) {
#pragma unroll
            for (int i = 0; i < NV; i++)
                if (opos == genus0 + i && sums[i] == osum) hit &= ~(1u << i);
        }
        if (hit) {
#pragma unroll
            for (int i = 0; i < NV; i++)
                if ((hit >> i) & 1u) pg_emit(ncand, cand, task, genus0 + i, sums[i]);
        }
    }
    // the displaced champion stays a candidate if it is within the margin of the new one
    if (leader && took && old != PG_CHAMP_INIT && osum <= bm + margin) pg_emit(ncand, cand, task, opos, osum);
}

// ------------------------------------------------------------------ phase 1

#define PG_QADD4(v) c0 += (v).x; c1 += (v).y; c2 += (v).z; c3 += (v).w;
#define PG_QSPILL()                                                                   \
    s0 += c0 & 0xFFFFu; s1 += c0 >> 16; s2 += c1 & 0xFFFFu; s3 += c1 >> 16;           \
    s4 += c2 & 0xFFFFu; s5 += c2 >> 16; s6 += c3 & 0xFFFFu; s7 += c3 >> 16;           \
    c0 = c1 = c2 = c3 = 0u;

// MINB = CTAs/SM the register budget must allow (a 4- or 5-CTA budget for short reads was measured:
// the spills cost more than the occupancy brings)
template <int BLOCK, int MINB, bool PRUNE>
__global__ void __launch_bounds__(BLOCK, MINB)
k_classify_q(const uint16_t *__restrict__ qtable, const uint16_t *__restrict__ words,
             const int64_t *__restrict__ off, const int32_t *__restrict__ nwords,
             const uint8_t *__restrict__ flags, const int32_t *__restrict__ order, int64_t slot0,
             const uint32_t *__restrict__ boot_pool, const int32_t *__restrict__ boot_off, int min_boot,
             const unsigned long long *__restrict__ blockmask,
             double vmax, unsigned long long *__restrict__ champ, unsigned int *__restrict__ ncand,
             unsigned long long *__restrict__ cand, const int32_t *__restrict__ guess)
{
    constexpr int LPR = 8;                          // lanes per 128-byte row (64 genera x 16 bit)
    constexpr int NGR = (BLOCK - 32) / LPR;
    constexpr int IL = 4;
    extern __shared__ uint4 sQ[];                   // (n+1) rows x 8 uint4; row n is all zero

    const int tid = threadIdx.x;
    const int64_t read = order[blockIdx.x];
    if (flags[2 * read + 1]) return;                // short read (A2)
    const int n = nwords[read];
    if (n == 0) return;                             // no word: phase 2 writes genus 0 directly
    const size_t rc = (size_t)slot0 + blockIdx.x;   // slot = position in the chunk's order array
    // Branch and bound over genus blocks: grid row 0 takes the read's most promising block
    // (k_guess_block) and runs it in full, which seeds the champion slots; every other block
    // then stops a replicate as soon as ALL its 64 partial sums exceed champion + margin --
    // deficits are non-negative, so a partial sum is a lower bound of the final one and no
    // genus of the block can be the winner or a near-tie any more.
    // PRUNE = false (grid of one row): only the guessed block, in full -- the first pass of plan 2.
    int blk = blockIdx.y;
    bool prune_on = false;
    if (guess) {
        const int gs = guess[rc];
        blk = (blockIdx.y == 0) ? gs : (((int)blockIdx.y - 1 < gs) ? (int)blockIdx.y - 1 : (int)blockIdx.y);
        prune_on = PRUNE && blockIdx.y != 0;
    }
    const int gbase = blk * 64;
    const unsigned long long bmask = blockmask[blk];           // which of the block's 64 positions hold a genus
    const uint16_t *tbase = qtable + (size_t)blk * PG_NWORDS * 64;
    const uint16_t *w = words + off[read];
    unsigned long long *mychamp = champ + rc * (PG_NUM_BOOT + 1);
    unsigned int *mync = ncand + rc;
    unsigned long long *mycand = cand + rc * PG_CANDCAP;

    for (int c = tid; c < n * LPR; c += BLOCK) {
        const int r = c / LPR, l = c % LPR;
        pg_cp_async16(&sQ[c], tbase + (size_t)w[r] * 64 + l * 8);
    }
    if (tid < LPR) sQ[n * LPR + tid] = make_uint4(0u, 0u, 0u, 0u);
    pg_cp_async_wait_all();
    __syncthreads();

    if (tid < 32) {
        // ---- full sum (task 0): lane = one packed pair of genera, 16 rows per spill
        const uint32_t *col = reinterpret_cast<const uint32_t *>(sQ) + tid;
        const uint32_t margin_full = pg_margin(n, vmax);
        unsigned long long thr_full = ~0ULL;
        if (prune_on) {
            const unsigned long long c0v = *reinterpret_cast<volatile unsigned long long *>(mychamp);
            if (c0v != PG_CHAMP_INIT) thr_full = (c0v >> 32) + margin_full;
        }
        uint32_t lo = 0u, hi = 0u;
        int j = 0;
        for (; j + 16 <= n; j += 16) {
            uint32_t v[16], c = 0u;
#pragma unroll
            for (int u = 0; u < 16; u++) v[u] = col[(j + u) * 32];
#pragma unroll
            for (int u = 0; u < 16; u++) c += v[u];
            lo += c & 0xFFFFu;
            hi += c >> 16;
            if (prune_on && (unsigned long long)__reduce_min_sync(0xffffffffu, min(lo, hi)) > thr_full) return;
        }
        uint32_t c = 0u;
        for (; j < n; j++) c += col[j * 32];
        lo += c & 0xFFFFu;
        hi += c >> 16;
        const uint32_t sums[2] = {lo, hi};
        PgPending pp;
        const uint32_t vb2 = (uint32_t)(bmask >> (2 * tid)) & 3u;
        pg_epilogue_begin<2>(0xffffffffu, tid == 0, sums, (uint32_t)(gbase + 2 * tid), (uint32_t)gbase, vb2, mychamp, pp);
        pg_epilogue_finish<2>(0xffffffffu, tid == 0, 0, sums, (uint32_t)(gbase + 2 * tid), (uint32_t)gbase, vb2, 0,
                              margin_full, pp, mync, mycand);
        return;
    }

    // ---- replicates (tasks 1..100)
    int k = n >> 3;
    if (k < min_boot) k = min_boot;
    const int nb = (k + 3) >> 2;
    if (nb == 0) return;                            // k == 0: every sum is 0, genus 0 wins (phase 2)
    const int t2 = tid - 32;
    const int group = t2 / LPR, l = t2 % LPR;
    const int lane = tid & 31;
    const unsigned gmask = 0xFFu << ((lane / LPR) * LPR);
    const uint4 *lists = reinterpret_cast<const uint4 *>(boot_pool + boot_off[n]);
    const char *lane_base = reinterpret_cast<const char *>(sQ) + l * 16;
    const uint32_t margin = pg_margin(k, vmax);
#define PG_QROW(o) (*reinterpret_cast<const uint4 *>(lane_base + (o)))

    const int leader_lane = (lane / LPR) * LPR;
    const uint32_t genus0 = (uint32_t)(gbase + l * 8);
    const uint32_t vb8 = (uint32_t)(bmask >> (8 * l)) & 0xFFu;
    PgPending pend;
    uint32_t psum[8];
    int ptask = -1;
#define PG_LD4(d, q) d##0 = PG_QROW((q).x); d##1 = PG_QROW((q).y); d##2 = PG_QROW((q).z); d##3 = PG_QROW((q).w);
#define PG_ADD16(d) PG_QADD4(d##0) PG_QADD4(d##1) PG_QADD4(d##2) PG_QADD4(d##3)
    const int nquad = nb >> 2, tail = nb & 3;
#define PG_LD4P(d, q) if (active) { PG_LD4(d, q) }
    // lower bound of the group's smallest total: spilled part + packed part (min of sums >= sum of mins)
#define PG_PRUNE_CHECK()                                                                        \
    {                                                                                           \
        const uint32_t m2 = __vminu2(__vminu2(c0, c1), __vminu2(c2, c3));                        \
        const uint32_t lb = smin + min(m2 & 0xFFFFu, m2 >> 16);                                  \
        /* the group is pruned when every one of its 8 lanes is above the threshold: one vote */ \
        const unsigned over = __ballot_sync(0xffffffffu, (unsigned long long)lb > thr);          \
        if (((over >> gshift) & 0xFFu) == 0xFFu) active = false;                                 \
    }
    const int gshift = (lane / LPR) * LPR;
    unsigned long long champ_next = PG_CHAMP_INIT;
    if (prune_on && group < PG_NUM_BOOT) champ_next = *reinterpret_cast<volatile unsigned long long *>(mychamp + 1 + group);
    for (int task = group; task < PG_NUM_BOOT; task += NGR) {
        uint32_t c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;
        uint32_t s0 = 0u, s1 = 0u, s2 = 0u, s3 = 0u, s4 = 0u, s5 = 0u, s6 = 0u, s7 = 0u;
        uint32_t smin = 0u;
        bool active = true;
        // champion of this replicate as of a moment ago (stale values are larger: conservative)
        const unsigned long long thr = (champ_next == PG_CHAMP_INIT) ? ~0ULL : (champ_next >> 32) + margin;
        if (prune_on && task + NGR < PG_NUM_BOOT)
            champ_next = *reinterpret_cast<volatile unsigned long long *>(mychamp + 1 + task + NGR);
        const uint4 *lp = lists + (size_t)(task / IL) * nb * IL + (task % IL);
        uint4 qa = __ldg(lp), qb = __ldg(lp + IL);
        uint4 x0, x1, x2, x3, y0, y1, y2, y3;
        x0 = x1 = x2 = x3 = y0 = y1 = y2 = y3 = make_uint4(0u, 0u, 0u, 0u);
        PG_LD4(x, qa)
        // four 4-draw batches per trip: loads run one batch ahead of the adds, list reads two;
        // 16 rows x 4095 < 2^16, so the packed accumulators are spilled once per trip.
        // The four groups of a warp walk in lockstep; a pruned group stops issuing LDS.
        int b = 0;
        for (int it = 0; it < nquad; it++, b += 4) {
            PG_LD4P(y, qb) qa = __ldg(lp + (b + 2) * IL); PG_ADD16(x)
            PG_LD4P(x, qa) qb = __ldg(lp + (b + 3) * IL); PG_ADD16(y)
            if (prune_on) {
                PG_PRUNE_CHECK()
                if (!__any_sync(0xffffffffu, active)) break;
            }
            PG_LD4P(y, qb) qa = __ldg(lp + (b + 4) * IL); PG_ADD16(x)
            if (b + 4 < nb) { PG_LD4P(x, qa) }
            qb = __ldg(lp + (b + 5) * IL);
            PG_ADD16(y)
            PG_QSPILL()
            if (prune_on) {
                smin = min(min(min(s0, s1), min(s2, s3)), min(min(s4, s5), min(s6, s7)));
                PG_PRUNE_CHECK()
                if (!__any_sync(0xffffffffu, active)) break;
            }
        }
        if (tail && __any_sync(0xffffffffu, active)) {   // x holds batch b, qb the list entry of batch b+1
            if (tail >= 2) { PG_LD4P(y, qb) qa = __ldg(lp + (b + 2) * IL); }
            PG_ADD16(x)
            if (tail >= 2) {
                if (tail == 3) { PG_LD4P(x, qa) }
                PG_ADD16(y)
                if (tail == 3) { PG_ADD16(x) }
            }
            PG_QSPILL()
        }
        // the previous replicate's atomic has had a whole main loop to come back
        if (ptask >= 0)
            pg_epilogue_finish<8>(gmask, l == 0, leader_lane, psum, genus0, (uint32_t)gbase, vb8, 1 + ptask, margin, pend,
                                  mync, mycand);
        ptask = -1;
        if (active) {                                   // not pruned: this block may hold the winner or a near-tie
            psum[0] = s0; psum[1] = s1; psum[2] = s2; psum[3] = s3;
            psum[4] = s4; psum[5] = s5; psum[6] = s6; psum[7] = s7;
            ptask = task;
            pg_epilogue_begin<8>(gmask, l == 0, psum, genus0, (uint32_t)gbase, vb8, mychamp + 1 + task, pend);
        }
    }
    if (ptask >= 0)
        pg_epilogue_finish<8>(gmask, l == 0, leader_lane, psum, genus0, (uint32_t)gbase, vb8, 1 + ptask, margin, pend, mync,
                              mycand);
#undef PG_LD4P
#undef PG_PRUNE_CHECK
#undef PG_QROW
}

// ------------------------------------------------------------------ plan 3: one part of the best block, PG_PARTS reads per CTA
//
// k_classify_q spends one 128-byte shared-memory wavefront per (task, draw): a row of 64 positions.  Most of
// those 64 are irrelevant even inside the best block, so plan 3 evaluates only the best PART of it
// (PG_PART_POS = 16 positions, a 32-byte row) and leaves the other parts to the lower bounds like any other
// block (k_bound gets their minima in the spare slots of its rows).  PG_PARTS reads share a CTA and interleave
// their rows -- row r = [read 0: 32 bytes | read 1 | read 2 | read 3] -- so that the groups of a quarter-warp
// (the same task of the four reads) always hit different banks: one wavefront serves four (task, draw) pairs,
// and reads of equal length (adjacent in the order array) load the same sample-list entries.
template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_classify_h(const uint16_t *__restrict__ qtable, const uint16_t *__restrict__ words,
             const int64_t *__restrict__ off, const int32_t *__restrict__ nwords,
             const uint8_t *__restrict__ flags, const int32_t *__restrict__ order, int nreads_b, int64_t slot0,
             const uint32_t *__restrict__ boot_pool, const int32_t *__restrict__ boot_off, int min_boot,
             const unsigned long long *__restrict__ blockmask, double vmax,
             unsigned long long *__restrict__ champ, unsigned int *__restrict__ ncand,
             unsigned long long *__restrict__ cand, const int32_t *__restrict__ guess /* part ids */)
{
    constexpr int R = PG_PARTS;                     // reads per CTA
    constexpr int GL = 8 / R;                       // lanes per (read, task) group: GL x 8 positions
    constexpr int NGR = (BLOCK - 32) / 8;           // task slots per read: a quarter-warp = one slot of every read
    constexpr int FL = 32 / R;                      // full-sum lanes per read
    extern __shared__ uint4 sQ[];                   // (nmax+1) rows x 8 uint4: GL per read; row n of a read is zero
    const int tid = threadIdx.x, lane = tid & 31;

    // ---- the reads of this CTA (some may be missing, short or wordless: n = 0); this thread's read = sel
    const int sel = tid < 32 ? lane / FL : (lane / GL) % R;
    int n = 0, pb = 0, nmaxr = 0;
    size_t rc = 0;
#pragma unroll
    for (int s = 0; s < R; s++) {
        const int slot = R * (int)blockIdx.x + s;
        int ns = 0, ps = 0;
        const uint16_t *ws = words;
        size_t rcs = 0;
        if (slot < nreads_b) {
            const int64_t read = order[slot];
            if (!flags[2 * read + 1]) {
                ns = nwords[read];
                ws = words + off[read];
                rcs = (size_t)slot0 + slot;
                ps = guess[rcs];
            }
        }
        // stage read s: its part of the best block, GL 16-byte chunks per row
        const uint16_t *tb = qtable + (size_t)(ps / R) * PG_NWORDS * 64 + (ps % R) * PG_PART_POS;
        for (int c = tid; c < ns * GL; c += BLOCK) {
            const int r = c / GL, l = c % GL;
            pg_cp_async16(&sQ[r * 8 + s * GL + l], tb + (size_t)ws[r] * 64 + l * 8);
        }
        if (tid < GL) sQ[ns * 8 + s * GL + tid] = make_uint4(0u, 0u, 0u, 0u);
        if (s == sel) { n = ns; pb = ps; rc = rcs; }
        nmaxr = ns > nmaxr ? ns : nmaxr;
    }
    if (nmaxr == 0) return;
    pg_cp_async_wait_all();
    __syncthreads();
    const uint32_t gbase = (uint32_t)(pb / R) * 64u, pbase = gbase + (uint32_t)(pb % R) * PG_PART_POS;
    const unsigned long long bmask = blockmask[pb / R];
    unsigned long long *mychamp = champ + rc * (PG_NUM_BOOT + 1);
    unsigned int *mync = ncand + rc;
    unsigned long long *mycand = cand + rc * PG_CANDCAP;

    if (tid < 32) {
        // ---- full sums (task 0): FL lanes per read, lane = one packed pair of positions = word `lane` of the row
        if (n == 0) return;
        const int hl = lane % FL;
        const unsigned hmask = (FL == 32 ? 0xFFFFFFFFu : ((1u << FL) - 1u)) << (sel * FL);
        const uint32_t *col = reinterpret_cast<const uint32_t *>(sQ) + lane;
        uint32_t lo = 0u, hi = 0u;
        int j = 0;
        for (; j + 16 <= n; j += 16) {
            uint32_t c = 0u;
#pragma unroll
            for (int u = 0; u < 16; u++) c += col[(j + u) * 32];
            lo += c & 0xFFFFu;
            hi += c >> 16;
        }
        uint32_t c = 0u;
        for (; j < n; j++) c += col[j * 32];
        lo += c & 0xFFFFu;
        hi += c >> 16;
        const uint32_t sums[2] = {lo, hi};
        const uint32_t g0 = pbase + 2u * hl;
        const uint32_t vb2 = (uint32_t)(bmask >> (g0 - gbase)) & 3u;
        PgPending pp;
        pg_epilogue_begin<2>(hmask, hl == 0, sums, g0, gbase, vb2, mychamp, pp);
        pg_epilogue_finish<2>(hmask, hl == 0, sel * FL, sums, g0, gbase, vb2, 0, pg_margin(n, vmax), pp, mync, mycand);
        return;
    }

    // ---- replicates (tasks 1..100): GL lanes per (read, task); quarter-warp = the task slot, groups = the reads
    const int lg = lane % GL;
    const int group = ((tid - 32) >> 5) * 4 + (lane >> 3);
    int k = n >> 3;
    if (k < min_boot) k = min_boot;
    const int nb = n > 0 ? (k + 3) >> 2 : 0;
    // the reads of a CTA walk the same number of batches (a finished one just stops loading)
    int kmax = nmaxr >> 3;
    if (kmax < min_boot) kmax = min_boot;
    const int nbmax = (kmax + 3) >> 2;
    if (nbmax == 0) return;
    const unsigned gmask = ((1u << GL) - 1u) << (lane - lg);
    const uint4 *lists = reinterpret_cast<const uint4 *>(boot_pool + boot_off[n]);
    const char *lane_base = reinterpret_cast<const char *>(sQ) + sel * (GL * 16) + lg * 16;
    const uint32_t margin = pg_margin(k, vmax);
    const uint32_t genus0 = pbase + 8u * lg;
    const uint32_t vb8 = (uint32_t)(bmask >> (genus0 - gbase)) & 0xFFu;
    const uint32_t zoff = (uint32_t)n * PG_ROW_PITCH;
    const uint4 zq = make_uint4(zoff, zoff, zoff, zoff);
#define PG_HROW(o) (*reinterpret_cast<const uint4 *>(lane_base + (o)))
#define PG_HLQ(bb) ((bb) < nb ? __ldg(lp + (bb) * 4) : zq)
    PgPending pend;
    uint32_t psum[8];
    int ptask = -1;
    for (int task = group; task < PG_NUM_BOOT; task += NGR) {
        uint32_t c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;
        uint32_t s0 = 0u, s1 = 0u, s2 = 0u, s3 = 0u, s4 = 0u, s5 = 0u, s6 = 0u, s7 = 0u;
        const uint4 *lp = lists + (size_t)(task >> 2) * nb * 4 + (task & 3);
        // list entries one trip (four batches) ahead of the row loads; batches past this read's end are the zero row
        uint4 n0 = PG_HLQ(0), n1 = PG_HLQ(1), n2q = PG_HLQ(2), n3 = PG_HLQ(3);
        for (int b = 0; b < nbmax; b += 4) {
            const uint4 q0 = n0, q1 = n1, q2 = n2q, q3 = n3;
            if (b + 4 < nbmax) { n0 = PG_HLQ(b + 4); n1 = PG_HLQ(b + 5); n2q = PG_HLQ(b + 6); n3 = PG_HLQ(b + 7); }
            if (b < nb) {
                uint4 x0, x1, x2, x3;
                x0 = PG_HROW(q0.x); x1 = PG_HROW(q0.y); x2 = PG_HROW(q0.z); x3 = PG_HROW(q0.w);
                PG_QADD4(x0) PG_QADD4(x1) PG_QADD4(x2) PG_QADD4(x3)
                x0 = PG_HROW(q1.x); x1 = PG_HROW(q1.y); x2 = PG_HROW(q1.z); x3 = PG_HROW(q1.w);
                PG_QADD4(x0) PG_QADD4(x1) PG_QADD4(x2) PG_QADD4(x3)
                x0 = PG_HROW(q2.x); x1 = PG_HROW(q2.y); x2 = PG_HROW(q2.z); x3 = PG_HROW(q2.w);
                PG_QADD4(x0) PG_QADD4(x1) PG_QADD4(x2) PG_QADD4(x3)
                x0 = PG_HROW(q3.x); x1 = PG_HROW(q3.y); x2 = PG_HROW(q3.z); x3 = PG_HROW(q3.w);
                PG_QADD4(x0) PG_QADD4(x1) PG_QADD4(x2) PG_QADD4(x3)
                PG_QSPILL()                          // 16 rows x 4095 < 2^16
            }
        }
        if (nb == 0) continue;                          // k == 0: every sum is 0, genus 0 wins (phase 2)
        // the previous replicate's atomic has had a whole main loop to come back
        if (ptask >= 0)
            pg_epilogue_finish<8>(gmask, lg == 0, lane - lg, psum, genus0, gbase, vb8, 1 + ptask, margin, pend, mync, mycand);
        psum[0] = s0; psum[1] = s1; psum[2] = s2; psum[3] = s3;
        psum[4] = s4; psum[5] = s5; psum[6] = s6; psum[7] = s7;
        ptask = task;
        pg_epilogue_begin<8>(gmask, lg == 0, psum, genus0, gbase, vb8, mychamp + 1 + task, pend);
    }
    if (ptask >= 0)
        pg_epilogue_finish<8>(gmask, lg == 0, lane - lg, psum, genus0, gbase, vb8, 1 + ptask, margin, pend, mync, mycand);
#undef PG_HROW
#undef PG_HLQ
}

// ------------------------------------------------------------------ phase 0: most promising block

// One warp per read: deficit sums over 16 evenly spaced words for every genus block; the block
// holding the smallest sum is processed first and in full.  A wrong guess costs pruning
// efficiency only, never correctness.
__global__ void __launch_bounds__(256)
k_guess_block(const uint16_t *__restrict__ qtable, const uint16_t *__restrict__ words,
              const int64_t *__restrict__ off, const int32_t *__restrict__ nwords,
              const int32_t *__restrict__ order, int nreads_b, int64_t slot0, int ntile64, int32_t *__restrict__ guess)
{
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (slot >= nreads_b) return;
    const int64_t read = order[slot];
    const int n = nwords[read];
    int best = 0;
    if (n > 0) {
        const uint16_t *w = words + off[read];
        const int ns = n < 16 ? n : 16, stride = n / ns;
        uint32_t wv[16];
#pragma unroll
        for (int j = 0; j < 16; j++) wv[j] = j < ns ? (uint32_t)__ldg(w + j * stride) : 0u;
        const uint32_t *q32 = reinterpret_cast<const uint32_t *>(qtable);
        uint32_t bestv = 0xFFFFFFFFu;
        for (int b = 0; b < ntile64; b++) {
            uint32_t acc = 0u;                           // 16 rows x 4095 < 2^16: no carry between the halves
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (j < ns) acc += __ldg(q32 + ((size_t)b * PG_NWORDS + wv[j]) * 32 + lane);
            const uint32_t v = __reduce_min_sync(0xffffffffu, min(acc & 0xFFFFu, acc >> 16));
            if (v < bestv) { bestv = v; best = b; }
        }
    }
    if (lane == 0) guess[slot0 + slot] = best;
}

// ------------------------------------------------------------------ certified v2: block lower bounds
//
// The all-block kernel above spends most of its time proving, draw by draw, that a genus block
// does NOT hold a replicate's winner.  v2 proves it from a 64-times smaller table instead:
//   bm[w][blk] = min over the 64 genera of block blk of q[w][g]      (k_blockmin)
//   LB(task, blk) = sum over the task's draws of bm[w_draw][blk]  <=  Sq(g)  for every g in blk.
// If LB > champion + margin no genus of the block is the winner or a near-tie (same argument as
// the partial-sum prune).  With the genera laid out in lineage order (pg_model_set_lineage) the
// relatives of the read's genus share its block, and on the bench workload 98 % of the
// (replicate, block) pairs are dismissed by the bound; the survivors ("items") are evaluated
// exactly by k_light straight from L2.  Order of one bucket:
//   k_guess_bm   warp per read: 32 sampled words x bm -> the block to run in full
//   k_classify_q grid (reads, 1): that block in full; seeds the champion slots
//   k_bound      CTA per (read, group of 32 blocks): LB of the 101 tasks -> items / heavy flag
//   k_light      group of 8 lanes per item: exact sums of the item's block, champion update
//   k_resolve    as before; reads flagged heavy (too many items) are redone by the all-block kernel.


__global__ void __launch_bounds__(256)
k_guess_bm(const uint16_t *__restrict__ bm, const uint16_t *__restrict__ words, const int64_t *__restrict__ off,
           const int32_t *__restrict__ nwords, const int32_t *__restrict__ order, int nreads_b, int64_t slot0,
           int count, int per_group, int ngroup, int32_t *__restrict__ guess, int nsample /* 32, or 64 for a second try */)
{
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (slot >= nreads_b) return;
    const int64_t read = order[slot];
    const int n = nwords[read];
    uint32_t best = 0u;
    if (n > 0) {
        const uint16_t *w = words + off[read];
        const int ns = n < nsample ? n : nsample, stride = n / ns;
        const uint32_t wv = lane < ns ? (uint32_t)__ldg(w + lane * stride) : 0u;
        const uint32_t wv2 = lane + 32 < ns ? (uint32_t)__ldg(w + (lane + 32) * stride) : 0u;
        uint32_t bestkey = 0xFFFFFFFFu;
        for (int grp = 0; grp < ngroup; grp++) {
            uint32_t acc = 0u;
            for (int j = 0; j < ns; j++) {
                const uint32_t wj = __shfl_sync(0xffffffffu, j < 32 ? wv : wv2, j & 31);
                acc += __ldg(bm + ((size_t)grp * PG_NWORDS + wj) * 32 + lane);
            }
            // the table holds `count` units (blocks: PG_GB per group; parts: 32 per group)
            const uint32_t blk = (uint32_t)grp * (uint32_t)per_group + lane;
            const uint32_t key = ((int)lane < per_group && (int)blk < count) ? ((acc << 12) | blk) : 0xFFFFFFFFu;   // acc < 2^18, blk < 2^12
            bestkey = min(bestkey, __reduce_min_sync(0xffffffffu, key));
        }
        best = bestkey & 0xFFFu;
    }
    if (lane == 0) guess[slot0 + slot] = (int32_t)best;
}

// PART = true: the columns are the 16-position PARTS of the table (hm rows, 32 parts per group) instead of its 64-position
// blocks.  Four times the columns, but a part's minimum is much larger than its block's when every genus holds
// thousands of once-seen words (a training set with hundreds of members per genus): there the block bounds are too
// weak to dismiss far blocks and the part bounds are not.  The open pairs come out as part items.
template <int BLOCK, bool PART>
__global__ void __launch_bounds__(BLOCK)      // (a 64-register budget for 6 CTAs/SM was measured: no gain)
k_bound(const uint16_t *__restrict__ bm, const uint16_t *__restrict__ words, const int64_t *__restrict__ off,
        const int32_t *__restrict__ nwords, const uint8_t *__restrict__ flags, const int32_t *__restrict__ order,
        int64_t slot0, const uint32_t *__restrict__ boot_pool, const int32_t *__restrict__ boot_off, int min_boot,
        int ntile64, double vmax, const unsigned long long *__restrict__ champ, const int32_t *__restrict__ guess,
        unsigned long long *__restrict__ items, unsigned int *__restrict__ item_count, unsigned int item_cap,
        uint8_t *__restrict__ heavy, unsigned int light_max, const uint16_t *__restrict__ hm /* plan 3: guess[] holds part ids */)
{
    constexpr int NUNIT = BLOCK / 8;                // quarter-warps: one task each, lane = four blocks (one LDS.64 per draw)
    extern __shared__ uint4 sB[];                   // (n+1) rows x 4 uint4 (32 blocks x 16 bit); row n is zero
    __shared__ uint32_t s_list[PG_LIGHT_MAX];
    __shared__ uint32_t s_full[32];
    __shared__ uint32_t s_thr[PG_NUM_BOOT];         // champion + margin of the 100 replicates, staged once (a dependent global
                                                    // load at the head of every task was 8 % of the kernel's stall samples)
    __shared__ unsigned int s_cnt, s_base;

    const int tid = threadIdx.x;
    const int64_t read = order[blockIdx.x];
    if (flags[2 * read + 1]) return;
    const int n = nwords[read];
    if (n == 0) return;
    const size_t rc = (size_t)slot0 + blockIdx.x;
    const int grp = blockIdx.y;
    const uint16_t *w = words + off[read];
    const uint16_t *tb = bm + (size_t)grp * PG_NWORDS * 32;
    for (int c = tid; c < n * 4; c += BLOCK) pg_cp_async16(&sB[c], tb + (size_t)w[c >> 2] * 32 + (c & 3) * 8);
    // plan 3: the best block was evaluated on one part only; its other parts compete here like blocks of their
    // own, their minima written into the spare slots PG_GB..31 of the rows of the best block's group (gathered
    // while the cp.async copies are in flight, stored once they have landed)
    const int gs = guess[rc];
    const int best = hm ? gs / PG_PARTS : gs, own = hm ? gs % PG_PARTS : 0;
    const bool sib_here = !PART && hm && best / PG_GB == grp;
    const uint16_t *hrow = hm ? hm + (size_t)((best * PG_PARTS) >> 5) * PG_NWORDS * 32 + ((best * PG_PARTS) & 31) : NULL;
    uint2 hv[4];
    if (sib_here) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = tid + u * BLOCK;
            hv[u] = j < n ? __ldg(reinterpret_cast<const uint2 *>(hrow + (size_t)w[j] * 32)) : make_uint2(0u, 0u);
        }
    }
    if (tid < 4) sB[n * 4 + tid] = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 32) s_full[tid] = 0u;
    if (tid == 0) s_cnt = 0u;
    {
        int k0 = n >> 3;
        if (k0 < min_boot) k0 = min_boot;
        const uint32_t margin0 = pg_margin(k0, vmax);
        for (int t = tid; t < PG_NUM_BOOT; t += BLOCK) {
            const unsigned long long cv = __ldg(champ + rc * (PG_NUM_BOOT + 1) + 1 + t);
            // sums stay below 2^26: the threshold is kept in 32 bits (an empty slot = no threshold)
            s_thr[t] = (cv == PG_CHAMP_INIT) ? 0xFFFFFFFFu : (uint32_t)(cv >> 32) + margin0;
        }
    }
    pg_cp_async_wait_all();
    __syncthreads();
    if (sib_here) {
        uint16_t *rows16 = reinterpret_cast<uint16_t *>(sB);
        // slots PG_GB .. 31 of row j <- the block's four part minima (the read's own part is never tested); one
        // 8-byte store per row: a column of a row-major array is a 16-way bank conflict whatever the width
        auto put = [&](int j, uint2 v) { *reinterpret_cast<uint2 *>(rows16 + j * 32 + PG_GB) = v; };
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (tid + u * BLOCK < n) put(tid + u * BLOCK, hv[u]);
        for (int j = tid + 4 * BLOCK; j < n; j += BLOCK) put(j, __ldg(reinterpret_cast<const uint2 *>(hrow + (size_t)w[j] * 32)));
        __syncthreads();
    }

    const int unit = tid >> 3, hl = tid & 7;
    const int b0 = grp * (PART ? 32 : PG_GB) + 4 * hl;
    bool ok[4];
    int okblk[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int slot = 4 * hl + i;
        if (PART) {                                 // column = part b0 + i of the table; the read's own part is never tested
            const int p = b0 + i;
            ok[i] = p < PG_PARTS * ntile64 && p != gs;
            okblk[i] = 0x8000 | ((p % PG_PARTS) << 13) | (p / PG_PARTS);
            continue;
        }
        okblk[i] = b0 + i;
        ok[i] = slot < PG_GB && b0 + i < ntile64 && b0 + i != best;
        if (slot >= PG_GB) {                        // part item: that sibling part of the best block only
            const int part = slot - PG_GB;
            ok[i] = sib_here && part != own;
            okblk[i] = 0x8000 | (part << 13) | best;
        }
    }
    const char *base = reinterpret_cast<const char *>(sB) + hl * 8;
    const unsigned long long *mychamp = champ + rc * (PG_NUM_BOOT + 1);
#define PG_BROW(o) (*reinterpret_cast<const uint2 *>(base + (o)))
#define PG_BADD(o) { const uint2 v_ = PG_BROW(o); cx += v_.x; cy += v_.y; }
#define PG_BSPILL() { s[0] += cx & 0xFFFFu; s[1] += cx >> 16; s[2] += cy & 0xFFFFu; s[3] += cy >> 16; cx = cy = 0u; }
#define PG_SURVIVE(task, blk)                                                  \
    {                                                                          \
        const unsigned int pos = atomicAdd(&s_cnt, 1u);                        \
        if (pos < PG_LIGHT_MAX) s_list[pos] = ((uint32_t)(task) << 16) | (uint32_t)(blk); \
    }

    // ---- task 0 (full sum): the units split the rows, shared-memory atomics combine them
    {
        uint32_t s[4] = {0u, 0u, 0u, 0u}, cx = 0u, cy = 0u;
        int cnt = 0;
        for (int j = unit; j < n; j += NUNIT) {
            PG_BADD(j * 64)
            if (++cnt == 16) { PG_BSPILL() cnt = 0; }
        }
        PG_BSPILL()
#pragma unroll
        for (int i = 0; i < 4; i++) atomicAdd(&s_full[4 * hl + i], s[i]);
    }

    // ---- tasks 1..100: one replicate per quarter-warp
    int k = n >> 3;
    if (k < min_boot) k = min_boot;
    const int nb = (k + 3) >> 2;
    if (nb > 0) {
        const uint4 *lists = reinterpret_cast<const uint4 *>(boot_pool + boot_off[n]);
        for (int t = unit; t < PG_NUM_BOOT; t += NUNIT) {
            const uint4 *lp = lists + (size_t)(t >> 2) * nb * 4 + (t & 3);
            const uint32_t thr = s_thr[t];
            uint32_t s[4] = {0u, 0u, 0u, 0u}, cx = 0u, cy = 0u;
            bool lane_open = true;                  // some block of this lane may still be within the threshold
            // The four units of a warp walk in lockstep.  Partial sums only grow, so a lane whose four blocks are
            // all above the threshold stops loading (its frozen sums stay above it), and once every lane of the
            // warp has stopped the four tasks are settled.  The list entries run one trip (four batches = 16
            // draws) ahead of the loads that need them, in two register sets used alternately; batches past the
            // task's end point at the zero row.  (A check every 8 draws was measured: no gain.)
            const uint32_t zoff = (uint32_t)n * PG_ROW_PITCH;
            const uint4 zq = make_uint4(zoff, zoff, zoff, zoff);
#define PG_LQ(bb) ((bb) < nb ? __ldg(lp + (bb) * 4) : zq)
#define PG_LQ4(d, bb)                                                                                      \
    if ((bb) + 4 <= nb) { d##0 = __ldg(lp + ((bb) + 0) * 4); d##1 = __ldg(lp + ((bb) + 1) * 4);            \
                          d##2 = __ldg(lp + ((bb) + 2) * 4); d##3 = __ldg(lp + ((bb) + 3) * 4); }          \
    else { d##0 = PG_LQ((bb) + 0); d##1 = PG_LQ((bb) + 1); d##2 = PG_LQ((bb) + 2); d##3 = PG_LQ((bb) + 3); }
#define PG_TRIP(d)                                                                                         \
    if (lane_open) {                                                                                       \
        PG_BADD((d##0).x >> 1) PG_BADD((d##0).y >> 1) PG_BADD((d##0).z >> 1) PG_BADD((d##0).w >> 1)                \
        PG_BADD((d##1).x >> 1) PG_BADD((d##1).y >> 1) PG_BADD((d##1).z >> 1) PG_BADD((d##1).w >> 1)                \
        PG_BADD((d##2).x >> 1) PG_BADD((d##2).y >> 1) PG_BADD((d##2).z >> 1) PG_BADD((d##2).w >> 1)                \
        PG_BADD((d##3).x >> 1) PG_BADD((d##3).y >> 1) PG_BADD((d##3).z >> 1) PG_BADD((d##3).w >> 1)                \
        PG_BSPILL()                                 /* 16 rows x 4095 < 2^16: one spill per trip */        \
    }                                                                                                      \
    {                                                                                                      \
        bool mine = false;                                                                                 \
        _Pragma("unroll") for (int i = 0; i < 4; i++) mine |= ok[i] && s[i] <= thr;                        \
        lane_open = mine && lane_open;                                                                     \
    }
            uint4 A0, A1, A2, A3, B0, B1, B2, B3;
            B0 = B1 = B2 = B3 = zq;
            PG_LQ4(A, 0)
            for (int b = 0; b < nb; b += 8) {
                if (lane_open && b + 4 < nb) { PG_LQ4(B, b + 4) }
                PG_TRIP(A)
                if (__ballot_sync(0xffffffffu, lane_open) == 0u || b + 4 >= nb) break;
                if (lane_open && b + 8 < nb) { PG_LQ4(A, b + 8) }
                PG_TRIP(B)
                if (__ballot_sync(0xffffffffu, lane_open) == 0u) break;
            }
#undef PG_LQ
#undef PG_LQ4
#undef PG_TRIP
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (lane_open && ok[i] && s[i] <= thr) PG_SURVIVE(1 + t, okblk[i])
        }
    }
    __syncthreads();
    if (tid < 32) {
        const int blk = grp * PG_GB + tid;
        const unsigned long long cv = __ldg(mychamp);
        const unsigned long long thr = (cv == PG_CHAMP_INIT) ? ~0ULL : (cv >> 32) + pg_margin(n, vmax);
        if (PART) {
            const int p = grp * 32 + tid;
            if (p < PG_PARTS * ntile64 && p != gs && (unsigned long long)s_full[tid] <= thr)
                PG_SURVIVE(0, 0x8000 | ((p % PG_PARTS) << 13) | (p / PG_PARTS))
        } else {
        const bool real = tid < PG_GB && blk < ntile64 && blk != best;
        const int part = tid - PG_GB;
        if ((real || (tid >= PG_GB && sib_here && part != own)) && (unsigned long long)s_full[tid] <= thr)
            PG_SURVIVE(0, real ? blk : (0x8000 | (part << 13) | best))
        }
    }
    __syncthreads();
    const unsigned int cnt = s_cnt;
    if (cnt == 0u) return;
    if (tid == 0) {
        unsigned int b = 0xFFFFFFFFu;
        if (cnt <= light_max) {
            b = atomicAdd(item_count, cnt);
            if (b > item_cap || cnt > item_cap - b) {          // buffer full: blank what fits, redo the read
                for (unsigned int i = b; i < item_cap && i < b + cnt; i++) items[i] = (unsigned long long)PG_ITEM_NULL << 16;
                b = 0xFFFFFFFFu;
            }
        }
        if (b == 0xFFFFFFFFu) heavy[rc] = 1;
        s_base = b;
    }
    __syncthreads();
    const unsigned int gb = s_base;
    if (gb == 0xFFFFFFFFu) return;
    for (unsigned int i = tid; i < cnt; i += BLOCK)
        items[gb + i] = ((unsigned long long)rc << 32) | s_list[i];      // rc << 32 | task << 16 | block
#undef PG_BROW
#undef PG_BADD
#undef PG_BSPILL
#undef PG_SURVIVE
}


// ------------------------------------------------------------------ large models: coarse first-level bounds
//
// With more than one group of blocks (G > ~1 800) k_bound would stage the read's rows and walk the 100 sample lists
// once per group of 28 blocks (8 times for 10 000 genera, 121 ns per read).  Nearly all of those blocks are FAR from
// the read, and for them a much cheaper table proves the same thing: one byte per (word, block), c = min(bm >> 6, 15)
// (k_bm8).  c << 6 <= bm, so LB8(task, block) = (sum of c over the task's draws) << 6 is still a lower bound of every
// genus of the block.  A row of ALL blocks is ntile64 bytes (184 for 10 000 genera), so one CTA stages the read once
// and a lane adds a 16-byte segment (16 blocks) per draw: sixteen draws accumulate in packed 8-bit fields without a
// carry (16 x 15 < 256), then spill into packed 16-bit sums.  Early exits would not pay (a (task, block) pair closes
// after ~20 of 60 draws, the 16 pairs of a lane never together), so the loop is straight-line.
// Level 2: the few pairs the coarse bound leaves open (~2 % on the rdp_scale workload) are bounded again with the exact
// 16-bit minima, gathered from the L2-resident bm / hm tables by 8 lanes per pair; what survives are the items.
#define PG_B8_LIST 1536             // (task, segment, mask) entries per CTA after level 1
#define PG_B8_KEEP 4096             // pairs per CTA after level 2 (kept in the dead row area)
#define PG_B8_MAXSEG 16             // 16-byte segments (of 16 blocks) per CTA
#define PG_B8_MAXBLOCK 448

// grid (reads, column chunks); block = any multiple of 32 up to PG_B8_MAXBLOCK (the host picks the size that wastes
// the fewest lanes in the last round of units).  Several CTAs per SM, so that the staging, the two barriers and the
// latency-bound second level of one CTA hide behind the shared-memory loop of another.
__global__ void __launch_bounds__(PG_B8_MAXBLOCK, 2)
k_bound8(const uint8_t *__restrict__ bm8, int pitch_g, int nsegc, int pitch_s, const uint16_t *__restrict__ bm,
         const uint16_t *__restrict__ hm, const uint16_t *__restrict__ words, const int64_t *__restrict__ off,
         const int32_t *__restrict__ nwords, const uint8_t *__restrict__ flags, const int32_t *__restrict__ order,
         int64_t slot0, const uint32_t *__restrict__ boot_pool, const int32_t *__restrict__ boot_off, int min_boot,
         int ntile64, int sib0, double vmax, const unsigned long long *__restrict__ champ,
         const int32_t *__restrict__ guess, unsigned long long *__restrict__ items, unsigned int *__restrict__ item_count,
         unsigned int item_cap, uint8_t *__restrict__ heavy, unsigned int light_max)
{
    extern __shared__ uint4 sR8[];                  // (n+1) rows x pitch_s bytes; row n is zero
    __shared__ uint32_t s_list[PG_B8_LIST];         // level 1: task << 21 | segment << 16 | mask of open cells
    __shared__ uint32_t s_thr[PG_NUM_BOOT + 1];     // champion + margin per task (0xFFFFFFFF = no champion yet)
    __shared__ uint32_t s_full[PG_B8_MAXSEG * 16];  // task 0: coarse sums per column
    __shared__ unsigned int s_cnt1, s_cnt2, s_base;

    const int tid = threadIdx.x, BLOCK = blockDim.x;
    const int64_t read = order[blockIdx.x];
    if (flags[2 * read + 1]) return;
    const int n = nwords[read];
    if (n == 0) return;
    const size_t rc = (size_t)slot0 + blockIdx.x;
    const uint16_t *w = words + off[read];
    const int segbase = (int)blockIdx.y * nsegc;
    const int nsc = min(nsegc, pitch_g / 16 - segbase);
    char *rows = reinterpret_cast<char *>(sR8);

    // ---- stage this chunk of the read's coarse rows: the word ids first (they are needed again by level 2), then one
    // 16-byte copy per (row, segment)
    uint16_t *sw = reinterpret_cast<uint16_t *>(rows + (size_t)(n + 1) * pitch_s);
    for (int j = tid; j < n; j += BLOCK) sw[j] = w[j];
    __syncthreads();
    {
        const unsigned long long magic = (0x100000000ULL + (unsigned long long)nsc - 1ULL) / (unsigned long long)nsc;   // c / nsc, exact for c < 2^28
        const uint8_t *src = bm8 + (size_t)segbase * 16;
        for (int c = tid; c < n * nsc; c += BLOCK) {
            const int r = (int)(((unsigned long long)c * magic) >> 32), l = c - r * nsc;
            pg_cp_async16(rows + (size_t)r * pitch_s + l * 16, src + (size_t)sw[r] * pitch_g + l * 16);
        }
    }
    const int gs = guess[rc];
    const int best = gs / PG_PARTS, own = gs % PG_PARTS;
    const int col0 = segbase * 16, col1 = col0 + nsc * 16;
    const bool sib_here = sib0 >= col0 && sib0 < col1;
    const bool best_here = best >= col0 && best < col1;
    int k = n >> 3;
    if (k < min_boot) k = min_boot;
    const int nb = (k + 3) >> 2;
    for (int t = tid; t <= PG_NUM_BOOT; t += BLOCK) {
        const unsigned long long cv = __ldg(champ + rc * (PG_NUM_BOOT + 1) + t);
        s_thr[t] = (cv == PG_CHAMP_INIT) ? 0xFFFFFFFFu : (uint32_t)(cv >> 32) + pg_margin(t == 0 ? n : k, vmax);
    }
    for (int c = tid; c < PG_B8_MAXSEG * 16; c += BLOCK) s_full[c] = 0u;
    if (tid < nsc * 4) reinterpret_cast<uint32_t *>(rows + (size_t)n * pitch_s)[tid] = 0u;
    if (tid == 0) { s_cnt1 = 0u; s_cnt2 = 0u; }
    // the best block was evaluated on one part only: its other parts compete like blocks of their own, in the four
    // sibling columns (coarse part minima from hm, gathered while the copies are in flight); the best block's own
    // column and the own part never compete: PG_C8_MAX
    uint32_t sibv[2] = {0u, 0u};
    if (sib_here) {
        const uint16_t *hrow = hm + (size_t)((best * PG_PARTS) >> 5) * PG_NWORDS * 32 + ((best * PG_PARTS) & 31);
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int j = tid + u * BLOCK;
            if (j < n) {
                const uint2 hv = __ldg(reinterpret_cast<const uint2 *>(hrow + (size_t)sw[j] * 32));
                uint32_t p0 = min((hv.x & 0xFFFFu) >> PG_C8_SHIFT, PG_C8_MAX), p1 = min((hv.x >> 16) >> PG_C8_SHIFT, PG_C8_MAX);
                uint32_t p2 = min((hv.y & 0xFFFFu) >> PG_C8_SHIFT, PG_C8_MAX), p3 = min((hv.y >> 16) >> PG_C8_SHIFT, PG_C8_MAX);
                if (own == 0) p0 = PG_C8_MAX;
                if (own == 1) p1 = PG_C8_MAX;
                if (own == 2) p2 = PG_C8_MAX;
                if (own == 3) p3 = PG_C8_MAX;
                sibv[u] = p0 | (p1 << 8) | (p2 << 16) | (p3 << 24);
            }
        }
    }
    pg_cp_async_wait_all();
    __syncthreads();
    if (sib_here || best_here) {
        const uint16_t *hrow = hm + (size_t)((best * PG_PARTS) >> 5) * PG_NWORDS * 32 + ((best * PG_PARTS) & 31);
        for (int j = tid, u = 0; j < n; j += BLOCK, u++) {
            char *row = rows + (size_t)j * pitch_s;
            if (best_here) row[best - col0] = (char)PG_C8_MAX;
            if (sib_here) {
                uint32_t pv;
                if (u < 2) pv = sibv[u];
                else {
                    const uint2 hv = __ldg(reinterpret_cast<const uint2 *>(hrow + (size_t)sw[j] * 32));
                    uint32_t p0 = min((hv.x & 0xFFFFu) >> PG_C8_SHIFT, PG_C8_MAX), p1 = min((hv.x >> 16) >> PG_C8_SHIFT, PG_C8_MAX);
                    uint32_t p2 = min((hv.y & 0xFFFFu) >> PG_C8_SHIFT, PG_C8_MAX), p3 = min((hv.y >> 16) >> PG_C8_SHIFT, PG_C8_MAX);
                    if (own == 0) p0 = PG_C8_MAX;
                    if (own == 1) p1 = PG_C8_MAX;
                    if (own == 2) p2 = PG_C8_MAX;
                    if (own == 3) p3 = PG_C8_MAX;
                    pv = p0 | (p1 << 8) | (p2 << 16) | (p3 << 24);
                }
                *reinterpret_cast<uint32_t *>(row + (sib0 - col0)) = pv;
            }
        }
        __syncthreads();
    }

#define PG8_ADD(v_) { const uint4 t_ = (v_); a0 += t_.x; a1 += t_.y; a2 += t_.z; a3 += t_.w; }
#define PG8_SPILL()                                                                                   \
    { s[0] += a0 & 0x00FF00FFu; s[1] += (a0 >> 8) & 0x00FF00FFu; s[2] += a1 & 0x00FF00FFu; s[3] += (a1 >> 8) & 0x00FF00FFu;   \
      s[4] += a2 & 0x00FF00FFu; s[5] += (a2 >> 8) & 0x00FF00FFu; s[6] += a3 & 0x00FF00FFu; s[7] += (a3 >> 8) & 0x00FF00FFu;   \
      a0 = a1 = a2 = a3 = 0u; }
    // sum of cell c (column 16*segment + c) in the packed 16-bit registers
#define PG8_CELL(c) ((s[2 * ((c) >> 2) + ((c) & 1)] >> (16 * (((c) >> 1) & 1))) & 0xFFFFu)
    // is column `col` a competitor, and as which block field?
    auto field_of = [&](int col) -> int {
        if (col < ntile64) return col != best ? col : -1;
        const int part = col - sib0;
        if (part >= 0 && part < PG_PARTS && part != own) return 0x8000 | (part << 13) | best;
        return -1;
    };

    // ---- level 1.  Units: task 0 in chunks of 64 rows (combined through s_full), then one unit per (replicate, segment)
    const uint4 *lists = reinterpret_cast<const uint4 *>(boot_pool + boot_off[n]);
    const int n0u = ((n + 63) >> 6) * nsc;
    const int nunit = n0u + (nb > 0 ? PG_NUM_BOOT * nsc : 0);
    for (int u = tid; u < nunit; u += BLOCK) {
        uint32_t s[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}, a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;
        if (u < n0u) {
            const int ch = u / nsc, seg = u - ch * nsc;
            const char *base = rows + seg * 16;
            const int j1 = min(n, ch * 64 + 64);
            for (int j = ch * 64; j < j1; j += 16) {
#pragma unroll
                for (int uu = 0; uu < 16; uu++)
                    if (j + uu < j1) PG8_ADD(*reinterpret_cast<const uint4 *>(base + (size_t)(j + uu) * pitch_s))
                PG8_SPILL()
            }
#pragma unroll
            for (int c = 0; c < 16; c++) atomicAdd(&s_full[seg * 16 + c], PG8_CELL(c));
            continue;
        }
        const int v = u - n0u;
        const int task = v / nsc, seg = v - task * nsc;
        const char *base = rows + seg * 16;
        const uint4 *lp = lists + (size_t)(task >> 2) * nb * 4 + (task & 3);
#define PG8_ROW(o) (*reinterpret_cast<const uint4 *>(base + ((o) >> 7) * pitch_s))
        // list entries run one batch (four draws) ahead of the rows; a spill every four batches (16 draws x 15 < 256)
        uint4 q = __ldg(lp);
        for (int b = 0; b < nb; b++) {
            const uint4 qc = q;
            if (b + 1 < nb) q = __ldg(lp + (size_t)(b + 1) * 4);
            const uint4 x0 = PG8_ROW(qc.x), x1 = PG8_ROW(qc.y), x2 = PG8_ROW(qc.z), x3 = PG8_ROW(qc.w);
            PG8_ADD(x0) PG8_ADD(x1) PG8_ADD(x2) PG8_ADD(x3)
            if ((b & 3) == 3) PG8_SPILL()
        }
        PG8_SPILL()
#undef PG8_ROW
        const uint32_t thr = s_thr[1 + task];
        const uint32_t thr8 = min(thr >> PG_C8_SHIFT, 0xFFFFu);        // (sum << 6) <= thr  <=>  sum <= thr >> 6
        const uint32_t m2 = __vminu2(__vminu2(__vminu2(s[0], s[1]), __vminu2(s[2], s[3])),
                                     __vminu2(__vminu2(s[4], s[5]), __vminu2(s[6], s[7])));
        if (min(m2 & 0xFFFFu, m2 >> 16) <= thr8) {
            uint32_t mask = 0u;
#pragma unroll
            for (int c = 0; c < 16; c++)
                if (PG8_CELL(c) <= thr8 && field_of(col0 + seg * 16 + c) >= 0) mask |= 1u << c;
            if (mask) {
                const unsigned int pos = atomicAdd(&s_cnt1, 1u);
                if (pos < PG_B8_LIST) s_list[pos] = ((uint32_t)(1 + task) << 21) | ((uint32_t)seg << 16) | mask;
            }
        }
    }
    __syncthreads();
    if (tid < nsc * 16) {                               // task 0
        const uint32_t thr = s_thr[0];
        if (s_full[tid] <= (thr >> PG_C8_SHIFT) && field_of(col0 + tid) >= 0) {
            const unsigned int pos = atomicAdd(&s_cnt1, 1u);
            if (pos < PG_B8_LIST) s_list[pos] = ((uint32_t)(tid >> 4) << 16) | (1u << (tid & 15));
        }
    }
    __syncthreads();

    // ---- level 2: exact 16-bit bound of every open pair, 8 lanes per list entry.  The staged rows are dead now:
    // their memory takes the read's word ids and the list of surviving pairs.
    const unsigned int cnt1 = s_cnt1;
    const bool overflow = cnt1 > PG_B8_LIST;
    if (!overflow && cnt1 == 0u) return;
    uint32_t *s_keep = reinterpret_cast<uint32_t *>(rows);
    const unsigned int keep_cap = min((unsigned int)PG_B8_KEEP, (unsigned int)(((size_t)(n + 1) * pitch_s) / 4));
    if (!overflow) {
        const int l = tid & 7;
        const unsigned gmask = 0xFFu << (tid & 24);
        const uint32_t *lists32 = reinterpret_cast<const uint32_t *>(lists);
        for (unsigned int e = tid >> 3; e < cnt1; e += BLOCK >> 3) {
            const uint32_t en = s_list[e];
            const int task = (int)(en >> 21), seg = (int)((en >> 16) & 31u);
            const uint32_t thr = s_thr[task];
            const int t1 = task > 0 ? task - 1 : 0;
            const uint32_t *lp32 = lists32 + ((size_t)(t1 >> 2) * nb * 4 + (t1 & 3)) * 4;
            const int terms = task == 0 ? n : k;
            // 64 terms per pass: lane l takes terms l, l + 8, ...; their word ids are looked up once for all the open
            // cells of the entry, and a lane's gathers are in flight together
            uint32_t sum[16];
#pragma unroll
            for (int c = 0; c < 16; c++) sum[c] = 0u;
            for (int j0 = 0; j0 < terms; j0 += 64) {
                uint32_t wj[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int j = j0 + l + 8 * i;
                    wj[i] = 0xFFFFFFFFu;
                    if (j < terms) {
                        const uint32_t r = task == 0 ? (uint32_t)j : __ldg(lp32 + (size_t)(j >> 2) * 16 + (j & 3)) / PG_ROW_PITCH;
                        wj[i] = sw[r];
                    }
                }
                uint32_t mask = en & 0xFFFFu;
                while (mask) {
                    const int c = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int fld = field_of(col0 + seg * 16 + c);
                    const uint16_t *tb;
                    if (fld & 0x8000) {
                        const int pb = best * PG_PARTS + ((fld >> 13) & 3);
                        tb = hm + (size_t)(pb >> 5) * PG_NWORDS * 32 + (pb & 31);
                    } else {
                        tb = bm + (size_t)(fld / PG_GB) * PG_NWORDS * 32 + (fld % PG_GB);
                    }
                    uint32_t v[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) v[i] = wj[i] != 0xFFFFFFFFu ? (uint32_t)__ldg(tb + (size_t)wj[i] * 32) : 0u;
                    uint32_t acc = 0u;
#pragma unroll
                    for (int i = 0; i < 8; i++) acc += v[i];
#pragma unroll
                    for (int cc = 0; cc < 16; cc++)
                        if (cc == c) sum[cc] += acc;
                }
            }
            uint32_t mask = en & 0xFFFFu;
            while (mask) {
                const int c = __ffs(mask) - 1;
                mask &= mask - 1;
                uint32_t tot = 0u;
#pragma unroll
                for (int cc = 0; cc < 16; cc++)
                    if (cc == c) tot = sum[cc];
                tot += __shfl_xor_sync(gmask, tot, 1);
                tot += __shfl_xor_sync(gmask, tot, 2);
                tot += __shfl_xor_sync(gmask, tot, 4);
                if (l == 0 && tot <= thr) {
                    const unsigned int pos = atomicAdd(&s_cnt2, 1u);
                    if (pos < keep_cap) s_keep[pos] = ((uint32_t)task << 16) | (uint32_t)field_of(col0 + seg * 16 + c);
                }
            }
        }
    }
    __syncthreads();
    const unsigned int cnt = s_cnt2;
    if (!overflow && cnt == 0u) return;
    if (tid == 0) {
        unsigned int b = 0xFFFFFFFFu;
        if (!overflow && cnt <= light_max && cnt <= keep_cap) {
            b = atomicAdd(item_count, cnt);
            if (b > item_cap || cnt > item_cap - b) {          // buffer full: blank what fits, redo the read
                for (unsigned int i = b; i < item_cap && i < b + cnt; i++) items[i] = (unsigned long long)PG_ITEM_NULL << 16;
                b = 0xFFFFFFFFu;
            }
        }
        if (b == 0xFFFFFFFFu) heavy[rc] = 1;
        s_base = b;
    }
    __syncthreads();
    const unsigned int gb = s_base;
    if (gb == 0xFFFFFFFFu) return;
    for (unsigned int i = tid; i < cnt; i += BLOCK)
        items[gb + i] = ((unsigned long long)rc << 32) | s_keep[i];      // rc << 32 | task << 16 | block
#undef PG8_ADD
#undef PG8_SPILL
#undef PG8_CELL
}

// The block to evaluate first, from the coarse table: 32 sampled words x every block column -> the PG_G8_CAND blocks
// with the smallest coarse sums -> exact part minima (hm) of those blocks' parts -> the best part.  (k_guess_bm reads
// the part minima of EVERY part: 23 groups x 32 words for 10 000 genera.)
#define PG_G8_CAND 8
__global__ void __launch_bounds__(256)
k_guess8(const uint8_t *__restrict__ bm8, int pitch_g, const uint16_t *__restrict__ hm, const uint16_t *__restrict__ words,
         const int64_t *__restrict__ off, const int32_t *__restrict__ nwords, const int32_t *__restrict__ order,
         int nreads_b, int64_t slot0, int ntile64, int32_t *__restrict__ guess)
{
    extern __shared__ uint16_t s_sum[];             // [8 warps][pitch_g] coarse sums per column
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * (blockDim.x >> 5) + warp;
    if (slot >= nreads_b) return;
    const int64_t read = order[slot];
    const int n = nwords[read];
    uint32_t best = 0u;
    if (n > 0) {
        uint16_t *mysum = s_sum + (size_t)warp * pitch_g;
        const uint16_t *w = words + off[read];
        const int ns = n < 32 ? n : 32, stride = n / ns;
        const uint32_t wv = lane < ns ? (uint32_t)__ldg(w + lane * stride) : 0u;
        // lane owns the 4-byte column groups lane, lane + 32, ... (every lane takes part in the shuffles)
        const int ncg = pitch_g / 4;
        for (int cg0 = 0; cg0 < ncg; cg0 += 32) {
            const int cg = cg0 + lane;
            uint32_t lo = 0u, hi = 0u;              // packed 16-bit sums of bytes {0,2} and {1,3}
            for (int j = 0; j < ns; j++) {
                const uint32_t wj = __shfl_sync(0xffffffffu, wv, j);
                if (cg < ncg) {
                    const uint32_t x = __ldg(reinterpret_cast<const uint32_t *>(bm8 + (size_t)wj * pitch_g) + cg);
                    lo += x & 0x00FF00FFu;
                    hi += (x >> 8) & 0x00FF00FFu;
                }
            }
            if (cg < ncg) {
                mysum[4 * cg + 0] = (uint16_t)(lo & 0xFFFFu); mysum[4 * cg + 1] = (uint16_t)(hi & 0xFFFFu);
                mysum[4 * cg + 2] = (uint16_t)(lo >> 16);     mysum[4 * cg + 3] = (uint16_t)(hi >> 16);
            }
        }
        __syncwarp();
        // the PG_G8_CAND smallest (sum, column) keys, one per round; keys are distinct
        uint32_t last = 0u, mycand = 0u;
        bool first = true;
        for (int r = 0; r < PG_G8_CAND; r++) {
            uint32_t kmin = 0xFFFFFFFFu;
            for (int c = lane; c < ntile64; c += 32) {
                const uint32_t key = ((uint32_t)mysum[c] << 12) | (uint32_t)c;
                if ((first || key > last) && key < kmin) kmin = key;
            }
            kmin = __reduce_min_sync(0xffffffffu, kmin);
            if (kmin == 0xFFFFFFFFu) kmin = last;   // fewer than PG_G8_CAND blocks: repeat the last one
            last = kmin;
            first = false;
            if ((lane >> 2) == r) mycand = kmin & 0xFFFu;
        }
        // lane = (candidate, part): exact part minima over the sampled words
        const uint32_t pb = mycand * PG_PARTS + (lane & 3);
        const uint16_t *tb = hm + (size_t)(pb >> 5) * PG_NWORDS * 32 + (pb & 31);
        uint32_t acc = 0u;
        for (int j = 0; j < ns; j++) {
            const uint32_t wj = __shfl_sync(0xffffffffu, wv, j);
            acc += __ldg(tb + (size_t)wj * 32);
        }
        best = __reduce_min_sync(0xffffffffu, (acc << 12) | pb) & 0xFFFu;      // acc < 2^17, pb < 2^12
    }
    if (lane == 0) guess[slot0 + slot] = (int32_t)best;
}

// One group of 8 lanes per item (read, block, task): the block's 64 exact sums straight from the
// L2-resident table (an item is 60 rows of 128 bytes; staging the read's 486 rows would cost more),
// stopping as soon as every partial sum is above champion + margin.
__global__ void __launch_bounds__(256)
k_light(const uint16_t *__restrict__ qtable, const uint16_t *__restrict__ words, const int64_t *__restrict__ off,
        const int32_t *__restrict__ nwords, const int32_t *__restrict__ order_base,
        const uint32_t *__restrict__ boot_pool,
        const int32_t *__restrict__ boot_off, int min_boot, const unsigned long long *__restrict__ blockmask,
        double vmax, const unsigned long long *__restrict__ items, const unsigned int *__restrict__ item_count,
        unsigned int item_cap, unsigned long long *__restrict__ champ, unsigned int *__restrict__ ncand,
        unsigned long long *__restrict__ cand, unsigned int *__restrict__ items_total,
        const uint16_t *__restrict__ hm /* plan 3: part minima, or NULL */)
{
    const int lane = threadIdx.x & 31, l = lane & 7;
    const int gshift = lane & ~7;
    const unsigned gmask = 0xFFu << gshift;
    const unsigned int ngroups = gridDim.x * (blockDim.x >> 3);
    unsigned int cnt = *item_count;
    if (cnt > item_cap) cnt = item_cap;
    if (items_total && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(items_total, cnt);   // NULL: the list holds blanks, its writer counted
    for (unsigned int it = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3); it < cnt; it += ngroups) {
        const unsigned long long e = items[it];
        const int task = (int)((e >> 16) & 0xFFFFu);
        if (task == (int)PG_ITEM_NULL) continue;
        // block field: bit 15 = part item (plan 3: only one part of the block is open), bits 13-14 = which part
        const bool part_item = ((e >> 15) & 1ULL) != 0ULL;
        const int which = (int)((e >> 13) & 3ULL);
        const int blk = (int)(e & 0x1FFFu);
        bool lane_on = !part_item || l / (8 / PG_PARTS) == which;
        const size_t rc = (size_t)(e >> 32);
        const int64_t read = order_base[rc];
        const int n = nwords[read];
        const uint16_t *w = words + off[read];
        int k = n >> 3;
        if (k < min_boot) k = min_boot;
        const int nb = (k + 3) >> 2;
        const int terms = task == 0 ? n : k;
        const uint32_t margin = pg_margin(terms, vmax);
        unsigned long long *slot = champ + rc * (PG_NUM_BOOT + 1) + task;
        const unsigned long long cv = *reinterpret_cast<volatile unsigned long long *>(slot);
        const unsigned long long thr = (cv == PG_CHAMP_INIT) ? ~0ULL : (cv >> 32) + margin;
        const uint4 *tb = reinterpret_cast<const uint4 *>(qtable + (size_t)blk * PG_NWORDS * 64) + l;
        uint32_t s0 = 0u, s1 = 0u, s2 = 0u, s3 = 0u, s4 = 0u, s5 = 0u, s6 = 0u, s7 = 0u;
        bool pruned = false;
        const uint4 *lp = reinterpret_cast<const uint4 *>(boot_pool + boot_off[n]) +
                          (task > 0 ? (size_t)((task - 1) >> 2) * nb * 4 + ((task - 1) & 3) : 0);
        if (hm && !part_item) {
            // A whole block is open because its 64-position bound was not enough.  Bound its four parts first
            // (8 bytes per draw instead of a 128-byte row; the 8 lanes split the draws): a part whose bound is
            // above the threshold switches its two lanes off, a block with no part left is dropped here.
            const uint16_t *hrow = hm + (size_t)((blk * PG_PARTS) >> 5) * PG_NWORDS * 32 + ((blk * PG_PARTS) & 31);
            uint32_t lb0 = 0u, lb1 = 0u, lb2 = 0u, lb3 = 0u;
#define PG_PART_ADD(r_)                                                                                     \
    if ((int)(r_) < n) {                                                                                    \
        const uint2 hv_ = __ldg(reinterpret_cast<const uint2 *>(hrow + (size_t)__ldg(w + (r_)) * 32));      \
        lb0 += hv_.x & 0xFFFFu; lb1 += hv_.x >> 16; lb2 += hv_.y & 0xFFFFu; lb3 += hv_.y >> 16;             \
    }
            if (task == 0) {
                for (int j = l; j < n; j += 8) PG_PART_ADD((uint32_t)j)
            } else {
                for (int b = l; b < nb; b += 8) {
                    const uint4 q = __ldg(lp + (size_t)b * 4);
                    PG_PART_ADD(q.x / PG_ROW_PITCH) PG_PART_ADD(q.y / PG_ROW_PITCH)
                    PG_PART_ADD(q.z / PG_ROW_PITCH) PG_PART_ADD(q.w / PG_ROW_PITCH)
                }
            }
#undef PG_PART_ADD
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                lb0 += __shfl_xor_sync(gmask, lb0, o); lb1 += __shfl_xor_sync(gmask, lb1, o);
                lb2 += __shfl_xor_sync(gmask, lb2, o); lb3 += __shfl_xor_sync(gmask, lb3, o);
            }
            const int part = l / (8 / PG_PARTS);
            const uint32_t mine = part == 0 ? lb0 : (part == 1 ? lb1 : (part == 2 ? lb2 : lb3));
            lane_on = (unsigned long long)mine <= thr;
            if ((__ballot_sync(gmask, lane_on) & gmask) == 0u) continue;
        }
        const int nstep = task == 0 ? (n + 7) >> 3 : (nb + 1) >> 1;
        // (prefetching the next step's list entries and word ids while this step's rows are in flight was
        // measured: 80 registers, a quarter fewer resident warps, 12 % slower)
        for (int st = 0; st < nstep; st++) {
            // row indices of this step's 8 draws (index n = padding = no row)
            uint32_t r[8];
            if (task == 0) {
#pragma unroll
                for (int u = 0; u < 8; u++) r[u] = (uint32_t)min(st * 8 + u, n);
            } else {
                const uint4 qa = __ldg(lp + (size_t)(2 * st) * 4);
                uint4 qb = make_uint4((uint32_t)n * PG_ROW_PITCH, (uint32_t)n * PG_ROW_PITCH, (uint32_t)n * PG_ROW_PITCH,
                                      (uint32_t)n * PG_ROW_PITCH);
                if (2 * st + 1 < nb) qb = __ldg(lp + (size_t)(2 * st + 1) * 4);
                r[0] = qa.x / PG_ROW_PITCH; r[1] = qa.y / PG_ROW_PITCH; r[2] = qa.z / PG_ROW_PITCH; r[3] = qa.w / PG_ROW_PITCH;
                r[4] = qb.x / PG_ROW_PITCH; r[5] = qb.y / PG_ROW_PITCH; r[6] = qb.z / PG_ROW_PITCH; r[7] = qb.w / PG_ROW_PITCH;
            }
            uint32_t wi[8];
#pragma unroll
            for (int u = 0; u < 8; u++) wi[u] = (int)r[u] < n ? (uint32_t)__ldg(w + r[u]) : 0xFFFFFFFFu;
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++)
                v[u] = (wi[u] != 0xFFFFFFFFu && lane_on) ? __ldg(tb + (size_t)wi[u] * 8) : make_uint4(0u, 0u, 0u, 0u);
            uint32_t c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;
#pragma unroll
            for (int u = 0; u < 8; u++) { PG_QADD4(v[u]) }
            PG_QSPILL()
            const uint32_t smin = lane_on ? min(min(min(s0, s1), min(s2, s3)), min(min(s4, s5), min(s6, s7))) : 0xFFFFFFFFu;
            const unsigned over = __ballot_sync(gmask, (unsigned long long)smin > thr);
            if (((over >> gshift) & 0xFFu) == 0xFFu) { pruned = true; break; }
        }
        if (pruned) continue;
        const uint32_t sums[8] = {s0, s1, s2, s3, s4, s5, s6, s7};
        const uint32_t gbase = (uint32_t)blk * 64u;
        PgPending pend;
        const uint32_t vb8 = lane_on ? (uint32_t)(__ldg(blockmask + blk) >> (8 * l)) & 0xFFu : 0u;
        pg_epilogue_begin<8>(gmask, l == 0, sums, gbase + (uint32_t)l * 8u, gbase, vb8, slot, pend);
        pg_epilogue_finish<8>(gmask, l == 0, gshift, sums, gbase + (uint32_t)l * 8u, gbase, vb8, task, margin, pend,
                              ncand + rc, cand + rc * PG_CANDCAP);
    }
}

// ------------------------------------------------------------------ phase 2

__device__ __forceinline__ int pg_pos2genus(const int32_t *__restrict__ perm, uint32_t pos, int npos, int G)
{
    const int g = pos < (uint32_t)npos ? perm[pos] : G;     // padding positions hold a value >= G
    return g < G ? g : 0;
}

// strict fp32 sum of genus g over the task's rows, in the reference's order
__device__ float pg_strict_sum(const float *__restrict__ table, const uint16_t *sw, int n, int k, int nb,
                               const uint32_t *__restrict__ list, int task, uint32_t g)
{
    const float *tb = table + ((size_t)(g >> 5) * PG_NWORDS) * PG_GENUS_TILE + (g & 31);
    float a = 0.f;
    if (task == 0) {
#pragma unroll 4
        for (int j = 0; j < n; j++) a = __fadd_rn(a, __ldg(tb + (size_t)sw[j] * PG_GENUS_TILE));
    } else {
        const int t = task - 1;
        const uint32_t *lp = list + ((size_t)(t / 4) * nb * 4 + (t % 4)) * 4;
#pragma unroll 4
        for (int j = 0; j < k; j++) {
            const uint32_t r = __ldg(lp + (size_t)(j >> 2) * 16 + (j & 3)) / PG_ROW_PITCH;
            a = __fadd_rn(a, __ldg(tb + (size_t)sw[r] * PG_GENUS_TILE));
        }
    }
    return a;
}

// the same sum by the whole warp.  Lane l gathers the 16 consecutive terms 16l .. 16l+15 of a 512-term pass (all
// loads in flight at once); the running sum then hops from lane to lane: every lane adds its own 16 terms in
// order to the value it receives from its predecessor, so the adds stay one chain in the reference's order at
// the price of ONE shuffle per 16 terms (a shuffle per term kept the LSU pipe 80 % busy).  Every lane returns
// the result.
__device__ float pg_strict_sum_warp(const float *__restrict__ table, const uint16_t *sw, int n, int k, int nb,
                                    const uint32_t *__restrict__ list, int task, uint32_t g, int lane)
{
    const float *tb = table + ((size_t)(g >> 5) * PG_NWORDS) * PG_GENUS_TILE + (g & 31);
    const int terms = task == 0 ? n : k;
    const int t1 = task > 0 ? task - 1 : 0;
    const uint32_t *lp = list + ((size_t)(t1 / 4) * nb * 4 + (t1 % 4)) * 4;
    float a = 0.f;
    for (int base = 0; base < terms; base += 512) {
        float v[16];
        const int j0 = base + lane * 16;
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const int j = j0 + u;
            v[u] = 0.f;
            if (j < terms) {
                const uint32_t r = task == 0 ? (uint32_t)j : __ldg(lp + (size_t)(j >> 2) * 16 + (j & 3)) / PG_ROW_PITCH;
                v[u] = __ldg(tb + (size_t)sw[r] * PG_GENUS_TILE);
            }
        }
        const int cnt = terms - j0;                 // terms this lane owns in this pass (may be <= 0 or > 16)
        const int hops = (terms - base + 15) / 16 < 32 ? (terms - base + 15) / 16 : 32;
        for (int h = 0; h < hops; h++) {
            // lane h continues the chain from lane h-1 (lane 0 from the previous pass); the other lanes idle along
            const float in = __shfl_sync(0xffffffffu, a, h == 0 ? (base == 0 ? 0 : 31) : h - 1);
            if (lane == h) {
                float x = (h == 0 && base == 0) ? 0.f : in;
#pragma unroll
                for (int u = 0; u < 16; u++)
                    if (u < cnt) x = __fadd_rn(x, v[u]);
                a = x;
            }
        }
        // the pass ends in lane hops-1: park its value in lane 31 for the next pass / the result
        a = __shfl_sync(0xffffffffu, a, hops - 1);
    }
    return a;
}

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS)
k_resolve(const float *__restrict__ table, const uint16_t *__restrict__ words, const int64_t *__restrict__ off,
          const int32_t *__restrict__ nwords, const uint8_t *__restrict__ flags, const int32_t *__restrict__ order,
          int nreads_b, int64_t slot0, int nmax, const uint32_t *__restrict__ boot_pool,
          const int32_t *__restrict__ boot_off, int min_boot, int G, double vmax,
          const unsigned long long *__restrict__ champ, const unsigned int *__restrict__ ncand,
          const unsigned long long *__restrict__ cand, const int32_t *__restrict__ anc, int depth,
          pg_result *__restrict__ results, int32_t *__restrict__ boot_winners, int *__restrict__ fb_count,
          int32_t *__restrict__ fb_list, const int32_t *__restrict__ perm, int npos, const uint8_t *__restrict__ heavy,
          int *__restrict__ hv_count, int32_t *__restrict__ hv_list)
{
    extern __shared__ unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * WARPS + warp;
    if (slot >= nreads_b) return;
    const size_t per_warp = (((size_t)nmax * 2 + 15) & ~(size_t)15) + (PG_NUM_BOOT + 1) * 4 + 16;
    unsigned char *mine = smem_raw + warp * per_warp;
    uint16_t *sw = reinterpret_cast<uint16_t *>(mine);
    int *winners = reinterpret_cast<int *>(mine + (((size_t)nmax * 2 + 15) & ~(size_t)15));
    unsigned int *need = reinterpret_cast<unsigned int *>(winners + PG_NUM_BOOT + 1);

    const int64_t read = order[slot];
    pg_result *res = results + read;
    uint32_t *raw = reinterpret_cast<uint32_t *>(res);
    if (flags[2 * read + 1]) {                                  // A2: short read
        if (lane < 16) raw[lane] = (lane == 0) ? 0xFFFFFFFFu : (lane == 3 ? (1u << 8) : 0u);
        if (boot_winners)
            for (int r = lane; r < PG_NUM_BOOT; r += 32) boot_winners[read * PG_NUM_BOOT + r] = -1;
        return;
    }
    const size_t rc = (size_t)slot0 + slot;
    if (heavy && heavy[rc]) {                                   // too many items: the all-block kernel redoes it
        if (lane == 0) hv_list[atomicAdd(hv_count, 1)] = (int32_t)read;
        return;
    }
    const unsigned int nc = ncand[rc];
    if (nc > PG_CANDCAP) {                                      // too many near-ties: strict kernels take it
        if (lane == 0) fb_list[atomicAdd(fb_count, 1)] = (int32_t)read;
        return;
    }
    const int n = nwords[read];
    int k = n >> 3;
    if (k < min_boot) k = min_boot;
    const int nb = (k + 3) >> 2;
    const uint16_t *w = words + off[read];
    for (int j = lane; j < n; j += 32) sw[j] = w[j];
    const unsigned long long *mychamp = champ + rc * (PG_NUM_BOOT + 1);
    const unsigned long long *mycand = cand + rc * PG_CANDCAP;
    const uint32_t *list = boot_pool + boot_off[n];
    const uint32_t margin_full = pg_margin(n, vmax), margin_rep = pg_margin(k, vmax);
    const bool trivial_rep = (nb == 0 || n == 0);               // k == 0: all sums 0 -> genus 0

    for (int t = lane; t <= PG_NUM_BOOT; t += 32)
        winners[t] = (n == 0 || (t > 0 && trivial_rep)) ? 0 : pg_pos2genus(perm, (uint32_t)mychamp[t], npos, G);   // table position -> genus
    if (lane < 4) need[lane] = 0u;
    __syncwarp();
    if (n > 0) {
        if (lane == 0) need[0] = 1u;                            // task 0: the score is always strict
        for (unsigned int e = lane; e < nc; e += 32) {
            const unsigned long long en = mycand[e];
            const int t = (int)(en >> 56);
            if (t > 0 && trivial_rep) continue;
            const unsigned long long csum = mychamp[t] >> 32;
            if ((en & 0xFFFFFFFFu) <= csum + (t == 0 ? margin_full : margin_rep)) atomicOr(&need[t >> 5], 1u << (t & 31));
        }
    }
    __syncwarp();

    float score = 0.f;
    for (int wd = 0; wd < 4; wd++) {
        unsigned int bits = need[wd];
        while (bits) {
            const int t = wd * 32 + __ffs(bits) - 1;
            bits &= bits - 1;
            const unsigned long long ch = mychamp[t];
            const unsigned long long thr = (ch >> 32) + (t == 0 ? margin_full : margin_rep);
            uint32_t best_ord = 0u, best_g = 0xFFFFFFFFu;
            // virtual list: entries 0..nc-1, then the champion at index nc
            for (unsigned int e0 = 0; e0 <= nc; e0 += 32) {
                const unsigned int e = e0 + lane;
                bool match = false;
                uint32_t g = 0;
                if (e < nc) {
                    const unsigned long long en = mycand[e];
                    match = ((int)(en >> 56) == t) && ((en & 0xFFFFFFFFu) <= thr);
                    g = (uint32_t)(en >> 32) & 0xFFFFFFu;
                } else if (e == nc) {
                    match = true;
                    g = (uint32_t)ch;
                }
                if (match) g = (uint32_t)pg_pos2genus(perm, g, npos, G);    // ties resolve on the genus index, not the position
                const unsigned int bal = __ballot_sync(0xffffffffu, match);
                float a = 0.f;
                if (__popc(bal) <= 6) {
                    // few survivors (the usual case): the warp sums them one after the other
                    unsigned int todo = bal;
                    while (todo) {
                        const int src = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const float r = pg_strict_sum_warp(table, sw, n, k, nb, list, t, __shfl_sync(0xffffffffu, g, src), lane);
                        if (lane == src) a = r;
                    }
                } else if (match) {
                    a = pg_strict_sum(table, sw, n, k, nb, list, t, g);      // many: one survivor per lane
                }
                if (match) {
                    const uint32_t ob = pg_ord(a);
                    const uint32_t gm = __reduce_max_sync(bal, ob);
                    const uint32_t gi = __reduce_min_sync(bal, ob == gm ? g : 0xFFFFFFFFu);
                    if (gm > best_ord || (gm == best_ord && gi < best_g)) { best_ord = gm; best_g = gi; }
                }
                if (bal) {
                    const int src = __ffs(bal) - 1;
                    best_ord = __shfl_sync(0xffffffffu, best_ord, src);
                    best_g = __shfl_sync(0xffffffffu, best_g, src);
                }
            }
            if (lane == 0) winners[t] = (int)best_g;
            if (t == 0) score = pg_unord(best_ord);
        }
    }
    __syncwarp();

    // ---- A9 vote
    const int genus = winners[0];
    int gb[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int r = q * 32 + lane;
        gb[q] = r < PG_NUM_BOOT ? winners[1 + r] : -1;
        if (boot_winners && r < PG_NUM_BOOT) boot_winners[read * PG_NUM_BOOT + r] = gb[q];
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
    const uint32_t packed = (uint32_t)myvote & 0xFFu;
    const uint32_t w0 = __shfl_sync(0xffffffffu, packed, (lane & 7) * 4 + 0);
    const uint32_t w1 = __shfl_sync(0xffffffffu, packed, (lane & 7) * 4 + 1);
    const uint32_t w2 = __shfl_sync(0xffffffffu, packed, (lane & 7) * 4 + 2);
    const uint32_t w3 = __shfl_sync(0xffffffffu, packed, (lane & 7) * 4 + 3);
    const uint32_t vw = w0 | (w1 << 8) | (w2 << 16) | (w3 << 24);
    if (lane == 0) {
        raw[0] = (uint32_t)genus;
        raw[1] = (uint32_t)n;
        raw[2] = __float_as_uint(score);
        raw[3] = (uint32_t)flags[2 * read] | ((uint32_t)levels << 16);
    }
    if (lane < 8) raw[4 + lane] = vw;
    if (lane >= 8 && lane < 12) raw[4 + lane] = 0u;
}

// ------------------------------------------------------------------ host

template <int BLOCK, int MINB>
static int launch_q(pg_ctx *ctx, const pg_model *md, unsigned nreads_b, unsigned nblk_y, size_t smem, const uint16_t *d_words,
                    const int64_t *d_off, const int32_t *d_nwords, const uint8_t *d_flags, const int32_t *d_order,
                    int64_t slot0, int min_boot, unsigned long long *d_champ, unsigned int *d_ncand,
                    unsigned long long *d_cand, const int32_t *d_guess)
{
    dim3 grid(nreads_b, nblk_y);
    if (nblk_y > 1) {
        PG_CUDA(ctx, pg_smem_unlock(ctx, k_classify_q<BLOCK, MINB, true>));
        k_classify_q<BLOCK, MINB, true><<<grid, BLOCK, smem, ctx->stream>>>(md->d_qtable, d_words, d_off, d_nwords, d_flags, d_order,
                                                              slot0, ctx->d_boot_pool, ctx->d_boot_off, min_boot, md->d_blockmask,
                                                              md->vmax, d_champ, d_ncand, d_cand, d_guess);
    } else {
        PG_CUDA(ctx, pg_smem_unlock(ctx, k_classify_q<BLOCK, MINB, false>));
        k_classify_q<BLOCK, MINB, false><<<grid, BLOCK, smem, ctx->stream>>>(md->d_qtable, d_words, d_off, d_nwords, d_flags, d_order,
                                                              slot0, ctx->d_boot_pool, ctx->d_boot_off, min_boot, md->d_blockmask,
                                                              md->vmax, d_champ, d_ncand, d_cand, d_guess);
    }
    PG_LAUNCHED(ctx);
    return PG_OK;
}

// phase 1 for one bucket (timed by the caller).  version 2 = best block + lower bounds + items
// (reads with too many items are flagged in cb.heavy); version 1 = every block, partial-sum pruning.
bool pg_mma_usable(const pg_model *md, int nmax);                                    // pg_mma.cu
int pg_mma_build_tables(pg_model *md);
int pg_mma_launch(pg_ctx *ctx, const pg_model *md, unsigned nreads_b, int nmax, const uint16_t *d_words, const int64_t *d_off,
                  const int32_t *d_nwords, const uint8_t *d_flags, const int32_t *d_order, int64_t slot0, int min_boot,
                  const PgCertBufs &cb, const int32_t *d_guess, unsigned int light_max);

int pg_certified_phase1(pg_ctx *ctx, const pg_model *md, const Bucket &bk, unsigned nreads_b, int nmax,
                        const uint16_t *d_words, const int64_t *d_off, const int32_t *d_nwords,
                        const uint8_t *d_flags, const int32_t *d_order, int64_t slot0, int min_boot,
                        const PgCertBufs &cb, int version)
{
    const size_t smem = (size_t)(nmax + 1) * 128;
    static int noprune = -1;                            // PG_NO_PRUNE=1: every block in full (A/B switch for profiling)
    if (noprune < 0) { const char *e = getenv("PG_NO_PRUNE"); noprune = (e && atoi(e)) ? 1 : 0; }
    int32_t *d_guess = cb.guess;
    if (noprune || md->ntile64 < 2) { d_guess = NULL; version = 1; }
    // version 4 = plan 4 (pg_mma.cu): plan 3's best part and bounds as one tensor-core product per read; reads it
    // cannot take (more than 640 words, a model without the byte tables, an explicit bound_level) go through plan 3
    const bool use_mma = d_guess && version == 4 && cb.bound_level == 0 && cb.force_part < 0 && pg_mma_usable(md, nmax);
    if (version == 4) version = 3;
    // version 3 = plan 3 (best part of the best block, PG_PARTS reads per CTA, the default); 2 = plan 2 (whole best block);
    // 1 = plan 1 (every block, partial-sum pruning)
    unsigned nblk_y = (unsigned)md->ntile64;
    // large models (more than one group of blocks): coarse first-level bounds over all blocks at once (k_bound8)
    static int env_b8 = -2;                             // PG_BOUND8=0: never, 1: always (A/B switch)
    if (env_b8 == -2) { const char *e = getenv("PG_BOUND8"); env_b8 = e ? atoi(e) : -1; }
    const int blevel = env_b8 >= 0 ? (env_b8 ? 2 : 1) : cb.bound_level;
    int nsegc = 0, pitch_s = 0, nchunk8 = 0, block8 = PG_B8_MAXBLOCK;
    bool use8 = d_guess && version == 3 && md->d_bm8 && blevel == 2;   // measured on 10 000 genera: k_bound 4.66 M reads/s, k_bound8 4.49 M
    if (use8) {
        // column chunks sized for two CTAs of up to 448 threads per SM (a 10 000-genus model with 486-word reads: one
        // chunk of 12 segments, 102 KB); four smaller CTAs per SM were measured: the 96-byte row pitch that fits costs
        // 70 % more shared-memory wavefronts in bank conflicts
        const int nseg_total = md->bm8_pitch / 16;
        const size_t budget = 103 * 1024 - (size_t)nmax * 2;
        int maxseg = (int)(budget / ((size_t)(nmax + 1) * 16));
        if (maxseg > PG_B8_MAXSEG) maxseg = PG_B8_MAXSEG;
        if (maxseg < 1) use8 = false;
        else {
            nchunk8 = (nseg_total + maxseg - 1) / maxseg;
            nsegc = (nseg_total + nchunk8 - 1) / nchunk8;
            pitch_s = nsegc * 16;
            // an odd number of 16-byte units per row spreads the rows of a warp's tasks over all bank offsets
            if (!(nsegc & 1) && (size_t)(nmax + 1) * (pitch_s + 16) <= budget) pitch_s += 16;
            // block size: the one that leaves the fewest idle lanes in the last round of (task, segment) units
            const int units = nsegc * (((nmax + 63) >> 6) + PG_NUM_BOOT);
            double best_waste = 1e30;
            for (int bsz = PG_B8_MAXBLOCK; bsz >= 256; bsz -= 32) {
                const double waste = (double)((units + bsz - 1) / bsz) * bsz / units;
                if (waste < best_waste - 1e-9) { best_waste = waste; block8 = bsz; }
            }
        }
    }
    // the coarse table also picks the block to evaluate first (k_guess8), whichever kernel bounds the rest
    static int env_g8 = -2;                             // PG_GUESS8=0: k_guess_bm even for large models (A/B switch)
    if (env_g8 == -2) { const char *e = getenv("PG_GUESS8"); env_g8 = e ? atoi(e) : -1; }
    // a second try (reads the first guess left "heavy"): the exact part minima of EVERY part over 64 sampled words
    const bool guess8 = d_guess && version == 3 && md->d_bm8 && (use8 || (env_g8 != 0 && md->ngroup > 1)) && !cb.retry;
    // cb.stage (plan 4 only): 1 = guess + tensor-core kernel, 2 = the item kernel of reads that went through stage 1 --
    // the caller runs stage 2 of one slice on a second stream under stage 1 of the next; 0 = everything, in turn
    const bool back_only = cb.stage == 2 && use_mma;
    if (back_only) {
    } else if (guess8) {
        k_guess8<<<(nreads_b + 7) / 8, 256, (size_t)8 * md->bm8_pitch * 2, ctx->stream>>>(
            md->d_bm8, md->bm8_pitch, md->d_hmtable, d_words, d_off, d_nwords, d_order, (int)nreads_b, slot0, md->ntile64, d_guess);
        PG_LAUNCHED(ctx);
    } else if (d_guess && version == 3) {
        k_guess_bm<<<(nreads_b + 7) / 8, 256, 0, ctx->stream>>>(md->d_hmtable, d_words, d_off, d_nwords, d_order, (int)nreads_b,
                                                               slot0, PG_PARTS * md->ntile64, 32, md->ngroup_h, d_guess, cb.retry ? 64 : 32);
        PG_LAUNCHED(ctx);
    } else if (d_guess && version == 2) {
        k_guess_bm<<<(nreads_b + 7) / 8, 256, 0, ctx->stream>>>(md->d_bmtable, d_words, d_off, d_nwords, d_order, (int)nreads_b,
                                                               slot0, md->ntile64, PG_GB, md->ngroup, d_guess, 32);
        PG_LAUNCHED(ctx);
        nblk_y = 1;                                     // grid row 0 = the guessed block, in full
    } else if (d_guess) {
        k_guess_block<<<(nreads_b + 7) / 8, 256, 0, ctx->stream>>>(md->d_qtable, d_words, d_off, d_nwords, d_order,
                                                                  (int)nreads_b, slot0, md->ntile64, d_guess);
        PG_LAUNCHED(ctx);
    }
    static int qblock = -1;                             // PG_Q_BLOCK=448: experiment switch for the first bucket
    if (qblock < 0) { const char *e = getenv("PG_Q_BLOCK"); qblock = e ? atoi(e) : 0; }
    int rc = PG_OK;
    if (back_only) {
    } else if (use_mma) {
        PG_CUDA(ctx, cudaMemsetAsync(cb.item_count, 0, 4, ctx->stream));
        const unsigned int lm = cb.light_max == 0 ? 2048u : (cb.light_max < 0 ? 0u : (unsigned int)cb.light_max);
        PG_TRY(pg_mma_launch(ctx, md, nreads_b, nmax, d_words, d_off, d_nwords, d_flags, d_order, slot0, min_boot, cb, d_guess, lm));
        if (cb.count_mma) ctx->st_mma += nreads_b;
    } else if (d_guess && version == 3) {
        const unsigned npair = (nreads_b + PG_PARTS - 1) / PG_PARTS;
#define PG_LAUNCH_H(B, M)                                                                                               \
    {                                                                                                                   \
        PG_CUDA(ctx, pg_smem_unlock(ctx, k_classify_h<B, M>)); \
        k_classify_h<B, M><<<npair, B, smem, ctx->stream>>>(md->d_qtable, d_words, d_off, d_nwords, d_flags, d_order,   \
                                                            (int)nreads_b, slot0, ctx->d_boot_pool, ctx->d_boot_off,    \
                                                            min_boot, md->d_blockmask, md->vmax, cb.champ, cb.ncand,    \
                                                            cb.cand, d_guess);                                          \
        PG_LAUNCHED(ctx);                                                                                               \
    }
        if (bk.block == 192) PG_LAUNCH_H(256, 3)      /* 192 / 256 / 448 threads measured: 16.43 / 16.54 / 16.34 M reads/s */
        else if (bk.block == 448) PG_LAUNCH_H(352, 2)
        else PG_LAUNCH_H(832, 1)
#undef PG_LAUNCH_H
    } else if (bk.block == 192 && qblock != 448)
        rc = launch_q<192, 3>(ctx, md, nreads_b, nblk_y, smem, d_words, d_off, d_nwords, d_flags, d_order, slot0, min_boot, cb.champ, cb.ncand, cb.cand, d_guess);
    else if (bk.block == 448 || bk.block == 192)
        rc = launch_q<448, 2>(ctx, md, nreads_b, nblk_y, smem, d_words, d_off, d_nwords, d_flags, d_order, slot0, min_boot, cb.champ, cb.ncand, cb.cand, d_guess);
    else
        rc = launch_q<832, 1>(ctx, md, nreads_b, nblk_y, smem, d_words, d_off, d_nwords, d_flags, d_order, slot0, min_boot, cb.champ, cb.ncand, cb.cand, d_guess);
    PG_TRY(rc);
    if (!(d_guess && version >= 2)) return PG_OK;
    if (cb.stage == 1 && use_mma) return PG_OK;

    int light_max = cb.light_max == 0 ? PG_LIGHT_MAX : (cb.light_max < 0 ? 0 : cb.light_max);
    if (light_max > PG_LIGHT_MAX) light_max = PG_LIGHT_MAX;
    if (!use_mma) PG_CUDA(ctx, cudaMemsetAsync(cb.item_count, 0, 4, ctx->stream));
    // (a two-reads-per-CTA k_bound with interleaved rows, like k_classify_h, was measured: same wavefronts,
    // 46 % more instructions, 28 % slower -- LDS.64 rows of 64 bytes gain nothing from the interleave)
    if (use_mma) {
        // best part and bounds are done (pg_mma_launch above)
    } else if (use8) {
        const size_t bsmem = (size_t)(nmax + 1) * pitch_s + (((size_t)nmax * 2 + 15) & ~(size_t)15);
        PG_CUDA(ctx, pg_smem_unlock(ctx, k_bound8));
        k_bound8<<<dim3(nreads_b, (unsigned)nchunk8), block8, bsmem, ctx->stream>>>(
            md->d_bm8, md->bm8_pitch, nsegc, pitch_s, md->d_bmtable, md->d_hmtable, d_words, d_off, d_nwords, d_flags, d_order,
            slot0, ctx->d_boot_pool, ctx->d_boot_off, min_boot, md->ntile64, md->sib0, md->vmax, cb.champ, d_guess, cb.items,
            cb.item_count, cb.item_cap, cb.heavy, (unsigned int)(cb.light_max == 0 ? PG_B8_KEEP : light_max));
        PG_LAUNCHED(ctx);
    } else {
    const size_t bsmem = (size_t)(nmax + 1) * 64;
    static int env_part = -2;                           // PG_BOUND_PART=0/1: block / part columns (A/B switch)
    if (env_part == -2) { const char *e = getenv("PG_BOUND_PART"); env_part = e ? atoi(e) : -1; }
    const bool part_cols = version == 3 && (env_part >= 0 ? env_part != 0 : (blevel == 3 || (blevel == 0 && (cb.force_part >= 0 ? cb.force_part != 0 : md->part_bounds))));
    if (part_cols) {
        PG_CUDA(ctx, pg_smem_unlock(ctx, k_bound<160, true>));
        k_bound<160, true><<<dim3(nreads_b, (unsigned)md->ngroup_h), 160, bsmem, ctx->stream>>>(
            md->d_hmtable, d_words, d_off, d_nwords, d_flags, d_order, slot0, ctx->d_boot_pool, ctx->d_boot_off, min_boot,
            md->ntile64, md->vmax, cb.champ, d_guess, cb.items, cb.item_count, cb.item_cap, cb.heavy,
            (unsigned int)light_max, md->d_hmtable);
    } else {
    PG_CUDA(ctx, pg_smem_unlock(ctx, k_bound<160, false>));
    k_bound<160, false><<<dim3(nreads_b, (unsigned)md->ngroup), 160, bsmem, ctx->stream>>>(
        md->d_bmtable, d_words, d_off, d_nwords, d_flags, d_order, slot0, ctx->d_boot_pool, ctx->d_boot_off, min_boot,
        md->ntile64, md->vmax, cb.champ, d_guess, cb.items, cb.item_count, cb.item_cap, cb.heavy,
        (unsigned int)light_max, version == 3 ? md->d_hmtable : NULL);
    }
    PG_LAUNCHED(ctx);
    }
    static int light_ctas = 0;                          // resident CTAs per SM of the persistent item kernel
    if (!light_ctas) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&light_ctas, k_light, 256, 0) != cudaSuccess || light_ctas < 1) light_ctas = 2;
    }
    k_light<<<ctx->sm_count * light_ctas, 256, 0, ctx->stream>>>(md->d_qtable, d_words, d_off, d_nwords, d_order - slot0, ctx->d_boot_pool,
                                                        ctx->d_boot_off, min_boot, md->d_blockmask, md->vmax, cb.items,
                                                        cb.item_count, cb.item_cap, cb.champ, cb.ncand, cb.cand, use_mma ? NULL : cb.counters + 3,
                                                        version == 3 ? md->d_hmtable : NULL);
    PG_LAUNCHED(ctx);
    return PG_OK;
}

// phase 2 for one bucket: strict re-check of survivors + vote; overflowing reads go to cb.fb_list,
// heavy reads (use_heavy) to cb.hv_list
int pg_certified_phase2(pg_ctx *ctx, const pg_model *md, unsigned nreads_b, int nmax, const uint16_t *d_words,
                        const int64_t *d_off, const int32_t *d_nwords, const uint8_t *d_flags,
                        const int32_t *d_order, int64_t slot0, int min_boot, const PgCertBufs &cb, bool use_heavy,
                        pg_result *d_results, int32_t *d_boot_winners)
{
    constexpr int WARPS = 4;
    const size_t per_warp = (((size_t)nmax * 2 + 15) & ~(size_t)15) + (PG_NUM_BOOT + 1) * 4 + 16;
    const size_t smem = per_warp * WARPS;
    PG_CUDA(ctx, pg_smem_unlock(ctx, k_resolve<WARPS>));
    k_resolve<WARPS><<<(nreads_b + WARPS - 1) / WARPS, 32 * WARPS, smem, ctx->stream>>>(
        md->d_table, d_words, d_off, d_nwords, d_flags, d_order, (int)nreads_b, slot0, nmax, ctx->d_boot_pool,
        ctx->d_boot_off, min_boot, md->G, md->vmax, cb.champ, cb.ncand, cb.cand, md->d_anc, md->depth, d_results,
        d_boot_winners, (int *)cb.counters, cb.fb_list, md->d_perm, md->ntile64 * 64, use_heavy ? cb.heavy : NULL,
        (int *)cb.counters + 1, cb.hv_list);
    PG_LAUNCHED(ctx);
    return PG_OK;
}

