// pg_classify_common.cuh -- device helpers shared by the strict (pg_classify.cu) and
// certified (pg_certified.cu) classification kernels.
#pragma once
#include "pg_internal.cuh"

// ------------------------------------------------------------------ A8 sample lists

#define JR_MULT 0x5DEECE66DULL
#define JR_MASK ((1ULL << 48) - 1)

__device__ __forceinline__ int32_t jr_next(unsigned long long &s, int bits)
{
    s = (s * JR_MULT + 0xBULL) & JR_MASK;
    return (int32_t)(s >> (48 - bits));
}

__device__ __forceinline__ int32_t jr_next_int(unsigned long long &s, int32_t n)
{
    if ((n & -n) == n) return (int32_t)(((long long)n * (long long)jr_next(s, 31)) >> 31);
    int32_t bits, val;
    do {
        bits = jr_next(s, 31);
        val = bits % n;
    } while ((long long)bits - val + (n - 1) > 0x7FFFFFFFLL);   // Java: int overflow => redraw
    return val;
}

// One thread per distinct n: the 100 x k draws of java.util.Random(1).nextInt(n),
// stored as shared-memory byte offsets (row * 128, the row pitch of a 32-genus
// block) so the inner loop needs one IADD per draw.  Each replicate is padded to
// nb = ceil(k/4) batches of 4 draws with row n (the all-zero row).  The IL = 32/LPR
// replicates that the groups of one warp walk at the same time are interleaved
// batch by batch, so the warp's list load is one contiguous 16*IL-byte segment:
//     uint4 index of (task, batch) = ((task / IL) * nb + batch) * IL + task % IL
// Two zero-row batches per lane follow the last block for the pipelined read-ahead.
#define PG_ROW_PITCH 128u
__host__ __device__ static inline size_t pg_boot_list_entries(int k, int il)
{
    const int nb = (k + 3) >> 2;
    const int tblocks = (PG_NUM_BOOT + il - 1) / il;
    return ((size_t)tblocks * nb + 2) * il * 4;            // uint32 entries
}
static __global__ void k_boot_indices(const int32_t *__restrict__ ns, const int32_t *__restrict__ offs,
                               const int32_t *__restrict__ ils, int cnt, int min_boot,
                               uint32_t *__restrict__ pool)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    const int n = ns[t], il = ils[t];
    int k = n >> 3;
    if (k < min_boot) k = min_boot;
    const int nb = (k + 3) >> 2;
    uint32_t *out = pool + offs[t];
    const size_t total = pg_boot_list_entries(k, il);
    for (size_t e = 0; e < total; e++) out[e] = (uint32_t)n * PG_ROW_PITCH;
    unsigned long long s = (1ULL ^ JR_MULT) & JR_MASK;          // setSeed(1)
    for (int run = 0; run < PG_NUM_BOOT; run++) {
        const size_t base = ((size_t)(run / il) * nb * il + (run % il)) * 4;
        for (int j = 0; j < k; j++) {
            const uint32_t r = n > 0 ? (uint32_t)jr_next_int(s, n) : 0u;
            out[base + (size_t)(j >> 2) * il * 4 + (j & 3)] = r * PG_ROW_PITCH;
        }
    }
}

// ------------------------------------------------------------------ keys

// order-preserving map fp32 -> u32
__device__ __forceinline__ uint32_t pg_ord(float f)
{
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float pg_unord(uint32_t u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
// max over keys == larger score first, then SMALLER genus index (first strict max)
__device__ __forceinline__ unsigned long long pg_key(float score, uint32_t genus)
{
    return ((unsigned long long)pg_ord(score) << 32) | (unsigned long long)(0xFFFFFFFFu - genus);
}

__device__ __forceinline__ void pg_cp_async16(void *smem, const void *gmem)
{
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void pg_cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}


// Reads are bucketed by word count so each launch sizes its shared memory (and so
// its CTAs/SM) for the reads it actually carries.
struct Bucket { int nmax; int lpr; int block; };

// device buffers of certified mode (one chunk of reads)
struct PgCertBufs {
    unsigned long long *champ;       // [reads][101] champion slots: sum << 32 | table position
    unsigned int       *ncand;       // [reads]
    unsigned long long *cand;        // [reads][PG_CANDCAP] near-tie lists
    int32_t            *guess;       // [reads] block run in full
    unsigned long long *items;       // [item_cap] (read, block, task) still to evaluate exactly
    unsigned int        item_cap;
    int                 light_max;   // pg_classify_opts.light_max
    int                 bound_level; // pg_classify_opts.bound_level
    unsigned int       *item_count;  // cursor of the item list in flight (counters + 2; + 4 for the second half of a pipelined bucket)
    int                 stage;       // plan 4: 0 = all of phase 1, 1 = up to the tensor-core kernel, 2 = the item kernel only
    bool                retry;       // second try of reads the first guess left heavy: a more careful guess (k_guess_bm, 64 words)
    bool                count_mma;   // this pass counts for pg_classify_stats3 (not the one-off trials)
    int                 force_part;  // -1: the model's choice; 0 / 1: block / part columns in k_bound (the one-off trial of a model)
    unsigned int       *counters;    // [0] strict fallbacks, [1] heavy reads, [2] items of the bucket in flight, [3] items so far
    uint8_t            *heavy;       // [reads of the chunk] flag
    int32_t            *fb_list, *hv_list;   // [reads of the call] read indices, appended across chunks
};
