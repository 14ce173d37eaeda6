// pg_internal.cuh -- shared host/device plumbing of libpangea_b200.so.
// Not part of the ABI; the ABI is include/pangea_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <utility>

#include "pangea_b200.h"

struct pg_ctx {
    int          device;
    int          sm_count;
    // opt-in maximum of dynamic shared memory per block.  Every kernel that needs more than 48 KB is given THIS limit,
    // not the size of the launch at hand: the attribute is per function and device, and two host threads that share a
    // device (rdp_classifier runs two contexts per GPU) would otherwise lower it under each other's launches
    int          smem_optin;
    cudaStream_t stream;
    cudaStream_t own_stream;
    cudaStream_t copy_stream, down_stream;    // uploads / downloads of pg_classify(), created on first use
    cudaStream_t aux_stream;                  // plan 4: item kernel + phase 2 of one slice run here under the tensor-core kernel of the next
    cudaEvent_t  ev_pipe[4];                  // [0..1] stage 1 of slice parity p done, [2..3] stages 2 + phase 2 done
    mutable char err[512];
    int64_t      launches;

    // java.util.Random(1) sample lists, cached per n (row A8): pool of uint32 smem offsets,
    // h_boot_off[n] = offset of the list for n words (or -1), mirrored on device.
    uint32_t            *d_boot_pool;
    size_t               boot_cap, boot_used;     // uint32 units
    int32_t             *d_boot_off;              // [PG_MAX_WORDS+1]
    std::vector<int32_t> h_boot_off;
    int                  boot_min_words;

    // plan 4 (pg_mma.cu): draw-count images of the replicates, one slot per word count n <= 640
    uint8_t             *d_cnt_img;
    std::vector<char>    cnt_built;
    int                  cnt_min_words;

    // classify-kernel timing (CUDA events on the launching stream)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pending;
    std::vector<cudaEvent_t>                         ev_free;
    double  classify_ms;
    int64_t classify_launches;
    int64_t st_certified, st_strict, st_handed_back;   // routing of the last classify call

    // grow-only device scratch for classification
    struct Scratch {
        void  *p;
        size_t cap;
    } s_words, s_nwords, s_flags, s_best, s_order, s_results, s_boot, s_bytes, s_off, s_cand,
      s_champ, s_ncand, s_candl, s_fb, s_guess, s_items, s_heavy, s_hist, s_meta;
    static const int kNumScratch = 19;
    int64_t st_heavy, st_items;                        // certified v2: reads redone by the all-block kernel, light items
    int64_t st_mma;                                    // reads whose best part and bounds came from the tensor-core kernel (plan 4)
    // pinned host staging
    void  *h_pin;
    size_t h_pin_cap;
};

struct pg_model {
    pg_ctx  *ctx;
    int      G, ntile, depth;
    int64_t  N;                 // host copy, valid after commit
    int32_t *d_m;               // [ntile][65536][32]   A5 counts, genus-tiled
    int32_t *d_nw;              // [65536]
    int32_t *d_M;               // [ntile*32]
    unsigned long long *d_N;    // [1]
    float   *d_table;           // [ntile][65536][32]   A4/A6 dense log table, genus-tiled
    float   *d_logPrior;        // [65536]
    float   *d_pdiff;           // [65536]  logPrior[w] - logPrior[rc(w)]: the orientation test sums these (A3)
    float   *d_Pw;              // [65536]
    float   *d_logLeave;        // [ntile*32]
    int32_t *d_anc;             // [G][depth]
    // certified mode (pg_certified.cu): deficits below the per-word row maximum in units of
    // 2^-7 nat, 64 genera per 128-byte row segment
    uint16_t *d_qtable;         // [ntile64][65536][64]
    float   *d_rowmax;          // [65536]
    int      ntile64;           // 64-position blocks of the quantised table (>= ceil(G/64): clade-aligned blocks are padded)
    // certified v2: genera are laid out in the quantised table in LINEAGE order (relatives share a
    // 64-genus block), and every block has a per-word minimum used as a lower bound of all its genera
    int32_t  *d_perm;           // [ntile64*64] table position -> genus index (>= G: padding)
    unsigned long long *d_blockmask;   // [ntile64] bit i: position blk*64+i holds a genus
    uint16_t *d_bmtable;        // [ngroup][65536][32]   min over the 64 genera of block (group*32 + i)
    uint16_t *d_hmtable;        // [ngroup_h][65536][32]  min over each 32-position half block (2*block + half)
    // coarse first-level bound of models with more than one group of blocks (k_bound8): per word one byte per block,
    // min(bm >> PG_C8_SHIFT, PG_C8_MAX), row pitch bm8_pitch bytes (a multiple of 16); columns sib0 .. sib0+3 are left
    // zero in the table -- k_bound8 fills them per read with the part minima of the read's best block
    uint8_t  *d_bm8;            // [65536][bm8_pitch]
    int      bm8_pitch, sib0;
    // bound the table's 16-position parts instead of its 64-position blocks (k_bound<.., true>)?  Decided once per model
    // by timing both on the first chunk of reads it classifies (results do not depend on it); mutable for that reason
    mutable bool part_bounds, bounds_tuned;
    // plan 4 (pg_mma.cu): byte planes the tensor-core kernel gathers its B operand from
    uint8_t  *d_qx;             // [ntile64 * 4][65536][32]  low bytes | high bytes of a part's 16 deficits
    uint8_t  *d_bm8x;           // [65536][pitch8]           min(bm >> 2, 255) per block
    uint8_t  *d_hm8x;           // [65536][hpitch]           min(hm >> 2, 255) per part (4 * block + part)
    int      pitch8, hpitch;
    int      x8_blocks;         // ntile64 the three tables were built for (0: none)
    int      ngroup;            // ceil(ntile64 / 31): slot 31 of a bm row is spare (k_bound, plan 3)
    int      ngroup_h;          // ceil(2*ntile64 / 32)
    double   vmax;              // max |table entry| over real genera (fp32 error bound)
    bool     q_ok;              // every deficit fits the 12-bit field: certificates are valid
    bool     committed;
    bool     tables_only;       // built by pg_model_from_tables: log tables given, no counts behind them
};

struct pg_reads {
    pg_ctx   *ctx;
    int64_t   count;
    int64_t   total_bytes;
    int64_t  *d_off;            // [count+1] byte offsets of the original records
    uint32_t *d_planes;         // 3 x uint32 per 32-base chunk: {lo, hi, valid}
    int64_t   nchunks_cap;
};

// chunk index (32 bases per chunk) at which read i starts in the plane store:
// floor(off[i]/32) + i never overlaps the previous read and needs no scan.
__host__ __device__ static inline int64_t pg_chunk_start(int64_t off_i, int64_t i)
{
    return (off_i >> 5) + i;
}

int  pg_fail(const pg_ctx *ctx, int code, const char *fmt, ...);
int  pg_scratch(pg_ctx *ctx, pg_ctx::Scratch *s, size_t bytes);
int  pg_pinned(pg_ctx *ctx, size_t bytes);
// d_out[0..n) = exclusive scan of d_in, d_out[n] = total; synchronises the stream
int  pg_device_scan(pg_ctx *ctx, const int64_t *d_in, int64_t n, int64_t *d_out);
int  pg_pack_launch(pg_ctx *ctx, const char *d_bytes, const int64_t *d_off, int64_t count, uint32_t *d_planes);
int  pg_prior_diff_launch(pg_ctx *ctx, pg_model *md);        // pg_reads.cu: d_pdiff from d_logPrior

// A copy between host and device ORDERED WITH the context's stream: every kernel of the library runs on ctx->stream,
// which is created non-blocking, so a plain cudaMemcpy (legacy stream) is ordered with nothing.  Enqueued on
// ctx->stream and waited for, so pageable/local host buffers may be reused or freed on return.
static inline cudaError_t pg_copy_sync(const pg_ctx *ctx, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind)
{
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(ctx->stream);
}

// Transient device buffers: the stream-ordered allocator (cudaMallocAsync on the context's stream, pool kept warm by
// pg_init) instead of cudaMalloc / cudaFree -- cudaFree synchronises the WHOLE device, which serialises the contexts
// that share a GPU in a host pipeline (rdp_classifier runs two per device).
static inline cudaError_t pg_dev_alloc(const pg_ctx *ctx, void **p, size_t bytes)
{
    return cudaMallocAsync(p, bytes ? bytes : 1, ctx->stream);
}
static inline void pg_dev_free(const pg_ctx *ctx, void *p)
{
    if (p) cudaFreeAsync(p, ctx->stream);
}

// every launch of a kernel with more than 48 KB of dynamic shared memory is preceded by this: the function's limit is
// raised to ALL the device allows beside its static shared memory -- the same value from every thread, see smem_optin
template <typename K>
static inline cudaError_t pg_smem_unlock(const pg_ctx *ctx, K kernel)
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin - (int)a.sharedSizeBytes);
}

#define PG_CUDA(ctx, call)                                                              \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess)                                                         \
            return pg_fail((ctx), PG_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,  \
                           cudaGetErrorString(e__));                                    \
    } while (0)

#define PG_LAUNCHED(ctx)                                                                \
    do {                                                                                \
        (ctx)->launches++;                                                              \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess)                                                         \
            return pg_fail((ctx), PG_ECUDA, "%s:%d launch: %s", __FILE__, __LINE__,     \
                           cudaGetErrorString(e__));                                    \
    } while (0)

#define PG_TRY(call)                                                                    \
    do {                                                                                \
        int r__ = (call);                                                               \
        if (r__ != PG_OK) return r__;                                                   \
    } while (0)
