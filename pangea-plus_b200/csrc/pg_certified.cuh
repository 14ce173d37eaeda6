// pg_certified.cuh -- constants and device helpers shared by the certified-mode kernels (pg_certified.cu, pg_mma.cu).
#pragma once
#include "pg_classify_common.cuh"

#define PG_Q_SCALE 128.0            // units per nat
#define PG_Q_MAX   4095             // 12-bit field: 16 rows x 4095 < 65536
#define PG_CANDCAP 128              // near-tie entries kept per read
#define PG_CHAMP_INIT 0xFFFFFFFFFFFFFFFFULL

#define PG_PARTS 4                       // plan 3 evaluates one part (64 / PG_PARTS positions) of the best block
#define PG_PART_POS (64 / PG_PARTS)
#define PG_GB (32 - PG_PARTS)

#define PG_LIGHT_MAX 768            // items per (read, group) above which the read is "heavy"
#define PG_ITEM_NULL 0xFFFFu

// margin in quantisation units: terms (floor error) + 2E*128 (fp32 order error) + 1
__device__ __forceinline__ uint32_t pg_margin(int terms, double vmax)
{
    if (terms <= 1) return (uint32_t)terms + 1u;
    const double u = 5.9604644775390625e-08;                 // 2^-24
    const double m = (double)(terms - 1) * u;
    const double E = m / (1.0 - m) * (double)terms * vmax * 1.0001;
    return (uint32_t)terms + (uint32_t)ceil(2.0 * E * PG_Q_SCALE) + 1u;
}

__device__ __forceinline__ void pg_emit(unsigned int *ncand, unsigned long long *cand, int task, uint32_t genus,
                                        uint32_t sum)
{
    const unsigned int slot = atomicAdd(ncand, 1u);
    if (slot < PG_CANDCAP)
        cand[slot] = ((unsigned long long)task << 56) | ((unsigned long long)genus << 32) | sum;
}
