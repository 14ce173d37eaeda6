// tc_probe.cu -- stand-alone check of the tcgen05 pieces the bound kernel (pg_mma.cu) is built from.
//   D[128 x N] (int32, TMEM) = A[128 x K] (u8, K-major, no swizzle) x B[K x N] (u8, MN-major, no swizzle, gathered rows)
// two byte planes (hi, lo) of 12-bit table values, two accumulators, result 256*Dh + Dl compared with a CPU sum.
// Also times the three phases (row gather by cp.async, MMA issue -> commit, TMEM read-out) with clock64 on every SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu ; run: ./tc_probe [N] [n] [iters]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t s, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16_ca(uint32_t s, const void *g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ULL << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity)
{
    for (int spin = 0; spin < (1 << 22); spin++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
#define TMEM_LD16(taddr, v)                                                                                       \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),     \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
                 : "r"(taddr))

// tabH/tabL: [65536][N] byte planes; words: [nreads][n]; Aimg: [Kpad/16][128][16]; out: [nreads][128][N]
__global__ void __launch_bounds__(1024, 1)
k_probe(const uint8_t *__restrict__ tabH, const uint8_t *__restrict__ tabL, const uint16_t *__restrict__ words, int n, int Kpad, int N,
        const uint8_t *__restrict__ Aimg, const uint8_t *__restrict__ Arow, int32_t *__restrict__ out, int reads_per_cta, int variant, long long *__restrict__ cyc,
        int *__restrict__ fail)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ uint16_t sw[1024];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *sA = smem, *sBh = smem + (size_t)128 * Kpad, *sBl = sBh + (size_t)Kpad * N;
    const bool ts = (variant & 4) != 0, timing = (variant & 8) != 0, interleave = (variant & 16) != 0, quiet = (variant & 32) != 0;
    const int gm = (variant >> 8) & 3;
    const uint32_t need = 2 * N + (ts ? Kpad / 4 : 0);
    const uint32_t ncols = need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512)));

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int i = tid; i < 128 * Kpad / 16; i += blockDim.x) cp_async16(smem_u32(sA + (size_t)i * 16), Aimg + (size_t)i * 16);
    cp_async_wait_all();
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t tmemA = tmem + 2u * (uint32_t)N;
    if (ts) {
        // A into tensor memory: lane = row, column = four consecutive K bytes; Arow is the plain [128][Kpad] matrix
        if (warp < 4) {
            const uint32_t *arow = reinterpret_cast<const uint32_t *>(Arow + (size_t)(warp * 32 + lane) * Kpad);
            for (int c = 0; c < Kpad / 4; c += 8) {
                uint32_t v[8];
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = arow[c + i];
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(tmemA + ((uint32_t)(warp * 32) << 16) + (uint32_t)c),
                             "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    }
    const uint32_t idesc = (2u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const int nchunk = N / 16;
    uint32_t phase = 0;
    long long t_gather = 0, t_mma = 0, t_epi = 0;

    for (int r = 0; r < reads_per_cta; r++) {
        const int read = blockIdx.x * reads_per_cta + r;
        const uint16_t *w = words + (size_t)read * n;
        for (int j = tid; j < n; j += blockDim.x) sw[j] = w[j];
        __syncthreads();
        long long c0 = clock64();
        // ---- gather: plane[c][j][16] <- tab[w_j][16c .. 16c+15]
        if (gm == 3) {
            // 32 bytes per load (LDG.256), two 16-byte stores: do requests or bytes bound the gather?
            const int total = n * (nchunk / 2);
            for (int i0 = tid; i0 < total; i0 += 2 * blockDim.x) {
                uint32_t vh[2][8], vl[2][8];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int i = i0 + u * blockDim.x;
                    if (i < total) {
                        const int j = i / (nchunk / 2), c2 = i - j * (nchunk / 2);
                        const size_t src = (size_t)sw[j] * N + c2 * 32;
                        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(vh[u][0]), "=r"(vh[u][1]), "=r"(vh[u][2]), "=r"(vh[u][3]),
                                     "=r"(vh[u][4]), "=r"(vh[u][5]), "=r"(vh[u][6]), "=r"(vh[u][7]) : "l"(tabH + src));
                        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(vl[u][0]), "=r"(vl[u][1]), "=r"(vl[u][2]), "=r"(vl[u][3]),
                                     "=r"(vl[u][4]), "=r"(vl[u][5]), "=r"(vl[u][6]), "=r"(vl[u][7]) : "l"(tabL + src));
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int i = i0 + u * blockDim.x;
                    if (i < total) {
                        const int j = i / (nchunk / 2), c2 = i - j * (nchunk / 2);
                        const uint32_t d0 = (uint32_t)((2 * c2) * Kpad + j) * 16, d1 = (uint32_t)((2 * c2 + 1) * Kpad + j) * 16;
                        *reinterpret_cast<uint4 *>(sBh + d0) = make_uint4(vh[u][0], vh[u][1], vh[u][2], vh[u][3]);
                        *reinterpret_cast<uint4 *>(sBh + d1) = make_uint4(vh[u][4], vh[u][5], vh[u][6], vh[u][7]);
                        *reinterpret_cast<uint4 *>(sBl + d0) = make_uint4(vl[u][0], vl[u][1], vl[u][2], vl[u][3]);
                        *reinterpret_cast<uint4 *>(sBl + d1) = make_uint4(vl[u][4], vl[u][5], vl[u][6], vl[u][7]);
                    }
                }
            }
        } else if (gm == 2) {
            const int total = n * nchunk;
            for (int i0 = tid; i0 < total; i0 += 4 * blockDim.x) {
                uint4 vh[4], vl[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = i0 + u * blockDim.x;
                    if (i < total) {
                        const int j = i / nchunk, c = i - j * nchunk;
                        const size_t src = (size_t)sw[j] * N + c * 16;
                        vh[u] = __ldg(reinterpret_cast<const uint4 *>(tabH + src));
                        vl[u] = __ldg(reinterpret_cast<const uint4 *>(tabL + src));
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = i0 + u * blockDim.x;
                    if (i < total) {
                        const int j = i / nchunk, c = i - j * nchunk;
                        const uint32_t dst = (uint32_t)(c * Kpad + j) * 16;
                        *reinterpret_cast<uint4 *>(sBh + dst) = vh[u];
                        *reinterpret_cast<uint4 *>(sBl + dst) = vl[u];
                    }
                }
            }
        } else {
            for (int i = tid; i < n * nchunk; i += blockDim.x) {
                const int j = i / nchunk, c = i - j * nchunk;
                const size_t src = (size_t)sw[j] * N + c * 16;
                const uint32_t dst = (uint32_t)(c * Kpad + j) * 16;
                if (gm == 1) { cp_async16_ca(smem_u32(sBh) + dst, tabH + src); cp_async16_ca(smem_u32(sBl) + dst, tabL + src); }
                else { cp_async16(smem_u32(sBh) + dst, tabH + src); cp_async16(smem_u32(sBl) + dst, tabL + src); }
            }
            cp_async_wait_all();
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();
        long long c1 = clock64();
        // ---- MMA: one thread
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t a_lbo = (variant & 1) ? 128u : 2048u, a_sbo = (variant & 1) ? 2048u : 128u;
            const uint32_t b_lbo = (variant & 2) ? (uint32_t)Kpad * 16u : 128u, b_sbo = (variant & 2) ? 128u : (uint32_t)Kpad * 16u;
            for (int it = 0; it < 2 * (Kpad / 32); it++) {
                const int p = interleave ? (it & 1) : it / (Kpad / 32);
                const int ks = interleave ? (it >> 1) : it % (Kpad / 32);
                const uint32_t bbase = smem_u32(p ? sBl : sBh);
                {
                    const uint64_t ad = make_desc(smem_u32(sA) + (uint32_t)ks * 4096u, a_lbo, a_sbo);
                    const uint64_t bd = make_desc(bbase + (uint32_t)ks * 512u, b_lbo, b_sbo);
                    if (ts) mma_i8_ts(tmem + (uint32_t)(p * N), tmemA + (uint32_t)ks * 8u, bd, idesc, ks > 0 ? 1u : 0u);
                    else mma_i8(tmem + (uint32_t)(p * N), ad, bd, idesc, ks > 0 ? 1u : 0u);
                }
            }
            mma_commit(smem_u32(&bar));
        }
        if (quiet) {
            if (warp == 0 && !mbar_wait(smem_u32(&bar), phase)) { if (tid == 0) atomicAdd(fail, 1); }
            __syncthreads();
        } else if (!mbar_wait(smem_u32(&bar), phase)) { if (tid == 0) atomicAdd(fail, 1); break; }
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        long long c2 = clock64();
        // ---- read-out: warp q of 0..3 owns TMEM lanes 32q..32q+31 (thread = row of D)
        if (warp < 4) {
            const int row = warp * 32 + lane;
            int32_t *o = out + ((size_t)read * 128 + row) * N;
            uint32_t chk = 0u;
            for (int c = 0; c < nchunk; c++) {
                uint32_t vh[16], vl[16];
                const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 16);
                TMEM_LD16(ta, vh);
                TMEM_LD16(ta + (uint32_t)N, vl);
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const uint32_t v = vh[i] * 256u + vl[i];
                    if (timing) chk += v <= 1000u ? 1u : 0u; else o[c * 16 + i] = (int32_t)v;
                }
            }
            if (timing && chk == 0xFFFFFFFFu) o[0] = 1;
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        long long c3 = clock64();
        t_gather += c1 - c0; t_mma += c2 - c1; t_epi += c3 - c2;
    }
    if (tid == 0) { cyc[blockIdx.x * 3 + 0] = t_gather; cyc[blockIdx.x * 3 + 1] = t_mma; cyc[blockIdx.x * 3 + 2] = t_epi; }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(ncols) : "memory");
}

int main(int argc, char **argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 48, n = argc > 2 ? atoi(argv[2]) : 486, rpc = argc > 3 ? atoi(argv[3]) : 8;
    const int ncta = argc > 4 ? atoi(argv[4]) : 148, nthr = argc > 5 ? atoi(argv[5]) : 256;
    const int Kpad = (n + 31) & ~31, nreads = ncta * rpc;
    printf("probe: N=%d n=%d Kpad=%d reads/cta=%d ctas=%d\n", N, n, Kpad, rpc, ncta);
    std::vector<uint8_t> tH((size_t)65536 * N), tL((size_t)65536 * N), A((size_t)128 * Kpad, 0), Aimg((size_t)128 * Kpad, 0);
    std::vector<uint16_t> words((size_t)nreads * n);
    srand(12345);
    for (size_t i = 0; i < tH.size(); i++) { const int q = rand() % 4096; tH[i] = (uint8_t)(q >> 8); tL[i] = (uint8_t)(q & 255); }
    for (size_t i = 0; i < words.size(); i++) words[i] = (uint16_t)(rand() & 0xFFFF);
    for (int j = 0; j < n; j++) A[j] = 1;                                   // row 0: every word once
    for (int t = 1; t <= 100; t++)
        for (int d = 0; d < n / 8; d++) A[(size_t)t * Kpad + rand() % n]++;   // rows 1..100: n/8 draws with replacement
    for (int t = 101; t < 128; t++)
        for (int j = 0; j < Kpad; j++) A[(size_t)t * Kpad + j] = (uint8_t)(rand() % 7);   // junk rows incl. the K padding
    for (int m = 0; m < 128; m++)
        for (int k = 0; k < Kpad; k++) Aimg[((size_t)(k / 16) * 128 + m) * 16 + k % 16] = A[(size_t)m * Kpad + k];
    uint8_t *dH, *dL, *dA, *dAr; uint16_t *dW; int32_t *dOut; long long *dCyc; int *dFail;
    CK(cudaMalloc(&dH, tH.size())); CK(cudaMalloc(&dL, tL.size())); CK(cudaMalloc(&dA, Aimg.size())); CK(cudaMalloc(&dAr, A.size())); CK(cudaMemcpy(dAr, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dW, words.size() * 2)); CK(cudaMalloc(&dOut, (size_t)nreads * 128 * N * 4));
    CK(cudaMalloc(&dCyc, ncta * 3 * 8)); CK(cudaMalloc(&dFail, 4));
    CK(cudaMemcpy(dH, tH.data(), tH.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dL, tL.data(), tL.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA, Aimg.data(), Aimg.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dW, words.data(), words.size() * 2, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)128 * Kpad + 2 * (size_t)Kpad * N + 1024;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // CPU reference for a sample of reads
    std::vector<int32_t> out((size_t)nreads * 128 * N);
    const int variants[4] = {0, 8 + 16 + 32, 768, 8 + 16 + 32 + 768};
    for (int vi = 0; vi < 4; vi++) {
        const int variant = variants[vi];
        if ((variant & 4) && 2 * N + Kpad / 4 > 512) { printf("variant %d: A does not fit tensor memory beside D\n", variant); continue; }
        CK(cudaMemset(dOut, 0xEE, out.size() * 4)); CK(cudaMemset(dFail, 0, 4));
        k_probe<<<ncta, nthr, smem>>>(dH, dL, dW, n, Kpad, N, dA, dAr, dOut, rpc, variant, dCyc, dFail);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: kernel error %s\n", variant, cudaGetErrorString(e)); return 2; }
        int fail = 0; CK(cudaMemcpy(&fail, dFail, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
        long long good = 0, total = 0;
        const int sample[4] = {0, 1, nreads / 2, nreads - 1};
        for (int s = 0; s < 4; s++) {
            const int r = sample[s];
            for (int t = 0; t <= 100; t++)
                for (int c = 0; c < N; c++) {
                    long long acc = 0;
                    for (int j = 0; j < n; j++) {
                        const size_t cell = (size_t)words[(size_t)r * n + j] * N + c;
                        acc += (long long)A[(size_t)t * Kpad + j] * (tH[cell] * 256 + tL[cell]);
                    }
                    good += (long long)out[((size_t)r * 128 + t) * N + c] == acc;
                    total++;
                }
        }
        std::vector<long long> cyc(ncta * 3);
        CK(cudaMemcpy(cyc.data(), dCyc, cyc.size() * 8, cudaMemcpyDeviceToHost));
        double g = 0, m = 0, ep = 0;
        for (int i = 0; i < ncta; i++) { g += cyc[i * 3]; m += cyc[i * 3 + 1]; ep += cyc[i * 3 + 2]; }
        printf("variant %3d (A in %s, %s, planes %s, wait %s, gather %d): match %lld / %lld, barrier timeouts %d; cycles per read per CTA: gather %.0f  mma %.0f  readout %.0f\n",
               variant, (variant & 4) ? "tmem" : "smem", (variant & 8) ? "timing only" : "checked", (variant & 16) ? "interleaved" : "in turn", (variant & 32) ? "quiet" : "all spin", (variant >> 8) & 3, good, total, fail,
               g / ncta / rpc, m / ncta / rpc, ep / ncta / rpc);
    }
    return 0;
}
