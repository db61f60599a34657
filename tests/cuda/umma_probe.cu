// tcgen05 / TMEM / TMA building-block probe (test infrastructure for the v2 kernels).
//
//   umma_probe <test> <variant>
//     test 0: D[128,128] = A[128,K] * B[128,K]^T      both K-major, TMA SWIZZLE_128B tiles, kind::f16, K=256
//     test 1: same as 0 but as the split product Ah*Bh + Ah*Bl + Al*Bh (three accumulating MMA passes)
//     test 2: D[128,32]  = P[128,K] * V[K,32]          A K-major SW128, B MN-major ([K][32] row-major) SWIZZLE_64B
//     test 3: D[128,128] = Q[128,32] * K[128,32]^T     both K-major with 64-byte rows (SWIZZLE_64B), K=32
//   variant selects descriptor field guesses (LBO/SBO/k-step) for the layouts I could not pin from headers.
// Prints "PROBE test=.. variant=.. max_err=.. PASS|FAIL".  Exit code 0 on PASS.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);    \
            exit(2);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum));
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

struct ProbeParams {
    int n;             // UMMA N (columns of D)
    int kblocks;       // number of TMA k-blocks
    int kb_elems;      // K elements per k-block
    int passes;        // 1, or 3 for the split product
    uint32_t idesc;
    uint32_t a_hi, b_hi;       // upper 32 bits of the smem descriptors (SBO, version, layout type)
    uint32_t a_lbo, b_lbo;     // encoded LBO (>>4)
    uint32_t a_kstep, b_kstep; // bytes added to the start address per UMMA_K=16 step
    uint32_t a_bytes, b_bytes; // TMA bytes per tile
    int a_box_k0, b_is_mn;     // unused / B coordinate order
};

// grid 1, block 128
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                                                    const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
                                                    float* __restrict__ D, ProbeParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                  // 16 KB (1024-aligned)
    uint8_t* sB = smem + 16384;          // 16 KB
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_tma = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        mbar_init(bar_tma, 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    uint32_t phase = 0;
    int first = 1;
    for (int pass = 0; pass < p.passes; ++pass) {
        // split product passes: (Al,Bh), (Ah,Bl), (Ah,Bh)
        const CUtensorMap* mA = (p.passes == 3 && pass == 0) ? &mapAl : &mapAh;
        const CUtensorMap* mB = (p.passes == 3 && pass == 1) ? &mapBl : &mapBh;
        for (int kb = 0; kb < p.kblocks; ++kb) {
            if (tid == 0) {
                mbar_expect_tx(bar_tma, p.a_bytes + p.b_bytes);
                tma_load_2d(smem_u32(sA), mA, bar_tma, kb * p.kb_elems, 0);
                if (p.b_is_mn) tma_load_2d(smem_u32(sB), mB, bar_tma, 0, kb * p.kb_elems);
                else           tma_load_2d(smem_u32(sB), mB, bar_tma, kb * p.kb_elems, 0);
            }
            mbar_wait(bar_tma, phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == 0) {
                const uint64_t da0 = ((uint64_t)p.a_hi << 32) | ((uint64_t)p.a_lbo << 16) | ((smem_u32(sA) >> 4) & 0x3FFF);
                const uint64_t db0 = ((uint64_t)p.b_hi << 32) | ((uint64_t)p.b_lbo << 16) | ((smem_u32(sB) >> 4) & 0x3FFF);
                for (int k = 0; k < p.kb_elems / 16; ++k) {
                    umma_f16(tmem, da0 + (uint64_t)((k * p.a_kstep) >> 4), db0 + (uint64_t)((k * p.b_kstep) >> 4), p.idesc,
                             first ? 0u : 1u);
                    first = 0;
                }
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, phase);
            first = 0;
            phase ^= 1;
            __syncthreads();
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < p.n; c0 += 32) {
        uint32_t v[32];
        const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) D[(size_t)(warp * 32 + lane) * p.n + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}


// ---- test 4: TS-mode PV.  P (fp32 in global) -> split -> tcgen05.st to TMEM (packed halves), V MN-major via TMA.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum));
}
__device__ __forceinline__ void tmem_st32(uint32_t addr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
          "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),
          "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld32p(uint32_t addr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr));
}

// grid 1, block 128.  P [128][128] fp32 row-major; V planes [128][32] fp16 via TMA (box 32 x 128, SWIZZLE_64B); D [128][32]
__global__ void __launch_bounds__(128) probe_ts_kernel(const __grid_constant__ CUtensorMap mapVh, const __grid_constant__ CUtensorMap mapVl,
                                                       const float* __restrict__ P, float* __restrict__ D, int concat, uint32_t lbo_bytes) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_tma = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) { mbar_init(bar_tma, 1); mbar_init(bar_mma, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        mbar_expect_tx(bar_tma, 2 * 128 * 64);
        tma_load_2d(smem_u32(smem), &mapVh, bar_tma, 0, 0);
        tma_load_2d(smem_u32(smem) + 8192, &mapVl, bar_tma, 0, 0);
    }
    // every thread owns one row of P: split into packed fp16 hi / lo and store to TMEM columns [0,64) / [64,128)
    const float* prow = P + (size_t)tid * 128;
    for (int half = 0; half < 2; ++half) {      // 64 elements = 32 packed columns per tcgen05.st
        uint32_t hi[32], lo[32];
        for (int i = 0; i < 32; ++i) {
            const float x0 = prow[half * 64 + 2 * i], x1 = prow[half * 64 + 2 * i + 1];
            __half2 h = __floats2half2_rn(x0, x1);
            float2 hf = __half22float2(h);
            __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
            hi[i] = *reinterpret_cast<uint32_t*>(&h);
            lo[i] = *reinterpret_cast<uint32_t*>(&l);
        }
        tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + half * 32, hi);
        tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 64 + half * 32, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    mbar_wait(bar_tma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // N=32, B MN-major
        const uint32_t hi_bits = ((512u >> 4) & 0x3FFF) | (1u << 14) | (4u << 29);
        const uint64_t dvh = ((uint64_t)hi_bits << 32) | (1ull << 16) | ((smem_u32(smem) >> 4) & 0x3FFF);
        const uint64_t dvl = ((uint64_t)hi_bits << 32) | (1ull << 16) | (((smem_u32(smem) + 8192) >> 4) & 0x3FFF);
        const uint32_t d_t = tmem + 128;
        if (!concat) {
            for (int k = 0; k < 8; ++k) {           // 16 kv rows per slice: A advances 8 columns, V advances 1024 bytes
                const uint64_t o = (uint64_t)((k * 1024) >> 4);
                umma_f16_ts(d_t, tmem + 64 + k * 8, dvh + o, idesc, k != 0);     // Pl * Vh
                umma_f16_ts(d_t, tmem + k * 8, dvl + o, idesc, 1u);             // Ph * Vl
                umma_f16_ts(d_t, tmem + k * 8, dvh + o, idesc, 1u);             // Ph * Vh
            }
        } else {
            // [Vh | Vl] as one MN-major B operand of N = 64: second 32-column atom at LBO bytes (the Vl tile)
            const uint32_t idesc64 = (1u << 4) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t dcat = ((uint64_t)hi_bits << 32) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((smem_u32(smem) >> 4) & 0x3FFF);
            for (int k = 0; k < 8; ++k) {
                const uint64_t o = (uint64_t)((k * 1024) >> 4);
                umma_f16_ts(d_t, tmem + k * 8, dcat + o, idesc64, k != 0);      // Ph * [Vh | Vl]  -> D[:, 0:64]
                umma_f16_ts(d_t, tmem + 64 + k * 8, dvh + o, idesc, 1u);        // Pl * Vh         -> D[:, 0:32]
            }
        }
        umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32], v2[32];
    tmem_ld32p(tmem + ((uint32_t)(warp * 32) << 16) + 128, v);
    tmem_ld32p(tmem + ((uint32_t)(warp * 32) << 16) + 160, v2);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) D[(size_t)tid * 32 + i] = __uint_as_float(v[i]) + (concat ? __uint_as_float(v2[i]) : 0.f);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// ---- test 5: per-SM rate micro-benchmarks (cycles via clock64); grid = 148 CTAs so every SM is busy.
//   mode 0: tcgen05.ld x32 by 4 warps   mode 1: tcgen05.st x32 by 4 warps
//   mode 2: SS MMA M128 N128 K16        mode 3: SS MMA M128 N32 K16 (B MN-major)   mode 4: TS MMA M128 N32 K16
//   mode 5: SS MMA M128 N64 K16
__global__ void __launch_bounds__(128) probe_rate_kernel(int mode, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[1];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar = smem_u32(&bars[0]);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    long long t0 = clock64();
    if (mode == 0) {
        uint32_t acc = 0;
        for (int it = 0; it < iters; ++it) {
            uint32_t v[32];
            tmem_ld32p(lane_base + (it & 7) * 32, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[0] + v[31];
        }
        if (acc == 0x12345678u) out[1000] = acc;
    } else if (mode == 1) {
        uint32_t v[32];
        for (int i = 0; i < 32; ++i) v[i] = tid + i;
        for (int it = 0; it < iters; ++it) tmem_st32(lane_base + (it & 7) * 32, v);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else {
        if (tid == 0) {
            // modes >= 6: interleave independent accumulators / other shapes
            //  6: SS N128, 2 accumulators   7: TS N32, 4 accumulators   8: SS N256   9: SS N32 (MN), 4 accumulators
            // 10: SS N128 accumulate=0     11: TS N32, 8 accumulators  12: SS N128 + TS N32 alternating (attention mix)
            int N = 128; bool mn = false, ts = false; int nacc = 1, dstride = 0; uint32_t accf = 1u;
            switch (mode) {
                case 2: N = 128; break;
                case 3: N = 32; mn = true; break;
                case 4: N = 32; mn = true; ts = true; break;
                case 5: N = 64; break;
                case 6: N = 128; nacc = 2; dstride = 128; break;
                case 7: N = 32; mn = true; ts = true; nacc = 4; dstride = 32; break;
                case 8: N = 256; break;
                case 9: N = 32; mn = true; nacc = 4; dstride = 32; break;
                case 10: N = 128; accf = 0u; break;
                case 11: N = 32; mn = true; ts = true; nacc = 8; dstride = 32; break;
                default: break;
            }
            const uint32_t hi_bits = ((512u >> 4) & 0x3FFF) | (1u << 14) | (4u << 29);
            const uint64_t da = ((uint64_t)hi_bits << 32) | (1ull << 16) | ((smem_u32(smem) >> 4) & 0x3FFF);
            const uint64_t db = ((uint64_t)hi_bits << 32) | (1ull << 16) | (((smem_u32(smem) + 16384) >> 4) & 0x3FFF);
            if (mode == 12) {
                const uint32_t id128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                const uint32_t id32 = (1u << 4) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                for (int it = 0; it < iters; it += 2) {
                    umma_f16(tmem + 256 + ((it >> 1) & 1) * 128, da, db, id128, 1u);
                    umma_f16_ts(tmem + 128 + ((it >> 1) & 3) * 32, tmem + (it & 7) * 8, db, id32, 1u);
                }
            } else {
                const uint32_t idesc = (1u << 4) | (mn ? (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                for (int it = 0; it < iters; ++it) {
                    const uint32_t d = tmem + 256 + (uint32_t)((it % nacc) * dstride);
                    if (ts) umma_f16_ts(d, tmem + (it & 7) * 8, db, idesc, accf);
                    else    umma_f16(d, da, db, idesc, accf);
                }
            }
            umma_commit(bar);
        }
        mbar_wait(bar, 0);
    }
    long long t1 = clock64();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}


// ---- test 6: MMA issue-rate benchmark with compile-time shapes, 8 MMAs unrolled per iteration (like a real k-loop)
template <int N, bool TS, bool MN, int NACC>
__global__ void __launch_bounds__(128) probe_rate2_kernel(int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[1];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar = smem_u32(&bars[0]);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid; i < 49152 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    long long t0 = clock64();
    if (tid == 0) {
        constexpr uint32_t idesc = (1u << 4) | (MN ? (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t hi_bits = ((512u >> 4) & 0x3FFF) | (1u << 14) | (4u << 29);
        const uint64_t da = ((uint64_t)hi_bits << 32) | (1ull << 16) | ((smem_u32(smem) >> 4) & 0x3FFF);
        const uint64_t db = ((uint64_t)hi_bits << 32) | (1ull << 16) | (((smem_u32(smem) + 16384) >> 4) & 0x3FFF);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t d = tmem + 256 + (uint32_t)((u % NACC) * (N <= 64 ? 32 : 128));
                if constexpr (TS) umma_f16_ts(d, tmem + u * 8, db + (uint64_t)(u & 1) * 2, idesc, 1u);
                else              umma_f16(d, da + (uint64_t)(u & 1) * 2, db + (uint64_t)(u & 1) * 2, idesc, 1u);
            }
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    long long t1 = clock64();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N, bool TS, bool MN, int NACC>
static void run_rate2(const char* name, long long* dout) {
    CK(cudaFuncSetAttribute(probe_rate2_kernel<N, TS, MN, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024));
    const int iters = 500;
    probe_rate2_kernel<N, TS, MN, NACC><<<148, 128, 49152 + 1024>>>(iters, dout);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(148);
    CK(cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (auto v : h) avg += (double)v; avg /= 148.0;
    printf("RATE2 %-40s cycles/MMA=%.1f  (floor %d)\n", name, avg / (iters * 8.0), 128 * N / 256);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
    return (EncodeFn)fn;
}

// 2D fp16 row-major [rows][cols] tensor, box [box_rows][box_cols]
static CUtensorMap make_map(EncodeFn enc, void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols,
                            CUtensorMapSwizzle sw) {
    CUtensorMap m;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(2); }
    return m;
}


// ---- test 7: TMA streaming rate per SM.  Each CTA streams `iters` stages of NLOADS x [128 rows x 64 B] tiles through a
// DEPTH-deep ring (no consumer work: a stage is recycled as soon as it lands).  mode 0: all CTAs read the same rows
// (L2-resident), mode 1: every CTA reads its own rows of a large tensor (HBM stream).
template <int DEPTH, int NLOADS>
__global__ void __launch_bounds__(64) probe_tma_kernel(const __grid_constant__ CUtensorMap map, int iters, int mode, int rows_total,
                                                       long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[DEPTH];
    const int tid = threadIdx.x;
    if (tid == 0) { for (int i = 0; i < DEPTH; ++i) mbar_init(smem_u32(&bars[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    long long t0 = clock64();
    if (tid == 0) {
        const uint32_t sb = (smem_u32(smem) + 1023u) & ~1023u;
        int row_base = mode == 0 ? 0 : (int)(((long long)blockIdx.x * 8191) % (rows_total / 128 - NLOADS)) * 128;
        for (int it = 0; it < iters + DEPTH; ++it) {
            const int s = it % DEPTH;
            if (it >= DEPTH) mbar_wait(smem_u32(&bars[s]), ((it / DEPTH) - 1) & 1);
            if (it < iters) {
                mbar_expect_tx(smem_u32(&bars[s]), NLOADS * 8192);
                for (int l = 0; l < NLOADS; ++l) {
                    int row = row_base + l * 128;
                    if (mode == 1) { row_base = (row_base + 128 * NLOADS * 37) % (rows_total - 128 * NLOADS); row = row_base + l * 128; }
                    tma_load_2d(sb + s * (NLOADS * 8192) + l * 8192, &map, smem_u32(&bars[s]), 0, row);
                }
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
}

template <int DEPTH, int NLOADS>
static void run_tma(EncodeFn enc, __half* buf, int rows_total, int mode, long long* dout, bool wide = false) {
    // wide: view the same memory as [rows/2][64 halves] and fetch [64 rows x 128 B] boxes (also 8 KB each)
    CUtensorMap m = wide ? make_map(enc, buf, rows_total / 2, 64, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)
                         : make_map(enc, buf, rows_total, 32, 128, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (wide) { printf("[128B-wide boxes] "); rows_total /= 2; }
    const int smem = DEPTH * NLOADS * 8192 + 1024;
    CK(cudaFuncSetAttribute(probe_tma_kernel<DEPTH, NLOADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 400;
    probe_tma_kernel<DEPTH, NLOADS><<<148, 64, smem>>>(m, iters, mode, rows_total, dout);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(148);
    CK(cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (auto v : h) avg += (double)v; avg /= 148.0;
    printf("TMA depth=%d loads/stage=%d (%d KB/stage) mode=%s: %.1f B/clk/SM  (%.0f cycles/stage)\n", DEPTH, NLOADS, NLOADS * 8,
           mode ? "HBM-distinct" : "L2-shared", (double)iters * NLOADS * 8192 / avg, avg / iters);
}


// ---- test 8: TMA multicast within a cluster.  Each CTA of a cluster of size C loads 1/C of every stage and multicasts it to
// all C CTAs; a stage is recycled when ALL CTAs have seen it land (remote mbarrier arrives), as a real GEMM ring would.
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_bar), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
template <int C, int DEPTH>
__global__ void __launch_bounds__(64) probe_mc_kernel(const __grid_constant__ CUtensorMap map, int iters, int rows_total, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[DEPTH], empty[DEPTH];
    const int tid = threadIdx.x;
    const uint32_t rank = cluster_rank();
    if (tid == 0) {
        for (int i = 0; i < DEPTH; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), C); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    long long t0 = clock64();
    if (tid == 0) {
        const uint32_t sb = (smem_u32(smem) + 1023u) & ~1023u;
        constexpr int STAGE = 32768, SLICE = STAGE / C;           // each CTA loads SLICE bytes = SLICE/64 rows
        int row_base = (int)((((long long)(blockIdx.x / C)) * 8191) % (rows_total / 512 - 1)) * 512;
        for (int it = 0; it < iters + DEPTH; ++it) {
            const int s = it % DEPTH;
            if (it >= DEPTH) {           // consume: wait for the stage issued DEPTH iterations ago, then release it everywhere
                mbar_wait(smem_u32(&full[s]), ((it / DEPTH) - 1) & 1);
                for (uint32_t c = 0; c < (uint32_t)C; ++c) mbar_arrive_remote(smem_u32(&empty[s]), c);
            }
            if (it < iters) {
                if (it >= DEPTH) mbar_wait(smem_u32(&empty[s]), ((it / DEPTH) - 1) & 1);
                mbar_expect_tx(smem_u32(&full[s]), STAGE);
                row_base = (row_base + 512 * 37) % (rows_total - 512);
                // rows [row_base + rank*SLICE/64, +SLICE/64) of this stage; box = 128 rows x 64 B = 8 KB
                for (int l = 0; l < SLICE / 8192; ++l) {
                    const int piece = rank * (SLICE / 8192) + l;
                    tma_load_2d_mc(sb + s * STAGE + piece * 8192, &map, smem_u32(&full[s]), 0, row_base + piece * 128, (uint16_t)((1u << C) - 1));
                }
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (tid == 0) out[blockIdx.x] = t1 - t0;
}
template <int C, int DEPTH>
static void run_mc(EncodeFn enc, __half* buf, int rows_total, long long* dout) {
    CUtensorMap m = make_map(enc, buf, rows_total, 32, 128, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    const int smem = DEPTH * 32768 + 1024;
    CK(cudaFuncSetAttribute(probe_mc_kernel<C, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 400;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148 / C * C); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, probe_mc_kernel<C, DEPTH>, m, iters, rows_total, dout));
    CK(cudaDeviceSynchronize());
    const int n = 148 / C * C;
    std::vector<long long> h(n);
    CK(cudaMemcpy(h.data(), dout, n * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (auto v : h) avg += (double)v; avg /= n;
    printf("MULTICAST cluster=%d depth=%d: delivered %.1f B/clk/SM (L2 reads %.1f B/clk/SM), %.0f cycles/stage\n", C, DEPTH,
           (double)iters * 32768 / avg, (double)iters * 32768 / C / avg, avg / iters);
}


// ---- test 9: interleaved operands.  Q,K rows = [hi 32 | lo 32] halves (128 B, K-major SWIZZLE_128B): S = Qh Kh^T + Qh Kl^T + Ql Kh^T
// by pointing the K=16 slices at byte offsets 0/32 (hi) and 64/96 (lo).  V rows = [Vh 32 | Vl 32] (MN-major SWIZZLE_128B):
// O = Ph [Vh|Vl] (N=64) + Pl Vh (N=32 sub-read of the same tile), P from TMEM.  D0 = S [128x128], D1 = O [128x32].
__global__ void __launch_bounds__(128) probe_il_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                                                       const __grid_constant__ CUtensorMap mapV, const float* __restrict__ P,
                                                       float* __restrict__ Sout, float* __restrict__ Oout) {
    extern __shared__ __align__(1024) uint8_t smem_[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t smem = (smem_u32(smem_) + 1023u) & ~1023u;
    const uint32_t bar_tma = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) { mbar_init(bar_tma, 1); mbar_init(bar_mma, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t sQ = smem, sK = smem + 16384, sV = smem + 32768;     // each 128 rows x 128 B = 16 KB
    if (tid == 0) {
        mbar_expect_tx(bar_tma, 3 * 16384);
        tma_load_2d(sQ, &mapQ, bar_tma, 0, 0);
        tma_load_2d(sK, &mapK, bar_tma, 0, 0);
        tma_load_2d(sV, &mapV, bar_tma, 0, 0);
    }
    const float* prow = P + (size_t)tid * 128;
    for (int half = 0; half < 2; ++half) {
        uint32_t hi[32], lo[32];
        for (int i = 0; i < 32; ++i) {
            const float x0 = prow[half * 64 + 2 * i], x1 = prow[half * 64 + 2 * i + 1];
            __half2 h = __floats2half2_rn(x0, x1);
            float2 hf = __half22float2(h);
            __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
            hi[i] = *reinterpret_cast<uint32_t*>(&h);
            lo[i] = *reinterpret_cast<uint32_t*>(&l);
        }
        tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + half * 32, hi);          // P_hi cols [0,64)
        tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 64 + half * 32, lo);     // P_lo cols [64,128)
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    mbar_wait(bar_tma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t hb = ((1024u >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);     // SBO 1024, SWIZZLE_128B
        auto desc = [&](uint32_t a) { return ((uint64_t)hb << 32) | (1ull << 16) | (uint64_t)((a >> 4) & 0x3FFF); };
        const uint32_t id_qk = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t id_pv64 = (1u << 4) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t id_pv32 = (1u << 4) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t dS = tmem + 128, dO = tmem + 256;
        for (int ks = 0; ks < 2; ++ks) {
            umma_f16(dS, desc(sQ + 64 + ks * 32), desc(sK + ks * 32), id_qk, ks != 0);        // Ql Kh
            umma_f16(dS, desc(sQ + ks * 32), desc(sK + 64 + ks * 32), id_qk, 1u);             // Qh Kl
            umma_f16(dS, desc(sQ + ks * 32), desc(sK + ks * 32), id_qk, 1u);                  // Qh Kh
        }
        for (int k = 0; k < 8; ++k) {      // 16 kv rows of 128 B per slice
            umma_f16_ts(dO, tmem + k * 8, desc(sV + k * 2048), id_pv64, k != 0);              // Ph [Vh|Vl]
            umma_f16_ts(dO, tmem + 64 + k * 8, desc(sV + k * 2048), id_pv32, 1u);             // Pl Vh
        }
        umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32p(tmem + ((uint32_t)(warp * 32) << 16) + 128 + c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) Sout[(size_t)tid * 128 + c0 + i] = __uint_as_float(v[i]);
    }
    uint32_t v[32], v2[32];
    tmem_ld32p(tmem + ((uint32_t)(warp * 32) << 16) + 256, v);
    tmem_ld32p(tmem + ((uint32_t)(warp * 32) << 16) + 288, v2);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) Oout[(size_t)tid * 32 + i] = __uint_as_float(v[i]) + __uint_as_float(v2[i]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}


// ---- test 10: cta_group::2 pair GEMM.  Cluster of 2 CTAs computes D[256 x 128] = A[256 x K] B[128 x K]^T: each CTA holds its 128
// rows of A and 64 of the 128 rows of B; the leader issues M=256 MMAs that read both CTAs' smem and write both CTAs' TMEM.
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum));
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
probe_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* __restrict__ D, int kblocks) {
    extern __shared__ __align__(1024) uint8_t smem_[];
    __shared__ __align__(8) uint64_t bars[2];        // [0] full (used in the leader), [1] mma_done (arrives in both CTAs)
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_rank();
    const uint32_t smem = (smem_u32(smem_) + 1023u) & ~1023u;
    const uint32_t sA = smem, sB = smem + 16384;      // A: 128 rows x 128 B; B half: 64 rows x 128 B
    const uint32_t bar_full = smem_u32(&bars[0]), bar_done = smem_u32(&bars[1]);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    if (tid == 0) { mbar_init(bar_full, 1); mbar_init(bar_done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    uint32_t phase = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
        if (tid == 0) {
            if (rank == 0) mbar_expect_tx(bar_full, 2 * (16384 + 8192));          // bytes landing in BOTH CTAs
            tma_load_2d_2sm(sA, &mapA, bar_full, kb * 64, (int)rank * 128);      // my 128 rows of A
            tma_load_2d_2sm(sB, &mapB, bar_full, kb * 64, (int)rank * 64);       // my 64 rows of B
            if (rank == 0) {
                mbar_wait(bar_full, phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t hb = ((1024u >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
                const uint64_t da = ((uint64_t)hb << 32) | (1ull << 16) | (uint64_t)((sA >> 4) & 0x3FFF);
                const uint64_t db = ((uint64_t)hb << 32) | (1ull << 16) | (uint64_t)((sB >> 4) & 0x3FFF);
                const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
                for (int k = 0; k < 4; ++k) umma_f16_2sm(tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                umma_commit_2sm(bar_done, 3);
            }
        }
        mbar_wait(bar_done, phase);       // both CTAs: the MMAs have finished reading this stage's smem
        phase ^= 1;
        __syncthreads();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32p(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) D[(size_t)(rank * 128 + tid) * 128 + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

static uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
    return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (layout_type << 29);   // SBO | version=1 (bit 46) | layout (bits 61..63)
}

int main(int argc, char** argv) {
    const int test = argc > 1 ? atoi(argv[1]) : 0;
    const int variant = argc > 2 ? atoi(argv[2]) : 0;
    EncodeFn enc = get_encode();

    if (test == 4) {
        std::vector<float> Pm(128 * 128), Vm(128 * 32);
        srand(77);
        for (auto& v : Pm) v = (rand() % 10001) / 10000.0f * (rand() % 7 == 0 ? 1.0f : 0.01f);
        for (auto& v : Vm) v = (rand() % 2001 - 1000) / 500.0f;
        std::vector<__half> Vh(Vm.size()), Vl(Vm.size());
        for (size_t i = 0; i < Vm.size(); ++i) { Vh[i] = __float2half_rn(Vm[i]); Vl[i] = __float2half_rn(Vm[i] - __half2float(Vh[i])); }
        float *dP, *dD; __half *dVh, *dVl;
        CK(cudaMalloc(&dP, Pm.size() * 4)); CK(cudaMalloc(&dD, 128 * 32 * 4)); CK(cudaMalloc(&dVh, Vm.size() * 2)); CK(cudaMalloc(&dVl, Vm.size() * 2));
        CK(cudaMemcpy(dP, Pm.data(), Pm.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dVh, Vh.data(), Vm.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dVl, Vl.data(), Vm.size() * 2, cudaMemcpyHostToDevice));
        CUtensorMap mVh = make_map(enc, dVh, 128, 32, 128, 32, CU_TENSOR_MAP_SWIZZLE_64B);
        CUtensorMap mVl = make_map(enc, dVl, 128, 32, 128, 32, CU_TENSOR_MAP_SWIZZLE_64B);
        CK(cudaFuncSetAttribute(probe_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024));
        probe_ts_kernel<<<1, 128, 16384 + 1024>>>(mVh, mVl, dP, dD, variant > 0 ? 1 : 0, variant == 2 ? 512u : (variant == 3 ? 64u : 8192u));
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<float> Dm(128 * 32);
        CK(cudaMemcpy(Dm.data(), dD, Dm.size() * 4, cudaMemcpyDeviceToHost));
        double max_err = 0, max_ref = 0;
        for (int i = 0; i < 128; ++i)
            for (int j = 0; j < 32; ++j) {
                double ref = 0;
                for (int k = 0; k < 128; ++k) ref += (double)Pm[i * 128 + k] * (double)Vm[k * 32 + j];
                max_err = fmax(max_err, fabs(ref - (double)Dm[i * 32 + j]));
                max_ref = fmax(max_ref, fabs(ref));
            }
        const bool pass = max_err <= 2e-5 * max_ref + 1e-6;
        printf("PROBE test=4 variant=%d (TS-mode PV, P via tcgen05.st; variant>0: [Vh|Vl] N=64 concat) max_err=%.3e max_ref=%.3f %s\n", max_err, max_ref, pass ? "PASS" : "FAIL");
        return pass ? 0 : 1;
    }





    if (test == 10) {
        const int M2 = 256, N2 = 128, K2 = 256;
        std::vector<float> A2((size_t)M2 * K2), B2((size_t)N2 * K2);
        srand(4242);
        for (auto& v : A2) v = (rand() % 2001 - 1000) / 500.0f;
        for (auto& v : B2) v = (rand() % 2001 - 1000) / 500.0f;
        std::vector<__half> Ah2(A2.size()), Bh2(B2.size());
        for (size_t i = 0; i < A2.size(); ++i) Ah2[i] = __float2half_rn(A2[i]);
        for (size_t i = 0; i < B2.size(); ++i) Bh2[i] = __float2half_rn(B2[i]);
        __half *dA2, *dB2; float* dD2;
        CK(cudaMalloc(&dA2, A2.size() * 2)); CK(cudaMalloc(&dB2, B2.size() * 2)); CK(cudaMalloc(&dD2, (size_t)M2 * N2 * 4));
        CK(cudaMemcpy(dA2, Ah2.data(), A2.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB2, Bh2.data(), B2.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemset(dD2, 0xFF, (size_t)M2 * N2 * 4));
        CUtensorMap mA2 = make_map(enc, dA2, M2, K2, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B);
        CUtensorMap mB2 = make_map(enc, dB2, N2, K2, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
        CK(cudaFuncSetAttribute(probe_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 24576 + 1024));
        probe_pair_kernel<<<2, 128, 24576 + 1024>>>(mA2, mB2, dD2, K2 / 64);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<float> D2((size_t)M2 * N2);
        CK(cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost));
        double me = 0, mr = 0; int bi = -1, bj = -1;
        for (int i = 0; i < M2; ++i)
            for (int j = 0; j < N2; ++j) {
                double ref = 0;
                for (int k = 0; k < K2; ++k) ref += (double)__half2float(Ah2[(size_t)i * K2 + k]) * (double)__half2float(Bh2[(size_t)j * K2 + k]);
                const double e = fabs(ref - (double)D2[(size_t)i * N2 + j]);
                if (!(e <= me)) { me = e; bi = i; bj = j; }
                mr = fmax(mr, fabs(ref));
            }
        const bool pass = me <= 1e-4 * mr + 1e-4;
        printf("PROBE test=10 (cta_group::2 pair GEMM 256x128x256) max_err=%.3e at (%d,%d) max_ref=%.2f %s\n", me, bi, bj, mr, pass ? "PASS" : "FAIL");
        return pass ? 0 : 1;
    }
    if (test == 9) {
        std::vector<float> Q(128 * 32), K(128 * 32), V(128 * 32), Pm(128 * 128);
        srand(99);
        for (auto& v : Q) v = (rand() % 2001 - 1000) / 400.0f;
        for (auto& v : K) v = (rand() % 2001 - 1000) / 400.0f;
        for (auto& v : V) v = (rand() % 2001 - 1000) / 500.0f;
        for (auto& v : Pm) v = (rand() % 10001) / 10000.0f * (rand() % 7 == 0 ? 1.0f : 0.01f);
        auto interleave = [](const std::vector<float>& x) {        // [128][hi 32 | lo 32]
            std::vector<__half> o(128 * 64);
            for (int r = 0; r < 128; ++r)
                for (int c = 0; c < 32; ++c) {
                    const __half h = __float2half_rn(x[r * 32 + c]);
                    o[r * 64 + c] = h;
                    o[r * 64 + 32 + c] = __float2half_rn(x[r * 32 + c] - __half2float(h));
                }
            return o;
        };
        auto Qi = interleave(Q), Ki = interleave(K), Vi = interleave(V);
        __half *dQ, *dK, *dV; float *dP, *dS, *dO;
        CK(cudaMalloc(&dQ, 16384)); CK(cudaMalloc(&dK, 16384)); CK(cudaMalloc(&dV, 16384));
        CK(cudaMalloc(&dP, Pm.size() * 4)); CK(cudaMalloc(&dS, 128 * 128 * 4)); CK(cudaMalloc(&dO, 128 * 32 * 4));
        CK(cudaMemcpy(dQ, Qi.data(), 16384, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dK, Ki.data(), 16384, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dV, Vi.data(), 16384, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dP, Pm.data(), Pm.size() * 4, cudaMemcpyHostToDevice));
        CUtensorMap mQ = make_map(enc, dQ, 128, 64, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B);
        CUtensorMap mK = make_map(enc, dK, 128, 64, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B);
        CUtensorMap mV = make_map(enc, dV, 128, 64, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B);
        CK(cudaFuncSetAttribute(probe_il_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024));
        probe_il_kernel<<<1, 128, 49152 + 1024>>>(mQ, mK, mV, dP, dS, dO);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<float> Sm(128 * 128), Om(128 * 32);
        CK(cudaMemcpy(Sm.data(), dS, Sm.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(Om.data(), dO, Om.size() * 4, cudaMemcpyDeviceToHost));
        double es = 0, rs = 0, eo = 0, ro = 0;
        for (int i = 0; i < 128; ++i) {
            for (int j = 0; j < 128; ++j) {
                double ref = 0;
                for (int k = 0; k < 32; ++k) ref += (double)Q[i * 32 + k] * (double)K[j * 32 + k];
                es = fmax(es, fabs(ref - Sm[i * 128 + j])); rs = fmax(rs, fabs(ref));
            }
            for (int j = 0; j < 32; ++j) {
                double ref = 0;
                for (int k = 0; k < 128; ++k) ref += (double)Pm[i * 128 + k] * (double)V[k * 32 + j];
                eo = fmax(eo, fabs(ref - Om[i * 32 + j])); ro = fmax(ro, fabs(ref));
            }
        }
        const bool ps = es <= 2e-5 * rs + 1e-6, po = eo <= 2e-5 * ro + 1e-6;
        printf("PROBE test=9 interleaved QK: max_err=%.3e (ref %.2f) %s | interleaved PV: max_err=%.3e (ref %.2f) %s\n", es, rs,
               ps ? "PASS" : "FAIL", eo, ro, po ? "PASS" : "FAIL");
        return (ps && po) ? 0 : 1;
    }
    if (test == 8) {
        long long* dout; CK(cudaMalloc(&dout, 2048 * 8));
        const int rows_total = 1 << 23;
        __half* buf; CK(cudaMalloc(&buf, (size_t)rows_total * 64));
        CK(cudaMemset(buf, 0, (size_t)rows_total * 64));
        run_mc<1, 6>(enc, buf, rows_total, dout);
        run_mc<2, 6>(enc, buf, rows_total, dout);
        run_mc<4, 6>(enc, buf, rows_total, dout);
        return 0;
    }
    if (test == 7) {
        long long* dout; CK(cudaMalloc(&dout, 2048 * 8));
        const int rows_total = 1 << 23;            // 8M rows x 64 B = 512 MB
        __half* buf; CK(cudaMalloc(&buf, (size_t)rows_total * 64));
        CK(cudaMemset(buf, 0, (size_t)rows_total * 64));
        for (int mode = 0; mode < 2; ++mode) {
            run_tma<3, 4>(enc, buf, rows_total, mode, dout);
            run_tma<6, 4>(enc, buf, rows_total, mode, dout);
            run_tma<12, 2>(enc, buf, rows_total, mode, dout);
            run_tma<24, 1>(enc, buf, rows_total, mode, dout);
            run_tma<6, 4>(enc, buf, rows_total, mode, dout, true);
            run_tma<12, 4>(enc, buf, rows_total, mode, dout, true);
            run_tma<6, 4>(enc, buf, 1 << 18, mode, dout);        // 16 MB working set: L2-resident, distinct per CTA
        }
        return 0;
    }
    if (test == 6) {
        long long* dout; CK(cudaMalloc(&dout, 2048 * 8));
        run_rate2<128, false, false, 1>("SS N128 1 acc", dout);
        run_rate2<128, false, false, 2>("SS N128 2 acc", dout);
        run_rate2<64, false, false, 1>("SS N64 1 acc", dout);
        run_rate2<64, false, false, 4>("SS N64 4 acc", dout);
        run_rate2<32, false, true, 1>("SS N32 (B MN) 1 acc", dout);
        run_rate2<32, false, true, 4>("SS N32 (B MN) 4 acc", dout);
        run_rate2<32, true, true, 1>("TS N32 (B MN) 1 acc", dout);
        run_rate2<32, true, true, 4>("TS N32 (B MN) 4 acc", dout);
        run_rate2<256, false, false, 1>("SS N256 1 acc", dout);
        run_rate2<128, true, false, 1>("TS N128 1 acc", dout);
        return 0;
    }
    if (test == 5) {
        long long* dout; CK(cudaMalloc(&dout, 2048 * 8));
        CK(cudaFuncSetAttribute(probe_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024));
        const int iters = 2000;
        const char* names[] = {"tcgen05.ld 32x32b.x32 (4 warps)", "tcgen05.st 32x32b.x32 (4 warps)", "SS MMA M128 N128 K16",
                               "SS MMA M128 N32 K16 (B MN)", "TS MMA M128 N32 K16 (B MN)", "SS MMA M128 N64 K16",
                               "SS N128, 2 accumulators", "TS N32, 4 accumulators", "SS N256", "SS N32 MN, 4 accumulators",
                               "SS N128 accumulate=0", "TS N32, 8 accumulators", "SS N128 + TS N32 alternating (per pair)"};
        for (int mode = 0; mode < 13; ++mode) {
            probe_rate_kernel<<<148, 128, 32768 + 1024>>>(mode, iters, dout);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            std::vector<long long> h(148);
            CK(cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost));
            double avg = 0; for (auto v : h) avg += (double)v; avg /= 148.0;
            const double per = avg / iters;
            double bytes = mode < 2 ? 4.0 * 4096.0 : 0.0;      // 4 warps x 4 KB per iteration
            printf("RATE mode=%d %-34s cycles/iter=%.1f", mode, names[mode], per);
            if (mode < 2) printf("  => %.1f B/clk/SM", bytes / per);
            printf("\n");
        }
        return 0;
    }
    const int M = 128;
    int N = 128, K = 256, kb_elems = 64;
    if (test == 2) { N = 32; K = 128; }
    if (test == 3) { N = 128; K = 32; kb_elems = 32; }
    const int passes = test == 1 ? 3 : 1;
    const bool b_mn = test == 2;
    // data
    std::vector<float> A((size_t)M * K), B((size_t)N * K);
    srand(1234 + test);
    for (auto& v : A) v = (rand() % 2001 - 1000) / 500.0f;
    for (auto& v : B) v = (rand() % 2001 - 1000) / 500.0f;
    std::vector<__half> Ah(A.size()), Al(A.size()), Bh(B.size()), Bl(B.size());
    auto split = [](const std::vector<float>& x, std::vector<__half>& h, std::vector<__half>& l) {
        for (size_t i = 0; i < x.size(); ++i) { h[i] = __float2half_rn(x[i]); l[i] = __float2half_rn(x[i] - __half2float(h[i])); }
    };
    split(A, Ah, Al);
    split(B, Bh, Bl);
    // B storage: K-major -> [N][K]; MN-major -> [K][N]
    std::vector<__half> Bh_s(Bh), Bl_s(Bl);
    if (b_mn)
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) { Bh_s[(size_t)k * N + n] = Bh[(size_t)n * K + k]; Bl_s[(size_t)k * N + n] = Bl[(size_t)n * K + k]; }
    __half *dAh, *dAl, *dBh, *dBl;
    float* dD;
    CK(cudaMalloc(&dAh, A.size() * 2)); CK(cudaMalloc(&dAl, A.size() * 2));
    CK(cudaMalloc(&dBh, B.size() * 2)); CK(cudaMalloc(&dBl, B.size() * 2));
    CK(cudaMalloc(&dD, (size_t)M * N * 4));
    CK(cudaMemcpy(dAh, Ah.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dAl, Al.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBh, Bh_s.data(), B.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBl, Bl_s.data(), B.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, (size_t)M * N * 4));

    ProbeParams p{};
    p.n = N; p.kblocks = K / kb_elems; p.kb_elems = kb_elems; p.passes = passes; p.b_is_mn = b_mn;
    p.idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24) | (b_mn ? (1u << 16) : 0u);
    CUtensorMap mAh, mAl, mBh, mBl;
    if (test == 3) {          // 64-byte rows, SWIZZLE_64B, K-major both
        mAh = make_map(enc, dAh, M, K, 128, 32, CU_TENSOR_MAP_SWIZZLE_64B); mAl = make_map(enc, dAl, M, K, 128, 32, CU_TENSOR_MAP_SWIZZLE_64B);
        mBh = make_map(enc, dBh, N, K, 128, 32, CU_TENSOR_MAP_SWIZZLE_64B); mBl = make_map(enc, dBl, N, K, 128, 32, CU_TENSOR_MAP_SWIZZLE_64B);
        p.a_hi = p.b_hi = desc_hi(512, 4);     // SBO = 8 rows * 64 B, layout SWIZZLE_64B
        p.a_lbo = p.b_lbo = variant == 1 ? 0 : 1;
        p.a_kstep = p.b_kstep = 32;
        p.a_bytes = 128 * 64; p.b_bytes = 128 * 64;
    } else {
        mAh = make_map(enc, dAh, M, K, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B); mAl = make_map(enc, dAl, M, K, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B);
        p.a_hi = desc_hi(1024, 2); p.a_lbo = 1; p.a_kstep = 32; p.a_bytes = 128 * 128;
        if (!b_mn) {
            mBh = make_map(enc, dBh, N, K, N, 64, CU_TENSOR_MAP_SWIZZLE_128B); mBl = make_map(enc, dBl, N, K, N, 64, CU_TENSOR_MAP_SWIZZLE_128B);
            p.b_hi = desc_hi(1024, 2); p.b_lbo = variant == 1 ? 0 : 1; p.b_kstep = 32; p.b_bytes = N * 128;
            if (variant == 1) p.a_lbo = 0;
        } else {              // V tile: [64 k rows][32 n] = 64-byte rows, SWIZZLE_64B, MN-major
            mBh = make_map(enc, dBh, K, N, 64, 32, CU_TENSOR_MAP_SWIZZLE_64B); mBl = make_map(enc, dBl, K, N, 64, 32, CU_TENSOR_MAP_SWIZZLE_64B);
            p.b_bytes = 64 * 64;
            p.b_kstep = 16 * 64;             // 16 k rows per UMMA_K step
            // canonical MN-major B64: ((4,n),(8,k)):((1,LBO),(4,SBO)): SBO = stride between 8-row k groups = 512 B
            const uint32_t sbo[] = {512, 512, 512, 1024, 256, 512};
            const uint32_t lbo[] = {0, 512, 256, 512, 512, 1024};
            p.b_hi = desc_hi(sbo[variant % 6], 4);
            p.b_lbo = lbo[variant % 6] >> 4;
        }
    }
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024));
    probe_kernel<<<1, 128, 32768 + 1024>>>(mAh, mAl, mBh, mBl, dD, p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D((size_t)M * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_ref = 0;
    int bad_i = -1, bad_j = -1;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            double ref = 0;
            for (int k = 0; k < K; ++k) {
                const double a = passes == 3 ? (double)A[(size_t)i * K + k] : (double)__half2float(Ah[(size_t)i * K + k]);
                const double b = passes == 3 ? (double)B[(size_t)j * K + k] : (double)__half2float(Bh[(size_t)j * K + k]);
                ref += a * b;
            }
            const double e = fabs(ref - (double)D[(size_t)i * N + j]);
            if (!(e <= max_err)) { max_err = e; bad_i = i; bad_j = j; }
            max_ref = fmax(max_ref, fabs(ref));
        }
    const double tol = (passes == 3 ? 3e-6 : 2e-6) * max_ref * 4 + 1e-4;
    const bool pass = max_err <= tol;
    printf("PROBE test=%d variant=%d N=%d K=%d max_err=%.3e (at %d,%d; got %.5f) max_ref=%.3f tol=%.2e %s\n", test, variant, N, K,
           max_err, bad_i, bad_j, D[(size_t)(bad_i < 0 ? 0 : bad_i) * N + (bad_j < 0 ? 0 : bad_j)], max_ref, tol, pass ? "PASS" : "FAIL");
    return pass ? 0 : 1;
}
