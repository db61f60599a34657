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

static uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
    return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (layout_type << 29);   // SBO | version=1 (bit 46) | layout (bits 61..63)
}

int main(int argc, char** argv) {
    const int test = argc > 1 ? atoi(argv[1]) : 0;
    const int variant = argc > 2 ? atoi(argv[2]) : 0;
    EncodeFn enc = get_encode();
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
