// Split-fp16 GEMM on the 5th-gen tensor cores: TMA -> smem ring -> tcgen05.mma (TMEM accumulator) -> fused epilogue.
//
// Replaces every `Linear.forward -> F.linear` call site on the hot path (reference
// PhysDock/models/primitives/linear.py:146-161): q/k/v projections (attentions.py:248-250), out projection
// (:263), SwiGLU w1/w3/w2 (feed_forward.py:30-31), linear_downscale/upscale (layers/transformers.py:206,215).
//
// C[M,N] = A[M,K] W[N,K]^T with A, W stored as (hi, lo) fp16 planes; per K=16 slice three MMAs
// (Al*Wh + Ah*Wl + Ah*Wh) accumulate in fp32 in TMEM.
//
// PERSISTENT kernel: one CTA per SM walks 128x128 output tiles (n fastest, so consecutive CTAs share the A tile
// in L2); 320 threads, warp-specialised:
//   warp 0   : TMA producer: 4 plane tiles [128 rows x 64 halves] per stage, SWIZZLE_128B, 3-stage ring (192 KB)
//              that runs ahead across tile boundaries
//   warp 1   : TMEM allocator + MMA issuer: 12 x tcgen05.mma.kind::f16 M128 N128 K16 per stage into one of TWO
//              128-column accumulators, so the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2-9: epilogue (two warps per TMEM lane quarter, two 32-column chunks each).  A THREAD owns one output
//              row and reads 32 consecutive accumulator columns per tcgen05.ld -- exactly one attention head / one
//              SwiGLU column block, so the per-head RMSNorm and the SwiGLU product need no cross-thread traffic.
// (The first version launched one CTA per tile: tensor pipe 40% busy, the rest was per-tile prologue/epilogue and
//  wave tails -- profiles/r01_gemm_v2a_ncu.txt.)
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace pdk {

namespace {

// BK = 64 halves = 128-byte rows: TMA boxes narrower than 128 B run at less than half rate (28 vs 65 B/clk/SM,
// tests/cuda/umma_probe.cu test 7), which made the BK = 32 version L2-fabric bound.
constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3, ACC = 2;
constexpr int TILE_BYTES = 128 * BK * 2;            // 16 KB: one fp16 plane tile, 128-byte rows (SWIZZLE_128B)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;         // A_hi, A_lo, W_hi, W_lo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;   // + slack to 1024-align the ring
constexpr int NTHREADS = 320;
constexpr int EPI_THREADS = 256;
constexpr uint32_t TMEM_COLS = ACC * BN;

// Debug-only compile switches used by tools/gemm_variants.sh to attribute time (never defined in the product build):
//   PDK_DBG_NO_STORE  epilogue skips its global stores      PDK_DBG_NO_TMAWAIT  MMA warp does not wait for TMA data
//   PDK_DBG_NO_EPI    epilogue only drains TMEM (no math, no stores)
PDK_DEV void store8(float* dst, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// x * sigmoid(x) with MUFU ex2 / rcp (relative error ~3e-7; the IEEE expf + divide version cost ~50 instructions
// per value and made the SwiGLU epilogue the bottleneck of the K=128 GEMMs)
PDK_DEV float silu_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + ex2(-kLog2e * x)));
    return x * r;
}

template <int EPI>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                 const __grid_constant__ CUtensorMap mWh, const __grid_constant__ CUtensorMap mWl, const GemmArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * STAGES + 2 * ACC];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KT = p.K / BK;
    const int num_n = p.N / BN;
    const int num_tiles = (p.M / BM) * num_n;
    const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = smem_u32(&bars[0]);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull = [&](int b) { return bar0 + 8u * (2 * STAGES + b); };
    auto tempty = [&](int b) { return bar0 + 8u * (2 * STAGES + ACC + b); };

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
#pragma unroll
        for (int b = 0; b < ACC; ++b) { mbar_init(tfull(b), 1); mbar_init(tempty(b), EPI_THREADS); }
        mbar_fence_init();
        tma_prefetch_desc(&mAh); tma_prefetch_desc(&mAl); tma_prefetch_desc(&mWh); tma_prefetch_desc(&mWl);
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        // ================================================================= TMA producer
        int s = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int m0 = (t / num_n) * BM, n0 = (t % num_n) * BN;
            for (int kt = 0; kt < KT; ++kt) {
                mbar_wait(empty(s), ph ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(full(s), STAGE_BYTES);
                    const uint32_t dst = ring + s * STAGE_BYTES;
                    tma_load_2d(dst, &mAh, full(s), kt * BK, m0);
                    tma_load_2d(dst + TILE_BYTES, &mAl, full(s), kt * BK, m0);
                    tma_load_2d(dst + 2 * TILE_BYTES, &mWh, full(s), kt * BK, n0);
                    tma_load_2d(dst + 3 * TILE_BYTES, &mWl, full(s), kt * BK, n0);
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
        int s = 0, lt = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
            const int buf = lt & 1;
            mbar_wait(tempty(buf), (((uint32_t)lt >> 1) & 1u) ^ 1u);      // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d = tmem + buf * BN;
            for (int kt = 0; kt < KT; ++kt) {
#ifndef PDK_DBG_NO_TMAWAIT
                mbar_wait(full(s), ph);
#endif
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t src = ring + s * STAGE_BYTES;
                    const uint64_t ah = smem_desc(src, 1024, kLayoutSw128), al = smem_desc(src + TILE_BYTES, 1024, kLayoutSw128);
                    const uint64_t wh = smem_desc(src + 2 * TILE_BYTES, 1024, kLayoutSw128);
                    const uint64_t wl = smem_desc(src + 3 * TILE_BYTES, 1024, kLayoutSw128);
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        const uint64_t o = (uint64_t)(ks * 2);                  // +32 bytes (>>4) per K=16 slice
                        umma_f16(d, al + o, wh + o, idesc, (kt | ks) != 0);      // small terms first
                        umma_f16(d, ah + o, wl + o, idesc, 1u);
                        umma_f16(d, ah + o, wh + o, idesc, 1u);
                    }
                    umma_commit(empty(s));                   // frees the smem stage once these MMAs have read it
                    if (kt == KT - 1) umma_commit(tfull(buf));   // accumulator complete
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ================================================================= epilogue: 8 warps, thread = (row, 2 chunks)
        const int ew = warp - 2;
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int chunk0 = (ew >> 2) * 2;             // warps 2-5: chunks 0,1; warps 6-9: chunks 2,3
        int lt = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
            const int buf = lt & 1;
            const int m0 = (t / num_n) * BM, n0 = (t % num_n) * BN;
            mbar_wait(tfull(buf), ((uint32_t)lt >> 1) & 1u);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            const uint32_t taddr = tmem + buf * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int ch = chunk0; ch < chunk0 + 2; ++ch) {
                uint32_t raw[32];
                tmem_ld32(taddr + ch * 32, raw);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                const int col = n0 + ch * 32;
#ifdef PDK_DBG_NO_EPI
                if (v[0] == 123.456f && col == -1) p.out[row] = v[1];
                continue;
#endif
                if constexpr (EPI == EPI_STORE) {
                    if (p.bias) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + col + i);
                    }
                    if (p.act_silu) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = silu_fast(v[i]);
                    }
                    store8(p.out + (size_t)row * p.ldo + col, v);
                } else if constexpr (EPI == EPI_GATE_RESID) {
                    const float* gate = p.gate + (size_t)(row / p.rows_per_sample) * p.gate_stride + col;
                    float* x = p.out + (size_t)row * p.ldo + col;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4 xv = reinterpret_cast<float4*>(x)[i];
                        const float4 gv = __ldg(reinterpret_cast<const float4*>(gate) + i);
                        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + col) + i);
                        xv.x += (v[4 * i] + bv.x) * gv.x;
                        xv.y += (v[4 * i + 1] + bv.y) * gv.y;
                        xv.z += (v[4 * i + 2] + bv.z) * gv.z;
                        xv.w += (v[4 * i + 3] + bv.w) * gv.w;
                        reinterpret_cast<float4*>(x)[i] = xv;
                    }
                } else if constexpr (EPI == EPI_SWIGLU) {
                    // W rows interleaved in blocks of 16: columns [0,16) = w1 rows, [16,32) = w3 rows of hidden j0..j0+15
                    const int j0 = col / 2;
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        split2(silu_fast(v[2 * i]) * v[16 + 2 * i], silu_fast(v[2 * i + 1]) * v[16 + 2 * i + 1], hi[i], lo[i]);
                    uint4* dh = reinterpret_cast<uint4*>(p.ph + (size_t)row * p.ldp + j0);
                    uint4* dl = reinterpret_cast<uint4*>(p.pl + (size_t)row * p.ldp + j0);
#ifdef PDK_DBG_NO_STORE
                    if (hi[0] == 0x12345678u && lo[7] == 0x9abcdef0u) {
#endif
                    dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]); dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
#ifdef PDK_DBG_NO_STORE
                    }
#endif
                } else {   // EPI_QKV: this chunk is one head of q, k or v
                    const int which = col / p.c;
                    const int head = (col % p.c) / kHeadDim;
                    const int H = p.c / kHeadDim;
                    if (which < 2) {    // per-head RMSNorm (rms_norm.py:14-19); q additionally carries log2e/sqrt(32)
                        float ss = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) ss = fmaf(v[i], v[i], ss);
                        const float inv = (1.0f / sqrtf(ss * (1.0f / kHeadDim) + p.rms_eps)) * (which == 0 ? p.q_scale : 1.0f);
                        const float* gain = which == 0 ? p.norm_q : p.norm_k;
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = v[i] * inv * __ldg(gain + i);
                    }
                    __half* dh = which == 0 ? p.q : (which == 1 ? p.k : p.v);      // row = [hi 32 | lo 32] halves
                    __half* dl = dh + kHeadDim;
                    const size_t d0 = ((size_t)((row / p.rows_per_sample) * H + head) * p.rows_per_sample +
                                       (row % p.rows_per_sample)) * (2 * kHeadDim);
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        reinterpret_cast<uint4*>(dh + d0)[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                        reinterpret_cast<uint4*>(dl + d0)[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty(buf));                 // this thread is done reading accumulator `buf`
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
}

template <int EPI>
cudaError_t launch_one(const GemmArgs& a, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_umma_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    CUtensorMap mAh, mAl, mWh, mWl;
    cudaError_t e;
    if ((e = get_tensor_map_f16(a.Ah, a.M, a.K, a.lda, BM, BK, 128, &mAh)) != cudaSuccess) return e;
    if ((e = get_tensor_map_f16(a.Al, a.M, a.K, a.lda, BM, BK, 128, &mAl)) != cudaSuccess) return e;
    if ((e = get_tensor_map_f16(a.Wh, a.N, a.K, a.ldw, BN, BK, 128, &mWh)) != cudaSuccess) return e;
    if ((e = get_tensor_map_f16(a.Wl, a.N, a.K, a.ldw, BN, BK, 128, &mWl)) != cudaSuccess) return e;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    }
    const int tiles = (a.M / BM) * (a.N / BN);
    gemm_umma_kernel<EPI><<<tiles < num_sms ? tiles : num_sms, NTHREADS, SMEM_BYTES, st>>>(mAh, mAl, mWh, mWl, a);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_gemm(GemmEpilogue epi, const GemmArgs& a, cudaStream_t st) {
    if (a.M <= 0 || a.N <= 0 || a.K <= 0 || a.M % BM || a.N % BN || a.K % BK) return cudaErrorInvalidValue;
    if (a.lda % 8 || a.ldw % 8) return cudaErrorInvalidValue;   // 16-byte global strides for TMA
    switch (epi) {
        case EPI_STORE: return launch_one<EPI_STORE>(a, st);
        case EPI_GATE_RESID: return launch_one<EPI_GATE_RESID>(a, st);
        case EPI_SWIGLU: return launch_one<EPI_SWIGLU>(a, st);
        case EPI_QKV:
            if (a.N != 3 * a.c || a.c % BN) return cudaErrorInvalidValue;
            return launch_one<EPI_QKV>(a, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace pdk
