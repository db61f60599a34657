// Split-fp16 GEMM on the 5th-gen tensor cores: TMA -> smem ring -> tcgen05.mma (TMEM accumulator) -> fused epilogue.
//
// Replaces every `Linear.forward -> F.linear` call site on the hot path (reference
// PhysDock/models/primitives/linear.py:146-161): q/k/v projections (attentions.py:248-250), out projection
// (:263), SwiGLU w1/w3/w2 (feed_forward.py:30-31), linear_downscale/upscale (layers/transformers.py:206,215).
//
// C[M,N] = A[M,K] W[N,K]^T with A, W stored as (hi, lo) fp16 planes; per K=16 slice three MMAs
// (Al*Wh + Ah*Wl + Ah*Wh) accumulate in fp32 in TMEM.
//
// PERSISTENT kernel, 576 threads per CTA, warp-specialised.  Two tilings:
//   PAIR (M % 256 == 0 and K*N >= 250k, i.e. the token SwiGLU / w2 / QKV / out-proj shapes): a CLUSTER OF TWO CTAs owns a 256 x 128
//        tile; each CTA loads its 128 rows of A and 64 of the 128 W rows, the leader issues tcgen05.mma.cta_group::2
//        (M = 256) into both CTAs' TMEM.  Per CTA a stage is 48 KB for 768 MMA cycles instead of 64 KB (the single-CTA
//        tiling is bound by TMA ingest, ~48 B/clk/SM against the 83 B/clk it needs).  Measured (tools/time_gemm.py):
//        token SwiGLU 28.9 -> 26.4 us, token w2 21.0 -> 19.7 us, token out-proj (K = N = 512: 384 instead of 512 KB of
//        operands per CTA) 10.5 -> 10.2 us at B = 16 and 9.7 -> 8.8 us at 4 samples of 64 tokens; the K = 128 atom shapes
//        gain or lose a few tenths of a microsecond (atom out-proj always loses 0.6: cluster launch / two-CTA handshakes on
//        tiles with 2 K-steps): they use pair tiles only when N >= 384 and M >= 16384 (profiles/r02_gemm_tiling_policy.txt).
//   single CTA: 128 x 128 tiles.
// Tiles are walked n-fastest so concurrently running CTAs share the A tile in L2.  Roles:
//   warp 0   : TMA producer: 4 plane tiles [128 rows x 64 halves] per stage, SWIZZLE_128B, ring of 3 stages (single CTA,
//              64 KB each) / 4 stages (pair, 48 KB each), one fewer with WIDE epilogue staging; runs ahead across tiles
//   warp 1   : TMEM allocator + MMA issuer: 12 x tcgen05.mma.kind::f16 M128 N128 K16 per stage into one of TWO
//              128-column accumulators, so the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2-17: epilogue (four warps per TMEM lane quarter, one 32-column chunk each).  A THREAD owns one output
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
constexpr int BM = 128, BN = 128, BK = 64, ACC = 2;
constexpr int TILE_BYTES = 128 * BK * 2;            // 16 KB: one fp16 plane tile of 128 rows, 128-byte rows (SWIZZLE_128B)
constexpr int RING_BYTES = 192 * 1024;            // NARROW staging: 3 stages (single CTA) / 4 stages (pair)
// Epilogue staging: a thread owns one output ROW, so storing straight from registers makes every store instruction
// scatter 16 bytes to 32 different rows (measured: 14 of 36 us on the atom SwiGLU GEMM).  Each epilogue warp therefore
// parks 8 words per row in a private padded smem tile (48-byte rows: conflict-free 128-bit writes and reads) and
// writes it out row-contiguously, 16 rows x 32 bytes per instruction.
constexpr int EPI_WARPS = 16;
constexpr int STG_ROW_BYTES = 48, STG_WARP_BYTES = 32 * STG_ROW_BYTES;
// WIDE staging (template flag): the warp parks all 32 words of its rows (144-byte rows, conflict-free) and writes them
// out as 4 rows x 128 contiguous bytes per instruction.  The 8-word / 32-byte version costs 16 L1 wavefronts per store
// instruction (16 different lines), and the store-heavy epilogues (QKV: 64 KB per tile) were LSU-bound
// (profiles/r01_gemm_qkv_ncu.txt: l1tex 50%, 9 k cycles per tile against 1.5-6 k of MMA).  It needs 72 KB of staging, so
// the ring drops one stage: used where the ring is deep enough anyway (pair tiling) or K = 128 (two stages = a whole tile).
constexpr int WSTG_ROW_BYTES = 144, WSTG_WARP_BYTES = 32 * WSTG_ROW_BYTES;
template <bool PAIR, bool WIDE> __host__ __device__ constexpr int ring_stages() { return PAIR ? (WIDE ? 3 : 4) : (WIDE ? 2 : 3); }
template <bool PAIR, bool WIDE> __host__ __device__ constexpr int smem_bytes() {
    return ring_stages<PAIR, WIDE>() * (PAIR ? 48 : 64) * 1024 + EPI_WARPS * (WIDE ? WSTG_WARP_BYTES : STG_WARP_BYTES) + 1024;
}
constexpr int NTHREADS = (2 + EPI_WARPS) * 32;
constexpr uint32_t TMEM_COLS = ACC * BN;

// Debug-only compile switches used by tools/gemm_variants.sh to attribute time (never defined in the product build):
//   PDK_DBG_NO_STORE  epilogue skips its global stores      PDK_DBG_NO_TMAWAIT  MMA warp does not wait for TMA data
//   PDK_DBG_NO_EPI    epilogue only drains TMEM (no math, no stores)
// x * sigmoid(x) with MUFU ex2 / rcp (relative error ~3e-7; the IEEE expf + divide version cost ~50 instructions
// per value and made the SwiGLU epilogue the bottleneck of the K=128 GEMMs)
PDK_DEV float silu_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + ex2(-kLog2e * x)));
    return x * r;
}

template <int EPI, bool PAIR, bool WIDE>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                 const __grid_constant__ CUtensorMap mWh, const __grid_constant__ CUtensorMap mWl, const GemmArgs p) {
    constexpr int W_ROWS = PAIR ? BN / 2 : BN;                  // W rows this CTA loads per stage
    constexpr int W_TILE = W_ROWS * BK * 2;
    constexpr int STAGE_BYTES = 2 * TILE_BYTES + 2 * W_TILE;    // A_hi, A_lo, W_hi, W_lo: 48 KB (pair) / 64 KB
    constexpr int STAGES = ring_stages<PAIR, WIDE>();
    static_assert(STAGES * STAGE_BYTES <= RING_BYTES, "ring");
    constexpr int TILE_M = PAIR ? 2 * BM : BM;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * STAGES + 2 * ACC];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;        // 0 = leader of the pair
    const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int nworkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int KT = p.K / BK;
    const int num_n = p.N / BN;
    const int num_tiles = (p.M / TILE_M) * num_n;
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
        for (int b = 0; b < ACC; ++b) { mbar_init(tfull(b), 1); mbar_init(tempty(b), PAIR ? 2 * EPI_WARPS : EPI_WARPS); }   // one arrival per epilogue warp
        mbar_fence_init();
        tma_prefetch_desc(&mAh); tma_prefetch_desc(&mAl); tma_prefetch_desc(&mWh); tma_prefetch_desc(&mWl);
    }
    griddep_launch();                 // PDL: let the next kernel's launch + prologue overlap this kernel
    if (warp == 1) {
        if constexpr (PAIR) tmem_alloc_2sm(smem_u32(&tmem_slot), TMEM_COLS);
        else tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync();       // the peer's barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    griddep_wait();                   // everything above overlapped the previous kernel's tail; its outputs are visible now

    if (warp == 0) {
        // ================================================================= TMA producer
        int s = 0;
        uint32_t ph = 0;
        for (int t = worker; t < num_tiles; t += nworkers) {
            const int m0 = (t / num_n) * TILE_M + (int)rank * BM, w0 = (t % num_n) * BN + (int)rank * W_ROWS;
            for (int kt = 0; kt < KT; ++kt) {
                mbar_wait(empty(s), ph ^ 1u);
                if (elect_one()) {
                    const uint32_t dst = ring + s * STAGE_BYTES;
                    if constexpr (PAIR) {      // both CTAs' bytes land on the LEADER's barrier
                        if (rank == 0) mbar_expect_tx(full(s), 2 * STAGE_BYTES);
                        tma_load_2d_2sm(dst, &mAh, full(s), kt * BK, m0);
                        tma_load_2d_2sm(dst + TILE_BYTES, &mAl, full(s), kt * BK, m0);
                        tma_load_2d_2sm(dst + 2 * TILE_BYTES, &mWh, full(s), kt * BK, w0);
                        tma_load_2d_2sm(dst + 2 * TILE_BYTES + W_TILE, &mWl, full(s), kt * BK, w0);
                    } else {
                        mbar_expect_tx(full(s), STAGE_BYTES);
                        tma_load_2d(dst, &mAh, full(s), kt * BK, m0);
                        tma_load_2d(dst + TILE_BYTES, &mAl, full(s), kt * BK, m0);
                        tma_load_2d(dst + 2 * TILE_BYTES, &mWh, full(s), kt * BK, w0);
                        tma_load_2d(dst + 2 * TILE_BYTES + W_TILE, &mWl, full(s), kt * BK, w0);
                    }
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(TILE_M, BN);
        int s = 0, lt = 0;
        uint32_t ph = 0;
        for (int t = worker; t < num_tiles && rank == 0; t += nworkers, ++lt) {      // only the leader issues
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
                    const uint64_t wl = smem_desc(src + 2 * TILE_BYTES + W_TILE, 1024, kLayoutSw128);
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        const uint64_t o = (uint64_t)(ks * 2);                  // +32 bytes (>>4) per K=16 slice
                        if constexpr (PAIR) {
                            umma_f16_2sm(d, al + o, wh + o, idesc, (kt | ks) != 0);
                            umma_f16_2sm(d, ah + o, wl + o, idesc, 1u);
                            umma_f16_2sm(d, ah + o, wh + o, idesc, 1u);
                        } else {
                            umma_f16(d, al + o, wh + o, idesc, (kt | ks) != 0);      // small terms first
                            umma_f16(d, ah + o, wl + o, idesc, 1u);
                            umma_f16(d, ah + o, wh + o, idesc, 1u);
                        }
                    }
                    if constexpr (PAIR) {                    // arrive in BOTH CTAs: stage free / accumulator complete
                        umma_commit_2sm(empty(s));
                        if (kt == KT - 1) umma_commit_2sm(tfull(buf));
                    } else {
                        umma_commit(empty(s));
                        if (kt == KT - 1) umma_commit(tfull(buf));
                    }
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ================================================================= epilogue: 16 warps, thread = (row, one 32-column chunk)
        // Four warps per TMEM lane quarter, one chunk each.  With 8 warps x 2 chunks the epilogue was a serial latency
        // chain per warp (tcgen05.ld -> math -> staging -> residual load -> store, twice) and cost 9-12 us on every atom
        // GEMM whose main loop takes 5-10 us (tools/time_gemm.py with -DPDK_DBG_NO_EPI); the residual tile is now also
        // fetched BEFORE the accumulator is complete.
        const int ew = warp - 2;
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int ch = ew >> 2;                       // warps 2-5: chunk 0, 6-9: 1, 10-13: 2, 14-17: 3 (each group covers all quarters)
        const uint32_t stg = ring + STAGES * STAGE_BYTES + ew * (WIDE ? WSTG_WARP_BYTES : STG_WARP_BYTES);
        const int rr = lane >> 1, rc = lane & 1;      // write-out role: row (within a group of 16) and 16-byte piece
        const int wr = lane >> 3, wc = lane & 7;      // WIDE write-out role: row (within a group of 4) and 16-byte piece of 128
        // WIDE: park all 32 words of this thread's row; lane then copies piece wc of rows it*4 + wr, it = 0..7
        auto stage32 = [&](const uint32_t (&w)[32]) {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(stg + lane * WSTG_ROW_BYTES + i * 16), "r"(w[4 * i]),
                             "r"(w[4 * i + 1]), "r"(w[4 * i + 2]), "r"(w[4 * i + 3]) : "memory");
            __syncwarp();
        };
        auto unstage_w = [&](int it) {
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(stg + (it * 4 + wr) * WSTG_ROW_BYTES + wc * 16));
            return v;
        };
        // park 8 words of this thread's row, then hand each lane 4 consecutive words of rows rr and 16 + rr
        auto stage8 = [&](uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5, uint32_t w6, uint32_t w7) {
            __syncwarp();
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(stg + lane * STG_ROW_BYTES), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(stg + lane * STG_ROW_BYTES + 16), "r"(w4), "r"(w5), "r"(w6), "r"(w7) : "memory");
            __syncwarp();
        };
        auto unstage = [&](int it) {
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(stg + (it * 16 + rr) * STG_ROW_BYTES + rc * 16));
            return v;
        };
        // Tile coordinates are advanced incrementally and the sample index uses a multiply-high reciprocal: the integer
        // divisions of the first version (t / num_n, m0 / rows_per_sample, col / c, ...) were ~200 of the ~600
        // instructions each epilogue thread executed per tile (IABS / I2F.RP chains in profiles/r01_gemm_qkv_ncu.txt).
        int tm = worker / num_n, tn = worker % num_n;
        const int dm = nworkers / num_n, dn = nworkers % num_n;
        const uint32_t tiles_per_sample = (uint32_t)(p.rows_per_sample > 0 ? p.rows_per_sample / BM : 1);
        const uint32_t tps_magic = (uint32_t)((0x100000000ull + tiles_per_sample - 1) / tiles_per_sample);   // exact for mt, d < 2^16; d == 1 handled separately (2^32 does not fit)
        int lt = 0;
        for (int t = worker; t < num_tiles; t += nworkers, ++lt) {
            const int buf = lt & 1;
            const int mt = (PAIR ? 2 * tm : tm) + (int)rank;      // index of this CTA's 128-row tile
            const int m0 = mt * BM, n0 = tn * BN;
            tm += dm; tn += dn;
            if (tn >= num_n) { tn -= num_n; ++tm; }
            const int row0 = m0 + q * 32;             // first row of this warp; this thread computes row0 + lane
            const int col = n0 + ch * 32;
            // a 128-row tile never straddles samples (rows_per_sample % 128 == 0)
            const int sample = (EPI == EPI_GATE_RESID || EPI == EPI_QKV) ? (tiles_per_sample == 1 ? mt : (int)__umulhi((uint32_t)mt, tps_magic)) : 0;
            float4 xres[8];                           // EPI_GATE_RESID: the residual values this lane will update
            if constexpr (EPI == EPI_GATE_RESID) {
                if constexpr (WIDE) {
#pragma unroll
                    for (int it = 0; it < 8; ++it)
                        xres[it] = *reinterpret_cast<const float4*>(p.out + (size_t)(row0 + it * 4 + wr) * p.ldo + col + wc * 4);
                } else {
#pragma unroll
                    for (int ps = 0; ps < 4; ++ps)
#pragma unroll
                        for (int it = 0; it < 2; ++it)
                            xres[ps * 2 + it] = *reinterpret_cast<const float4*>(p.out + (size_t)(row0 + it * 16 + rr) * p.ldo + col + ps * 8 + rc * 4);
                }
            }
            mbar_wait(tfull(buf), ((uint32_t)lt >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem + buf * BN + ((uint32_t)(q * 32) << 16);
            uint32_t raw[32];
            tmem_ld32(taddr + ch * 32, raw);
            tmem_ld_wait();
            // the accumulator chunk is in registers: release it to the MMA warp before the math / stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_cluster(tempty(buf), 0);      // the leader's MMA warp waits for both CTAs
                else mbar_arrive(tempty(buf));
            }
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
#ifdef PDK_DBG_NO_EPI
            if (v[0] == 123.456f && col == -1) p.out[row0] = v[1];
            continue;
#endif
            if constexpr (EPI == EPI_STORE || EPI == EPI_GATE_RESID) {
                if (p.bias) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col) + i);
                        v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
                    }
                }
                if constexpr (EPI == EPI_STORE) {
                    if (p.act_silu) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = silu_fast(v[i]);
                    }
                } else {
                    const float* gate = p.gate + (size_t)sample * p.gate_stride + col;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gate) + i);
                        v[4 * i] *= g4.x; v[4 * i + 1] *= g4.y; v[4 * i + 2] *= g4.z; v[4 * i + 3] *= g4.w;
                    }
                }
                if constexpr (WIDE) {
                    uint32_t w[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(v[i]);
                    stage32(w);
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const uint4 u = unstage_w(it);
                        float4* dst = reinterpret_cast<float4*>(p.out + (size_t)(row0 + it * 4 + wr) * p.ldo + col + wc * 4);
                        float4 o = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
                        if constexpr (EPI == EPI_GATE_RESID) {
                            const float4 x = xres[it];
                            o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
                        }
                        *dst = o;
                    }
                } else
#pragma unroll
                for (int ps = 0; ps < 4; ++ps) {
                    stage8(__float_as_uint(v[ps * 8]), __float_as_uint(v[ps * 8 + 1]), __float_as_uint(v[ps * 8 + 2]), __float_as_uint(v[ps * 8 + 3]),
                           __float_as_uint(v[ps * 8 + 4]), __float_as_uint(v[ps * 8 + 5]), __float_as_uint(v[ps * 8 + 6]), __float_as_uint(v[ps * 8 + 7]));
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const uint4 u = unstage(it);
                        float4* dst = reinterpret_cast<float4*>(p.out + (size_t)(row0 + it * 16 + rr) * p.ldo + col + ps * 8 + rc * 4);
                        float4 o = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
                        if constexpr (EPI == EPI_GATE_RESID) {
                            const float4 x = xres[ps * 2 + it];
                            o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
                        }
                        *dst = o;
                    }
                }
            } else if constexpr (EPI == EPI_SWIGLU) {
                // W rows interleaved in blocks of 16: columns [0,16) = w1 rows, [16,32) = w3 rows of hidden j0..j0+15
                const int j0 = col / 2;
                uint32_t w[16];     // words 0-7: hi halves of the 16 hidden values, 8-15: lo halves
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    split2(silu_fast(v[2 * i]) * v[16 + 2 * i], silu_fast(v[2 * i + 1]) * v[16 + 2 * i + 1], w[i], w[8 + i]);
#pragma unroll
                for (int ps = 0; ps < 2; ++ps) {
                    stage8(w[ps * 8], w[ps * 8 + 1], w[ps * 8 + 2], w[ps * 8 + 3], w[ps * 8 + 4], w[ps * 8 + 5], w[ps * 8 + 6], w[ps * 8 + 7]);
#ifndef PDK_DBG_NO_STORE
                    __half* plane = ps == 0 ? p.ph : p.pl;
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const uint4 u = unstage(it);
                        *reinterpret_cast<uint4*>(plane + (size_t)(row0 + it * 16 + rr) * p.ldp + j0 + rc * 8) = u;
                    }
#endif
                }
            } else {   // EPI_QKV: this chunk is one head of q, k or v
                const int which = col >> p.c_shift;                          // c is a power of two (128 or 512)
                const int head = (col & (p.c - 1)) / kHeadDim;
                const int H = p.c / kHeadDim;
                if (which < 2) {    // per-head RMSNorm (rms_norm.py:14-19); q additionally carries log2e/sqrt(32)
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;      // four partial sums: a 32-deep FMA chain is pure latency
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        s0 = fmaf(v[i], v[i], s0); s1 = fmaf(v[i + 1], v[i + 1], s1);
                        s2 = fmaf(v[i + 2], v[i + 2], s2); s3 = fmaf(v[i + 3], v[i + 3], s3);
                    }
                    const float ss = (s0 + s1) + (s2 + s3);
                    const float inv = (inv_sqrt(ss * (1.0f / kHeadDim) + p.rms_eps)) * (which == 0 ? p.q_scale : 1.0f);
                    const float* gain = which == 0 ? p.norm_q : p.norm_k;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gain) + i);
                        v[4 * i] *= inv * g4.x; v[4 * i + 1] *= inv * g4.y; v[4 * i + 2] *= inv * g4.z; v[4 * i + 3] *= inv * g4.w;
                    }
                }
                __half* dbase = which == 0 ? p.q : (which == 1 ? p.k : p.v);      // row = [hi 32 | lo 32] halves = 128 bytes
                const size_t tile_row = (size_t)(sample * H + head) * p.rows_per_sample + (row0 - sample * p.rows_per_sample);
                if constexpr (WIDE) {
                    uint32_t w[32];                   // the 128-byte output row: 16 words hi | 16 words lo
#pragma unroll
                    for (int i = 0; i < 16; ++i) split2(v[2 * i], v[2 * i + 1], w[i], w[16 + i]);
                    stage32(w);
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const uint4 u = unstage_w(it);
                        *reinterpret_cast<uint4*>(dbase + (tile_row + it * 4 + wr) * (2 * kHeadDim) + wc * 8) = u;
                    }
                } else {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
#pragma unroll
                for (int ps = 0; ps < 4; ++ps) {      // 32-byte pieces of the 128-byte row: hi 0-15, hi 16-31, lo 0-15, lo 16-31
                    const uint32_t* w = ps < 2 ? hi + (ps & 1) * 8 : lo + (ps & 1) * 8;
                    stage8(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const uint4 u = unstage(it);
                        *reinterpret_cast<uint4*>(dbase + (tile_row + it * 16 + rr) * (2 * kHeadDim) + ps * 16 + rc * 8) = u;
                    }
                }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync();       // nobody leaves while the partner may still touch its smem / TMEM / barriers
    if (warp == 1) {
        if constexpr (PAIR) tmem_dealloc_2sm(tmem, TMEM_COLS);
        else tmem_dealloc(tmem, TMEM_COLS);
    }
}

template <int EPI, bool PAIR, bool WIDE>
cudaError_t launch_variant(const GemmArgs& a, cudaStream_t st, int num_sms) {
    constexpr int SMEM_BYTES = smem_bytes<PAIR, WIDE>();
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
    static PerDevice configured;
    cudaError_t e;
    if ((e = ensure_smem(configured, gemm_umma_kernel<EPI, PAIR, WIDE>, SMEM_BYTES)) != cudaSuccess) return e;
    CUtensorMap mAh, mAl, mWh, mWl;
    constexpr int W_ROWS = PAIR ? BN / 2 : BN;
    if ((e = get_tensor_map_f16(a.Ah, a.M, a.K, a.lda, BM, BK, 128, &mAh)) != cudaSuccess) return e;
    if ((e = get_tensor_map_f16(a.Al, a.M, a.K, a.lda, BM, BK, 128, &mAl)) != cudaSuccess) return e;
    if ((e = get_tensor_map_f16(a.Wh, a.N, a.K, a.ldw, W_ROWS, BK, 128, &mWh)) != cudaSuccess) return e;
    if ((e = get_tensor_map_f16(a.Wl, a.N, a.K, a.ldw, W_ROWS, BK, 128, &mWl)) != cudaSuccess) return e;
    const int tiles = (a.M / (PAIR ? 2 * BM : BM)) * (a.N / BN);
    const int workers = PAIR ? num_sms / 2 : num_sms;
    const int grid = (tiles < workers ? tiles : workers) * (PAIR ? 2 : 1);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = !measure_switch("PDK_NO_PDL");
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = PAIR ? 2 : 1;
    PDK_LAUNCH_CHECK(cudaLaunchKernelEx(&cfg, gemm_umma_kernel<EPI, PAIR, WIDE>, mAh, mAl, mWh, mWl, a));
    return cudaGetLastError();
}

template <int EPI>
cudaError_t launch_one(const GemmArgs& a, cudaStream_t st) {
    int num_sms = 0;
    cudaError_t e;
    if ((e = device_sm_count(&num_sms)) != cudaSuccess) return e;
    // measurement switches: compile-time false in the release build (common.cuh)
    static const bool allow_pair = !measure_switch("PDK_NO_PAIR");
    static const bool force_pair = measure_switch("PDK_FORCE_PAIR");
    // see the header: of the K = 128 atom shapes only the wide ones at large M gain (atom QKV 16.4 -> 15.8, downscale 17.6 -> 16.8 us
    // at 32768 rows; +0.2 us at 2048-10240 rows); the atom out-proj (N = 128) always loses
    const bool big = (long long)a.K * a.N >= 250000 || (a.M >= 16384 && a.N >= 384);
    static const bool allow_wide = !measure_switch("PDK_NO_WIDE");
    const bool pair = allow_pair && a.M % (2 * BM) == 0 && (big || force_pair);
    if constexpr (EPI == EPI_SWIGLU) {        // 32-byte rows per plane either way: always the narrow staging
        return pair ? launch_variant<EPI, true, false>(a, st, num_sms) : launch_variant<EPI, false, false>(a, st, num_sms);
    } else {
        static const bool wide_all = measure_switch("PDK_WIDE_ALL");
        const bool wide = allow_wide && (pair || a.K <= 2 * BK || wide_all);
        if (pair) return wide ? launch_variant<EPI, true, true>(a, st, num_sms) : launch_variant<EPI, true, false>(a, st, num_sms);
        return wide ? launch_variant<EPI, false, true>(a, st, num_sms) : launch_variant<EPI, false, false>(a, st, num_sms);
    }
}

}  // namespace

cudaError_t launch_gemm(GemmEpilogue epi, const GemmArgs& a, cudaStream_t st) {
    if (a.M <= 0 || a.N <= 0 || a.K <= 0 || a.M % BM || a.N % BN || a.K % BK) return cudaErrorInvalidValue;
    if (a.lda % 8 || a.ldw % 8) return cudaErrorInvalidValue;   // 16-byte global strides for TMA
    if ((epi == EPI_GATE_RESID || epi == EPI_QKV) && (a.rows_per_sample <= 0 || a.rows_per_sample % BM))
        return cudaErrorInvalidValue;                            // a row tile must not straddle samples
    switch (epi) {
        case EPI_STORE: return launch_one<EPI_STORE>(a, st);
        case EPI_GATE_RESID: return launch_one<EPI_GATE_RESID>(a, st);
        case EPI_SWIGLU: return launch_one<EPI_SWIGLU>(a, st);
        case EPI_QKV: {
            if (a.N != 3 * a.c || a.c % BN || (a.c & (a.c - 1))) return cudaErrorInvalidValue;      // c: power of two >= 128
            GemmArgs b = a;
            b.c_shift = 0;
            while ((1 << b.c_shift) < a.c) ++b.c_shift;
            return launch_one<EPI_QKV>(b, st);
        }
    }
    return cudaErrorInvalidValue;
}

}  // namespace pdk
