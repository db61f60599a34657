// Dense pair-biased attention on tcgen05 / TMEM / TMA  (replaces F.scaled_dot_product_attention(q,k,v,attn_bias)
// in DiTAttention, reference PhysDock/models/primitives/attentions.py:259-260).
//
// softmax(q k^T / sqrt(32) + bias) v with a dense fp32 bias [H,S,S] shared by all samples.  q (pre-scaled) and
// bias arrive multiplied by log2(e): the softmax runs on exp2.  All operands are split-fp16 planes.
//
// CTA = (head h, 128 query rows, a group of up to 4 samples).  The samples of a group share every bias tile: it
// is TMA-loaded once per key tile and read by all of them from shared memory, which divides the dominant L2
// stream (bias, 4 B per score) by the group size.  Work unit u = (key tile j of 64 keys, sample g).
//
//   warp 16 TMA producer   : Q of the group (once); per key tile the bias tile [128 x 64] fp32 (two SWIZZLE_128B
//                            boxes, double buffered) and per unit K, V [64 keys x 64 halves] (6-stage ring).  q/k/v
//                            rows are stored interleaved [hi 32 | lo 32] = 128 bytes because TMA boxes narrower than
//                            128 B run at less than half rate (tests/cuda/umma_probe.cu test 7); the hi and lo K=16
//                            slices of a row are then just byte offsets 0/32 and 64/96 into a SWIZZLE_128B tile.
//   warp 17 QK issuer      : S_g = Q_g K^T   (SS, M128 N64 K16 x 2 slices x 3 split products) into TMEM buffer g
//   warp 18 PV issuer      : O_g += P_g V    (TS: P from TMEM, V MN-major from smem; 4 slices x 3 products)
//                            Two issuing warps because one warp's serialized waits/commits left the tensor pipe
//                            idle ~40% of the time; QK(j,g) is ordered after
//                            PV(j-1,g) through the pv_done barrier (S and P share a TMEM buffer).
//   warps 4g .. 4g+3      : softmax warpgroup of sample g (FOUR warpgroups; with two, each serving two samples, the
//                            softmax warps were busy 87% of the time and set the pace; now they wait ~1000 cycles per unit for S:
//                            profiles/r01_attention_timeline.txt).
//                            Thread = one query row: reads its 64 scores with tcgen05.ld, adds the bias row, row max /
//                            exp2 / row sum entirely in registers (no shuffles), splits P into fp16 hi/lo and writes it
//                            back over S with tcgen05.st (S and P alias) in 16-column pieces so that the thread stays
//                            within 104 registers.  O is rescaled lazily (only when the row max grows by more than
//                            2^8), directly in TMEM.
// TMEM: 4 x 64 columns S/P + 4 x 64 columns O (P_hi V_hi + P_lo V_hi | P_hi V_lo, summed in the epilogue).
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

#include <algorithm>
#include <mutex>
#include <vector>

namespace pdk {

namespace {

constexpr int G = 4;                 // samples per CTA
constexpr int BQ = 128, BKV = 64, D = kHeadDim;
constexpr int NS = 6;                // K/V ring stages
constexpr int Q_TILE = BQ * 2 * D * 2;               // 16 KB: 128 rows x [hi|lo] 128 B
constexpr int KV_TILE = BKV * 2 * D * 2;             // 8 KB: 64 rows x 128 B
constexpr int KV_STAGE = 2 * KV_TILE;                // K, V
constexpr int BIAS_HALF = BQ * 32 * 4;               // 16 KB: [128 rows][32 fp32]
constexpr int BIAS_TILE = 2 * BIAS_HALF;             // 32 KB
constexpr int OFF_Q = 0;
constexpr int OFF_BIAS = G * Q_TILE;                 // 64 KB
constexpr int OFF_KV = OFF_BIAS + 2 * BIAS_TILE;     // 128 KB
constexpr int SMEM_BYTES = OFF_KV + NS * KV_STAGE + 1024;   // 225 KB
constexpr int NTHREADS = (4 * G + 3) * 32;       // 16 softmax warps + TMA + QK + PV = 608 threads (<= 104 registers each)
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_S = 0, COL_O = 256;
constexpr float kRescaleThreshold = 8.0f;            // log2 domain: P <= 2^8 stays exact enough in fp16 hi/lo

// debug timeline: slot layout trace[(unit * 8 + k)]
#define TRACE(unit, k) do { if (p.trace != nullptr && blockIdx.x == 0 && lane == 0) p.trace[(unit) * 8 + (k)] = clock64(); } while (0)

struct Bars {
    uint64_t q_full;
    uint64_t kv_full[NS], kv_empty[NS];
    uint64_t bias_full[2], bias_empty[2];
    uint64_t s_full[G], p_ready[G], pv_done[G];
};

PDK_DEV void tmem_st32(uint32_t addr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
          "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),
          "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
PDK_DEV void tmem_ld16(uint32_t addr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(addr));
}
PDK_DEV void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
PDK_DEV void tmem_st16(uint32_t addr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
PDK_DEV void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
PDK_DEV void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate));
}
PDK_DEV float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// ---- packed fp32x2 arithmetic (FADD2 / FMNMX3 on sm_100): halves the issue slots of the softmax inner loops
PDK_DEV uint64_t pack2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
PDK_DEV void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
PDK_DEV uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
PDK_DEV float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
// p0, p1 in [0, 2^8] -> packed fp16 (hi, lo) planes.  hi = p truncated to 11 significant bits (cvt.rz), whose fp32
// value is p & 0xffffe000 (exact for p >= 2^-14; below that the mismatch is < 2^-25 absolute); lo = fp16(p - hi).
PDK_DEV void split2_pos(float p0, float p1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rz.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(p1), "f"(p0));
    const float n0 = __uint_as_float((__float_as_uint(p0) & 0xffffe000u) ^ 0x80000000u);
    const float n1 = __uint_as_float((__float_as_uint(p1) & 0xffffe000u) ^ 0x80000000u);
    float l0, l1;
    unpack2(add2(pack2(p0, p1), pack2(n0, n1)), l0, l1);
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(l1), "f"(l0));
}

// Work list: one entry per CTA = (head, query tile, first sample, number of samples <= G), packed 8 bits each, ordered
// by decreasing sample count.  With the plain (B/G, S/128, H) grid the atom attention of the benchmark shape is 256
// equal CTAs on 148 SMs = 1.73 waves (13% of the machine idle).  The host instead splits the B samples of some
// (head, query tile) pairs into one more, smaller group (16 = 4+4+4+4 or 4+3+3+3+3) so that every SM gets one
// 4-sample and one 3-sample CTA: 7 sample-units per SM instead of 8 (launch_attention below).
constexpr int kMaxWork = 800;       // keeps the kernel parameters under 4 KB
struct AttnWork {
    uint32_t e[kMaxWork];
};

__global__ void __launch_bounds__(NTHREADS, 1)
attention_umma_kernel(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mK,
                      const __grid_constant__ CUtensorMap mV, const __grid_constant__ CUtensorMap mBias, const AttnArgs p,
                      const __grid_constant__ AttnWork work) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) Bars bars;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t we = work.e[blockIdx.x];
    const int h = (int)(we & 0xffu), qt = (int)((we >> 8) & 0xffu), b0 = p.b_base + (int)((we >> 16) & 0xffu), ng = (int)(we >> 24);
    const int S = p.S_pad;
    const int NJ = S / BKV;
    const uint32_t sm = (smem_u32(smem_raw) + 1023u) & ~1023u;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bars.q_full), 1);
        for (int s = 0; s < NS; ++s) { mbar_init(smem_u32(&bars.kv_full[s]), 1); mbar_init(smem_u32(&bars.kv_empty[s]), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&bars.bias_full[b]), 1); mbar_init(smem_u32(&bars.bias_empty[b]), ng * 128); }
        for (int g = 0; g < G; ++g) {
            mbar_init(smem_u32(&bars.s_full[g]), 1);
            mbar_init(smem_u32(&bars.p_ready[g]), 128);
            mbar_init(smem_u32(&bars.pv_done[g]), 1);
        }
        mbar_fence_init();
    }
    griddep_launch();                 // PDL (see common.cuh)
    if (warp == 4 * G + 1) tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    griddep_wait();

    if (warp == 4 * G) {
        // ================================================================= TMA producer
        if (elect_one()) {
            tma_prefetch_desc(&mQ); tma_prefetch_desc(&mK); tma_prefetch_desc(&mV); tma_prefetch_desc(&mBias);
            const uint32_t qbar = smem_u32(&bars.q_full);
            mbar_expect_tx(qbar, ng * Q_TILE);
            for (int g = 0; g < ng; ++g)
                tma_load_2d(sm + OFF_Q + g * Q_TILE, &mQ, qbar, 0, ((b0 + g) * p.H + h) * S + qt * BQ);
        }
        __syncwarp();
        int st = 0;
        uint32_t kv_par = 0;
        for (int j = 0; j < NJ; ++j) {
            const int bb = j & 1;
            mbar_wait(smem_u32(&bars.bias_empty[bb]), (((uint32_t)j >> 1) & 1u) ^ 1u);
            if (elect_one()) {
                const uint32_t bbar = smem_u32(&bars.bias_full[bb]);
                mbar_expect_tx(bbar, BIAS_TILE);
                tma_load_2d(sm + OFF_BIAS + bb * BIAS_TILE, &mBias, bbar, j * BKV, h * S + qt * BQ);
                tma_load_2d(sm + OFF_BIAS + bb * BIAS_TILE + BIAS_HALF, &mBias, bbar, j * BKV + 32, h * S + qt * BQ);
            }
            __syncwarp();
            for (int g = 0; g < ng; ++g) {
                mbar_wait(smem_u32(&bars.kv_empty[st]), kv_par ^ 1u);
                if (elect_one()) {
                    const uint32_t kbar = smem_u32(&bars.kv_full[st]);
                    mbar_expect_tx(kbar, KV_STAGE);
                    const int row = ((b0 + g) * p.H + h) * S + j * BKV;
                    const uint32_t dst = sm + OFF_KV + st * KV_STAGE;
                    tma_load_2d(dst, &mK, kbar, 0, row);
                    tma_load_2d(dst + KV_TILE, &mV, kbar, 0, row);
                }
                __syncwarp();
                if (++st == NS) { st = 0; kv_par ^= 1u; }
            }
        }
    } else if (warp == 4 * G + 1) {
        // ================================================================= QK issuer (whole warp loops, one lane issues)
        // S_g(j) = Q_g K_g(j)^T.  The S buffer of sample g doubles as its P buffer, so QK(j,g) waits for PV(j-1,g).
        constexpr uint32_t idesc_qk = umma_idesc_f16(BQ, BKV);          // M128 N64, both K-major
        mbar_wait(smem_u32(&bars.q_full), 0);
        int g = 0, st = 0;
        uint32_t kv_par = 0;
        for (int j = 0; j < NJ; ++j) {
            for (g = 0; g < ng; ++g) {
                mbar_wait(smem_u32(&bars.kv_full[st]), kv_par);
                if (j > 0) mbar_wait(smem_u32(&bars.pv_done[g]), (uint32_t)(j - 1) & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t q = sm + OFF_Q + g * Q_TILE, k = sm + OFF_KV + st * KV_STAGE;
                    // K-major SWIZZLE_128B tiles; hi halves at byte 0 of each row, lo halves at byte 64
                    const uint64_t qh = smem_desc(q, 1024, kLayoutSw128), ql = smem_desc(q + 64, 1024, kLayoutSw128);
                    const uint64_t kh = smem_desc(k, 1024, kLayoutSw128), kl = smem_desc(k + 64, 1024, kLayoutSw128);
                    const uint32_t d = tmem + COL_S + g * BKV;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint64_t o = (uint64_t)(ks * 2);
                        umma_f16(d, ql + o, kh + o, idesc_qk, ks != 0);
                        umma_f16(d, qh + o, kl + o, idesc_qk, 1u);
                        umma_f16(d, qh + o, kh + o, idesc_qk, 1u);
                    }
                    umma_commit(smem_u32(&bars.s_full[g]));
                }
                __syncwarp();
                TRACE(j * ng + g, 2);                          // QK issued
                if (++st == NS) { st = 0; kv_par ^= 1u; }
            }
        }
    } else if (warp == 4 * G + 2) {
        // ================================================================= PV issuer
        // O_g += P_g(j) V_g(j): P from TMEM (written by the softmax threads over S), V MN-major from smem.
        // A V row is [V_hi 32 | V_lo 32], i.e. the MN-major SWIZZLE_128B tile IS the N = 64 operand [V_hi | V_lo]:
        // P_hi [V_hi|V_lo] is one MMA (O columns 0-31 and 32-63), P_lo V_hi a second one (N = 32 sub-read of the same
        // tile) onto columns 0-31.  2 MMAs per slice instead of 3 (every M128 MMA costs >= 45 cycles).
        constexpr uint32_t idesc_pv = umma_idesc_f16(BQ, D, true);      // M128 N32, B (= V) MN-major
        constexpr uint32_t idesc_pv2 = umma_idesc_f16(BQ, 2 * D, true); // M128 N64
        int g = 0, st = 0;
        uint32_t kv_par = 0;
        for (int j = 0; j < NJ; ++j) {
            for (g = 0; g < ng; ++g) {
                mbar_wait(smem_u32(&bars.kv_full[st]), kv_par);        // visibility of the V tile to this warp
                mbar_wait(smem_u32(&bars.p_ready[g]), (uint32_t)j & 1u);
                TRACE(j * ng + g, 0);                          // PV warp saw p_ready
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t vd = smem_desc(sm + OFF_KV + st * KV_STAGE + KV_TILE, 1024, kLayoutSw128);
                    const uint32_t pa = tmem + COL_S + g * BKV, d = tmem + COL_O + g * 2 * D;
#pragma unroll
                    for (int ks = 0; ks < BKV / 16; ++ks) {
                        const uint64_t o = (uint64_t)((ks * 16 * 128) >> 4);    // 16 key rows of 128 bytes
                        umma_f16_ts(d, pa + ks * 8, vd + o, idesc_pv2, (j | ks) != 0);      // P_hi [V_hi | V_lo]
                        umma_f16_ts(d, pa + 32 + ks * 8, vd + o, idesc_pv, 1u);             // P_lo V_hi
                    }
                    umma_commit(smem_u32(&bars.kv_empty[st]));
                    umma_commit(smem_u32(&bars.pv_done[g]));
                }
                __syncwarp();
                TRACE(j * ng + g, 1);                          // PV issued
                if (++st == NS) { st = 0; kv_par ^= 1u; }
            }
        }
    } else {
        // ================================================================= softmax warpgroups (one per sample)
        const int g = warp >> 2;
        const int row = (warp & 3) * 32 + lane;                 // query row inside the tile == TMEM lane
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t bias_row = (uint32_t)row * 128u;
        const uint32_t sw = (uint32_t)(row & 7);
        if (g < ng) {
            float m_run = 0.f, l_run = 0.f;
            const uint32_t ts = tl + COL_S + g * BKV, to = tl + COL_O + g * 2 * D;
            for (int j = 0; j < NJ; ++j) {
                if ((warp & 3) == 0) TRACE(j * ng + g, 3);     // softmax starts waiting for S
                mbar_wait(smem_u32(&bars.s_full[g]), (uint32_t)j & 1u);
                if ((warp & 3) == 0) TRACE(j * ng + g, 4);     // S arrived
                tc_fence_after();
                uint64_t s2[32];                         // the 64 scores of this row as 32 fp32x2 pairs
                float mx = -INFINITY;
                const uint32_t bt = sm + OFF_BIAS + (j & 1) * BIAS_TILE + bias_row;
                mbar_wait(smem_u32(&bars.bias_full[j & 1]), ((uint32_t)j >> 1) & 1u);
                // both 32-column halves of S in flight with ONE wait (two ld + wait pairs exposed the TMEM read latency twice per
                // unit: atom attention 127.1 -> 125.2 us, token 13.3 -> 12.7 us, and the register allocator no longer spills)
                uint32_t r[64];
                tmem_ld32(ts, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
                tmem_ld32(ts + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
                tmem_ld_wait();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 b4 = lds128(bt + half * BIAS_HALF + (((uint32_t)c ^ sw) << 4));
                        s2[16 * half + 2 * c] = add2(pack2(__uint_as_float(r[32 * half + 4 * c]), __uint_as_float(r[32 * half + 4 * c + 1])), pack2(b4.x, b4.y));
                        s2[16 * half + 2 * c + 1] = add2(pack2(__uint_as_float(r[32 * half + 4 * c + 2]), __uint_as_float(r[32 * half + 4 * c + 3])), pack2(b4.z, b4.w));
                        float a0, a1, a2, a3;
                        unpack2(s2[16 * half + 2 * c], a0, a1);
                        unpack2(s2[16 * half + 2 * c + 1], a2, a3);
                        mx = max3(mx, a0, a1);
                        mx = max3(mx, a2, a3);
                    }
                }
                mbar_arrive(smem_u32(&bars.bias_empty[j & 1]));
                if ((warp & 3) == 0) TRACE(j * ng + g, 5);     // S + bias in registers, row max known
                // ---- running max with lazy rescale of O (in TMEM)
                if (j == 0) {
                    m_run = mx;
                } else {
                    mbar_wait(smem_u32(&bars.pv_done[g]), (uint32_t)(j - 1) & 1u);     // O_g is stable
                    const bool need = mx > m_run + kRescaleThreshold;
                    if (__any_sync(0xffffffffu, need)) {
                        tc_fence_after();
                        const float c = need ? ex2(m_run - mx) : 1.0f;
                        if (need) m_run = mx;
                        l_run *= c;
#pragma unroll 1
                        for (int q4 = 0; q4 < 4; ++q4) {       // 16 columns at a time: rare path, keep it out of the register budget
                            uint32_t o[16];
                            tmem_ld16(to + q4 * 16, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * c);
                            tmem_st16(to + q4 * 16, o);
                        }
                    }
                }
                // ---- P = exp2(S - m), row sum, split to fp16 hi/lo, back to TMEM over S (P_hi: columns 0-31, P_lo: 32-63)
                const uint64_t negm = pack2(-m_run, -m_run);
                uint64_t sum2 = pack2(0.f, 0.f);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float x0, x1;
                        unpack2(add2(s2[q4 * 8 + k], negm), x0, x1);
                        const float p0 = ex2(x0), p1 = ex2(x1);
                        sum2 = add2(sum2, pack2(p0, p1));
                        split2_pos(p0, p1, hi[k], lo[k]);
                    }
                    tmem_st8(ts + q4 * 8, hi);
                    tmem_st8(ts + 32 + q4 * 8, lo);
                }
                float sum, sum_b;
                unpack2(sum2, sum, sum_b);
                l_run += sum + sum_b;
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(smem_u32(&bars.p_ready[g]));
                if ((warp & 3) == 0) TRACE(j * ng + g, 6);     // P published
            }
            // ---- epilogue: O / l -> split planes (O columns 0-31: P_hi V_hi + P_lo V_hi, 32-63: P_hi V_lo)
            mbar_wait(smem_u32(&bars.pv_done[g]), (uint32_t)(NJ - 1) & 1u);
            tc_fence_after();
            const float inv = 1.0f / l_run;
            // Row-contiguous write-out: a thread owns one row, so storing from registers makes every store instruction touch
            // 32 different 128-byte lines (32 L1 wavefronts).  The rows are parked in this sample's Q tile (dead: its last
            // QK MMA has completed) with 80-byte rows and written as 8 rows x 64 bytes per instruction.
            const uint32_t stg = sm + OFF_Q + g * Q_TILE + (warp & 3) * (32 * 80);
            const int wr = lane >> 2, wc = lane & 3;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t o[16], o2[16];
                tmem_ld16(to + half * 16, o);
                tmem_ld16(to + 32 + half * 16, o2);
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    split2((__uint_as_float(o[2 * k]) + __uint_as_float(o2[2 * k])) * inv,
                           (__uint_as_float(o[2 * k + 1]) + __uint_as_float(o2[2 * k + 1])) * inv, hi[half * 8 + k], lo[half * 8 + k]);
            }
#pragma unroll
            for (int plane = 0; plane < 2; ++plane) {       // 0: hi, 1: lo
                const uint32_t* w = plane == 0 ? hi : lo;
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(stg + lane * 80 + i * 16), "r"(w[4 * i]), "r"(w[4 * i + 1]),
                                 "r"(w[4 * i + 2]), "r"(w[4 * i + 3]) : "memory");
                __syncwarp();
                __half* dst = plane == 0 ? p.oh : p.ol;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    uint4 u;
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                                 : "r"(stg + (it * 8 + wr) * 80 + wc * 16));
                    const size_t off = ((size_t)(b0 + g) * S + (size_t)qt * BQ + (warp & 3) * 32 + it * 8 + wr) * p.c + (size_t)h * D + wc * 8;
                    *reinterpret_cast<uint4*>(dst + off) = u;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4 * G + 1) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace

static long long* g_trace = nullptr;
void set_attention_trace(long long* buf) { g_trace = buf; }

namespace {

// Greedy list scheduling (CTAs are dispatched in order to the first SM that frees up) of a work list in which `y`
// of the (head, query tile) pairs use the finer sample partition: returns the makespan in sample-units.
struct Partition { int n; int sizes[64]; };
Partition make_partition(int B, int groups) {       // `groups` groups, sizes as even as possible
    Partition pt{};
    pt.n = groups;
    for (int i = 0; i < groups; ++i) pt.sizes[i] = B / groups + (i < B % groups ? 1 : 0);
    return pt;
}
double makespan(const std::vector<int>& sizes_desc, int sms, double fixed) {
    std::vector<double> sm_free(sms, 0.0);
    for (int sz : sizes_desc) {
        int best = 0;
        for (int i = 1; i < sms; ++i) if (sm_free[i] < sm_free[best]) best = i;
        sm_free[best] += sz + fixed;
    }
    double m = 0;
    for (double v : sm_free) m = v > m ? v : m;
    return m;
}

struct WorkList { int B, H, QT, sms; int n; AttnWork w; };

// Builds (and caches) the work list for a shape.  Returns nullptr when the shape needs more than kMaxWork CTAs.
const WorkList* get_work_list(int B, int H, int QT, int sms) {
    static std::mutex mu;
    static std::vector<WorkList*> cache;
    std::lock_guard<std::mutex> lock(mu);
    for (const WorkList* w : cache) if (w->B == B && w->H == H && w->QT == QT && w->sms == sms) return w;
    const int pairs = H * QT;
    const int k0 = (B + G - 1) / G;
    if (H > 255 || QT > 255 || B > 255 || (long long)pairs * (k0 + 1) > kMaxWork) return nullptr;
    const Partition coarse = make_partition(B, k0);
    const Partition fine = (k0 + 1 <= B) ? make_partition(B, k0 + 1) : coarse;
    static const bool no_balance = measure_switch("PDK_NO_ATTN_BALANCE");
    const double fixed = 0.15;        // per-CTA prologue + epilogue in sample-units (one unit = all key tiles of one sample)
    int best_y = 0;
    double best = 1e30;
    for (int y = 0; y <= (no_balance ? 0 : pairs); ++y) {
        std::vector<int> sizes;
        for (int pr = 0; pr < pairs; ++pr) {
            const Partition& pt = pr < y ? fine : coarse;
            for (int i = 0; i < pt.n; ++i) sizes.push_back(pt.sizes[i]);
        }
        std::sort(sizes.begin(), sizes.end(), [](int a, int b) { return a > b; });
        const double m = makespan(sizes, sms, fixed);
        if (m < best - 1e-9) { best = m; best_y = y; }
    }
    WorkList* wl = new WorkList{};
    wl->B = B; wl->H = H; wl->QT = QT; wl->sms = sms;
    struct Item { int ng, h, qt, b0; };
    std::vector<Item> items;
    // the pairs that use the fine partition are spread over heads and query tiles (pair index strided)
    for (int pr = 0; pr < pairs; ++pr) {
        const Partition& pt = pr < best_y ? fine : coarse;
        const int h = pr % H, qt = pr / H;
        int b0 = 0;
        for (int i = 0; i < pt.n; ++i) { items.push_back({pt.sizes[i], h, qt, b0}); b0 += pt.sizes[i]; }
    }
    std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.ng > b.ng; });
    wl->n = (int)items.size();
    for (int i = 0; i < wl->n; ++i)
        wl->w.e[i] = (uint32_t)items[i].h | ((uint32_t)items[i].qt << 8) | ((uint32_t)items[i].b0 << 16) | ((uint32_t)items[i].ng << 24);
    cache.push_back(wl);
    return wl;
}

}  // namespace

// Host-only view of the work list (C ABI: pdk_attention_work_list; checked without a GPU in tests/test_cabi.py)
int attention_work_list(int B, int H, int QT, int sms, uint32_t* out, int cap) {
    const WorkList* wl = get_work_list(B, H, QT, sms);
    if (wl == nullptr) return -1;
    if (wl->n > cap) return -2;
    for (int i = 0; i < wl->n; ++i) out[i] = wl->w.e[i];
    return wl->n;
}

cudaError_t launch_attention(const AttnArgs& a_in, cudaStream_t st) {
    AttnArgs a = a_in;
    a.trace = g_trace;
    if (a.S_pad <= 0 || a.S_pad % BQ || a.c != a.H * D || a.B <= 0) return cudaErrorInvalidValue;
    static PerDevice configured;
    int num_sms = 0;
    cudaError_t e;
    if ((e = ensure_smem(configured, attention_umma_kernel, SMEM_BYTES)) != cudaSuccess) return e;
    if ((e = device_sm_count(&num_sms)) != cudaSuccess) return e;
    const uint64_t rows = (uint64_t)a.B * a.H * a.S_pad;
    CUtensorMap mQ, mK, mV, mBias;
    if ((e = get_tensor_map_f16(a.q, rows, 2 * D, 2 * D, BQ, 2 * D, 128, &mQ)) != cudaSuccess) return e;
    if ((e = get_tensor_map_f16(a.k, rows, 2 * D, 2 * D, BKV, 2 * D, 128, &mK)) != cudaSuccess) return e;
    if ((e = get_tensor_map_f16(a.v, rows, 2 * D, 2 * D, BKV, 2 * D, 128, &mV)) != cudaSuccess) return e;
    if ((e = get_tensor_map_f32(a.bias, (uint64_t)a.H * a.S_pad, a.S_pad, a.S_pad, BQ, 32, 128, &mBias)) != cudaSuccess) return e;
    // Samples are processed in chunks of <= 255 (8-bit fields of the work list); shapes whose list does not fit in the
    // kernel parameters are launched as several lists.
    const int QT = a.S_pad / BQ;
    int b_done = 0;
    while (b_done < a.B) {
        int bc = a.B - b_done;
        if (bc > 252) bc = 252;
        while ((long long)a.H * QT * ((bc + G - 1) / G + 1) > kMaxWork && bc > G) bc = ((bc / 2 + G - 1) / G) * G;
        const WorkList* wl = get_work_list(bc, a.H, QT, num_sms);
        if (wl == nullptr) return cudaErrorInvalidValue;      // H * QT alone exceeds the list: not a PhysDock shape
        AttnArgs c = a;
        c.b_base = b_done;
        PDK_LAUNCH_CHECK(launch_pdl(attention_umma_kernel, dim3(wl->n), dim3(NTHREADS), (size_t)SMEM_BYTES, st, mQ, mK, mV, mBias, c, wl->w));
        b_done += bc;
    }
    return cudaGetLastError();
}

}  // namespace pdk
