// Dense pair-biased attention (v1: cp.async + ldmatrix + mma.sync flash kernel, split-fp16 operands).
//
// Replaces `F.scaled_dot_product_attention(q, k, v, attn_bias)` in DiTAttention
// (reference PhysDock/models/primitives/attentions.py:259-260): softmax(q k^T / sqrt(32) + bias) v with
// a dense additive bias [H,S,S] shared by all samples.  The bias (and the 1/sqrt(32) folded into q) arrive
// pre-multiplied by log2(e), so the softmax runs on exp2.
//
// CTA = 128 query rows of one (sample, head); 8 warps x 16 rows; key tiles of 64; two-stage cp.async ring
// holding K/V planes and the fp32 bias tile.  Grid is sample-fastest so the CTAs that share a bias tile
// are co-resident and hit it in L2.  QK^T and PV each run as 3 fp16 MMAs (hi*hi + hi*lo + lo*hi).
#include "common.cuh"
#include "kernels.h"

namespace pdk {

namespace {

constexpr int BQ = 128, BKV = 64, D = kHeadDim;
constexpr int Q_PLANE = BQ * D * 2;          // 8 KB
constexpr int KV_PLANE = BKV * D * 2;        // 4 KB
constexpr int BIAS_LD = BKV + 8;             // 72 floats: conflict-free float2 reads in C-fragment order
constexpr int BIAS_BYTES = BQ * BIAS_LD * 4; // 36 KB
constexpr int STAGE = 4 * KV_PLANE + BIAS_BYTES;      // K_hi, K_lo, V_hi, V_lo, bias = 52 KB
constexpr int SMEM = 2 * Q_PLANE + 2 * STAGE;         // 120 KB

__global__ void __launch_bounds__(256, 1) attention_kernel(const AttnArgs p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.x, qt = blockIdx.y, h = blockIdx.z;
    const int S = p.S_pad;
    const size_t head_off = ((size_t)b * p.H + h) * S * D;
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sStage = sQ + 2 * Q_PLANE;
    const float* bias_s0 = reinterpret_cast<const float*>(smem + 2 * Q_PLANE + 4 * KV_PLANE);

    const float* bias_g = p.bias + ((size_t)h * S + (size_t)qt * BQ) * S;

    auto load_kv = [&](int stage, int j) {
        const uint32_t sb = sStage + stage * STAGE;
        const size_t kv0 = head_off + (size_t)j * BKV * D;
        {   // K/V planes: 64 rows x 4 chunks = 256 chunks per plane -> one per thread
            const int row = tid >> 2, ch = tid & 3;
            const uint32_t off = swz64(row, ch);
            const size_t go = kv0 + (size_t)row * D + ch * 8;
            cp_async16(sb + off, p.kh + go);
            cp_async16(sb + KV_PLANE + off, p.kl + go);
            cp_async16(sb + 2 * KV_PLANE + off, p.vh + go);
            cp_async16(sb + 3 * KV_PLANE + off, p.vl + go);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {   // bias tile: 128 rows x 16 chunks of 4 floats
            const int q = tid + i * 256;
            const int row = q >> 4, ch = q & 15;
            cp_async16(sb + 4 * KV_PLANE + (row * BIAS_LD + ch * 4) * 4,
                       bias_g + (size_t)row * S + (size_t)j * BKV + ch * 4);
        }
    };

    // Q tile (both planes) + first K/V/bias tile
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int q = tid + i * 256;
        const int row = q >> 2, ch = q & 3;
        const size_t go = head_off + ((size_t)qt * BQ + row) * D + ch * 8;
        cp_async16(sQ + swz64(row, ch), p.qh + go);
        cp_async16(sQ + Q_PLANE + swz64(row, ch), p.ql + go);
    }
    load_kv(0, 0);
    cp_async_commit();

    uint32_t qh[2][4], ql[2][4];
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) o[i][k] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    const int NT = S / BKV;
    for (int j = 0; j < NT; ++j) {
        cp_async_wait<0>();
        __syncthreads();
        if (j + 1 < NT) load_kv((j + 1) & 1, j + 1);
        cp_async_commit();
        if (j == 0) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const int row = warp * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                const int ch = ks * 2 + (lane >> 4);
                ldmatrix_x4(qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], sQ + swz64(row, ch));
                ldmatrix_x4(ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3], sQ + Q_PLANE + swz64(row, ch));
            }
        }
        const uint32_t sb = sStage + (j & 1) * STAGE;
        const float* bias_s = bias_s0 + (size_t)(j & 1) * (STAGE / 4);

        // ---- S = Q K^T (log2 domain) ------------------------------------------------------
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) s[i][k] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t kh[4], kl[4];
                const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int ch = ks * 2 + ((lane >> 3) & 1);
                const uint32_t off = swz64(row, ch);
                ldmatrix_x4(kh[0], kh[1], kh[2], kh[3], sb + off);
                ldmatrix_x4(kl[0], kl[1], kl[2], kl[3], sb + KV_PLANE + off);
                mma_f16(s[2 * np], ql[ks], kh[0], kh[1]);
                mma_f16(s[2 * np], qh[ks], kl[0], kl[1]);
                mma_f16(s[2 * np], qh[ks], kh[0], kh[1]);
                mma_f16(s[2 * np + 1], ql[ks], kh[2], kh[3]);
                mma_f16(s[2 * np + 1], qh[ks], kl[2], kl[3]);
                mma_f16(s[2 * np + 1], qh[ks], kh[2], kh[3]);
            }
        }
        // ---- + bias, online softmax ---------------------------------------------------------
        float mx0 = -INFINITY, mx1 = -INFINITY;
        {
            const float* br0 = bias_s + (warp * 16 + g) * BIAS_LD + 2 * t;
            const float* br1 = br0 + 8 * BIAS_LD;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float2 b0 = *reinterpret_cast<const float2*>(br0 + nt * 8);
                const float2 b1 = *reinterpret_cast<const float2*>(br1 + nt * 8);
                s[nt][0] += b0.x; s[nt][1] += b0.y; s[nt][2] += b1.x; s[nt][3] += b1.y;
                mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
                mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
            }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float c0 = ex2(m0 - mn0), c1 = ex2(m1 - mn1);    // first tile: ex2(-inf) = 0
        m0 = mn0; m1 = mn1;
        l0 *= c0; l1 *= c1;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1; }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = ex2(s[nt][0] - mn0); s[nt][1] = ex2(s[nt][1] - mn0);
            s[nt][2] = ex2(s[nt][2] - mn1); s[nt][3] = ex2(s[nt][3] - mn1);
            l0 += s[nt][0] + s[nt][1];
            l1 += s[nt][2] + s[nt][3];
        }
        // ---- O += P V ---------------------------------------------------------------------------
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {    // 16 keys per step
            uint32_t ph[4], pl[4];
            split2_unit(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
            split2_unit(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
            split2_unit(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
            split2_unit(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
            for (int dp = 0; dp < 2; ++dp) {
                uint32_t vh[4], vl[4];
                const int row = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                const int ch = dp * 2 + (lane >> 4);
                const uint32_t off = swz64(row, ch);
                ldmatrix_x4_trans(vh[0], vh[1], vh[2], vh[3], sb + 2 * KV_PLANE + off);
                ldmatrix_x4_trans(vl[0], vl[1], vl[2], vl[3], sb + 3 * KV_PLANE + off);
                mma_f16(o[2 * dp], pl, vh[0], vh[1]);
                mma_f16(o[2 * dp], ph, vl[0], vl[1]);
                mma_f16(o[2 * dp], ph, vh[0], vh[1]);
                mma_f16(o[2 * dp + 1], pl, vh[2], vh[3]);
                mma_f16(o[2 * dp + 1], ph, vl[2], vl[3]);
                mma_f16(o[2 * dp + 1], ph, vh[2], vh[3]);
            }
        }
    }
    cp_async_wait<0>();

    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const size_t r0 = (size_t)b * S + (size_t)qt * BQ + warp * 16 + g;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int col = h * D + nt * 8 + 2 * t;
        uint32_t hi, lo;
        split2(o[nt][0] * i0, o[nt][1] * i0, hi, lo);
        *reinterpret_cast<uint32_t*>(p.oh + r0 * p.c + col) = hi;
        *reinterpret_cast<uint32_t*>(p.ol + r0 * p.c + col) = lo;
        split2(o[nt][2] * i1, o[nt][3] * i1, hi, lo);
        *reinterpret_cast<uint32_t*>(p.oh + (r0 + 8) * p.c + col) = hi;
        *reinterpret_cast<uint32_t*>(p.ol + (r0 + 8) * p.c + col) = lo;
    }
}

}  // namespace

cudaError_t launch_attention(const AttnArgs& a, cudaStream_t st) {
    if (a.S_pad <= 0 || a.S_pad % BQ || a.c != a.H * D || a.B <= 0) return cudaErrorInvalidValue;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid(a.B, a.S_pad / BQ, a.H);
    attention_kernel<<<grid, 256, SMEM, st>>>(a);
    return cudaGetLastError();
}

}  // namespace pdk
