// Split-fp16 GEMM with fused epilogues (v1: cp.async + ldmatrix + mma.sync, fp32 accumulate).
//
// Replaces every `Linear.forward -> F.linear` call site on the hot path (reference
// PhysDock/models/primitives/linear.py:146-161): q/k/v projections (attentions.py:248-250), out
// projection (:263), SwiGLU w1/w3/w2 (feed_forward.py:30-31), linear_downscale/upscale
// (layers/transformers.py:206,215).
//
// CTA tile 128x128x32, 8 warps as 2(M) x 4(N), warp tile 64x32, 3-stage cp.async ring (96 KB, 2 CTAs/SM).
// Every warp's 32 output columns are one attention head / four SwiGLU column blocks, which is what lets
// the per-head RMSNorm and the SwiGLU product run on the accumulator fragments without a round trip.
#include "common.cuh"
#include "kernels.h"

namespace pdk {

namespace {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int PLANE_BYTES = BM * BK * 2;           // 8 KB per fp16 plane tile (64-byte rows)
constexpr int STAGE_BYTES = 4 * PLANE_BYTES;       // A_hi, A_lo, W_hi, W_lo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES;   // 96 KB

template <int EPI>
__global__ void __launch_bounds__(256, 2) gemm_split_kernel(const GemmArgs p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int KT = p.K / BK;
    const uint32_t sbase = smem_u32(smem);

    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

    auto load_stage = [&](int stage, int kt) {
        const uint32_t sb = sbase + stage * STAGE_BYTES;
        const int k0 = kt * BK;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int q = tid + i * 256;
            const int row = q >> 2, ch = q & 3;
            const uint32_t off = swz64(row, ch);
            const size_t ga = (size_t)(m0 + row) * p.lda + k0 + ch * 8;
            const size_t gw = (size_t)(n0 + row) * p.ldw + k0 + ch * 8;
            cp_async16(sb + off, p.Ah + ga);
            cp_async16(sb + PLANE_BYTES + off, p.Al + ga);
            cp_async16(sb + 2 * PLANE_BYTES + off, p.Wh + gw);
            cp_async16(sb + 3 * PLANE_BYTES + off, p.Wl + gw);
        }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < KT) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const uint32_t sb = sbase + (kt % STAGES) * STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t bh[4][2], bl[4][2];
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                const int row = wn * 32 + np * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int ch = ks * 2 + ((lane >> 3) & 1);
                const uint32_t off = swz64(row, ch);
                ldmatrix_x4(bh[2 * np][0], bh[2 * np][1], bh[2 * np + 1][0], bh[2 * np + 1][1],
                            sb + 2 * PLANE_BYTES + off);
                ldmatrix_x4(bl[2 * np][0], bl[2 * np][1], bl[2 * np + 1][0], bl[2 * np + 1][1],
                            sb + 3 * PLANE_BYTES + off);
            }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                uint32_t ah[4], al[4];
                const int row = wm * 64 + mt * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                const int ch = ks * 2 + (lane >> 4);
                const uint32_t off = swz64(row, ch);
                ldmatrix_x4(ah[0], ah[1], ah[2], ah[3], sb + off);
                ldmatrix_x4(al[0], al[1], al[2], al[3], sb + PLANE_BYTES + off);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    mma_f16(acc[mt][nt], al, bh[nt][0], bh[nt][1]);
                    mma_f16(acc[mt][nt], ah, bl[nt][0], bl[nt][1]);
                    mma_f16(acc[mt][nt], ah, bh[nt][0], bh[nt][1]);
                }
            }
        }
    }
    cp_async_wait<0>();

    // ------------------------------------------------------------------------------ epilogues
    const int n_base = n0 + wn * 32;
    if constexpr (EPI == EPI_STORE) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int r0 = m0 + wm * 64 + mt * 16 + g;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int col = n_base + nt * 8 + 2 * t;
                float2 b = make_float2(0.f, 0.f);
                if (p.bias) b = *reinterpret_cast<const float2*>(p.bias + col);
                float v0 = acc[mt][nt][0] + b.x, v1 = acc[mt][nt][1] + b.y;
                float v2 = acc[mt][nt][2] + b.x, v3 = acc[mt][nt][3] + b.y;
                if (p.act_silu) { v0 = silu(v0); v1 = silu(v1); v2 = silu(v2); v3 = silu(v3); }
                *reinterpret_cast<float2*>(p.out + (size_t)r0 * p.ldo + col) = make_float2(v0, v1);
                *reinterpret_cast<float2*>(p.out + (size_t)(r0 + 8) * p.ldo + col) = make_float2(v2, v3);
            }
        }
    } else if constexpr (EPI == EPI_GATE_RESID) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int r0 = m0 + wm * 64 + mt * 16 + g;
            const int r1 = r0 + 8;
            const float* g0 = p.gate + (size_t)(r0 / p.rows_per_sample) * p.gate_stride;
            const float* g1 = p.gate + (size_t)(r1 / p.rows_per_sample) * p.gate_stride;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int col = n_base + nt * 8 + 2 * t;
                float2 b = make_float2(0.f, 0.f);
                if (p.bias) b = *reinterpret_cast<const float2*>(p.bias + col);
                const float2 ga = *reinterpret_cast<const float2*>(g0 + col);
                const float2 gb = *reinterpret_cast<const float2*>(g1 + col);
                float2* x0 = reinterpret_cast<float2*>(p.out + (size_t)r0 * p.ldo + col);
                float2* x1 = reinterpret_cast<float2*>(p.out + (size_t)r1 * p.ldo + col);
                float2 a0 = *x0, a1 = *x1;
                a0.x += (acc[mt][nt][0] + b.x) * ga.x;
                a0.y += (acc[mt][nt][1] + b.y) * ga.y;
                a1.x += (acc[mt][nt][2] + b.x) * gb.x;
                a1.y += (acc[mt][nt][3] + b.y) * gb.y;
                *x0 = a0;
                *x1 = a1;
            }
        }
    } else if constexpr (EPI == EPI_SWIGLU) {
        // W rows come in blocks of 8: even block = w1 rows, odd block = w3 rows of the same hidden columns
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int r0 = m0 + wm * 64 + mt * 16 + g;
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                const int j = ((n_base >> 3) + 2 * np) / 2 * 8 + 2 * t;   // hidden column
                const float* h1 = acc[mt][2 * np];
                const float* h3 = acc[mt][2 * np + 1];
                uint32_t hi, lo;
                split2(silu(h1[0]) * h3[0], silu(h1[1]) * h3[1], hi, lo);
                *reinterpret_cast<uint32_t*>(p.ph + (size_t)r0 * p.ldp + j) = hi;
                *reinterpret_cast<uint32_t*>(p.pl + (size_t)r0 * p.ldp + j) = lo;
                split2(silu(h1[2]) * h3[2], silu(h1[3]) * h3[3], hi, lo);
                *reinterpret_cast<uint32_t*>(p.ph + (size_t)(r0 + 8) * p.ldp + j) = hi;
                *reinterpret_cast<uint32_t*>(p.pl + (size_t)(r0 + 8) * p.ldp + j) = lo;
            }
        }
    } else {   // EPI_QKV
        const int which = n_base / p.c;                 // 0 q, 1 k, 2 v   (warp-uniform)
        const int head = (n_base % p.c) / kHeadDim;
        const int H = p.c / kHeadDim;
        __half* dh = which == 0 ? p.qh : (which == 1 ? p.kh : p.vh);
        __half* dl = which == 0 ? p.ql : (which == 1 ? p.kl : p.vl);
        const float* gain = which == 0 ? p.norm_q : p.norm_k;
        const float post = which == 0 ? p.q_scale : 1.0f;
        float w[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            w[nt][0] = which < 2 ? gain[nt * 8 + 2 * t] * post : 1.f;
            w[nt][1] = which < 2 ? gain[nt * 8 + 2 * t + 1] * post : 1.f;
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            float inv0 = 1.f, inv1 = 1.f;
            if (which < 2) {    // RMSNorm over the head's 32 channels (rms_norm.py:14-19)
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    s0 += acc[mt][nt][0] * acc[mt][nt][0] + acc[mt][nt][1] * acc[mt][nt][1];
                    s1 += acc[mt][nt][2] * acc[mt][nt][2] + acc[mt][nt][3] * acc[mt][nt][3];
                }
                s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
                s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
                s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
                s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
                inv0 = 1.0f / sqrtf(s0 * (1.0f / kHeadDim) + p.rms_eps);
                inv1 = 1.0f / sqrtf(s1 * (1.0f / kHeadDim) + p.rms_eps);
            }
            const int r0 = m0 + wm * 64 + mt * 16 + g;
            const int r1 = r0 + 8;
            const size_t d0 = ((size_t)((r0 / p.rows_per_sample) * H + head) * p.rows_per_sample +
                               (r0 % p.rows_per_sample)) * kHeadDim;
            const size_t d1 = ((size_t)((r1 / p.rows_per_sample) * H + head) * p.rows_per_sample +
                               (r1 % p.rows_per_sample)) * kHeadDim;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int d = nt * 8 + 2 * t;
                uint32_t hi, lo;
                split2(acc[mt][nt][0] * inv0 * w[nt][0], acc[mt][nt][1] * inv0 * w[nt][1], hi, lo);
                *reinterpret_cast<uint32_t*>(dh + d0 + d) = hi;
                *reinterpret_cast<uint32_t*>(dl + d0 + d) = lo;
                split2(acc[mt][nt][2] * inv1 * w[nt][0], acc[mt][nt][3] * inv1 * w[nt][1], hi, lo);
                *reinterpret_cast<uint32_t*>(dh + d1 + d) = hi;
                *reinterpret_cast<uint32_t*>(dl + d1 + d) = lo;
            }
        }
    }
}

template <int EPI>
cudaError_t launch_one(const GemmArgs& a, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_split_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid(a.N / BN, a.M / BM);
    gemm_split_kernel<EPI><<<grid, 256, SMEM_BYTES, st>>>(a);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_gemm(GemmEpilogue epi, const GemmArgs& a, cudaStream_t st) {
    if (a.M <= 0 || a.N <= 0 || a.K <= 0 || a.M % BM || a.N % BN || a.K % BK) return cudaErrorInvalidValue;
    if (a.lda % 8 || a.ldw % 8) return cudaErrorInvalidValue;   // 16-byte cp.async source alignment
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
