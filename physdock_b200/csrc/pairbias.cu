// Pair-bias prepass: the step- and sample-invariant half of DiTAttention, hoisted out of the loop.
//
// Reference (PhysDock/models/primitives/attentions.py:246,254-255) recomputes, in EVERY block of EVERY
// denoising step,   attn_bias = Linear_nobias(LayerNorm_affine(z)).permute(2,0,1) + gen_attn_mask(mask, -inf)
// although z (= ap [Na,Na,16] or z [Nt,Nt,128]) is constant for the whole sampling call.  Here one pass over
// the pair tensor normalises each pair vector once and emits the bias of ALL blocks of a stack
// ([L*H] outputs per pair, LayerNorm affine folded into the projection on the host), scaled by log2(e),
// mask folded in, padded to S_pad (pad key columns = kPadBias, pad query rows = 0).
#include "common.cuh"
#include "kernels.h"

namespace pdk {

namespace {

// ---- C = 16 (atom pairs): one thread per pair, HBM-streaming ---------------------------------
__global__ void __launch_bounds__(256) pair_bias_c16_kernel(const float* __restrict__ pair,
                                                            const float* __restrict__ mask,
                                                            const float* __restrict__ wfoldT,   // [16][LH]
                                                            const float* __restrict__ bfold,    // [LH]
                                                            float* __restrict__ bias, int S, int S_pad, int LH,
                                                            float ln_eps, float inf_) {
    extern __shared__ __align__(16) float sw[];     // [16*LH] weights, then [LH] offsets
    for (int i = threadIdx.x; i < 16 * LH; i += blockDim.x) sw[i] = wfoldT[i];
    for (int i = threadIdx.x; i < LH; i += blockDim.x) sw[16 * LH + i] = bfold[i];
    __syncthreads();
    const int i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= S_pad) return;
    const size_t plane = (size_t)S_pad * S_pad;
    float* dst = bias + (size_t)i * S_pad + j;
    if (j >= S || i >= S) {
        const float v = (j >= S) ? kPadBias : 0.f;
        for (int o = 0; o < LH; ++o) dst[o * plane] = v;
        return;
    }
    float x[16];
    const float4* src = reinterpret_cast<const float4*>(pair + ((size_t)i * S + j) * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 v = __ldg(src + q);
        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
    float mean = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) mean += x[c];
    mean *= (1.f / 16.f);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) { x[c] -= mean; var += x[c] * x[c]; }
    const float rstd = 1.0f / sqrtf(var * (1.f / 16.f) + ln_eps);
#pragma unroll
    for (int c = 0; c < 16; ++c) x[c] *= rstd;
    const float mterm = (mask[(size_t)i * S + j] == 0.f) ? -inf_ : 0.f;
    for (int o = 0; o < LH; ++o) {
        float acc = sw[16 * LH + o];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc = fmaf(x[c], sw[c * LH + o], acc);
        dst[o * plane] = (acc + mterm) * kLog2e;
    }
}

// ---- C = 128 (token pairs): CTA = one query row i x 32 keys; [32 x 128] x [128 x LH] in smem ------
__global__ void __launch_bounds__(256) pair_bias_c128_kernel(const float* __restrict__ pair,
                                                             const float* __restrict__ mask,
                                                             const float* __restrict__ wfoldT,  // [128][LH]
                                                             const float* __restrict__ bfold,
                                                             float* __restrict__ bias, int S, int S_pad, int LH,
                                                             float ln_eps, float inf_) {
    extern __shared__ __align__(16) float sm[];
    float* sw = sm;                      // [128][LH]
    float* sx = sm + 128 * LH;           // [128][33]  normalised pair vectors, channel-major
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int i = blockIdx.y, j0 = blockIdx.x * 32;
    const size_t plane = (size_t)S_pad * S_pad;
    if (i >= S || j0 >= S) {             // whole tile is padding
        const int j = j0 + lane;
        for (int o = warp; o < LH; o += 8) bias[o * plane + (size_t)i * S_pad + j] = (j >= S) ? kPadBias : 0.f;
        return;
    }
    for (int k = tid; k < 128 * LH; k += 256) sw[k] = wfoldT[k];
    // phase 1: each warp normalises 4 pair vectors (lane = 4 channels)
    for (int q = 0; q < 4; ++q) {
        const int jj = warp * 4 + q, j = j0 + jj;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < S) v = __ldg(reinterpret_cast<const float4*>(pair + ((size_t)i * S + j) * 128) + lane);
        const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.f / 128.f);
        v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
        const float var = warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w) * (1.f / 128.f);
        const float rstd = 1.0f / sqrtf(var + ln_eps);
        sx[(4 * lane + 0) * 33 + jj] = v.x * rstd;
        sx[(4 * lane + 1) * 33 + jj] = v.y * rstd;
        sx[(4 * lane + 2) * 33 + jj] = v.z * rstd;
        sx[(4 * lane + 3) * 33 + jj] = v.w * rstd;
    }
    __syncthreads();
    // phase 2: thread = (pair lane, 4 consecutive outputs); weights are warp-broadcast float4 reads
    const int j = j0 + lane;
    const float mterm = (j < S && mask[(size_t)i * S + j] == 0.f) ? -inf_ : 0.f;
    for (int o4 = warp * 4; o4 < LH; o4 += 32) {
        float a0 = bfold[o4], a1 = bfold[o4 + 1], a2 = bfold[o4 + 2], a3 = bfold[o4 + 3];
#pragma unroll 8
        for (int c = 0; c < 128; ++c) {
            const float xv = sx[c * 33 + lane];
            const float4 w = *reinterpret_cast<const float4*>(sw + c * LH + o4);
            a0 = fmaf(xv, w.x, a0); a1 = fmaf(xv, w.y, a1); a2 = fmaf(xv, w.z, a2); a3 = fmaf(xv, w.w, a3);
        }
        float* dst = bias + (size_t)o4 * plane + (size_t)i * S_pad + j;
        if (j < S) {
            dst[0] = (a0 + mterm) * kLog2e;
            dst[plane] = (a1 + mterm) * kLog2e;
            dst[2 * plane] = (a2 + mterm) * kLog2e;
            dst[3 * plane] = (a3 + mterm) * kLog2e;
        } else {
            dst[0] = kPadBias; dst[plane] = kPadBias; dst[2 * plane] = kPadBias; dst[3 * plane] = kPadBias;
        }
    }
}

}  // namespace

cudaError_t launch_pair_bias(const float* pair, const float* mask, const float* wfoldT, const float* bfold,
                             float* bias, int S, int S_pad, int C, int LH, float ln_eps, float inf_,
                             cudaStream_t st) {
    if (S <= 0 || S_pad < S || S_pad % 128 || LH <= 0 || LH % 4) return cudaErrorInvalidValue;
    if (C == 16) {
        dim3 grid((S_pad + 255) / 256, S_pad);
        const size_t smem = (size_t)(16 * LH + LH) * sizeof(float);
        pair_bias_c16_kernel<<<grid, 256, smem, st>>>(pair, mask, wfoldT, bfold, bias, S, S_pad, LH, ln_eps, inf_);
    } else if (C == 128) {
        const size_t smem = (size_t)(128 * LH + 128 * 33) * sizeof(float);
        if (smem > 220 * 1024) return cudaErrorInvalidValue;
        cudaError_t e = cudaFuncSetAttribute(pair_bias_c128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
        dim3 grid(S_pad / 32, S_pad);
        pair_bias_c128_kernel<<<grid, 256, smem, st>>>(pair, mask, wfoldT, bfold, bias, S, S_pad, LH, ln_eps, inf_);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace pdk
