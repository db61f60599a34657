// Conditioning and glue kernels of the denoiser: everything in AF3DiT.forward that is not a big GEMM or
// the attention core (reference PhysDock/models/layers/transformers.py:205-233 and the AdaLN-Zero path,
// primitives/adaptive_layer_norm_zero.py:18-21).  All are HBM/latency bound: coalesced, vectorised,
// warp-shuffle reductions, fp32 arithmetic in the reference's operation order.
#include "common.cuh"
#include "kernels.h"

namespace pdk {

namespace {

// ------------------------------------------------------------------------------------------------
// precond scalars + TimestepEmbeddings (transformers.py:219-224, timestep_embeddings.py:35-86,127-166).
// One CTA per sample.  The sinusoid argument t_hat*c_noise reaches ~6.5e3 rad, where a 1-ulp change of the
// fp32 log moves the phase by ~4e-4 rad; the log is therefore evaluated in fp64 and rounded once (a correctly
// rounded fp32 log; the oracle's CPU libm is within 1 ulp of that).  sin/cos use the accurate fp32 sincosf
// (full-range argument reduction, <= 2 ulp): their fp64 versions cost ~45 us per step for no measurable gain.
__global__ void __launch_bounds__(256) time_embed_kernel(const float* __restrict__ t_hat,
                                                         const float* __restrict__ freq,
                                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                                         float sigma_data, float* __restrict__ tsilu,
                                                         __half* __restrict__ ts_h, __half* __restrict__ ts_l,
                                                         float* __restrict__ coef, int coef_ld, int B) {
    griddep_launch();
    griddep_wait();
    __shared__ __align__(16) float proj[kTimeDim];
    __shared__ __align__(16) float hid[kTimeDim];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (b >= B) {      // zero padding rows of the [128 x 256] operand planes of the modulation GEMM
        if (ts_h != nullptr) { ts_h[(size_t)b * kTimeDim + tid] = __float2half(0.f); ts_l[(size_t)b * kTimeDim + tid] = __float2half(0.f); }
        return;
    }
    const float t = t_hat[b];
    const float sd2 = sigma_data * sigma_data;
    const float t2 = t * t;
    if (tid == 0) {
        float* cf = coef + (size_t)coef_ld * b;
        cf[0] = 1.0f / sqrtf(t2 + sd2);                      // c_in   (:219)
        cf[1] = sd2 / (sd2 + t2);                            // c_skip (:229)
        cf[2] = sigma_data * t / sqrtf(sd2 + t2);            // c_out  (:230)
        cf[3] = t;
    }
    const float c_noise = (float)log((double)(t / sigma_data)) / 4.0f;  // (:220)
    const float t_in = t * c_noise;                                     // (:223) sic
    if (tid < 128) {
        const float arg = t_in * freq[tid];
        float sn, cs;
        sincosf(arg, &sn, &cs);
        proj[tid] = cs;                               // flip_sin_to_cos=True -> [cos | sin]
        proj[128 + tid] = sn;
    }
    __syncthreads();
    // thread = one output feature; 64 independent float4 loads per layer keep the (L2-resident) weight rows streaming
    auto matvec = [&](const float* __restrict__ w, const float* x) {
        const float4* row = reinterpret_cast<const float4*>(w + (size_t)tid * kTimeDim);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 16
        for (int k = 0; k < kTimeDim / 4; ++k) {
            const float4 wv = __ldg(row + k);
            const float4 xv = reinterpret_cast<const float4*>(x)[k];
            a0 = fmaf(wv.x, xv.x, a0); a1 = fmaf(wv.y, xv.y, a1); a2 = fmaf(wv.z, xv.z, a2); a3 = fmaf(wv.w, xv.w, a3);
        }
        return (a0 + a1) + (a2 + a3);
    };
    hid[tid] = silu(matvec(w1, proj) + b1[tid]);      // linear_1 + SiLU
    __syncthreads();
    const float out = silu(matvec(w2, hid) + b2[tid]);   // linear_2, then the SiLU every AdaLN applies first
    if (tsilu != nullptr) tsilu[(size_t)b * kTimeDim + tid] = out;
    if (ts_h != nullptr) {
        const __half h = __float2half_rn(out);
        ts_h[(size_t)b * kTimeDim + tid] = h;
        ts_l[(size_t)b * kTimeDim + tid] = __float2half_rn(out - __half2float(h));
    }
}

// ------------------------------------------------------------------------------------------------
// mod[b, n] = sum_k tsilu[b,k] * wmod[n,k] + bmod[n]: the 36 AdaLN-Zero `Linear(256 -> 3c)` of the model in
// one weight-streaming pass (adaptive_layer_norm_zero.py:19).  Warp per output column, samples in registers.
constexpr int MOD_BCHUNK = 16;
constexpr int MOD_COLS = 64;       // output columns per CTA (8 per warp): amortises staging tsilu in shared memory
__global__ void __launch_bounds__(256) mod_gemv_kernel(const float* __restrict__ tsilu,
                                                       const float* __restrict__ wmod,
                                                       const float* __restrict__ bmod, float* __restrict__ mod,
                                                       int B, int Nmod) {
    griddep_launch();
    griddep_wait();
    __shared__ float ts[MOD_BCHUNK][kTimeDim];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * MOD_COLS + warp * (MOD_COLS / 8);
    for (int b0 = 0; b0 < B; b0 += MOD_BCHUNK) {
        const int nb = min(MOD_BCHUNK, B - b0);
        __syncthreads();
        for (int i = tid; i < nb * kTimeDim; i += 256) ts[i / kTimeDim][i % kTimeDim] = tsilu[(size_t)b0 * kTimeDim + i];
        __syncthreads();
#pragma unroll 2
        for (int c = 0; c < MOD_COLS / 8; ++c) {
            const int n = n0 + c;
            if (n >= Nmod) break;
            float w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = __ldg(wmod + (size_t)n * kTimeDim + lane + 32 * i);
            const float bias = bmod[n];
            for (int bb = 0; bb < nb; ++bb) {
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc = fmaf(ts[bb][lane + 32 * i], w[i], acc);
                acc = warp_sum(acc);
                if (lane == 0) mod[(size_t)(b0 + bb) * Nmod + n] = acc + bias;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// AdaLN-Zero modulate: planes = split( LN_noaffine(x) * (1 + scale) + shift )   (adaptive_layer_norm_zero.py:20)
// Warp per row; C/32 values per lane held in registers (C = 128 or 512).
// The first AdaLN of each atom stack also PRODUCES its input row (SRC != 0, C = 128), which removes one launch and one
// 16.8 MB read per stack:
//   SRC_PRECOND  x = Linear_{3->c_a}(x_hat * c_in) + a, pad rows 0        (AF3DiT.precond, transformers.py:222)
//   SRC_UPSCALE  x += up[b, atom_id_to_token_id[s], :] for s < Na          (AF3DiT.upscale, transformers.py:214-216)
// with the same arithmetic as the stand-alone precond / gather_add kernels (bit-identical); the row is stored back to x.
enum { SRC_X = 0, SRC_PRECOND = 1, SRC_UPSCALE = 2 };
struct AdalnSrc {
    const float* x_hat; const float* coef; int coef_stride; const float* a; const float* wx; const float* bx;   // SRC_PRECOND
    const float* up; const int* atom2tok; int St_pad;                                                           // SRC_UPSCALE
    int Na;
};
template <int C, int SRC>
__global__ void __launch_bounds__(256) adaln_kernel(float* __restrict__ x, const float* __restrict__ mod,
                                                    int mod_stride, int mod_off, __half* __restrict__ xh,
                                                    __half* __restrict__ xl, int rows, int S_pad, float eps, AdalnSrc src) {
    griddep_launch();
    griddep_wait();
    constexpr int V = C / 128;     // float4 per lane
    static_assert(SRC == SRC_X || C == 128, "fused sources exist for the atom width only");
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float4* xrow = reinterpret_cast<float4*>(x + (size_t)row * C);
    float4 v[V];
    if constexpr (SRC == SRC_PRECOND) {
        const int b = row / S_pad, s = row - b * S_pad;
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < src.Na) {
            const float c_in = src.coef[src.coef_stride * b];
            const float* xp = src.x_hat + ((size_t)b * src.Na + s) * 3;
            const float x0 = xp[0] * c_in, x1 = xp[1] * c_in, x2 = xp[2] * c_in;
            const float4 av = *reinterpret_cast<const float4*>(src.a + (size_t)s * C + 4 * lane);
            const float4 bv = *reinterpret_cast<const float4*>(src.bx + 4 * lane);
            // the 12 weights of this lane's four channels as three 16-byte loads (twelve scalar loads with a 48-byte lane stride
            // cost 144 L1 wavefronts per row: 13.4 us for this kernel against 8.5 us for the upscale variant, which moves more)
            const float4* w4 = reinterpret_cast<const float4*>(src.wx + (size_t)(4 * lane) * 3);
            const float4 wa = __ldg(w4), wb = __ldg(w4 + 1), wc = __ldg(w4 + 2);
            out.x = fmaf(x2, wa.z, fmaf(x1, wa.y, x0 * wa.x)) + bv.x + av.x;
            out.y = fmaf(x2, wb.y, fmaf(x1, wb.x, x0 * wa.w)) + bv.y + av.y;
            out.z = fmaf(x2, wc.x, fmaf(x1, wb.w, x0 * wb.z)) + bv.z + av.z;
            out.w = fmaf(x2, wc.w, fmaf(x1, wc.z, x0 * wc.y)) + bv.w + av.w;
        }
        v[0] = out;
        xrow[lane] = out;
    } else if constexpr (SRC == SRC_UPSCALE) {
        const int b = row / S_pad, s = row - b * S_pad;
        float4 cur = xrow[lane];
        if (s < src.Na) {
            const int tok = src.atom2tok[s];
            const float4 u = *(reinterpret_cast<const float4*>(src.up + ((size_t)b * src.St_pad + tok) * C) + lane);
            cur.x += u.x; cur.y += u.y; cur.z += u.z; cur.w += u.w;
            xrow[lane] = cur;
        }
        v[0] = cur;
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = xrow[lane + 32 * i];
    }
    // explicit rounding (no compiler-chosen FMA contraction): every instantiation of this kernel produces the same bits
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) sum = __fadd_rn(sum, __fadd_rn(__fadd_rn(v[i].x, v[i].y), __fadd_rn(v[i].z, v[i].w)));
    const float mean = __fmul_rn(warp_sum(sum), 1.f / C);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        v[i].x = __fsub_rn(v[i].x, mean); v[i].y = __fsub_rn(v[i].y, mean);
        v[i].z = __fsub_rn(v[i].z, mean); v[i].w = __fsub_rn(v[i].w, mean);
        sq = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, fmaf(v[i].w, v[i].w, sq))));
    }
    const float rstd = inv_sqrt(__fadd_rn(__fmul_rn(warp_sum(sq), 1.f / C), eps));
    const float* shift = mod + (size_t)(row / S_pad) * mod_stride + mod_off;
    const float* scale = shift + C;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int c0 = 4 * (lane + 32 * i);
        const float4 sh = *reinterpret_cast<const float4*>(shift + c0);
        const float4 sc = *reinterpret_cast<const float4*>(scale + c0);
        const float y0 = fmaf(__fmul_rn(v[i].x, rstd), __fadd_rn(1.f, sc.x), sh.x);
        const float y1 = fmaf(__fmul_rn(v[i].y, rstd), __fadd_rn(1.f, sc.y), sh.y);
        const float y2 = fmaf(__fmul_rn(v[i].z, rstd), __fadd_rn(1.f, sc.z), sh.z);
        const float y3 = fmaf(__fmul_rn(v[i].w, rstd), __fadd_rn(1.f, sc.w), sh.w);
        uint2 hi, lo;
        split2(y0, y1, hi.x, lo.x);
        split2(y2, y3, hi.y, lo.y);
        *reinterpret_cast<uint2*>(xh + (size_t)row * C + c0) = hi;
        *reinterpret_cast<uint2*>(xl + (size_t)row * C + c0) = lo;
    }
}

__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ x, __half* __restrict__ xh,
                                                    __half* __restrict__ xl, size_t n4) {
    griddep_launch();
    griddep_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        uint2 hi, lo;
        split2(v.x, v.y, hi.x, lo.x);
        split2(v.z, v.w, hi.y, lo.y);
        reinterpret_cast<uint2*>(xh)[i] = hi;
        reinterpret_cast<uint2*>(xl)[i] = lo;
    }
}

// ba = Linear_{3->c_a}(x_hat * c_in) + a   (transformers.py:222); pad rows zeroed.
__global__ void __launch_bounds__(256) precond_kernel(const float* __restrict__ x_hat, const float* __restrict__ coef,
                                                      const float* __restrict__ a, const float* __restrict__ wx,
                                                      const float* __restrict__ bx, float* __restrict__ ba, int B,
                                                      int Na, int S_pad, int c_a, int coef_stride) {
    griddep_launch();
    griddep_wait();
    const int per_row = c_a / 4;
    const uint32_t total = (uint32_t)B * S_pad * per_row;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {      // 32-bit: launch checks total < 2^31
        const int c4 = (int)(i % (uint32_t)per_row);
        const uint32_t r = i / (uint32_t)per_row;
        const int s = (int)(r % (uint32_t)S_pad), b = (int)(r / (uint32_t)S_pad);
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < Na) {
            const float c_in = coef[coef_stride * b];
            const float* xp = x_hat + ((size_t)b * Na + s) * 3;
            const float x0 = xp[0] * c_in, x1 = xp[1] * c_in, x2 = xp[2] * c_in;
            const float4 av = *reinterpret_cast<const float4*>(a + (size_t)s * c_a + 4 * c4);
            const float4 bv = *reinterpret_cast<const float4*>(bx + 4 * c4);
            const float* w = wx + (size_t)(4 * c4) * 3;
            out.x = fmaf(x2, w[2], fmaf(x1, w[1], x0 * w[0])) + bv.x + av.x;
            out.y = fmaf(x2, w[5], fmaf(x1, w[4], x0 * w[3])) + bv.y + av.y;
            out.z = fmaf(x2, w[8], fmaf(x1, w[7], x0 * w[6])) + bv.z + av.z;
            out.w = fmaf(x2, w[11], fmaf(x1, w[10], x0 * w[9])) + bv.w + av.w;
        }
        reinterpret_cast<float4*>(ba)[i] = out;
    }
}

// Token pooling of AF3DiT.downscale (transformers.py:207-212).  The reference forms per-token sums as a
// difference of an fp32 cumsum over all atoms; summing each contiguous chunk directly is the same quantity
// without the cancellation (documented deviation, ~1e-6 relative).
__global__ void __launch_bounds__(256) segment_mean_kernel(const float* __restrict__ h, const int* __restrict__ tok_start,
                                                           const float* __restrict__ s, float* __restrict__ bs, int B,
                                                           int Nt, int Sa_pad, int St_pad, int c_s) {
    griddep_launch();
    griddep_wait();
    const int per_row = c_s / 4;
    const uint32_t total = (uint32_t)B * St_pad * per_row;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {      // 32-bit: launch checks total < 2^31
        const int c4 = (int)(i % (uint32_t)per_row);
        const uint32_t r = i / (uint32_t)per_row;
        const int tok = (int)(r % (uint32_t)St_pad), b = (int)(r / (uint32_t)St_pad);
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tok < Nt) {
            const int a0 = tok_start[tok], a1 = tok_start[tok + 1];
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4* src = reinterpret_cast<const float4*>(h + ((size_t)b * Sa_pad) * c_s) + c4;
            for (int at = a0; at < a1; ++at) {
                const float4 v = src[(size_t)at * per_row];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            const float den = (float)(a1 - a0) + 1e-3f;
            const float4 sv = *reinterpret_cast<const float4*>(s + (size_t)tok * c_s + 4 * c4);
            out.x = acc.x / den + sv.x; out.y = acc.y / den + sv.y;
            out.z = acc.z / den + sv.z; out.w = acc.w / den + sv.w;
        }
        reinterpret_cast<float4*>(bs)[i] = out;
    }
}

// AF3DiT.upscale (transformers.py:214-216): ba[b,i,:] += up[b, atom_id_to_token_id[i], :]
__global__ void __launch_bounds__(256) gather_add_kernel(float* __restrict__ ba, const float* __restrict__ up,
                                                         const int* __restrict__ atom2tok, int B, int Na, int Sa_pad,
                                                         int St_pad, int c_a) {
    griddep_launch();
    griddep_wait();
    const int per_row = c_a / 4;
    const uint32_t total = (uint32_t)B * Na * per_row;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {      // 32-bit: launch checks total < 2^31
        const int c4 = (int)(i % (uint32_t)per_row);
        const uint32_t r = i / (uint32_t)per_row;
        const int at = (int)(r % (uint32_t)Na), b = (int)(r / (uint32_t)Na);
        const int tok = atom2tok[at];
        float4* dst = reinterpret_cast<float4*>(ba + ((size_t)b * Sa_pad + at) * c_a) + c4;
        const float4 u = *(reinterpret_cast<const float4*>(up + ((size_t)b * St_pad + tok) * c_a) + c4);
        float4 v = *dst;
        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        *dst = v;
    }
}

// AF3DiT.denoise (transformers.py:228-233): x_den = c_skip x_hat + c_out Linear_{c_a->3}(LN_affine(ba)).
// A warp takes EIGHT atom rows at a time (c_a = 128: one float4 per lane and row): the eight loads are in flight together and the
// five row reductions (mean, variance, three output dots) go through the multi-value butterfly, 85 shuffles per 8 rows instead of
// 200 -- one row per warp with five dependent butterflies was pure latency (12.7 us for 16.8 MB).
constexpr int kDenoiseRows = 8;
__global__ void __launch_bounds__(256) denoise_out_kernel(const float* __restrict__ ba, const float* __restrict__ x_hat,
                                                          const float* __restrict__ coef, const float* __restrict__ ln_w,
                                                          const float* __restrict__ ln_b, const float* __restrict__ wr,
                                                          float* __restrict__ x_den, int B, int Na, int S_pad,
                                                          float eps, int coef_stride, float* __restrict__ x_next) {
    griddep_launch();
    griddep_wait();
    const int lane = threadIdx.x & 31;
    const int rows = B * Na;
    const int r0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * kDenoiseRows;
    if (r0 >= rows) return;
    float4 v[kDenoiseRows];
    float red[kDenoiseRows];
#pragma unroll
    for (int i = 0; i < kDenoiseRows; ++i) {
        const int r = min(r0 + i, rows - 1);              // the tail warp recomputes the last row (never stored twice)
        const int b = r / Na, s = r - b * Na;
        v[i] = reinterpret_cast<const float4*>(ba + ((size_t)b * S_pad + s) * 128)[lane];
        red[i] = (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    warp_sum8(red);
#pragma unroll
    for (int i = 0; i < kDenoiseRows; ++i) {
        const float mean = red[i] * (1.f / 128.f);
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        red[i] = fmaf(v[i].x, v[i].x, v[i].y * v[i].y) + fmaf(v[i].z, v[i].z, v[i].w * v[i].w);
    }
    warp_sum8(red);
    const float4 g = reinterpret_cast<const float4*>(ln_w)[lane];
    const float4 be = reinterpret_cast<const float4*>(ln_b)[lane];
#pragma unroll
    for (int i = 0; i < kDenoiseRows; ++i) {
        const float rstd = inv_sqrt(red[i] * (1.f / 128.f) + eps);
        v[i].x = v[i].x * rstd * g.x + be.x; v[i].y = v[i].y * rstd * g.y + be.y;
        v[i].z = v[i].z * rstd * g.z + be.z; v[i].w = v[i].w * rstd * g.w + be.w;
    }
    float out[3][kDenoiseRows];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float4 w = reinterpret_cast<const float4*>(wr + j * 128)[lane];
#pragma unroll
        for (int i = 0; i < kDenoiseRows; ++i) out[j][i] = fmaf(v[i].w, w.w, fmaf(v[i].z, w.z, fmaf(v[i].y, w.y, v[i].x * w.x)));
        warp_sum8(out[j]);
    }
    // lane = 3 * row + component writes one coordinate (24 lanes active)
    const int i = lane / 3, j = lane - 3 * i;
    if (lane < 3 * kDenoiseRows && r0 + i < rows) {
        float r_j = 0.f;
#pragma unroll
        for (int jj = 0; jj < 3; ++jj)
#pragma unroll
            for (int ii = 0; ii < kDenoiseRows; ++ii)
                if (jj == j && ii == i) r_j = out[jj][ii];
        const int r = r0 + i;
        const int b = r / Na;
        const float* cf = coef + (size_t)coef_stride * b;
        const float c_skip = cf[1], c_out = cf[2];
        const size_t o = (size_t)r * 3 + j;
        const float xh = x_hat[o];
        const float xd = __fadd_rn(__fmul_rn(c_skip, xh), __fmul_rn(c_out, r_j));
        x_den[o] = xd;
        if (x_next != nullptr) {      // fused Euler update without physics (model.py:263-264,278-281), reference operation order
            const float th = cf[3], t_next = cf[4], eta = cf[5];      // per-step scalars live in the conditioning row
            const float d = __fsub_rn(xh, xd) / th;
            x_next[o] = __fadd_rn(xh, __fmul_rn(__fmul_rn(eta, __fsub_rn(t_next, th)), d));
        }
    }
}

inline int grid_for(size_t n, int block = 256, int cap = 148 * 16) {
    size_t g = (n + block - 1) / block;
    return (int)(g < (size_t)cap ? (g ? g : 1) : cap);
}

}  // namespace

cudaError_t launch_time_embed(const float* t_hat, const float* freq, const float* w1, const float* b1,
                              const float* w2, const float* b2, float sigma_data, float* tsilu, __half* ts_h,
                              __half* ts_l, int rows_padded, float* coef, int coef_ld, int B, cudaStream_t st) {
    if (B <= 0 || coef_ld < 4 || (ts_h != nullptr && rows_padded < B)) return cudaErrorInvalidValue;
    PDK_LAUNCH_CHECK(launch_pdl(time_embed_kernel, dim3(ts_h != nullptr ? rows_padded : B), dim3(256), (size_t)(0), st, t_hat, freq, w1, b1, w2, b2, sigma_data, tsilu, ts_h,
                                                                       ts_l, coef, coef_ld, B));
    return cudaGetLastError();
}

cudaError_t launch_mod_gemv(const float* tsilu, const float* wmod, const float* bmod, float* mod, int B,
                            int Nmod, cudaStream_t st) {
    if (B <= 0 || Nmod <= 0) return cudaErrorInvalidValue;
    PDK_LAUNCH_CHECK(launch_pdl(mod_gemv_kernel, dim3((Nmod + MOD_COLS - 1) / MOD_COLS), dim3(256), (size_t)(0), st, tsilu, wmod, bmod, mod, B, Nmod));
    return cudaGetLastError();
}

cudaError_t launch_adaln(const float* x, const float* mod, int mod_stride, int mod_off, __half* xh, __half* xl,
                         int B, int S_pad, int c, float eps, cudaStream_t st) {
    const int rows = B * S_pad;
    if (rows <= 0 || (mod_off % 4) || (mod_stride % 4)) return cudaErrorInvalidValue;
    const AdalnSrc none{};
    float* xm = const_cast<float*>(x);       // SRC_X never writes x
    if (c == 128)
        PDK_LAUNCH_CHECK(launch_pdl(adaln_kernel<128, SRC_X>, dim3((rows + 7) / 8), dim3(256), (size_t)(0), st, xm, mod, mod_stride, mod_off, xh, xl, rows, S_pad, eps, none));
    else if (c == 512)
        PDK_LAUNCH_CHECK(launch_pdl(adaln_kernel<512, SRC_X>, dim3((rows + 7) / 8), dim3(256), (size_t)(0), st, xm, mod, mod_stride, mod_off, xh, xl, rows, S_pad, eps, none));
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t launch_precond_adaln(const float* x_hat, const float* coef, int coef_stride, const float* a, const float* wx,
                                 const float* bx, float* ba, const float* mod, int mod_stride, int mod_off, __half* xh,
                                 __half* xl, int B, int Na, int S_pad, int c_a, float eps, cudaStream_t st) {
    const int rows = B * S_pad;
    if (rows <= 0 || c_a != 128 || Na > S_pad || (mod_off % 4) || (mod_stride % 4)) return cudaErrorInvalidValue;
    AdalnSrc src{};
    src.x_hat = x_hat; src.coef = coef; src.coef_stride = coef_stride; src.a = a; src.wx = wx; src.bx = bx; src.Na = Na;
    PDK_LAUNCH_CHECK(launch_pdl(adaln_kernel<128, SRC_PRECOND>, dim3((rows + 7) / 8), dim3(256), (size_t)(0), st, ba, mod, mod_stride, mod_off, xh, xl, rows, S_pad, eps, src));
    return cudaGetLastError();
}

cudaError_t launch_upscale_adaln(float* ba, const float* up, const int* atom2tok, const float* mod, int mod_stride,
                                 int mod_off, __half* xh, __half* xl, int B, int Na, int Sa_pad, int St_pad, int c_a,
                                 float eps, cudaStream_t st) {
    const int rows = B * Sa_pad;
    if (rows <= 0 || c_a != 128 || Na > Sa_pad || (mod_off % 4) || (mod_stride % 4)) return cudaErrorInvalidValue;
    AdalnSrc src{};
    src.up = up; src.atom2tok = atom2tok; src.St_pad = St_pad; src.Na = Na;
    PDK_LAUNCH_CHECK(launch_pdl(adaln_kernel<128, SRC_UPSCALE>, dim3((rows + 7) / 8), dim3(256), (size_t)(0), st, ba, mod, mod_stride, mod_off, xh, xl, rows, Sa_pad, eps, src));
    return cudaGetLastError();
}

cudaError_t launch_split(const float* x, __half* xh, __half* xl, size_t n, cudaStream_t st) {
    if (n == 0 || n % 4) return cudaErrorInvalidValue;
    PDK_LAUNCH_CHECK(launch_pdl(split_kernel, dim3(grid_for(n / 4)), dim3(256), (size_t)(0), st, x, xh, xl, n / 4));
    return cudaGetLastError();
}

cudaError_t launch_precond(const float* x_hat, const float* coef, int coef_stride, const float* a, const float* wx,
                           const float* bx, float* ba, int B, int Na, int S_pad, int c_a, cudaStream_t st) {
    if (c_a % 4 || Na > S_pad || (size_t)B * S_pad * (c_a / 4) >= (1ull << 31)) return cudaErrorInvalidValue;
    PDK_LAUNCH_CHECK(launch_pdl(precond_kernel, dim3(grid_for((size_t)B * S_pad * (c_a / 4))), dim3(256), (size_t)(0), st, x_hat, coef, a, wx, bx, ba, B, Na, S_pad, c_a, coef_stride));
    return cudaGetLastError();
}

cudaError_t launch_segment_mean(const float* h, const int* tok_start, const float* s, float* bs, int B, int Nt,
                                int Sa_pad, int St_pad, int c_s, cudaStream_t st) {
    if (c_s % 4 || Nt > St_pad || (size_t)B * St_pad * (c_s / 4) >= (1ull << 31)) return cudaErrorInvalidValue;
    PDK_LAUNCH_CHECK(launch_pdl(segment_mean_kernel, dim3(grid_for((size_t)B * St_pad * (c_s / 4))), dim3(256), (size_t)(0), st, h, tok_start, s, bs, B, Nt, Sa_pad,
                                                                               St_pad, c_s));
    return cudaGetLastError();
}

cudaError_t launch_gather_add(float* ba, const float* up, const int* atom2tok, int B, int Na, int Sa_pad,
                              int St_pad, int c_a, cudaStream_t st) {
    if (c_a % 4 || (size_t)B * Na * (c_a / 4) >= (1ull << 31)) return cudaErrorInvalidValue;
    PDK_LAUNCH_CHECK(launch_pdl(gather_add_kernel, dim3(grid_for((size_t)B * Na * (c_a / 4))), dim3(256), (size_t)(0), st, ba, up, atom2tok, B, Na, Sa_pad, St_pad, c_a));
    return cudaGetLastError();
}

cudaError_t launch_denoise_out(const float* ba, const float* x_hat, const float* coef, int coef_stride, const float* ln_w,
                               const float* ln_b, const float* wr, float* x_den, int B, int Na, int S_pad,
                               int c_a, float eps, float* x_next, cudaStream_t st) {
    if (c_a != 128) return cudaErrorInvalidValue;
    PDK_LAUNCH_CHECK(launch_pdl(denoise_out_kernel, dim3((B * Na + 8 * kDenoiseRows - 1) / (8 * kDenoiseRows)), dim3(256), (size_t)(0), st, ba, x_hat, coef, ln_w, ln_b, wr, x_den, B, Na, S_pad, eps,
                                coef_stride, x_next));
    return cudaGetLastError();
}

}  // namespace pdk
