// Shared device helpers for the physdock_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>

#define PDK_DEV __device__ __forceinline__

namespace pdk {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kPadBias = -1.0e30f;     // bias of padded key columns: exp2(kPadBias - m) == 0, stays finite
constexpr int   kHeadDim = 32;           // attentions.py:223 (c_hidden)
constexpr int   kTimeDim = 256;          // timestep_embeddings.py:157

// ---- Programmatic Dependent Launch ---------------------------------------------------------------------------
// Every kernel of the sampling step is launched with cudaLaunchAttributeProgrammaticStreamSerialization: it calls
// griddep_launch() at entry (the next kernel in the stream may be scheduled as soon as all CTAs of this grid have
// started) and griddep_wait() before it reads anything a predecessor wrote (returns when the preceding grid has
// completed and flushed).  Launch latency and kernel prologues thus overlap the previous kernel's tail.
PDK_DEV void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
PDK_DEV void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

PDK_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- async copies --------------------------------------------------------------------------
PDK_DEV void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
}
PDK_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> PDK_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- warp-level matrix fragments (legacy tensor path; see DESIGN.md "v1 kernels") -----------
PDK_DEV void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
PDK_DEV void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// D(16x8,f32) += A(16x16,f16,row) * B(16x8,f16,col)
PDK_DEV void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---- split-fp16 number format ----------------------------------------------------------------
// x (fp32) ~= hi + lo with hi = fp16(x), lo = fp16(x - hi): 22 significant bits.  A product a*b is
// evaluated on the tensor cores as a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with fp32 accumulation, which
// tools/precision_probe.py shows is indistinguishable from fp32 at the 1e-3 Angstrom parity budget.
constexpr float kHalfMax = 65504.f;
// Values beyond +-65504 saturate (cvt.satfinite: the documented limit of the format); in range the result is exactly
// hi = fp16(x), lo = fp16(x - hi).  6 instructions per pair (the explicit fminf/fmaxf clamp cost 4 more).
PDK_DEV void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - hf.y), "f"(x0 - hf.x));
}
// 1/sqrt(x) from MUFU rsqrt + one Newton step (relative error ~1e-7; the IEEE sqrt + divide pair is ~25 instructions)
PDK_DEV float inv_sqrt(float x) {
    const float y = rsqrtf(x);
    return y * fmaf(-0.5f * x * y, y, 1.5f);
}
// same, for values known to lie in [0, 1] (softmax probabilities)
PDK_DEV void split2_unit(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    __half2 h = __floats2half2_rn(x0, x1);
    float2 hf = __half22float2(h);
    __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = *reinterpret_cast<uint32_t*>(&l);
}

// 64-byte-row tile (32 halves per row): 16-byte chunk c of row r lives at chunk c ^ ((r>>1)&3).
// Makes both the cp.async fills and every 8-row ldmatrix phase bank-conflict free.
PDK_DEV uint32_t swz64(int row, int chunk) { return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); }

// ---- math --------------------------------------------------------------------------------------
PDK_DEV float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}
PDK_DEV float silu(float x) { return x / (1.0f + expf(-x)); }

PDK_DEV float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
PDK_DEV double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}


// Sums of EIGHT per-lane values over the warp at once: after three exchange steps each lane holds one row's partial sum
// (17 shuffles instead of the 40 of eight separate butterflies); every lane ends with all eight totals.
PDK_DEV void warp_sum8(float (&s)[8]) {
    const unsigned lane = threadIdx.x & 31u;
    const bool b16 = lane & 16u, b8 = lane & 8u, b4 = lane & 4u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b16 ? s[i] : s[i + 4], keep = b16 ? s[i + 4] : s[i];
        s[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b8 ? s[i] : s[i + 2], keep = b8 ? s[i + 2] : s[i];
        s[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
        const float send = b4 ? s[0] : s[1], keep = b4 ? s[1] : s[0];
        s[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], 2);
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], 1);
    // s[0] is now the total of row 4*[lane&16] + 2*[lane&8] + [lane&4]
    const float t = s[0];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = __shfl_sync(0xffffffffu, t, ((i >> 2) & 1) * 16 + ((i >> 1) & 1) * 8 + (i & 1) * 4);
}

#define PDK_LAUNCH_CHECK(expr)                 \
    do {                                       \
        cudaError_t e_ = (expr);               \
        if (e_ != cudaSuccess) return e_;      \
    } while (0)

// host side: <<<>>> with the PDL attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    static const int pdl = getenv("PDK_NO_PDL") == nullptr;     // measurement switch; PDL is on in production
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace pdk
