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

// Measurement switches (A/B toggles read from the environment) exist ONLY in debug builds compiled with -DPDK_MEASURE
// (tools/gemm_variants.sh); in the release library they are compile-time false, so no environment variable can change
// the hot path.
#ifdef PDK_MEASURE
inline bool measure_switch(const char* name) { return getenv(name) != nullptr; }
#else
constexpr bool measure_switch(const char*) { return false; }
#endif

// Per-device one-time setup (cudaFuncSetAttribute, SM count): keyed by the current device so that a process driving
// several GPUs configures each of them.
constexpr int kMaxDevices = 64;
struct PerDevice {
    bool done[kMaxDevices] = {};
    int sms[kMaxDevices] = {};
};
inline cudaError_t current_device(int* dev) {
    cudaError_t e = cudaGetDevice(dev);
    if (e != cudaSuccess) return e;
    return (*dev < 0 || *dev >= kMaxDevices) ? cudaErrorInvalidDevice : cudaSuccess;
}
inline cudaError_t device_sm_count(int* sms) {
    static PerDevice pd;
    int dev = 0;
    cudaError_t e = current_device(&dev);
    if (e != cudaSuccess) return e;
    if (!pd.done[dev]) {
        if ((e = cudaDeviceGetAttribute(&pd.sms[dev], cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        pd.done[dev] = true;
    }
    *sms = pd.sms[dev];
    return cudaSuccess;
}
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device); `pd` is the kernel's own static PerDevice
template <typename K>
inline cudaError_t ensure_smem(PerDevice& pd, K kernel, int bytes) {
    int dev = 0;
    cudaError_t e = current_device(&dev);
    if (e != cudaSuccess) return e;
    if (!pd.done[dev]) {
        if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
        pd.done[dev] = true;
    }
    return cudaSuccess;
}

PDK_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
    static const int pdl = !measure_switch("PDK_NO_PDL");
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace pdk
