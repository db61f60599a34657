// Blackwell (sm_100a) building blocks: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 MMA / TMEM, and the
// host-side tensor-map cache.  Every encoding used here was validated on a B200 by tests/cuda/umma_probe.cu
// (profiles/r01_umma_probe.log): K-major SWIZZLE_128B and SWIZZLE_64B operand tiles, MN-major SWIZZLE_64B B
// tiles (SBO = 512 B), kind::f16 with fp32 accumulation in TMEM, tcgen05.ld 32x32b.x32.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace pdk {

// One lane of a fully converged warp (elect.sync).  Issue code for TMA / tcgen05 must sit under this predicate
// in warp-uniform control flow: under a divergent `if (lane == 0)` the compiler cannot prove uniformity of the
// uniform-register operands and wraps every UTCHMMA / UTMALDG in an ELECT + BRA.U.ANY waterfall loop
// (~100 cycles per issued MMA, measured: profiles/r01_attention_v2a_ncu.txt).
PDK_DEV bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------------------------- mbarrier
PDK_DEV void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
PDK_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
PDK_DEV void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
PDK_DEV void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
PDK_DEV bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
PDK_DEV void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// ------------------------------------------------------------------------------------- TMA
PDK_DEV void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
PDK_DEV void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
PDK_DEV void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ------------------------------------------------------------------------------------- tcgen05 / TMEM
PDK_DEV void tmem_alloc(uint32_t dst_smem, uint32_t cols) {     // whole warp; cols = power of two >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
PDK_DEV void tmem_dealloc(uint32_t addr, uint32_t cols) {       // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}
PDK_DEV void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
PDK_DEV void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], fp16 inputs, fp32 accumulate.  One thread issues.
PDK_DEV void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate));
}
// mbarrier arrive once all previously issued MMAs of this thread have completed (implies fence::before_thread_sync)
PDK_DEV void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t <-> TMEM lane base+t)
PDK_DEV void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr));
}
PDK_DEV void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------- CTA pairs (cta_group::2)
// Validated by tests/cuda/umma_probe.cu test 10.  A cluster of two CTAs computes one M=256 tile: each CTA holds its 128
// rows of A and half of the B rows; the LEADER (cluster rank 0) issues MMAs that read both CTAs' shared memory and
// write both CTAs' TMEM.  Peer TMA loads signal the leader's mbarrier (address with the peer bit cleared); commits are
// multicast to the same barrier offset in both CTAs.
PDK_DEV uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
PDK_DEV void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
PDK_DEV void tmem_alloc_2sm(uint32_t dst_smem, uint32_t cols) {      // the same warp of BOTH CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
}
PDK_DEV void tmem_dealloc_2sm(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}
PDK_DEV void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
PDK_DEV void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate));
}
PDK_DEV void umma_commit_2sm(uint32_t bar) {      // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
PDK_DEV void mbar_arrive_cluster(uint32_t local_bar, uint32_t cta) {      // arrive on the same barrier offset in CTA `cta`
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_bar), "r"(cta));
    // default (.release.cta) semantics: the arrival only says "this warp has drained its accumulator chunk" (its tcgen05.ld has
    // completed, tcgen05.fence::before_thread_sync issued); no generic-proxy write has to become visible to the other CTA.  The
    // .release.cluster form cost a cluster-scope fence per epilogue warp and tile (ncu: membar = 4.7 stall cycles per issue):
    // token QKV 18.55 -> 18.2 us, SwiGLU 25.4 -> 24.95, w2 17.25 -> 16.6.
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64) (2 = SWIZZLE_128B, 4 = SWIZZLE_64B).
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw64 = 4;
PDK_DEV uint64_t smem_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout) {
    return ((uint64_t)(((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (layout << 29)) << 32) | (1ull << 16) |
           (uint64_t)((smem_addr >> 4) & 0x3FFF);
}
// Same with an explicit leading-dimension byte offset: for MN-major operands LBO is the distance between
// consecutive swizzle atoms along MN (validated: tests/cuda/umma_probe.cu test 4 variant 1).
PDK_DEV uint64_t smem_desc_lbo(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return ((uint64_t)(((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (layout << 29)) << 32) |
           ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | (uint64_t)((smem_addr >> 4) & 0x3FFF);
}
// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1, a/b format F16=0,
// a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, bool b_mn_major = false) {
    return (1u << 4) | (b_mn_major ? (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------- host: tensor maps
// 2-D fp16 row-major [rows][cols] (leading dimension ld elements), box [box_rows][box_cols], swizzle 64/128.
// Maps are cached by value of their arguments; pointers into torch allocations stay valid while referenced.
cudaError_t get_tensor_map_f16(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                               uint32_t box_cols, int swizzle_bytes, CUtensorMap* out);
// 2-D fp32 map (bias tiles)
cudaError_t get_tensor_map_f32(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                               uint32_t box_cols, int swizzle_bytes, CUtensorMap* out);

}  // namespace pdk
