// tcgen05 / TMEM / UMMA-descriptor wrappers shared by the encoder kernels and the tensor-core top-k screen (sm_100a).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace ebsd {

// ---------------------------------------------------------------- tcgen05 wrappers
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// e4m3 x e4m3 -> fp32, K = 32 per instruction (twice the fp16 rate); operands byte-packed in the same K-major layouts
__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

#ifndef EBSD_WAIT_HINT_NS
#define EBSD_WAIT_HINT_NS 20000u
#endif
#ifdef EBSD_DEBUG_NOTRAP
// debugging build (tools/): a timed-out wait records (block, thread, barrier address, parity) and RETURNS, so that the
// kernel drains and the host can read which wait starved (ebsd_debug_timeout_info)
__device__ unsigned long long g_timeout_info[4];
#define EBSD_TIMEOUT_ACTION(bar_u32, parity)                                                                        \
    do {                                                                                                            \
        if (atomicCAS(&g_timeout_info[0], 0ull, 1ull) == 0ull) {                                                     \
            g_timeout_info[1] = ((unsigned long long)blockIdx.x << 32) | threadIdx.x;                                \
            g_timeout_info[2] = (unsigned long long)(bar_u32);                                                       \
            g_timeout_info[3] = (unsigned long long)(parity);                                                        \
        }                                                                                                           \
        return;                                                                                                     \
    } while (0)
#define EBSD_TIMEOUT_CYCLES 400000000ll
#else
#define EBSD_TIMEOUT_ACTION(bar_u32, parity)                                                              \
    do {                                                                                                  \
        printf("ebsd: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);                 \
        __trap();                                                                                         \
    } while (0)
#define EBSD_TIMEOUT_CYCLES 4000000000ll   // ~2 s
#endif
// Bounded mbarrier wait: a mis-programmed pipeline must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_hint(bar, parity, EBSD_WAIT_HINT_NS)) {
        if (clock64() - t0 > EBSD_TIMEOUT_CYCLES) EBSD_TIMEOUT_ACTION(smem_u32(bar), parity);
    }
}

// The same wait by 32-bit shared-window address (for hot loops: no generic-pointer conversion per call).
__device__ __forceinline__ void mbar_wait_bounded_u32(uint32_t bar_u32, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar_u32), "r"(parity)
        : "memory");
    if (ok) return;
    const long long t0 = clock64();
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar_u32), "r"(parity), "r"(EBSD_WAIT_HINT_NS)
            : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) {
            printf("ebsd: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// UMMA shared-memory descriptor, K-major operand whose rows are SWB bytes wide (SWB = 64 or 128 = swizzle span).
template <int SWB>
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {
    constexpr uint64_t layout = SWB == 128 ? 2ull : 4ull;  // SWIZZLE_128B / SWIZZLE_64B
    constexpr uint64_t sbo = (8ull * SWB) >> 4;            // 8-row group stride
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (0ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
// Instruction descriptor: fp16 x fp16 -> fp32 (kind::f16) or e4m3 x e4m3 -> fp32 (kind::f8f6f4: format code 0 is
// E4M3 there), both operands K-major, M = 128.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int n) {
    return (1u << 4)                      // D format: F32
           | (0u << 7) | (0u << 10)       // A, B format: F16
           | ((uint32_t)(n >> 3) << 17)   // N
           | ((uint32_t)(128 >> 4) << 24);  // M
}

}  // namespace ebsd
