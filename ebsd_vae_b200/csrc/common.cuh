// Shared helpers for libebsd_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ebsd_b200.h"

namespace ebsd {

// thread-local error message behind ebsd_last_error()
void set_error(const char *fmt, ...);
int check_device_arch();  // EBSD_OK when the current device is compute capability 10.x
int sm_count();

#define EBSD_CUDA_TRY(expr)                                                                        \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ::ebsd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return EBSD_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

// after every kernel launch: count it (ebsd_launch_count) and surface launch errors
void note_launch();
#define EBSD_LAUNCH_CHECK()                \
    do {                                   \
        ::ebsd::note_launch();             \
        EBSD_CUDA_TRY(cudaGetLastError()); \
    } while (0)

#define EBSD_REQUIRE(cond, ...)                \
    do {                                       \
        if (!(cond)) {                         \
            ::ebsd::set_error(__VA_ARGS__);    \
            return EBSD_ERR_ARG;               \
        }                                      \
    } while (0)

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
typedef CUresult (*tensormap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);
tensormap_encode_fn get_tensormap_encode();

// ---------------------------------------------------------------- device-side PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint (ns): the thread sleeps in hardware until the phase completes or the hint expires,
// instead of coming back every few dozen cycles -- polling warps otherwise take issue slots from the working warps
// (ncu on the top-k screen: 3/4 of all executed instructions were polling loops).
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t *bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// One lane of a converged warp; unlike `lane == 0` the compiler knows the branch holds a single thread and keeps
// tcgen05 / TMA operands in uniform registers (no per-instruction ELECT + branch loop).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, int c3,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
            "r"(smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// Programmatic dependent launch (the encoder chain: a kernel launched with the programmatic-stream-serialization
// attribute may become resident while its predecessor still runs).  griddep_launch_dependents: this CTA no longer
// objects to the NEXT kernel's CTAs being scheduled (they take SMs as this grid's CTAs exit and run their prologue:
// barrier init, TMEM allocation, weight loads).  griddep_wait: blocks until the PREVIOUS grid has completed and its
// memory is visible -- every thread executes it before touching anything an earlier kernel of the chain wrote or
// still reads.  Both are no-ops in a launch without the attribute.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

}  // namespace ebsd
