// Microbenchmark: issue rate of tcgen05.mma.cta_group::1.kind::f16 (M = 128, K = 16, operands in shared memory)
// as a function of N, and the cost of alternating two instruction descriptors (N1, N2) in runs of `group` MMAs.
// All 148 SMs run the same loop; cycles are clock64() around issue + commit + wait on one SM.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mma_rate mma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <utility>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc(uint32_t saddr, int rowb) {
    const uint64_t layout = rowb == 128 ? 2ull : (rowb == 64 ? 4ull : 6ull);   // SWIZZLE_128B / 64B / 32B
    const uint64_t sbo = (8ull * rowb) >> 4;
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ uint32_t idesc(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// KIND 1: kind::f8f6f4 (e4m3, K = 32 per instruction) instead of kind::f16
template <int KIND>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    if (KIND)
        asm volatile("{.reg .pred p; setp.eq.b32 p, 0, 0; tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;}"
                     ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
    else
        asm volatile("{.reg .pred p; setp.eq.b32 p, 0, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}"
                     ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
}

// mode 0: all MMAs use N1.  mode 1: runs of `group` MMAs with N1 then `group` MMAs with N2 (different TMEM columns).
template <int KIND>
__global__ void __launch_bounds__(128, 1) rate(int n1, int n2, int group, int reps, int rowb, int shift_rows,
                                               long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t mbar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tslot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = s32(smem), b0 = s32(smem) + 48 * 1024;
        const uint32_t id1 = idesc(n1), id2 = idesc(n2);
        const uint64_t da = desc(a0 + shift_rows * rowb, rowb), db = desc(b0, rowb);
        const uint64_t da2 = desc(a0 + 16384 + shift_rows * rowb, rowb);
        const int kstep = rowb >= 128 ? 2 : (rowb == 64 ? ((0 + 1) & 1) * 2 : 0);   // 32-byte K steps inside a row (0: re-read)
        const long long t0 = clock64();
        if (group >= 8) {
            for (int r = 0; r < reps; ++r) {
                for (int g = 0; g < group; g += 8) {
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        mma<KIND>(tmem, da + (u & 3) * kstep + (u >> 2) * 64, db + (u & 3) * kstep, id1, 1u);
                }
                if (n2 > 0)
                    for (int g = 0; g < group; g += 8) {
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            mma<KIND>(tmem + 256, da2 + (u & 3) * kstep + (u >> 2) * 64, db + (u & 3) * kstep, id2, 1u);
                    }
            }
        } else {
            for (int r = 0; r < reps; ++r) {
                for (int g = 0; g < group; ++g) mma<KIND>(tmem, da + g * 2, db + g * 2, id1, 1u);
                if (n2 > 0)
                    for (int g = 0; g < group; ++g) mma<KIND>(tmem + 256, da2 + g * 2, db + g * 2, id2, 1u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mbar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(&mbar)));
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// runs of `group` kind::f16 MMAs (N = n1) alternating with runs of `group` kind::f8f6f4 MMAs (N = n2): the cost of a
// KIND switch in the issue stream (the encoder blocks issue both kinds into one accumulator pair)
__global__ void __launch_bounds__(128, 1) rate_kinds(int n1, int n2, int group, int reps, long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t mbar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tslot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = s32(smem), b0 = s32(smem) + 48 * 1024;
        const uint32_t id1 = idesc(n1), id2 = idesc(n2);
        const uint64_t da = desc(a0 + 128, 128), db = desc(b0, 128), da2 = desc(a0 + 16384 + 128, 128);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            for (int g = 0; g < group; ++g) mma<0>(tmem, da + (g & 3) * 2, db + (g & 3) * 2, id1, 1u);
            for (int g = 0; g < group; ++g) mma<1>(tmem + 256, da2 + (g & 3) * 2, db + (g & 3) * 2, id2, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mbar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(&mbar)));
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main(int argc, char **argv) {
    long long *d, h;
    CK(cudaMalloc(&d, 8));
    CK(cudaFuncSetAttribute(rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(rate_kinds, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    if (argc > 1 && argv[1][0] == 'k') {
        // kind alternation only: ./mma_rate k
        for (auto pr : {std::pair<int, int>{128, 128}, {64, 64}, {64, 32}, {256, 256}})
            for (int group : {1, 2, 4, 8, 12, 16, 36, 72}) {
                const int reps = 2304 / group;
                for (int it = 0; it < 2; ++it) {
                    rate_kinds<<<148, 128, 100 * 1024>>>(pr.first, pr.second, group, reps, d);
                    CK(cudaDeviceSynchronize());
                }
                CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
                const double per = (double)h / (2.0 * reps * group);
                const double floor_ = (pr.first / 2.0 > 46 ? pr.first / 2.0 : 46.0) * 0.5 + (pr.second / 2.0 > 46 ? pr.second / 2.0 : 46.0) * 0.5;
                printf("f16 N=%3d x %d  <->  f8 N=%3d x %d : %7.1f cycles per MMA (unswitched ~%5.1f) -> %6.1f cycles per switch\n",
                       pr.first, group, pr.second, group, per, floor_, (per - floor_) * group);
            }
        cudaFree(d);
        return 0;
    }
    int kind = 0;
    auto run = [&](int n1, int n2, int group, int rowb, int shift) {
        const int reps = 4096 / group;
        for (int it = 0; it < 2; ++it) {
            if (kind) rate<1><<<148, 128, 100 * 1024>>>(n1, n2, group, reps, rowb, shift, d);
            else rate<0><<<148, 128, 100 * 1024>>>(n1, n2, group, reps, rowb, shift, d);
            CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
        const int total = reps * group * (n2 > 0 ? 2 : 1);
        const double ideal = reps * group * (n1 / 2.0 + (n2 > 0 ? n2 / 2.0 : 0.0));
        printf("rowb=%3d shift=%2d N1=%3d N2=%3d group=%4d : %9lld cycles, %7.1f per MMA, floor %7.1f -> %5.1f %% of floor rate\n",
               rowb, shift, n1, n2, group, h, (double)h / total, ideal / total, 100.0 * ideal / (double)h);
    };
    for (kind = 0; kind < 2; ++kind) {
        printf("---- %s\n", kind ? "kind::f8f6f4 (K = 32)" : "kind::f16 (K = 16)");
        for (int rowb : {128, 64, 32})
            for (int n : {16, 32, 64, 96, 128, 256}) run(n, 0, 64, rowb, 1);
        // the front-end block's mix: fp16 N = 64 twice + (second kind set by g_kind_f8 only: run separately)
    }
    kind = 0;
    for (int group : {1, 2, 4, 8, 16, 32, 72})
        for (auto pr : {std::pair<int, int>{64, 32}, {128, 64}, {256, 128}}) run(pr.first, pr.second, group, 128, 1);
    for (int group : {1, 4, 16}) for (int n : {64, 128, 256}) run(n, n, group, 128, 1);
    cudaFree(d);
    return 0;
}
