// Probe for the front-end block (encoder_front.cuh): K-major operands with 32-BYTE rows under SWIZZLE_32B, written by
// threads (not TMA), read by tcgen05.mma through descriptors whose start is shifted by whole rows and whose 8-row
// group stride (SBO) is an arbitrary multiple of 32 bytes -- for kind::f16 (K = 16 halves per row) and kind::f8f6f4
// (K = 32 e4m3 bytes per row).
//
// A [R rows][32 B] is written so that the 16-byte chunk c of row r lands at  r*32 + ((c ^ ((addr >> 7) & 1)) << 4)
// (addr = absolute shared-memory address of the row), B selects one K element per column (B[n][k] = (k == n)), so
// D[m][n] = A_view[m][n]: the values tell which shared-memory element the tensor core actually read.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma32 probe_umma32.cu
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int RA = 384;   // rows of A in shared memory
constexpr int ROWB = 32;

// value of A element (r, k) as the test pattern: mode 0 -> r % 16, mode 1 -> (r / 16) % 16, mode 2 -> k % 16
__device__ __host__ inline int pattern(int mode, int r, int k) { return mode == 0 ? r % 16 : (mode == 1 ? (r / 16) % 16 : k % 16); }

template <bool FP8>
__global__ void probe(int mode, int shift, int gstride, float *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = smem + RA * ROWB;   // 12288: a multiple of 1024
    __shared__ uint64_t mbar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int KE = FP8 ? 32 : 16;   // K elements per row
    // ---- write A and B with the assumed swizzle
    for (int i = threadIdx.x; i < (RA + 32) * 2; i += blockDim.x) {
        const int r = i >> 1, c = i & 1;
        const bool isB = r >= RA;
        const int rr = isB ? r - RA : r;
        uint8_t *base = isB ? sB : sA;
        const uint32_t row_addr = s32(base) + rr * ROWB;
        uint8_t *dst = base + rr * ROWB + (((c ^ ((row_addr >> 7) & 1))) << 4);
        if (FP8) {
            uint8_t v[16];
            for (int j = 0; j < 16; ++j) {
                const int k = c * 16 + j;
                const float f = isB ? (k == rr ? 1.f : 0.f) : (float)pattern(mode, rr, k);
                v[j] = (uint8_t)__nv_cvt_float_to_fp8(f, __NV_SATFINITE, __NV_E4M3);
            }
            *(uint4 *)dst = *(const uint4 *)v;
        } else {
            __half v[8];
            for (int j = 0; j < 8; ++j) {
                const int k = c * 8 + j;
                v[j] = __float2half(isB ? (k == rr ? 1.f : 0.f) : (float)pattern(mode, rr, k));
            }
            *(uint4 *)dst = *(const uint4 *)v;
        }
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(s32(&tslot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        constexpr uint64_t layout = 6ull;   // SWIZZLE_32B
        const uint64_t sbo = ((uint64_t)gstride * ROWB) >> 4;
        const uint32_t a_start = s32(sA) + shift * ROWB;
        const uint32_t b_start = s32(sB);
        const uint64_t adesc = (uint64_t)((a_start & 0x3ffff) >> 4) | (sbo << 32) | (1ull << 46) | (layout << 61);
        const uint64_t bdesc = (uint64_t)((b_start & 0x3ffff) >> 4) | ((uint64_t)((8 * ROWB) >> 4) << 32) | (1ull << 46) | (layout << 61);
        const uint32_t idesc = (1u << 4) | ((uint32_t)(KE >> 3) << 17) | ((128u >> 4) << 24);   // N = KE
        if (FP8)
            asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;}"
                         ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0) : "memory");
        else
            asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}"
                         ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mbar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(&mbar)));
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp < 4) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 32 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

template <bool FP8>
void run() {
    float *dOut;
    CK(cudaMalloc(&dOut, 128 * 32 * 4));
    CK(cudaFuncSetAttribute(probe<FP8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
    std::vector<float> out(128 * 32);
    constexpr int KE = FP8 ? 32 : 16;
    int total_bad = 0;
    for (int mode = 0; mode < 3; ++mode)
        for (int gstride : {8, 10, 18, 20})
            for (int shift = 0; shift <= 40; ++shift) {
                if (15 * gstride + 8 + shift > RA) continue;
                probe<FP8><<<1, 128, 32 * 1024>>>(mode, shift, gstride, dOut);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
                int bad = 0, fm = -1, fn = -1;
                float got = 0, want = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < KE; ++n) {
                        const int r = (m % 8) + (m / 8) * gstride + shift;
                        const float w = (float)pattern(mode, r, n);
                        if (out[m * 32 + n] != w) {
                            if (!bad) { fm = m; fn = n; got = out[m * 32 + n]; want = w; }
                            ++bad;
                        }
                    }
                total_bad += bad;
                if (bad || shift == 0)
                    printf("%s mode=%d gstride=%2d shift=%2d : %s (%d bad, first m=%d n=%d got %.1f want %.1f)\n",
                           FP8 ? "fp8 " : "fp16", mode, gstride, shift, bad ? "MISMATCH" : "ok", bad, fm, fn, got, want);
            }
    printf("%s: %s\n", FP8 ? "kind::f8f6f4 32-byte rows SWIZZLE_32B" : "kind::f16 32-byte rows SWIZZLE_32B",
           total_bad ? "MISMATCHES" : "all shifts / group strides ok");
    cudaFree(dOut);
}

int main() {
    CK(cudaFree(0));
    run<false>();
    run<true>();
    return 0;
}
