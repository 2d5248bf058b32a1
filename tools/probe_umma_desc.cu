// Probe: how does tcgen05.mma address a K-major swizzled operand whose descriptor start is shifted by whole rows?
//
// A [R rows][64 fp16] (128-byte rows) is loaded by TMA with SWIZZLE_128B into 1024-aligned smem.  B [16][64] selects
// column n (B[n][k] = k == n).  One MMA (M=128, N=16, K=16) with the A descriptor start = base + shift*128 + koff*32
// gives D[m][n] = A_smem_view[m][n]: the values tell which smem element the tensor core actually read.
// Variants: base_offset field = 0, or (start >> 7) & 7.
// Same for 64-byte rows / SWIZZLE_64B.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma_desc probe_umma_desc.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ROWB>  // row bytes: 128 or 64
__global__ void probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int shift,
                      int koff, int use_base_offset, int gstride, float *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    constexpr int RA = 320;
    uint8_t *sA = smem;
    uint8_t *sB = smem + RA * ROWB + ((RA * ROWB) % 1024 ? 1024 - (RA * ROWB) % 1024 : 0);
    __shared__ uint64_t bar, mbar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(s32(&tslot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(RA * ROWB + 16 * ROWB));
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(sA)), "l"(&mapA), "r"(0), "r"(0), "r"(s32(&bar)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(sA + (RA / 2) * ROWB)), "l"(&mapA), "r"(0), "r"(RA / 2), "r"(s32(&bar)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(sB)), "l"(&mapB), "r"(0), "r"(0), "r"(s32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(&bar)));
        asm volatile("tcgen05.fence::after_thread_sync;");
        constexpr uint64_t layout = ROWB == 128 ? 2ull : 4ull;
        const uint64_t sbo = ((uint64_t)gstride * ROWB) >> 4;  // 8-row group stride, in rows of the window
        const uint32_t a_start = s32(sA) + shift * ROWB + koff * 32;
        const uint32_t b_start = s32(sB) + koff * 32;
        uint64_t bo = use_base_offset ? (uint64_t)((a_start >> 7) & 7) : 0ull;
        const uint64_t adesc = (uint64_t)((a_start & 0x3ffff) >> 4) | (sbo << 32) | (1ull << 46) | (bo << 49) | (layout << 61);
        const uint64_t bdesc = (uint64_t)((b_start & 0x3ffff) >> 4) | ((uint64_t)((8 * ROWB) >> 4) << 32) | (1ull << 46) | (layout << 61);
        const uint32_t idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}"
                     ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mbar)) : "memory");
        ok = 0;
        while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(&mbar)));
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp < 4) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 16 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROWB>
void run(enc_fn enc) {
    constexpr int KW = ROWB / 2;  // fp16 per row
    constexpr int RA = 320;
    std::vector<__half> hA(RA * KW), hB(16 * KW);
    // (values) A[r][k] = r + k/64 exactly representable: r < 192 (8 bits) + k/64 (6 fractional bits) fits fp16's 11-bit mantissa
    __half *dA, *dB; float *dOut;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dOut, 128 * 16 * 4));
    CUtensorMap mA, mB;
    cuuint64_t gdA[2] = {(cuuint64_t)KW, RA}, gs[1] = {(cuuint64_t)ROWB};
    cuuint32_t bxA[2] = {(cuuint32_t)KW, RA / 2}, es[2] = {1, 1};
    CUtensorMapSwizzle sw = ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    if (enc(&mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, gdA, gs, bxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode A failed\n"); exit(1); }
    cuuint64_t gdB[2] = {(cuuint64_t)KW, 16};
    cuuint32_t bxB[2] = {(cuuint32_t)KW, 16};
    if (enc(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, gdB, gs, bxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode B failed\n"); exit(1); }
    CK(cudaFuncSetAttribute(probe<ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    std::vector<float> out(128 * 16);
    const int nk = KW / 16;
    for (int mode = 0; mode < 2; ++mode) {
    // mode 0: A[r][k] = r (checks row addressing), mode 1: A[r][k] = k (checks column / swizzle addressing)
    for (int r = 0; r < RA; ++r) for (int k = 0; k < KW; ++k) hA[r * KW + k] = __float2half(mode == 0 ? (float)r : (float)k);
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    for (int gstride : {8, 10, 18, 34}) {
    for (int ubo = 0; ubo < 1; ++ubo) {
        for (int koff = 0; koff < nk; ++koff) {
            // B[n][k] = 1 where k == koff*16 + n
            for (auto &x : hB) x = __float2half(0.f);
            for (int n = 0; n < 16; ++n) hB[n * KW + koff * 16 + n] = __float2half(1.f);
            CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
            for (int shift = 0; shift <= 17; ++shift) {
                if (15 * gstride + 8 + shift > RA) continue;
                probe<ROWB><<<1, 128, 64 * 1024>>>(mA, mB, shift, koff, ubo, gstride, dOut);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
                int bad = 0, first_m = -1, first_n = -1; float got = 0, want = 0;
                for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
                    const float w = mode == 0 ? (float)((m % 8) + (m / 8) * gstride + shift) : (float)(koff * 16 + n);
                    if (out[m * 16 + n] != w) { if (!bad) { first_m = m; first_n = n; got = out[m * 16 + n]; want = w; } ++bad; }
                }
                printf("rowB=%d mode=%d gstride=%d koff=%d shift=%2d : %s", ROWB, mode, gstride, koff, shift, bad ? "MISMATCH" : "ok");
                if (bad) printf(" (%d bad, first m=%d n=%d got %.4f want %.4f)", bad, first_m, first_n, got, want);
                printf("\n");
            }
        }
    }
    }
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dOut);
}

int main() {
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    run<128>((enc_fn)p);
    run<64>((enc_fn)p);
    return 0;
}
