// Probe: 3-D TMA load of a uint8 patch (box W x H x 1) from [n][128][128] patterns with out-of-bounds start coordinates.
// usage: probe_tma_u8 BOXW BOXH L2PROMO(0..3) X Y   -- prints the checksum of the patch or the CUDA error
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_tma_u8 probe_tma_u8.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, int x, int y, int n, int bytes, uint8_t *out) {
    __shared__ __align__(1024) uint8_t buf[4096];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(s32(buf)), "l"(&map), "r"(x), "r"(y), "r"(n), "r"(s32(&bar)) : "memory");
        uint32_t ok = 0;
        long long t0 = clock64();
        while (!ok && clock64() - t0 < 1000000000ll)
            asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(&bar)));
        if (!ok) printf("timeout\n");
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
    const int bw = atoi(argv[1]), bh = atoi(argv[2]), promo = atoi(argv[3]), x = atoi(argv[4]), y = atoi(argv[5]);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    const int nimg = 3;
    std::vector<uint8_t> h(nimg * 16384);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(1 + (i % 251));
    uint8_t *d, *dout;
    CK(cudaMalloc(&d, h.size())); CK(cudaMalloc(&dout, 4096));
    CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    CUtensorMap m;
    cuuint64_t gd[3] = {128, 128, (cuuint64_t)nimg}, gs[2] = {128, 16384};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
    CUresult cr = ((enc_fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr) { printf("encode failed %d\n", (int)cr); return 1; }
    k<<<1, 128>>>(m, x, y, 1, bw * bh, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("box %dx%d promo %d at (%d,%d): %s\n", bw, bh, promo, x, y, cudaGetErrorString(e)); return 2; }
    std::vector<uint8_t> o(bw * bh);
    CK(cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r = 0; r < bh; ++r) for (int c = 0; c < bw; ++c) {
        const int yy = y + r, xx = x + c;
        const uint8_t want = (yy < 0 || yy >= 128 || xx < 0 || xx >= 128) ? 0 : h[16384 + yy * 128 + xx];
        if (o[r * bw + c] != want) ++bad;
    }
    printf("box %dx%d promo %d at (%d,%d): ok, %d mismatches\n", bw, bh, promo, x, y, bad);
    return 0;
}
