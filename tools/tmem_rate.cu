// Microbenchmark: how fast can the epilogue warps of one CTA drain TMEM (tcgen05.ld 32x32b) on B200?
// The top-k screen (csrc/topk_screen.cuh) reads one fp32 per (query, row) pair, 128 lanes x 256 columns = 128 KiB per
// tile, so its floor is this number.  Variants: loads of 16/32/64/128 columns, one or two loads in flight per warp,
// 4/8/16 warps (every warp reads its own lane quarter = warp & 3; warps of a quarter split the columns), with and
// without a concurrent MMA stream (3 x M128 N256 K16 per 256 columns read, like the screen) writing the other half.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tmem_rate tmem_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (((8ull * 64) >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void mma256(uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
    const uint32_t id = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}"
                 ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}

template <int X>
struct Ld;
#define LD_BODY(X, REGS_OUT, ...)                                                                     \
    template <> struct Ld<X> {                                                                         \
        static __device__ __forceinline__ void go(uint32_t taddr, uint32_t *r) {                       \
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x" #X ".b32 {" REGS_OUT "}, [%" #X "];"       \
                         : __VA_ARGS__ : "r"(taddr) : "memory");                                       \
        }                                                                                              \
    };
#define R4(b) "=r"(r[b]), "=r"(r[b + 1]), "=r"(r[b + 2]), "=r"(r[b + 3])
#define R16(b) R4(b), R4(b + 4), R4(b + 8), R4(b + 12)
#define R64(b) R16(b), R16(b + 16), R16(b + 32), R16(b + 48)
LD_BODY(16, "%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15", R16(0))
LD_BODY(32, "%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31",
        R16(0), R16(16))
LD_BODY(64, "%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
            "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63",
        R64(0))

__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Every loader warp reads `cols_per_rep` columns of its lane quarter per repetition, X columns per load, INFLIGHT loads
// between waits, and folds the values with fmax (the screen's chunk_max) so nothing is optimised away.
template <int X, int INFLIGHT>
__global__ void __launch_bounds__(544, 1) drain(int groups, int reps, int with_mma, long long *out, float *sink) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t mbar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5;
    const int nload = 4 * groups;   // loader warps: 0 .. nload-1; warp nload = MMA issuer
    for (int i = threadIdx.x; i < 32 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0u;
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
    const long long t0 = clock64();
    float m = -1e30f;
    if (warp < nload) {
        const int quarter = warp & 3, grp = warp >> 2;
        const int gcols = 256 / groups;   // this warp's share of a 256-column accumulator
        const uint32_t t_row = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(grp * gcols);
        for (int r = 0; r < reps; ++r) {
            for (int c0 = 0; c0 < gcols; c0 += X * INFLIGHT) {
                uint32_t v[X * INFLIGHT];
#pragma unroll
                for (int u = 0; u < INFLIGHT; ++u) Ld<X>::go(t_row + c0 + u * X, v + u * X);
                ld_wait();
                float a = __uint_as_float(v[0]), b = __uint_as_float(v[1]), c = __uint_as_float(v[2]), d = __uint_as_float(v[3]);
#pragma unroll
                for (int i = 4; i < X * INFLIGHT; i += 4) {
                    a = fmaxf(a, __uint_as_float(v[i]));
                    b = fmaxf(b, __uint_as_float(v[i + 1]));
                    c = fmaxf(c, __uint_as_float(v[i + 2]));
                    d = fmaxf(d, __uint_as_float(v[i + 3]));
                }
                m = fmaxf(m, fmaxf(fmaxf(a, b), fmaxf(c, d)));
            }
        }
    } else if (warp == nload && with_mma) {
        if ((threadIdx.x & 31) == 0) {
            const uint64_t da = desc(s32(smem)), db = desc(s32(smem) + 8192);
            for (int r = 0; r < reps; ++r) {
                mma256(tmem + 256, da, db, 0u);
                mma256(tmem + 256, da, db + 2, 1u);
                mma256(tmem + 256, da + 2, db, 1u);
                if ((r & 7) == 7 || r == reps - 1) {
                    // keep the MMA stream roughly in step with the loaders instead of queueing everything up front
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mbar)) : "memory");
                    uint32_t ok = 0;
                    const uint32_t par = (uint32_t)((r >> 3) & 1);
                    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(&mbar)), "r"(par));
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    const long long t1 = clock64();
    if (m == 12345.f) sink[threadIdx.x] = m;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int X, int INFLIGHT>
static void run(int groups, int with_mma, long long *d, float *sink) {
    if (X * INFLIGHT > 256 / groups) return;
    const int reps = 2048;
    const int threads = 32 * (4 * groups + 1);
    long long h;
    CK(cudaFuncSetAttribute(drain<X, INFLIGHT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
    for (int i = 0; i < 2; ++i) {
        drain<X, INFLIGHT><<<148, threads, 40 * 1024>>>(groups, reps, with_mma, d, sink);
        CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
    printf("x%-3d inflight %d  loader warps %2d  mma %d : %8.1f cycles per 128x256 fp32 tile  (%5.1f B/clk/SM)\n", X, INFLIGHT,
           4 * groups, with_mma, (double)h / reps, 131072.0 * reps / (double)h);
}

int main() {
    long long *d;
    float *sink;
    CK(cudaMalloc(&d, 8));
    CK(cudaMalloc(&sink, 4096));
    for (int with_mma = 0; with_mma < 2; ++with_mma)
        for (int groups : {1, 2, 4}) {
            run<16, 1>(groups, with_mma, d, sink);
            run<16, 2>(groups, with_mma, d, sink);
            run<16, 4>(groups, with_mma, d, sink);
            run<32, 1>(groups, with_mma, d, sink);
            run<32, 2>(groups, with_mma, d, sink);
            run<64, 1>(groups, with_mma, d, sink);
        }
    cudaFree(d);
    return 0;
}
