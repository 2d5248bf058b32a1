// K1 tensor-core path: 3x3 convolution as an implicit GEMM on tcgen05 (sm_100a), fed by TMA.
//
//   D[pixel, co] = sum_{tap, ci} A[pixel shifted by tap, ci] * W[tap, ci, co]
//
// * M = 128 output pixels per tile (a block of image rows), N = Cout, K = 9 * Cin.
// * Operands are fp16 with a three-term split so that the result carries fp32 accuracy
//   (a = a_hi + a_lo, w = w_hi + w_lo;  a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi; SURVEY appendix B:
//   latent error 6e-6 vs 5e-3 for single-pass fp16).  The weight halves are stacked along N:
//       MMA 1:  A = a_hi tile,  B = [w_hi ; w_lo]  (N = 2*Cout)  -> D[:, 0:Cout] and D[:, Cout:2*Cout]
//       MMA 2:  A = a_lo tile,  B =  w_hi          (N =   Cout)  -> D[:, 0:Cout]
//   and the epilogue adds the two column halves.  This reads the activation tile twice per K step
//   instead of three times (the A re-read is what limits small-N tcgen05.mma).
// * The activation tile for tap (dy,dx) is one 4-D TMA box (channels, x, y, image) at coordinates shifted by
//   (dx-1, dy-1); out-of-bounds coordinates are zero-filled by TMA, which IS the convolution's zero padding.
//   TMA writes the 128B/64B-swizzled K-major layout that the UMMA shared-memory descriptor expects.
// * Accumulators live in TMEM (2 buffers x 2*Cout fp32 columns): the epilogue of tile j overlaps the MMAs of
//   tile j+1.  Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
//   warps 4-7 = epilogue (one TMEM lane = one pixel per thread).
// * Epilogue: tcgen05.ld -> add halves -> raw fp32 NHWC store + per-(image, channel) sum / sum of squares
//   (warp butterfly, then one double atomicAdd per channel per warp) for InstanceNorm.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace ebsd {

// ---------------------------------------------------------------- configuration
template <int CIN, int COUT, int W>
struct MmaConvCfg {
    static constexpr int KC = CIN < 64 ? CIN : 64;       // channels per K block (one swizzle span)
    static constexpr int SWB = KC * 2;                   // bytes per operand row
    static constexpr int NCHUNK = CIN / KC;
    static constexpr int ITERS = 9 * NCHUNK;             // K blocks per tile
    static constexpr int KSTEPS = KC / 16;
    static constexpr int TW = W < 128 ? W : 128;         // tile = TW x TH pixels of TB images
    static constexpr int TH = (128 / TW) < W ? (128 / TW) : W;
    static constexpr int TB = 128 / (TW * TH);
    static constexpr int A_BYTES = 128 * SWB;
    static constexpr int B_BYTES = 2 * COUT * SWB;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + B_BYTES;
    static constexpr int STAGES_FIT = (192 * 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
    static constexpr int TMEM_COLS = 4 * COUT;           // 2 buffers x (hi|lo halves)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int THREADS = 256;
    static_assert(TW * TH * TB == 128, "tile must hold 128 pixels");
    static_assert(TMEM_COLS >= 32 && TMEM_COLS <= 512, "TMEM budget");
};

struct MmaConvParams {
    float *raw;      // [B,H,W,COUT] fp32
    double *sums;    // [B,COUT,2]
    int nimg;
    int ntiles;
};

// Sum each of the 32 values over the 32 lanes; lane c returns the total of v[c].
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int step = 16, n = 32; step >= 1; step >>= 1, n >>= 1) {
        const bool upper = (lane & step) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = upper ? v[i] : v[i + n / 2];
            const float keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
    }
    return v[0];
}

template <int CIN, int COUT, int W>
__global__ void __launch_bounds__(256, 1)
conv3x3_mma_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                   const __grid_constant__ CUtensorMap map_w, const MmaConvParams p) {
    using C = MmaConvCfg<CIN, COUT, W>;
    constexpr int H = W;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t *full_bar = bars;
    uint64_t *empty_bar = bars + C::STAGES;
    uint64_t *tfull_bar = bars + 2 * C::STAGES;
    uint64_t *tempty_bar = tfull_bar + 2;
    uint32_t *tmem_slot = (uint32_t *)(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], 4);
        }
        mbar_fence_init();
        tma_prefetch_desc(&map_hi);
        tma_prefetch_desc(&map_lo);
        tma_prefetch_desc(&map_w);
    }
    if (warp == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    constexpr int tiles_per_img = (H * W) / 128;  // 0 when an image has only 64 pixels (W = 8): TB = 2

    if (warp == 0) {
        // ===================== TMA producer
        if (lane == 0) {
            unsigned it = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                int b0, y0;
                if (tiles_per_img > 0) {
                    b0 = tile / (tiles_per_img > 0 ? tiles_per_img : 1);
                    y0 = (tile - b0 * tiles_per_img) * C::TH;
                } else {
                    b0 = tile * C::TB;
                    y0 = 0;
                }
                for (int kb = 0; kb < C::ITERS; ++kb, ++it) {
                    const int s = it % C::STAGES;
                    const unsigned ph = (it / C::STAGES) & 1u;
                    const int tap = kb / C::NCHUNK, cc = kb - tap * C::NCHUNK;
                    const int dy = tap / 3, dx = tap - dy * 3;
                    mbar_wait_bounded(&empty_bar[s], ph ^ 1u);
                    uint8_t *st = smem + s * C::STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], C::STAGE_BYTES);
                    tma_load_4d(st, &map_hi, cc * C::KC, dx - 1, y0 + dy - 1, b0, &full_bar[s]);
                    tma_load_4d(st + C::A_BYTES, &map_lo, cc * C::KC, dx - 1, y0 + dy - 1, b0, &full_bar[s]);
                    tma_load_2d(st + 2 * C::A_BYTES, &map_w, 0, kb * 2 * COUT, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc_n2 = umma_idesc_f16(2 * COUT);
            constexpr uint32_t idesc_n1 = umma_idesc_f16(COUT);
            unsigned it = 0;
            int j = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++j) {
                const int buf = j & 1;
                mbar_wait_bounded(&tempty_bar[buf], (((unsigned)j >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 2 * COUT);
                for (int kb = 0; kb < C::ITERS; ++kb, ++it) {
                    const int s = it % C::STAGES;
                    const unsigned ph = (it / C::STAGES) & 1u;
                    mbar_wait_bounded(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + s * C::STAGE_BYTES);
                    const uint32_t a_lo = a_hi + C::A_BYTES;
                    const uint32_t b_w = a_hi + 2 * C::A_BYTES;
#pragma unroll
                    for (int k = 0; k < C::KSTEPS; ++k) {
                        const uint64_t dh = umma_smem_desc<C::SWB>(a_hi + k * 32);
                        const uint64_t dl = umma_smem_desc<C::SWB>(a_lo + k * 32);
                        const uint64_t db = umma_smem_desc<C::SWB>(b_w + k * 32);
                        umma_f16(d_tmem, dh, db, idesc_n2, (kb | k) != 0 ? 1u : 0u);
                        umma_f16(d_tmem, dl, db, idesc_n1, 1u);
                    }
                    umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
                }
                umma_commit(&tfull_bar[buf]);    // accumulator complete
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM -> registers -> raw store + plane statistics
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;  // pixel index inside the tile = TMEM lane
        int j = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++j) {
            const int buf = j & 1;
            int b0, y0;
            if (tiles_per_img > 0) {
                b0 = tile / (tiles_per_img > 0 ? tiles_per_img : 1);
                y0 = (tile - b0 * tiles_per_img) * C::TH;
            } else {
                b0 = tile * C::TB;
                y0 = 0;
            }
            const int x = m % C::TW;
            const int y = y0 + (m / C::TW) % C::TH;
            const int b = b0 + m / (C::TW * C::TH);
            const bool live = b < p.nimg;  // warp-uniform: a warp's 32 pixels belong to one image
            mbar_wait_bounded(&tfull_bar[buf], ((unsigned)j >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 2 * COUT);
            float *out = p.raw + (((long long)b * H + y) * W + x) * COUT;
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += 32) {
                float v[32], w[32];
                tmem_ld32(t_row + c0, v);
                tmem_ld32(t_row + COUT + c0, w);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += w[i];
                if (live) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *(float4 *)(out + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) w[i] = v[i] * v[i];
                const float s1 = warp_transpose_reduce32(v, lane);
                const float s2 = warp_transpose_reduce32(w, lane);
                if (live) {
                    double *dst = p.sums + ((long long)b * COUT + c0 + lane) * 2;
                    atomicAdd(dst, (double)s1);
                    atomicAdd(dst + 1, (double)s2);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Finisher for the tensor-core path: y = leaky((x - mean) * rstd), optional 2x2 max-pool, then the fp16
// hi / lo split written as two NHWC planes (the next layer's TMA source).
// ---------------------------------------------------------------------------------------------
template <int CH, bool POOL>
__global__ void __launch_bounds__(256) finish_split_kernel(const float *__restrict__ raw,
                                                           const double *__restrict__ sums, __half *__restrict__ hi,
                                                           __half *__restrict__ lo, int H, int W, long long B) {
    const int Ho = POOL ? H / 2 : H, Wo = POOL ? W / 2 : W;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = B * Ho * Wo * (CH / 4);
    if (gid >= total) return;
    const int c4 = (int)(gid % (CH / 4));
    long long r = gid / (CH / 4);
    const int xo = (int)(r % Wo);
    r /= Wo;
    const int yo = (int)(r % Ho);
    const long long n = r / Ho;
    const double inv_hw = 1.0 / ((double)H * (double)W);
    float mean[4], rstd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double s = sums[(n * CH + c4 * 4 + j) * 2 + 0], ss = sums[(n * CH + c4 * 4 + j) * 2 + 1];
        const double mm = s * inv_hw;
        double var = ss * inv_hw - mm * mm;
        if (var < 0.0) var = 0.0;
        mean[j] = (float)mm;
        rstd[j] = (float)(1.0 / sqrt(var + 1e-5));
    }
    float4 v;
    if (POOL) {
        const float *q = raw + ((n * H + yo * 2) * W + xo * 2) * CH + c4 * 4;
        const float4 a = *(const float4 *)q, b = *(const float4 *)(q + CH);
        const float4 c = *(const float4 *)(q + (long long)W * CH), d = *(const float4 *)(q + (long long)W * CH + CH);
        v.x = fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x));
        v.y = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
        v.z = fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z));
        v.w = fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w));
    } else {
        v = *(const float4 *)(raw + ((n * H + yo) * W + xo) * CH + c4 * 4);
    }
    float o[4] = {v.x, v.y, v.z, v.w};
    __half h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float t = (o[j] - mean[j]) * rstd[j];
        t = t >= 0.f ? t : t * 0.02f;
        h[j] = __float2half_rn(t);
        l[j] = __float2half_rn(t - __half2float(h[j]));
    }
    const long long off = ((n * Ho + yo) * Wo + xo) * CH + c4 * 4;
    *(uint2 *)(hi + off) = *(const uint2 *)h;
    *(uint2 *)(lo + off) = *(const uint2 *)l;
}

// Weight packing for the tensor path: torch [Cout,Cin,3,3] fp32 ->
//   rows (kb*2*Cout + r), kb = tap*NCHUNK + chunk;  r < Cout: fp16 hi of w[co=r], r >= Cout: fp16 lo of w[co=r-Cout];
//   each row holds KC channels (K-major).
__global__ void pack_conv_weights_mma_kernel(const float *__restrict__ w, __half *__restrict__ out, int cin, int cout,
                                             int kc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nchunk = cin / kc;
    const int total = 9 * nchunk * 2 * cout * kc;
    if (i >= total) return;
    const int k = i % kc;
    const int r = (i / kc) % (2 * cout);
    const int kb = i / (kc * 2 * cout);
    const int tap = kb / nchunk, cc = kb % nchunk;
    const int co = r < cout ? r : r - cout;
    const float val = w[((long long)co * cin + cc * kc + k) * 9 + tap];
    const __half h = __float2half_rn(val);
    out[i] = r < cout ? h : __float2half_rn(val - __half2float(h));
}

}  // namespace ebsd
