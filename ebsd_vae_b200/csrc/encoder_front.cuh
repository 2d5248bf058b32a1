// K1 front end: conv0 (1 -> 32) + InstanceNorm + LeakyReLU + conv1 (32 -> 32) + 2x2 max-pool of a 128 x 128 uint8
// pattern in ONE kernel, both convolutions on tcgen05 (latice/model.py:93-98, 109-112).
//
// Why its own kernel.  In the generic block kernel (encoder_fused.cuh) conv0 ran on CUDA cores inside the producer
// warps (288 FMAs per position, 72 weights per thread in registers, two producer warps per scheduler) and a helper
// warp staged the input pixels with plain loads; with every role switched off that structure still took 0.9 ms per 1184
// patterns (tools/time_fused.py), more than the tensor work itself.  Here
//   * the uint8 patch of a work item (20 rows x 48 bytes) arrives by TMA, four items ahead (OOB pixels = 0 = padding);
//   * two builder warps expand it to an im2col operand: one 32-byte row per window position = the 9 taps as fp16
//     (integers 0..255 are exact) + 7 zeros;
//   * conv0 is 6 small tcgen05.mma per item (3 M-tiles of 128 positions x [fp16 hi, fp16 lo] of the weights, N = 32,
//     K = 16) into TMEM: the pixel is exact and hi + lo carries the weight to 2^-22, so conv0 is fp32-accurate; the
//     weights are pre-scaled by s0 / 255 (s0 a power of two) so that both halves sit in fp16's normal range;
//   * eight producer warps read conv0's output from TMEM (a lane = one window position, 32 channels in registers), apply
//     (x - mean) * rstd (conv0's plane statistics come from conv0_stats_u8_kernel: exact integer autocorrelation) and
//     LeakyReLU, and write conv1's operands in the swizzled K-major layouts;
//   * conv1 per tile and tap:  fp16 window (64-byte rows) x [fp16(w) ; fp16(4096 s (w - fp16 w))]   N = 64, K = 2 x 16
//                              e4m3(4096 (a - fp16 a)) window (32-byte rows) x e4m3(s w)             N = 32, K = 32
//     the second MMA accumulates onto columns 32..63 of the first (both carry the factor 4096 s), the epilogue adds
//     columns [0,32) + [32,64) / (4096 s).  Compared with the generic kernel's arithmetic the a * (w - fp16 w) term
//     stays in fp16 (it rides on the N = 64 MMA for free: a 128 x N x 16 MMA costs max(N/2, (4096 + 32 N)/128, 46)
//     cycles, tools/mma_rate.cu) and the fp8 window is half as wide: 142 instead of 184 cycles per tile and tap.
//   * the epilogue is the generic one for pooled 32-channel output (plane sums, 2x2 max-pool, TMA store).
//
// Geometry: a work item = 16 x 16 output pixels = two tiles of 16 rows x 8 columns; window 18 x 18 positions (1-pixel
// halo), im2col rows and TMEM lanes in window order.  tools/probe_umma32.cu verified 32-byte rows under SWIZZLE_32B with
// row-shifted descriptor starts and arbitrary 8-row-group strides for kind::f16 and kind::f8f6f4.
//
// Warp roles (640 threads): warps 0-7 producers, warps 8-15 epilogue (tile 0 / tile 1), warps 16-17 im2col builders,
// warp 18 TMA (weights once, patches) and TMEM allocation, warp 19 MMA issuer -- the scheduler picks the eligible warp
// with the HIGHEST id first (B300_MICROARCH.md), so the one thread that feeds the tensor pipe sits in the last warp.
#pragma once
#include "encoder_fused.cuh"

namespace ebsd {

struct FrontCfg {
    static constexpr int NT = 2;                      // tiles per work item
    static constexpr int PITCH = 8 * NT + 2, WIN_H = 18, WIN_POS = WIN_H * PITCH;   // 18 x 18 = 324 window positions
    static constexpr int A16_BYTES = (WIN_POS * 64 + 1023) / 1024 * 1024;            // fp16 window, 64-byte rows
    static constexpr int A8_BYTES = (WIN_POS * 32 + 1023) / 1024 * 1024;             // e4m3 residual window, 32-byte rows
    static constexpr int A_STAGE = A16_BYTES + A8_BYTES;
    static constexpr int A_STAGES = 3;
    static constexpr int W16_TAP = 64 * 64, W8_TAP = 32 * 32, W0_BYTES = 64 * 32;
    static constexpr int W16_OFF = 0, W8_OFF = 9 * W16_TAP, W0_OFF = W8_OFF + 9 * W8_TAP;
    static constexpr int W_BYTES = W0_OFF + W0_BYTES;                                 // 48128: one bulk copy
    static constexpr int W_REGION = (W_BYTES + 1023) / 1024 * 1024;
    static constexpr int MT = (WIN_POS + 127) / 128;                                  // conv0 M-tiles per item (3)
    static constexpr int IM_BYTES = MT * 128 * 32, IM_BUFS = 2;
    // uint8 patch of an item: rows y0 - 2 .. y0 + 17, columns x0 - 16 .. x0 + 31.  TMA wants the box to start at a
    // 16-byte aligned global address, i.e. at a pixel column that is a multiple of 16 (tools/probe_tma_u8.cu: any other
    // start raises "illegal instruction"), so the 2-pixel halo on the left costs a 16-pixel margin
    static constexpr int PATCH_W = 48, PATCH_X0 = 16, PATCH_H = WIN_H + 2, PATCH_BYTES = PATCH_W * PATCH_H, NPATCH = 4;
    static constexpr int PATCH_STRIDE = 1024;   // TMA destinations are 128-byte aligned
    static constexpr int WSTG = 1024;                                                 // per epilogue warp and tile
    static constexpr int STAGING = 4 * NT * WSTG;
    static constexpr int OFF_W = A_STAGES * A_STAGE;
    static constexpr int OFF_IM = OFF_W + W_REGION;
    static constexpr int OFF_PATCH = OFF_IM + IM_BUFS * IM_BYTES;
    static constexpr int OFF_STG = OFF_PATCH + NPATCH * PATCH_STRIDE;
    static constexpr int OFF_SCR = OFF_STG + STAGING;                                 // [epilogue warp][32 lanes][32 floats]: plane-sum scratch
    static constexpr int OFF_X = OFF_SCR + 8 * 4096;                                  // barriers (256 B) + table (256 B)
    static constexpr int SMEM_BYTES = 1024 + OFF_X + 1024;
    static constexpr int ACC_COLS = NT * 64;                                          // conv1 accumulators of one item
    static constexpr int C0_COL = 2 * ACC_COLS, C0_COLS = MT * 32;                    // conv0 output: 2 buffers x MT x 32 columns
    static constexpr int TMEM_COLS = 512;
    static constexpr int THREADS = 640, PRODUCER_WARPS = 8, EPILOGUE_WARPS = 8;
    static constexpr int ITEMS_X = 128 / (8 * NT), ITEMS_PER_IMAGE = (128 / 16) * ITEMS_X;   // 8 x 8
#ifndef EBSD_FRONT_ZGROUP
#define EBSD_FRONT_ZGROUP 4
#endif
    static constexpr int ZGROUP = EBSD_FRONT_ZGROUP;   // items whose per-lane plane sums are reduced together
    static_assert(ITEMS_PER_IMAGE % ZGROUP == 0, "plane-sum groups must not straddle images");
    static_assert(C0_COL + 2 * C0_COLS <= TMEM_COLS, "TMEM budget");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(A16_BYTES % 1024 == 0 && A8_BYTES % 1024 == 0 && W16_TAP % 1024 == 0 && W8_TAP % 1024 == 0 &&
                  W8_OFF % 1024 == 0 && W0_OFF % 1024 == 0 && IM_BYTES % 1024 == 0, "swizzle atoms need aligned regions");
};

struct FrontParams {
    const double *sums0;      // [nimg,32,2] plane sums of conv0's output (conv0_stats_u8_kernel)
    const uint8_t *weights;   // FrontCfg::W_BYTES: pre-swizzled shared-memory image (pack_front_weights_kernel)
    double *sums;             // out: [nimg,32,2] plane sums of conv1's un-pooled output, zero on entry
    float c0_inv_scale;       // 1 / s0: conv0's TMEM values are s0 x the true ones
    float corr_scale;         // 1 / (4096 s) of conv1's correction columns
    int nimg, nitems;
#if defined(EBSD_DEBUG_NOTRAP) || defined(EBSD_ROLE_PROFILE)
    int dbg;                  // role switches of the debugging / role-profiling builds (tools/debug_front.py, time_front.py)
#endif
};
#if defined(EBSD_DEBUG_NOTRAP) || defined(EBSD_ROLE_PROFILE)
#define FRONT_DBG(p, bit) (((p).dbg & (bit)) != 0)
#else
#define FRONT_DBG(p, bit) false
#endif

// K-major operand with 32-byte rows, SWIZZLE_32B; GROUP_ROWS = rows between the starts of consecutive 8-row groups
template <int GROUP_ROWS>
__device__ __forceinline__ uint64_t umma_smem_desc32(uint32_t saddr) {
    constexpr uint64_t sbo = ((uint64_t)GROUP_ROWS * 32) >> 4;
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (sbo << 32) | (1ull << 46) | (6ull << 61);
}
__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// y = leaky(x * scale + shift) as fp16 plus the e4m3 residual 4096 (y - fp16 y) of each value of the pair
__device__ __forceinline__ void norm_split_pair_front(float2 x, const float4 &tt, __half2 &h, uint32_t &r8) {
    const float2 y = __ffma2_rn(x, make_float2(tt.x, tt.y), make_float2(tt.z, tt.w));
    const float2 z = __fmul2_rn(y, make_float2(0.02f, 0.02f));
    const float2 a = make_float2(fmaxf(y.x, z.x), fmaxf(y.y, z.y));  // LeakyReLU(0.02)
    h = __floats2half2_rn(a.x, a.y);
    const float2 d = __ffma2_rn(__half22float2(h), make_float2(-kResidualScale, -kResidualScale),
                                __fmul2_rn(a, make_float2(kResidualScale, kResidualScale)));   // exact
    r8 = pack_e4m3x2(d.x, d.y);
}

// map_pat: uint8 patterns as (x, y, n) with box (48, 20, 1), no swizzle, OOB = 0;  map_out: as in the generic kernel,
// box = (16 channels, 4 x, 2 y, 1 n) of the pooled fp32 [nimg,64,64,32] output, 64-byte rows / SWIZZLE_64B.
__global__ void __launch_bounds__(FrontCfg::THREADS, 1)
front_u8_kernel(const __grid_constant__ CUtensorMap map_pat, const __grid_constant__ CUtensorMap map_out,
                const FrontParams p) {
    using C = FrontCfg;
    constexpr int COUT = 32;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_u = smem_u32(smem);
    uint8_t *extra = smem + C::OFF_X;
    uint64_t *a_full = (uint64_t *)extra;             // [A_STAGES] producers -> MMA (also: conv0's TMEM slot is read)
    uint64_t *a_empty = a_full + C::A_STAGES;         // [A_STAGES] MMA -> producers
    uint64_t *im_full = a_empty + C::A_STAGES;        // [2] builders -> MMA
    uint64_t *im_empty = im_full + 2;                 // [2] MMA -> builders
    uint64_t *patch_full = im_empty + 2;              // [NPATCH] TMA -> builders
    uint64_t *patch_empty = patch_full + C::NPATCH;   // [NPATCH] builders -> TMA
    uint64_t *tfull_bar = patch_empty + C::NPATCH;    // [2] MMA -> epilogue
    uint64_t *tempty_bar = tfull_bar + 2;             // [2] epilogue -> MMA
    uint64_t *c0_full = tempty_bar + 2;               // [2] conv0 of an item is in its TMEM buffer (item & 1)
    uint64_t *w_full = c0_full + 2;                   // [1] weights landed
    uint32_t *tmem_slot = (uint32_t *)(w_full + 1);
    static_assert((2 * C::A_STAGES + 4 + 2 * C::NPATCH + 4 + 3) * 8 + 4 <= 256, "barrier area");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    griddep_launch_dependents();   // the next block's CTAs may take SMs as soon as this grid's CTAs leave them

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::A_STAGES; ++s) {
            mbar_init(&a_full[s], C::PRODUCER_WARPS);
            mbar_init(&a_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&im_full[b], 2);
            mbar_init(&im_empty[b], 1);
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], C::EPILOGUE_WARPS);
        }
        for (int b = 0; b < C::NPATCH; ++b) {
            mbar_init(&patch_full[b], 1);
            mbar_init(&patch_empty[b], 2);
        }
        mbar_init(&c0_full[0], 1);
        mbar_init(&c0_full[1], 1);
        mbar_init(w_full, 1);
        mbar_fence_init();
        tma_prefetch_desc(&map_pat);
        tma_prefetch_desc(&map_out);
    }
    // im2col rows beyond the window (324 .. 383) are read by the last M-tile: keep them zero
    for (int i = threadIdx.x; i < C::IM_BUFS * (C::IM_BYTES - C::WIN_POS * 32) / 16; i += C::THREADS) {
        const int per = (C::IM_BYTES - C::WIN_POS * 32) / 16;
        const int b = i / per, r = i - b * per;
        sts128(smem_u + C::OFF_IM + b * C::IM_BYTES + C::WIN_POS * 32 + r * 16, make_uint4(0u, 0u, 0u, 0u));
    }
    fence_proxy_async();
    if (warp == 18) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // a multiple of ZGROUP items per CTA: the epilogue reduces its per-lane plane sums once per aligned group of ZGROUP
    // items, and which items share a group must not depend on where an image sits in the batch (ITEMS_PER_IMAGE = 64 is
    // a multiple as well, so a group never straddles an image or a CTA range)
    const int per_cta = ((p.nitems + (int)gridDim.x - 1) / (int)gridDim.x + C::ZGROUP - 1) / C::ZGROUP * C::ZGROUP;
    const int item_begin = (int)blockIdx.x * per_cta;
    const int item_end = item_begin + per_cta < p.nitems ? item_begin + per_cta : p.nitems;
    const int n_items = item_end > item_begin ? item_end - item_begin : 0;
    auto decode = [&](int item, int &n, int &y0, int &x0) {
        n = item / C::ITEMS_PER_IMAGE;
        const int r = item - n * C::ITEMS_PER_IMAGE;
        const int yb = r / C::ITEMS_X;
        y0 = yb * 16;
        x0 = (r - yb * C::ITEMS_X) * 8 * C::NT;
    };

    // Programmatic dependent launch: the prologue above and the weight image (written once at create time) overlap the
    // tail of the kernel in front (the conv0 statistics); nothing below the wait runs before that kernel has completed.
    if (warp == 18) {
        if (elect_one_sync() && n_items > 0) {
            if (FRONT_DBG(p, 8)) {
                mbar_arrive(w_full);
            } else {
                mbar_expect_tx(w_full, C::W_BYTES);
                bulk_load_1d(smem_u + C::OFF_W, p.weights, C::W_BYTES, w_full);
            }
        }
        __syncwarp();
    }
    griddep_wait();

    if (warp == 19) {
        // ===================== MMA issuer
        if (elect_one_sync() && n_items > 0) {
            constexpr uint32_t idesc64 = umma_idesc_f16(64), idesc32 = umma_idesc_f16(32);
            const uint32_t w16 = smem_u + C::OFF_W + C::W16_OFF, w8 = smem_u + C::OFF_W + C::W8_OFF,
                           w0 = smem_u + C::OFF_W + C::W0_OFF;
            // conv0 of item jj: im2col buffer and TMEM buffer jj & 1
            auto issue_conv0 = [&](int ib) {
                const uint32_t im = smem_u + C::OFF_IM + ib * C::IM_BYTES;
#pragma unroll
                for (int m = 0; m < (FRONT_DBG(p, 1) ? 0 : C::MT); ++m) {
                    const uint32_t d = tmem_base + (uint32_t)(C::C0_COL + ib * C::C0_COLS + 32 * m);
                    const uint64_t da = umma_smem_desc32<8>(im + m * 128 * 32);
                    umma_f16(d, da, umma_smem_desc32<8>(w0), idesc32, 0u);               // weights, fp16 hi
                    umma_f16(d, da, umma_smem_desc32<8>(w0 + 32 * 32), idesc32, 1u);     // weights, fp16 lo
                }
                umma_commit(&im_empty[ib]);
                umma_commit(&c0_full[ib]);
            };
            mbar_wait_bounded(w_full, 0);
            mbar_wait_bounded(&im_full[0], 0);
            tc_fence_after();
            issue_conv0(0);
            for (int j = 0; j < n_items; ++j) {
                // conv0 of the NEXT item goes in front of this item's conv1 MMAs and does not wait for this item's
                // window: its TMEM buffer was last read for item j - 1, whose a_full the previous iteration waited for.
                // (With one conv0 buffer every item paid the commit -> producers -> a_full -> issue round trip.)
                if (j + 1 < n_items) {
                    const int ib = (j + 1) & 1;
                    mbar_wait_bounded(&im_full[ib], ((unsigned)(j + 1) >> 1) & 1u);
                    tc_fence_after();
                    issue_conv0(ib);
                }
                const int sa = j % C::A_STAGES;
                mbar_wait_bounded(&a_full[sa], (unsigned)(j / C::A_STAGES) & 1u);
                tc_fence_after();
                const int buf = j & 1;
                mbar_wait_bounded(&tempty_bar[buf], (((unsigned)j >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_item = tmem_base + (uint32_t)(buf * C::ACC_COLS);
                const uint32_t win16 = smem_u + sa * C::A_STAGE, win8 = win16 + C::A16_BYTES;
#pragma unroll 1
                for (int tap = 0; tap < (FRONT_DBG(p, 2) ? 0 : 9); ++tap) {
                    const int dy = tap / 3, dx = tap - dy * 3;
                    const uint32_t shift = (uint32_t)(dy * C::PITCH + dx);
#pragma unroll
                    for (int t = 0; t < C::NT; ++t)
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            umma_f16(d_item + t * 64, umma_smem_desc_g<64, C::PITCH>(win16 + (shift + 8 * t) * 64 + k * 32),
                                     umma_smem_desc_g<64, 8>(w16 + tap * C::W16_TAP + k * 32), idesc64, (tap | k) == 0 ? 0u : 1u);
                }
#pragma unroll 1
                for (int tap = 0; tap < (FRONT_DBG(p, 4) ? 0 : 9); ++tap) {
                    const int dy = tap / 3, dx = tap - dy * 3;
                    const uint32_t shift = (uint32_t)(dy * C::PITCH + dx);
#pragma unroll
                    for (int t = 0; t < C::NT; ++t)
                        umma_f8(d_item + t * 64 + 32, umma_smem_desc32<C::PITCH>(win8 + (shift + 8 * t) * 32),
                                umma_smem_desc32<8>(w8 + tap * C::W8_TAP), idesc32, 1u);
                }
                umma_commit(&a_empty[sa]);
                umma_commit(&tfull_bar[buf]);
            }
        }
    } else if (warp == 16 || warp == 17) {
        // ===================== im2col builders: uint8 patch -> [position][9 taps as fp16, 7 zeros], SWIZZLE_32B rows
        const int bw = warp - 16;
        for (int j = 0; j < n_items; ++j) {
            const int pb = j % C::NPATCH, ib = j & 1;
            mbar_wait_bounded(&patch_full[pb], (unsigned)(j / C::NPATCH) & 1u);
            mbar_wait_bounded(&im_empty[ib], (((unsigned)j >> 1) & 1u) ^ 1u);
            const uint32_t patch = smem_u + C::OFF_PATCH + pb * C::PATCH_STRIDE;
            const uint32_t im = smem_u + C::OFF_IM + ib * C::IM_BYTES;
#pragma unroll 2
            for (int pos = bw * 32 + lane; pos < (FRONT_DBG(p, 128) ? 0 : C::WIN_POS); pos += 64) {
                const int wy = pos / C::PITCH, wx = pos - wy * C::PITCH;
                // window position (wy, wx) = image pixel (y0 - 1 + wy, x0 - 1 + wx); the patch starts at (y0 - 2, x0 - 16)
                uint32_t px[9];
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) px[dy * 3 + dx] = lds_u8(patch + (wy + dy) * C::PATCH_W + wx + dx + (C::PATCH_X0 - 2));
                __half2 h[5];
#pragma unroll
                for (int i = 0; i < 4; ++i) h[i] = __halves2half2(__uint2half_rn(px[2 * i]), __uint2half_rn(px[2 * i + 1]));
                h[4] = __halves2half2(__uint2half_rn(px[8]), __ushort_as_half((unsigned short)0));
                const uint32_t row = im + (uint32_t)pos * 32u;
                const uint32_t sw = ((row >> 7) & 1u) << 4;
                sts128(row + sw, make_uint4(*(const uint32_t *)&h[0], *(const uint32_t *)&h[1], *(const uint32_t *)&h[2],
                                            *(const uint32_t *)&h[3]));
                sts128(row + (sw ^ 16u), make_uint4(*(const uint32_t *)&h[4], 0u, 0u, 0u));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&im_full[ib]);
                mbar_arrive(&patch_empty[pb]);
            }
        }
    } else if (warp == 18) {
        // ===================== TMA: one uint8 patch per item, NPATCH items ahead (the weight image was requested above)
        if (elect_one_sync() && n_items > 0) {
            for (int j = 0; j < n_items; ++j) {
                const int pb = j % C::NPATCH;
                mbar_wait_bounded(&patch_empty[pb], ((unsigned)(j / C::NPATCH) & 1u) ^ 1u);
                int n, y0, x0;
                decode(item_begin + j, n, y0, x0);
                if (FRONT_DBG(p, 16)) {
                    mbar_arrive(&patch_full[pb]);
                    continue;
                }
                mbar_expect_tx(&patch_full[pb], C::PATCH_BYTES);
                tma_load_3d(smem_u + C::OFF_PATCH + pb * C::PATCH_STRIDE, &map_pat, x0 - C::PATCH_X0, y0 - 2, n, &patch_full[pb]);
            }
        }
    } else if (warp >= 8 && warp < 16) {
        // ===================== epilogue: TMEM -> plane sums of the un-pooled output, 2x2 max-pool, TMA store.
        // Eight warps: a warp reads the TMEM lane quarter warp & 3; warps 8-11 take channels 0..15 of both tiles of
        // every item, warps 12-15 channels 16..31.  (With four warps walking everything the epilogue was the longest
        // chain of the kernel: ncu showed the MMA issuer waiting for a free accumulator buffer.  Splitting by channel
        // half rather than by tile keeps ONE plane-sum reduction per warp and item.)
        const int quarter = warp & 3, hf = (warp >> 2) & 1;
        // lane = pixel (image row y0 + 4 quarter + (lane >> 3), column x0 + 8 t + (lane & 7)) of tile t
        const uint32_t stg_u32 = smem_u + C::OFF_STG + (uint32_t)((hf * 4 + quarter) * C::WSTG);   // [tile][8 px x 64 B]
        // scratch box of this warp (4 KiB): [32 rows][64 B] for the pooling, later [32 rows][128 B] for the plane sums
        const uint32_t scr = smem_u + C::OFF_SCR + (uint32_t)((warp - 8) * 4096);
        auto pool_pos = [](int l) { return (l & ~3) | ((l & 1) << 1) | ((l >> 1) & 1); };   // bits 0 and 1 swapped
        const uint32_t pool_wr = scr + (uint32_t)(pool_pos(lane) * 64);
        const uint32_t pool_wr_sw = (uint32_t)((pool_pos(lane) >> 1) & 3);
        const int pp = lane >> 2, cc = lane & 3;   // pooled pixel ([2 y][4 x] of the warp's 4 x 8 pixels) and chunk this lane produces
        uint32_t pool_rd[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int src = (2 * (pp >> 2) + (k >> 1)) * 8 + 2 * (pp & 3) + (k & 1);   // source lane = pixel (row, column)
            const int pos = pool_pos(src);
            pool_rd[k] = scr + (uint32_t)(pos * 64) + (uint32_t)((cc ^ ((pos >> 1) & 3)) << 4);
        }
        const uint32_t stg_off = (uint32_t)(pp * 64) + (uint32_t)((cc ^ ((pp >> 1) & 3)) << 4);
        // running plane sums as (hi, lo) fp32 pairs (PairSum, encoder_aux.cuh): lane c < 16 holds the sum of channel hf*16 + c, lane
        // 16 + c its sum of squares
        PairSum accum;
        accum.clear();
        int cur_n = -1;
        auto flush = [&]() {
            if (cur_n >= 0 && cur_n < p.nimg)
                atomicAdd(p.sums + ((long long)cur_n * COUT + hf * 16 + (lane & 15)) * 2 + (lane >> 4), accum.value());
            accum.clear();
        };
        float z[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) z[i] = 0.f;
        for (int j = 0; j < n_items; ++j) {
            const int buf = j & 1;
            int n, y0, x0;
            decode(item_begin + j, n, y0, x0);
            if (n != cur_n) {
                flush();
                cur_n = n;
            }
            mbar_wait_bounded(&tfull_bar[buf], ((unsigned)j >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * C::ACC_COLS + hf * 16);
            // both boxes of the previous item have been read by their TMA stores before they are overwritten
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            // z: [0,16) sums, [16,32) sums of squares of this warp's channels over its rows of the tiles of ZGROUP items
            if (((item_begin + j) % C::ZGROUP) == 0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) z[i] = 0.f;
            }
#pragma unroll 1
            for (int t = 0; t < (FRONT_DBG(p, 64) ? 0 : C::NT); ++t) {
                float v[16], w[16];
                tmem_ld16(t_row + t * 64, v);
                tmem_ld16(t_row + t * 64 + 32, w);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i] = fmaf(w[i], p.corr_scale, v[i]);
                    z[i] += v[i];
                    z[16 + i] = fmaf(v[i], v[i], z[16 + i]);
                }
                // 2x2 max-pool THROUGH SHARED MEMORY: every lane parks its pixel's 16 channels in the warp's scratch box
                // (4 STS.128), then lane (pooled pixel pp, 16-byte chunk cc) fetches that chunk of the four source pixels
                // (4 LDS.128), takes 12 maxima and stores straight into the TMA staging box.  The transposing shuffle
                // butterfly this replaces (12 SHFL + 24 selects + 12 maxima) cost twice the instructions; rows sit at
                // position P(lane) = lane with bits 0 and 1 swapped so that the two pooled pixels of a store phase and
                // of a load phase fall into different halves of the 32 banks.
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    sts128(pool_wr + (uint32_t)((c ^ pool_wr_sw) << 4),
                           make_uint4(__float_as_uint(v[4 * c]), __float_as_uint(v[4 * c + 1]), __float_as_uint(v[4 * c + 2]),
                                      __float_as_uint(v[4 * c + 3])));
                __syncwarp();
                float4 o = lds128(pool_rd[0]);
#pragma unroll
                for (int k = 1; k < 4; ++k) {
                    const float4 x = lds128(pool_rd[k]);
                    o.x = fmaxf(o.x, x.x);
                    o.y = fmaxf(o.y, x.y);
                    o.z = fmaxf(o.z, x.z);
                    o.w = fmaxf(o.w, x.w);
                }
                // one box per (tile, 16-channel half): 8 pooled pixels x 64 B, 64B-swizzled
                const uint32_t stg = stg_u32 + (uint32_t)(t * 512);
                sts128(stg + stg_off, make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w)));
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_4d(&map_out, stg, hf * 16, (x0 + 8 * t) >> 1, (y0 + 4 * quarter) >> 1, n);
                    bulk_commit();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);   // the accumulator is free before the (shuffle-heavy) reduction
            if (!FRONT_DBG(p, 64) && (((item_begin + j) % C::ZGROUP) == C::ZGROUP - 1 || j == n_items - 1)) {
                // lane c needs the sum of z[c] over the 32 lanes: through a swizzled scratch box (8 STS.128, then 32
                // conflict-free LDS.32 + 32 FADD, all independent) instead of a transposing shuffle reduction (31 SHFL +
                // ~120 ALU instructions in five dependent rounds -- shuffles share the shared-memory data path with the
                // tensor core's operand reads, which is what bounds this kernel)
                const uint32_t rowa = scr + (uint32_t)(lane * 128);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    sts128(rowa + (uint32_t)((i ^ (lane & 7)) << 4),
                           make_uint4(__float_as_uint(z[4 * i]), __float_as_uint(z[4 * i + 1]), __float_as_uint(z[4 * i + 2]),
                                      __float_as_uint(z[4 * i + 3])));
                __syncwarp();
                const uint32_t col = scr + (uint32_t)((lane & 3) << 2);
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int r = 0; r < 32; r += 2) {
                    s0 += lds32(col + (uint32_t)(r * 128) + (uint32_t)(((lane >> 2) ^ (r & 7)) << 4));
                    s1 += lds32(col + (uint32_t)((r + 1) * 128) + (uint32_t)(((lane >> 2) ^ ((r + 1) & 7)) << 4));
                }
                __syncwarp();   // the box is rewritten by the next item
                accum.add(s0 + s1);
            }
        }
        flush();
        if (lane == 0) bulk_wait_all();  // the staging buffers must outlive the TMA reads; stores complete before exit
    } else if (warp < 8) {
        // ===================== producers: conv0 output (TMEM) -> InstanceNorm + LeakyReLU -> conv1 operands (smem)
        // A warp owns a TMEM lane quarter and one 16-channel half: per item up to three units (M-tile m, its quarter,
        // its half) = 32 window positions x 16 channels, a lane = one position.  (scale, shift) of the half's channels
        // stay in registers for a whole image.
        const int pw = warp, quarter = pw & 3, half = pw >> 2;
        float4 tt[8];   // channel pair c of this half: (scale_a, scale_b, shift_a, shift_b)
        int tab_n = -1;
        auto update_table = [&](int n) {
            if (n == tab_n) return;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float sc[2], sh[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    // biased variance, eps = 1e-5 (torch instance_norm); TMEM holds s0 x conv0
                    const double *q = p.sums0 + ((long long)n * 32 + half * 16 + 2 * c + e) * 2;
                    const double mm = q[0] * (1.0 / 16384.0);
                    double var = q[1] * (1.0 / 16384.0) - mm * mm;
                    if (var < 0.0) var = 0.0;
                    const double rstd = 1.0 / sqrt(var + 1e-5);
                    sc[e] = (float)(rstd * (double)p.c0_inv_scale);
                    sh[e] = (float)(-mm * rstd);
                }
                tt[c] = make_float4(sc[0], sc[1], sh[0], sh[1]);
            }
            tab_n = n;
        };
        for (int j = 0; j < n_items; ++j) {
            int n, y0, x0;
            decode(item_begin + j, n, y0, x0);
            update_table(n);
            const int sa = j % C::A_STAGES, cb = j & 1;
            mbar_wait_bounded(&c0_full[cb], ((unsigned)j >> 1) & 1u);
            tc_fence_after();
            mbar_wait_bounded(&a_empty[sa], ((unsigned)(j / C::A_STAGES) & 1u) ^ 1u);
            const uint32_t win16 = smem_u + sa * C::A_STAGE, win8 = win16 + C::A16_BYTES;
#pragma unroll 1
            for (int m = 0; m < C::MT; ++m) {
                if (m * 128 + quarter * 32 >= C::WIN_POS || FRONT_DBG(p, 32)) continue;   // warp-uniform
                const int pos = m * 128 + quarter * 32 + lane;
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(C::C0_COL + cb * C::C0_COLS + 32 * m + 16 * half), v);
                tmem_ld_wait();
                if (pos < C::WIN_POS) {
                    const int wy = pos / C::PITCH, wx = pos - wy * C::PITCH;
                    const int y = y0 - 1 + wy, x = x0 - 1 + wx;
                    const bool inside = y >= 0 && y < 128 && x >= 0 && x < 128;   // outside: conv1's zero padding
                    __half2 h[8];
                    uint32_t r8[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        norm_split_pair_front(make_float2(v[2 * c], v[2 * c + 1]), tt[c], h[c], r8[c]);
                        if (!inside) {
                            h[c] = __halves2half2(__ushort_as_half((unsigned short)0), __ushort_as_half((unsigned short)0));
                            r8[c] = 0u;
                        }
                    }
                    const uint32_t row16 = win16 + (uint32_t)pos * 64u, row8 = win8 + (uint32_t)pos * 32u;
                    const uint32_t s16 = (row16 >> 7) & 3u, s8 = (row8 >> 7) & 1u;
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        sts128(row16 + (((uint32_t)(2 * half + c) ^ s16) << 4),
                               make_uint4(*(const uint32_t *)&h[4 * c], *(const uint32_t *)&h[4 * c + 1],
                                          *(const uint32_t *)&h[4 * c + 2], *(const uint32_t *)&h[4 * c + 3]));
                    sts128(row8 + (((uint32_t)half ^ s8) << 4),
                           make_uint4(r8[0] | (r8[1] << 16), r8[2] | (r8[3] << 16), r8[4] | (r8[5] << 16), r8[6] | (r8[7] << 16)));
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[sa]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 18) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// conv0 + conv1 weights of the front end as the shared-memory image front_u8_kernel copies in one piece:
//   [tap][64 rows x 64 B]  conv1: rows 0..31 fp16(w[co][ci]), rows 32..63 fp16(4096 s (w - fp16 w)); SWIZZLE_64B
//   [tap][32 rows x 32 B]  conv1: e4m3(s w[co][ci]); SWIZZLE_32B
//   [64 rows x 32 B]       conv0: rows 0..31 fp16 hi, 32..63 fp16 lo of s0 w0[c][tap] / 255, K = tap (9 of 16); SWIZZLE_32B
// one thread per 16-bit slot of the image
__global__ void pack_front_weights_kernel(const float *__restrict__ w0, const float *__restrict__ w1, float s0, float s1,
                                          uint8_t *__restrict__ out) {
    using C = FrontCfg;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // 16-bit slot
    if (i >= C::W_BYTES / 2) return;
    const int byte = i * 2;
    uint16_t val = 0;
    if (byte < C::W8_OFF) {
        const int tap = byte / C::W16_TAP, o = byte % C::W16_TAP;
        const int r = o / 64, pc = (o % 64) / 16, e = (o % 16) / 2;
        const int c = pc ^ ((r >> 1) & 3);          // logical chunk stored at physical chunk pc
        const int ci = c * 8 + e, co = r & 31;
        const float w = w1[((long long)co * 32 + ci) * 9 + tap];
        const __half hi = __float2half_rn(w);
        val = r < 32 ? __half_as_ushort(hi) : __half_as_ushort(__float2half_rn((w - __half2float(hi)) * (kResidualScale * s1)));
    } else if (byte < C::W0_OFF) {
        const int o0 = byte - C::W8_OFF;
        const int tap = o0 / C::W8_TAP, o = o0 % C::W8_TAP;
        const int r = o / 32, pc = (o % 32) / 16, e = o % 16;   // e even: two channels per slot
        const int c = pc ^ ((r >> 2) & 1);
        const int ci = c * 16 + e;
        const float wa = w1[((long long)r * 32 + ci) * 9 + tap], wb = w1[((long long)r * 32 + ci + 1) * 9 + tap];
        val = (uint16_t)pack_e4m3x2(wa * s1, wb * s1);
    } else {
        const int o = byte - C::W0_OFF;
        const int r = o / 32, pc = (o % 32) / 16, e = (o % 16) / 2;
        const int c = pc ^ ((r >> 2) & 1);
        const int k = c * 8 + e, ch = r & 31;
        if (k < 9) {
            const float w = w0[(long long)ch * 9 + k] * s0 / 255.0f;
            const __half hi = __float2half_rn(w);
            val = r < 32 ? __half_as_ushort(hi) : __half_as_ushort(__float2half_rn(w - __half2float(hi)));
        }
    }
    ((uint16_t *)out)[i] = val;
}

}  // namespace ebsd
