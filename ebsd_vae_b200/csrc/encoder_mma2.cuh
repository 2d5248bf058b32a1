// K1 tensor-core path, second generation: the convolution as nine row-SHIFTED views of ONE shared-memory window.
//
// Activations are stored zero-padded and flattened: image n is a (H+2) x (W+2) grid of pixels, each pixel a row of
// Cin fp16 values (two planes: hi and lo of the fp16 split, see encoder_mma.cuh).  With the padding in memory,
// output pixel P (flattened padded index) needs input pixels P + (dy-1)*(W+2) + (dx-1): every tap is the same
// 1-D stream shifted by a constant.  A tile is 128 consecutive padded positions; its inputs are the window of
// WIN = 130 + 2*(W+2) rows around it, loaded ONCE per K chunk by TMA (instead of nine times, one per tap, as in
// the first-generation kernel, which was L2-bandwidth bound).  The nine A operands are UMMA descriptors whose start
// address is the window base plus (dy*(W+2) + dx) rows -- tools/probe_umma_desc.cu shows that tcgen05.mma applies
// the 128B/64B swizzle to absolute shared-memory address bits, so row-shifted descriptors with base_offset = 0 read
// exactly what TMA wrote.  Outputs at padding positions are computed and discarded (3-6 % of the work for the
// large layers; the small late layers keep the first-generation kernel).
//
// Weights: small layers keep all nine taps resident in shared memory for the lifetime of the (persistent) CTA;
// large layers stream [w_hi; w_lo] tiles through a second ring.
//
// Everything else (three-term split stacked along N, TMEM double buffering, warp roles, fused plane statistics)
// is as in encoder_mma.cuh.
#pragma once
#include "encoder_mma.cuh"

namespace ebsd {

template <int CIN, int COUT, int W>
struct Mma2Cfg {
    static constexpr int H = W;
    static constexpr int WP = W + 2;
    static constexpr int HP = H + 2;
    static constexpr int KC = CIN < 64 ? CIN : 64;
    static constexpr int ROWB = KC * 2;                      // bytes per window row (one K chunk of one pixel)
    static constexpr int NCHUNK = CIN / KC;
    static constexpr int KSTEPS = KC / 16;
    static constexpr int WIN = 130 + 2 * WP;                 // rows a tile can touch
    static constexpr int NBOX = (WIN + 255) / 256;
    static constexpr int BOXR = (((WIN + NBOX - 1) / NBOX) + 7) / 8 * 8;   // TMA box rows (multiple of 8, <= 256)
    static constexpr int WINR = NBOX * BOXR;
    static constexpr int A_PLANE_BYTES = (WINR * ROWB + 1023) / 1024 * 1024;
    static constexpr int A_STAGE_BYTES = 2 * A_PLANE_BYTES;  // hi + lo
    static constexpr int A_STAGES = 2;
    static constexpr int B_TILE_BYTES = 2 * COUT * ROWB;     // [w_hi; w_lo] for one (tap, chunk)
    static constexpr int B_TOTAL_BYTES = 9 * NCHUNK * B_TILE_BYTES;
    static constexpr bool RESIDENT_B = B_TOTAL_BYTES <= 80 * 1024;
    static constexpr int B_STAGES_FIT = (200 * 1024 - A_STAGES * A_STAGE_BYTES) / B_TILE_BYTES;
    static constexpr int B_STAGES = RESIDENT_B ? 9 * NCHUNK : (B_STAGES_FIT > 6 ? 6 : B_STAGES_FIT);
    static constexpr int B_BYTES = B_STAGES * B_TILE_BYTES;
    static constexpr int TMEM_COLS = 4 * COUT;
    static constexpr int SMEM_BYTES = A_STAGES * A_STAGE_BYTES + B_BYTES + 1024 + 512;
    static constexpr int THREADS = 256;
    static_assert(BOXR <= 256, "TMA box too tall");
    static_assert(RESIDENT_B || B_STAGES >= 2, "not enough shared memory for the weight ring");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct Mma2Params {
    float *raw;      // [nimg,H,W,COUT] fp32, un-padded
    double *sums;    // [nimg,COUT,2]
    int nimg;
    int ntiles;      // ceil(nimg*HP*WP / 128)
    int dbg;         // profiling switches (ebsd_debug_set_flags): 1 no atomics, 2 no raw store, 4 no statistics,
                     // 8 no TMEM loads, 16 no MMAs
};

template <int CIN, int COUT, int W>
__global__ void __launch_bounds__(256, 1)
conv3x3_mma2_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                    const __grid_constant__ CUtensorMap map_w, const Mma2Params p) {
    using C = Mma2Cfg<CIN, COUT, W>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *smem_b = smem + C::A_STAGES * C::A_STAGE_BYTES;
    uint64_t *bars = (uint64_t *)(smem_b + C::B_BYTES);
    uint64_t *a_full = bars;                       // [A_STAGES]
    uint64_t *a_empty = a_full + C::A_STAGES;      // [A_STAGES]
    uint64_t *b_full = a_empty + C::A_STAGES;      // [B_STAGES]
    uint64_t *b_empty = b_full + C::B_STAGES;      // [B_STAGES]
    uint64_t *tfull_bar = b_empty + C::B_STAGES;   // [2]
    uint64_t *tempty_bar = tfull_bar + 2;          // [2]
    uint32_t *tmem_slot = (uint32_t *)(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::A_STAGES; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < C::B_STAGES; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
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

    // Each CTA owns a CONTIGUOUS range of tiles: consecutive tiles are consecutive image rows, so the plane
    // statistics can be accumulated in registers and flushed once per image instead of once per tile.
    const int tiles_per_cta = (p.ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int tile_begin = (int)blockIdx.x * tiles_per_cta;
    const int tile_end = tile_begin + tiles_per_cta < p.ntiles ? tile_begin + tiles_per_cta : p.ntiles;

    if (warp == 0) {
        // ===================== TMA producer
        if (lane == 0) {
            if (C::RESIDENT_B) {
                for (int kb = 0; kb < 9 * C::NCHUNK; ++kb) {
                    mbar_expect_tx(&b_full[kb], C::B_TILE_BYTES);
                    tma_load_2d(smem_b + kb * C::B_TILE_BYTES, &map_w, 0, kb * 2 * COUT, &b_full[kb]);
                }
            }
            unsigned ait = 0, bit = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int row0 = tile * 128 - C::WP - 1;  // first padded position of the window (may be negative)
                for (int cc = 0; cc < C::NCHUNK; ++cc, ++ait) {
                    const int sa = ait % C::A_STAGES;
                    mbar_wait_bounded(&a_empty[sa], ((ait / C::A_STAGES) & 1u) ^ 1u);
                    uint8_t *st = smem + sa * C::A_STAGE_BYTES;
                    mbar_expect_tx(&a_full[sa], 2 * C::NBOX * C::BOXR * C::ROWB);
#pragma unroll
                    for (int bx = 0; bx < C::NBOX; ++bx) {
                        tma_load_2d(st + bx * C::BOXR * C::ROWB, &map_hi, cc * C::KC, row0 + bx * C::BOXR, &a_full[sa]);
                        tma_load_2d(st + C::A_PLANE_BYTES + bx * C::BOXR * C::ROWB, &map_lo, cc * C::KC,
                                    row0 + bx * C::BOXR, &a_full[sa]);
                    }
                    if (!C::RESIDENT_B) {
                        for (int tap = 0; tap < 9; ++tap, ++bit) {
                            const int sb = bit % C::B_STAGES;
                            mbar_wait_bounded(&b_empty[sb], ((bit / C::B_STAGES) & 1u) ^ 1u);
                            mbar_expect_tx(&b_full[sb], C::B_TILE_BYTES);
                            tma_load_2d(smem_b + sb * C::B_TILE_BYTES, &map_w, 0, (tap * C::NCHUNK + cc) * 2 * COUT,
                                        &b_full[sb]);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc_n2 = umma_idesc_f16(2 * COUT);
            constexpr uint32_t idesc_n1 = umma_idesc_f16(COUT);
            if (C::RESIDENT_B) {
                for (int kb = 0; kb < 9 * C::NCHUNK; ++kb) mbar_wait_bounded(&b_full[kb], 0);
            }
            unsigned ait = 0, bit = 0;
            int j = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile, ++j) {
                const int buf = j & 1;
                mbar_wait_bounded(&tempty_bar[buf], (((unsigned)j >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 2 * COUT);
                for (int cc = 0; cc < C::NCHUNK; ++cc, ++ait) {
                    const int sa = ait % C::A_STAGES;
                    mbar_wait_bounded(&a_full[sa], (ait / C::A_STAGES) & 1u);
                    tc_fence_after();
                    const uint32_t win_hi = smem_u32(smem + sa * C::A_STAGE_BYTES);
                    const uint32_t win_lo = win_hi + C::A_PLANE_BYTES;
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        const int dy = tap / 3, dx = tap - dy * 3;
                        const uint32_t shift = (uint32_t)((dy * C::WP + dx) * C::ROWB);
                        uint32_t b_w;
                        int sb = 0;
                        if (C::RESIDENT_B) {
                            b_w = smem_u32(smem_b + (tap * C::NCHUNK + cc) * C::B_TILE_BYTES);
                        } else {
                            sb = bit % C::B_STAGES;
                            mbar_wait_bounded(&b_full[sb], (bit / C::B_STAGES) & 1u);
                            tc_fence_after();
                            b_w = smem_u32(smem_b + sb * C::B_TILE_BYTES);
                            ++bit;
                        }
                        // same-shape MMAs back to back: switching the instruction descriptor between consecutive
                        // tcgen05.mma costs a pipeline drain (measured ~150 cycles per switch)
#pragma unroll
                        for (int k = 0; k < C::KSTEPS; ++k) {
                            const uint64_t dh = umma_smem_desc<C::ROWB>(win_hi + shift + k * 32);
                            const uint64_t db = umma_smem_desc<C::ROWB>(b_w + k * 32);
                            if (!(p.dbg & 16)) umma_f16(d_tmem, dh, db, idesc_n2, (cc | tap | k) != 0 ? 1u : 0u);
                        }
#pragma unroll
                        for (int k = 0; k < C::KSTEPS; ++k) {
                            const uint64_t dl = umma_smem_desc<C::ROWB>(win_lo + shift + k * 32);
                            const uint64_t db = umma_smem_desc<C::ROWB>(b_w + k * 32);
                            if (!(p.dbg & 16)) umma_f16(d_tmem, dl, db, idesc_n1, 1u);
                        }
                        if (!C::RESIDENT_B) umma_commit(&b_empty[sb]);
                    }
                    umma_commit(&a_empty[sa]);
                }
                umma_commit(&tfull_bar[buf]);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        constexpr int NCB = COUT / 32;
        float acc1[NCB], acc2[NCB];  // running sum / sum of squares of channel cb*32 + lane for image cur_n
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) acc1[cb] = acc2[cb] = 0.f;
        int cur_n = -1;
        auto flush = [&]() {
            if (cur_n >= 0 && cur_n < p.nimg && !(p.dbg & 1)) {
#pragma unroll
                for (int cb = 0; cb < NCB; ++cb) {
                    double *dst = p.sums + ((long long)cur_n * COUT + cb * 32 + lane) * 2;
                    atomicAdd(dst, (double)acc1[cb]);
                    atomicAdd(dst + 1, (double)acc2[cb]);
                }
            }
#pragma unroll
            for (int cb = 0; cb < NCB; ++cb) acc1[cb] = acc2[cb] = 0.f;
        };
        int j = 0;
        for (int tile = tile_begin; tile < tile_end; ++tile, ++j) {
            const int buf = j & 1;
            const int pos = tile * 128 + m;  // flattened padded output position
            const int n = pos / (C::HP * C::WP);
            const int rem = pos - n * (C::HP * C::WP);
            const int yp = rem / C::WP, xp = rem - yp * C::WP;
            const bool interior = n < p.nimg && yp >= 1 && yp <= C::H && xp >= 1 && xp <= W;
            const int n_first = __shfl_sync(0xffffffffu, n, 0);
            const int n_last = __shfl_sync(0xffffffffu, n, 31);
            const bool straddle = n_first != n_last;  // the warp's 32 positions span two images
            if (straddle || n_first != cur_n) {
                flush();
                cur_n = straddle ? -1 : n_first;
            }
            mbar_wait_bounded(&tfull_bar[buf], ((unsigned)j >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 2 * COUT);
            float *out = p.raw + (((long long)n * C::H + (yp - 1)) * W + (xp - 1)) * COUT;
#pragma unroll
            for (int cb = 0; cb < NCB; ++cb) {
                const int c0 = cb * 32;
                float v[32], w[32];
                if (!(p.dbg & 8)) {
                    tmem_ld32(t_row + c0, v);
                    tmem_ld32(t_row + COUT + c0, w);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = w[i] = (float)(i + lane);
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = interior ? v[i] + w[i] : 0.f;
                if (interior && !(p.dbg & 2)) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *(float4 *)(out + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
                if (p.dbg & 4) continue;
                if (!straddle) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) w[i] = v[i] * v[i];
                    acc1[cb] += warp_transpose_reduce32(v, lane);
                    acc2[cb] += warp_transpose_reduce32(w, lane);
                } else {
                    // rare: reduce each image's lanes separately and add them straight to global memory
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        const int nn = half == 0 ? n_first : n_last;
                        float a[32], b[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            a[i] = n == nn ? v[i] : 0.f;
                            b[i] = a[i] * a[i];
                        }
                        const float s1 = warp_transpose_reduce32(a, lane);
                        const float s2 = warp_transpose_reduce32(b, lane);
                        if (nn < p.nimg) {
                            double *dst = p.sums + ((long long)nn * COUT + c0 + lane) * 2;
                            atomicAdd(dst, (double)s1);
                            atomicAdd(dst + 1, (double)s2);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        }
        flush();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Layer 0 without a raw round trip: conv0 is 288 FMAs per pixel, cheaper to recompute than to store.
//   pass 1 (conv0_stats_kernel):  patterns -> per-(image, channel) sum / sum of squares
//   pass 2 (conv0_finish_kernel): patterns -> conv -> normalise -> LeakyReLU -> fp16 hi/lo padded planes [n,130,130,32]
// One thread = one pixel x 32 channels; one CTA = one image row pair.
// ---------------------------------------------------------------------------------------------
// Four horizontally adjacent pixels x 32 channels per thread (each weight float4 feeds 16 FMAs).
template <bool U8>
__device__ __forceinline__ void conv0_quad(const void *__restrict__ patterns, const float *ws, long long n, int y,
                                           int x0, float (&acc)[4][32]) {
    float in[3][6];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int yy = y + dy - 1;
#pragma unroll
        for (int dx = 0; dx < 6; ++dx) {
            const int xx = x0 + dx - 1;
            float v = 0.f;
            if (yy >= 0 && yy < 128 && xx >= 0 && xx < 128) {
                const long long off = (n * 128 + yy) * 128 + xx;
                if (U8) v = (float)((const uint8_t *)patterns)[off] / 255.0f;
                else v = ((const float *)patterns)[off];
            }
            in[dy][dx] = v;
        }
    }
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[px][c] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 w = *(const float4 *)(ws + (dy * 3 + dx) * 32 + c4 * 4);
#pragma unroll
                for (int px = 0; px < 4; ++px) {
                    const float a = in[dy][dx + px];
                    acc[px][c4 * 4 + 0] = fmaf(a, w.x, acc[px][c4 * 4 + 0]);
                    acc[px][c4 * 4 + 1] = fmaf(a, w.y, acc[px][c4 * 4 + 1]);
                    acc[px][c4 * 4 + 2] = fmaf(a, w.z, acc[px][c4 * 4 + 2]);
                    acc[px][c4 * 4 + 3] = fmaf(a, w.w, acc[px][c4 * 4 + 3]);
                }
            }
        }
}

// grid = (16 row groups, nimg); block = 256 threads = 8 rows x 32 pixel quads
template <bool U8>
__global__ void __launch_bounds__(256) conv0_stats_kernel(const void *__restrict__ patterns,
                                                          const float *__restrict__ w0, double *__restrict__ sums) {
    __shared__ float ws[9 * 32];
    __shared__ float red[8][32][2];
    for (int i = threadIdx.x; i < 9 * 32; i += 256) ws[i] = w0[i];
    __syncthreads();
    const long long n = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int y = blockIdx.x * 8 + warp, x0 = lane * 4;
    float acc[4][32];
    conv0_quad<U8>(patterns, ws, n, y, x0, acc);
    float s1[32], s2[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        s1[c] = (acc[0][c] + acc[1][c]) + (acc[2][c] + acc[3][c]);
        s2[c] = fmaf(acc[0][c], acc[0][c], acc[1][c] * acc[1][c]) + fmaf(acc[2][c], acc[2][c], acc[3][c] * acc[3][c]);
    }
    const float t1 = warp_transpose_reduce32(s1, lane);
    const float t2 = warp_transpose_reduce32(s2, lane);
    red[warp][lane][0] = t1;
    red[warp][lane][1] = t2;
    __syncthreads();
    if (threadIdx.x < 64) {
        const int c = threadIdx.x >> 1, which = threadIdx.x & 1;
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][c][which];
        atomicAdd(&sums[(n * 32 + c) * 2 + which], (double)t);
    }
}

template <bool U8>
__global__ void __launch_bounds__(256) conv0_finish_kernel(const void *__restrict__ patterns,
                                                           const float *__restrict__ w0,
                                                           const double *__restrict__ sums, __half *__restrict__ hi,
                                                           __half *__restrict__ lo) {
    __shared__ float ws[9 * 32];
    __shared__ float s_mean[32], s_rstd[32];
    const long long n = blockIdx.y;
    for (int i = threadIdx.x; i < 9 * 32; i += 256) ws[i] = w0[i];
    if (threadIdx.x < 32) {
        const double inv = 1.0 / 16384.0;
        const double mm = sums[(n * 32 + threadIdx.x) * 2] * inv;
        double var = sums[(n * 32 + threadIdx.x) * 2 + 1] * inv - mm * mm;
        if (var < 0.0) var = 0.0;
        s_mean[threadIdx.x] = (float)mm;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + 1e-5));
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int y = blockIdx.x * 8 + warp, x0 = lane * 4;
    float acc[4][32];
    conv0_quad<U8>(patterns, ws, n, y, x0, acc);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int px = 0; px < 4; ++px) {
        const long long off = ((n * 130 + (y + 1)) * 130 + (x0 + px + 1)) * 32;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
            __half h[8], l[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = c8 * 8 + j;
                float t = (acc[px][c] - s_mean[c]) * s_rstd[c];
                t = t >= 0.f ? t : t * 0.02f;
                h[j] = __float2half_rn(t);
                l[j] = __float2half_rn(t - __half2float(h[j]));
            }
            *(uint4 *)(hi + off + c8 * 8) = *(const uint4 *)h;
            *(uint4 *)(lo + off + c8 * 8) = *(const uint4 *)l;
        }
    }
    // zero borders of the padded [130,130] plane
    auto zero_px = [&](int yp, int xp) {
        const long long o2 = ((n * 130 + yp) * 130 + xp) * 32;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
            *(uint4 *)(hi + o2 + c8 * 8) = z;
            *(uint4 *)(lo + o2 + c8 * 8) = z;
        }
    };
    if (lane == 0) zero_px(y + 1, 0);
    if (lane == 31) zero_px(y + 1, 129);
    if (y == 0 || y == 127) {
        const int yb = y == 0 ? 0 : 129;
#pragma unroll
        for (int px = 0; px < 4; ++px) zero_px(yb, x0 + px + 1);
        if (lane == 0) zero_px(yb, 0);
        if (lane == 31) zero_px(yb, 129);
    }
}

// f32 NHWC (un-padded) -> zero-padded fp16 hi / lo planes (debug hook)
__global__ void split_pad_f32_kernel(const float *__restrict__ x, __half *__restrict__ hi, __half *__restrict__ lo,
                                     int H, int W, int CH, long long B) {
    const int Hp = H + 2, Wp = W + 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = B * Hp * Wp * CH;
    if (i >= total) return;
    const int c = (int)(i % CH);
    long long r = i / CH;
    const int xp = (int)(r % Wp);
    r /= Wp;
    const int yp = (int)(r % Hp);
    const long long n = r / Hp;
    float v = 0.f;
    if (yp >= 1 && yp <= H && xp >= 1 && xp <= W) v = x[((n * H + yp - 1) * W + xp - 1) * CH + c];
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
}

}  // namespace ebsd
