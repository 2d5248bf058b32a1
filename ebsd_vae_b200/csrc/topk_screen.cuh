// K2s: tensor-core SCREEN for the exact top-k.
//
// The exact search ranks rows by the canonical fp32 dot s = q.d (fma chain, oracle/topk_ref.c).  On CUDA cores that
// costs 16 FMAs per (query, row) pair and is bound by the fp32 pipe (topk_kernel, ~42 TFLOP/s).  Here the tensor
// cores compute an APPROXIMATE dot s~ for every pair, only pairs that can still matter survive, and the survivors are
// re-ranked with the canonical arithmetic -- the final lists are bit-identical to the exact kernel's.
//
//   operands   rows and queries as fp16 pairs [hi(16) | lo(16)] (64-byte K-major rows, SWIZZLE_64B): x = hi + lo up
//              to 2^-22 |x|;  s~ = q_hi.d_hi + q_hi.d_lo + q_lo.d_hi  = three tcgen05.mma (M = 128 queries,
//              N = 2 x 128 rows, K = 16 each) into fp32 TMEM accumulators.
//   error      |s~ - s| <= EPS = 1e-5 for unit vectors: split remainders (<= 4 * 2^-24), fp16 underflow of the lo parts
//              (<= 16 * 2^-25), 48 fp32 accumulations in the tensor core (<= 48 * 2^-23 even if it truncates) and the
//              rounding of the canonical chain itself (<= 16 * 2^-24); tests measure the actual maximum (~3e-7).
//   filter     a query keeps every row with s~ >= thr.  thr starts at tau0 - EPS, tau0 = the query's exact k-th best dot
//              over a prefix of the dictionary (found by the CUDA-core kernel first), and is raised whenever a
//              survivor buffer runs out of room: thr = (k-th largest s~ among rows already kept) - 2 EPS.
//              Those k rows have s >= kth - EPS, so the final k-th best exact dot S_k >= kth - EPS, and a dropped row
//              has s <= s~ + EPS < kth - EPS <= S_k: it cannot be in the top-k, ties included.
//   layout     TMEM lane = query, column = dictionary row.  A 256-row tile is computed as two 128-column halves with
//              their own accumulator and barriers; four accumulators = 512 TMEM columns = (tile parity, half).
//              16 epilogue warps = 4 lane quarters x 4 GROUPS, group g = (tile parity << 1) | half: a warp owns ONE
//              accumulator, visits every second tile and scans all 128 columns of its half there (4 tcgen05.ld of 32
//              columns; the barrier round trip, the address arithmetic and the loop overhead are paid once per 128
//              columns -- with 64 columns of every tile per warp they were half of the ~100 instructions per warp and
//              tile).  An epilogue thread owns one query (its threshold lives in a register): a 3-input max tree and
//              one compare per chunk; survivors go to a per-(work item, group, query) buffer of CAP entries in
//              global memory, compacted out of line when fewer than 32 entries are free (k-th largest -> new thr).
//              Tile parity is RELATIVE to the pass (t - tile_begin), so the re-rank knows which rows a group covered.
//   re-rank    topk_rerank_kernel: one warp per query gathers the surviving rows of all its buffers, recomputes the
//              canonical fp32 dot and inserts into the (dot desc, row asc) sorted list of the exact kernel.
#pragma once
#include <cuda_fp16.h>
#include <math.h>

#include "tcgen05.cuh"

namespace ebsd {

constexpr int kScrM = 128;           // queries per work item (TMEM lanes)
constexpr int kScrN = 256;           // dictionary rows per tile (TMEM columns of one accumulator)
constexpr int kScrRowB = 64;         // bytes per operand row: 16 fp16 hi | 16 fp16 lo
constexpr int kScrTileB = kScrN * kScrRowB;   // 16 KiB
constexpr int kScrQB = kScrM * kScrRowB;      // 8 KiB
constexpr int kScrStages = 6;
constexpr int kScrCap = 96;          // entries of a survivor buffer; compacted when fewer than 32 are free (CAP - 32 > EBSD_MAX_TOPK)
constexpr int kScrGroups = 4;        // survivor groups of a work item = accumulators: (tile parity << 1) | column half
constexpr int kScrGroupCols = kScrN / 2;   // columns a group scans in each tile of its parity
constexpr float kScrEps = 1e-5f;
#ifndef EBSD_SCREEN_KO
#define EBSD_SCREEN_KO 0   // compile-time role knock-outs, timing only (tools/roles_screen.sh): 1 no max tree, 2 no TMEM loads, 4 one MMA of three
#endif
constexpr int kScrThreads = 128 + 128 * kScrGroups;   // warp 0 TMA, 1 MMA, 2 TMEM alloc, 4.. epilogue (lane quarter x column group)
constexpr int kScrSmem = 1024 + kScrStages * kScrTileB + kScrQB + 256;

// Work distribution.  A pass covers the dictionary tiles [tile_begin, tile_end) for every query tile: U = n_qtiles * T
// (query tile, dictionary tile) units, query-tile major.  CTA c of n_ctas takes the contiguous span that starts at
// c * span_base + min(c, span_rem) with span_base = U / n_ctas, span_rem = U % n_ctas (host side) -- every CTA gets the
// same number of units to within one (with whole (query tile, dictionary split) items dealt round-robin, 316 items on
// 148 CTAs left 29 % of the kernel idle at 10 M x 10 k), and neither the screen nor the re-rank divides by run-time
// 64-bit values to find a span (the re-rank used ~9 such divisions per query: a quarter of its instructions).  A CTA
// walks its span as ITEMS = maximal pieces inside one query tile.  Along the chain of CTAs either the CTA or the query tile
// advances from one item to the next, so `cta + qt` numbers the items uniquely (< n_ctas + n_qtiles): that is the
// slot of the item's survivor buffers, which the re-rank finds again with the same arithmetic.
struct ScreenParams {
    long long Q, N;
    int k;
    int n_qtiles, n_ctas;
    long long span_base;   // units per CTA (floor)
    int span_rem;          // the first span_rem CTAs take one unit more
    long long tile_begin, tile_end;   // dictionary tiles (256 rows each) this pass covers
    const float *tau0;  // [Q][k]: exact top-k dots of a dictionary prefix (column k-1 seeds the threshold)
    float *cand_s;      // [items][groups][128][CAP] approximate dots
    int *cand_i;        // [items][groups][128][CAP] shard-local rows
    int *cand_n;        // [items][groups][128]      entries used
};

struct ScreenItem {
    int qt, id;
    long long tile0, tile1;   // absolute dictionary tiles
};
__host__ __device__ inline long long screen_span_begin(const ScreenParams &p, long long cta) {
    return cta * p.span_base + (cta < p.span_rem ? cta : (long long)p.span_rem);
}
// the CTA whose span holds unit u (one division by the 32-bit-sized span length)
__host__ __device__ inline long long screen_span_owner(const ScreenParams &p, long long u) {
    const long long big = (long long)p.span_rem * (p.span_base + 1);   // units held by the longer spans
    const long long num = u < big ? u : u - big, den = u < big ? p.span_base + 1 : (p.span_base > 0 ? p.span_base : 1);
    const long long quo = ((num | den) >> 32) == 0 ? (long long)((unsigned)num / (unsigned)den) : num / den;
    return u < big ? quo : p.span_rem + quo;
}
// the item that starts at unit u of CTA `cta`'s span [.., u_end); advances u past it
__device__ __forceinline__ ScreenItem screen_item_at(const ScreenParams &p, int cta, long long &u, long long u_end) {
    const long long T = p.tile_end - p.tile_begin;
    ScreenItem it;
    it.qt = (int)(u / T);
    const long long t0 = u - (long long)it.qt * T;
    long long t1 = t0 + (u_end - u);
    if (t1 > T) t1 = T;
    it.id = cta + it.qt;
    it.tile0 = p.tile_begin + t0;
    it.tile1 = p.tile_begin + t1;
    u += t1 - t0;
    return it;
}

// fp32 rows [n,16] -> fp16 pairs [n][hi(16) | lo(16)]
__global__ void split_rows_f16_kernel(const float *__restrict__ x, __half *__restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 16) return;
    const long long r = i >> 4;
    const int c = (int)(i & 15);
    const float v = x[i];
    const __half h = __float2half_rn(v);
    out[r * 32 + c] = h;
    out[r * 32 + 16 + c] = __float2half_rn(v - __half2float(h));
}

// k-th largest of vals[0..n) (n <= kScrCap, k <= n): repeated maximum below the previous one, counting duplicates.
__device__ __noinline__ float kth_largest(const float *vals, int n, int k) {
    float bound = INFINITY;
    int taken = 0;
    for (;;) {
        float best = -INFINITY;
        int cnt = 0;
        for (int i = 0; i < n; ++i) {
            const float v = vals[i];
            if (v < bound) {
                if (v > best) {
                    best = v;
                    cnt = 1;
                } else if (v == best) {
                    ++cnt;
                }
            }
        }
        taken += cnt;
        if (taken >= k || cnt == 0) return best;
        bound = best;
    }
}

// A survivor buffer is nearly full: raise the threshold to (k-th largest kept s~) - 2 EPS and drop what falls below it.
// If that does not make room for another 32-column chunk, more than CAP - 32 rows lie within 2 EPS of the k-th best
// (massive duplication): the screen cannot narrow this query down, cnt = -1 tells the re-rank to scan the range exactly.
// Returns (new threshold, new count as int bits): by value, so that the caller's cnt / thr stay in registers.
__device__ __noinline__ float2 screen_compact(float *cs, int *ci, int cnt, int k) {
    const float kth = kth_largest(cs, cnt, k);
    float thr = kth - 2.0f * kScrEps;
    int w = 0;
    for (int r = 0; r < cnt; ++r) {
        const float sv = cs[r];
        const int iv = ci[r];
        if (sv >= thr) {
            cs[w] = sv;
            ci[w] = iv;
            ++w;
        }
    }
    cnt = w;
    if (cnt > kScrCap - 32) {
        cnt = -1;
        thr = INFINITY;
    }
    return make_float2(thr, __int_as_float(cnt));
}

__global__ void __launch_bounds__(kScrThreads, 1)
topk_screen_kernel(const __grid_constant__ CUtensorMap map_d, const __grid_constant__ CUtensorMap map_q,
                   const ScreenParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *smem_d = smem;
    uint8_t *smem_q = smem + kScrStages * kScrTileB;
    uint64_t *d_full = (uint64_t *)(smem_q + kScrQB);   // [stages]
    uint64_t *d_empty = d_full + kScrStages;            // [stages]
    uint64_t *q_full = d_empty + kScrStages;            // [1]
    uint64_t *q_empty = q_full + 1;                     // [1]
    uint64_t *tfull = q_empty + 1;                      // [4]: accumulator (tile parity, column half) -- see the MMA loop
    uint64_t *tempty = tfull + 4;                       // [4]
    uint32_t *tmem_slot = (uint32_t *)(tempty + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kScrStages; ++s) {
            mbar_init(&d_full[s], 1);
            mbar_init(&d_empty[s], 1);
        }
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int b = 0; b < 4; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 4);   // the 4 warps (lane quarters) of the accumulator's group
        }
        mbar_fence_init();
        tma_prefetch_desc(&map_d);
        tma_prefetch_desc(&map_q);
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long u_begin = screen_span_begin(p, blockIdx.x), u_end = screen_span_begin(p, (long long)blockIdx.x + 1);

    if (warp == 0) {
        // ===================== TMA: the query tile of an item once, then its dictionary tiles
        if (elect_one_sync()) {
            unsigned it = 0, qit = 0;
            for (long long u = u_begin; u < u_end; ++qit) {
                const ScreenItem item = screen_item_at(p, blockIdx.x, u, u_end);
                mbar_wait_bounded(q_empty, (qit & 1u) ^ 1u);
                mbar_expect_tx(q_full, kScrQB);
                tma_load_2d(smem_q, &map_q, 0, item.qt * kScrM, q_full);
                for (long long t = item.tile0; t < item.tile1; ++t, ++it) {
                    const int s = it % kScrStages;
                    mbar_wait_bounded(&d_empty[s], ((it / kScrStages) & 1u) ^ 1u);
                    mbar_expect_tx(&d_full[s], kScrTileB);
                    tma_load_2d(smem_d + s * kScrTileB, &map_d, 0, (int)(t * kScrN), &d_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA: s~ = q_hi.d_hi + q_hi.d_lo + q_lo.d_hi, one accumulator per half tile
        if (elect_one_sync()) {
            // A 256-row dictionary tile is computed as two 128-column halves with their own TMEM accumulator and
            // barriers: the epilogue warps of half 0 and of half 1 form two independent double-buffered pipelines, so
            // the MMA -> drain -> release round trip of one half overlaps the other's (with one 256-column
            // accumulator per tile the hand-off was a latency chain: ~1400 cycles per tile against a 450-cycle drain).
            constexpr uint32_t idesc = umma_idesc_f16(kScrN / 2);
            unsigned it = 0, qit = 0, use0 = 0, use1 = 0;   // use*: tiles issued so far per relative tile parity
            for (long long u = u_begin; u < u_end; ++qit) {
                const ScreenItem item = screen_item_at(p, blockIdx.x, u, u_end);
                mbar_wait_bounded(q_full, qit & 1u);
                tc_fence_after();
                const uint32_t qa = smem_u32(smem_q);
                for (long long t = item.tile0; t < item.tile1; ++t, ++it) {
                    const int s = it % kScrStages;
                    mbar_wait_bounded(&d_full[s], (it / kScrStages) & 1u);
                    const int par = (int)((t - p.tile_begin) & 1);
                    const unsigned use = par ? use1 : use0;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int buf = par * 2 + half;
                        mbar_wait_bounded(&tempty[buf], (use & 1u) ^ 1u);
                        tc_fence_after();
                        const uint32_t db = smem_u32(smem_d + s * kScrTileB) + (uint32_t)(half * (kScrN / 2) * kScrRowB);
                        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * (kScrN / 2));
                        umma_f16(d_tmem, umma_smem_desc<kScrRowB>(qa), umma_smem_desc<kScrRowB>(db), idesc, 0u);            // hi.hi
#if !(EBSD_SCREEN_KO & 4)
                        umma_f16(d_tmem, umma_smem_desc<kScrRowB>(qa), umma_smem_desc<kScrRowB>(db + 32), idesc, 1u);       // hi.lo
                        umma_f16(d_tmem, umma_smem_desc<kScrRowB>(qa + 32), umma_smem_desc<kScrRowB>(db), idesc, 1u);       // lo.hi
#endif
                        if (half == 1) umma_commit(&d_empty[s]);
                        umma_commit(&tfull[buf]);
                    }
                    if (par) ++use1;
                    else ++use0;
                }
                umma_commit(q_empty);  // the query tile may be replaced once this item's MMAs have read it
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: threshold filter, one thread = one query x the 128 columns of one accumulator
        const int quarter = warp & 3, g = (warp - 4) >> 2;   // g = accumulator = (tile parity << 1) | column half
        const int par = g >> 1, half = g & 1;
        const int m = quarter * 32 + lane;
        // shared-window addresses of this warp's accumulator barriers, computed once
        // (the opaque mov keeps ptxas from rematerialising the address computation inside the loop)
        uint32_t tfull_u32, tempty_u32, t_row;
        asm volatile("mov.u32 %0, %1;" : "=r"(tfull_u32) : "r"(smem_u32(tfull) + (uint32_t)(g * 8)));
        asm volatile("mov.u32 %0, %1;" : "=r"(tempty_u32) : "r"(smem_u32(tempty) + (uint32_t)(g * 8)));
        asm volatile("mov.u32 %0, %1;" : "=r"(t_row) : "r"(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * kScrGroupCols)));
        const int nrows = p.N > 0x7fffffffll ? 0x7fffffff : (int)p.N;   // rows are 32-bit throughout (cand_i)
        unsigned use = 0;   // tiles of this parity drained so far
        for (long long u = u_begin; u < u_end;) {
            const ScreenItem item = screen_item_at(p, blockIdx.x, u, u_end);
            const int qt = item.qt;
            const int tile1 = (int)item.tile1;
            const int tile_first = (int)item.tile0 + ((((int)(item.tile0 - p.tile_begin) ^ par) & 1));
            const bool live = (long long)qt * kScrM + m < p.Q;
            const long long slot = ((long long)item.id * kScrGroups + g) * kScrM + m;
            float *cs = p.cand_s + slot * kScrCap;
            int *ci = p.cand_i + slot * kScrCap;
            int cnt = 0;
            // k prefix rows have exact dots >= tau0, so S_k >= tau0; a row with s~ < tau0 - EPS has s < tau0
            float thr = live ? p.tau0[((long long)qt * kScrM + m) * p.k + (p.k - 1)] - kScrEps : INFINITY;
            for (int t = tile_first; t < tile1; t += 2, ++use) {
                mbar_wait_bounded_u32(tfull_u32, use & 1u);
                tc_fence_after();
                const int row_base = t * kScrN + half * kScrGroupCols;
                // survivors of one 32-column chunk (rare).  The code is kept SMALL on purpose: the first version inlined
                // the buffer compaction into each of the 32 unrolled steps (17k SASS instructions, far beyond the
                // instruction cache) and every entry cost ~2000 cycles of instruction fetch.  Now: room for a whole
                // chunk is made up front (out of line), the steps are predicated appends.
                // The maximum tree keeps the maxima g[0..3] of the four 8-column quarters of the chunk, and the survivor
                // path only walks the quarters that hold one: on small shards with many queries most chunks have a survivor
                // in SOME lane (100 k rows x 80 k queries: 61 % of the warp-chunks, 2130 cycles per tile against 1000 at
                // 1.25 M rows), and walking all 32 columns for the one or two rows that pass was what they paid for.
                auto scan = [&](const float (&v)[32], const float (&g)[4], int c0) {
                    if (cnt > kScrCap - 32) {
                        const float2 r = screen_compact(cs, ci, cnt, p.k);
                        thr = r.x;
                        cnt = __float_as_int(r.y);
                    }
                    const int row0 = row_base + c0;
                    const int nvalid = nrows - row0;   // rows past the end are TMA zero fill
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (g[j] >= thr) {
#pragma unroll
                            for (int i = 8 * j; i < 8 * j + 8; ++i) {
                                if (v[i] >= thr && i < nvalid) {
                                    cs[cnt] = v[i];
                                    ci[cnt] = row0 + i;
                                    ++cnt;
                                }
                            }
                        }
                    }
                };
                auto quarter_max = [](const float (&v)[32], float (&g)[4]) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float a = fmaxf(fmaxf(v[8 * j], v[8 * j + 1]), v[8 * j + 2]);
                        const float b = fmaxf(fmaxf(v[8 * j + 3], v[8 * j + 4]), v[8 * j + 5]);
                        g[j] = fmaxf(fmaxf(a, b), fmaxf(v[8 * j + 6], v[8 * j + 7]));
                    }
                    return fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
                };
#pragma unroll 1
                for (int c0 = 0; c0 < kScrGroupCols; c0 += 32) {
                    float v[32];
#if EBSD_SCREEN_KO & 2
                    continue;
#endif
                    tmem_ld32(t_row + c0, v);
                    tmem_ld_wait();
#if EBSD_SCREEN_KO & 1
                    if (v[0] + v[31] == 12345.678f) cnt = -1;
                    continue;
#endif
                    float g[4];
                    if (quarter_max(v, g) >= thr) scan(v, g, c0);   // queries past Q have thr = +inf
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_u32) : "memory");
            }
            p.cand_n[slot] = cnt;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace ebsd
