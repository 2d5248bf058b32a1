// K2: exact cosine top-k over the 16-D latent dictionary (canonical fp32, see oracle/topk_ref.c).
//
// Replaces FaissLatentVectorDatabase.query_similar / _l2_normalize (latice/index/faiss_db.py:109-113,216-256)
// and ChromaLatentVectorDatabase.query_similar (latice/index/chroma_db.py:231-259).
//
// Shape of the kernel
//   * persistent CTAs (two per SM with TQ = 4) walk work items (query tile, dictionary split);
//   * one elected thread streams 128-row dictionary tiles (8 KiB) into a 4-stage shared-memory ring with TMA
//     (cp.async.bulk.tensor.2d, SWIZZLE_64B so that the consumers' LDS.128 are bank-conflict free); it runs
//     kStages-1 tiles ahead of the consumers and is itself lane 0 of warp 0;
//   * 8 warps each own 2*TQ queries; a thread holds a TQ x 8 register tile of dot products,
//     accumulated as fma chains in ascending j (bit-identical to the oracle);
//   * selection: per lane and query the best of its 8 rows is compared with the query's current k-th best (tau, in
//     registers); one vote skips the tile, one ballot per query slot finds the lanes with survivors, and only those
//     pay a ballot per row; survivors are inserted into a sorted per-query list (one entry per lane, ballot +
//     shuffle insertion).  No CTA-wide synchronisation in the steady state.
//   * ties: dot descending, then row index ascending -- per query the rows are visited in ascending order and tau
//     only lets strictly larger dots through once the list is full, which is exactly that rule.
//   * batched searches (N >= 65 536 rows, Q >= 2048) go through the tensor-core screen of topk_screen.cuh first and only
//     the survivors are re-ranked with this arithmetic (topk_rerank_kernel): same lists, bit for bit.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "topk_screen.cuh"

namespace ebsd {

constexpr int kD = 16;
constexpr int kTileRows = 128;
constexpr int kTileBytes = kTileRows * kD * 4;
constexpr int kStages = 4;
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kListLen = 32;                    // one entry per lane; k <= 32
constexpr int kIdxEmpty = 0x7fffffff;

struct __align__(8) Entry {
    float dot;
    int idx;
};

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ bool beats(float da, long long ia, float db, long long ib) {
    return (da > db) || (da == db && ia < ib);
}

// Insert candidate (cd, ci) into the warp-distributed sorted list (lane l holds entry l).
template <typename IdxT>
__device__ __forceinline__ void warp_insert(float &e_dot, IdxT &e_idx, float cd, IdxT ci, int lane) {
    const bool mine = beats(e_dot, (long long)e_idx, cd, (long long)ci);
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    const int pos = __popc(m);  // the list is sorted, so `mine` is true on a prefix of lanes
    const float up_dot = __shfl_up_sync(0xffffffffu, e_dot, 1);
    const IdxT up_idx = __shfl_up_sync(0xffffffffu, e_idx, 1);
    if (lane == pos) {
        e_dot = cd;
        e_idx = ci;
    } else if (lane > pos) {
        e_dot = up_dot;
        e_idx = up_idx;
    }
}

struct TopkParams {
    const float *queries;  // [Q,16] normalised
    long long Q;
    long long N;
    long long index_base;
    int k;
    int n_qtiles;
    int n_splits;
    int tiles_per_split;   // dictionary tiles per split (last split may be shorter)
    Entry *parts;          // [S,Q,k] when n_splits > 1
    float *out_dot;        // [Q,k] when n_splits == 1
    long long *out_idx;
    float *out_dist;
};

template <int TQ>
struct Smem {
    static constexpr int QT = kWarps * 2 * TQ;
    static constexpr int off_dtile = 0;
    static constexpr int off_qtile = off_dtile + kStages * kTileBytes;
    static constexpr int off_list = off_qtile + QT * kD * 4;
    static constexpr int off_bar = off_list + QT * kListLen * 8;
    static constexpr int bytes = off_bar + 2 * kStages * 8;
    static constexpr int alloc = bytes + 1024;  // slack for manual 1024-byte alignment
};

template <int TQ>
__global__ void __launch_bounds__(kThreads, TQ <= 4 ? 2 : 1)
topk_kernel(const __grid_constant__ CUtensorMap dict_map, const TopkParams p) {
    using S = Smem<TQ>;
    constexpr int QT = S::QT;
    constexpr int QPW = 2 * TQ;  // queries per warp

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float *qtile = (float *)(smem + S::off_qtile);
    Entry *lists = (Entry *)(smem + S::off_list);
    uint64_t *full_bar = (uint64_t *)(smem + S::off_bar);
    uint64_t *empty_bar = full_bar + kStages;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kWarps);
        }
        mbar_fence_init();
        tma_prefetch_desc(&dict_map);
    }
    __syncthreads();

    const long long total_tiles = (p.N + kTileRows - 1) / kTileRows;
    const int n_items = p.n_qtiles * p.n_splits;
    unsigned it = 0;  // running tile counter of this CTA: ring stage/phase bookkeeping (same on every warp)

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int split = item / p.n_qtiles;
        const int qt = item - split * p.n_qtiles;
        const long long tile0 = (long long)split * p.tiles_per_split;
        long long tile1 = tile0 + p.tiles_per_split;
        if (tile1 > total_tiles) tile1 = total_tiles;
        const int ntiles = (int)(tile1 > tile0 ? tile1 - tile0 : 0);
        const long long q0 = (long long)qt * QT;

        // ---- stage the query tile and reset the per-query state
        {
            // smem layout: query (w, qg, qi) lives at row (w*TQ + qi)*2 + qg so that the two lane halves of a
            // warp read different banks.
            for (int v = tid; v < QT * 4; v += kWarps * 32) {
                const int ql = v >> 2, c = v & 3;  // query within tile (w*QPW + qg*TQ + qi), 16-byte chunk
                const int w = ql / QPW, r = ql - w * QPW, qg = r / TQ, qi = r - qg * TQ;
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                if (q0 + ql < p.Q) val = *(const float4 *)(p.queries + (q0 + ql) * kD + c * 4);
                *(float4 *)(qtile + (((w * TQ + qi) * 2 + qg) * kD) + c * 4) = val;
            }
            for (int v = tid; v < QT * kListLen; v += kWarps * 32) {
                lists[v].dot = -INFINITY;
                lists[v].idx = kIdxEmpty;
            }
        }
        __syncthreads();

        // Producer duty (warp 0, lane 0): tile t of this item goes to ring slot (it + t) % kStages.
        auto issue_tile = [&](int t) {
            const unsigned pit = it + (unsigned)t;
            const int s = pit % kStages;
            const unsigned ph = (pit / kStages) & 1u;
            mbar_wait(&empty_bar[s], ph ^ 1u);
            mbar_expect_tx(&full_bar[s], kTileBytes);
            tma_load_2d(smem + S::off_dtile + s * kTileBytes, &dict_map, 0, (int)((tile0 + t) * kTileRows),
                        &full_bar[s]);
        };
        if (tid == 0) {
            for (int t = 0; t < kStages - 1 && t < ntiles; ++t) issue_tile(t);
        }
        {
            const int qg = lane >> 4;   // which half of the warp's queries
            const int dg = lane & 15;   // rows dg + 16*i of every tile
            const int wq0 = warp * QPW; // first query (tile-local) of this warp
            const int swz = (dg >> 1) & 3;
            float tau[TQ];
#pragma unroll
            for (int qi = 0; qi < TQ; ++qi) tau[qi] = -INFINITY;

            unsigned cit = it;
            for (int t = 0; t < ntiles; ++t, ++cit) {
                const int s = cit % kStages;
                const unsigned ph = (cit / kStages) & 1u;
                mbar_wait(&full_bar[s], ph);
                // explicit shared-window addresses: the manually aligned smem pointer has lost its address space and
                // would compile to generic 64-bit loads
                const uint32_t dt_u32 = smem_u32(smem + S::off_dtile + s * kTileBytes);
                const uint32_t qt_u32 = smem_u32(qtile);

                float acc[TQ][8];
#pragma unroll
                for (int qi = 0; qi < TQ; ++qi)
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[qi][i] = 0.f;

#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float4 dv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) dv[i] = lds_f4(dt_u32 + (uint32_t)(((dg + 16 * i) * 4 + (c ^ swz)) * 16));
#pragma unroll
                    for (int qi = 0; qi < TQ; ++qi) {
                        const float4 qv = lds_f4(qt_u32 + (uint32_t)(((((warp * TQ + qi) * 2 + qg) * kD) + c * 4) * 4));
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float a = acc[qi][i];
                            a = fmaf(qv.x, dv[i].x, a);
                            a = fmaf(qv.y, dv[i].y, a);
                            a = fmaf(qv.z, dv[i].z, a);
                            a = fmaf(qv.w, dv[i].w, a);
                            acc[qi][i] = a;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
                if (tid == 0 && t + kStages - 1 < ntiles) issue_tile(t + kStages - 1);

                // ---- selection: fast reject against tau; survivors are found with warp ballots and inserted one by
                // one, so the cost is proportional to the number of survivors (a few per query per tile once tau
                // has warmed up), not to the tile size.  For one query the visiting order is (i, lane) ascending =
                // row ascending, which is what the strict "dot > tau" rule needs for the lower-row-wins tie order.
                // Per lane and query: the best of its 8 rows.  One vote tells whether the tile can be skipped; if
                // not, one ballot per query slot finds the lanes to look at, and only those cost a ballot per row.
                float best[TQ];
                bool hit = false;
#pragma unroll
                for (int qi = 0; qi < TQ; ++qi) {
                    float b = acc[qi][0];
#pragma unroll
                    for (int i = 1; i < 8; ++i) b = fmaxf(b, acc[qi][i]);
                    best[qi] = b;
                    hit |= b > tau[qi];
                }
                if (__any_sync(0xffffffffu, hit)) {
                    const long long row_base = (tile0 + t) * kTileRows;  // shard-local row of tile row 0
                    long long valid_ll = p.N - row_base;
                    const int valid = valid_ll > kTileRows ? kTileRows : (int)valid_ll;
#pragma unroll
                    for (int qi = 0; qi < TQ; ++qi) {
                        if (__ballot_sync(0xffffffffu, best[qi] > tau[qi]) == 0u) continue;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int row = dg + 16 * i;
                            unsigned m = __ballot_sync(0xffffffffu, row < valid && acc[qi][i] > tau[qi]);
                            while (m) {  // warp-uniform
                                const int src = __ffs(m) - 1;
                                m &= m - 1;
                                const float cd = __shfl_sync(0xffffffffu, acc[qi][i], src);
                                const int sq = src >> 4;  // which half of the warp's queries the source lane serves
                                const int ql = wq0 + sq * TQ + qi;
                                // tau of that query may have risen since the ballot (earlier survivor of this tile)
                                const float cur_tau = __shfl_sync(0xffffffffu, tau[qi], sq << 4);
                                if (!(cd > cur_tau)) continue;
                                Entry e = lists[ql * kListLen + lane];
                                warp_insert<int>(e.dot, e.idx, cd, (int)(row_base + (src & 15) + 16 * i), lane);
                                lists[ql * kListLen + lane] = e;
                                const float kth = __shfl_sync(0xffffffffu, e.dot, p.k - 1);  // -inf until k entries
                                if (qg == sq) tau[qi] = kth;
                            }
                        }
                    }
                }
            }

            // ---- write this warp's lists
            __syncwarp();
            for (int ql = wq0; ql < wq0 + QPW; ++ql) {
                const long long qglob = q0 + ql;
                if (qglob >= p.Q || lane >= p.k) continue;
                const Entry e = lists[ql * kListLen + lane];
                if (p.n_splits > 1) {
                    p.parts[((long long)split * p.Q + qglob) * p.k + lane] = e;
                } else {
                    const bool empty = e.idx == kIdxEmpty;
                    p.out_dot[qglob * p.k + lane] = e.dot;
                    p.out_idx[qglob * p.k + lane] = empty ? -1ll : p.index_base + e.idx;
                    if (p.out_dist) p.out_dist[qglob * p.k + lane] = 1.0f - e.dot;
                }
            }
        }
        it += (unsigned)ntiles;
        __syncthreads();
    }
}

// Merge S partial lists per query (one warp per query). Partials hold shard-local int rows.
// The S*k entries of a query are read 32 at a time (coalesced, independent loads) and only those that beat the running
// k-th best are inserted: the first version walked them one by one with dependent loads and took 440 us for the
// 296 partial lists of a 10 M-row single-query search -- more than the search itself.
__global__ void __launch_bounds__(kThreads) topk_merge_parts_kernel(const Entry *parts, int S, long long Q, int k,
                                                                    long long index_base, float *out_dot,
                                                                    long long *out_idx, float *out_dist) {
    // one CTA per query: each warp merges a slice of the S*k entries, warp 0 merges the eight lists
    __shared__ Entry lists_s[kWarps][kListLen];
    const long long q = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float e_dot = -INFINITY;
    int e_idx = kIdxEmpty;
    const int total = S * k;
    const int per_warp = ((total + kWarps - 1) / kWarps + 63) / 64 * 64;
    const int t_end = (warp + 1) * per_warp < total ? (warp + 1) * per_warp : total;
    auto fetch = [&](int t) {
        Entry e;
        e.dot = -INFINITY;
        e.idx = kIdxEmpty;
        if (t < t_end) {
            const int sp = t / k, j = t - sp * k;
            e = parts[((long long)sp * Q + q) * k + j];
        }
        return e;
    };
    auto offer = [&](const Entry &e) {
        const float kd = __shfl_sync(0xffffffffu, e_dot, k - 1);
        const int ki = __shfl_sync(0xffffffffu, e_idx, k - 1);
        unsigned m = __ballot_sync(0xffffffffu, e.idx != kIdxEmpty && beats(e.dot, e.idx, kd, ki));
        while (m) {  // warp-uniform
            const int src = __ffs(m) - 1;
            m &= m - 1;
            warp_insert<int>(e_dot, e_idx, __shfl_sync(0xffffffffu, e.dot, src), __shfl_sync(0xffffffffu, e.idx, src), lane);
        }
    };
    // eight coalesced 32-entry loads in flight per round trip (per_warp is a multiple of 64; rounds of 256 entries)
    for (int base = warp * per_warp; base < t_end; base += 256) {
        Entry e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = fetch(base + 32 * j + lane);
#pragma unroll
        for (int j = 0; j < 8; ++j) offer(e[j]);
    }
    lists_s[warp][lane].dot = e_dot;
    lists_s[warp][lane].idx = e_idx;
    __syncthreads();
    if (warp != 0) return;
    for (int w = 1; w < kWarps; ++w) {
        Entry e = lists_s[w][lane];
        if (lane >= k) e.idx = kIdxEmpty;
        offer(e);
    }
    if (lane < k) {
        out_dot[q * k + lane] = e_dot;
        out_idx[q * k + lane] = e_idx == kIdxEmpty ? -1ll : index_base + e_idx;
        if (out_dist) out_dist[q * k + lane] = 1.0f - e_dot;
    }
}

// K2q: the interactive case (index_pattern: one query, or a handful) over a large dictionary is a pure HBM stream:
// 64 B per row against 32 FLOP per (query, row).  No shared-memory staging: every lane owns a row, reads its 64 bytes
// with four 16-byte loads that bypass L1, and the loads of the next batch are in flight while the current one is
// reduced (register double buffer), so each SM keeps >= 64 KiB outstanding.  Queries are broadcast from shared
// memory; the dot is the canonical fma chain; a warp walks a contiguous ascending row range and keeps TQ lists in
// registers (lane l = entry l), so the strict `dot > tau` filter gives the (dot desc, row asc) order; the 8 warps
// of a CTA merge through shared memory and each CTA writes one partial list per query for topk_merge_parts_kernel.
constexpr int kStreamU = 2;   // rows per lane and batch

template <int TQ>
__global__ void __launch_bounds__(kThreads, TQ <= 2 ? 3 : 2)
topk_stream_kernel(const float *__restrict__ dict, const TopkParams p, long long rows_per_warp) {
    __shared__ float4 qs[TQ][4];
    __shared__ Entry lists_s[kWarps][TQ][kListLen];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < TQ * 4) {
        const int qi = tid >> 2, c = tid & 3;
        qs[qi][c] = qi < p.Q ? *(const float4 *)(p.queries + (long long)qi * kD + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const long long gw = (long long)blockIdx.x * kWarps + warp;
    const long long r_begin = gw * rows_per_warp;
    long long r_end = r_begin + rows_per_warp;
    if (r_end > p.N) r_end = p.N;

    float e_dot[TQ], tau[TQ];
    int e_idx[TQ];
#pragma unroll
    for (int qi = 0; qi < TQ; ++qi) {
        e_dot[qi] = -INFINITY;
        e_idx[qi] = kIdxEmpty;
        tau[qi] = -INFINITY;
    }
    auto load = [&](long long r0, float4 (&d)[kStreamU][4]) {
#pragma unroll
        for (int u = 0; u < kStreamU; ++u) {
            const long long row = r0 + u * 32 + lane;
            if (row < r_end) {
                const float4 *src = (const float4 *)(dict + row * kD);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(d[u][c].x), "=f"(d[u][c].y), "=f"(d[u][c].z), "=f"(d[u][c].w)
                                 : "l"(src + c));
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) d[u][c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    float4 cur[kStreamU][4], nxt[kStreamU][4];
    if (r_begin < r_end) load(r_begin, cur);
    for (long long r0 = r_begin; r0 < r_end; r0 += kStreamU * 32) {
        if (r0 + kStreamU * 32 < r_end) load(r0 + kStreamU * 32, nxt);
        float acc[TQ][kStreamU];
        bool hit = false;
#pragma unroll
        for (int qi = 0; qi < TQ; ++qi) {
            const float4 q0 = qs[qi][0], q1 = qs[qi][1], q2 = qs[qi][2], q3 = qs[qi][3];
#pragma unroll
            for (int u = 0; u < kStreamU; ++u) {
                float a = 0.f;
                a = fmaf(q0.x, cur[u][0].x, a); a = fmaf(q0.y, cur[u][0].y, a); a = fmaf(q0.z, cur[u][0].z, a); a = fmaf(q0.w, cur[u][0].w, a);
                a = fmaf(q1.x, cur[u][1].x, a); a = fmaf(q1.y, cur[u][1].y, a); a = fmaf(q1.z, cur[u][1].z, a); a = fmaf(q1.w, cur[u][1].w, a);
                a = fmaf(q2.x, cur[u][2].x, a); a = fmaf(q2.y, cur[u][2].y, a); a = fmaf(q2.z, cur[u][2].z, a); a = fmaf(q2.w, cur[u][2].w, a);
                a = fmaf(q3.x, cur[u][3].x, a); a = fmaf(q3.y, cur[u][3].y, a); a = fmaf(q3.z, cur[u][3].z, a); a = fmaf(q3.w, cur[u][3].w, a);
                if (r0 + u * 32 + lane >= r_end) a = -INFINITY;
                acc[qi][u] = a;
                hit |= a > tau[qi];
            }
        }
        if (__any_sync(0xffffffffu, hit)) {
#pragma unroll
            for (int qi = 0; qi < TQ; ++qi)
#pragma unroll
                for (int u = 0; u < kStreamU; ++u) {
                    unsigned m = __ballot_sync(0xffffffffu, acc[qi][u] > tau[qi]);
                    while (m) {  // warp-uniform; rows in ascending order
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const float cd = __shfl_sync(0xffffffffu, acc[qi][u], src);
                        if (!(cd > tau[qi])) continue;   // tau may have risen since the ballot
                        warp_insert<int>(e_dot[qi], e_idx[qi], cd, (int)(r0 + u * 32 + src), lane);
                        tau[qi] = __shfl_sync(0xffffffffu, e_dot[qi], p.k - 1);  // -inf until k entries
                    }
                }
        }
#pragma unroll
        for (int u = 0; u < kStreamU; ++u)
#pragma unroll
            for (int c = 0; c < 4; ++c) cur[u][c] = nxt[u][c];
    }
    // ---- CTA merge: warp qi collects query qi's eight lists
#pragma unroll
    for (int qi = 0; qi < TQ; ++qi) {
        lists_s[warp][qi][lane].dot = e_dot[qi];
        lists_s[warp][qi][lane].idx = e_idx[qi];
    }
    __syncthreads();
    if (warp < TQ && warp < p.Q) {
        float m_dot = -INFINITY;
        int m_idx = kIdxEmpty;
        for (int w = 0; w < kWarps; ++w) {
            const Entry e = lists_s[w][warp][lane];
            for (int j = 0; j < p.k; ++j) {
                const int ci = __shfl_sync(0xffffffffu, e.idx, j);
                if (ci == kIdxEmpty) break;  // sorted: empties at the tail
                warp_insert<int>(m_dot, m_idx, __shfl_sync(0xffffffffu, e.dot, j), ci, lane);
            }
        }
        if (lane < p.k) {
            Entry o;
            o.dot = m_dot;
            o.idx = m_idx;
            p.parts[((long long)blockIdx.x * p.Q + warp) * p.k + lane] = o;
        }
    }
}

// Merge R per-shard lists with global int64 indices (K2m, runs after the all-gather).
__global__ void topk_merge_global_kernel(const float *dots, const long long *idx, int R, long long Q, int k,
                                         float *out_dot, long long *out_idx, float *out_dist) {
    const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= Q) return;
    float e_dot = -INFINITY;
    long long e_idx = 0x7fffffffffffffffll;
    for (int r = 0; r < R; ++r) {
        const long long base = ((long long)r * Q + q) * k;
        for (int j = 0; j < k; ++j) {
            const long long ci = idx[base + j];
            if (ci < 0) break;
            warp_insert<long long>(e_dot, e_idx, dots[base + j], ci, lane);
        }
    }
    if (lane < k) {
        const bool empty = e_idx == 0x7fffffffffffffffll;
        out_dot[q * k + lane] = e_dot;
        out_idx[q * k + lane] = empty ? -1ll : e_idx;
        if (out_dist) out_dist[q * k + lane] = 1.0f - e_dot;
    }
}

// x[i,:] /= ||x[i,:]||, bit for bit what FaissLatentVectorDatabase._l2_normalize (latice/index/faiss_db.py:109-113)
// computes with numpy on float32 rows: rounded squares s_j, numpy's pairwise sum r_j = s_j + s_{j+8},
// ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)), IEEE sqrt and division, zero norm -> 1 (tests/golden/l2_normalize.npz).
__global__ void normalize_rows_kernel(float *x, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 *row = (float4 *)(x + i * kD);
    float4 v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = row[c];
    const float *e = (const float *)v;
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(__fmul_rn(e[j], e[j]), __fmul_rn(e[j + 8], e[j + 8]));
    const float n2 = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                               __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    float norm = __fsqrt_rn(n2);
    if (norm == 0.f) norm = 1.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        v[c].x = __fdiv_rn(v[c].x, norm);
        v[c].y = __fdiv_rn(v[c].y, norm);
        v[c].z = __fdiv_rn(v[c].z, norm);
        v[c].w = __fdiv_rn(v[c].w, norm);
        row[c] = v[c];
    }
}

// Re-rank of the screen's survivors (topk_screen.cuh): one warp per query, canonical fp32 dots, the exact kernel's
// (dot desc, row asc) list.  A buffer flagged -1 (more than CAP rows within 2 EPS of the k-th best) is replaced by a
// brute-force scan of the rows that buffer covers.
__device__ __forceinline__ float canonical_dot(const float4 (&q)[4], const float *__restrict__ row) {
    float a = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 d = __ldg((const float4 *)row + c);
        a = fmaf(q[c].x, d.x, a);
        a = fmaf(q[c].y, d.y, a);
        a = fmaf(q[c].z, d.z, a);
        a = fmaf(q[c].w, d.w, a);
    }
    return a;
}

// Threshold seed of the screen (run_screen): the k-th largest canonical dot of every query over the first n0 <= 512
// rows, written to out_dot[q][k - 1] -- the only entry the first screen pass reads (its re-rank rebuilds the lists from
// the survivors).  One warp per query, 8 or 16 rows per lane, k rounds of (lane maximum, warp maximum, the lowest owning lane
// drops ONE occurrence): no sorted lists, no insertions.  The tiled CUDA-core kernel spends its time there on such a
// short prefix (every query inserts k (1 + ln(n0 / k)) ~ 50 rows before its threshold bites): 376 us for 512 rows x
// 80 k queries, 0.83 ms for 4096 rows.
constexpr int kSeedRowsMax = 512;
template <int R>   // rows per lane: n0 <= 32 R
__global__ void __launch_bounds__(256) topk_seed_kth_kernel(const float *__restrict__ dict, int n0,
                                                            const float *__restrict__ queries, long long Q, int k,
                                                            float *__restrict__ out_dot) {
    const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= Q) return;
    float4 qv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) qv[c] = __ldg((const float4 *)(queries + q * kD) + c);
    float d[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const int row = j * 32 + lane;
        d[j] = row < n0 ? canonical_dot(qv, dict + (long long)row * kD) : -INFINITY;
    }
    float kth = -INFINITY;
    for (int t = 0; t < k; ++t) {
        float m = d[0];
#pragma unroll
        for (int j = 1; j < R; ++j) m = fmaxf(m, d[j]);
        float w = m;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, o));
        kth = w;
        const unsigned owners = __ballot_sync(0xffffffffu, m == w);
        if (lane == __ffs(owners) - 1) {   // drop one occurrence (duplicates count separately)
            bool done = false;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                if (!done && d[j] == w) {
                    d[j] = -INFINITY;
                    done = true;
                }
            }
        }
    }
    if (lane == 0) out_dot[q * k + (k - 1)] = kth;   // -inf when n0 < k: every row passes the first screen pass
}

// init_from_out: the output arrays already hold the exact top-k of the rows below the pass (global indices); they
// seed the list, so the result is the top-k over both ranges.
__global__ void __launch_bounds__(256) topk_rerank_kernel(const float *__restrict__ dict, const float *__restrict__ queries,
                                                          const ScreenParams p, long long index_base, int init_from_out,
                                                          float *out_dot, long long *out_idx, float *out_dist) {
    const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= p.Q) return;
    float4 qv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) qv[c] = __ldg((const float4 *)(queries + q * kD) + c);
    const int qt = (int)(q / kScrM), m = (int)(q % kScrM);
    float e_dot = -INFINITY;
    int e_idx = kIdxEmpty;
    if (init_from_out && lane < p.k) {   // a sorted list: lane l takes entry l
        const long long gi = out_idx[q * p.k + lane];
        if (gi >= 0) {
            e_dot = out_dot[q * p.k + lane];
            e_idx = (int)(gi - index_base);
        }
    }
    __syncwarp();
    // the items of this query tile: the CTAs whose unit spans intersect [qt T, (qt + 1) T) (ScreenParams)
    const long long T = p.tile_end - p.tile_begin, u_lo = (long long)qt * T, u_hi = u_lo + T;
    long long c = screen_span_owner(p, u_lo);
    long long s1 = screen_span_begin(p, c);
    for (; c < p.n_ctas; ++c) {
        const long long s0 = s1;
        s1 = s0 + p.span_base + (c < p.span_rem ? 1 : 0);
        if (s0 >= u_hi) break;
        const long long lo = s0 > u_lo ? s0 : u_lo, hi = s1 < u_hi ? s1 : u_hi;
        if (lo >= hi) continue;   // a CTA without units (fewer units than CTAs)
        const int item = (int)c + qt;
        for (int half = 0; half < kScrGroups; ++half) {
            const long long slot = ((long long)item * kScrGroups + half) * kScrM + m;
            const int n = p.cand_n[slot];
            if (n >= 0) {
                for (int base = 0; base < n; base += 32) {
                    const int j = base + lane;
                    int row = kIdxEmpty;
                    float d = -INFINITY;
                    if (j < n) {
                        row = p.cand_i[slot * kScrCap + j];
                        d = canonical_dot(qv, dict + (long long)row * kD);
                    }
                    // only candidates that beat the current k-th entry can enter the list: with the list of the previous
                    // passes as a start that is ~k ln 8 of the ~8 k survivors of a pass (ncu: the insertions of ALL
                    // survivors were 35 % of this kernel's instructions)
                    const float kd = __shfl_sync(0xffffffffu, e_dot, p.k - 1);
                    const int ki = __shfl_sync(0xffffffffu, e_idx, p.k - 1);
                    unsigned mask = __ballot_sync(0xffffffffu, beats(d, (long long)row, kd, (long long)ki));
                    while (mask) {
                        const int j2 = __ffs(mask) - 1;
                        mask &= mask - 1;
                        warp_insert<int>(e_dot, e_idx, __shfl_sync(0xffffffffu, d, j2), __shfl_sync(0xffffffffu, row, j2), lane);
                    }
                }
            } else {
                // the group's rows: tiles of relative parity half >> 1 in the item's range, column half (half & 1)
                const long long tile0 = p.tile_begin + (lo - u_lo), tile1 = p.tile_begin + (hi - u_lo);
                const long long tile_first = tile0 + (((tile0 - p.tile_begin) ^ (half >> 1)) & 1);
                for (long long t = tile_first; t < tile1; t += 2)
                    for (int c0 = 0; c0 < kScrGroupCols; c0 += 32) {
                        const long long row = t * kScrN + (half & 1) * kScrGroupCols + c0 + lane;
                        const float d = row < p.N ? canonical_dot(qv, dict + row * kD) : -INFINITY;
                        const float kth = __shfl_sync(0xffffffffu, e_dot, p.k - 1);
                        unsigned mask = __ballot_sync(0xffffffffu, row < p.N && d >= kth);
                        while (mask) {
                            const int src = __ffs(mask) - 1;
                            mask &= mask - 1;
                            warp_insert<int>(e_dot, e_idx, __shfl_sync(0xffffffffu, d, src), (int)(t * kScrN + (half & 1) * kScrGroupCols + c0 + src), lane);
                        }
                    }
            }
        }
    }
    if (lane < p.k) {
        out_dot[q * p.k + lane] = e_dot;
        out_idx[q * p.k + lane] = e_idx == kIdxEmpty ? -1ll : index_base + e_idx;
        if (out_dist) out_dist[q * p.k + lane] = 1.0f - e_dot;
    }
}

// ------------------------------------------------------------------------------------------ host side
struct TopkPlan {
    int tq;               // queries per thread (8 or 1)
    int qt;               // queries per CTA
    int n_qtiles;
    int n_splits;
    int tiles_per_split;
};

// queries per thread of the batched kernel: 4 (two CTAs of 8 warps per SM; default -- 16 resident warps hide the
// shared-memory and FMA latencies, measured 1.2-2.3x faster than one CTA of TQ = 8, profiles/README.md) or 8 (one CTA
// per SM, 64 accumulators per thread).  EBSD_TOPK_TQ=8 selects the latter for A/B timing.
static int topk_tq() {
    static int tq = 0;
    if (tq == 0) {
        const char *e = getenv("EBSD_TOPK_TQ");
        tq = (e && atoi(e) == 8) ? 8 : 4;
    }
    return tq;
}

static TopkPlan make_plan(long long N, long long Q, int sms) {
    TopkPlan pl;
    pl.tq = Q <= 16 ? 1 : topk_tq();
    sms *= pl.tq <= 4 ? 2 : 1;  // CTA slots: the TQ <= 4 kernels run two CTAs per SM
    pl.qt = kWarps * 2 * pl.tq;
    pl.n_qtiles = (int)((Q + pl.qt - 1) / pl.qt);
    const long long total_tiles = (N + kTileRows - 1) / kTileRows;
    // Choose the number of dictionary splits: fill whole waves of `sms` persistent CTAs while keeping each
    // split long enough to amortise the warm-up of the selection (tau starts at -inf in every split).
    const double overhead_tiles = 24.0;
    int best_s = 1;
    double best_eff = -1.0;
    const long long max_s = total_tiles < 1 ? 1 : (total_tiles < 4096 ? total_tiles : 4096);
    for (long long s = 1; s <= max_s && s <= 4096; ++s) {
        const long long tps = (total_tiles + s - 1) / s;
        if (tps < 4 && s > 1) break;
        const long long s_eff = (total_tiles + tps - 1) / tps;  // splits actually used
        const long long items = (long long)pl.n_qtiles * s_eff;
        const long long waves = (items + sms - 1) / sms;
        const double eff = ((double)items / (double)(waves * sms)) * ((double)tps / ((double)tps + overhead_tiles));
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best_s = (int)s_eff;
        }
        if (items >= 16ll * sms) break;
    }
    pl.n_splits = best_s;
    pl.tiles_per_split = (int)((total_tiles + best_s - 1) / best_s);
    if (pl.tiles_per_split < 1) pl.tiles_per_split = 1;
    pl.n_splits = (int)((total_tiles + pl.tiles_per_split - 1) / pl.tiles_per_split);
    if (pl.n_splits < 1) pl.n_splits = 1;
    return pl;
}

// ---- tensor-core screen (topk_screen.cuh): plan, workspace carving, launch
struct ScreenPlan {
    int n_qtiles, items;
    size_t off_dpairs, off_qpairs, off_cs, off_ci, off_cn, off_parts, bytes;
};

// Fewest rows of the CUDA-core seeding search, and the smallest prefix a further screen level is added for.  The seed
// costs ~0.2 ms per 1024 rows at 80 k queries (0.83 ms of a 2.7 ms search of a 100 k-row shard with 4096 rows), an
// extra level only a short screen pass plus one re-rank launch (~0.2 ms at 80 k queries, 0.03 ms at 10 k).
constexpr long long kScreenSeedMinRows = 256;
constexpr long long kScreenLevelMinRows = 1024;

// Rows of the seeding search.  With tau0 the k-th best of N/8 rows a query keeps about 8k survivors over the whole
// dictionary (k / n0 per row), spread over its (splits x column groups) buffers: the CAP-entry buffers practically
// never fill, so the costly in-place compaction stays an exception; the seeding search itself costs 1/8 of a
// CUDA-core search.  (N/16 measured slower: 41.7 vs 34.6 ms at 1M x 65536 -- twice the survivors, and every survivor
// makes its whole warp walk the 32-column chunk.)
static long long screen_prefix_rows(long long N) {
    long long n0 = N / 8;   // called with N = the rows of the first screen pass
    if (n0 < kScreenSeedMinRows) n0 = kScreenSeedMinRows;
    return n0 < N ? n0 : N;
}

// Stage boundaries of the screen in 256-row tiles, ascending: [0] = rows of the exact CUDA-core seeding search (not
// tile aligned), then the END tile of every screen pass; each pass covers 8x the rows of the one before
// (..., N/512, N/64, N/8, N -- as many levels as leave a prefix of >= 1024 rows; the seed is >= 256 rows).  The CUDA-core search costs ~12 ns per
// row and 10k queries against ~1 ns for the screen, so it should only ever see a few thousand rows: at 1.25 M x 80 k
// (one shard of the 10 M-row dictionary on 8 GPUs) the N/64 seed alone was 2 of 16.8 ms.
// EBSD_TOPK_SCREEN_LEVELS=n caps the number of screen passes (A/B timing; 2 = the round-1 staging).
struct ScreenStages {
    long long seed_rows;
    int n_pass;
    long long pass_end[8];   // tiles
};
static ScreenStages screen_stages(long long N) {
    static int max_levels = -1;
    if (max_levels < 0) {
        const char *e = getenv("EBSD_TOPK_SCREEN_LEVELS");
        max_levels = (e && atoi(e) >= 1 && atoi(e) <= 8) ? atoi(e) : 8;
    }
    const long long total_tiles = (N + kScrN - 1) / kScrN;
    long long ends[8];
    int n = 0;
    ends[n++] = total_tiles;
    long long rows = N;
    while (n < max_levels && rows / 8 >= kScreenLevelMinRows) {
        rows /= 8;
        ends[n++] = (rows + kScrN - 1) / kScrN;
    }
    ScreenStages st;
    st.seed_rows = screen_prefix_rows(ends[n - 1] * kScrN);
    if (st.seed_rows > N) st.seed_rows = N;
    st.n_pass = n;
    for (int i = 0; i < n; ++i) st.pass_end[i] = ends[n - 1 - i];
    return st;
}

constexpr long long kScreenQueryChunk = 131072;   // queries per screen pass (bounds the survivor buffers at ~256 MiB)

static bool screen_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("EBSD_TOPK_SCREEN");
        on = (e && atoi(e) == 0) ? 0 : 1;
    }
    return on == 1;
}

// The screen pays off once there are enough (query, row) pairs to amortise the operand conversion, the seeding search
// and the re-rank.  Measured switch-over (gpurun_out/sweep_screen.log, profiles/README.md): at 65 536 rows the screen
// wins from ~2k queries on (0.30 vs 0.41 ms at Q = 4096, 0.68 vs 1.12 ms at Q = 16 384); at 30 000 rows or 1024
// queries the CUDA-core kernel is as fast or faster.  EBSD_TOPK_SCREEN_MIN_ROWS moves the row threshold for A/B timing.
static bool screen_applies(long long N, long long Q, int k) {
    static long long min_rows = -1;
    if (min_rows < 0) {
        const char *e = getenv("EBSD_TOPK_SCREEN_MIN_ROWS");
        min_rows = (e && atoll(e) > 0) ? atoll(e) : 65536;
        if (min_rows < 8192) min_rows = 8192;   // stage A searches >= 4096 rows and stage B whole 256-row tiles
    }
    return screen_enabled() && k < kScrCap / 2 && N >= min_rows && Q >= 2048 && N < 0x7fffff00ll;
}

static ScreenPlan make_screen_plan(long long N, long long Q, int sms) {
    ScreenPlan pl;
    pl.n_qtiles = (int)((Q + kScrM - 1) / kScrM);
    const long long total_tiles = (N + kScrN - 1) / kScrN;
    (void)total_tiles;
    pl.items = sms + pl.n_qtiles;   // balanced unit spans: item ids are cta + query tile (ScreenParams)
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t r = off;
        off += (bytes + 1023) & ~(size_t)1023;
        return r;
    };
    const size_t slots = (size_t)pl.items * kScrGroups * kScrM;
    pl.off_dpairs = take((size_t)total_tiles * kScrN * kScrRowB);      // padded to whole tiles
    pl.off_qpairs = take((size_t)pl.n_qtiles * kScrM * kScrRowB);
    pl.off_cs = take(slots * kScrCap * sizeof(float));
    pl.off_ci = take(slots * kScrCap * sizeof(int));
    pl.off_cn = take(slots * sizeof(int));
    // partial lists of the seeding search (the CUDA-core kernel over the first few thousand rows, see run_screen)
    const TopkPlan ex = make_plan(screen_stages(N).seed_rows, Q, sms);
    pl.off_parts = take(ex.n_splits > 1 ? (size_t)ex.n_splits * (size_t)Q * 32 * sizeof(Entry) : 1024);
    pl.bytes = off + 1024;
    return pl;
}

static int make_pairs_map(CUtensorMap *map, const void *base, long long rows, int box_rows) {
    tensormap_encode_fn encode = get_tensormap_encode();
    if (!encode) {
        set_error("ebsd_topk: cuTensorMapEncodeTiled entry point not available");
        return EBSD_ERR_CUDA;
    }
    const cuuint64_t gdim[2] = {32, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)kScrRowB};
    const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)base, gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("ebsd_topk: cuTensorMapEncodeTiled(fp16 pairs) failed with %d", (int)cr);
        return EBSD_ERR_CUDA;
    }
    return EBSD_OK;
}

constexpr int kMaxDevices = 64;
static int current_device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

template <int TQ>
static int launch_topk(const CUtensorMap &map, const TopkParams &p, int sms, cudaStream_t st) {
    using S = Smem<TQ>;
    static bool configured[kMaxDevices] = {};   // function attributes are per device
    const int dev = current_device_slot();
    if (!configured[dev]) {
        EBSD_CUDA_TRY(cudaFuncSetAttribute(topk_kernel<TQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::alloc));
        configured[dev] = true;
    }
    const int n_items = p.n_qtiles * p.n_splits;
    const int slots = sms * (TQ <= 4 ? 2 : 1);  // resident CTAs
    const int grid = n_items < slots ? n_items : slots;
    topk_kernel<TQ><<<grid, kThreads, S::alloc, st>>>(map, p);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

// K2q applies to a handful of queries over a dictionary long enough to give every warp a few batches.
constexpr long long kStreamMaxQ = 8;
static bool stream_applies(long long N, long long Q) {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("EBSD_TOPK_STREAM");   // 0: keep the tiled kernel (A/B timing)
        on = (e && atoi(e) == 0) ? 0 : 1;
    }
    return on == 1 && Q <= kStreamMaxQ && N >= 32768;
}
static int stream_tq(long long Q) { return Q <= 1 ? 1 : (Q <= 2 ? 2 : (Q <= 4 ? 4 : 8)); }
// CTAs of the stream kernel (= partial lists per query): every resident slot, but at least 4 batches per warp
static int stream_grid(long long N, long long Q, int sms) {
    const int slots = sms * (stream_tq(Q) <= 2 ? 3 : 2);
    const long long by_rows = N / (kWarps * kStreamU * 32 * 4);
    return (int)(by_rows < 1 ? 1 : (by_rows < slots ? by_rows : slots));
}
static size_t exact_workspace_bytes(long long N, long long Q, int k, int sms) {
    if (stream_applies(N, Q)) return (size_t)stream_grid(N, Q, sms) * (size_t)Q * (size_t)k * sizeof(Entry);
    const TopkPlan pl = make_plan(N, Q, sms);
    return pl.n_splits > 1 ? (size_t)pl.n_splits * (size_t)Q * (size_t)k * sizeof(Entry) : 0;
}

template <int TQ>
static int launch_stream(const float *dict, const TopkParams &p, int grid, long long rows_per_warp, cudaStream_t st) {
    topk_stream_kernel<TQ><<<grid, kThreads, 0, st>>>(dict, p, rows_per_warp);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

static int run_stream(const float *dict, long long N, long long index_base, const float *queries, long long Q, int k,
                      float *out_dot, long long *out_idx, float *out_dist, void *workspace, int sms, cudaStream_t st) {
    const int grid = stream_grid(N, Q, sms);
    const long long warps = (long long)grid * kWarps;
    const long long batch = kStreamU * 32;
    const long long rows_per_warp = ((N + warps - 1) / warps + batch - 1) / batch * batch;
    TopkParams p;
    p.queries = queries;
    p.Q = Q;
    p.N = N;
    p.index_base = index_base;
    p.k = k;
    p.n_qtiles = 1;
    p.n_splits = grid;
    p.tiles_per_split = 0;
    p.parts = (Entry *)workspace;
    p.out_dot = out_dot;
    p.out_idx = out_idx;
    p.out_dist = out_dist;
    const int tq = stream_tq(Q);
    int rc = tq == 1 ? launch_stream<1>(dict, p, grid, rows_per_warp, st)
                     : (tq == 2 ? launch_stream<2>(dict, p, grid, rows_per_warp, st)
                                : (tq == 4 ? launch_stream<4>(dict, p, grid, rows_per_warp, st)
                                           : launch_stream<8>(dict, p, grid, rows_per_warp, st)));
    if (rc) return rc;
    topk_merge_parts_kernel<<<(unsigned)Q, kThreads, 0, st>>>(p.parts, grid, Q, k, index_base, out_dot, out_idx, out_dist);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

// The exact CUDA-core search of `N` rows (workspace: n_splits * Q * k entries when the plan splits the dictionary).
static int run_exact(const float *dict, long long N, long long index_base, const float *queries, long long Q, int k,
                     float *out_dot, long long *out_idx, float *out_dist, void *workspace, int sms, cudaStream_t st) {
    if (stream_applies(N, Q)) return run_stream(dict, N, index_base, queries, Q, k, out_dot, out_idx, out_dist, workspace, sms, st);
    const TopkPlan pl = make_plan(N, Q, sms);
    tensormap_encode_fn encode = get_tensormap_encode();
    if (!encode) {
        set_error("ebsd_topk: cuTensorMapEncodeTiled entry point not available");
        return EBSD_ERR_CUDA;
    }
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)kD, (cuuint64_t)N};
    const cuuint64_t gstride[1] = {(cuuint64_t)(kD * 4)};
    const cuuint32_t box[2] = {(cuuint32_t)kD, (cuuint32_t)kTileRows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)dict, gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("ebsd_topk: cuTensorMapEncodeTiled failed with %d", (int)cr);
        return EBSD_ERR_CUDA;
    }
    TopkParams p;
    p.queries = queries;
    p.Q = Q;
    p.N = N;
    p.index_base = index_base;
    p.k = k;
    p.n_qtiles = pl.n_qtiles;
    p.n_splits = pl.n_splits;
    p.tiles_per_split = pl.tiles_per_split;
    p.parts = (Entry *)workspace;
    p.out_dot = out_dot;
    p.out_idx = out_idx;
    p.out_dist = out_dist;
    int rc = pl.tq == 8 ? launch_topk<8>(map, p, sms, st)
                        : (pl.tq == 4 ? launch_topk<4>(map, p, sms, st) : launch_topk<1>(map, p, sms, st));
    if (rc) return rc;
    if (pl.n_splits > 1) {
        topk_merge_parts_kernel<<<(unsigned)Q, kThreads, 0, st>>>(p.parts, pl.n_splits, Q, k, index_base, out_dot, out_idx,
                                                                 out_dist);
        EBSD_LAUNCH_CHECK();
    }
    return EBSD_OK;
}

static int run_screen(const float *dict, long long N, long long index_base, const float *queries, long long Q, int k,
                      float *out_dot, long long *out_idx, float *out_dist, void *workspace, size_t workspace_bytes, int sms,
                      cudaStream_t st) {
    const ScreenPlan pl = make_screen_plan(N, Q, sms);
    // every chunk of a long query batch has its own plan (a short last chunk splits the seeding search more ways)
    if (workspace == nullptr || workspace_bytes < pl.bytes) {
        set_error("ebsd_topk: workspace too small (%zu < %zu)", workspace_bytes, pl.bytes);
        return EBSD_ERR_WORKSPACE;
    }
    uint8_t *ws = (uint8_t *)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
    __half *dpairs = (__half *)(ws + pl.off_dpairs);
    __half *qpairs = (__half *)(ws + pl.off_qpairs);
    const long long total_tiles = (N + kScrN - 1) / kScrN;
    const long long drows = total_tiles * kScrN, qrows = (long long)pl.n_qtiles * kScrM;
    // operands as fp16 (hi | lo) pairs; the padding rows of the last tile are zero
    EBSD_CUDA_TRY(cudaMemsetAsync(dpairs + N * 32, 0, (size_t)(drows - N) * kScrRowB, st));
    EBSD_CUDA_TRY(cudaMemsetAsync(qpairs + Q * 32, 0, (size_t)(qrows - Q) * kScrRowB, st));
    split_rows_f16_kernel<<<(unsigned)((N * 16 + 255) / 256), 256, 0, st>>>(dict, dpairs, N);
    EBSD_LAUNCH_CHECK();
    split_rows_f16_kernel<<<(unsigned)((Q * 16 + 255) / 256), 256, 0, st>>>(queries, qpairs, Q);
    EBSD_LAUNCH_CHECK();
    CUtensorMap map_d, map_q;
    int rc;
    if ((rc = make_pairs_map(&map_d, dpairs, drows, kScrN))) return rc;
    if ((rc = make_pairs_map(&map_q, qpairs, qrows, kScrM))) return rc;
    // Thresholds are seeded in stages, each giving k rows whose exact dots bound the final k-th best from below:
    //   seed    CUDA-core search of the first few thousand rows                  -> exact top-k of [0, n0)
    //   pass 0  screen + re-rank of the first 8 n0 rows, thr from the seed        -> exact top-k of that prefix
    //   pass i  screen of the next 8x rows, thr from pass i-1; the re-rank starts from its list (~8k survivors each)
    // out_dot / out_idx carry the running exact lists between the stages (each kernel reads them before the next
    // one overwrites them: same stream).
    static bool configured[kMaxDevices] = {};   // function attributes are per device
    const int dev = current_device_slot();
    if (!configured[dev]) {
        EBSD_CUDA_TRY(cudaFuncSetAttribute(topk_screen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kScrSmem));
        configured[dev] = true;
    }
    const ScreenStages stages = screen_stages(N);
    if (stages.seed_rows <= kSeedRowsMax) {
        if (stages.seed_rows <= 256)
            topk_seed_kth_kernel<8><<<(unsigned)((Q + 7) / 8), 256, 0, st>>>(dict, (int)stages.seed_rows, queries, Q, k, out_dot);
        else
            topk_seed_kth_kernel<16><<<(unsigned)((Q + 7) / 8), 256, 0, st>>>(dict, (int)stages.seed_rows, queries, Q, k, out_dot);
        EBSD_LAUNCH_CHECK();
    } else if ((rc = run_exact(dict, stages.seed_rows, index_base, queries, Q, k, out_dot, out_idx, nullptr, ws + pl.off_parts,
                               sms, st))) {
        return rc;
    }
    auto pass = [&](long long tile_begin, long long tile_end, int init_from_out) -> int {
        ScreenParams p;
        p.Q = Q;
        p.N = N;
        p.k = k;
        p.n_qtiles = pl.n_qtiles;
        p.tile_begin = tile_begin;
        p.tile_end = tile_end;
        const long long units = (long long)p.n_qtiles * (tile_end - tile_begin);
        const int grid = units < sms ? (int)(units < 1 ? 1 : units) : sms;
        p.n_ctas = grid;
        p.span_base = units > 0 ? units / grid : 0;
        p.span_rem = units > 0 ? (int)(units % grid) : 0;
        p.tau0 = out_dot;
        p.cand_s = (float *)(ws + pl.off_cs);
        p.cand_i = (int *)(ws + pl.off_ci);
        p.cand_n = (int *)(ws + pl.off_cn);
        topk_screen_kernel<<<grid, kScrThreads, kScrSmem, st>>>(map_d, map_q, p);
        EBSD_LAUNCH_CHECK();
        const int wpb = 8;
        topk_rerank_kernel<<<(unsigned)((Q + wpb - 1) / wpb), wpb * 32, 0, st>>>(dict, queries, p, index_base, init_from_out,
                                                                              out_dot, out_idx, out_dist);
        EBSD_LAUNCH_CHECK();
        return EBSD_OK;
    };
    long long begin = 0;
    for (int i = 0; i < stages.n_pass; ++i) {
        if (stages.pass_end[i] <= begin) continue;
        if ((rc = pass(begin, stages.pass_end[i], i > 0 ? 1 : 0))) return rc;
        begin = stages.pass_end[i];
    }
    return EBSD_OK;
}

// Workspace of the screen path for Q queries: ebsd_topk runs it in chunks of kScreenQueryChunk queries and every
// chunk size has its own plan -- the largest of the (at most two) distinct plans counts, not the first chunk's.
static size_t screen_workspace_bytes(long long N, long long Q, int sms) {
    size_t need = 0;
    if (Q >= kScreenQueryChunk) need = make_screen_plan(N, kScreenQueryChunk, sms).bytes;
    const long long rem = Q % kScreenQueryChunk;
    if (rem > 0) {
        const size_t b = make_screen_plan(N, rem, sms).bytes;
        if (b > need) need = b;
    }
    return need;
}

// ---- candidate exchange between shards: one 64-bit word per candidate = (dot bits << 32) | (global row + 1), so that
// the per-shard lists travel in ONE collective; rows fit 32 bits for dictionaries below 2^32 - 1 rows.
__global__ void pack_candidates_kernel(const float *__restrict__ dot, const long long *__restrict__ idx, long long n,
                                       unsigned long long *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long r = idx[i];
    out[i] = ((unsigned long long)__float_as_uint(dot[i]) << 32) | (unsigned long long)(unsigned)(r < 0 ? 0u : (unsigned)(r + 1));
}

// k-way merge of R packed lists [R,Q,k] (order: dot descending, global row ascending), one warp per query
__global__ void topk_merge_packed_kernel(const unsigned long long *__restrict__ packed, int R, long long Q, int k,
                                         float *out_dot, long long *out_idx, float *out_dist) {
    const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= Q) return;
    float e_dot = -INFINITY;
    long long e_idx = 0x7fffffffffffffffll;
    for (int r = 0; r < R; ++r) {
        const long long base = ((long long)r * Q + q) * k;
        for (int j = 0; j < k; ++j) {
            const unsigned long long w = packed[base + j];
            const unsigned row1 = (unsigned)(w & 0xffffffffull);
            if (row1 == 0u) break;
            warp_insert<long long>(e_dot, e_idx, __uint_as_float((unsigned)(w >> 32)), (long long)row1 - 1, lane);
        }
    }
    if (lane < k) {
        const bool empty = e_idx == 0x7fffffffffffffffll;
        out_dot[q * k + lane] = e_dot;
        out_idx[q * k + lane] = empty ? -1ll : e_idx;
        if (out_dist) out_dist[q * k + lane] = 1.0f - e_dot;
    }
}

}  // namespace ebsd

using namespace ebsd;

extern "C" {

int ebsd_normalize_rows(float *x, int64_t n, int d, void *stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    EBSD_REQUIRE(d == kD, "ebsd_normalize_rows: d must be %d, got %d", kD, d);
    EBSD_REQUIRE(n >= 0, "ebsd_normalize_rows: negative n");
    EBSD_REQUIRE(n == 0 || x != nullptr, "ebsd_normalize_rows: null pointer");
    EBSD_REQUIRE(((uintptr_t)x & 15) == 0, "ebsd_normalize_rows: x must be 16-byte aligned");
    if (n == 0) return EBSD_OK;
    const int threads = 256;
    normalize_rows_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(x, n);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

size_t ebsd_topk_workspace_bytes(int64_t N, int64_t Q, int k) {
    if (N <= 0 || Q <= 0 || k <= 0) return 0;
    if (screen_applies(N, Q, k)) return screen_workspace_bytes(N, Q, sm_count());
    return exact_workspace_bytes(N, Q, k, sm_count());
}

int ebsd_topk(const float *dict, int64_t N, int64_t index_base, const float *queries, int64_t Q, int k,
              float *out_dot, int64_t *out_idx, float *out_dist, void *workspace, size_t workspace_bytes,
              void *stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    EBSD_REQUIRE(k >= 1 && k <= EBSD_MAX_TOPK, "ebsd_topk: k must be in [1,%d], got %d", EBSD_MAX_TOPK, k);
    EBSD_REQUIRE(N >= 0 && Q >= 0, "ebsd_topk: negative size");
    EBSD_REQUIRE(N < 0x7fffff00ll, "ebsd_topk: a shard holds at most 2^31-257 rows");
    if (Q == 0) return EBSD_OK;
    EBSD_REQUIRE(queries && out_dot && out_idx, "ebsd_topk: null pointer");
    EBSD_REQUIRE(((uintptr_t)queries & 15) == 0, "ebsd_topk: queries must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {
        const int wpb = 8;
        topk_merge_global_kernel<<<(unsigned)((Q + wpb - 1) / wpb), wpb * 32, 0, st>>>(
            nullptr, nullptr, 0, Q, k, out_dot, (long long *)out_idx, out_dist);
        EBSD_LAUNCH_CHECK();
        return EBSD_OK;
    }
    EBSD_REQUIRE(dict != nullptr, "ebsd_topk: null dictionary");
    EBSD_REQUIRE(((uintptr_t)dict & 15) == 0, "ebsd_topk: dict must be 16-byte aligned");

    const int sms = sm_count();
    if (screen_applies(N, Q, k)) {
        const size_t need_s = screen_workspace_bytes(N, Q, sms);
        if (workspace == nullptr || workspace_bytes < need_s) {
            set_error("ebsd_topk: workspace too small (%zu < %zu)", workspace_bytes, need_s);
            return EBSD_ERR_WORKSPACE;
        }
        // very large query batches go through in chunks: the survivor buffers scale with the number of query tiles
        for (long long q0 = 0; q0 < Q; q0 += kScreenQueryChunk) {
            const long long qn = Q - q0 < kScreenQueryChunk ? Q - q0 : kScreenQueryChunk;
            if ((rc = run_screen(dict, N, index_base, queries + q0 * kD, qn, k, out_dot + q0 * k, (long long *)out_idx + q0 * k,
                                 out_dist ? out_dist + q0 * k : nullptr, workspace, workspace_bytes, sms, st)))
                return rc;
        }
        return EBSD_OK;
    }
    const size_t need = exact_workspace_bytes(N, Q, k, sms);
    if (need > 0 && (workspace == nullptr || workspace_bytes < need)) {
        set_error("ebsd_topk: workspace too small (%zu < %zu)", workspace_bytes, need);
        return EBSD_ERR_WORKSPACE;
    }
    return run_exact(dict, N, index_base, queries, Q, k, out_dot, (long long *)out_idx, out_dist, workspace, sms, st);
}

int ebsd_topk_merge(const float *dots, const int64_t *idx, int R, int64_t Q, int k, float *out_dot,
                    int64_t *out_idx, float *out_dist, void *stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    EBSD_REQUIRE(k >= 1 && k <= EBSD_MAX_TOPK, "ebsd_topk_merge: k must be in [1,%d], got %d", EBSD_MAX_TOPK, k);
    EBSD_REQUIRE(R >= 0 && Q >= 0, "ebsd_topk_merge: negative size");
    if (Q == 0) return EBSD_OK;
    EBSD_REQUIRE(out_dot && out_idx && (R == 0 || (dots && idx)), "ebsd_topk_merge: null pointer");
    const int wpb = 8;
    topk_merge_global_kernel<<<(unsigned)((Q + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        dots, (const long long *)idx, R, Q, k, out_dot, (long long *)out_idx, out_dist);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

int ebsd_topk_pack(const float *dots, const int64_t *idx, int64_t n, uint64_t *packed, void *stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    EBSD_REQUIRE(n >= 0, "ebsd_topk_pack: negative size");
    if (n == 0) return EBSD_OK;
    EBSD_REQUIRE(dots && idx && packed, "ebsd_topk_pack: null pointer");
    pack_candidates_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        dots, (const long long *)idx, n, (unsigned long long *)packed);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

int ebsd_topk_merge_packed(const uint64_t *packed, int R, int64_t Q, int k, float *out_dot, int64_t *out_idx,
                           float *out_dist, void *stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    EBSD_REQUIRE(k >= 1 && k <= EBSD_MAX_TOPK, "ebsd_topk_merge_packed: k must be in [1,%d], got %d", EBSD_MAX_TOPK, k);
    EBSD_REQUIRE(R >= 0 && Q >= 0, "ebsd_topk_merge_packed: negative size");
    if (Q == 0) return EBSD_OK;
    EBSD_REQUIRE(out_dot && out_idx && (R == 0 || packed), "ebsd_topk_merge_packed: null pointer");
    const int wpb = 8;
    topk_merge_packed_kernel<<<(unsigned)((Q + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        (const unsigned long long *)packed, R, Q, k, out_dot, (long long *)out_idx, out_dist);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

}  // extern "C"
