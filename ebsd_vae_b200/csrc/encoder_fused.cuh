// K1: one kernel per convolution block that
//   (1) BUILDS its own tensor-core operands in shared memory: "producer" warps read the previous block's raw fp32
//       output (or, for the first tensor-core block, the uint8 pattern itself and run conv0 on CUDA cores), apply
//       InstanceNorm + LeakyReLU(0.02) (latice/model.py:93-98) with the plane statistics the previous block left
//       behind, and store the result in the swizzled K-major layouts tcgen05.mma reads (see "Arithmetic");
//   (2) runs the 3x3 convolution as nine row-SHIFTED views of that window (implicit GEMM, fp32 accumulation in TMEM);
//   (3) finishes in the epilogue warps: TMEM -> registers, plane statistics of the un-pooled output (sum, sum of
//       squares, fp64 atomics), optional 2x2 max-pool with warp shuffles (pooling commutes with the increasing map
//       x -> leaky((x-mean)*rstd), so it is applied to the raw values), raw fp32 NHWC store.
// Nothing but the (pooled) raw output and 2 numbers per (image, channel) goes to memory between blocks: there are no
// finisher kernels, no fp16 planes in HBM/L2 and no TMA window re-reads.
//
// Arithmetic.  The latents must match torch fp32 within 1e-3 relative; a single fp16 (or bf16 / tf32) product per MAC
// misses that by 3-5x, the exact three-term fp16 split a_hi*w_hi + a_hi*w_lo + a_lo*w_hi costs
// three tensor-core passes.  Here each MAC is ONE fp16 pass plus ONE fp8 pass (fp8 runs at twice the fp16 rate, so
// two units of tensor time instead of three):
//     a*w  ~=  fp16(a) * fp16(w)                                             kind::f16,    N = Cout, K = 16 per MMA
//            + [ e4m3(a) , e4m3(4096 (a - fp16 a)) ] . [ e4m3(4096 s (w - fp16 w)) ; e4m3(s w) ] / (4096 s)
//                                                                            kind::f8f6f4, N = Cout, K = 32 per MMA
// The two correction products are first-order error terms (2^-12 relative to a*w), so the 2^-4 relative rounding of
// their e4m3 operands leaves 2^-16 per product: measured 1.2e-4 relative on the latents through all ten blocks
// (tests/test_gpu_encoder.py; tools/emulate_split.py is the CPU emulation the design was chosen with), against 6e-6
// for the three-term split and 3e-3 for one fp16 pass.  s is a per-layer power of two that brings max|w| into
// (64, 128].  The fp16 products accumulate in TMEM columns [0, Cout), the scaled fp8 products in [Cout, 2 Cout); the
// epilogue adds them as main + corr / (4096 s).  Per (tap, K chunk) the operand bytes are what the three-term split
// needed -- an fp16 window and an equally sized window of (a, residual) byte pairs; [w_fp16 ; w_fp8] weight rows.
//
// Tile geometry.  A tile is 128 output positions = 16 groups of 8 horizontally adjacent pixels.  A UMMA K-major
// operand is 16 eight-row groups at a constant stride (SBO), and tools/probe_umma_desc.cu shows that stride may be
// any multiple of 16 bytes while the swizzle is applied to absolute shared-memory address bits.  So:
//   * W >= 16: tile = 16 image rows x 8 columns.  The window is [18][8*NT + 2] positions (NT tiles side by side,
//     1-pixel halo); group g of tile t under tap (dy,dx) starts at row (g + dy) * PITCH + 8 t + dx: stride PITCH.
//   * W == 8:  tile = 2 images x 8 rows x 8 columns, window [10][2 images][10]; group j = 2 y + image starts at
//     (y + dy) * 20 + image * 10 + dx = 10 j + ...: stride 10.
// Zero padding is simply zeros the producers write at halo positions outside the image.
//
// Warp roles (512 threads): warp 0 weight TMA, warp 1 MMA issuer, warp 2 TMEM allocator, warp 3 conv0 patch staging
// (front-end block only), warps 4-7 epilogue
// (TMEM lane quarter = warp & 3), warps 8-15 producers (two per scheduler: global-load and ALU latency overlap), joined
// by warps 2 and 3 in the blocks that read a raw plane.
#pragma once
#include "encoder_aux.cuh"
#include "tcgen05.cuh"

namespace ebsd {

enum FusedSrc { SRC_U8 = 0, SRC_F32 = 1, SRC_RAW = 2 };

// 1 = the blocks whose weights are streamed run as CTA pairs (tcgen05 cta_group::2, see FusedCfg::PAIR); measured
// +4 % on the whole encoder against single CTAs (profiles/README.md).  A four-term variant of the pair (one B region,
// N = 2*COUT for both planes) was tried first and lost: under tensor load the chip is power-limited and the extra
// term costs more than the halved weight traffic saves.  0 = single CTAs (make EXTRA=-DEBSD_PAIR=0).
// tiles per work item of the 128x128 front-end block: 4 (two window stages of 18 x 34 positions) or 2 (four stages of
// 18 x 18: deeper producer -> MMA decoupling, 6 % more halo work -- measured slower, 1911 vs 1491 us per 1184 patterns)
#ifndef EBSD_FRONT_NT
#define EBSD_FRONT_NT 4
#endif
// staging boxes per epilogue warp of the un-pooled blocks (4 KiB each): a warp waits for the TMA store issued NSB
// blocks earlier to have read its box before it refills it
#ifndef EBSD_NSB_RESIDENT
#define EBSD_NSB_RESIDENT 2
#endif
#ifndef EBSD_NSB_STREAMED
#define EBSD_NSB_STREAMED 1
#endif
#ifndef EBSD_PAIR
#define EBSD_PAIR 1
#endif

template <int CIN_, int COUT_, int W_, int SRC_, bool POOL_>
struct FusedCfg {
    static constexpr int CIN = CIN_, COUT = COUT_, W = W_, SRC = SRC_;
    static constexpr bool POOL = POOL_;                      // 2x2 max-pool of the raw output in the epilogue
    static constexpr bool FIRST = SRC_ != SRC_RAW;           // conv0 is computed by the producers (CIN = 32, W = 128)
    static constexpr int NI = W == 8 ? 2 : 1;                // images interleaved in one window row
    // CTA pairs (tcgen05 cta_group::2, see below) for the blocks with >= 64 input channels.  The two 32-channel blocks
    // were measured slower as pairs (1490 -> 1568 and 543 -> 627 us per 1184 patterns): their MMA streams are short and
    // the lock step of two producer groups costs more than the halved B reads save.
    static constexpr bool PAIR = EBSD_PAIR != 0 && CIN >= 64;
    static constexpr int KC = CIN < 64 ? CIN : 64;
    static constexpr int ROWB = KC * 2;
    static constexpr int NCHUNK = CIN / KC;
    static constexpr int B_TILE = 2 * COUT * ROWB;           // [w_fp16; w_fp8] of one (tap, K chunk)
    // per (tap, K chunk) a CTA holds X = the fp16 rows and Y = the fp8 rows of the weights, COUT rows each -- or,
    // PAIR, its half of either (COUT/2 rows: cta_group::2 splits the N rows of B between the two CTAs)
    static constexpr int B_X = (PAIR ? COUT / 2 : COUT) * ROWB;
    static constexpr int B_Y = B_X;
    static constexpr int B_CTA = B_X + B_Y;                  // weight bytes one CTA holds per (tap, K chunk)
    // all nine taps stay in shared memory when they fit next to two windows: the 32-channel blocks always, the
    // 64 -> 64 block only as a pair (108 KB per CTA) and with one tile per window
    static constexpr bool RESIDENT_B = 9 * NCHUNK * B_CTA <= (PAIR ? 112 : 80) * 1024;
    static constexpr int NT = W >= 128 ? EBSD_FRONT_NT : (W >= 64 ? ((RESIDENT_B && CIN == 64) ? 1 : 2) : 1);  // tiles per work item
    static constexpr int TR = 16 / NI;                       // image rows per tile
    static constexpr int WIN_H = TR + 2;
    static constexpr int PITCH = NI == 1 ? 8 * NT + 2 : 10 * NI;
    static constexpr int GSTRIDE = NI == 1 ? PITCH : 10;     // window rows between consecutive 8-row groups
    static constexpr int WIN_POS = WIN_H * PITCH;
    static constexpr int SWMASK = ROWB == 128 ? 7 : 3;
    static constexpr int KSTEPS = KC / 16;
    static constexpr int A_PLANE = (WIN_POS * ROWB + 1023) / 1024 * 1024;
    static constexpr int A_STAGE = 2 * A_PLANE;              // fp16 window + fp8 (value, residual) window
    // CTA PAIRS (tcgen05 cta_group::2, M = 256): each CTA builds the window of its own work item and holds HALF of
    // every weight tile; the leader issues one MMA stream for both.  25 % fewer weight bytes per SM and half the B
    // operand reads of the tensor core, which matters because the single-CTA blocks are shared-memory-bandwidth bound.
    static constexpr int CL = PAIR ? 2 : 1;                  // cluster size
    static constexpr int B_BOX_ROWS = PAIR ? COUT / 2 : 2 * COUT;  // rows of one weight TMA box
    static constexpr int A_STAGES = (W >= 128 && EBSD_FRONT_NT == 2) ? 4 : 2;
    static constexpr int PATCH_W = 8 * NT + 4, PATCH_H = 20;   // FIRST: input pixels a window needs (conv0 + conv1 halos)
    // row stride of the patch in shared memory.  A producer warp reads 32 consecutive window positions, which wrap to
    // the next window row after PITCH of them; with a stride of PITCH + 32 words the lanes after the wrap stay on the
    // banks they would have had without it (stride 36: the last two lanes collided with the first two on nearly every
    // load, 1.8 wavefronts per LDS, ncu)
    static constexpr int PATCH_S = PITCH + 32;
    // output staging for the TMA stores: per epilogue warp one [32 or 8 rows][128 B] box, 128B-swizzled
    static constexpr int WSTG = POOL ? 1024 : 4096;
    // COUT = 32 with several tiles per item (the front end): the epilogue walks 16-channel halves outside the tile
    // loop so that the plane sums stay in registers across the NT tiles; every tile then needs its own box
    static constexpr bool ACCUM = COUT == 32 && NT > 1;
    static constexpr int NSB = ACCUM ? NT : (!POOL ? (RESIDENT_B ? EBSD_NSB_RESIDENT : EBSD_NSB_STREAMED) : 1);  // staging boxes per warp
    static constexpr int STAGING = 4 * NSB * WSTG;
    // pooled blocks: a [32 pixels][32 channels] fp32 scratch box per epilogue warp holding the UN-pooled values of the
    // current (tile, channel block), so that the plane statistics are column sums read back with LDS (see the epilogue)
#ifndef EBSD_POOL_STAT_LDS
#define EBSD_POOL_STAT_LDS 1
#endif
    static constexpr bool STAT_LDS = EBSD_POOL_STAT_LDS != 0 && POOL && !ACCUM;
    static constexpr int STAT_SCRATCH = STAT_LDS ? 4 * 4096 : 0;
    // barriers + tables (+ FIRST: two conv0 patches of PATCH_BYTES, written by a helper warp one item ahead) | staging
    static constexpr int PATCH_BYTES = (PATCH_H * PATCH_S * 4 + 127) / 128 * 128;
    static constexpr int XBASE = FIRST ? (2560 + 2 * PATCH_BYTES + 1023) / 1024 * 1024 : 8192;
    static constexpr int EXTRA = XBASE + STAGING + STAT_SCRATCH;
    static constexpr int B_FIT = (226 * 1024 - 1024 - EXTRA - A_STAGES * A_STAGE) / B_CTA;
    static constexpr int B_STAGES = RESIDENT_B ? 9 * NCHUNK : (B_FIT > 8 ? 8 : B_FIT);
    static constexpr int B_BYTES = B_STAGES * B_CTA;
    static constexpr int ACC_COLS = NT * 2 * COUT;           // TMEM columns of one work item
    static constexpr int TMEM_COLS = 512;
    static constexpr int SMEM_BYTES = 1024 + A_STAGES * A_STAGE + B_BYTES + EXTRA;
    static constexpr int THREADS = 512;
    // producer threads: warps 8..15, plus the otherwise idle warps 2 and 3 in the blocks that read a raw plane (in the
    // front-end block warp 3 stages the conv0 patch and the channel-group mapping needs a multiple of four warps)
#ifndef EBSD_EXTRA_PRODUCERS
#define EBSD_EXTRA_PRODUCERS 1
#endif
    static constexpr int PRODUCERS = (FIRST || !EBSD_EXTRA_PRODUCERS) ? 256 : 320;
    static constexpr int ITEMS_X = NI == 1 ? W / (8 * NT) : 1;
    static constexpr int ITEMS_PER_IMAGE = NI == 1 ? (W / 16) * ITEMS_X : 1;   // NI == 2: one item = 2 images
    static_assert(2 * ACC_COLS <= TMEM_COLS, "TMEM budget");
    static_assert(RESIDENT_B || B_STAGES >= 2, "weight ring does not fit");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(!FIRST || (CIN == 32 && W == 128), "conv0 fusion is for the 1->32->32 @128x128 front end");
};

struct FusedParams {
    const void *src;         // SRC_RAW: fp32 [nimg,W,W,CIN] raw output of the previous block (pooled to this block's
                             //          size); SRC_U8 / SRC_F32: patterns [nimg,128,128]
    const double *src_sums;  // [nimg,CIN,2] plane sums (sum, sum of squares) of the block that produced src
    double inv_src_plane;    // 1 / number of pixels those sums run over
    const float *w0;         // FIRST: conv0 weights [tap][32] fp32
    double *sums;            // out: [nimg,COUT,2], must be zero on entry
    float corr_scale;        // 1 / (4096 * weight scale): brings the fp8 correction sum to the scale of the fp16 sum
    int nimg;
    int nitems;
#ifdef EBSD_ROLE_PROFILE
    int dbg;                 // role-profiling build only (tools/time_fused.py): 1 producers write nothing, 2 no MMAs,
                             // 4 epilogue does nothing but release TMEM, 8 no plane statistics, 16 no stores,
                             // 32 statistics without the fp64 accumulation (W >= 16 blocks with register sums)
#endif
};
#ifdef EBSD_ROLE_PROFILE
#define EBSD_DBG(p) ((p).dbg)
#else
#define EBSD_DBG(p) 0
#endif

template <int ROWB, int GROUP_ROWS>
__device__ __forceinline__ uint64_t umma_smem_desc_g(uint32_t saddr) {
    constexpr uint64_t layout = ROWB == 128 ? 2ull : 4ull;
    constexpr uint64_t sbo = ((uint64_t)GROUP_ROWS * ROWB) >> 4;
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, uint32_t smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
                 "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// Fire-and-forget L2 prefetch of a 4-D box (out-of-range coordinates are simply skipped by the TMA unit).
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap *map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1),
                 "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) helpers.  `bar` arguments are shared-window addresses of the issuing CTA; the leader
// (cluster rank 0) owns the barriers the MMA issuer waits on, so followers signal the leader's copy.
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
// wait with cluster-scope acquire: the arrivals come from the peer CTA's threads (bounded like mbar_wait_bounded)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(EBSD_WAIT_HINT_NS)
            : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) {
            printf("ebsd encoder: pair mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// this CTA's half of a weight tile; the transaction bytes are credited to the LEADER's barrier (cluster address)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const CUtensorMap *map, int c0, int c1,
                                                 uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(smem_dst), "l"(map), "r"(c0), "r"(c1), "r"(leader_bar)
        : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f8_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at the same offset in BOTH CTAs of the pair once the preceding MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__host__ __device__ constexpr uint32_t umma_idesc_f16_m256(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Shared-memory accesses by 32-bit shared-window address: pointers derived from the aligned dynamic-smem base lose
// their address space and would compile to generic LD/ST with 64-bit address arithmetic.
__device__ __forceinline__ void sts128(uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// y = leaky(x * scale + shift) as fp16 plus the fp8 (value, residual) pairs of the correction product; eight values ->
// two 16-byte chunks (fp16 window: 8 halves; fp8 window: per channel pair the bytes a(k), a(k+1), res(k), res(k+1)).
// The table holds one float4 (scale_a, scale_b, shift_a, shift_b) per channel PAIR, laid out [pair j = 0..3][8-channel
// group] so that the eight lanes that handle the eight channel groups of one position read 128 contiguous bytes (no
// bank conflicts): pair j of group c8 sits at tab + j * jstride + c8 * 16.  tab_u32 already includes c8 * 16.
// Packed fp32 pairs (FFMA2 / FMUL2, sm_100) halve the issue slots of the arithmetic.
__device__ __forceinline__ void norm_split_pair(float2 x, const float4 &tt, __half2 &h, uint32_t &q) {
    const float2 y = __ffma2_rn(x, make_float2(tt.x, tt.y), make_float2(tt.z, tt.w));
    const float2 z = __fmul2_rn(y, make_float2(0.02f, 0.02f));
    const float2 a = make_float2(fmaxf(y.x, z.x), fmaxf(y.y, z.y));  // LeakyReLU(0.02)
    h = __floats2half2_rn(a.x, a.y);
    // exact: (a - fp16 a) * 4096
    const float2 d = __ffma2_rn(__half22float2(h), make_float2(-kResidualScale, -kResidualScale),
                                __fmul2_rn(a, make_float2(kResidualScale, kResidualScale)));
    q = pack_e4m3x2(a.x, a.y) | (pack_e4m3x2(d.x, d.y) << 16);
}
__device__ __forceinline__ void norm_split8(const float (&x)[8], uint32_t tab_u32, uint32_t jstride, uint4 &hi,
                                            uint4 &lo) {
    __half2 h[4];
    uint32_t l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) norm_split_pair(make_float2(x[2 * j], x[2 * j + 1]), lds128(tab_u32 + j * jstride), h[j], l[j]);
    hi = *(const uint4 *)h;
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// the same with the four table entries already in registers (a thread whose channel group never changes)
__device__ __forceinline__ void norm_split8_t(const float (&x)[8], const float4 (&tt)[4], uint4 &hi, uint4 &lo) {
    __half2 h[4];
    uint32_t l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) norm_split_pair(make_float2(x[2 * j], x[2 * j + 1]), tt[j], h[j], l[j]);
    hi = *(const uint4 *)h;
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <class C>
__device__ __forceinline__ void store_chunk(uint32_t stage_u32, int pos, int c8, const uint4 &hi, const uint4 &lo) {
    const uint32_t a_hi = stage_u32 + pos * C::ROWB + c8 * 16;
    const uint32_t a_lo = a_hi + C::A_PLANE;
    sts128(a_hi ^ (((a_hi >> 7) & C::SWMASK) << 4), hi);
    sts128(a_lo ^ (((a_lo >> 7) & C::SWMASK) << 4), lo);
}

// map_out: fp32 [nimg,Wo,Wo,COUT] raw output (Wo = W/2 when pooling), box = one epilogue warp's share of a tile
template <int CIN, int COUT, int W, int SRC, bool POOL>
__global__ void __launch_bounds__(512, 1)
conv3x3_fused_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out,
                     const __grid_constant__ CUtensorMap map_src, const FusedParams p) {
    // map_src (SRC_RAW only): the source raw tensor (c, x, y, n) with box = one window; used for L2 prefetches
    // map_w: box = [KC, 2*COUT / CL] rows of the packed weights (the whole tile when CL = 1)
    using C = FusedCfg<CIN, COUT, W, SRC, POOL>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *smem_b = smem + C::A_STAGES * C::A_STAGE;
    uint8_t *extra = smem_b + C::B_BYTES;
    uint64_t *a_full = (uint64_t *)extra;            // [A_STAGES]
    uint64_t *a_empty = a_full + C::A_STAGES;        // [A_STAGES]
    uint64_t *b_full = a_empty + C::A_STAGES;        // [B_STAGES] (<= 18)
    uint64_t *b_empty = b_full + C::B_STAGES;        // [B_STAGES]
    uint64_t *tfull_bar = b_empty + C::B_STAGES;     // [2]
    uint64_t *tempty_bar = tfull_bar + 2;            // [2]
    uint32_t *tmem_slot = (uint32_t *)(tempty_bar + 2);
    const uint32_t tab_u32 = smem_u32(extra + 512);     // float2 [NI][CIN] (scale, shift) of the source planes: <= 2 KB
    uint64_t *patch_full = (uint64_t *)(extra + 256);   // [2] FIRST: patch buffer written (helper warp -> producers)
    uint64_t *patch_empty = patch_full + 2;             // [2] FIRST: patch buffer read by all producer warps
    const uint32_t patch_base_u32 = smem_u32(extra + 2560);  // FIRST: 2 x float [PATCH_H][PATCH_S] input pixels
    static_assert(!C::FIRST || (2560 + 2 * C::PATCH_BYTES <= C::XBASE && C::PATCH_S >= C::PATCH_W), "conv0 patches do not fit");
    static_assert((4 * C::A_STAGES + 2 * C::B_STAGES + 4) * 8 + 4 <= 256, "barrier area");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    griddep_launch_dependents();   // the next block's CTAs may take SMs as soon as this grid's CTAs leave them

    // PAIR: signals towards the leader's MMA issuer are collected LOCALLY in each CTA (one CTA-scope arrival per producer
    // / epilogue warp) and the follower forwards each completed phase with ONE cluster-scope arrival from its otherwise
    // idle warp 1 (the "relay").  A release.cluster arrival compiles to MEMBAR.ALL.GPU + ERRBAR: issued by every producer
    // warp it waited for the global loads the warp had just put in flight for the next batch (ncu: 12 % membar + 12 % mio
    // stalls in the 64 -> 64 block).
    const uint32_t cta_rank = C::PAIR ? cluster_ctarank() : 0u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < C::A_STAGES; ++s) {
            mbar_init(&a_full[s], C::PAIR ? C::PRODUCERS / 32 + (cta_rank == 0 ? 1 : 0) : C::PRODUCERS);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < C::B_STAGES; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], (C::PAIR && cta_rank == 0) ? 5 : 4);  // PAIR leader: own epilogue warps + the relay
            if (C::FIRST) {
                mbar_init(&patch_full[b], 1);
                mbar_init(&patch_empty[b], C::PRODUCERS / 32);
            }
        }
        mbar_fence_init();
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_out);
    }
    if (C::PAIR) {
        __syncthreads();
        cluster_sync_all();  // both CTAs are resident and their barriers initialised before the pair allocates TMEM
        if (warp == 2) tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
    } else {
        if (warp == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    if (C::PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // contiguous item range per CTA: consecutive items belong to the same image, so plane statistics are
    // flushed once per image per warp instead of once per tile
    const int per_cta = (p.nitems + (int)gridDim.x - 1) / (int)gridDim.x;
    const int item_begin = (int)blockIdx.x * per_cta;
    // PAIR: both CTAs of a pair run the same number of items (one M = 256 MMA stream serves both); items
    // beyond nitems decode to images >= nimg, which every role already treats as "nothing to load or store".
    const int item_end = C::CL > 1 ? item_begin + per_cta
                                   : (item_begin + per_cta < p.nitems ? item_begin + per_cta : p.nitems);

    // one weight tile: a whole [w_fp16; w_fp8] box, or (PAIR) this CTA's halves of the two as boxes of COUT/2
    // rows, with the bytes of both CTAs credited to the leader's barrier
    auto load_weights = [&](uint8_t *dst_ptr, int kb, uint64_t *bar) {
        if (!C::PAIR) {
            mbar_expect_tx(bar, C::B_TILE);
            tma_load_2d(dst_ptr, &map_w, 0, kb * 2 * COUT, bar);
        } else {
            if (cta_rank == 0) mbar_expect_tx(bar, 2 * C::B_CTA);
            const int row0 = kb * 2 * COUT;
            const uint32_t dst = smem_u32(dst_ptr);
            const uint32_t lbar = map_to_cta(smem_u32(bar), 0);
            tma_load_2d_pair(dst, &map_w, 0, row0 + (int)cta_rank * (COUT / 2), lbar);
            tma_load_2d_pair(dst + C::B_X, &map_w, 0, row0 + COUT + (int)cta_rank * (COUT / 2), lbar);
        }
    };
    // Programmatic dependent launch: everything above (and the resident weights, which no kernel of the chain writes)
    // overlaps the tail of the previous block's kernel; nothing below the wait runs before that kernel has completed.
    if (C::RESIDENT_B && warp == 0) {
        if (elect_one_sync())
            for (int kb = 0; kb < 9 * C::NCHUNK; ++kb) load_weights(smem_b + kb * C::B_CTA, kb, &b_full[kb]);
        __syncwarp();
    }
    griddep_wait();

    if (warp == 0) {
        // ===================== weight loads (TMA) + L2 prefetch of the windows the producers will read
        if (elect_one_sync()) {
            // items ahead (the producers themselves run up to A_STAGES items ahead of the MMAs).  Measured on one box, whole
            // bench step, round 1: PF = 8: 223 k patterns/s, 4: 225-229 k, 2: 233 k, 1: 231-234 k, 0: 224-226 k -- windows
            // prefetched too early are evicted again by the blocks' own output stream before the producers read them.
            // Round 2 (leaner producers): PF = 1 beats 2 by 0.5 % on the step (292.5 vs 290.9 k; 32->64 block 451 -> 439 us),
            // 3 and 4 lose 1-10 % on the 64x64 blocks.
#ifndef EBSD_PF
#define EBSD_PF 1
#endif
            constexpr int PF = EBSD_PF;
            auto prefetch_item = [&](int item) {
                if (C::FIRST || item >= item_end || item >= p.nitems) return;
                int n, y0, x0;
                if (C::NI == 1) {
                    n = item / C::ITEMS_PER_IMAGE;
                    const int r = item - n * C::ITEMS_PER_IMAGE;
                    const int yb = r / C::ITEMS_X;
                    y0 = yb * 16 - 1;
                    x0 = (r - yb * C::ITEMS_X) * 8 * C::NT - 1;
                } else {
                    n = item * 2;
                    y0 = 0;
                    x0 = 0;
                }
                tma_prefetch_l2_4d(&map_src, 0, x0, y0, n);
            };
            for (int j = 0; j < PF + (C::RESIDENT_B ? C::A_STAGES : 0); ++j) prefetch_item(item_begin + j);
            if (C::RESIDENT_B) {
                if (!C::FIRST) {
                    // pace the prefetches with the windows being consumed (a_empty arrives in both CTAs of a pair)
                    unsigned ait = 0;
                    for (int item = item_begin; item < item_end; ++item)
                        for (int cc = 0; cc < C::NCHUNK; ++cc, ++ait) {
                            mbar_wait_bounded(&a_empty[ait % C::A_STAGES], (ait / C::A_STAGES) & 1u);
                            if (cc == 0) prefetch_item(item + PF + C::A_STAGES);
                        }
                }
            } else {
                unsigned bit = 0;
                for (int item = item_begin; item < item_end; ++item) {
                    prefetch_item(item + PF);
                    for (int cc = 0; cc < C::NCHUNK; ++cc)
                        for (int tap = 0; tap < 9; ++tap, ++bit) {
                            const int sb = bit % C::B_STAGES;
                            mbar_wait_bounded(&b_empty[sb], ((bit / C::B_STAGES) & 1u) ^ 1u);
                            load_weights(smem_b + sb * C::B_CTA, tap * C::NCHUNK + cc, &b_full[sb]);
                        }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (PAIR: the leader CTA issues for both)
        if (cta_rank == 0 && elect_one_sync()) {
            // fp16 window x fp16 weights (X region) -> columns [0, COUT) of the tile's accumulator; fp8 window x fp8
            // weights (Y region) -> columns [COUT, 2 COUT).  N = COUT for both; the instruction descriptor is the same
            // word for the two kinds (format code 0 = F16 / E4M3).  PAIR: M = 256 (this CTA's window rows + the
            // peer's) with the B rows split between the two CTAs.
            constexpr uint32_t idesc = C::PAIR ? umma_idesc_f16_m256(COUT) : umma_idesc_f16(COUT);
            auto mma = [&](bool f8, uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
                if (EBSD_DBG(p) & 2) return;
                if (f8) {
                    if (C::PAIR) umma_f8_pair(d, da, db, idesc, acc);
                    else umma_f8(d, da, db, idesc, acc);
                } else {
                    if (C::PAIR) umma_f16_pair(d, da, db, idesc, acc);
                    else umma_f16(d, da, db, idesc, acc);
                }
            };
            auto commit = [&](uint64_t *bar) {
                if (C::PAIR) umma_commit_pair(bar);
                else umma_commit(bar);
            };
            auto wait = [&](uint64_t *bar, uint32_t parity) {
                if (C::PAIR) mbar_wait_cluster(bar, parity);
                else mbar_wait_bounded(bar, parity);
            };
            // weight barriers complete by TMA transaction bytes only (both CTAs' loads are credited to the leader): no
            // cluster-scope acquire, which costs an L1 invalidation (CCTL.IVALL) per wait
            auto wait_tx = [&](uint64_t *bar, uint32_t parity) { mbar_wait_bounded(bar, parity); };
            // all MMAs of one plane for one tap: NT tiles x KSTEPS
            auto tap_mmas = [&](bool f8, uint32_t d_item, uint32_t win, uint32_t b_w, bool first) {
#pragma unroll
                for (int t = 0; t < C::NT; ++t)
#pragma unroll
                    for (int k = 0; k < C::KSTEPS; ++k)
                        mma(f8, d_item + t * 2 * COUT, umma_smem_desc_g<C::ROWB, C::GSTRIDE>(win + t * 8 * C::ROWB + k * 32),
                            umma_smem_desc_g<C::ROWB, 8>(b_w + k * 32), (first && k == 0) ? 0u : 1u);
            };
            if (C::RESIDENT_B) {
                for (int kb = 0; kb < 9 * C::NCHUNK; ++kb) wait_tx(&b_full[kb], 0);
                tc_fence_after();
            }
            unsigned ait = 0, bit = 0;
            int j = 0;
            for (int item = item_begin; item < item_end; ++item, ++j) {
                const int buf = j & 1;
                wait(&tempty_bar[buf], (((unsigned)j >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_item = tmem_base + (uint32_t)(buf * C::ACC_COLS);
                for (int cc = 0; cc < C::NCHUNK; ++cc, ++ait) {
                    const int sa = ait % C::A_STAGES;
                    wait(&a_full[sa], (ait / C::A_STAGES) & 1u);
                    tc_fence_after();
                    const uint32_t win_hi = smem_u32(smem + sa * C::A_STAGE);
                    const uint32_t win_lo = win_hi + C::A_PLANE;
                    if (C::RESIDENT_B) {
                        // all fp16 MMAs, then all fp8 MMAs: two kind switches per window
#pragma unroll 1
                        for (int tap = 0; tap < 9; ++tap) {
                            const int dy = tap / 3, dx = tap - dy * 3;
                            const uint32_t shift = (uint32_t)((dy * C::PITCH + dx) * C::ROWB);
                            const uint32_t b_w = smem_u32(smem_b + (tap * C::NCHUNK + cc) * C::B_CTA);
                            tap_mmas(false, d_item, win_hi + shift, b_w, (cc | tap) == 0);
                        }
#pragma unroll 1
                        for (int tap = 0; tap < 9; ++tap) {
                            const int dy = tap / 3, dx = tap - dy * 3;
                            const uint32_t shift = (uint32_t)((dy * C::PITCH + dx) * C::ROWB);
                            const uint32_t b_w = smem_u32(smem_b + (tap * C::NCHUNK + cc) * C::B_CTA);
                            tap_mmas(true, d_item + COUT, win_lo + shift, b_w + C::B_X, (cc | tap) == 0);
                        }
                    } else {
#pragma unroll 1
                        for (int tap = 0; tap < 9; ++tap, ++bit) {
                            const int dy = tap / 3, dx = tap - dy * 3;
                            const uint32_t shift = (uint32_t)((dy * C::PITCH + dx) * C::ROWB);
                            const int sb = bit % C::B_STAGES;
                            wait_tx(&b_full[sb], (bit / C::B_STAGES) & 1u);
                            tc_fence_after();
                            const uint32_t b_w = smem_u32(smem_b + sb * C::B_CTA);
                            tap_mmas(false, d_item, win_hi + shift, b_w, (cc | tap) == 0);
                            tap_mmas(true, d_item + COUT, win_lo + shift, b_w + C::B_X, (cc | tap) == 0);
                            commit(&b_empty[sb]);
                        }
                    }
                    commit(&a_empty[sa]);
                }
                commit(&tfull_bar[buf]);
            }
        }
        if (C::PAIR && cta_rank != 0 && elect_one_sync()) {
            // ===================== relay (follower CTA): forward locally completed phases to the leader's barriers
            const unsigned total_a = (unsigned)(item_end - item_begin) * C::NCHUNK, total_t = (unsigned)(item_end - item_begin);
            unsigned na = 0, nt = 0;
            const long long t0 = clock64();
            while (na < total_a || nt < total_t) {
                if (na < total_a && mbar_try_wait_hint(&a_full[na % C::A_STAGES], (na / C::A_STAGES) & 1u, 200u)) {
                    mbar_arrive_cluster(map_to_cta(smem_u32(&a_full[na % C::A_STAGES]), 0));
                    ++na;
                }
                if (nt < total_t && mbar_try_wait(&tempty_bar[nt & 1], (nt >> 1) & 1u)) {
                    mbar_arrive_cluster(map_to_cta(smem_u32(&tempty_bar[nt & 1]), 0));
                    ++nt;
                }
                if (clock64() - t0 > EBSD_TIMEOUT_CYCLES) EBSD_TIMEOUT_ACTION(smem_u32(&a_full[0]), na);
            }
        }
    } else if (warp == 3 && C::FIRST) {
        // ===================== FIRST: conv0 patch staging.  The 20 x 36 input pixels a window needs (conv0 halo on top
        // of the conv1 halo) go to one of two shared-memory patches, one item ahead of the producers, which used to do
        // this themselves between two 256-thread named barriers per item (10 % of their time in barrier stalls, ncu).
        if constexpr (C::FIRST) {
            constexpr int NPX = C::PATCH_H * C::PATCH_W, NPL = (NPX + 31) / 32;
            unsigned pit = 0;
            for (int item = item_begin; item < item_end; ++item, ++pit) {
                const int pb = pit & 1;
                const int n = item / C::ITEMS_PER_IMAGE;
                const int r = item - n * C::ITEMS_PER_IMAGE;
                const int yb = r / C::ITEMS_X;
                const int y0 = yb * 16, x0 = (r - yb * C::ITEMS_X) * 8 * C::NT;
                uint32_t pv[NPL];
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    const int idx = lane + 32 * i;
                    const int py = idx / C::PATCH_W, px = idx - py * C::PATCH_W;
                    const int gy = y0 - 2 + py, gx = x0 - 2 + px;
                    uint32_t v = 0u;
                    if (idx < NPX && n < p.nimg && gy >= 0 && gy < 128 && gx >= 0 && gx < 128) {
                        const long long off = ((long long)n * 128 + gy) * 128 + gx;
                        if (SRC == SRC_U8) v = __ldg((const uint8_t *)p.src + off);
                        else v = __float_as_uint(__ldg((const float *)p.src + off));
                    }
                    pv[i] = v;
                }
                mbar_wait_bounded(&patch_empty[pb], ((pit >> 1) & 1u) ^ 1u);
                const uint32_t dst = patch_base_u32 + (uint32_t)(pb * C::PATCH_BYTES);
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    const int idx = lane + 32 * i;
                    if (idx < NPX) {
                        const int py = idx / C::PATCH_W, px = idx - py * C::PATCH_W;
                        // ToTensor: uint8 -> float32, true division by 255 (latice/data_module.py:31)
                        const float v = SRC == SRC_U8 ? (float)pv[i] / 255.0f : __uint_as_float(pv[i]);
                        sts32(dst + (uint32_t)((py * C::PATCH_S + px) * 4), __float_as_uint(v));
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&patch_full[pb]);
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================== epilogue
        const int quarter = warp & 3;
        const int g = quarter * 4 + (lane >> 3);  // 8-row group of this lane inside the tile
        const int xl = lane & 7;
        constexpr int NCB = COUT / 32;
        constexpr int PV = C::NI == 1 ? 8 : 16;   // lane distance of the vertical pooling partner
        const int a_par = xl & 1;
        const int b_par = C::NI == 1 ? (g & 1) : ((g >> 1) & 1);
        const int im = C::NI == 1 ? 0 : (g & 1);  // image slot of this lane (NI == 2)
        const int yl = C::NI == 1 ? g : (g >> 1); // image row of this lane relative to the tile
        // staging rows of this lane: pooled pixel / un-pooled pixel in the warp's TMA box ([n][y][x] order)
        const int prow = C::NI == 1 ? ((lane >> 4) * 4 + (xl >> 1)) : (im * 4 + (xl >> 1));
        const int urow = C::NI == 1 ? lane : (im * 16 + ((lane >> 4) & 1) * 8 + xl);
        const uint32_t stg_u32 = smem_u32(extra + C::XBASE) + (uint32_t)(quarter * C::NSB * C::WSTG);
        const uint32_t scr_u32 = smem_u32(extra + C::XBASE + C::STAGING) + (uint32_t)(quarter * 4096);   // STAT_LDS
        auto scr_key = [](int l) { return ((l & 1) << 2) | ((l >> 1) & 3); };   // swizzle key of scratch row l
        // STAT_LDS pooling: this lane produces chunks 2 cq, 2 cq + 1 (8 channels) of pooled pixel pp = staging row pp;
        // its four source pixels are scratch rows (= lanes) src_k
        const int pool_pp = lane >> 2, pool_cq = lane & 3;
        uint32_t pool_row[4], pool_key[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int src = C::NI == 1 ? (2 * (pool_pp >> 2) + (k >> 1)) * 8 + 2 * (pool_pp & 3) + (k & 1)
                                       : (k >> 1) * 16 + (pool_pp >> 2) * 8 + 2 * (pool_pp & 3) + (k & 1);
            pool_row[k] = scr_u32 + (uint32_t)(src * 128);
            pool_key[k] = (uint32_t)scr_key(src);
        }
        int sbuf = 0;
        // Running plane sums in fp64: the per-tile fp32 partial sums are fixed by the tile, but WHICH tiles of an image a
        // CTA handles depends on where the image sits in the batch -- fp32 running sums made equal patterns differ by
        // ~1e-7 in their statistics, fp64 ones (and fp64 atomics) agree to the last bit in practice.
        // The 128-channel blocks would need 16-32 more registers for them: they add every tile's sums to memory directly
        // (TILE_FLUSH; 4-8x more fp64 reductions at the L2, still a few hundred per image).
        constexpr bool TILE_FLUSH = NCB >= 4;
        constexpr int NACC = TILE_FLUSH ? 1 : NCB;
        PairSum acc1[NACC][C::NI], acc2[NACC][C::NI];   // (hi, lo) fp32 pairs, fp64 only at the flush (encoder_aux.cuh)
#pragma unroll
        for (int cb = 0; cb < NACC; ++cb)
#pragma unroll
            for (int s = 0; s < C::NI; ++s) {
                acc1[cb][s].clear();
                acc2[cb][s].clear();
            }
        PairSum accum_lo, accum_hi;  // ACCUM: lane c < 16: sum of channel c (lo) / 16 + c (hi); lanes >= 16: squares
        accum_lo.clear();
        accum_hi.clear();
        int cur_n = -1;
        auto flush = [&]() {
            if (C::ACCUM) {
                if (cur_n >= 0 && cur_n < p.nimg) {
                    double *dst = p.sums + ((long long)cur_n * COUT + (lane & 15)) * 2 + (lane >> 4);
                    atomicAdd(dst, accum_lo.value());
                    atomicAdd(dst + 32, accum_hi.value());
                }
                accum_lo.clear();
                accum_hi.clear();
                return;
            }
            if (TILE_FLUSH) return;
            if (cur_n >= 0) {
#pragma unroll
                for (int s = 0; s < C::NI; ++s) {
                    if (cur_n + s < p.nimg) {
#pragma unroll
                        for (int cb = 0; cb < NACC; ++cb) {
                            double *dst = p.sums + ((long long)(cur_n + s) * COUT + cb * 32 + lane) * 2;
                            atomicAdd(dst, acc1[cb][s].value());
                            atomicAdd(dst + 1, acc2[cb][s].value());
                        }
                    }
                }
            }
#pragma unroll
            for (int cb = 0; cb < NACC; ++cb)
#pragma unroll
                for (int s = 0; s < C::NI; ++s) {
                    acc1[cb][s].clear();
                    acc2[cb][s].clear();
                }
        };
        int j = 0;
        for (int item = item_begin; item < item_end; ++item, ++j) {
            const int buf = j & 1;
            int n, y0, x0;
            if (C::NI == 1) {
                n = item / C::ITEMS_PER_IMAGE;
                const int r = item - n * C::ITEMS_PER_IMAGE;
                const int yb = r / C::ITEMS_X;
                y0 = yb * 16;
                x0 = (r - yb * C::ITEMS_X) * 8 * C::NT;
            } else {
                n = item * 2;
                y0 = 0;
                x0 = 0;
            }
            if (n != cur_n) {
                flush();
                cur_n = n;
            }
            const bool valid = n + im < p.nimg;
            mbar_wait_bounded(&tfull_bar[buf], ((unsigned)j >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * C::ACC_COLS);
            if constexpr (C::ACCUM) {
                static_assert(!C::ACCUM || (POOL && C::NI == 1), "the accumulating epilogue is written for the pooled front end");
                // all boxes of the previous item have been read by their TMA stores before they are overwritten
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
#pragma unroll 1
                for (int hf = 0; hf < ((EBSD_DBG(p) & 4) ? 0 : 2); ++hf) {
                    float z[32];  // [0,16): sums, [16,32): sums of squares of channels hf*16 + i over the item's tiles
#pragma unroll
                    for (int i = 0; i < 32; ++i) z[i] = 0.f;
#pragma unroll 1
                    for (int t = 0; t < C::NT; ++t) {
                        float v[16], w[16];
                        tmem_ld16(t_row + t * 2 * COUT + hf * 16, v);
                        tmem_ld16(t_row + t * 2 * COUT + COUT + hf * 16, w);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            v[i] = valid ? fmaf(w[i], p.corr_scale, v[i]) : 0.f;
                            z[i] += v[i];
                            z[16 + i] = fmaf(v[i], v[i], z[16 + i]);
                        }
                        // 2x2 max: transposing butterfly, 16 -> 8 -> 4 channels per lane
                        float r[8], o[4];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float send = a_par ? v[i] : v[8 + i];
                            const float keep = a_par ? v[8 + i] : v[i];
                            r[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float send = b_par ? r[i] : r[4 + i];
                            const float keep = b_par ? r[4 + i] : r[i];
                            o[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, PV));
                        }
                        // one box per (tile, 16-channel half): 8 pooled pixels x 64 B, 64B-swizzled.  With 128-byte rows
                        // the eight lanes of a store phase hit only four 16-byte bank groups (two wavefronts per
                        // phase, ncu); with 64-byte rows they cover all eight.
                        const uint32_t stg = stg_u32 + (uint32_t)(t * C::WSTG + hf * 512);
                        const int ch = a_par * 2 + b_par;  // 16-byte chunk of this lane's 4 channels inside the half
                        sts128(stg + (uint32_t)(prow * 64) + (uint32_t)((ch ^ ((prow >> 1) & 3)) << 4),
                               make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3])));
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0 && n < p.nimg && !(EBSD_DBG(p) & 16)) {
                            tma_store_4d(&map_out, stg, hf * 16, (x0 + 8 * t) >> 1, (y0 + 4 * quarter) >> 1, n);
                            bulk_commit();
                        }
                    }
                    if (!(EBSD_DBG(p) & 8)) {
                        // one 32-wide transposing reduction serves both quantities: lane c < 16 ends up with the sum
                        // of channel hf*16 + c, lane 16 + c with its sum of squares
                        const float red = warp_transpose_reduce32(z, lane);
                        if (hf == 0) accum_lo.add(red);
                        else accum_hi.add(red);
                    }
                }
            } else {
#pragma unroll 1
            for (int t = 0; t < ((EBSD_DBG(p) & 4) ? 0 : C::NT); ++t) {
#pragma unroll
                for (int cb = 0; cb < NCB; ++cb) {
                    float v[32], w[32];
                    tmem_ld32(t_row + t * 2 * COUT + cb * 32, v);
                    tmem_ld32(t_row + t * 2 * COUT + COUT + cb * 32, w);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = valid ? fmaf(w[i], p.corr_scale, v[i]) : 0.f;
                    // ---- store through shared memory: one TMA box per warp, block of 32 channels and tile
                    const uint32_t stg = stg_u32 + (uint32_t)(sbuf * C::WSTG);
                    if (lane == 0) bulk_wait_read<C::NSB - 1>();
                    __syncwarp();
                    if (C::STAT_LDS) {
                        // un-pooled values -> scratch row `lane`, 16-byte chunks XOR-swizzled with scr_key(lane): the
                        // eight lanes of a store phase cover all banks, and so do the two pooled pixels (source rows
                        // two apart) of a load phase below
                        const uint32_t rowa = scr_u32 + (uint32_t)(lane * 128);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            sts128(rowa + (uint32_t)((i ^ scr_key(lane)) << 4),
                                   make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]),
                                              __float_as_uint(v[4 * i + 2]), __float_as_uint(v[4 * i + 3])));
                        // 2x2 max-pool through the scratch box: lane (pooled pixel pp, chunk pair cq) fetches its two
                        // 16-byte chunks of the four source pixels (8 LDS.128), 24 maxima, two stores into the TMA
                        // staging box -- the transposing shuffle butterfly (24 SHFL + 48 selects + 24 maxima) cost
                        // three times the instructions
                        __syncwarp();
                        const uint32_t rowp = stg + (uint32_t)(pool_pp * 128);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            float4 o = lds128(pool_row[0] + (uint32_t)(((2 * pool_cq + e) ^ pool_key[0]) << 4));
#pragma unroll
                            for (int k = 1; k < 4; ++k) {
                                const float4 x = lds128(pool_row[k] + (uint32_t)(((2 * pool_cq + e) ^ pool_key[k]) << 4));
                                o.x = fmaxf(o.x, x.x);
                                o.y = fmaxf(o.y, x.y);
                                o.z = fmaxf(o.z, x.z);
                                o.w = fmaxf(o.w, x.w);
                            }
                            sts128(rowp + (uint32_t)(((2 * pool_cq + e) ^ (pool_pp & 7)) << 4),
                                   make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w)));
                        }
                    } else if (POOL) {
                        // transposing butterfly: after the x-pair step a lane keeps 16 channels, after the row-pair
                        // step 8 channels, each the maximum over the 2x2 block
                        float r[16], o[8];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float send = a_par ? v[i] : v[16 + i];
                            const float keep = a_par ? v[16 + i] : v[i];
                            r[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float send = b_par ? r[i] : r[8 + i];
                            const float keep = b_par ? r[8 + i] : r[i];
                            o[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, PV));
                        }
                        const int ch = a_par * 4 + b_par * 2;  // first 16-byte chunk of this lane's 8 channels
                        const uint32_t rowa = stg + (uint32_t)(prow * 128);
                        sts128(rowa + (uint32_t)(((ch) ^ (prow & 7)) << 4),
                               make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3])));
                        sts128(rowa + (uint32_t)(((ch + 1) ^ (prow & 7)) << 4),
                               make_uint4(__float_as_uint(o[4]), __float_as_uint(o[5]), __float_as_uint(o[6]), __float_as_uint(o[7])));
                    } else {
                        const uint32_t rowa = stg + (uint32_t)(urow * 128);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            sts128(rowa + (uint32_t)((i ^ (urow & 7)) << 4),
                                   make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]),
                                              __float_as_uint(v[4 * i + 2]), __float_as_uint(v[4 * i + 3])));
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && n < p.nimg && !(EBSD_DBG(p) & 16)) {
                        int cx, cy;
                        if (C::NI == 1) {
                            cx = POOL ? (x0 + 8 * t) >> 1 : x0 + 8 * t;
                            cy = POOL ? (y0 + 4 * quarter) >> 1 : y0 + 4 * quarter;
                        } else {
                            cx = 0;
                            cy = POOL ? quarter : 2 * quarter;
                        }
                        tma_store_4d(&map_out, stg, cb * 32, cx, cy, n);
                        bulk_commit();
                    }
                    sbuf = (sbuf + 1) % C::NSB;
                    // plane statistics of the un-pooled output
                    if (EBSD_DBG(p) & 8) continue;
                    if constexpr (!POOL || C::STAT_LDS) {
                        // Un-pooled blocks: the warp's 32 pixels x 32 channels are already staged as a swizzled
                        // [pixel row][channel] box for the TMA store, so lane c sums COLUMN c of it: 32 conflict-free
                        // LDS.32 + 64 FADD/FFMA, all independent of each other, instead of two transposing shuffle
                        // reductions (62 SHFL + ~190 ALU instructions in five dependent rounds).  The TMA store only
                        // reads the box; the __syncwarp before the next refill orders these reads before its stores.
                        // Pooled blocks (STAT_LDS): the same over a scratch box the un-pooled values were written to
                        // before the pooling butterfly (row = lane).
                        const uint32_t col = (POOL ? scr_u32 : stg) + (uint32_t)((lane & 3) << 2);
                        float s[C::NI == 1 ? 2 : C::NI], q[C::NI == 1 ? 2 : C::NI];
#pragma unroll
                        for (int i = 0; i < (C::NI == 1 ? 2 : C::NI); ++i) s[i] = q[i] = 0.f;
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            const float x = lds32(col + (uint32_t)(r * 128) + (uint32_t)((((lane >> 2) ^ (POOL ? scr_key(r) : (r & 7)))) << 4));
                            // NI == 2: un-pooled box rows 0-15 = image slot 0, 16-31 = slot 1; scratch rows = lanes, slot (r >> 3) & 1
                            const int a = C::NI == 1 ? (r & 1) : (POOL ? ((r >> 3) & 1) : (r >> 4));
                            s[a] += x;
                            q[a] = fmaf(x, x, q[a]);
                        }
                        if (C::NI == 1) {
                            const float t1 = s[0] + s[1], t2 = q[0] + q[1];
                            if (TILE_FLUSH) {
                                if (n < p.nimg) {
                                    double *dst = p.sums + ((long long)n * COUT + cb * 32 + lane) * 2;
                                    atomicAdd(dst, (double)t1);
                                    atomicAdd(dst + 1, (double)t2);
                                }
                            } else {
                                acc1[TILE_FLUSH ? 0 : cb][0].add(t1);
                                acc2[TILE_FLUSH ? 0 : cb][0].add(t2);
                            }
                        } else {
#pragma unroll
                            for (int sl = 0; sl < C::NI; ++sl) {
                                if (TILE_FLUSH) {
                                    if (n + sl < p.nimg) {
                                        double *dst = p.sums + ((long long)(n + sl) * COUT + cb * 32 + lane) * 2;
                                        atomicAdd(dst, (double)s[sl]);
                                        atomicAdd(dst + 1, (double)q[sl]);
                                    }
                                } else {
                                    acc1[TILE_FLUSH ? 0 : cb][sl].add(s[sl]);
                                    acc2[TILE_FLUSH ? 0 : cb][sl].add(q[sl]);
                                }
                            }
                        }
                        continue;
                    }
                    if (C::NI == 1) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) w[i] = v[i] * v[i];
                        const float t1 = warp_transpose_reduce32(v, lane), t2 = warp_transpose_reduce32(w, lane);
                        if (TILE_FLUSH) {
                            if (n < p.nimg) {
                                double *dst = p.sums + ((long long)n * COUT + cb * 32 + lane) * 2;
                                atomicAdd(dst, (double)t1);
                                atomicAdd(dst + 1, (double)t2);
                            }
                        } else if (EBSD_DBG(p) & 32) {   // role profiling: the reductions without the fp64 accumulation
                            if (t1 + t2 == 12345.678f) acc1[0][0].add(1.0f);
                        } else {
                            acc1[TILE_FLUSH ? 0 : cb][0].add(t1);
                            acc2[TILE_FLUSH ? 0 : cb][0].add(t2);
                        }
                    } else {
#pragma unroll
                        for (int s = 0; s < C::NI; ++s) {
                            float a[32], b[32];
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                a[i] = im == s ? v[i] : 0.f;
                                b[i] = a[i] * a[i];
                            }
                            const float t1 = warp_transpose_reduce32(a, lane), t2 = warp_transpose_reduce32(b, lane);
                            if (TILE_FLUSH) {
                                if (n + s < p.nimg) {
                                    double *dst = p.sums + ((long long)(n + s) * COUT + cb * 32 + lane) * 2;
                                    atomicAdd(dst, (double)t1);
                                    atomicAdd(dst + 1, (double)t2);
                                }
                            } else {
                                acc1[TILE_FLUSH ? 0 : cb][s].add(t1);
                                acc2[TILE_FLUSH ? 0 : cb][s].add(t2);
                            }
                        }
                    }
                }
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);   // PAIR: the follower's relay forwards the completed phase
        }
        flush();
        if (lane == 0) bulk_wait_all();  // the staging buffers must outlive the TMA reads; stores complete before exit
    } else if (warp >= 8 || (C::PRODUCERS > 256 && (warp == 2 || warp == 3))) {
        // ===================== producers: build the fp16 window and the fp8 (value, residual) window in shared memory
        const int ptid = warp >= 8 ? (int)threadIdx.x - 256 : 256 + (int)threadIdx.x - 64;
        const uint32_t smem_base_u32 = smem_u32(smem);
        auto decode = [&](int item, int &n, int &y0, int &x0) {
            if (C::NI == 1) {
                n = item / C::ITEMS_PER_IMAGE;
                const int r = item - n * C::ITEMS_PER_IMAGE;
                const int yb = r / C::ITEMS_X;
                y0 = yb * 16;
                x0 = (r - yb * C::ITEMS_X) * 8 * C::NT;
            } else {
                n = item * 2;
                y0 = 0;
                x0 = 0;
            }
        };
        int tab_n = -1;
        // (scale, shift) of the source planes: biased variance, eps = 1e-5 (torch instance_norm)
        auto update_table = [&](int n) {
            if (n == tab_n) return;
            named_bar_sync(1, C::PRODUCERS);
            for (int i = ptid; i < C::NI * CIN; i += C::PRODUCERS) {
                const int s = i / CIN, c = i - s * CIN;
                float2 t = make_float2(0.f, 0.f);
                if (n + s < p.nimg) {
                    const double *q = p.src_sums + ((long long)(n + s) * CIN + c) * 2;
                    const double mm = q[0] * p.inv_src_plane;
                    double var = q[1] * p.inv_src_plane - mm * mm;
                    if (var < 0.0) var = 0.0;
                    const double rstd = 1.0 / sqrt(var + 1e-5);
                    t = make_float2((float)rstd, (float)(-mm * rstd));
                }
                // channel c -> float4 [(c % 8) / 2][c / 8], scale in component c % 2, shift in 2 + c % 2 (see norm_split8)
                const uint32_t slot = tab_u32 + (uint32_t)(s * CIN * 8 + (((c & 7) >> 1) * (CIN / 8) + (c >> 3)) * 16 + (c & 1) * 4);
                sts32(slot, __float_as_uint(t.x));
                sts32(slot + 8, __float_as_uint(t.y));
            }
            tab_n = n;
            named_bar_sync(1, C::PRODUCERS);
        };
        unsigned ait = 0;
        if constexpr (C::FIRST) {
            // conv0 on CUDA cores, weight-stationary: a producer warp owns one group of 8 output channels and keeps
            // its 72 weights in registers (as 36 channel pairs for FFMA2); a unit = (window position, channel group).
            const int pw = warp - 8, cg = pw & 3, phalf = pw >> 2;
            float2 wreg[9][4];
#pragma unroll
            for (int tp = 0; tp < 9; ++tp)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    wreg[tp][j] = make_float2(__ldg(p.w0 + tp * 32 + cg * 8 + 2 * j), __ldg(p.w0 + tp * 32 + cg * 8 + 2 * j + 1));
            for (int item = item_begin; item < item_end; ++item, ++ait) {
                int n, y0, x0;
                decode(item, n, y0, x0);
                update_table(n);
                const int pb = ait & 1;   // patch buffer of this item (staged by warp 3)
                mbar_wait_bounded(&patch_full[pb], (ait >> 1) & 1u);
                const uint32_t patch_u32 = patch_base_u32 + (uint32_t)(pb * C::PATCH_BYTES);
                const int sa = ait % C::A_STAGES;
                mbar_wait_bounded(&a_empty[sa], ((ait / C::A_STAGES) & 1u) ^ 1u);
                const uint32_t stage_u32 = smem_base_u32 + sa * C::A_STAGE;
                float4 tt[4];  // (scale, shift) of this warp's 8 channels for the current image
#pragma unroll
                for (int j = 0; j < 4; ++j) tt[j] = lds128(tab_u32 + (uint32_t)((j * (CIN / 8) + cg) * 16));
#pragma unroll 2
                for (int pos = phalf * 32 + lane; pos < ((EBSD_DBG(p) & 1) ? 0 : C::WIN_POS); pos += 64) {
                    const int wy = pos / C::PITCH, wx = pos - wy * C::PITCH;
                    const int y = y0 - 1 + wy, x = x0 - 1 + wx;
                    uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
                    if (y >= 0 && y < 128 && x >= 0 && x < 128) {
                        float2 acc[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
                                const float v = lds32(patch_u32 + ((wy + dy) * C::PATCH_S + wx + dx) * 4);
                                const float2 vv = make_float2(v, v);
#pragma unroll
                                for (int j = 0; j < 4; ++j) acc[j] = __ffma2_rn(vv, wreg[dy * 3 + dx][j], acc[j]);
                            }
                        __half2 h[4];
                        uint32_t l[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) norm_split_pair(acc[j], tt[j], h[j], l[j]);
                        hi = *(const uint4 *)h;
                        lo = make_uint4(l[0], l[1], l[2], l[3]);
                    }
                    store_chunk<C>(stage_u32, pos, cg, hi, lo);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&patch_empty[pb]);   // this warp is done reading the patch
                fence_proxy_async();
                if (C::PAIR) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[sa]);
                } else {
                    mbar_arrive(&a_full[sa]);
                }
            }
        } else {
            // raw fp32 -> normalise -> LeakyReLU -> fp16 + fp8 pairs.  One unit = 8 channels of one window position; the
            // global loads of the next batch of units are in flight while the current batch is converted, across
            // stage and item boundaries (a stage = one K chunk of one window).
            constexpr int C8 = C::KC / 8;
            constexpr int UNITS = C::WIN_POS * C8;
            // units per thread and batch.  1296 (block 2) and 1440 (blocks 3-7) units per stage are two batches of 768 with 3 per
            // thread; with 4 the second batch ran 27-40 % full (measured: blocks 2/3 -6 %, block 5 -3 %)
            constexpr int NB3 = (UNITS + C::PRODUCERS * 3 - 1) / (C::PRODUCERS * 3), NB4 = (UNITS + C::PRODUCERS * 4 - 1) / (C::PRODUCERS * 4);
            constexpr int BATCH = 3 * NB3 < 4 * NB4 ? 3 : 4;   // fewer (mostly empty) unit slots per stage
            constexpr int NBATCH = BATCH == 3 ? NB3 : NB4;
            struct Cursor {
                int item, cc, b;
            };
            auto advance = [&](Cursor &c) {
                if (++c.b == NBATCH) {
                    c.b = 0;
                    if (++c.cc == C::NCHUNK) {
                        c.cc = 0;
                        ++c.item;
                    }
                }
            };
            // Thread-constant geometry.  Unit slot k of a thread is u = ptid + PRODUCERS k; when PRODUCERS is a multiple of
            // C8 * PITCH (blocks 3-7: 320 = 8 channel groups x 10 window columns x 4 rows) the channel group and the window
            // column of a thread never change and its window row advances by a constant per slot, so the divisions by
            // C8 and PITCH, done twice per unit (issue and consume), drop out of the loop: the producers of these blocks
            // are bound by their own instruction stream (block 3 alone: 660 us with MMAs and epilogue idle, 435 us with
            // the conversion skipped as well).
            constexpr bool FIXED_GEOM = C::NI == 1 && C::PRODUCERS % (C8 * C::PITCH) == 0;
            constexpr int SLOT_ROWS = C::PRODUCERS / (C8 * C::PITCH);
            const int fg_c8 = ptid % C8, fg_wx = (ptid / C8) % C::PITCH, fg_wy0 = ptid / (C8 * C::PITCH);
            // geometry of unit u of the stage (item, cc): window position, channel group, source pointer (or null)
            auto geom = [&](int n, int y0, int x0, int cc, int k, int &pos, int &c8, int &s) -> const float * {
                const int u = ptid + C::PRODUCERS * k;   // k = unit slot of this thread
                int wy, wx;
                if constexpr (FIXED_GEOM) {
                    c8 = fg_c8;
                    wx = fg_wx;
                    wy = fg_wy0 + SLOT_ROWS * k;
                    pos = wy * C::PITCH + wx;
                } else {
                    pos = u / C8;
                    c8 = u - pos * C8;
                    wy = pos / C::PITCH;
                    wx = pos - wy * C::PITCH;
                }
                int nn, y, x;
                if (C::NI == 1) {
                    s = 0;
                    nn = n;
                    y = y0 - 1 + wy;
                    x = x0 - 1 + wx;
                } else {
                    s = wx / 10;
                    nn = n + s;
                    y = wy - 1;
                    x = wx - s * 10 - 1;
                }
                if (u >= UNITS || y < 0 || y >= W || x < 0 || x >= W || nn >= p.nimg) return nullptr;
                return (const float *)p.src + ((((long long)nn * W + y) * W + x) * CIN + cc * C::KC + c8 * 8);
            };
            auto issue = [&](const Cursor &c, float4 (&buf)[BATCH][2]) {
                int n, y0, x0;
                decode(c.item, n, y0, x0);
#pragma unroll
                for (int i = 0; i < BATCH; ++i) {
                    int pos, c8, s;
                    const float *src = geom(n, y0, x0, c.cc, c.b * BATCH + i, pos, c8, s);
                    if (src) {
                        buf[i][0] = __ldg((const float4 *)src);
                        buf[i][1] = __ldg((const float4 *)src + 1);
                    }
                }
            };
            auto consume = [&](const Cursor &c, float4 (&buf)[BATCH][2], uint32_t stage_u32) {
                int n, y0, x0;
                decode(c.item, n, y0, x0);
                float4 tt[4];   // FIXED_GEOM: one table read per batch instead of one per unit
                if constexpr (FIXED_GEOM) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) tt[j] = lds128(tab_u32 + (uint32_t)((c.cc * C8 + fg_c8) * 16) + (uint32_t)j * (CIN * 2));
                }
#pragma unroll
                for (int i = 0; i < BATCH; ++i) {
                    const int u = ptid + C::PRODUCERS * (c.b * BATCH + i);
                    int pos, c8, s;
                    const float *src = geom(n, y0, x0, c.cc, c.b * BATCH + i, pos, c8, s);
                    if (u >= UNITS) continue;
                    uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
                    if (src) {
                        const float xv[8] = {buf[i][0].x, buf[i][0].y, buf[i][0].z, buf[i][0].w,
                                             buf[i][1].x, buf[i][1].y, buf[i][1].z, buf[i][1].w};
                        if constexpr (FIXED_GEOM) norm_split8_t(xv, tt, hi, lo);
                        else norm_split8(xv, tab_u32 + (uint32_t)(s * CIN * 8 + (c.cc * C8 + c8) * 16), CIN * 2, hi, lo);
                    }
                    store_chunk<C>(stage_u32, pos, c8, hi, lo);
                }
            };
            float4 buf0[BATCH][2], buf1[BATCH][2];
            Cursor cur = {item_begin, 0, 0}, nxt = cur;
            if (cur.item < item_end) issue(nxt, buf0);
            advance(nxt);
            uint32_t stage_u32 = smem_base_u32;
            auto step = [&](float4 (&bc)[BATCH][2], float4 (&bn)[BATCH][2]) {
                if (nxt.item < item_end) issue(nxt, bn);
                if (cur.b == 0) {
                    if (cur.cc == 0) {
                        int n, y0, x0;
                        decode(cur.item, n, y0, x0);
                        update_table(n);
                    }
                    const int sa = ait % C::A_STAGES;
                    mbar_wait_bounded(&a_empty[sa], ((ait / C::A_STAGES) & 1u) ^ 1u);
                    stage_u32 = smem_base_u32 + sa * C::A_STAGE;
                }
                if (!(EBSD_DBG(p) & 1)) consume(cur, bc, stage_u32);
                if (cur.b == NBATCH - 1) {
                    fence_proxy_async();
                    if (C::PAIR) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&a_full[ait % C::A_STAGES]);
                    } else {
                        mbar_arrive(&a_full[ait % C::A_STAGES]);
                    }
                    ++ait;
                }
                advance(cur);
                advance(nxt);
            };
            while (cur.item < item_end) {
                step(buf0, buf1);
                if (cur.item >= item_end) break;
                step(buf1, buf0);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (C::PAIR) {
        cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still signal it or run MMAs on it
        if (warp == 2) tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    } else {
        if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// Plane statistics of conv0's output WITHOUT running conv0 (uint8 patterns).  conv0 has one input channel, so for
// output channel c:   sum_p o_c(p)   = sum_t w[t][c] S[t] / 255,        S[t]    = sum_p k(p + t)
//                     sum_p o_c(p)^2 = sum_{t,t'} w[t][c] w[t'][c] R[t][t'] / 255^2,  R[t][t'] = sum_p k(p + t) k(p + t')
// with k the uint8 pixel (zero outside the image) and t, t' the nine taps.  S and R are exact integers: each lane
// holds four horizontally adjacent pixels per tap as one packed word and a DP4A accumulates four products.
// One CTA per image, 8 warps x 16 rows; writes sums0[n][32][2] (no atomics).
__global__ void __launch_bounds__(256) conv0_stats_u8_kernel(const uint8_t *__restrict__ pats, const float *__restrict__ w0,
                                                             double *__restrict__ sums0) {
    __shared__ unsigned red[8][54];
    __shared__ unsigned tot[54];
    griddep_launch_dependents();
    griddep_wait();   // sums0 was zeroed / read by earlier work of the stream
    const long long n = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned *img = (const unsigned *)(pats + n * 16384);
    auto load_row = [&](int y, unsigned &wm, unsigned &wc, unsigned &wp) {
        unsigned w = 0u;
        if (y >= 0 && y < 128) w = __ldg(img + y * 32 + lane);
        unsigned prev = __shfl_up_sync(0xffffffffu, w, 1), next = __shfl_down_sync(0xffffffffu, w, 1);
        if (lane == 0) prev = 0u;
        if (lane == 31) next = 0u;
        wm = __funnelshift_l(prev, w, 8);   // pixels x-1 .. x+2
        wc = w;                             // pixels x   .. x+3
        wp = __funnelshift_r(w, next, 8);   // pixels x+1 .. x+4
    };
    unsigned acc[54];
#pragma unroll
    for (int i = 0; i < 54; ++i) acc[i] = 0u;
    unsigned t[9];
    const int y0 = warp * 16;
    load_row(y0 - 1, t[0], t[1], t[2]);
    load_row(y0, t[3], t[4], t[5]);
#pragma unroll 1
    for (int y = y0; y < y0 + 16; ++y) {
        load_row(y + 1, t[6], t[7], t[8]);
        int k = 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            acc[i] = __dp4a(t[i], 0x01010101u, acc[i]);
#pragma unroll
            for (int j = i; j < 9; ++j, ++k) acc[k] = __dp4a(t[i], t[j], acc[k]);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) t[i] = t[i + 3];
    }
#pragma unroll
    for (int i = 0; i < 54; ++i) {
        unsigned v = acc[i];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 54) {
        unsigned v = 0u;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        tot[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int c = threadIdx.x;
        double wt[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) wt[i] = (double)w0[i * 32 + c];
        double s1 = 0.0, s2 = 0.0;
        int k = 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            s1 += wt[i] * (double)tot[i];
#pragma unroll
            for (int j = i; j < 9; ++j, ++k) s2 += (i == j ? 1.0 : 2.0) * wt[i] * wt[j] * (double)tot[k];
        }
        sums0[(n * 32 + c) * 2 + 0] = s1 / 255.0;
        sums0[(n * 32 + c) * 2 + 1] = s2 / 65025.0;
    }
}

// mu / logvar heads (latice/model.py:57-58, 127-129) straight from the last block's pooled raw output
// raw9 [n,4,4,128] + its plane sums (64 pixels): InstanceNorm + LeakyReLU here, then the two 2048 -> 16 products.
__global__ void __launch_bounds__(256) heads_norm_kernel(const float *__restrict__ raw9, const double *__restrict__ sums9,
                                                         const float *__restrict__ wh, const float *__restrict__ bh,
                                                         float *__restrict__ mu, float *__restrict__ logvar) {
    __shared__ float feat[2048];
    __shared__ float2 tb[128];
    griddep_launch_dependents();
    griddep_wait();   // raw9 / sums9 come from the last block's kernel
    const long long n = blockIdx.x;
    if (threadIdx.x < 128) {
        const double *q = sums9 + (n * 128 + threadIdx.x) * 2;
        const double mm = q[0] * (1.0 / 64.0);
        double var = q[1] * (1.0 / 64.0) - mm * mm;
        if (var < 0.0) var = 0.0;
        const double rstd = 1.0 / sqrt(var + 1e-5);
        tb[threadIdx.x] = make_float2((float)rstd, (float)(-mm * rstd));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += 256) {
        const float2 t = tb[i & 127];
        const float a = fmaf(raw9[n * 2048 + i], t.x, t.y);
        feat[i] = fmaxf(a, 0.02f * a);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4 *f = (const float4 *)feat;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = lane; i < 512; i += 32) {
        const float4 x = f[i];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const float4 w = ((const float4 *)(wh + (warp * 4 + o) * 2048))[i];
            acc[o] = fmaf(x.x, w.x, acc[o]);
            acc[o] = fmaf(x.y, w.y, acc[o]);
            acc[o] = fmaf(x.z, w.z, acc[o]);
            acc[o] = fmaf(x.w, w.w, acc[o]);
        }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], s);
    }
    if (lane == 0) {
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int oi = warp * 4 + o;
            const float v = acc[o] + bh[oi];
            if (oi < 16) mu[n * 16 + oi] = v;
            else if (logvar) logvar[n * 16 + (oi - 16)] = v;
        }
    }
}

}  // namespace ebsd
