// K1 host side: VAE encoder forward (latice/model.py:55-58) as a chain of kernels over image chunks.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "encoder_finish.cuh"
#include "encoder_mma.cuh"
#include "encoder_mma2.cuh"
#include "encoder_fused.cuh"
#include "encoder_simt.cuh"

namespace ebsd {

struct LayerPlan {
    int cin, cout, hw;  // input spatial size (= output, stride 1 pad 1)
    bool pool;
};
static const LayerPlan kPlan[EBSD_N_CONV] = {
    {1, 32, 128, false},  {32, 32, 128, true},  {32, 64, 64, false}, {64, 64, 64, true},   {64, 128, 32, false},
    {128, 128, 32, true}, {128, 128, 16, false}, {128, 128, 16, true}, {128, 128, 8, false}, {128, 128, 8, true},
};

constexpr int kChunk = 32;                            // images per pass, SIMT path
constexpr int kChunkMma = 256;                        // images per pass, tensor-core path (fills 148 SMs in late layers)
constexpr size_t kRawFloats = 128ull * 128 * 32;      // largest raw / activation plane set per image
constexpr size_t kSumsDoubles = 128 * 2;

}  // namespace ebsd

struct ebsd_encoder {
    int device;
    int use_mma;                 // EBSD_ENCODER_PATH: 3 = fused producer/tcgen05/epilogue blocks (default, "fused"),
                                 // 2 = tcgen05 shifted-window path with finisher kernels ("mma"), 1 = first-generation
                                 // tcgen05 path ("mma1"), 0 = fp32 CUDA-core path ("simt")
    int chunk, sub;              // images per pass / per early sub-chunk of the fused path (EBSD_ENCODER_CHUNK, _SUB)
    float *w_simt[EBSD_N_CONV];  // [tap][ci][co] fp32
    __half *w_mma[EBSD_N_CONV];  // tensor-path packing (encoder_mma.cuh), layers 1..9
    CUtensorMap w_map[EBSD_N_CONV];
    CUtensorMap w_map_fused[EBSD_N_CONV];  // box = the slice one CTA of a cluster fetches (encoder_fused.cuh)
    float *wh;                   // [32][2048] permuted heads
    float *bh;                   // [32]
};

using namespace ebsd;

namespace {

template <int CIN, int COUT>
int launch_simt_conv(const float *in, const float *wt, float *raw, int hw, int nimg, cudaStream_t st) {
    using C = SimtConvCfg<CIN, COUT>;
    dim3 grid((hw / 8) * (hw / 8), COUT / C::CO_TILE, nimg);
    conv3x3_simt_kernel<CIN, COUT><<<grid, C::THREADS, C::smem_floats * sizeof(float), st>>>(in, wt, raw, hw, hw);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

int simt_conv_dispatch(int layer, const float *in, const float *wt, float *raw, int nimg, cudaStream_t st) {
    const LayerPlan &L = kPlan[layer];
    if (L.cin == 32 && L.cout == 32) return launch_simt_conv<32, 32>(in, wt, raw, L.hw, nimg, st);
    if (L.cin == 32 && L.cout == 64) return launch_simt_conv<32, 64>(in, wt, raw, L.hw, nimg, st);
    if (L.cin == 64 && L.cout == 64) return launch_simt_conv<64, 64>(in, wt, raw, L.hw, nimg, st);
    if (L.cin == 64 && L.cout == 128) return launch_simt_conv<64, 128>(in, wt, raw, L.hw, nimg, st);
    if (L.cin == 128 && L.cout == 128) return launch_simt_conv<128, 128>(in, wt, raw, L.hw, nimg, st);
    set_error("encoder: no SIMT kernel for layer %d", layer);
    return EBSD_ERR_ARG;
}

template <int C, int MODE>
int launch_finish_generic(const float *raw, const double *sums, void *out_a, void *out_b, int hw, bool pool, int nimg,
                          cudaStream_t st) {
    const int ho = pool ? hw / 2 : hw;
    const int hq = ho + (MODE == FIN_SPLIT_PAD ? 2 : 0);
    const int items = hq * hq * (C / 4);
    const dim3 grid((items + 256 * kFinishItemsPerThread - 1) / (256 * kFinishItemsPerThread), nimg);
    if (pool) finish_kernel<C, true, MODE><<<grid, 256, 0, st>>>(raw, sums, out_a, out_b, hw, hw);
    else finish_kernel<C, false, MODE><<<grid, 256, 0, st>>>(raw, sums, out_a, out_b, hw, hw);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

template <int C>
int launch_stats(const float *raw, double *sums, int hw, int nimg, cudaStream_t st) {
    EBSD_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)nimg * C * 2 * sizeof(double), st));
    const int pixels = hw * hw;
    int slices = pixels / 512;
    if (slices < 1) slices = 1;
    if (slices > 32) slices = 32;
    plane_stats_kernel<C><<<dim3(nimg, slices), 256, 0, st>>>(raw, sums, pixels);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

template <int C>
int launch_finish(const float *raw, const double *sums, float *out, int hw, bool pool, int nimg, cudaStream_t st) {
    return launch_finish_generic<C, FIN_F32>(raw, sums, out, nullptr, hw, pool, nimg, st);
}

int stats_and_finish(int layer, const float *raw, double *sums, float *out, int nimg, cudaStream_t st) {
    const LayerPlan &L = kPlan[layer];
    int rc;
    switch (L.cout) {
        case 32:
            if ((rc = launch_stats<32>(raw, sums, L.hw, nimg, st))) return rc;
            return launch_finish<32>(raw, sums, out, L.hw, L.pool, nimg, st);
        case 64:
            if ((rc = launch_stats<64>(raw, sums, L.hw, nimg, st))) return rc;
            return launch_finish<64>(raw, sums, out, L.hw, L.pool, nimg, st);
        default:
            if ((rc = launch_stats<128>(raw, sums, L.hw, nimg, st))) return rc;
            return launch_finish<128>(raw, sums, out, L.hw, L.pool, nimg, st);
    }
}


int make_act_map(CUtensorMap *map, const __half *base, int cin, int hw, int nimg, int kc, int tw, int th, int tb) {
    tensormap_encode_fn encode = get_tensormap_encode();
    if (!encode) {
        set_error("encoder: cuTensorMapEncodeTiled entry point not available");
        return EBSD_ERR_CUDA;
    }
    const cuuint64_t gdim[4] = {(cuuint64_t)cin, (cuuint64_t)hw, (cuuint64_t)hw, (cuuint64_t)nimg};
    const cuuint64_t gstride[3] = {(cuuint64_t)cin * 2, (cuuint64_t)hw * cin * 2, (cuuint64_t)hw * hw * cin * 2};
    const cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tb};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult cr = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void *)base, gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE,
                               kc * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("encoder: cuTensorMapEncodeTiled(activations) failed with %d", (int)cr);
        return EBSD_ERR_CUDA;
    }
    return EBSD_OK;
}

int make_weight_map(CUtensorMap *map, const __half *base, int cin, int cout, int kc, int box_rows = 0) {
    tensormap_encode_fn encode = get_tensormap_encode();
    if (!encode) {
        set_error("encoder: cuTensorMapEncodeTiled entry point not available");
        return EBSD_ERR_CUDA;
    }
    const int rows = 9 * (cin / kc) * 2 * cout;
    const cuuint64_t gdim[2] = {(cuuint64_t)kc, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)kc * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)(box_rows > 0 ? box_rows : 2 * cout)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)base, gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE,
                               kc * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("encoder: cuTensorMapEncodeTiled(weights) failed with %d", (int)cr);
        return EBSD_ERR_CUDA;
    }
    return EBSD_OK;
}

template <int CIN, int COUT, int W>
int launch_mma_conv(const ebsd_encoder *enc, int layer, const __half *hi, const __half *lo, float *raw, double *sums,
                    int nimg, cudaStream_t st) {
    using C = MmaConvCfg<CIN, COUT, W>;
    static bool configured = false;
    if (!configured) {
        EBSD_CUDA_TRY(cudaFuncSetAttribute(conv3x3_mma_kernel<CIN, COUT, W>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured = true;
    }
    CUtensorMap map_hi, map_lo;
    int rc;
    if ((rc = make_act_map(&map_hi, hi, CIN, W, nimg, C::KC, C::TW, C::TH, C::TB))) return rc;
    if ((rc = make_act_map(&map_lo, lo, CIN, W, nimg, C::KC, C::TW, C::TH, C::TB))) return rc;
    EBSD_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)nimg * COUT * 2 * sizeof(double), st));
    MmaConvParams p;
    p.raw = raw;
    p.sums = sums;
    p.nimg = nimg;
    p.ntiles = (W * W >= 128) ? nimg * ((W * W) / 128) : (nimg + C::TB - 1) / C::TB;
    const int sms = sm_count();
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    conv3x3_mma_kernel<CIN, COUT, W><<<grid, C::THREADS, C::SMEM_BYTES, st>>>(map_hi, map_lo, enc->w_map[layer], p);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

int mma_conv_dispatch(const ebsd_encoder *enc, int layer, const __half *hi, const __half *lo, float *raw, double *sums,
                      int nimg, cudaStream_t st) {
    switch (layer) {
        case 1: return launch_mma_conv<32, 32, 128>(enc, layer, hi, lo, raw, sums, nimg, st);
        case 2: return launch_mma_conv<32, 64, 64>(enc, layer, hi, lo, raw, sums, nimg, st);
        case 3: return launch_mma_conv<64, 64, 64>(enc, layer, hi, lo, raw, sums, nimg, st);
        case 4: return launch_mma_conv<64, 128, 32>(enc, layer, hi, lo, raw, sums, nimg, st);
        case 5: return launch_mma_conv<128, 128, 32>(enc, layer, hi, lo, raw, sums, nimg, st);
        case 6:
        case 7: return launch_mma_conv<128, 128, 16>(enc, layer, hi, lo, raw, sums, nimg, st);
        case 8:
        case 9: return launch_mma_conv<128, 128, 8>(enc, layer, hi, lo, raw, sums, nimg, st);
    }
    set_error("encoder: no tensor-core kernel for layer %d", layer);
    return EBSD_ERR_ARG;
}

template <int C>
int launch_finish_split(const float *raw, const double *sums, __half *hi, __half *lo, int hw, bool pool, int nimg,
                        cudaStream_t st) {
    return launch_finish_generic<C, FIN_SPLIT>(raw, sums, hi, lo, hw, pool, nimg, st);
}

int finish_split_dispatch(int layer, const float *raw, const double *sums, __half *hi, __half *lo, int nimg,
                          cudaStream_t st) {
    const LayerPlan &L = kPlan[layer];
    switch (L.cout) {
        case 32: return launch_finish_split<32>(raw, sums, hi, lo, L.hw, L.pool, nimg, st);
        case 64: return launch_finish_split<64>(raw, sums, hi, lo, L.hw, L.pool, nimg, st);
        default: return launch_finish_split<128>(raw, sums, hi, lo, L.hw, L.pool, nimg, st);
    }
}

int finish_f32_dispatch(int layer, const float *raw, const double *sums, float *out, int nimg, cudaStream_t st) {
    const LayerPlan &L = kPlan[layer];
    switch (L.cout) {
        case 32: return launch_finish<32>(raw, sums, out, L.hw, L.pool, nimg, st);
        case 64: return launch_finish<64>(raw, sums, out, L.hw, L.pool, nimg, st);
        default: return launch_finish<128>(raw, sums, out, L.hw, L.pool, nimg, st);
    }
}

template <int C>
int stats_only(const float *raw, double *sums, int hw, int nimg, cudaStream_t st) {
    return launch_stats<C>(raw, sums, hw, nimg, st);
}

// f32 NHWC -> fp16 hi / lo planes (debug hook only)
__global__ void split_f32_kernel(const float *__restrict__ x, __half *__restrict__ hi, __half *__restrict__ lo,
                                 long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
}

int g_debug_flags = 0;  // ebsd_debug_set_flags (profiling switches of conv3x3_mma2_kernel)

// ------------------------------------------------------------------ second-generation tensor path (encoder_mma2.cuh)
int make_plane_map(CUtensorMap *map, const __half *base, int cin, long long rows, int kc, int box_rows) {
    tensormap_encode_fn encode = get_tensormap_encode();
    if (!encode) {
        set_error("encoder: cuTensorMapEncodeTiled entry point not available");
        return EBSD_ERR_CUDA;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)cin, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)cin * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)base, gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE,
                               kc * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("encoder: cuTensorMapEncodeTiled(padded plane) failed with %d", (int)cr);
        return EBSD_ERR_CUDA;
    }
    return EBSD_OK;
}

template <int CIN, int COUT, int W>
int launch_mma2_conv(const ebsd_encoder *enc, int layer, const __half *hi, const __half *lo, float *raw, double *sums,
                     int nimg, cudaStream_t st) {
    using C = Mma2Cfg<CIN, COUT, W>;
    static bool configured = false;
    if (!configured) {
        EBSD_CUDA_TRY(cudaFuncSetAttribute(conv3x3_mma2_kernel<CIN, COUT, W>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured = true;
    }
    const long long rows = (long long)nimg * C::HP * C::WP;
    CUtensorMap map_hi, map_lo;
    int rc;
    if ((rc = make_plane_map(&map_hi, hi, CIN, rows, C::KC, C::BOXR))) return rc;
    if ((rc = make_plane_map(&map_lo, lo, CIN, rows, C::KC, C::BOXR))) return rc;
    EBSD_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)nimg * COUT * 2 * sizeof(double), st));
    Mma2Params p;
    p.raw = raw;
    p.sums = sums;
    p.nimg = nimg;
    p.ntiles = (int)((rows + 127) / 128);
    p.dbg = g_debug_flags;
    const int sms = sm_count();
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    conv3x3_mma2_kernel<CIN, COUT, W><<<grid, C::THREADS, C::SMEM_BYTES, st>>>(map_hi, map_lo, enc->w_map[layer], p);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

// layers 1..5 (W >= 32) have a second-generation kernel
int mma2_conv_dispatch(const ebsd_encoder *enc, int layer, const __half *hi, const __half *lo, float *raw,
                       double *sums, int nimg, cudaStream_t st) {
    switch (layer) {
        case 1: return launch_mma2_conv<32, 32, 128>(enc, layer, hi, lo, raw, sums, nimg, st);
        case 2: return launch_mma2_conv<32, 64, 64>(enc, layer, hi, lo, raw, sums, nimg, st);
        case 3: return launch_mma2_conv<64, 64, 64>(enc, layer, hi, lo, raw, sums, nimg, st);
        case 4: return launch_mma2_conv<64, 128, 32>(enc, layer, hi, lo, raw, sums, nimg, st);
        case 5: return launch_mma2_conv<128, 128, 32>(enc, layer, hi, lo, raw, sums, nimg, st);
    }
    set_error("encoder: no second-generation kernel for layer %d", layer);
    return EBSD_ERR_ARG;
}

template <int C>
int launch_finish_split_padded(const float *raw, const double *sums, __half *hi, __half *lo, int hw, bool pool,
                               int nimg, cudaStream_t st) {
    return launch_finish_generic<C, FIN_SPLIT_PAD>(raw, sums, hi, lo, hw, pool, nimg, st);
}

int finish_split_padded_dispatch(int layer, const float *raw, const double *sums, __half *hi, __half *lo, int nimg,
                                 cudaStream_t st) {
    const LayerPlan &L = kPlan[layer];
    switch (L.cout) {
        case 32: return launch_finish_split_padded<32>(raw, sums, hi, lo, L.hw, L.pool, nimg, st);
        case 64: return launch_finish_split_padded<64>(raw, sums, hi, lo, L.hw, L.pool, nimg, st);
        default: return launch_finish_split_padded<128>(raw, sums, hi, lo, L.hw, L.pool, nimg, st);
    }
}

// Workspace of the second-generation path for a chunk of `chunk` images, early layers in sub-chunks of `sub`:
//   raw | early planes (hi, lo) | late planes (hi, lo) | sums
struct Mma2Workspace {
    float *raw;
    __half *early_hi, *early_lo;   // inputs of conv 1..3 for one sub-chunk
    __half *late_hi, *late_lo;     // inputs of conv 4..9 (and the final fp32 features) for the whole chunk
    double *sums;
    size_t bytes;
};
constexpr int kSub = 24;                                     // images per early sub-chunk (keeps ~100 MB in L2)
constexpr size_t kEarlyPlaneHalfs = 130ull * 130 * 32;       // largest early plane per image (conv1 input)
constexpr size_t kLatePlaneHalfs = 34ull * 34 * 128;         // largest late plane per image (conv5 input)
constexpr size_t kLateRawFloats = 32ull * 32 * 128;          // largest raw output of conv 4..9 per image

Mma2Workspace carve_mma2(void *workspace, size_t chunk) {
    const size_t sub = chunk < (size_t)kSub ? chunk : (size_t)kSub;
    uint8_t *p = (uint8_t *)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
    auto take = [&](size_t bytes) {
        uint8_t *r = p;
        p += (bytes + 1023) & ~(size_t)1023;
        return r;
    };
    Mma2Workspace w;
    const size_t raw_floats = sub * kRawFloats > chunk * kLateRawFloats ? sub * kRawFloats : chunk * kLateRawFloats;
    w.raw = (float *)take(raw_floats * sizeof(float));
    w.early_hi = (__half *)take(sub * kEarlyPlaneHalfs * sizeof(__half));
    w.early_lo = (__half *)take(sub * kEarlyPlaneHalfs * sizeof(__half));
    w.late_hi = (__half *)take(chunk * kLatePlaneHalfs * sizeof(__half));
    w.late_lo = (__half *)take(chunk * kLatePlaneHalfs * sizeof(__half));
    w.sums = (double *)take(chunk * kSumsDoubles * sizeof(double));
    w.bytes = (size_t)(p - (uint8_t *)workspace);
    return w;
}

// One chunk (<= kChunkMma images) of the second-generation path.
int forward_chunk_mma2(const ebsd_encoder *enc, const void *pin, int dtype, int nimg, float *mu, float *logvar,
                       const Mma2Workspace &w, cudaStream_t st) {
    int rc;
    const size_t px_bytes = dtype == EBSD_PATTERN_U8 ? 1 : 4;
    // ---- layers 0..3 in L2-sized sub-chunks; conv3's finisher deposits into the chunk-level planes
    for (int s0 = 0; s0 < nimg; s0 += kSub) {
        const int ns = nimg - s0 < kSub ? nimg - s0 : kSub;
        const void *pats = (const uint8_t *)pin + (size_t)s0 * 128 * 128 * px_bytes;
        EBSD_CUDA_TRY(cudaMemsetAsync(w.sums, 0, (size_t)ns * 32 * 2 * sizeof(double), st));
        if (dtype == EBSD_PATTERN_U8) conv0_stats_kernel<true><<<dim3(16, ns), 256, 0, st>>>(pats, enc->w_simt[0], w.sums);
        else conv0_stats_kernel<false><<<dim3(16, ns), 256, 0, st>>>(pats, enc->w_simt[0], w.sums);
        EBSD_LAUNCH_CHECK();
        if (dtype == EBSD_PATTERN_U8)
            conv0_finish_kernel<true><<<dim3(16, ns), 256, 0, st>>>(pats, enc->w_simt[0], w.sums, w.early_hi, w.early_lo);
        else
            conv0_finish_kernel<false><<<dim3(16, ns), 256, 0, st>>>(pats, enc->w_simt[0], w.sums, w.early_hi, w.early_lo);
        EBSD_LAUNCH_CHECK();
        for (int l = 1; l <= 3; ++l) {
            if ((rc = mma2_conv_dispatch(enc, l, w.early_hi, w.early_lo, w.raw, w.sums, ns, st))) return rc;
            if (l < 3) {
                if ((rc = finish_split_padded_dispatch(l, w.raw, w.sums, w.early_hi, w.early_lo, ns, st))) return rc;
            } else {  // conv3 -> pooled 32x32x64, padded 34x34: chunk-level buffer at image offset s0
                const size_t off = (size_t)s0 * 34 * 34 * 64;
                if ((rc = finish_split_padded_dispatch(l, w.raw, w.sums, w.late_hi + off, w.late_lo + off, ns, st)))
                    return rc;
            }
        }
    }
    // ---- layers 4..9 over the whole chunk
    for (int l = 4; l < EBSD_N_CONV; ++l) {
        if (l <= 5) {
            if ((rc = mma2_conv_dispatch(enc, l, w.late_hi, w.late_lo, w.raw, w.sums, nimg, st))) return rc;
        } else {
            if ((rc = mma_conv_dispatch(enc, l, w.late_hi, w.late_lo, w.raw, w.sums, nimg, st))) return rc;
        }
        if (l == 4) {  // next layer (5) reads padded planes
            if ((rc = finish_split_padded_dispatch(l, w.raw, w.sums, w.late_hi, w.late_lo, nimg, st))) return rc;
        } else if (l < EBSD_N_CONV - 1) {  // layers 6..9 read un-padded planes (first-generation kernel)
            if ((rc = finish_split_dispatch(l, w.raw, w.sums, w.late_hi, w.late_lo, nimg, st))) return rc;
        } else {
            if ((rc = finish_f32_dispatch(l, w.raw, w.sums, (float *)w.late_hi, nimg, st))) return rc;
        }
    }
    heads_kernel<<<nimg, 256, 0, st>>>((const float *)w.late_hi, enc->wh, enc->bh, mu, logvar);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}


// ------------------------------------------------------------------ third-generation path (encoder_fused.cuh)
// Two-level chunking.  Blocks 0..3 (128x128 and 64x64 planes, 1.75 MB of raw fp32 per image) run over SUB-chunks
// small enough that a block's output is still in the 126 MB L2 when the next block reads it and is overwritten by
// the next sub-chunk before it is ever written back; blocks 4..9 (<= 0.5 MB per image) run over the whole chunk so
// that even the 8x8 blocks fill all SMs.
constexpr int kChunkFused = 1480;                      // most images per pass (10 per SM): long launches amortise pipeline fill
constexpr int kSubFused = 1480;                        // images per early sub-chunk (blocks 0..3); measured: L2 residency
                                                       // gains less than the extra launches cost, so sub = chunk
constexpr size_t kFusedRaw1Floats = 64ull * 64 * 32;   // per image: pooled output of conv1
constexpr size_t kFusedRaw2Floats = 64ull * 64 * 64;   // per image: output of conv2
constexpr size_t kFusedRaw3Floats = 32ull * 32 * 64;   // per image: pooled output of conv3 (largest tenant of late buffer 0)
constexpr size_t kFusedRaw4Floats = 32ull * 32 * 128;  // per image: output of conv4 (largest tenant of late buffer 1)

// Images per pass for a batch of B: as few passes as the cap allows, all of (nearly) equal size -- a short last
// pass would run the same 12 launches with a fraction of the work.
size_t fused_chunk_for(int64_t B, int cap) {
    if (B <= cap) return (size_t)B;
    const int64_t passes = (B + cap - 1) / cap;
    int64_t c = (B + passes - 1) / passes;
    c += c & 1;  // the 8x8 blocks take images in pairs
    return (size_t)(c < cap ? c : cap);
}

struct FusedWorkspace {
    float *raw1, *raw2;    // sub-chunk level
    float *late0, *late1;  // chunk level: conv3/5/7/9 outputs, conv4/6/8 outputs
    double *sums;          // [EBSD_N_CONV][chunk][128][2]
    size_t bytes;
};

FusedWorkspace carve_fused(void *workspace, size_t chunk, size_t sub) {
    if (sub > chunk) sub = chunk;
    uint8_t *p = (uint8_t *)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
    auto take = [&](size_t bytes) {
        uint8_t *r = p;
        p += (bytes + 1023) & ~(size_t)1023;
        return r;
    };
    FusedWorkspace w;
    w.raw1 = (float *)take(sub * kFusedRaw1Floats * sizeof(float));
    w.raw2 = (float *)take(sub * kFusedRaw2Floats * sizeof(float));
    w.late0 = (float *)take(chunk * kFusedRaw3Floats * sizeof(float));
    w.late1 = (float *)take(chunk * kFusedRaw4Floats * sizeof(float));
    w.sums = (double *)take((size_t)EBSD_N_CONV * chunk * kSumsDoubles * sizeof(double));
    w.bytes = (size_t)(p - (uint8_t *)workspace);
    return w;
}

// rows of the packed weights one CTA fetches per (tap, K chunk): the whole [w_hi; w_lo] tile, or its 1/CL slice
// when the block runs as a CTA pair (FusedCfg::RESIDENT_B / PAIR)
int fused_weight_box_rows(int layer) {
    const int cout = kPlan[layer].cout;
    return (EBSD_PAIR && kPlan[layer].cin >= 64) ? cout / 2 : 2 * cout;  // FusedCfg::PAIR, B_BOX_ROWS
}

// fp32 [nimg,Wo,Wo,COUT] output as a 4-D tensor (c, x, y, n); box = (32 channels, bx, by, bn), 128B-swizzled in smem,
// or (bc = 16 channels, ...) with 64-byte rows and the 64B swizzle (the front-end block's half-channel boxes)
int make_out_map(CUtensorMap *map, const float *base, int cout, int wo, int nimg, int bx, int by, int bn, int bc = 32) {
    tensormap_encode_fn encode = get_tensormap_encode();
    if (!encode) {
        set_error("encoder: cuTensorMapEncodeTiled entry point not available");
        return EBSD_ERR_CUDA;
    }
    const cuuint64_t gdim[4] = {(cuuint64_t)cout, (cuuint64_t)wo, (cuuint64_t)wo, (cuuint64_t)nimg};
    const cuuint64_t gstride[3] = {(cuuint64_t)cout * 4, (cuuint64_t)wo * cout * 4, (cuuint64_t)wo * wo * cout * 4};
    const cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bn};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult cr = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, bc == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("encoder: cuTensorMapEncodeTiled(raw output) failed with %d", (int)cr);
        return EBSD_ERR_CUDA;
    }
    return EBSD_OK;
}

// fp32 [nimg,W,W,CIN] source as (c, x, y, n) for L2 prefetches of whole windows (no shared-memory destination)
int make_src_map(CUtensorMap *map, const float *base, int cin, int w, int nimg, int bx, int by, int bn) {
    tensormap_encode_fn encode = get_tensormap_encode();
    if (!encode) {
        set_error("encoder: cuTensorMapEncodeTiled entry point not available");
        return EBSD_ERR_CUDA;
    }
    const cuuint64_t gdim[4] = {(cuuint64_t)cin, (cuuint64_t)w, (cuuint64_t)w, (cuuint64_t)nimg};
    const cuuint64_t gstride[3] = {(cuuint64_t)cin * 4, (cuuint64_t)w * cin * 4, (cuuint64_t)w * w * cin * 4};
    const cuuint32_t box[4] = {(cuuint32_t)cin, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bn};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult cr = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("encoder: cuTensorMapEncodeTiled(raw source) failed with %d", (int)cr);
        return EBSD_ERR_CUDA;
    }
    return EBSD_OK;
}

template <int CIN, int COUT, int W, int SRC, bool POOL>
int launch_fused(const ebsd_encoder *enc, int layer, const void *src, const double *src_sums, int src_plane,
                 float *raw, double *sums, int nimg, cudaStream_t st) {
    using C = FusedCfg<CIN, COUT, W, SRC, POOL>;
    static bool configured = false;
    if (!configured) {
        EBSD_CUDA_TRY(cudaFuncSetAttribute(conv3x3_fused_kernel<CIN, COUT, W, SRC, POOL>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured = true;
    }
    CUtensorMap map_out, map_src;
    int rc;
    if (SRC == SRC_RAW) {
        // the window a work item reads: (all channels, PITCH or 8 columns, WIN_H or 8 rows, NI images)
        if (C::NI == 1) rc = make_src_map(&map_src, (const float *)src, CIN, W, nimg, C::PITCH, C::WIN_H, 1);
        else rc = make_src_map(&map_src, (const float *)src, CIN, W, nimg, W, W, C::NI);
        if (rc) return rc;
    } else {
        memset(&map_src, 0, sizeof(map_src));
    }
    if (C::NI == 1) rc = make_out_map(&map_out, raw, COUT, POOL ? W / 2 : W, nimg, POOL ? 4 : 8, POOL ? 2 : 4, 1, C::ACCUM ? 16 : 32);
    else rc = make_out_map(&map_out, raw, COUT, POOL ? W / 2 : W, nimg, POOL ? 4 : 8, POOL ? 1 : 2, 2);
    if (rc) return rc;
    FusedParams p;
    p.src = src;
    p.src_sums = src_sums;
    p.inv_src_plane = 1.0 / (double)src_plane;
    p.w0 = enc->w_simt[0];
    p.sums = sums;
    p.nimg = nimg;
    p.nitems = C::NI == 1 ? nimg * C::ITEMS_PER_IMAGE : (nimg + C::NI - 1) / C::NI;
    p.dbg = g_debug_flags;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C::CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent grid: as many CTAs as can be co-resident (clusters must fit inside a GPC), balanced over the items
    static int max_ctas = 0;
    if (max_ctas == 0) {
        max_ctas = sm_count();
        if (C::CL > 1) {
            cfg.gridDim = dim3(max_ctas / C::CL * C::CL);
            int nclusters = 0;
            EBSD_CUDA_TRY(cudaOccupancyMaxActiveClusters(&nclusters, conv3x3_fused_kernel<CIN, COUT, W, SRC, POOL>, &cfg));
            if (nclusters < 1) {
                set_error("encoder: a cluster of %d CTAs of the fused block kernel does not fit on this device", C::CL);
                return EBSD_ERR_CUDA;
            }
            if (nclusters * C::CL < max_ctas) max_ctas = nclusters * C::CL;
        }
    }
    const int per = (p.nitems + max_ctas - 1) / max_ctas;
    int grid = (p.nitems + per - 1) / per;
    grid = (grid + C::CL - 1) / C::CL * C::CL;
    cfg.gridDim = dim3(grid);
    EBSD_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv3x3_fused_kernel<CIN, COUT, W, SRC, POOL>, enc->w_map_fused[layer], map_out, map_src, p));
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

// src: layer 1 -> patterns (dtype), layers 2..9 -> raw output of layer-1 ... ; sums must be zeroed by the caller
int fused_dispatch(const ebsd_encoder *enc, int layer, int dtype, const void *src, const double *src_sums,
                   int src_plane, float *raw, double *sums, int nimg, cudaStream_t st) {
    switch (layer) {
        case 1:
            if (dtype == EBSD_PATTERN_U8)
                return launch_fused<32, 32, 128, SRC_U8, true>(enc, layer, src, src_sums, src_plane, raw, sums, nimg, st);
            return launch_fused<32, 32, 128, SRC_F32, true>(enc, layer, src, src_sums, src_plane, raw, sums, nimg, st);
        case 2: return launch_fused<32, 64, 64, SRC_RAW, false>(enc, layer, src, src_sums, src_plane, raw, sums, nimg, st);
        case 3: return launch_fused<64, 64, 64, SRC_RAW, true>(enc, layer, src, src_sums, src_plane, raw, sums, nimg, st);
        case 4: return launch_fused<64, 128, 32, SRC_RAW, false>(enc, layer, src, src_sums, src_plane, raw, sums, nimg, st);
        case 5: return launch_fused<128, 128, 32, SRC_RAW, true>(enc, layer, src, src_sums, src_plane, raw, sums, nimg, st);
        case 6: return launch_fused<128, 128, 16, SRC_RAW, false>(enc, layer, src, src_sums, src_plane, raw, sums, nimg, st);
        case 7: return launch_fused<128, 128, 16, SRC_RAW, true>(enc, layer, src, src_sums, src_plane, raw, sums, nimg, st);
        case 8: return launch_fused<128, 128, 8, SRC_RAW, false>(enc, layer, src, src_sums, src_plane, raw, sums, nimg, st);
        case 9: return launch_fused<128, 128, 8, SRC_RAW, true>(enc, layer, src, src_sums, src_plane, raw, sums, nimg, st);
    }
    set_error("encoder: no fused kernel for layer %d", layer);
    return EBSD_ERR_ARG;
}

int conv0_stats(const ebsd_encoder *enc, const void *pats, int dtype, int nimg, double *sums0, cudaStream_t st) {
    // uint8: exact integer autocorrelation (patterns must be 4-byte aligned); float32: conv0 recomputed on CUDA cores
    if (dtype == EBSD_PATTERN_U8 && ((uintptr_t)pats & 3) == 0)
        conv0_stats_u8_kernel<<<nimg, 256, 0, st>>>((const uint8_t *)pats, enc->w_simt[0], sums0);
    else if (dtype == EBSD_PATTERN_U8) conv0_stats_kernel<true><<<dim3(16, nimg), 256, 0, st>>>(pats, enc->w_simt[0], sums0);
    else conv0_stats_kernel<false><<<dim3(16, nimg), 256, 0, st>>>(pats, enc->w_simt[0], sums0);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

// One chunk of the fused path: per sub-chunk conv0 statistics + blocks 1..3, then blocks 4..9 and the heads.
int forward_chunk_fused(const ebsd_encoder *enc, const void *pin, int dtype, int nimg, size_t chunk, size_t sub,
                        float *mu, float *logvar, const FusedWorkspace &w, cudaStream_t st) {
    int rc;
    const size_t lstride = chunk * kSumsDoubles;  // doubles per layer in w.sums
    const size_t px_bytes = dtype == EBSD_PATTERN_U8 ? 1 : 4;
    auto layer_sums = [&](int l, int n0) { return w.sums + (size_t)l * lstride + (size_t)n0 * kPlan[l].cout * 2; };
    auto plane = [&](int l) { return kPlan[l].hw * kPlan[l].hw; };
    EBSD_CUDA_TRY(cudaMemsetAsync(w.sums, 0, (size_t)EBSD_N_CONV * lstride * sizeof(double), st));
    for (int s0 = 0; s0 < nimg; s0 += (int)sub) {
        const int ns = nimg - s0 < (int)sub ? nimg - s0 : (int)sub;
        const void *pats = (const uint8_t *)pin + (size_t)s0 * 128 * 128 * px_bytes;
        if ((rc = conv0_stats(enc, pats, dtype, ns, layer_sums(0, s0), st))) return rc;
        if ((rc = fused_dispatch(enc, 1, dtype, pats, layer_sums(0, s0), plane(0), w.raw1, layer_sums(1, s0), ns, st)))
            return rc;
        if ((rc = fused_dispatch(enc, 2, dtype, w.raw1, layer_sums(1, s0), plane(1), w.raw2, layer_sums(2, s0), ns, st)))
            return rc;
        if ((rc = fused_dispatch(enc, 3, dtype, w.raw2, layer_sums(2, s0), plane(2), w.late0 + (size_t)s0 * kFusedRaw3Floats,
                                 layer_sums(3, s0), ns, st)))
            return rc;
    }
    const float *src = w.late0;
    for (int l = 4; l < EBSD_N_CONV; ++l) {
        float *dst = (l & 1) ? w.late0 : w.late1;
        if ((rc = fused_dispatch(enc, l, dtype, src, layer_sums(l - 1, 0), plane(l - 1), dst, layer_sums(l, 0), nimg, st)))
            return rc;
        src = dst;
    }
    heads_norm_kernel<<<nimg, 256, 0, st>>>(src, layer_sums(EBSD_N_CONV - 1, 0), enc->wh, enc->bh, mu, logvar);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

// One chunk of the tensor-core path. Workspace: raw fp32 | hi fp16 | lo fp16 | sums.
int forward_chunk_mma(const ebsd_encoder *enc, const void *pin, int dtype, int nimg, float *mu, float *logvar,
                      float *raw, __half *hi, __half *lo, double *sums, cudaStream_t st) {
    int rc;
    const long long pairs = (long long)nimg * 128 * 64;
    if (dtype == EBSD_PATTERN_U8)
        conv0_kernel<true><<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(pin, enc->w_simt[0], raw, nimg);
    else
        conv0_kernel<false><<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(pin, enc->w_simt[0], raw, nimg);
    EBSD_LAUNCH_CHECK();
    if ((rc = stats_only<32>(raw, sums, 128, nimg, st))) return rc;
    if ((rc = finish_split_dispatch(0, raw, sums, hi, lo, nimg, st))) return rc;
    for (int l = 1; l < EBSD_N_CONV; ++l) {
        if ((rc = mma_conv_dispatch(enc, l, hi, lo, raw, sums, nimg, st))) return rc;
        if (l < EBSD_N_CONV - 1) {
            if ((rc = finish_split_dispatch(l, raw, sums, hi, lo, nimg, st))) return rc;
        } else {
            if ((rc = finish_f32_dispatch(l, raw, sums, (float *)hi, nimg, st))) return rc;
        }
    }
    heads_kernel<<<nimg, 256, 0, st>>>((const float *)hi, enc->wh, enc->bh, mu, logvar);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

}  // namespace

extern "C" {

int ebsd_encoder_create(ebsd_encoder **out, const ebsd_weights *w, int device, void *stream) {
    EBSD_REQUIRE(out && w, "ebsd_encoder_create: null pointer");
    *out = nullptr;
    EBSD_CUDA_TRY(cudaSetDevice(device));
    int rc = check_device_arch();
    if (rc) return rc;
    for (int i = 0; i < EBSD_N_CONV; ++i) EBSD_REQUIRE(w->conv_w[i], "ebsd_encoder_create: conv_w[%d] is null", i);
    EBSD_REQUIRE(w->mu_w && w->mu_b && w->logvar_w && w->logvar_b, "ebsd_encoder_create: head weights are null");
    cudaStream_t st = (cudaStream_t)stream;
    ebsd_encoder *enc = (ebsd_encoder *)calloc(1, sizeof(ebsd_encoder));
    EBSD_REQUIRE(enc, "ebsd_encoder_create: out of host memory");
    enc->device = device;
    for (int i = 0; i < EBSD_N_CONV; ++i) {
        const int total = 9 * kPlan[i].cin * kPlan[i].cout;
        EBSD_CUDA_TRY(cudaMalloc(&enc->w_simt[i], (size_t)total * sizeof(float)));
        pack_conv_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(w->conv_w[i], enc->w_simt[i], kPlan[i].cin,
                                                                    kPlan[i].cout);
        EBSD_LAUNCH_CHECK();
    }
    const char *path = getenv("EBSD_ENCODER_PATH");
    enc->use_mma = 3;
    if (path && strcmp(path, "simt") == 0) enc->use_mma = 0;
    if (path && strcmp(path, "mma1") == 0) enc->use_mma = 1;
    if (path && strcmp(path, "mma") == 0) enc->use_mma = 2;
    enc->chunk = kChunkFused;
    enc->sub = kSubFused;
    if (const char *c = getenv("EBSD_ENCODER_CHUNK")) {
        const int v = atoi(c);
        if (v >= 1 && v <= 4096) enc->chunk = v;
    }
    if (const char *c = getenv("EBSD_ENCODER_SUB")) {
        const int v = atoi(c);
        if (v >= 1 && v <= 4096) enc->sub = v;
    }
    for (int i = 1; i < EBSD_N_CONV; ++i) {
        const int cin = kPlan[i].cin, cout = kPlan[i].cout, kc = cin < 64 ? cin : 64;
        const int total = 9 * (cin / kc) * 2 * cout * kc;
        EBSD_CUDA_TRY(cudaMalloc(&enc->w_mma[i], (size_t)total * sizeof(__half)));
        pack_conv_weights_mma_kernel<<<(total + 255) / 256, 256, 0, st>>>(w->conv_w[i], enc->w_mma[i], cin, cout, kc);
        EBSD_LAUNCH_CHECK();
        if ((rc = make_weight_map(&enc->w_map[i], enc->w_mma[i], cin, cout, kc))) return rc;
        if ((rc = make_weight_map(&enc->w_map_fused[i], enc->w_mma[i], cin, cout, kc, fused_weight_box_rows(i)))) return rc;
    }
    EBSD_CUDA_TRY(cudaMalloc(&enc->wh, 32 * 2048 * sizeof(float)));
    EBSD_CUDA_TRY(cudaMalloc(&enc->bh, 32 * sizeof(float)));
    pack_head_weights_kernel<<<(32 * 2048 + 255) / 256, 256, 0, st>>>(w->mu_w, w->logvar_w, w->mu_b, w->logvar_b,
                                                                     enc->wh, enc->bh);
    EBSD_LAUNCH_CHECK();
    EBSD_CUDA_TRY(cudaStreamSynchronize(st));  // the caller may free `w` right after create returns
    *out = enc;
    return EBSD_OK;
}

void ebsd_encoder_destroy(ebsd_encoder *enc) {
    if (!enc) return;
    for (int i = 0; i < EBSD_N_CONV; ++i) cudaFree(enc->w_simt[i]);
    for (int i = 1; i < EBSD_N_CONV; ++i) cudaFree(enc->w_mma[i]);
    cudaFree(enc->wh);
    cudaFree(enc->bh);
    free(enc);
}

size_t ebsd_encoder_workspace_bytes(const ebsd_encoder *enc, int64_t B) {
    if (B <= 0) return 0;
    const int chunk_cap = (enc && enc->use_mma) ? kChunkMma : kChunk;
    const size_t nimg = (size_t)(B < chunk_cap ? B : chunk_cap);
    if (enc && enc->use_mma == 3) {
        const size_t c = fused_chunk_for(B, enc->chunk);
        return carve_fused(nullptr, c, (size_t)enc->sub).bytes + 1024;
    }
    if (enc && enc->use_mma == 2) return carve_mma2(nullptr, nimg).bytes + 1024;
    // raw fp32 + (fp32 activations | fp16 hi + fp16 lo planes) + plane sums
    return nimg * (2 * kRawFloats * sizeof(float) + kSumsDoubles * sizeof(double)) + 256;
}

int ebsd_encoder_forward(ebsd_encoder *enc, const void *patterns, int dtype, int64_t B, float *mu, float *logvar,
                         void *workspace, size_t workspace_bytes, void *stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    EBSD_REQUIRE(enc != nullptr, "ebsd_encoder_forward: null encoder");
    EBSD_REQUIRE(dtype == EBSD_PATTERN_U8 || dtype == EBSD_PATTERN_F32, "ebsd_encoder_forward: bad dtype %d", dtype);
    EBSD_REQUIRE(B >= 0, "ebsd_encoder_forward: negative batch");
    if (B == 0) return EBSD_OK;
    EBSD_REQUIRE(patterns && mu, "ebsd_encoder_forward: null pointer");
    const size_t need = ebsd_encoder_workspace_bytes(enc, B);
    if (!workspace || workspace_bytes < need) {
        set_error("ebsd_encoder_forward: workspace too small (%zu < %zu)", workspace_bytes, need);
        return EBSD_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int chunk_cap = enc->use_mma ? kChunkMma : kChunk;
    const size_t chunk = (size_t)(B < chunk_cap ? B : chunk_cap);
    uint8_t *ws = (uint8_t *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float *raw = (float *)ws;
    float *act = raw + chunk * kRawFloats;
    double *sums = (double *)(act + chunk * kRawFloats);

    const size_t px_bytes = dtype == EBSD_PATTERN_U8 ? 1 : 4;
    if (enc->use_mma == 3) {
        const size_t fchunk = fused_chunk_for(B, enc->chunk);
        const FusedWorkspace wf = carve_fused(workspace, fchunk, (size_t)enc->sub);
        for (int64_t b0 = 0; b0 < B; b0 += (int64_t)fchunk) {
            const int nimg = (int)((B - b0) < (int64_t)fchunk ? (B - b0) : (int64_t)fchunk);
            const void *pin = (const uint8_t *)patterns + (size_t)b0 * 128 * 128 * px_bytes;
            if ((rc = forward_chunk_fused(enc, pin, dtype, nimg, fchunk, (size_t)enc->sub, mu + b0 * 16,
                                          logvar ? logvar + b0 * 16 : nullptr, wf, st)))
                return rc;
        }
        return EBSD_OK;
    }
    if (enc->use_mma == 2) {
        const Mma2Workspace w2 = carve_mma2(workspace, chunk);
        for (int64_t b0 = 0; b0 < B; b0 += kChunkMma) {
            const int nimg = (int)((B - b0) < kChunkMma ? (B - b0) : kChunkMma);
            const void *pin = (const uint8_t *)patterns + (size_t)b0 * 128 * 128 * px_bytes;
            if ((rc = forward_chunk_mma2(enc, pin, dtype, nimg, mu + b0 * 16, logvar ? logvar + b0 * 16 : nullptr, w2,
                                         st)))
                return rc;
        }
        return EBSD_OK;
    }
    if (enc->use_mma) {
        __half *hi = (__half *)act;
        __half *lo = hi + chunk * kRawFloats;
        for (int64_t b0 = 0; b0 < B; b0 += kChunkMma) {
            const int nimg = (int)((B - b0) < kChunkMma ? (B - b0) : kChunkMma);
            const void *pin = (const uint8_t *)patterns + (size_t)b0 * 128 * 128 * px_bytes;
            if ((rc = forward_chunk_mma(enc, pin, dtype, nimg, mu + b0 * 16, logvar ? logvar + b0 * 16 : nullptr, raw,
                                        hi, lo, sums, st)))
                return rc;
        }
        return EBSD_OK;
    }
    for (int64_t b0 = 0; b0 < B; b0 += kChunk) {
        const int nimg = (int)((B - b0) < kChunk ? (B - b0) : kChunk);
        const void *pin = (const uint8_t *)patterns + (size_t)b0 * 128 * 128 * px_bytes;
        const long long pairs = (long long)nimg * 128 * 64;
        if (dtype == EBSD_PATTERN_U8)
            conv0_kernel<true><<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(pin, enc->w_simt[0], raw, nimg);
        else
            conv0_kernel<false><<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(pin, enc->w_simt[0], raw, nimg);
        EBSD_LAUNCH_CHECK();
        if ((rc = stats_and_finish(0, raw, sums, act, nimg, st))) return rc;
        for (int l = 1; l < EBSD_N_CONV; ++l) {
            if ((rc = simt_conv_dispatch(l, act, enc->w_simt[l], raw, nimg, st))) return rc;
            if ((rc = stats_and_finish(l, raw, sums, act, nimg, st))) return rc;
        }
        heads_kernel<<<nimg, 256, 0, st>>>(act, enc->wh, enc->bh, mu + b0 * 16, logvar ? logvar + b0 * 16 : nullptr);
        EBSD_LAUNCH_CHECK();
    }
    return EBSD_OK;
}


int ebsd_debug_conv_layer(ebsd_encoder *enc, int layer, int use_mma, const float *act, int nimg, float *raw,
                          double *sums, void *workspace, size_t workspace_bytes, void *stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    EBSD_REQUIRE(enc && act && raw && sums, "ebsd_debug_conv_layer: null pointer");
    EBSD_REQUIRE(layer >= 1 && layer < EBSD_N_CONV, "ebsd_debug_conv_layer: layer must be in [1,9]");
    EBSD_REQUIRE(nimg >= 1, "ebsd_debug_conv_layer: nimg must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    const LayerPlan &L = kPlan[layer];
    if (!use_mma) {
        if ((rc = simt_conv_dispatch(layer, act, enc->w_simt[layer], raw, nimg, st))) return rc;
        switch (L.cout) {
            case 32: return launch_stats<32>(raw, sums, L.hw, nimg, st);
            case 64: return launch_stats<64>(raw, sums, L.hw, nimg, st);
            default: return launch_stats<128>(raw, sums, L.hw, nimg, st);
        }
    }
    if (use_mma == 2) {
        EBSD_REQUIRE(layer <= 5, "ebsd_debug_conv_layer: the shifted-window kernel covers layers 1..5");
        const long long np = (long long)nimg * (L.hw + 2) * (L.hw + 2) * L.cin;
        const size_t need2 = (size_t)np * 2 * sizeof(__half) + 2048;
        if (!workspace || workspace_bytes < need2) {
            set_error("ebsd_debug_conv_layer: workspace too small (%zu < %zu)", workspace_bytes, need2);
            return EBSD_ERR_WORKSPACE;
        }
        __half *phi = (__half *)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
        __half *plo = phi + ((np + 511) / 512) * 512;
        split_pad_f32_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(act, phi, plo, L.hw, L.hw, L.cin, nimg);
        EBSD_LAUNCH_CHECK();
        return mma2_conv_dispatch(enc, layer, phi, plo, raw, sums, nimg, st);
    }
    const long long n = (long long)nimg * L.hw * L.hw * L.cin;
    const size_t need = (size_t)n * 2 * sizeof(__half) + 256;
    if (!workspace || workspace_bytes < need) {
        set_error("ebsd_debug_conv_layer: workspace too small (%zu < %zu)", workspace_bytes, need);
        return EBSD_ERR_WORKSPACE;
    }
    __half *hi = (__half *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    __half *lo = hi + n;
    split_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(act, hi, lo, n);
    EBSD_LAUNCH_CHECK();
    return mma_conv_dispatch(enc, layer, hi, lo, raw, sums, nimg, st);
}

void ebsd_debug_set_flags(int flags) { g_debug_flags = flags; }

int ebsd_debug_fused_layer(ebsd_encoder *enc, int layer, int dtype, const void *src, double *src_sums, int src_plane,
                           int nimg, float *raw, double *sums, void *stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    EBSD_REQUIRE(enc && src && src_sums && raw && sums, "ebsd_debug_fused_layer: null pointer");
    EBSD_REQUIRE(layer >= 1 && layer < EBSD_N_CONV, "ebsd_debug_fused_layer: layer must be in [1,9]");
    EBSD_REQUIRE(nimg >= 1, "ebsd_debug_fused_layer: nimg must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    EBSD_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)nimg * kPlan[layer].cout * 2 * sizeof(double), st));
    if (layer == 1) {
        EBSD_CUDA_TRY(cudaMemsetAsync(src_sums, 0, (size_t)nimg * 32 * 2 * sizeof(double), st));
        if ((rc = conv0_stats(enc, src, dtype, nimg, src_sums, st))) return rc;
        src_plane = 128 * 128;
    }
    return fused_dispatch(enc, layer, dtype, src, src_sums, src_plane, raw, sums, nimg, st);
}

}  // extern "C"
