// K1 host side: VAE encoder forward (latice/model.py:55-58) as a chain of kernels over image chunks.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
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

constexpr int kChunk = 32;                            // images per pass (keeps inter-layer traffic near L2)
constexpr size_t kRawFloats = 128ull * 128 * 32;      // largest raw / activation plane set per image
constexpr size_t kSumsDoubles = 128 * 2;

}  // namespace ebsd

struct ebsd_encoder {
    int device;
    float *w_simt[EBSD_N_CONV];  // [tap][ci][co] fp32
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
    const int ho = pool ? hw / 2 : hw;
    const long long total = (long long)nimg * ho * ho * (C / 4);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (pool) finish_f32_kernel<C, true><<<blocks, 256, 0, st>>>(raw, sums, out, hw, hw, nimg);
    else finish_f32_kernel<C, false><<<blocks, 256, 0, st>>>(raw, sums, out, hw, hw, nimg);
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
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
    cudaFree(enc->wh);
    cudaFree(enc->bh);
    free(enc);
}

size_t ebsd_encoder_workspace_bytes(const ebsd_encoder *enc, int64_t B) {
    (void)enc;
    if (B <= 0) return 0;
    const size_t nimg = (size_t)(B < kChunk ? B : kChunk);
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
    const size_t chunk = (size_t)(B < kChunk ? B : kChunk);
    uint8_t *ws = (uint8_t *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float *raw = (float *)ws;
    float *act = raw + chunk * kRawFloats;
    double *sums = (double *)(act + chunk * kRawFloats);

    const size_t px_bytes = dtype == EBSD_PATTERN_U8 ? 1 : 4;
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

}  // extern "C"
