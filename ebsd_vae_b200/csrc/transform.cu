// Input transform on the device: 8-bit quantise + centre crop / zero pad to 128 x 128.
//
// Replaces create_default_transform (latice/data_module.py:17-33 = ToPILImage -> Grayscale -> CenterCrop ->
// ToTensor) as applied by DPdataset.__getitem__ (data_module.py:122-133) and encode_pattern(s_batch)
// (latice/index/dp_indexer.py:124-126, 150-163), up to the uint8 image; ToTensor's k/255 happens inside the encoder.
//
//   quantise: (x * 255).astype(uint8) as numpy does it on x86-64 (torchvision to_pil_image): the product is
//             rounded in the input's own precision, truncated toward zero to a 32-bit integer and the low byte kept;
//             NaN and |x*255| >= 2^31 give 0 (cvttsd2si overflow pattern 0x80000000).  uint8 input is taken as is.
//             The dictionary path (DPdataset.__getitem__, data_module.py:132) casts every frame to float64 FIRST, whatever
//             the file holds -- uint8 / int16 / uint16 / int32 / float32 files included, so a stored 255 becomes
//             65025 mod 256 = 1: src_dtype | EBSD_SRC_VIA_F64 does that cast here, per pixel, instead of a float64
//             copy of the chunk on the host or the device.
//   crop/pad: torchvision center_crop -- the host passes, per axis, the first source index, the first destination
//             index and the length of the copied span (ebsd_vae_b200/transform.py:_axis_window); the rest is zero.
#include "common.cuh"

namespace ebsd {

__device__ __forceinline__ uint8_t quantise_f64(double x) {
    const double v = x * 255.0;
    if (!(fabs(v) < 2147483648.0)) return 0;
    return (uint8_t)(int)v;
}
__device__ __forceinline__ uint8_t quantise_f32(float x) {
    const float v = __fmul_rn(x, 255.0f);
    if (!(fabsf(v) < 2147483648.0f)) return 0;
    return (uint8_t)(int)v;
}

struct CropParams {
    long long B;
    int H, W;          // source frame
    int sy, dy, ly;    // rows: source start, destination start, length
    int sx, dx, lx;    // columns
};

template <int DTYPE>
__device__ __forceinline__ double load_as_f64(const void *src, long long off) {
    if (DTYPE == 0) return (double)((const uint8_t *)src)[off];
    if (DTYPE == 1) return (double)((const float *)src)[off];
    if (DTYPE == 2) return ((const double *)src)[off];
    if (DTYPE == 3) return (double)((const int16_t *)src)[off];
    if (DTYPE == 4) return (double)((const uint16_t *)src)[off];
    if (DTYPE == 5) return (double)((const int32_t *)src)[off];
    return (double)((const long long *)src)[off];
}

// one thread = four horizontally adjacent output pixels (one 32-bit store)
template <int DTYPE, bool VIA_F64>  // EBSD_SRC_* (include/ebsd_b200.h)
__global__ void __launch_bounds__(256) quantise_crop_kernel(const void *__restrict__ src, uint8_t *__restrict__ dst,
                                                            const CropParams p) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = p.B * 128 * 32;
    if (gid >= total) return;
    const int q = (int)(gid & 31);
    const int oy = (int)((gid >> 5) & 127);
    const long long n = gid >> 12;
    uint32_t packed = 0u;
    const int yy = oy - p.dy;
    if (yy >= 0 && yy < p.ly) {
        const long long row = (n * p.H + (p.sy + yy)) * p.W;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int xx = q * 4 + j - p.dx;
            if (xx < 0 || xx >= p.lx) continue;
            const long long off = row + p.sx + xx;
            uint8_t v;
            if (VIA_F64) v = quantise_f64(load_as_f64<DTYPE>(src, off));
            else if (DTYPE == 0) v = ((const uint8_t *)src)[off];
            else if (DTYPE == 1) v = quantise_f32(((const float *)src)[off]);
            else v = quantise_f64(((const double *)src)[off]);
            packed |= (uint32_t)v << (8 * j);
        }
    }
    ((uint32_t *)dst)[gid] = packed;
}

}  // namespace ebsd

using namespace ebsd;

extern "C" int ebsd_quantize_crop(const void *src, int src_dtype, int64_t B, int H, int W, int sy, int dy, int ly,
                                  int sx, int dx, int lx, uint8_t *dst, void *stream) {
    int rc = check_device_arch();
    if (rc) return rc;
    const bool via = (src_dtype & EBSD_SRC_VIA_F64) != 0;
    const int code = src_dtype & ~EBSD_SRC_VIA_F64;
    EBSD_REQUIRE(code >= 0 && code <= (via ? EBSD_SRC_I64 : EBSD_SRC_F64), "ebsd_quantize_crop: bad dtype %d", src_dtype);
    EBSD_REQUIRE(B >= 0 && H > 0 && W > 0, "ebsd_quantize_crop: bad shape");
    EBSD_REQUIRE(ly >= 0 && lx >= 0 && sy >= 0 && sx >= 0 && dy >= 0 && dx >= 0 && sy + ly <= H && sx + lx <= W &&
                     dy + ly <= 128 && dx + lx <= 128,
                 "ebsd_quantize_crop: window does not fit (source %dx%d, rows %d+%d->%d, columns %d+%d->%d)", H, W, sy,
                 ly, dy, sx, lx, dx);
    if (B == 0) return EBSD_OK;
    EBSD_REQUIRE(src && dst, "ebsd_quantize_crop: null pointer");
    EBSD_REQUIRE(((uintptr_t)dst & 3) == 0, "ebsd_quantize_crop: dst must be 4-byte aligned");
    CropParams p;
    p.B = B;
    p.H = H;
    p.W = W;
    p.sy = sy;
    p.dy = dy;
    p.ly = ly;
    p.sx = sx;
    p.dx = dx;
    p.lx = lx;
    const long long total = B * 128 * 32;
    const unsigned grid = (unsigned)((total + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (!via) {
        if (code == 0) quantise_crop_kernel<0, false><<<grid, 256, 0, st>>>(src, dst, p);
        else if (code == 1) quantise_crop_kernel<1, false><<<grid, 256, 0, st>>>(src, dst, p);
        else quantise_crop_kernel<2, false><<<grid, 256, 0, st>>>(src, dst, p);
    } else {
        switch (code) {
            case 0: quantise_crop_kernel<0, true><<<grid, 256, 0, st>>>(src, dst, p); break;
            case 1: quantise_crop_kernel<1, true><<<grid, 256, 0, st>>>(src, dst, p); break;
            case 2: quantise_crop_kernel<2, true><<<grid, 256, 0, st>>>(src, dst, p); break;
            case 3: quantise_crop_kernel<3, true><<<grid, 256, 0, st>>>(src, dst, p); break;
            case 4: quantise_crop_kernel<4, true><<<grid, 256, 0, st>>>(src, dst, p); break;
            case 5: quantise_crop_kernel<5, true><<<grid, 256, 0, st>>>(src, dst, p); break;
            default: quantise_crop_kernel<6, true><<<grid, 256, 0, st>>>(src, dst, p); break;
        }
    }
    EBSD_LAUNCH_CHECK();
    return EBSD_OK;
}

// ---------------------------------------------------------------------------------------- angle file (host code)
// The text of an angle file after its two header lines -> float64 [rows][3] (latice/data_module.py:100-110: fields
// separated by single spaces, empty fields dropped, float(token)).  Runs without the Python GIL (ctypes releases it), so
// build_dictionary parses the angles while the GPU encodes; the Python parser needed 0.35 s per 100k rows and held the
// GIL the launching thread needs.  Only the REGULAR case is handled here: every line exactly three plain decimal numbers,
// nothing but spaces around them.  Anything else (short or long rows, blank lines, tabs, "nan", "1_0", non-ASCII, lone
// carriage returns) returns -1 and the caller falls back to the Python restatement, which reproduces the reference's
// padding and error messages.  std::from_chars is locale independent and correctly rounded, like Python's float().
#include <charconv>

namespace {
inline bool plain_decimal(const char *b, const char *e) {
    const char *p = b;
    if (p < e && (*p == '+' || *p == '-')) ++p;
    int digits = 0;
    while (p < e && *p >= '0' && *p <= '9') ++p, ++digits;
    if (p < e && *p == '.') {
        ++p;
        while (p < e && *p >= '0' && *p <= '9') ++p, ++digits;
    }
    if (digits == 0) return false;
    if (p < e && (*p == 'e' || *p == 'E')) {
        ++p;
        if (p < e && (*p == '+' || *p == '-')) ++p;
        int ed = 0;
        while (p < e && *p >= '0' && *p <= '9') ++p, ++ed;
        if (ed == 0) return false;
    }
    return p == e;
}
}  // namespace

extern "C" int64_t ebsd_parse_angle_text(const char *text, size_t len, double *out, int64_t capacity_rows) {
    if (!text || !out || capacity_rows < 0) return -1;
    const char *p = text, *end = text + len;
    int64_t rows = 0;
    while (p < end) {
        const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *line_end = eol ? eol : end;
        const char *le = line_end;
        if (le > p && le[-1] == '\r' && eol) --le;   // "\r\n" (universal newlines)
        if (rows >= capacity_rows) return -1;
        int ntok = 0;
        const char *q = p;
        while (true) {
            while (q < le && *q == ' ') ++q;
            if (q >= le) break;
            const char *tb = q;
            while (q < le && *q != ' ') {
                const unsigned char c = (unsigned char)*q;
                if (c < 0x21 || c > 0x7e) return -1;   // tabs, control characters, lone CR, non-ASCII: generic parser
                ++q;
            }
            if (ntok == 3 || !plain_decimal(tb, q)) return -1;
            const char *fb = (*tb == '+') ? tb + 1 : tb;   // from_chars does not take a leading plus
            double v = 0.0;
            const std::from_chars_result r = std::from_chars(fb, q, v);
            if (r.ec == std::errc::result_out_of_range) {
                return -1;   // float() gives inf / 0.0 here: leave it to the generic parser
            }
            if (r.ec != std::errc() || r.ptr != q) return -1;
            out[rows * 3 + ntok] = v;
            ++ntok;
        }
        if (ntok != 3) return -1;
        ++rows;
        p = eol ? eol + 1 : end;
    }
    return rows;
}
