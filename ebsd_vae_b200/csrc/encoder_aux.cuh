// K1 auxiliary kernels: weight packing (run once by ebsd_encoder_create), the conv0 plane statistics for inputs the
// integer-autocorrelation kernel does not take (float32 tensors, unaligned uint8), and a warp reduction the fused block
// kernels and these kernels share.
#pragma once
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"

namespace ebsd {

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

// Two floats -> two e4m3 bytes (round to nearest even, saturating at +-448); x in the low byte.
__device__ __forceinline__ uint32_t pack_e4m3x2(float x, float y) {
    return (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(x, y), __NV_SATFINITE, __NV_E4M3);
}

// Scale of the fp8 correction operands (see encoder_fused.cuh, "Arithmetic"): activation residuals are stored as
// e4m3((a - fp16(a)) * kResidualScale); |a| < 128 after InstanceNorm of a plane of <= 16384 pixels, so the scaled
// residual stays below 256 < 448.
constexpr float kResidualScale = 4096.0f;

// conv0 weights: torch [32,1,3,3] -> [tap][32] fp32 (general form: [tap][ci][co]).
__global__ void pack_conv_weights_kernel(const float *__restrict__ w, float *__restrict__ out, int cin, int cout) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = 9 * cin * cout;
    if (i >= total) return;
    const int co = i % cout;
    const int ci = (i / cout) % cin;
    const int tap = i / (cout * cin);
    out[i] = w[((long long)co * cin + ci) * 9 + tap];
}

// Weight packing for the fused tensor-core blocks: torch [Cout,Cin,3,3] fp32 -> rows of KC 16-bit slots (K-major),
//   row (kb*2*Cout + r), kb = tap*NCHUNK + chunk:
//     r <  Cout: fp16(w[co = r]) of the chunk's KC input channels                      (operand of the fp16 MMA)
//     r >= Cout: for each channel PAIR (k, k+1) of co = r - Cout the four e4m3 bytes
//                [ (w - fp16 w)(k), (w - fp16 w)(k+1) ] * 4096 * wscale,  [ w(k), w(k+1) ] * wscale
//                -- the K order of the activation side's [ a(k), a(k+1), res(k), res(k+1) ]  (operand of the fp8 MMA)
__global__ void pack_fused_weights_kernel(const float *__restrict__ w, uint16_t *__restrict__ out, int cin, int cout,
                                          int kc, float wscale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nchunk = cin / kc;
    const int total = 9 * nchunk * 2 * cout * kc;
    if (i >= total) return;
    const int k = i % kc;
    const int r = (i / kc) % (2 * cout);
    const int kb = i / (kc * 2 * cout);
    const int tap = kb / nchunk, cc = kb % nchunk;
    if (r < cout) {
        const float val = w[((long long)r * cin + cc * kc + k) * 9 + tap];
        out[i] = __half_as_ushort(__float2half_rn(val));
        return;
    }
    const int co = r - cout, k0 = k & ~1;
    const float v0 = w[((long long)co * cin + cc * kc + k0) * 9 + tap];
    const float v1 = w[((long long)co * cin + cc * kc + k0 + 1) * 9 + tap];
    if (k & 1) {
        out[i] = (uint16_t)pack_e4m3x2(v0 * wscale, v1 * wscale);
    } else {
        const float l0 = v0 - __half2float(__float2half_rn(v0)), l1 = v1 - __half2float(__float2half_rn(v1));
        out[i] = (uint16_t)pack_e4m3x2(l0 * (kResidualScale * wscale), l1 * (kResidualScale * wscale));
    }
}

// max |w| of a tensor (one block; the tensors are <= 147456 floats)
__global__ void __launch_bounds__(1024) absmax_kernel(const float *__restrict__ w, int n, float *__restrict__ out) {
    __shared__ float red[32];
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = red[threadIdx.x];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
        if (threadIdx.x == 0) *out = m;
    }
}

// Head packing: mu/logvar [16,2048] in NCHW-flatten order (c*16 + hw) -> wh [32][hw*128 + c].
__global__ void pack_head_weights_kernel(const float *__restrict__ mu_w, const float *__restrict__ lv_w,
                                         const float *__restrict__ mu_b, const float *__restrict__ lv_b,
                                         float *__restrict__ wh, float *__restrict__ bh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 32) bh[i] = i < 16 ? mu_b[i] : lv_b[i - 16];
    if (i >= 32 * 2048) return;
    const int o = i / 2048, f = i % 2048;
    const int hw = f / 128, c = f % 128;
    const float *src = o < 16 ? mu_w + o * 2048 : lv_w + (o - 16) * 2048;
    wh[i] = src[c * 16 + hw];
}

// ---------------------------------------------------------------------------------------------
// conv0 plane statistics by running conv0 on CUDA cores (float32 patterns, or uint8 patterns that are not 4-byte
// aligned; aligned uint8 takes the exact integer-autocorrelation kernel in encoder_fused.cuh).
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

// grid = (16 row groups, nimg); block = 256 threads = 8 rows x 32 pixel quads; sums must be zero on entry
template <bool U8>
__global__ void __launch_bounds__(256) conv0_stats_kernel(const void *__restrict__ patterns,
                                                          const float *__restrict__ w0, double *__restrict__ sums) {
    __shared__ float ws[9 * 32];
    __shared__ float red[8][32][2];
    griddep_launch_dependents();
    for (int i = threadIdx.x; i < 9 * 32; i += 256) ws[i] = w0[i];
    __syncthreads();
    griddep_wait();   // sums is accumulated atomically and was zeroed by earlier work of the stream
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

// Running sum of fp32 partial sums as an unevaluated (hi, lo) pair (Knuth's TwoSum: hi + lo carries the sum to ~2^-48
// relative, whatever the order the partials arrive in), converted to fp64 only when it is flushed to memory.  The plane
// statistics used to be accumulated as `double += (double)partial`: F2F.F64.F32 + DADD per (tile, channel block) cost the
// epilogue warps 5-7 % of a whole block (ncu: math-pipe throttle on exactly those lines; role profile flag 32).
struct PairSum {
    float hi, lo;
    __device__ __forceinline__ void clear() { hi = lo = 0.f; }
    __device__ __forceinline__ void add(float x) {
        const float s = __fadd_rn(hi, x);
        const float bb = __fadd_rn(s, -hi);
        const float e = __fadd_rn(__fadd_rn(hi, -__fadd_rn(s, -bb)), __fadd_rn(x, -bb));
        hi = s;
        lo = __fadd_rn(lo, e);
    }
    __device__ __forceinline__ double value() const { return (double)hi + (double)lo; }
};

}  // namespace ebsd
