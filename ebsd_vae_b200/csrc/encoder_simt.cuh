// Encoder building blocks that run on the CUDA cores (fp32): the first convolution (Cin = 1), the
// InstanceNorm plane statistics, the normalise + LeakyReLU + max-pool "finisher", the mu/logvar heads and
// a plain fp32 direct convolution used as the bring-up path (EBSD_ENCODER_PATH=simt) and as the in-library
// cross-check of the tensor-core path.
//
// Reference semantics: latice/model.py:93-98 (Conv3x3 s1 p1 -> InstanceNorm2d(eps=1e-5, biased variance,
// affine=False) -> LeakyReLU(0.02)), pools at latice/model.py:112-124, heads latice/model.py:57-58,127-129.
// The conv bias is dropped: InstanceNorm subtracts the per-(n,c) plane mean, which cancels it exactly
// (tests/test_oracle_encoder.py::test_conv_bias_is_cancelled_by_instance_norm).
// MaxPool is applied to the raw accumulators' normalised values after the plane statistics were taken over
// the un-pooled plane; since x -> leaky((x-mean)*rstd) is increasing, pool(f(x)) == f(pool(x)).
#pragma once
#include "common.cuh"

namespace ebsd {

constexpr float kLeaky = 0.02f;
constexpr double kInEps = 1e-5;

// ---------------------------------------------------------------------------------------------
// conv0: [B,128,128] (u8 or f32) -> raw [B,128,128,32] fp32 NHWC.  Weights w0[tap][co] (9 x 32).
// One thread = 2 horizontally adjacent pixels x 32 output channels.
// ---------------------------------------------------------------------------------------------
template <bool U8>
__global__ void __launch_bounds__(256) conv0_kernel(const void *__restrict__ patterns, const float *__restrict__ w0,
                                                    float *__restrict__ raw, long long B) {
    __shared__ float ws[9 * 32];
    for (int i = threadIdx.x; i < 9 * 32; i += blockDim.x) ws[i] = w0[i];
    __syncthreads();
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // pixel pair id
    const long long total = B * 128 * 64;
    if (gid >= total) return;
    const int xp = (int)(gid & 63);
    const int y = (int)((gid >> 6) & 127);
    const long long n = gid >> 13;
    const int x0 = xp * 2;
    float in[3][4];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int yy = y + dy - 1;
#pragma unroll
        for (int dx = 0; dx < 4; ++dx) {
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
    float4 *out0 = (float4 *)(raw + ((n * 128 + y) * 128 + x0) * 32);
    float4 *out1 = out0 + 8;
#pragma unroll
    for (int cg = 0; cg < 8; ++cg) {
        float4 a0 = make_float4(0, 0, 0, 0), a1 = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float4 w = *(const float4 *)(ws + (dy * 3 + dx) * 32 + cg * 4);
                const float p0 = in[dy][dx], p1 = in[dy][dx + 1];
                a0.x = fmaf(p0, w.x, a0.x); a0.y = fmaf(p0, w.y, a0.y); a0.z = fmaf(p0, w.z, a0.z); a0.w = fmaf(p0, w.w, a0.w);
                a1.x = fmaf(p1, w.x, a1.x); a1.y = fmaf(p1, w.y, a1.y); a1.z = fmaf(p1, w.z, a1.z); a1.w = fmaf(p1, w.w, a1.w);
            }
        out0[cg] = a0;
        out1[cg] = a1;
    }
}

// ---------------------------------------------------------------------------------------------
// Plain fp32 direct convolution (bring-up / cross-check path).
//   in  : finished activations [B,H,W,CIN] fp32 NHWC
//   wt  : [tap][ci][co] fp32
//   raw : [B,H,W,COUT] fp32 NHWC
// CTA = 8x8 output pixels x CO_TILE channels; thread = 4 pixels (along x) x 4 channels.
// ---------------------------------------------------------------------------------------------
template <int CIN, int COUT>
struct SimtConvCfg {
    static constexpr int CO_TILE = COUT < 64 ? COUT : 64;
    static constexpr int THREADS = (CO_TILE / 4) * 16;
    static constexpr int CI_CHUNK = 16;
    static constexpr int smem_floats = 10 * 10 * CI_CHUNK + 9 * CI_CHUNK * CO_TILE;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(SimtConvCfg<CIN, COUT>::THREADS)
conv3x3_simt_kernel(const float *__restrict__ in, const float *__restrict__ wt, float *__restrict__ raw, int H, int W) {
    using C = SimtConvCfg<CIN, COUT>;
    extern __shared__ float sm[];
    float *s_in = sm;                          // [10][10][CI_CHUNK]
    float *s_w = sm + 10 * 10 * C::CI_CHUNK;   // [9][CI_CHUNK][CO_TILE]
    const int tiles_x = W / 8;
    const int tile = blockIdx.x;
    const int ty0 = (tile / tiles_x) * 8, tx0 = (tile % tiles_x) * 8;
    const int co0 = blockIdx.y * C::CO_TILE;
    const long long n = blockIdx.z;
    const int tid = threadIdx.x;
    const int cg = tid % (C::CO_TILE / 4);  // 4-channel group
    const int pg = tid / (C::CO_TILE / 4);  // pixel group 0..15
    const int py = pg >> 1, px0 = (pg & 1) * 4;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int c0 = 0; c0 < CIN; c0 += C::CI_CHUNK) {
        __syncthreads();
        for (int v = tid; v < 10 * 10 * (C::CI_CHUNK / 4); v += C::THREADS) {
            const int c4 = v % (C::CI_CHUNK / 4);
            const int pos = v / (C::CI_CHUNK / 4);
            const int yy = ty0 + pos / 10 - 1, xx = tx0 + pos % 10 - 1;
            float4 val = make_float4(0, 0, 0, 0);
            if (yy >= 0 && yy < H && xx >= 0 && xx < W)
                val = *(const float4 *)(in + ((n * H + yy) * W + xx) * CIN + c0 + c4 * 4);
            *(float4 *)(s_in + pos * C::CI_CHUNK + c4 * 4) = val;
        }
        for (int v = tid; v < 9 * C::CI_CHUNK * (C::CO_TILE / 4); v += C::THREADS) {
            const int o4 = v % (C::CO_TILE / 4);
            const int rest = v / (C::CO_TILE / 4);
            const int ci = rest % C::CI_CHUNK, tap = rest / C::CI_CHUNK;
            *(float4 *)(s_w + (tap * C::CI_CHUNK + ci) * C::CO_TILE + o4 * 4) =
                *(const float4 *)(wt + ((long long)tap * CIN + c0 + ci) * COUT + co0 + o4 * 4);
        }
        __syncthreads();
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;
            const float *ip = s_in + ((py + dy) * 10 + px0 + dx) * C::CI_CHUNK;
            const float *wp = s_w + tap * C::CI_CHUNK * C::CO_TILE + cg * 4;
#pragma unroll 4
            for (int ci = 0; ci < C::CI_CHUNK; ++ci) {
                const float4 w = *(const float4 *)(wp + ci * C::CO_TILE);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float a = ip[i * C::CI_CHUNK + ci];
                    acc[i][0] = fmaf(a, w.x, acc[i][0]);
                    acc[i][1] = fmaf(a, w.y, acc[i][1]);
                    acc[i][2] = fmaf(a, w.z, acc[i][2]);
                    acc[i][3] = fmaf(a, w.w, acc[i][3]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int yy = ty0 + py, xx = tx0 + px0 + i;
        *(float4 *)(raw + ((n * H + yy) * W + xx) * COUT + co0 + cg * 4) =
            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

// ---------------------------------------------------------------------------------------------
// Plane statistics: raw [B,HW,C] fp32 -> sums [B,C,2] float64 (sum, sum of squares), accumulated with
// atomics over `gridDim.y` pixel slices.  `sums` must be zeroed first.
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) plane_stats_kernel(const float *__restrict__ raw, double *__restrict__ sums,
                                                          int HW) {
    constexpr int PG = 256 / C;  // pixel lanes per channel (C <= 256)
    __shared__ double red[2][256];
    const long long n = blockIdx.x;
    const int slice = blockIdx.y, nslices = gridDim.y;
    const int c = threadIdx.x % C, pl = threadIdx.x / C;
    const int per = (HW + nslices - 1) / nslices;
    const int p0 = slice * per, p1 = min(HW, p0 + per);
    double s = 0.0, ss = 0.0;
    for (int p = p0 + pl; p < p1; p += PG) {
        const double v = (double)raw[(n * HW + p) * C + c];
        s += v;
        ss += v * v;
    }
    red[0][threadIdx.x] = s;
    red[1][threadIdx.x] = ss;
    __syncthreads();
    if (threadIdx.x < C) {
        for (int g = 1; g < PG; ++g) {
            s += red[0][g * C + c];
            ss += red[1][g * C + c];
        }
        atomicAdd(&sums[(n * C + c) * 2 + 0], s);
        atomicAdd(&sums[(n * C + c) * 2 + 1], ss);
    }
}

// ---------------------------------------------------------------------------------------------
// Finisher: y = leaky((x - mean) * rstd), optional 2x2 max-pool, fp32 NHWC out.
// One thread = 4 channels of one output pixel.
// ---------------------------------------------------------------------------------------------
template <int C, bool POOL>
__global__ void __launch_bounds__(256) finish_f32_kernel(const float *__restrict__ raw, const double *__restrict__ sums,
                                                         float *__restrict__ out, int H, int W, long long B) {
    const int Ho = POOL ? H / 2 : H, Wo = POOL ? W / 2 : W;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = B * Ho * Wo * (C / 4);
    if (gid >= total) return;
    const int c4 = (int)(gid % (C / 4));
    long long r = gid / (C / 4);
    const int xo = (int)(r % Wo);
    r /= Wo;
    const int yo = (int)(r % Ho);
    const long long n = r / Ho;
    const double inv_hw = 1.0 / ((double)H * (double)W);
    float mean[4], rstd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double s = sums[(n * C + c4 * 4 + j) * 2 + 0], ss = sums[(n * C + c4 * 4 + j) * 2 + 1];
        const double m = s * inv_hw;
        double var = ss * inv_hw - m * m;
        if (var < 0.0) var = 0.0;
        mean[j] = (float)m;
        rstd[j] = (float)(1.0 / sqrt(var + kInEps));
    }
    float4 v;
    if (POOL) {
        const float *p = raw + ((n * H + yo * 2) * W + xo * 2) * C + c4 * 4;
        const float4 a = *(const float4 *)p, b = *(const float4 *)(p + C);
        const float4 c = *(const float4 *)(p + (long long)W * C), d = *(const float4 *)(p + (long long)W * C + C);
        v.x = fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x));
        v.y = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
        v.z = fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z));
        v.w = fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w));
    } else {
        v = *(const float4 *)(raw + ((n * H + yo) * W + xo) * C + c4 * 4);
    }
    float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float t = (o[j] - mean[j]) * rstd[j];
        o[j] = t >= 0.f ? t : t * kLeaky;
    }
    *(float4 *)(out + ((n * Ho + yo) * Wo + xo) * C + c4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
}

// ---------------------------------------------------------------------------------------------
// Heads: feat [B,2048] (NHWC flatten: (h*4+w)*128 + c) x wh [32,2048] (rows 0..15 mu, 16..31 logvar,
// permuted to the same order) + bias -> mu [B,16], logvar [B,16] (nullable).
// One CTA (8 warps) per image; warp w produces outputs 4w..4w+3.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) heads_kernel(const float *__restrict__ feat, const float *__restrict__ wh,
                                                    const float *__restrict__ bh, float *__restrict__ mu,
                                                    float *__restrict__ logvar) {
    const long long n = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4 *f = (const float4 *)(feat + n * 2048);
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

// Weight packing: torch [Cout,Cin,3,3] -> [tap][ci][co] fp32.
__global__ void pack_conv_weights_kernel(const float *__restrict__ w, float *__restrict__ out, int cin, int cout) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = 9 * cin * cout;
    if (i >= total) return;
    const int co = i % cout;
    const int ci = (i / cout) % cin;
    const int tap = i / (cout * cin);
    out[i] = w[((long long)co * cin + ci) * 9 + tap];
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

}  // namespace ebsd
