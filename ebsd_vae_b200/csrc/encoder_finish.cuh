// InstanceNorm "finisher": y = LeakyReLU_0.02((x - mean) * rstd), optional 2x2 max-pool, in the layout the next
// layer wants.  (latice/model.py:96-97 + the MaxPool2d modules at latice/model.py:112-124.)
//
//   mean, rstd come from the per-(image, channel) sum / sum of squares the conv epilogue accumulated (fp64):
//   biased variance over H*W, eps = 1e-5 added before the reciprocal square root (torch instance_norm semantics).
//   They are evaluated ONCE per CTA into shared memory (fp64 sqrt/div per output element would dominate).
//   Pooling is applied to the raw values: x -> leaky((x-mean)*rstd) is increasing, so pool(f(x)) = f(pool(x)).
//
// Output modes
//   FIN_F32        fp32 NHWC [n,Ho,Wo,C]                       (CUDA-core path, and the features for the heads)
//   FIN_SPLIT      fp16 hi / lo planes NHWC [n,Ho,Wo,C]        (first-generation tensor kernel)
//   FIN_SPLIT_PAD  fp16 hi / lo planes [n,Ho+2,Wo+2,C], zero borders (shifted-window tensor kernel)
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace ebsd {

enum FinishMode { FIN_F32 = 0, FIN_SPLIT = 1, FIN_SPLIT_PAD = 2 };

constexpr int kFinishItemsPerThread = 4;

// grid = (ceil(items_per_image / (256*4)), nimg); one item = 4 channels of one output position
template <int CH, bool POOL, int MODE>
__global__ void __launch_bounds__(256) finish_kernel(const float *__restrict__ raw, const double *__restrict__ sums,
                                                     void *__restrict__ out_a, void *__restrict__ out_b, int H,
                                                     int W) {
    __shared__ float s_mean[CH], s_rstd[CH];
    const long long n = blockIdx.y;
    for (int c = threadIdx.x; c < CH; c += 256) {
        const double inv_hw = 1.0 / ((double)H * (double)W);
        const double mm = sums[(n * CH + c) * 2 + 0] * inv_hw;
        double var = sums[(n * CH + c) * 2 + 1] * inv_hw - mm * mm;
        if (var < 0.0) var = 0.0;
        s_mean[c] = (float)mm;
        s_rstd[c] = (float)(1.0 / sqrt(var + 1e-5));
    }
    __syncthreads();
    const int Ho = POOL ? H / 2 : H, Wo = POOL ? W / 2 : W;
    constexpr int PAD = MODE == FIN_SPLIT_PAD ? 1 : 0;
    const int Hq = Ho + 2 * PAD, Wq = Wo + 2 * PAD;
    const int items = Hq * Wq * (CH / 4);
#pragma unroll
    for (int r = 0; r < kFinishItemsPerThread; ++r) {
        const int item = (blockIdx.x * kFinishItemsPerThread + r) * 256 + threadIdx.x;
        if (item >= items) break;
        const int c4 = item % (CH / 4);
        const int pos = item / (CH / 4);
        const int xq = pos % Wq, yq = pos / Wq;
        const long long off = ((n * Hq + yq) * Wq + xq) * CH + c4 * 4;
        if (PAD && (yq == 0 || yq == Hq - 1 || xq == 0 || xq == Wq - 1)) {
            *(uint2 *)((__half *)out_a + off) = make_uint2(0u, 0u);
            *(uint2 *)((__half *)out_b + off) = make_uint2(0u, 0u);
            continue;
        }
        const int yo = yq - PAD, xo = xq - PAD;
        float4 v;
        if (POOL) {
            const float *q = raw + ((n * H + yo * 2) * W + xo * 2) * CH + c4 * 4;
            const float4 a = *(const float4 *)q, b = *(const float4 *)(q + CH);
            const float4 c = *(const float4 *)(q + (long long)W * CH), d = *(const float4 *)(q + (long long)W * CH + CH);
            v.x = fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x));
            v.y = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
            v.z = fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z));
            v.w = fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w));
        } else {
            v = *(const float4 *)(raw + ((n * H + yo) * W + xo) * CH + c4 * 4);
        }
        float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float t = (o[j] - s_mean[c4 * 4 + j]) * s_rstd[c4 * 4 + j];
            o[j] = t >= 0.f ? t : t * 0.02f;
        }
        if (MODE == FIN_F32) {
            *(float4 *)((float *)out_a + off) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
            __half h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                h[j] = __float2half_rn(o[j]);
                l[j] = __float2half_rn(o[j] - __half2float(h[j]));
            }
            *(uint2 *)((__half *)out_a + off) = *(const uint2 *)h;
            *(uint2 *)((__half *)out_b + off) = *(const uint2 *)l;
        }
    }
}

}  // namespace ebsd
