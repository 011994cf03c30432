// Memory-bound pieces of the visual frontend (NHWC fp32): 3x3 max-pool after the stem, depthwise 3x3
// convolutions and the pass-through half of the ShuffleNetV2 blocks, global average pool + L2 norm.
// The 1x1 convolutions and the Conv3d stem go through gemm.cuh (reference video.py:68-87,
// shufflenetv2.py:26-104,151-152).
#pragma once
#include "common.cuh"

namespace l2s {

// MaxPool3d((1,3,3),(1,2,2),(0,1,1)) on NHWC [N,H,W,C] -> [N,H/2,W/2,C]; C % 4 == 0.
__global__ void maxpool3x3s2_kernel(const float* __restrict__ x, float* __restrict__ y,
                                    int N, int H, int W, int C, int Ho, int Wo) {
    const int c4n = C >> 2;
    const size_t total = (size_t)N * Ho * Wo * c4n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int c4 = i % c4n; size_t r = i / c4n;
        int wo = r % Wo; r /= Wo;
        int ho = r % Ho; int n = r / Ho;
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int dh = -1; dh <= 1; ++dh) {
            int hi = 2 * ho + dh;
            if (hi < 0 || hi >= H) continue;
#pragma unroll
            for (int dw = -1; dw <= 1; ++dw) {
                int wi = 2 * wo + dw;
                if (wi < 0 || wi >= W) continue;
                float4 v = *reinterpret_cast<const float4*>(x + (((size_t)n * H + hi) * W + wi) * C + c4 * 4);
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        *reinterpret_cast<float4*>(y + (((size_t)n * Ho + ho) * Wo + wo) * C + c4 * 4) = m;
    }
}

// Depthwise 3x3, pad 1, stride s, BN folded: y[n,ho,wo,c] = sum w[k][c]*x[...] + b[c].
// x rows have ldx floats (channels [xoff, xoff+C)), y rows ldy floats at yoff; C % 4 == 0; w is [9][C].
__global__ void dwconv3x3_kernel(const float* __restrict__ x, int ldx, int xoff, float* __restrict__ y, int ldy, int yoff,
                                 const float* __restrict__ w, const float* __restrict__ bias,
                                 int N, int H, int W, int C, int stride, int Ho, int Wo) {
    const int c4n = C >> 2;
    const size_t total = (size_t)N * Ho * Wo * c4n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int c4 = i % c4n; size_t r = i / c4n;
        int wo = r % Wo; r /= Wo;
        int ho = r % Ho; int n = r / Ho;
        float4 acc = __ldg(reinterpret_cast<const float4*>(bias + c4 * 4));
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            int hi = ho * stride + kh - 1;
            if (hi < 0 || hi >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                int wi = wo * stride + kw - 1;
                if (wi < 0 || wi >= W) continue;
                float4 v = *reinterpret_cast<const float4*>(x + (((size_t)n * H + hi) * W + wi) * ldx + xoff + c4 * 4);
                float4 k = __ldg(reinterpret_cast<const float4*>(w + (kh * 3 + kw) * C + c4 * 4));
                acc.x = fmaf(v.x, k.x, acc.x); acc.y = fmaf(v.y, k.y, acc.y);
                acc.z = fmaf(v.z, k.z, acc.z); acc.w = fmaf(v.w, k.w, acc.w);
            }
        }
        *reinterpret_cast<float4*>(y + (((size_t)n * Ho + ho) * Wo + wo) * ldy + yoff + c4 * 4) = acc;
    }
}

// Stride-1 variant: one thread produces WT consecutive outputs along w for its channel quad, keeping the nine weights and
// a (WT+2)-wide input window in registers: (WT+2)*3 activation loads per WT outputs instead of 9 per output.
template <int WT>
__global__ void dwconv3x3_s1_strip_kernel(const float* __restrict__ x, int ldx, int xoff, float* __restrict__ y, int ldy, int yoff,
                                          const float* __restrict__ w, const float* __restrict__ bias, int N, int H, int W, int C) {
    const int c4n = C >> 2, nstrip = W / WT;
    const size_t total = (size_t)N * H * nstrip * c4n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int c4 = i % c4n; size_t r = i / c4n;
        int st = r % nstrip; r /= nstrip;
        int ho = r % H; int n = r / H;
        const int wo0 = st * WT;
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c4 * 4));
        float4 acc[WT];
#pragma unroll
        for (int j = 0; j < WT; ++j) acc[j] = b;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int hi = ho + kh - 1;
            if (hi < 0 || hi >= H) continue;
            float4 k[3];
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) k[kw] = __ldg(reinterpret_cast<const float4*>(w + (kh * 3 + kw) * C + c4 * 4));
            const float* row = x + ((size_t)n * H + hi) * W * ldx + xoff + c4 * 4;
            float4 v[WT + 2];
#pragma unroll
            for (int j = 0; j < WT + 2; ++j) {
                const int wi = wo0 + j - 1;
                v[j] = (wi >= 0 && wi < W) ? *reinterpret_cast<const float4*>(row + (size_t)wi * ldx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < WT; ++j)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    acc[j].x = fmaf(v[j + kw].x, k[kw].x, acc[j].x); acc[j].y = fmaf(v[j + kw].y, k[kw].y, acc[j].y);
                    acc[j].z = fmaf(v[j + kw].z, k[kw].z, acc[j].z); acc[j].w = fmaf(v[j + kw].w, k[kw].w, acc[j].w);
                }
        }
#pragma unroll
        for (int j = 0; j < WT; ++j)
            *reinterpret_cast<float4*>(y + (((size_t)n * H + ho) * W + wo0 + j) * ldy + yoff + c4 * 4) = acc[j];
    }
}

// Pass-through half of a stride-1 block followed by channel_shuffle(2): logical output channel 2j takes
// logical input channel j (< half).  Physical channel of logical l: l < half ? l : l - half + hp.
__global__ void shuffle_passthrough_kernel(const float* __restrict__ x, float* __restrict__ y, size_t rows, int ld, int half, int hp) {
    const size_t total = rows * half;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int j = i % half; size_t m = i / half;
        int l = 2 * j;
        int ph = l < half ? l : l - half + hp;
        y[m * ld + ph] = x[m * ld + j];
    }
}

// conv_last output [N, P, C] (P spatial positions) -> mean over P -> L2-normalise over C -> out [N, C]
// (AvgPool2d(3) + F.normalize(p=2,dim=2), reference shufflenetv2.py:152, video.py:85).  One CTA per frame.
__global__ void avgpool_l2norm_kernel(const float* __restrict__ x, float* __restrict__ out, int P, int C) {
    extern __shared__ float sm[];
    const int n = blockIdx.x;
    float ss = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f;
        for (int q = 0; q < P; ++q) a += x[((size_t)n * P + q) * C + c];
        a /= (float)P;
        sm[c] = a;
        ss += a * a;
    }
    __shared__ float wsum[32];
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? wsum[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) wsum[0] = v;
    }
    __syncthreads();
    const float denom = fmaxf(sqrtf(wsum[0]), 1e-12f);
    for (int c = threadIdx.x; c < C; c += blockDim.x) out[(size_t)n * C + c] = sm[c] / denom;
}

}  // namespace l2s
