// Generic fp32 SIMT GEMM with implicit conv1d/conv3d-stem addressing and a fused epilogue.
//
//   C[m, cmap(n)] = epi( sum_kk A(m, kk) * W[n, kk] )            W is [N][Ktot] row-major
//
// A(m, kk) addressing modes
//   dense / conv1d : rows are grouped in sequences of L_out; kk = tap*Kc + ci; the source row is
//                    seq*L_in + t*stride + tap - pad (zero outside [0,L_in)), i.e. Conv1d over a
//                    time-major [B, L, C] activation (reference decoder.py:86-104,159-196,209-230).
//   stem           : implicit im2col of Conv3d(3->24,(5,7,7),s(1,2,2),p(2,3,3)) over the caller's
//                    NCDHW video (reference video.py:68-69); kk = ((ci*5+kt)*7+kh)*7+kw.
// Epilogue: +bias[n] +addrow[seq][n] -> act -> +addpos[t][n] +resid[m][n] -> store (optionally through
// a channel map that realises ShuffleNetV2's concat+channel_shuffle, or transposed to [B, N, L]).
//
// This is the parity-first path (exact fp32 FMA).  Tensor-core (tcgen05) variants replace it for the
// large convolutions once parity is pinned.
#pragma once
#include "common.cuh"

namespace l2s {

struct GemmParams {
    const float* A; int lda;
    const float* W;                 // [N][taps*Kc]
    float* C; int ldc;
    int M, N, Kc, taps;
    int L_out, L_in, pad, stride;   // dense: taps=1,pad=0,stride=1,L_out=L_in=M
    const float* bias;              // [N] or null
    int act; const float* act_w;    // activation + per-channel parameter
    const float* addrow;            // [M/L_out][N] or null (added before act)
    const float* addpos; int ldpos; // [L_out][ldpos] or null (added after act)
    const float* resid; int ldr;    // [M][ldr] or null (added after act)
    int cstride, coff, chalf, chp;  // channel map: l = n*cstride+coff; phys = l<chalf ? l : l-chalf+chp (chalf>0)
    int transposed;                 // 1: C[(seq*N + n)*L_out + t]
    // stem mode
    int stem; int T, H, Wd, Ho, Wo;
};

constexpr int GEMM_BK = 16;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_kernel(const GemmParams p) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int BK = GEMM_BK;
    constexpr int LDA = BM + 4, LDW = BN + 4;
    __shared__ __align__(16) float As[2][BK][LDA];
    __shared__ __align__(16) float Ws[2][BK][LDW];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int Ktot = p.taps * p.Kc;
    const int nkt = (Ktot + BK - 1) / BK;

    // ---- loader roles: each thread loads float4 (4 consecutive k) for fixed rows ----------------
    constexpr int A_F4 = BM * (BK / 4);            // float4 slots in the A tile
    constexpr int W_F4 = BN * (BK / 4);
    constexpr int A_PER = (A_F4 + NT - 1) / NT;
    constexpr int W_PER = (W_F4 + NT - 1) / NT;
    const bool a_vec = (!p.stem) && ((p.lda & 3) == 0) && ((p.Kc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
    const bool w_vec = ((Ktot & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.W) & 15) == 0);

    float4 areg[A_PER], wreg[W_PER];

    // per-slot row bookkeeping for A (independent of k-tile)
    int a_seqbase[A_PER], a_t[A_PER];
    bool a_ok[A_PER];
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
        int slot = tid + i * NT;
        int r = slot / (BK / 4);
        int m = m0 + r;
        a_ok[i] = (slot < A_F4) && (m < p.M);
        int mm = a_ok[i] ? m : 0;
        if (!p.stem) {
            int seq = mm / p.L_out, t = mm - seq * p.L_out;
            a_seqbase[i] = seq * p.L_in;
            a_t[i] = t * p.stride - p.pad;
        } else {
            // m -> (n, ho, wo); store n*? in seqbase and packed (ho,wo) in t
            int hw = p.Ho * p.Wo;
            int n = mm / hw, r2 = mm - n * hw;
            a_seqbase[i] = n;
            a_t[i] = r2;
        }
    }

    auto load_tiles = [&](int kt) {
        const int k0 = kt * BK;
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int slot = tid + i * NT;
            int kq = slot % (BK / 4);
            int kk = k0 + kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_ok[i] && kk < Ktot) {
                if (!p.stem) {
                    int tap = kk / p.Kc, ci = kk - tap * p.Kc;
                    int tt = a_t[i] + tap;
                    if (tt >= 0 && tt < p.L_in) {
                        const float* src = p.A + (size_t)(a_seqbase[i] + tt) * p.lda + ci;
                        if (a_vec) {
                            v = *reinterpret_cast<const float4*>(src);
                        } else {
                            v.x = src[0];
                            if (ci + 1 < p.Kc) v.y = src[1];
                            if (ci + 2 < p.Kc) v.z = src[2];
                            if (ci + 3 < p.Kc) v.w = src[3];
                        }
                    }
                } else {
                    int n = a_seqbase[i];
                    int b = n / p.T, t = n - b * p.T;
                    int ho = a_t[i] / p.Wo, wo = a_t[i] - ho * p.Wo;
                    float tmp[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        int k = kk + j;
                        float x = 0.f;
                        if (k < Ktot) {
                            int ci = k / 245, r = k - ci * 245;
                            int kt3 = r / 49, r2 = r - kt3 * 49;
                            int kh = r2 / 7, kw = r2 - kh * 7;
                            int ti = t + kt3 - 2, hi = 2 * ho + kh - 3, wi = 2 * wo + kw - 3;
                            if (ti >= 0 && ti < p.T && hi >= 0 && hi < p.H && wi >= 0 && wi < p.Wd)
                                x = __ldg(p.A + ((size_t)((b * 3 + ci) * p.T + ti) * p.H + hi) * p.Wd + wi);
                        }
                        tmp[j] = x;
                    }
                    v = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
                }
            }
            areg[i] = v;
        }
#pragma unroll
        for (int i = 0; i < W_PER; ++i) {
            int slot = tid + i * NT;
            int r = slot / (BK / 4), kq = slot % (BK / 4);
            int n = n0 + r, kk = k0 + kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (slot < W_F4 && n < p.N && kk < Ktot) {
                const float* src = p.W + (size_t)n * Ktot + kk;
                if (w_vec) {
                    v = __ldg(reinterpret_cast<const float4*>(src));
                } else {
                    v.x = __ldg(src);
                    if (kk + 1 < Ktot) v.y = __ldg(src + 1);
                    if (kk + 2 < Ktot) v.z = __ldg(src + 2);
                    if (kk + 3 < Ktot) v.w = __ldg(src + 3);
                }
            }
            wreg[i] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int slot = tid + i * NT;
            if (slot < A_F4) {
                int r = slot / (BK / 4), kq = slot % (BK / 4);
                As[buf][kq * 4 + 0][r] = areg[i].x;
                As[buf][kq * 4 + 1][r] = areg[i].y;
                As[buf][kq * 4 + 2][r] = areg[i].z;
                As[buf][kq * 4 + 3][r] = areg[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < W_PER; ++i) {
            int slot = tid + i * NT;
            if (slot < W_F4) {
                int r = slot / (BK / 4), kq = slot % (BK / 4);
                Ws[buf][kq * 4 + 0][r] = wreg[i].x;
                Ws[buf][kq * 4 + 1][r] = wreg[i].y;
                Ws[buf][kq * 4 + 2][r] = wreg[i].z;
                Ws[buf][kq * 4 + 3][r] = wreg[i].w;
            }
        }
    };

    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nkt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nkt) load_tiles(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                float4 v = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * TN + j]);
                b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nkt) store_tiles(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue -------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= p.M) continue;
        int seq = m / p.L_out, t = m - seq * p.L_out;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * TN + j;
            if (n >= p.N) continue;
            float v = acc[i][j];
            if (p.bias) v += __ldg(p.bias + n);
            if (p.addrow) v += p.addrow[(size_t)seq * p.N + n];
            v = apply_act(v, p.act, p.act_w ? __ldg(p.act_w + n) : 1.f);
            if (p.addpos) v += __ldg(p.addpos + (size_t)t * p.ldpos + n);
            if (p.resid) v += p.resid[(size_t)m * p.ldr + n];
            if (p.transposed) {
                p.C[((size_t)seq * p.N + n) * p.L_out + t] = v;
            } else {
                int l = n * p.cstride + p.coff;
                if (p.chalf > 0 && l >= p.chalf) l = l - p.chalf + p.chp;
                p.C[(size_t)m * p.ldc + l] = v;
            }
        }
    }
}

inline GemmParams gemm_defaults() {
    GemmParams p{};
    p.taps = 1; p.pad = 0; p.stride = 1; p.cstride = 1;
    return p;
}

inline cudaError_t launch_gemm(GemmParams p, cudaStream_t s) {
    if (p.L_out == 0) { p.L_out = p.M; p.L_in = p.M; }
    if (p.N <= 32) {
        dim3 grid(ceil_div(p.M, 128), ceil_div(p.N, 32));
        gemm_kernel<128, 32, 4, 4><<<grid, 256, 0, s>>>(p);
    } else if (p.M <= 512) {
        dim3 grid(ceil_div(p.M, 32), ceil_div(p.N, 64));
        gemm_kernel<32, 64, 4, 4><<<grid, 128, 0, s>>>(p);
    } else {
        dim3 grid(ceil_div(p.M, 64), ceil_div(p.N, 64));
        gemm_kernel<64, 64, 4, 4><<<grid, 256, 0, s>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace l2s
