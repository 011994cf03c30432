// Pointwise (1x1) convolution of the ShuffleNetV2 trunk as a STREAMING tensor-core GEMM (reference
// shufflenetv2.py:51-89: every 1x1 Conv2d + BatchNorm2d + ReLU of the InvertedResidual blocks).
//
//   C[m, map(n)] = relu( sum_k A[m, k] * W[n, k] + bias[n] ),   M = frames*h*w rows (8 K .. 535 K), K, N in 24..232
//
// These GEMMs are tiny in K and N and huge in M: HBM-bound (64 MB in+out at stage 2 = 10 us) with ~1 GFLOP of math.
// The tcgen05 pipeline (TMA -> smem split -> MMA -> TMEM -> transposed store) pays its per-tile latency chain 7 times
// per SM at K = 64 and ran them at 7x the memory bound.  Here every warp streams 16-row tiles straight from global
// memory into mma.sync fragments (no staging, no CTA-level synchronisation in the loop), so 16 warps per SM keep enough
// loads in flight to cover HBM latency (measured at B=32: 50 us vs 70 us at stage 2, 37 vs 40 us at stage 3; tensor pipe
// 28-43 % active, DRAM traffic 2x the compulsory bytes because the two branches of a block write interleaved channels
// of the same sectors at different times — still 3-5x above the memory bound, the next thing to fix here):
//  * A: lane (g = lane>>2, t = lane&3) reads A[row g / g+8][k0 + 4t .. +3] as two LDG.128 per 16-wide k chunk (each row
//    contributes one full 64-byte segment per chunk), splits them into tf32 hi / lo in registers;
//  * W: the slice's rows, pre-split hi / lo at pack time, resident in shared memory ([64][Kp + 16] each, LDS.128
//    conflict-free); N is cut into slices of 64 columns (grid.y) so the accumulators stay at 32 registers;
//  * 3xTF32 error compensation (a_lo*w_hi + a_hi*w_lo + a_hi*w_hi), fp32 accumulation: same accuracy class as the
//    tcgen05 path; MMA column t / t+4 of k-step s <-> real k = k0 + 4t + 2s / + 2s + 1 for both operands;
//  * epilogue straight from the accumulator fragments: bias + ReLU, ShuffleNet's concat + channel_shuffle folded into
//    the store address (cstride / coff / chalf / chp as in gemm.cuh).
#pragma once
#include "matvec.cuh"

namespace l2s {

constexpr int PW_THREADS = 256;
constexpr int PW_NT = 8;                  // n-tiles of 8 columns per slice
constexpr int PW_NSLICE = 8 * PW_NT;      // 64 output columns per CTA

struct PwParams {
    const float* A; int lda;
    int M, N, Kc;
    const float* Whi; const float* Wlo; int kcp;       // packed [N][kcp], zero-padded beyond Kc
    const float* bias; int relu;
    float* C; int ldc;
    int cstride, coff, chalf, chp;
    int Kp, ldw;                                        // Kp = round_up(K per split, 16); ldw = Kp or Kp + 16, whichever is 16 mod 32
    // split-K mode (few rows, long K: the decoder pre-loop's small GEMMs): grid.z splits of `ksplit` columns write raw
    // partial sums to partial[z][M][N]; pw_reduce_kernel adds them in a fixed order and applies the epilogue
    int ksplit; float* partial;
    // stride-1 InvertedResidual blocks (shufflenetv2.py:97-100 + channel_shuffle): the pass-through half x[:, j] -> logical
    // channel 2j is written here together with the branch output (logical 2j+1), so every 32-byte sector of the block's
    // output is completed by one warp at once (no partial-sector fills from DRAM, no separate copy kernel).
    // Requires cstride == 2, coff == 1, N even.
    const float* pass_x; int pass_ld;
};

__global__ void __launch_bounds__(PW_THREADS, 2) pw_mma_kernel(const PwParams p) {
    extern __shared__ __align__(16) float smem[];
    float* wsm = smem;                                       // full-precision weights (hi + lo is exact), split per use
    const int n_base = blockIdx.y * PW_NSLICE;
    const int kb = blockIdx.z * p.ksplit;                    // this CTA's K range: [kb, kb + Kc)
    const int Kc = min(p.Kc - kb, p.ksplit);
    const int Kp = (Kc + 15) / 16 * 16;
    for (int i = threadIdx.x; i < PW_NSLICE * (Kp / 4); i += PW_THREADS) {
        const int r = i / (Kp / 4), c4 = i - r * (Kp / 4);
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f), l = h;
        if (n_base + r < p.N) {
            h = __ldg(reinterpret_cast<const float4*>(p.Whi + (size_t)(n_base + r) * p.kcp + kb) + c4);
            l = __ldg(reinterpret_cast<const float4*>(p.Wlo + (size_t)(n_base + r) * p.kcp + kb) + c4);
        }
        *reinterpret_cast<float4*>(wsm + (size_t)r * p.ldw + 4 * c4) = make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nchunks = Kp / 16;
    const int ntiles = (p.M + 15) / 16;
    const int wstride = gridDim.x * (PW_THREADS / 32);
    // per-lane epilogue constants: this lane owns columns n_base + 8*nt + 2t, +1
    const float* w_lane = wsm + (size_t)g * p.ldw + 4 * t;

    for (int tile = blockIdx.x * (PW_THREADS / 32) + warp; tile < ntiles; tile += wstride) {
        const int r0 = tile * 16 + g, r1 = r0 + 8;
        const bool v0 = r0 < p.M, v1 = r1 < p.M;
        const float* a0p = p.A + (size_t)(v0 ? r0 : 0) * p.lda + kb + 4 * t;
        const float* a1p = p.A + (size_t)(v1 ? r1 : 0) * p.lda + kb + 4 * t;
        float acc[PW_NT][4];
#pragma unroll
        for (int nt = 0; nt < PW_NT; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
        float4 xa = make_float4(0.f, 0.f, 0.f, 0.f), xb = xa;
        if (4 * t < Kc) { xa = __ldg(reinterpret_cast<const float4*>(a0p)); xb = __ldg(reinterpret_cast<const float4*>(a1p)); }
        for (int c = 0; c < nchunks; ++c) {
            float4 na = make_float4(0.f, 0.f, 0.f, 0.f), nb = na;
            if (c + 1 < nchunks && (c + 1) * 16 + 4 * t < Kc) {          // next chunk in flight while this one is multiplied
                na = __ldg(reinterpret_cast<const float4*>(a0p + (c + 1) * 16));
                nb = __ldg(reinterpret_cast<const float4*>(a1p + (c + 1) * 16));
            }
            const float av[4] = {xa.x, xa.y, xa.z, xa.w}, bv[4] = {xb.x, xb.y, xb.z, xb.w};
            uint32_t ah[4], al[4], bh[4], bl[4];                           // rows g (a*) and g+8 (b*)
#pragma unroll
            for (int i = 0; i < 4; ++i) { split_tf32(av[i], ah[i], al[i]); split_tf32(bv[i], bh[i], bl[i]); }
#pragma unroll
            for (int nt = 0; nt < PW_NT; ++nt) {
                const float4 w4 = *reinterpret_cast<const float4*>(w_lane + (size_t)nt * 8 * p.ldw + c * 16);
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
                uint32_t whv[4], wlv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) split_tf32(wv[i], whv[i], wlv[i]);
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    mma_tf32(acc[nt], al[2 * s], bl[2 * s], al[2 * s + 1], bl[2 * s + 1], whv[2 * s], whv[2 * s + 1]);
                    mma_tf32(acc[nt], ah[2 * s], bh[2 * s], ah[2 * s + 1], bh[2 * s + 1], wlv[2 * s], wlv[2 * s + 1]);
                    mma_tf32(acc[nt], ah[2 * s], bh[2 * s], ah[2 * s + 1], bh[2 * s + 1], whv[2 * s], whv[2 * s + 1]);
                }
            }
            xa = na; xb = nb;
        }
        // epilogue: acc[nt] = {(r0, n), (r0, n+1), (r1, n), (r1, n+1)}, n = n_base + 8 nt + 2 t
        if (p.partial) {
            float* P = p.partial + (size_t)blockIdx.z * p.M * p.N;
#pragma unroll
            for (int nt = 0; nt < PW_NT; ++nt) {
                const int n = n_base + nt * 8 + 2 * t;
                if (n + 1 < p.N && !(p.N & 1)) {                       // 8-byte stores need an even row stride
                    if (v0) *reinterpret_cast<float2*>(P + (size_t)r0 * p.N + n) = make_float2(acc[nt][0], acc[nt][1]);
                    if (v1) *reinterpret_cast<float2*>(P + (size_t)r1 * p.N + n) = make_float2(acc[nt][2], acc[nt][3]);
                } else {
                    if (n < p.N) {
                        if (v0) P[(size_t)r0 * p.N + n] = acc[nt][0];
                        if (v1) P[(size_t)r1 * p.N + n] = acc[nt][2];
                    }
                    if (n + 1 < p.N) {
                        if (v0) P[(size_t)r0 * p.N + n + 1] = acc[nt][1];
                        if (v1) P[(size_t)r1 * p.N + n + 1] = acc[nt][3];
                    }
                }
            }
            continue;
        }
        if (p.pass_x) {
#pragma unroll
            for (int nt = 0; nt < PW_NT; ++nt) {
                const int n = n_base + nt * 8 + 2 * t;                 // even; n + 1 < N because N is even
                if (n < p.N) {
                    const float b0 = p.bias ? __ldg(p.bias + n) : 0.f, b1 = p.bias ? __ldg(p.bias + n + 1) : 0.f;
                    int l0 = 2 * n, l1 = 2 * n + 2;                    // logical channels (2n, 2n+1) and (2n+2, 2n+3)
                    if (l0 >= p.chalf) l0 = l0 - p.chalf + p.chp;
                    if (l1 >= p.chalf) l1 = l1 - p.chalf + p.chp;
                    float y00 = acc[nt][0] + b0, y01 = acc[nt][1] + b1, y10 = acc[nt][2] + b0, y11 = acc[nt][3] + b1;
                    if (p.relu) { y00 = fmaxf(y00, 0.f); y01 = fmaxf(y01, 0.f); y10 = fmaxf(y10, 0.f); y11 = fmaxf(y11, 0.f); }
                    if (v0) {
                        const float2 x = __ldg(reinterpret_cast<const float2*>(p.pass_x + (size_t)r0 * p.pass_ld + n));
                        *reinterpret_cast<float2*>(p.C + (size_t)r0 * p.ldc + l0) = make_float2(x.x, y00);
                        *reinterpret_cast<float2*>(p.C + (size_t)r0 * p.ldc + l1) = make_float2(x.y, y01);
                    }
                    if (v1) {
                        const float2 x = __ldg(reinterpret_cast<const float2*>(p.pass_x + (size_t)r1 * p.pass_ld + n));
                        *reinterpret_cast<float2*>(p.C + (size_t)r1 * p.ldc + l0) = make_float2(x.x, y10);
                        *reinterpret_cast<float2*>(p.C + (size_t)r1 * p.ldc + l1) = make_float2(x.y, y11);
                    }
                }
            }
            continue;
        }
        if (p.cstride == 1 && !((p.coff | p.ldc | p.chalf | p.chp | p.N) & 1)) {
            // contiguous channels: the two columns a lane owns are neighbours in memory -> one 8-byte store per row
#pragma unroll
            for (int nt = 0; nt < PW_NT; ++nt) {
                const int n = n_base + nt * 8 + 2 * t;
                if (n < p.N) {
                    const float b0 = p.bias ? __ldg(p.bias + n) : 0.f, b1 = p.bias ? __ldg(p.bias + n + 1) : 0.f;
                    int l = n + p.coff;
                    if (p.chalf > 0 && l >= p.chalf) l = l - p.chalf + p.chp;
                    float y00 = acc[nt][0] + b0, y01 = acc[nt][1] + b1, y10 = acc[nt][2] + b0, y11 = acc[nt][3] + b1;
                    if (p.relu) { y00 = fmaxf(y00, 0.f); y01 = fmaxf(y01, 0.f); y10 = fmaxf(y10, 0.f); y11 = fmaxf(y11, 0.f); }
                    if (v0) *reinterpret_cast<float2*>(p.C + (size_t)r0 * p.ldc + l) = make_float2(y00, y01);
                    if (v1) *reinterpret_cast<float2*>(p.C + (size_t)r1 * p.ldc + l) = make_float2(y10, y11);
                }
            }
            continue;
        }
#pragma unroll
        for (int nt = 0; nt < PW_NT; ++nt) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = n_base + nt * 8 + 2 * t + j;
                if (n < p.N) {
                    const float b = p.bias ? __ldg(p.bias + n) : 0.f;
                    int l = n * p.cstride + p.coff;
                    if (p.chalf > 0 && l >= p.chalf) l = l - p.chalf + p.chp;
                    float y0 = acc[nt][j] + b, y1 = acc[nt][2 + j] + b;
                    if (p.relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
                    if (v0) p.C[(size_t)r0 * p.ldc + l] = y0;
                    if (v1) p.C[(size_t)r1 * p.ldc + l] = y1;
                }
            }
        }
    }
}

inline int pw_ldw(int Kp) { return (Kp % 32 == 0) ? Kp + 16 : Kp; }      // rows 16 floats apart modulo the 32 banks: LDS.128 conflict-free
inline size_t pw_smem_bytes(int Kc) { const int Kp = (Kc + 15) / 16 * 16; return (size_t)PW_NSLICE * pw_ldw(Kp) * sizeof(float); }

// C[m][n] = act( sum_z partial[z][m][n] + bias[n] ), splits added in index order (deterministic)
// Optional epilogue operands as in gemm.cuh: + addrow[m / L][n] before the activation, + addpos[m % L][n] and + resid[m][n] after it.
__global__ void pw_reduce_kernel(const float* __restrict__ partial, int nsplits, int M, int N, const float* __restrict__ bias, int act,
                                 const float* __restrict__ act_w, float* __restrict__ C, int ldc,
                                 int L = 1, const float* __restrict__ addrow = nullptr, const float* __restrict__ addpos = nullptr, int ldpos = 0,
                                 const float* __restrict__ resid = nullptr, int ldr = 0) {
    const size_t total = (size_t)M * N;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = i % N; const size_t m = i / N;
        float v = 0.f;
        for (int z = 0; z < nsplits; ++z) v += partial[(size_t)z * total + i];
        if (bias) v += __ldg(bias + n);
        if (addrow) v += __ldg(addrow + (m / L) * N + n);
        v = apply_act(v, act, act_w ? __ldg(act_w + n) : 1.f);
        if (addpos) v += __ldg(addpos + (m % L) * ldpos + n);
        if (resid) v += resid[m * ldr + n];
        C[m * ldc + n] = v;
    }
}

// Returns nullptr on success.  p.ksplit == 0: single pass over K with the fused epilogue.
inline const char* launch_pw_mma(PwParams p, int num_sms, cudaStream_t s) {
    if ((p.lda & 3) || (p.Kc & 3) || (reinterpret_cast<uintptr_t>(p.A) & 15) || (p.kcp & 3)) return "pointwise conv needs 16-byte aligned rows";
    if (p.pass_x && (p.cstride != 2 || p.coff != 1 || (p.N & 1) || (p.chalf & 1) || (p.chp & 1) || (p.ldc & 1) || (p.pass_ld & 1) || p.ksplit > 0))
        return "fused pass-through needs the channel-shuffle store pattern (cstride 2, coff 1, even sizes)";
    if (p.ksplit <= 0) { p.ksplit = (p.Kc + 15) / 16 * 16; p.partial = nullptr; }
    if (p.ksplit & 15) return "K split must be a multiple of 16";
    const int nsplits = (p.Kc + p.ksplit - 1) / p.ksplit;
    p.Kp = (std::min(p.ksplit, p.Kc) + 15) / 16 * 16;
    p.ldw = pw_ldw(p.Kp);
    if ((p.Kc + 15) / 16 * 16 > p.kcp) return "packed weights narrower than the padded K";
    const size_t smem = pw_smem_bytes(std::min(p.ksplit, p.Kc));
    static size_t attr_smem = 0;                           // raised when a launch needs more, not set on every launch
    cudaError_t e = cudaSuccess;
    if (smem > attr_smem) {
        e = cudaFuncSetAttribute(pw_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cudaGetErrorString(e);
        attr_smem = smem;
    }
    const int nslices = (p.N + PW_NSLICE - 1) / PW_NSLICE;
    const int per_sm = smem * 2 + 4096 <= 227 * 1024 ? 2 : 1;
    const int ntiles = (p.M + 15) / 16;
    int gx = std::max(1, num_sms * per_sm / (nslices * nsplits));       // grid = a multiple of the SM count (resident CTAs)
    gx = std::min(gx, (ntiles + PW_THREADS / 32 - 1) / (PW_THREADS / 32));
    pw_mma_kernel<<<dim3(gx, nslices, nsplits), PW_THREADS, smem, s>>>(p);
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace l2s
