// Weight-stationary batched mat-vec pass used by the persistent recurrent kernels
// (decoder step loop, Bi-LSTM encoder, speaker-encoder LSTM).
//
// One CTA (512 threads = 16 warps) owns R rows of a weight matrix, resident in shared memory as
// Wsm[R][K] fp32.  Activations live in global memory (L2) FEATURE-MAJOR: X[k][ldb] with the clip
// index contiguous, so that 8 lanes x float4 cover 32 clips with one 128-byte line.  Up to two
// K-segments (e.g. [x ; h_prev]) may come from different buffers.
//
// Lane mapping inside a warp: kq = lane>>3 (4 k-quads), cg = lane&7 (8 groups of 4 clips).
// Warp w handles the 16-wide k groups {w, w+16, ...}.  Each thread accumulates R x 4 partial dot
// products; partials are reduced over kq by a halving shuffle exchange, over the 16 warps through
// shared memory, and thread (r = tid>>5, b = tid&31) receives the finished value for row r, clip b.
#pragma once
#include "common.cuh"

namespace l2s {

constexpr int MV_THREADS = 512;
constexpr int MV_WARPS = 16;
constexpr int MV_CLIPS = 32;      // clips per pass invocation

struct Seg {
    const float* x;   // feature-major [K][ldb], already offset to the first feature of the segment
    int K;            // multiple of 256
};

template <int R>
__device__ __forceinline__ void mv_accumulate(const float* __restrict__ Wsm, int ldw, int wcol0,
                                              const float* __restrict__ X, int K, int ldb, int b0,
                                              float (&acc)[R][4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kq = lane >> 3, cg = lane & 7;
    const int ngroups = K >> 4;                       // 16-wide k groups
    for (int g = warp; g < ngroups; g += MV_WARPS) {
        const int kb = g * 16 + kq * 4;
        float4 x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = ldcg4(X + (size_t)(kb + i) * ldb + b0 + cg * 4);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 w = *reinterpret_cast<const float4*>(Wsm + r * ldw + wcol0 + kb);
            acc[r][0] = fmaf(w.x, x[0].x, acc[r][0]); acc[r][1] = fmaf(w.x, x[0].y, acc[r][1]);
            acc[r][2] = fmaf(w.x, x[0].z, acc[r][2]); acc[r][3] = fmaf(w.x, x[0].w, acc[r][3]);
            acc[r][0] = fmaf(w.y, x[1].x, acc[r][0]); acc[r][1] = fmaf(w.y, x[1].y, acc[r][1]);
            acc[r][2] = fmaf(w.y, x[1].z, acc[r][2]); acc[r][3] = fmaf(w.y, x[1].w, acc[r][3]);
            acc[r][0] = fmaf(w.z, x[2].x, acc[r][0]); acc[r][1] = fmaf(w.z, x[2].y, acc[r][1]);
            acc[r][2] = fmaf(w.z, x[2].z, acc[r][2]); acc[r][3] = fmaf(w.z, x[2].w, acc[r][3]);
            acc[r][0] = fmaf(w.w, x[3].x, acc[r][0]); acc[r][1] = fmaf(w.w, x[3].y, acc[r][1]);
            acc[r][2] = fmaf(w.w, x[3].z, acc[r][2]); acc[r][3] = fmaf(w.w, x[3].w, acc[r][3]);
        }
    }
}

// Runs one pass.  `red` is a shared scratch of MV_WARPS*R*32 floats.  Returns the finished dot product
// for (row tid>>5, clip b0 + (tid&31)) in threads tid < R*32 (0 elsewhere).  Contains two
// __syncthreads(); the caller must __syncthreads() before `red` is reused by another pass.
template <int R>
__device__ __forceinline__ float mv_pass(const float* __restrict__ Wsm, int ldw,
                                         const Seg& s0, const Seg& s1, int ldb, int b0, float* red) {
    float acc[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
    mv_accumulate<R>(Wsm, ldw, 0, s0.x, s0.K, ldb, b0, acc);
    if (s1.K > 0) mv_accumulate<R>(Wsm, ldw, s0.K, s1.x, s1.K, ldb, b0, acc);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kq = lane >> 3, cg = lane & 7;
    // halving exchange over kq bit1 (lane^16) then bit0 (lane^8): afterwards this lane holds rows
    // [kq*R/4, (kq+1)*R/4) summed over the 4 k-quads.
    constexpr int H1 = R / 2, H2 = R / 4;
    float a1[H1][4];
    {
        const bool hi = (kq & 2) != 0;
#pragma unroll
        for (int r = 0; r < H1; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float keep = hi ? acc[r + H1][c] : acc[r][c];
                float send = hi ? acc[r][c] : acc[r + H1][c];
                a1[r][c] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
    }
    float a2[H2][4];
    {
        const bool hi = (kq & 1) != 0;
#pragma unroll
        for (int r = 0; r < H2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float keep = hi ? a1[r + H2][c] : a1[r][c];
                float send = hi ? a1[r][c] : a1[r + H2][c];
                a2[r][c] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
    }
    // rows held: base = (kq>>1)*H1 + (kq&1)*H2
    const int rbase = (kq >> 1) * H1 + (kq & 1) * H2;
#pragma unroll
    for (int r = 0; r < H2; ++r)
        *reinterpret_cast<float4*>(red + ((size_t)(warp * R + rbase + r) * MV_CLIPS) + cg * 4) =
            make_float4(a2[r][0], a2[r][1], a2[r][2], a2[r][3]);
    __syncthreads();
    float v = 0.f;
    if (threadIdx.x < R * MV_CLIPS) {
        const int r = threadIdx.x >> 5, b = threadIdx.x & 31;
#pragma unroll
        for (int w = 0; w < MV_WARPS; ++w) v += red[(size_t)(w * R + r) * MV_CLIPS + b];
    }
    return v;
}

}  // namespace l2s
