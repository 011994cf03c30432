// Weight-stationary batched mat-vec pass used by the persistent recurrent kernels
// (decoder step loop, Bi-LSTM encoder, speaker-encoder LSTM).
//
// One CTA (512 threads = 16 warps) owns R rows of a weight matrix, resident in shared memory as
// Wsm[R][ldw] fp32.  Activations live in global memory (L2) FEATURE-MAJOR: X[k][ldb] with the clip
// index contiguous: 16 lanes x float2 cover 32 clips of one feature with one 128-byte line.  A pass may
// consume several K-segments (e.g. [x ; h_prev]) from different buffers via mv_accumulate.
//
// Lane mapping inside a warp: kq = lane>>4 (2 k-quads), cg = lane&15 (16 pairs of clips).
// Warp w handles the 8-wide k groups {w, w+16, ...}.  Each thread accumulates R x 2 partial dot
// products (<= 32 registers at R=16), which leaves room for TWO groups of x loads in flight per warp
// (ld.global.cg.v2 straight from L2, consumed two iterations later) so L2 latency overlaps the FMAs.
// Partials are reduced over kq by one halving shuffle exchange, over the 16 warps through shared memory,
// and thread (r = tid>>5, b = tid&31) receives the finished value for row r, clip b.
#pragma once
#include "common.cuh"

namespace l2s {

constexpr int MV_THREADS = 512;
constexpr int MV_WARPS = 16;
constexpr int MV_CLIPS = 32;      // clips per pass invocation
constexpr int MV_GW = 8;          // k values per warp-group (2 k-quads x 4)

struct Seg {
    const float* x;   // feature-major [K][ldb], already offset to the first feature of the segment
    int K;            // multiple of 8
};

template <int R>
__device__ __forceinline__ void mv_zero(float (&acc)[R][2]) {
#pragma unroll
    for (int r = 0; r < R; ++r) { acc[r][0] = acc[r][1] = 0.f; }
}

__device__ __forceinline__ float2 ldcg2(const float* p) { return __ldcg(reinterpret_cast<const float2*>(p)); }

template <int R>
__device__ __forceinline__ void mv_group_fma(const float* __restrict__ wp, int ldw, const float2 (&x)[4], float (&acc)[R][2]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float4 w = *reinterpret_cast<const float4*>(wp + r * ldw);
        acc[r][0] = fmaf(w.x, x[0].x, acc[r][0]); acc[r][1] = fmaf(w.x, x[0].y, acc[r][1]);
        acc[r][0] = fmaf(w.y, x[1].x, acc[r][0]); acc[r][1] = fmaf(w.y, x[1].y, acc[r][1]);
        acc[r][0] = fmaf(w.z, x[2].x, acc[r][0]); acc[r][1] = fmaf(w.z, x[2].y, acc[r][1]);
        acc[r][0] = fmaf(w.w, x[3].x, acc[r][0]); acc[r][1] = fmaf(w.w, x[3].y, acc[r][1]);
    }
}

template <int R>
__device__ __forceinline__ void mv_accumulate(const float* __restrict__ Wsm, int ldw, int wcol0,
                                              const float* __restrict__ X, int K, int ldb, int b0,
                                              float (&acc)[R][2]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kq = lane >> 4, cg = lane & 15;
    const int ngroups = K / MV_GW;
    int g = warp;
    if (g >= ngroups) return;
    const size_t gstride = (size_t)MV_WARPS * MV_GW * ldb;          // elements between a warp's consecutive groups
    const float* xp = X + (size_t)(g * MV_GW + kq * 4) * ldb + b0 + cg * 2;
    const float* wp = Wsm + wcol0 + g * MV_GW + kq * 4;
    float2 xa[4], xb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xa[i] = ldcg2(xp + (size_t)i * ldb);
    if (g + MV_WARPS < ngroups) {
#pragma unroll
        for (int i = 0; i < 4; ++i) xb[i] = ldcg2(xp + gstride + (size_t)i * ldb);
    }
    for (; g < ngroups; g += 2 * MV_WARPS) {
        mv_group_fma<R>(wp, ldw, xa, acc);
        if (g + 2 * MV_WARPS < ngroups) {
#pragma unroll
            for (int i = 0; i < 4; ++i) xa[i] = ldcg2(xp + 2 * gstride + (size_t)i * ldb);
        }
        if (g + MV_WARPS < ngroups) {
            mv_group_fma<R>(wp + MV_WARPS * MV_GW, ldw, xb, acc);
            if (g + 3 * MV_WARPS < ngroups) {
#pragma unroll
                for (int i = 0; i < 4; ++i) xb[i] = ldcg2(xp + 3 * gstride + (size_t)i * ldb);
            }
        }
        xp += 2 * gstride;
        wp += 2 * MV_WARPS * MV_GW;
    }
}

// Reduces the per-thread partials.  Returns the finished dot product for (row tid>>5, clip tid&31) in
// threads tid < R*32 (0 elsewhere).
//   RED_ROWS == R : one round, `red` holds MV_WARPS*R*32 floats.
//   RED_ROWS == R/2 : two rounds through a half-size buffer (MV_WARPS*R/2*32 floats).
// Contains __syncthreads(); the caller must __syncthreads() again before `red` is reused.
template <int R, int RED_ROWS>
__device__ __forceinline__ float mv_reduce(float (&acc)[R][2], float* red) {
    static_assert(RED_ROWS == R || RED_ROWS * 2 == R, "RED_ROWS must be R or R/2");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kq = lane >> 4, cg = lane & 15;
    // halving exchange over kq (lane^16): afterwards this lane holds rows [kq*R/2, (kq+1)*R/2) summed over both k-quads
    constexpr int H = R / 2;
    float a[H][2];
    {
        const bool hi = kq != 0;
#pragma unroll
        for (int r = 0; r < H; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float keep = hi ? acc[r + H][c] : acc[r][c];
                float send = hi ? acc[r][c] : acc[r + H][c];
                a[r][c] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
    }
    float v = 0.f;
    if (RED_ROWS == R) {
#pragma unroll
        for (int r = 0; r < H; ++r)
            *reinterpret_cast<float2*>(red + ((size_t)(warp * R + kq * H + r) * MV_CLIPS) + cg * 2) = make_float2(a[r][0], a[r][1]);
        __syncthreads();
        if (threadIdx.x < R * MV_CLIPS) {
            const int r = threadIdx.x >> 5, b = threadIdx.x & 31;
#pragma unroll
            for (int w = 0; w < MV_WARPS; ++w) v += red[(size_t)(w * R + r) * MV_CLIPS + b];
        }
    } else {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            if (half == 1) __syncthreads();
            if (kq == half) {
#pragma unroll
                for (int r = 0; r < H; ++r)
                    *reinterpret_cast<float2*>(red + ((size_t)(warp * RED_ROWS + r) * MV_CLIPS) + cg * 2) = make_float2(a[r][0], a[r][1]);
            }
            __syncthreads();
            const int t = (int)threadIdx.x - half * RED_ROWS * MV_CLIPS;
            if (t >= 0 && t < RED_ROWS * MV_CLIPS) {
                const int r = t >> 5, b = t & 31;
#pragma unroll
                for (int w = 0; w < MV_WARPS; ++w) v += red[(size_t)(w * RED_ROWS + r) * MV_CLIPS + b];
            }
        }
    }
    return v;
}

// ---- narrow variant for B <= 2 (single-clip inference, BASELINE config 1) ---------------------------------------
// With one or two clips the 16 clip-pair lanes of the wide mapping would multiply padding.  Here all 32 lanes split k
// instead (lane = k-quad inside a 128-wide group, warp w takes groups w, w+16, ...), each lane keeps R x 2 partials
// for clips {0,1}, every partial is reduced with warp shuffles and the (at most 16) warp results are summed in smem.
template <int R>
__device__ __forceinline__ void mv_accumulate_narrow(const float* __restrict__ Wsm, int ldw, int wcol0,
                                                     const float* __restrict__ X, int K, int ldb, float (&acc)[R][2]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ngroups = K >> 7;
    for (int g = warp; g < ngroups; g += MV_WARPS) {
        const int kb = g * 128 + lane * 4;
        float2 x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = ldcg2(X + (size_t)(kb + i) * ldb);
        mv_group_fma<R>(Wsm + wcol0 + kb, ldw, x, acc);
    }
}

// Returns the finished value for (row tid>>5, clip tid&31) in threads with (tid&31) < 2 and tid < R*32; 0 elsewhere.
template <int R>
__device__ __forceinline__ float mv_reduce_narrow(float (&acc)[R][2], float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float v = warp_sum(acc[r][c]);
            if (lane == ((2 * r + c) & 31)) red[warp * (2 * R) + 2 * r + c] = v;
        }
    __syncthreads();
    float v = 0.f;
    const int r = threadIdx.x >> 5, b = threadIdx.x & 31;
    if (r < R && b < 2) {
#pragma unroll
        for (int w = 0; w < MV_WARPS; ++w) v += red[w * (2 * R) + 2 * r + b];
    }
    return v;
}

// Convenience: one- or two-segment pass with a full-size reduction buffer (used by lstm.cuh).
template <int R>
__device__ __forceinline__ float mv_pass(const float* __restrict__ Wsm, int ldw,
                                         const Seg& s0, const Seg& s1, int ldb, int b0, float* red) {
    float acc[R][2];
    mv_zero<R>(acc);
    mv_accumulate<R>(Wsm, ldw, 0, s0.x, s0.K, ldb, b0, acc);
    if (s1.K > 0) mv_accumulate<R>(Wsm, ldw, s0.K, s1.x, s1.K, ldb, b0, acc);
    return mv_reduce<R, R>(acc, red);
}

}  // namespace l2s
