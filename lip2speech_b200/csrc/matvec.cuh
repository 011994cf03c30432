// Weight-stationary batched mat-vec passes used by the persistent recurrent kernels: the FMA pass (mv_*, row-partitioned
// decode kernel), the 32-clip tensor-core pass (mv32_*, Bi-LSTM encoder and speaker-encoder LSTM) and the 8-clip
// tensor-core pass (mv8_*, stage-pipelined decode kernel).
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

// ---- tensor-core variant (3xTF32 on mma.sync.m16n8k8) -----------------------------------------------------------
// The FMA pass above is bound by shared-memory bandwidth: every LDS.128 of weights feeds only 8 FMAs per lane.  Here a
// warp owns a 16-row x 16-k weight tile per iteration and multiplies it with the [16 k][32 clips] activation tile on
// the tensor cores; each weight is read from shared memory ONCE per warp and reused for all 32 clips.  fp32-grade
// accuracy comes from the error-compensated split a = a_hi + a_lo (hi = top 19 bits, lo = exact remainder):
// D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi, split in registers (resident weights stay plain fp32, 4 B each).
// tcgen05 is not usable here: it takes operands from shared memory only (no room for pre-split weights: 2 x 25 MB
// > 148 x 227 KB) and needs M >= 64 rows per CTA while a CTA owns ~14 rows of each layer.
//
// (A 32-clip form of this pass — four n-tiles per weight tile on the row-partitioned layout — lives in
// tools/mvt_bench.cu: it measured 4.65 us against the FMA pass's 6.5 us and showed the L2 broadcast bound.)
constexpr int MV_KC = 16;         // k values per warp iteration of the tensor-core pass

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- 32-clip tensor-core pass on the feature-major layout (lstm.cuh) ----------------------------------------------
// RT 16-row weight tiles x four n-tiles of 8 clips.  MMA column n of n-tile j <-> clip 4n + j, so lane (g, t) reads
// X[k][4g..4g+3] as one LDG.128 per k (full 128-byte lines per warp) and ends up holding 8 consecutive clips of rows g and
// g+8; MMA column t / t+4 of k8-step s <-> real k = k0 + 4t + 2s / + 2s + 1, weights W[g][k0+4t..+3] as LDS.128.
// K % 256 == 0 and K <= 512 per segment: a warp owns K/256 chunks and requests them all before its first MMA.
template <int RT>
__device__ __forceinline__ void mv32_zero(float (&acc)[RT][4][4]) {
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[r][j][i] = 0.f;
}

template <int RT>
__device__ __forceinline__ void mv32_accumulate(const float* __restrict__ W, int ldw, int wcol0, int R,
                                                const float* __restrict__ X, int K, int ldb, int b0, float (&acc)[RT][4][4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int npw = K / (MV_KC * MV_WARPS);                  // 1 or 2 chunks per warp
    float4 x[2][4];
#pragma unroll
    for (int d = 0; d < 2; ++d)
        if (d < npw) {
            const float* xp = X + (size_t)((warp + d * MV_WARPS) * MV_KC + 4 * t) * ldb + b0 + 4 * g;
#pragma unroll
            for (int i = 0; i < 4; ++i) x[d][i] = ldcg4(xp + (size_t)i * ldb);
        }
#pragma unroll
    for (int d = 0; d < 2; ++d)
        if (d < npw) {
            const int k0 = wcol0 + (warp + d * MV_WARPS) * MV_KC + 4 * t;
            uint32_t xh[4][4], xl[4][4];                      // [k index i][clip j]
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                split_tf32(x[d][i].x, xh[i][0], xl[i][0]); split_tf32(x[d][i].y, xh[i][1], xl[i][1]);
                split_tf32(x[d][i].z, xh[i][2], xl[i][2]); split_tf32(x[d][i].w, xh[i][3], xl[i][3]);
            }
#pragma unroll
            for (int rt = 0; rt < RT; ++rt) {
                const float4 wa = *reinterpret_cast<const float4*>(W + (size_t)min(rt * 16 + g, R - 1) * ldw + k0);
                const float4 wb = *reinterpret_cast<const float4*>(W + (size_t)min(rt * 16 + g + 8, R - 1) * ldw + k0);
                const float wav[4] = {wa.x, wa.y, wa.z, wa.w}, wbv[4] = {wb.x, wb.y, wb.z, wb.w};
                uint32_t ah[4], al[4], bh[4], bl[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { split_tf32(wav[i], ah[i], al[i]); split_tf32(wbv[i], bh[i], bl[i]); }
#pragma unroll
                for (int s = 0; s < 2; ++s)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        mma_tf32(acc[rt][j], al[2 * s], bl[2 * s], al[2 * s + 1], bl[2 * s + 1], xh[2 * s][j], xh[2 * s + 1][j]);
                        mma_tf32(acc[rt][j], ah[2 * s], bh[2 * s], ah[2 * s + 1], bh[2 * s + 1], xl[2 * s][j], xl[2 * s + 1][j]);
                        mma_tf32(acc[rt][j], ah[2 * s], bh[2 * s], ah[2 * s + 1], bh[2 * s + 1], xh[2 * s][j], xh[2 * s + 1][j]);
                    }
            }
        }
}

// Cross-warp reduction of one row tile through `red` (MV_WARPS x 16 x 32 floats).  Thread tid receives the finished value
// for (row tid>>5 of the tile, clip tid&31).  Contains one __syncthreads(); the caller must __syncthreads() again before
// `red` is reused.
__device__ __forceinline__ float mv32_reduce_tile(const float (&a)[4][4], float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    float* base = red + (size_t)warp * 16 * MV_CLIPS;
    *reinterpret_cast<float4*>(base + g * MV_CLIPS + 8 * t) = make_float4(a[0][0], a[1][0], a[2][0], a[3][0]);
    *reinterpret_cast<float4*>(base + g * MV_CLIPS + 8 * t + 4) = make_float4(a[0][1], a[1][1], a[2][1], a[3][1]);
    *reinterpret_cast<float4*>(base + (g + 8) * MV_CLIPS + 8 * t) = make_float4(a[0][2], a[1][2], a[2][2], a[3][2]);
    *reinterpret_cast<float4*>(base + (g + 8) * MV_CLIPS + 8 * t + 4) = make_float4(a[0][3], a[1][3], a[2][3], a[3][3]);
    __syncthreads();
    const int r = threadIdx.x >> 5, b = threadIdx.x & 31;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < MV_WARPS; ++w) v += red[(size_t)(w * 16 + r) * MV_CLIPS + b];
    return v;
}

// ---- 8-clip variant for the stage-pipelined decode kernel (decode3.cuh) ------------------------------------------
// One n-tile (8 clips) and RT 16-row tiles per pass, so the activation fragment of a chunk (4 floats per lane) is reused
// by every row tile.  Activations are GROUP-MAJOR here: X[k][8] for one clip group, so the 16 k x 8 clips of a chunk are
// 512 contiguous bytes and each of the four LDG.32 of a warp reads exactly one 128-byte line (the feature-major
// [k][32] layout costs four line requests per LDG: the L1 wavefront queue, not the tensor pipe, set the pace).
// Lane (g = lane>>2, t = lane&3) owns clip g and the k values k0 + 4i + t, i = 0..3; the weight columns of every
// 16-chunk are permuted at pack time (position 4t + i <- column 4i + t) so its four weights are one LDS.128.
// K is a multiple of 256: every warp owns K/256 <= MV8_DEPTH chunks per segment and requests all of them from L2 before
// its first MMA (one exposed L2 round trip per segment).  RT is a compile-time constant: the body is branch-free, so
// the RT independent accumulator chains interleave and the tensor pipe (8.6 clk per m16n8k8 per scheduler, measured)
// stays busy.
constexpr int MV8_DEPTH = 4;

template <int RT>
__device__ __forceinline__ void mv8_zero(float (&acc)[RT][4]) {
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[r][i] = 0.f;
}

// Requests the chunks of one segment that this warp owns: x[d][i] = X[(16*(warp + 16 d) + 4 i + t)*8 + g], d < K/256.
// X: feature 0 of the segment for this clip group (X[k*8 + clip]).
template <int DEPTH>
__device__ __forceinline__ void mv8_load(const float* __restrict__ X, int K, float (&x)[DEPTH][4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int npw = K / (MV_KC * MV_WARPS);                  // chunks per warp
    const float* xp = X + (size_t)(warp * MV_KC + t) * 8 + g;
    constexpr int xstride = MV_WARPS * MV_KC * 8;
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
        if (d < npw) {
#pragma unroll
            for (int i = 0; i < 4; ++i) x[d][i] = __ldcg(xp + d * xstride + i * 32);
        }
}

// Multiplies the loaded chunks with the weight rows of the pass.  W: first row of the pass in shared memory; R real rows
// (rows >= R re-read row R-1; their results are never used).
template <int RT, int DEPTH>
__device__ __forceinline__ void mv8_mma(const float* __restrict__ W, int ldw, int wcol0, int R, int K,
                                        const float (&x)[DEPTH][4], float (&acc)[RT][4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int npw = K / (MV_KC * MV_WARPS);
    const float* wrow[RT][2];
#pragma unroll
    for (int rt = 0; rt < RT; ++rt) {
        wrow[rt][0] = W + (size_t)min(rt * 16 + g, R - 1) * ldw + wcol0 + warp * MV_KC + 4 * t;
        wrow[rt][1] = W + (size_t)min(rt * 16 + g + 8, R - 1) * ldw + wcol0 + warp * MV_KC + 4 * t;
    }
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
        if (d < npw) {
            uint32_t xh[4], xl[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) split_tf32(x[d][i], xh[i], xl[i]);
#pragma unroll
            for (int rt = 0; rt < RT; ++rt) {
                const float4 wa = *reinterpret_cast<const float4*>(wrow[rt][0] + d * MV_WARPS * MV_KC);
                const float4 wb = *reinterpret_cast<const float4*>(wrow[rt][1] + d * MV_WARPS * MV_KC);
                const float wav[4] = {wa.x, wa.y, wa.z, wa.w}, wbv[4] = {wb.x, wb.y, wb.z, wb.w};
                uint32_t ah[4], al[4], bh[4], bl[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { split_tf32(wav[i], ah[i], al[i]); split_tf32(wbv[i], bh[i], bl[i]); }
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    mma_tf32(acc[rt], al[2 * s], bl[2 * s], al[2 * s + 1], bl[2 * s + 1], xh[2 * s], xh[2 * s + 1]);
                    mma_tf32(acc[rt], ah[2 * s], bh[2 * s], ah[2 * s + 1], bh[2 * s + 1], xl[2 * s], xl[2 * s + 1]);
                    mma_tf32(acc[rt], ah[2 * s], bh[2 * s], ah[2 * s + 1], bh[2 * s + 1], xh[2 * s], xh[2 * s + 1]);
                }
            }
        }
}

template <int RT>
__device__ __forceinline__ void mv8_accumulate(const float* __restrict__ W, int ldw, int wcol0, int R,
                                               const float* __restrict__ X, int K, float (&acc)[RT][4]) {
    float x[MV8_DEPTH][4];
    mv8_load<MV8_DEPTH>(X, K, x);
    mv8_mma<RT, MV8_DEPTH>(W, ldw, wcol0, R, K, x, acc);
}

// Cross-warp reduction of row tiles 3*round .. 3*round+2 through `red` (MV_WARPS x 3 x 128 floats = 24 KB).  Threads
// tid < 384 receive the finished value for (row 48*round + (tid>>3), clip tid&7).  Contains one __syncthreads(); the
// caller must __syncthreads() again before `red` is reused.
constexpr int MV8_RTILES = 3;     // row tiles per reduction round
template <int RT>
__device__ __forceinline__ float mv8_reduce_round(const float (&acc)[RT][4], int round, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int rt = 0; rt < RT; ++rt)
        if ((rt / MV8_RTILES) == round) {
            float* base = red + (size_t)(warp * MV8_RTILES + (rt % MV8_RTILES)) * 128;
            *reinterpret_cast<float2*>(base + g * 8 + 2 * t) = make_float2(acc[rt][0], acc[rt][1]);
            *reinterpret_cast<float2*>(base + (g + 8) * 8 + 2 * t) = make_float2(acc[rt][2], acc[rt][3]);
        }
    __syncthreads();
    float v = 0.f;
    if (threadIdx.x < MV8_RTILES * 128) {
        const int j = threadIdx.x >> 7, e = threadIdx.x & 127;
        float part[MV_WARPS];
#pragma unroll
        for (int w = 0; w < MV_WARPS; ++w) part[w] = red[(size_t)(w * MV8_RTILES + j) * 128 + e];
#pragma unroll
        for (int w = 0; w < MV_WARPS; ++w) v += part[w];
    }
    return v;
}

}  // namespace l2s
