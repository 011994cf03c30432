// Micro-benchmark + correctness check of the tensor-core (3xTF32 mma.sync) weight-stationary pass against the FMA pass
// (csrc/matvec.cuh).  148 CTAs, 16 weight rows per CTA resident in shared memory, activations [K][32] in L2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lip2speech_b200/csrc tools/mvt_bench.cu -o /tmp/mvt_bench && /tmp/mvt_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "matvec.cuh"

using namespace l2s;

// ---- 32-clip tensor-core pass on the row-partitioned layout (bench only; the product uses the 8-clip mv8_* form) ----
// Fragment mapping (g = lane>>2, t = lane&3): MMA column t / t+4 of k8-step s <-> real k = k0 + 4t + 2s / + 2s + 1, so a lane
// reads W[g][k0+4t..+3] and W[g+8][k0+4t..+3] as two LDS.128; MMA column n of n-tile j <-> clip 4n + j, so a lane reads
// X[k][4g..4g+3] as one LDG.128 per k and ends up holding 8 consecutive clips of rows g and g+8.
namespace l2s {
__device__ __forceinline__ void mvt_zero(float (&acc)[4][4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
}

// One 16-k chunk: weights wa (row g) / wb (row g+8), activations x[i] = X[k0+4t+i][4g..4g+3].
__device__ __forceinline__ void mvt_chunk(const float4& wa, const float4& wb, const float4 (&x)[4], float (&acc)[4][4]) {
    const float wav[4] = {wa.x, wa.y, wa.z, wa.w}, wbv[4] = {wb.x, wb.y, wb.z, wb.w};
    uint32_t ah[4], al[4], bh[4], bl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { split_tf32(wav[i], ah[i], al[i]); split_tf32(wbv[i], bh[i], bl[i]); }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const float xs0[4] = {x[2 * s].x, x[2 * s].y, x[2 * s].z, x[2 * s].w};
        const float xs1[4] = {x[2 * s + 1].x, x[2 * s + 1].y, x[2 * s + 1].z, x[2 * s + 1].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t h0, l0, h1, l1;
            split_tf32(xs0[j], h0, l0);
            split_tf32(xs1[j], h1, l1);
            mma_tf32(acc[j], al[2 * s], bl[2 * s], al[2 * s + 1], bl[2 * s + 1], h0, h1);
            mma_tf32(acc[j], ah[2 * s], bh[2 * s], ah[2 * s + 1], bh[2 * s + 1], l0, l1);
            mma_tf32(acc[j], ah[2 * s], bh[2 * s], ah[2 * s + 1], bh[2 * s + 1], h0, h1);
        }
    }
}

// Accumulates rows [0,16) x clips [b0, b0+32) over one K segment (K % 16 == 0).  Rows >= R re-read row R-1 (their
// results are never used).  Warp w takes chunks w, w+16, ...; the activations of the next chunk are in flight while
// the current one is multiplied.
template <int R>
__device__ __forceinline__ void mvt_accumulate(const float* __restrict__ Wsm, int ldw, int wcol0,
                                               const float* __restrict__ X, int K, int ldb, int b0,
                                               float (&acc)[4][4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nchunks = K / MV_KC;
    int c = warp;
    if (c >= nchunks) return;
    const int ra = g < R ? g : R - 1, rb = g + 8 < R ? g + 8 : R - 1;
    const float* wpa = Wsm + (size_t)ra * ldw + wcol0 + c * MV_KC + 4 * t;
    const float* wpb = Wsm + (size_t)rb * ldw + wcol0 + c * MV_KC + 4 * t;
    const float* xp = X + (size_t)(c * MV_KC + 4 * t) * ldb + b0 + 4 * g;
    const size_t cstride = (size_t)MV_WARPS * MV_KC * ldb;
    float4 xa[4], xb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xa[i] = ldcg4(xp + (size_t)i * ldb);
    for (; c < nchunks; c += 2 * MV_WARPS) {
        const bool more1 = c + MV_WARPS < nchunks;
        if (more1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) xb[i] = ldcg4(xp + cstride + (size_t)i * ldb);
        }
        mvt_chunk(*reinterpret_cast<const float4*>(wpa), *reinterpret_cast<const float4*>(wpb), xa, acc);
        if (more1) {
            if (c + 2 * MV_WARPS < nchunks) {
#pragma unroll
                for (int i = 0; i < 4; ++i) xa[i] = ldcg4(xp + 2 * cstride + (size_t)i * ldb);
            }
            mvt_chunk(*reinterpret_cast<const float4*>(wpa + MV_WARPS * MV_KC), *reinterpret_cast<const float4*>(wpb + MV_WARPS * MV_KC), xb, acc);
        }
        xp += 2 * cstride;
        wpa += 2 * MV_WARPS * MV_KC;
        wpb += 2 * MV_WARPS * MV_KC;
    }
}

// Cross-warp reduction of the tensor-core partial tiles.  `red` holds MV_WARPS x RED_ROWS x 32 floats (RED_ROWS = 16:
// one round; RED_ROWS = 8: rows [0,8) then rows [8,16) through the same buffer).  Returns the finished value for
// (row tid>>5, clip tid&31) in threads tid < 16*32.  Contains __syncthreads(); the caller must __syncthreads() again
// before `red` is reused.
template <int RED_ROWS>
__device__ __forceinline__ float mvt_reduce(const float (&acc)[4][4], float* red) {
    static_assert(RED_ROWS == 16 || RED_ROWS == 8, "RED_ROWS must be 16 or 8");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    float v = 0.f;
    if (RED_ROWS == 16) {
        float* base = red + (size_t)warp * 16 * MV_CLIPS;
        *reinterpret_cast<float4*>(base + g * MV_CLIPS + 8 * t) = make_float4(acc[0][0], acc[1][0], acc[2][0], acc[3][0]);
        *reinterpret_cast<float4*>(base + g * MV_CLIPS + 8 * t + 4) = make_float4(acc[0][1], acc[1][1], acc[2][1], acc[3][1]);
        *reinterpret_cast<float4*>(base + (g + 8) * MV_CLIPS + 8 * t) = make_float4(acc[0][2], acc[1][2], acc[2][2], acc[3][2]);
        *reinterpret_cast<float4*>(base + (g + 8) * MV_CLIPS + 8 * t + 4) = make_float4(acc[0][3], acc[1][3], acc[2][3], acc[3][3]);
        __syncthreads();
        const int r = threadIdx.x >> 5, b = threadIdx.x & 31;
#pragma unroll
        for (int w = 0; w < MV_WARPS; ++w) v += red[(size_t)(w * 16 + r) * MV_CLIPS + b];
    } else {
        float* base = red + (size_t)warp * 8 * MV_CLIPS;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            if (half == 1) __syncthreads();
            *reinterpret_cast<float4*>(base + g * MV_CLIPS + 8 * t) =
                make_float4(acc[0][2 * half], acc[1][2 * half], acc[2][2 * half], acc[3][2 * half]);
            *reinterpret_cast<float4*>(base + g * MV_CLIPS + 8 * t + 4) =
                make_float4(acc[0][2 * half + 1], acc[1][2 * half + 1], acc[2][2 * half + 1], acc[3][2 * half + 1]);
            __syncthreads();
            const int tt = (int)threadIdx.x - half * 8 * MV_CLIPS;
            if (tt >= 0 && tt < 8 * MV_CLIPS) {
                const int r = tt >> 5, b = tt & 31;
#pragma unroll
                for (int w = 0; w < MV_WARPS; ++w) v += red[(size_t)(w * 8 + r) * MV_CLIPS + b];
            }
        }
    }
    return v;
}

}  // namespace l2s


// ---- experimental chunk variants (bench only) -------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// (hi, mid) bf16x2 split of the pair (a -> low half, b -> high half)
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& mid) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(rb), "f"(ra));
}
template <int MODE>
__device__ __forceinline__ void chunk_x(const float4& wa, const float4& wb, const float4 (&x)[4], float (&acc)[4][4], float (&acc2)[4][4]) {
    if (MODE == 3) {   // bf16x2 split, m16n8k16
        uint32_t a0h, a0m, a1h, a1m, a2h, a2m, a3h, a3m;
        split_bf16x2(wa.x, wa.y, a0h, a0m); split_bf16x2(wb.x, wb.y, a1h, a1m);
        split_bf16x2(wa.z, wa.w, a2h, a2m); split_bf16x2(wb.z, wb.w, a3h, a3m);
        const float x0[4] = {x[0].x, x[0].y, x[0].z, x[0].w}, x1[4] = {x[1].x, x[1].y, x[1].z, x[1].w};
        const float x2[4] = {x[2].x, x[2].y, x[2].z, x[2].w}, x3[4] = {x[3].x, x[3].y, x[3].z, x[3].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t b0h, b0m, b1h, b1m;
            split_bf16x2(x0[j], x1[j], b0h, b0m); split_bf16x2(x2[j], x3[j], b1h, b1m);
            mma_bf16(acc2[j], a0m, a1m, a2m, a3m, b0h, b1h);
            mma_bf16(acc2[j], a0h, a1h, a2h, a3h, b0m, b1m);
            mma_bf16(acc[j], a0h, a1h, a2h, a3h, b0h, b1h);
        }
        return;
    }
    const float wav[4] = {wa.x, wa.y, wa.z, wa.w}, wbv[4] = {wb.x, wb.y, wb.z, wb.w};
    uint32_t ah[4], al[4], bh[4], bl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { split_tf32(wav[i], ah[i], al[i]); split_tf32(wbv[i], bh[i], bl[i]); }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const float xs0[4] = {x[2 * s].x, x[2 * s].y, x[2 * s].z, x[2 * s].w};
        const float xs1[4] = {x[2 * s + 1].x, x[2 * s + 1].y, x[2 * s + 1].z, x[2 * s + 1].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t h0, l0, h1, l1;
            split_tf32(xs0[j], h0, l0);
            split_tf32(xs1[j], h1, l1);
            if (MODE == 1) {
                mma_tf32(acc2[j], al[2 * s], bl[2 * s], al[2 * s + 1], bl[2 * s + 1], h0, h1);
                mma_tf32(acc2[j], ah[2 * s], bh[2 * s], ah[2 * s + 1], bh[2 * s + 1], l0, l1);
            }
            mma_tf32(acc[j], ah[2 * s], bh[2 * s], ah[2 * s + 1], bh[2 * s + 1], h0, h1);
        }
    }
}
__device__ int g_rot;
template <int MODE>
__device__ __forceinline__ void acc_x(const float* __restrict__ Wsm, int ldw, const float* __restrict__ X, int K, int ldb, float (&acc)[4][4], float (&acc2)[4][4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nchunks = K / 16;
    const float* wpa = Wsm + (size_t)g * ldw + 4 * t;
    const float* wpb = Wsm + (size_t)(g + 8) * ldw + 4 * t;
    float4 xa[4], xb[4];
    int c = warp;
    if (c >= nchunks) return;
    const int rot = g_rot ? (int)((blockIdx.x * 37u) % (unsigned)nchunks) : 0;
#define RC(cc) (((cc) + rot) % nchunks)
#pragma unroll
    for (int i = 0; i < 4; ++i) xa[i] = ldcg4(X + (size_t)(RC(c) * 16 + 4 * t + i) * ldb + 4 * g);
    for (; c < nchunks; c += 32) {
        if (c + 16 < nchunks) {
#pragma unroll
            for (int i = 0; i < 4; ++i) xb[i] = ldcg4(X + (size_t)(RC(c + 16) * 16 + 4 * t + i) * ldb + 4 * g);
        }
        chunk_x<MODE>(*reinterpret_cast<const float4*>(wpa + RC(c) * 16), *reinterpret_cast<const float4*>(wpb + RC(c) * 16), xa, acc, acc2);
        if (c + 16 < nchunks) {
            if (c + 32 < nchunks) {
#pragma unroll
                for (int i = 0; i < 4; ++i) xa[i] = ldcg4(X + (size_t)(RC(c + 32) * 16 + 4 * t + i) * ldb + 4 * g);
            }
            chunk_x<MODE>(*reinterpret_cast<const float4*>(wpa + RC(c + 16) * 16), *reinterpret_cast<const float4*>(wpb + RC(c + 16) * 16), xb, acc, acc2);
        }
    }
#undef RC
}

template <int VARIANT, int R>
__global__ void __launch_bounds__(512, 1) bench_kernel(const float* __restrict__ Wg, const float* __restrict__ X, float* __restrict__ out,
                                                        int K, int ldw, int iters) {
    extern __shared__ __align__(16) float smem[];
    float* wsm = smem;                       // [R][ldw]
    float* red = wsm + R * ldw;              // [16][R][32]
    for (int i = threadIdx.x; i < R * K; i += 512) wsm[(i / K) * ldw + (i % K)] = Wg[(size_t)blockIdx.x * R * K + i];
    __syncthreads();
    float sink = 0.f;
    for (int it = 0; it < iters; ++it) {
        const float* x = X + (size_t)(it & 3) * K * 32;
        float v;
        if (VARIANT == 0) {
            float acc[R][2]; mv_zero<R>(acc);
            mv_accumulate<R>(wsm, ldw, 0, x, K, 32, 0, acc);
            v = mv_reduce<R, R>(acc, red);
        } else if (VARIANT == 1) {
            float acc[4][4]; mvt_zero(acc);
            mvt_accumulate<R>(wsm, ldw, 0, x, K, 32, 0, acc);
            v = mvt_reduce<16>(acc, red);
        } else {
            float acc[4][4], acc2[4][4]; mvt_zero(acc); mvt_zero(acc2);
            acc_x<VARIANT - 10>(wsm, ldw, x, K, 32, acc, acc2);
            for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) acc[j][i] += acc2[j][i];
            v = mvt_reduce<16>(acc, red);
        }
        sink += v;
        __syncthreads();
    }
    out[(size_t)blockIdx.x * 512 + threadIdx.x] = sink;
}

template <int VARIANT, int R>
void run(const char* name, int K, int ldw, int iters, const float* W, const float* X, float* out, const std::vector<float>& hW, const std::vector<float>& hX) {
    size_t smem = (size_t)(R * ldw + MV_WARPS * 16 * 32) * sizeof(float);
    cudaFuncSetAttribute(bench_kernel<VARIANT, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench_kernel<VARIANT, R><<<148, 512, smem>>>(W, X, out, K, ldw, 1);
    std::vector<float> h(148 * 512);
    cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int cta = 0; cta < 148; cta += 49)
        for (int r = 0; r < R; ++r)
            for (int b = 0; b < 32; ++b) {
                double s = 0;
                for (int k = 0; k < K; ++k) s += (double)hW[((size_t)cta * R + r) * K + k] * hX[(size_t)k * 32 + b];
                maxerr = fmax(maxerr, fabs(s - h[cta * 512 + r * 32 + b]));
                maxref = fmax(maxref, fabs(s));
            }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench_kernel<VARIANT, R><<<148, 512, smem>>>(W, X, out, K, ldw, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    printf("%-26s R=%2d K=%4d ldw=%4d : %7.3f us/pass   rel err %.2e  %s\n", name, R, K, ldw, ms * 1e3 / iters, maxerr / maxref,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const int KMAX = 1536;
    std::vector<float> hW((size_t)148 * 16 * KMAX), hX((size_t)4 * KMAX * 32);
    srand(1);
    for (auto& v : hW) v = (rand() / (float)RAND_MAX - 0.5f) * 0.1f;
    for (auto& v : hX) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    float *W, *X, *out;
    cudaMalloc(&W, hW.size() * 4); cudaMalloc(&X, hX.size() * 4); cudaMalloc(&out, 148 * 512 * 4);
    const int iters = 2000;
    for (int K : {256, 512, 1024, 1536}) {
        // the activations of pass `it` are X + (it&3)*K*32: the check uses it = 0
        std::vector<float> w2((size_t)148 * 16 * K);
        for (size_t i = 0; i < w2.size(); ++i) w2[i] = hW[i];
        cudaMemcpy(W, w2.data(), w2.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(X, hX.data(), hX.size() * 4, cudaMemcpyHostToDevice);
        run<0, 16>("fma  2-clip lanes", K, K, iters, W, X, out, w2, hX);
        run<1, 16>("mma  3xTF32", K, K + 16, iters, W, X, out, w2, hX);
        run<1, 16>("mma  3xTF32 (ldw=K)", K, K, iters, W, X, out, w2, hX);
        { int one = 1, zero = 0; cudaMemcpyToSymbol(g_rot, &one, 4);
          run<10, 16>("mma  3xTF32 rotated k", K, K + 16, iters, W, X, out, w2, hX);
          run<12, 16>("mma  1xTF32 rotated k", K, K + 16, iters, W, X, out, w2, hX);
          cudaMemcpyToSymbol(g_rot, &zero, 4); }
        run<11, 16>("mma  3xTF32 2 acc sets", K, K + 16, iters, W, X, out, w2, hX);
        run<12, 16>("mma  1xTF32 (rate probe)", K, K + 16, iters, W, X, out, w2, hX);
        run<13, 16>("mma  bf16x2 split k16", K, K + 16, iters, W, X, out, w2, hX);
    }
    return 0;
}
