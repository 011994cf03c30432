// Micro-benchmark of the weight-stationary mat-vec pass (csrc/matvec.cuh) in isolation: 148 CTAs, weights in
// shared memory, activations [K][32] in L2, `iters` passes back to back.  Build & run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lip2speech_b200/csrc tools/mv_bench.cu -o /tmp/mv_bench && /tmp/mv_bench
#include <cstdio>
#include <vector>
#include "matvec.cuh"

using namespace l2s;

// ---- variant B: 4 clips per lane (kq=4 x cg=8), LDS.128 weights, x double-buffered one group ahead -------------
template <int R>
__device__ __forceinline__ void acc4(const float* __restrict__ Wsm, int ldw, const float* __restrict__ X, int K, int ldb, float (&acc)[R][4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kq = lane >> 3, cg = lane & 7;
    const int ngroups = K >> 4;
    for (int g = warp; g < ngroups; g += MV_WARPS) {
        const int kb = g * 16 + kq * 4;
        float4 x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = ldcg4(X + (size_t)(kb + i) * ldb + cg * 4);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 w = *reinterpret_cast<const float4*>(Wsm + r * ldw + kb);
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

// ---- variant C: x staged in shared memory by cp.async (whole segment), 4 clips per lane, LDS for both operands ----
template <int R>
__device__ __forceinline__ void acc4_smemx(const float* __restrict__ Wsm, int ldw, const float* __restrict__ Xs /*smem [K][32]*/, int K, float (&acc)[R][4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kq = lane >> 3, cg = lane & 7;
    const int ngroups = K >> 4;
    for (int g = warp; g < ngroups; g += MV_WARPS) {
        const int kb = g * 16 + kq * 4;
        float4 x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const float4*>(Xs + (kb + i) * 32 + cg * 4);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 w = *reinterpret_cast<const float4*>(Wsm + r * ldw + kb);
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

template <int VARIANT, int R>
__global__ void __launch_bounds__(512, 1) bench_kernel(const float* __restrict__ Wg, const float* __restrict__ X, float* __restrict__ out,
                                                        int K, int iters, int do_reduce) {
    extern __shared__ __align__(16) float smem[];
    float* wsm = smem;                       // [R][K]
    float* red = wsm + R * K;                // [16][R][32]
    float* xs = red + MV_WARPS * R * 32;     // [K][32] (variant C only)
    for (int i = threadIdx.x; i < R * K; i += 512) wsm[i] = Wg[(size_t)blockIdx.x * R * K + i];
    __syncthreads();
    float sink = 0.f;
    for (int it = 0; it < iters; ++it) {
        const float* x = X + (size_t)(it & 3) * K * 32;       // rotate between 4 activation buffers (all L2-resident)
        if (VARIANT == 0) {
            float acc[R][2]; mv_zero<R>(acc);
            mv_accumulate<R>(wsm, K, 0, x, K, 32, 0, acc);
            if (do_reduce) { float v = mv_reduce<R, R>(acc, red); sink += v; __syncthreads(); }
            else { for (int r = 0; r < R; ++r) sink += acc[r][0] + acc[r][1]; }
        } else if (VARIANT == 1) {
            float acc[R][4];
            for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
            acc4<R>(wsm, K, x, K, 32, acc);
            for (int r = 0; r < R; ++r) sink += acc[r][0] + acc[r][1] + acc[r][2] + acc[r][3];
            if (do_reduce) __syncthreads();
        } else {
            // stage x with cp.async (16 B per thread per step), then compute from smem
            for (int i = threadIdx.x; i < K * 8; i += 512)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(xs + i * 4)), "l"(x + i * 4) : "memory");
            asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            float acc[R][4];
            for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
            acc4_smemx<R>(wsm, K, xs, K, acc);
            for (int r = 0; r < R; ++r) sink += acc[r][0] + acc[r][1] + acc[r][2] + acc[r][3];
            __syncthreads();
        }
    }
    out[blockIdx.x * 512 + threadIdx.x] = sink;
}

template <int VARIANT, int R>
void run(const char* name, int K, int iters, int do_reduce, const float* W, const float* X, float* out) {
    size_t smem = (size_t)(R * K + MV_WARPS * R * 32 + (VARIANT == 2 ? K * 32 : 0)) * sizeof(float);
    cudaFuncSetAttribute(bench_kernel<VARIANT, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench_kernel<VARIANT, R><<<148, 512, smem>>>(W, X, out, K, 10, do_reduce);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench_kernel<VARIANT, R><<<148, 512, smem>>>(W, X, out, K, iters, do_reduce);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    double fma_us = (double)R * K * 32 / 128.0 / 1965.0;      // FMA-issue bound per pass (128 FMA/clk/SM @1.965 GHz)
    printf("%-34s R=%2d K=%4d reduce=%d : %7.3f us/pass  (FMA bound %.3f us)  %s\n", name, R, K, do_reduce, ms * 1e3 / iters, fma_us,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

// raw FFMA issue rate: 32 independent accumulators per thread, operands in registers
__global__ void __launch_bounds__(512, 1) fma_kernel(float* out, int iters, float a0, float b0) {
    float acc[32];
    float a[4] = {a0, a0 + 1.f, a0 + 2.f, a0 + 3.f}, b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = b0 + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = (float)i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = fmaf(a[i & 3], b[i >> 2], acc[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    out[blockIdx.x * 512 + threadIdx.x] = s;
}

int main() {
    {
        float* o; cudaMalloc(&o, 148 * 512 * 4);
        fma_kernel<<<148, 512>>>(o, 1000, 1.f, 2.f);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int iters = 20000;
        cudaEventRecord(e0);
        fma_kernel<<<148, 512>>>(o, iters, 1.f, 2.f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma_per_sm = 512.0 * 32 * iters;
        printf("raw FFMA: %.1f FMA/clk/SM (assuming 1.965 GHz), %.2f TFLOP/s chip\n", fma_per_sm / (ms * 1e-3 * 1.965e9), 2 * fma_per_sm * 148 / (ms * 1e-3) / 1e12);
    }
    const int KMAX = 1536;
    float *W, *X, *out;
    cudaMalloc(&W, (size_t)148 * 16 * KMAX * 4); cudaMalloc(&X, (size_t)4 * KMAX * 32 * 4); cudaMalloc(&out, 148 * 512 * 4);
    cudaMemset(W, 0, (size_t)148 * 16 * KMAX * 4); cudaMemset(X, 0, (size_t)4 * KMAX * 32 * 4);
    const int iters = 2000;
    for (int K : {512, 1024}) {
        run<0, 16>("v0 2-clip lanes, depth-2 prefetch", K, iters, 0, W, X, out);
        run<0, 16>("v0 2-clip lanes, depth-2 prefetch", K, iters, 1, W, X, out);
        run<0, 8>("v0 2-clip lanes, depth-2 prefetch", K, iters, 1, W, X, out);
        run<1, 16>("v1 4-clip lanes, no prefetch", K, iters, 0, W, X, out);
        run<1, 8>("v1 4-clip lanes, no prefetch", K, iters, 0, W, X, out);
        run<2, 16>("v2 4-clip lanes, x via cp.async smem", K, iters, 0, W, X, out);
    }
    return 0;
}
