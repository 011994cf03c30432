// Micro-benchmark of the 8-clip tensor-core pass (matvec.cuh: mv8_*) and of the raw mma.sync tf32 issue rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lip2speech_b200/csrc tools/mv8_bench.cu -o tools/mv8_bench.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "matvec.cuh"
using namespace l2s;

__global__ void __launch_bounds__(512, 1) mma_rate_kernel(float* out, int iters, int nacc) {
    float acc[8][4];
    for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
    uint32_t a = threadIdx.x, b = blockIdx.x + 1;
    for (int it = 0; it < iters; ++it) {
        if (nacc == 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) mma_tf32(acc[j], a, a + 1, a + 2, a + 3, b, b + 1);
        } else if (nacc == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) mma_tf32(acc[j & 1], a, a + 1, a + 2, a + 3, b, b + 1);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) mma_tf32(acc[0], a, a + 1, a + 2, a + 3, b, b + 1);
        }
    }
    float s = 0.f;
    for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) s += acc[j][i];
    out[blockIdx.x * 512 + threadIdx.x] = s;
}

// distinct operands per MMA (8 A fragments x 8 B fragments from registers), 8 independent accumulators
__global__ void __launch_bounds__(512, 1) mma_rate_distinct_kernel(float* out, int iters, const uint32_t* __restrict__ src) {
    float acc[8][4];
    for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
    uint32_t a[8][4], b[8][2];
    for (int j = 0; j < 8; ++j) { for (int i = 0; i < 4; ++i) a[j][i] = src[(threadIdx.x + 37 * (4 * j + i)) & 1023]; for (int i = 0; i < 2; ++i) b[j][i] = src[(threadIdx.x + 91 * (2 * j + i) + 5) & 1023]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_tf32(acc[k], a[j][0], a[j][1], a[j][2], a[j][3], b[(j + k) & 7][0], b[(j + k) & 7][1]);
    }
    float s = 0.f;
    for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) s += acc[j][i];
    out[blockIdx.x * 512 + threadIdx.x] = s;
}

// mv8_mma without the hi/lo splits (same loads, same MMA count; operands used raw): isolates the cost of the split ALU work
template <int RT>
__device__ __forceinline__ void mv8_mma_nosplit(const float* __restrict__ W, int ldw, int R, int K, const float (&x)[MV8_DEPTH][4], float (&acc)[RT][4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int npw = K / (MV_KC * MV_WARPS);
#pragma unroll
    for (int d = 0; d < MV8_DEPTH; ++d)
        if (d < npw) {
            uint32_t xh[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xh[i] = __float_as_uint(x[d][i]);
#pragma unroll
            for (int rt = 0; rt < RT; ++rt) {
                const float4 wa = *reinterpret_cast<const float4*>(W + (size_t)min(rt * 16 + g, R - 1) * ldw + (warp + d * MV_WARPS) * MV_KC + 4 * t);
                const float4 wb = *reinterpret_cast<const float4*>(W + (size_t)min(rt * 16 + g + 8, R - 1) * ldw + (warp + d * MV_WARPS) * MV_KC + 4 * t);
                const uint32_t ah[4] = {__float_as_uint(wa.x), __float_as_uint(wa.y), __float_as_uint(wa.z), __float_as_uint(wa.w)};
                const uint32_t bh[4] = {__float_as_uint(wb.x), __float_as_uint(wb.y), __float_as_uint(wb.z), __float_as_uint(wb.w)};
#pragma unroll
                for (int s = 0; s < 2; ++s)
#pragma unroll
                    for (int rep = 0; rep < 3; ++rep)
                        mma_tf32(acc[rt], ah[2 * s], bh[2 * s], ah[2 * s + 1], bh[2 * s + 1], xh[2 * s], xh[2 * s + 1]);
            }
        }
}

template <int MAXRT>
__global__ void __launch_bounds__(512, 1) pass_kernel(const float* __restrict__ Wg, const float* __restrict__ X, float* __restrict__ out,
                                                       int R, int K, int ldw, int iters, int do_reduce) {
    extern __shared__ __align__(16) float smem[];
    float* red = smem;
    float* wsm = smem + 6144;
    for (int i = threadIdx.x; i < R * K; i += 512) wsm[(i / K) * ldw + (i % K)] = Wg[(size_t)blockIdx.x * R * K + i];
    __syncthreads();
    float sink = 0.f;
    for (int it = 0; it < iters; ++it) {
        const float* x = X + (size_t)(it & 3) * K * 8;
        float acc[MAXRT][4];
        mv8_zero<MAXRT>(acc);
        if (do_reduce >= 2) { float xx[MV8_DEPTH][4]; mv8_load<MV8_DEPTH>(x, K, xx); mv8_mma_nosplit<MAXRT>(wsm, ldw, R, K, xx, acc); }   // mode 2: no splits, no reduce
        else mv8_accumulate<MAXRT>(wsm, ldw, 0, R, x, K, acc);
        if (do_reduce == 1) { for (int rd = 0; rd < (MAXRT + 2) / 3; ++rd) { sink += mv8_reduce_round<MAXRT>(acc, rd, red); __syncthreads(); } }
        else for (int r = 0; r < MAXRT; ++r) sink += acc[r][0] + acc[r][1] + acc[r][2] + acc[r][3];
    }
    out[(size_t)blockIdx.x * 512 + threadIdx.x] = sink;
}

int main() {
    float* o; cudaMalloc(&o, 148 * 512 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int nacc : {8, 2, 1}) {
        mma_rate_kernel<<<148, 512>>>(o, 100, nacc);
        const int iters = 20000;
        cudaEventRecord(e0);
        mma_rate_kernel<<<148, 512>>>(o, iters, nacc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double mmas_per_smsp = 4.0 * 8 * iters;        // 4 warps per scheduler
        printf("mma.sync m16n8k8 tf32, %d independent accumulators/warp, 16 warps/SM: %.2f clk per MMA per SMSP (1.965 GHz) -> %.0f MAC/clk/SM\n",
               nacc, ms * 1e-3 * 1.965e9 / mmas_per_smsp, 1024.0 * 4 / (ms * 1e-3 * 1.965e9 / mmas_per_smsp));
    }
    {
        uint32_t* src; cudaMalloc(&src, 4096); cudaMemset(src, 0x3c, 4096);
        for (int warps : {16, 8, 4}) {
            mma_rate_distinct_kernel<<<148, warps * 32>>>(o, 10, src);
            const int iters = 4000;
            cudaEventRecord(e0);
            mma_rate_distinct_kernel<<<148, warps * 32>>>(o, iters, src);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double mmas_per_smsp = (warps / 4.0) * 64 * iters;
            printf("mma.sync tf32 distinct operands, %d warps/SM: %.2f clk per MMA per SMSP\n", warps, ms * 1e-3 * 1.965e9 / mmas_per_smsp);
        }
    }
    const int KMAX = 1536, RMAX = 48;
    std::vector<float> hW((size_t)148 * RMAX * KMAX), hX((size_t)4 * KMAX * 32);
    srand(1);
    for (auto& v : hW) v = (rand() / (float)RAND_MAX - 0.5f) * 0.1f;
    for (auto& v : hX) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    float *W, *X;
    cudaMalloc(&W, hW.size() * 4); cudaMalloc(&X, hX.size() * 4);
    cudaMemcpy(W, hW.data(), hW.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(X, hX.data(), hX.size() * 4, cudaMemcpyHostToDevice);
    struct Cfg { int R, K; } cfgs[] = {{32, 1536}, {32, 1024}, {32, 512}, {48, 1024}, {48, 512}, {16, 512}, {16, 1024}, {24, 1024}, {48, 256}};
    for (auto cf : cfgs)
        for (int red = 0; red < 3; ++red) {
            const int ldw = cf.K + 16, iters = 2000;
            size_t smem = (size_t)(6144 + cf.R * ldw) * 4;
            cudaFuncSetAttribute(pass_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (cf.R <= 32) { cudaFuncSetAttribute(pass_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }
            cudaEventRecord(e0);
            if (cf.R <= 32) pass_kernel<2><<<148, 512, smem>>>(W, X, o, cf.R, cf.K, ldw, iters, red);
            else pass_kernel<3><<<148, 512, smem>>>(W, X, o, cf.R, cf.K, ldw, iters, red);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            cudaError_t e = cudaGetLastError();
            printf("mv8 pass R=%2d K=%4d reduce=%d : %6.3f us/pass %s\n", cf.R, cf.K, red, ms * 1e3 / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
