// raw fma.rn.f32x2 (FFMA2) issue rate on sm_100a vs scalar FFMA
#include <cstdio>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    ra = *reinterpret_cast<unsigned long long*>(&a); rb = *reinterpret_cast<unsigned long long*>(&b); rc = *reinterpret_cast<unsigned long long*>(&c);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__global__ void __launch_bounds__(512, 1) k2(float* out, int iters, float a0, float b0) {
    float2 acc[16];
    float2 a[4], b[4];
    for (int i = 0; i < 4; ++i) { a[i] = make_float2(a0 + i, a0 + i); b[i] = make_float2(b0 + i + threadIdx.x, b0 - i); }
    for (int i = 0; i < 16; ++i) acc[i] = make_float2((float)i, (float)-i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma2(a[i & 3], b[i >> 2], acc[i]);
    }
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * 512 + threadIdx.x] = s;
}
__global__ void __launch_bounds__(512, 1) k1(float* out, int iters, float a0, float b0) {
    float acc[32]; float a[4], b[8];
    for (int i = 0; i < 4; ++i) a[i] = a0 + i;
    for (int i = 0; i < 8; ++i) b[i] = b0 + i + threadIdx.x;
    for (int i = 0; i < 32; ++i) acc[i] = (float)i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = fmaf(a[i & 3], b[i >> 2], acc[i]);
    }
    float s = 0.f;
    for (int i = 0; i < 32; ++i) s += acc[i];
    out[blockIdx.x * 512 + threadIdx.x] = s;
}
int main() {
    float* o; cudaMalloc(&o, 148 * 512 * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    k1<<<148, 512>>>(o, 100, 1.f, 2.f);
    cudaEventRecord(e0); k1<<<148, 512>>>(o, iters, 1.f, 2.f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA : %.1f FMA/clk/SM @1.965GHz  (%.3f ms)\n", 512.0 * 32 * iters / (ms * 1e-3 * 1.965e9), ms);
    k2<<<148, 512>>>(o, 100, 1.f, 2.f);
    cudaEventRecord(e0); k2<<<148, 512>>>(o, iters, 1.f, 2.f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA2: %.1f FMA/clk/SM @1.965GHz  (%.3f ms)  %s\n", 512.0 * 32 * iters / (ms * 1e-3 * 1.965e9), ms, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
