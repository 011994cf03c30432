// Shared helpers for the lip2speech_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>

namespace l2s {

constexpr int ACT_NONE = 0;
constexpr int ACT_RELU = 1;
constexpr int ACT_SILU = 2;
constexpr int ACT_PSINE = 3;   // sin(x) * w[channel]   (reference decoder.py:43-70)
constexpr int ACT_PRELU = 4;   // x>=0 ? x : w[channel]*x (reference video.py:66)

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float siluf_acc(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float v, int act, float w) {
    switch (act) {
        case ACT_RELU: return v > 0.f ? v : 0.f;
        case ACT_SILU: return siluf_acc(v);
        case ACT_PSINE: return sinf(v) * w;
        case ACT_PRELU: return v >= 0.f ? v : w * v;
        default: return v;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// L2-only loads for data produced by other CTAs of the same persistent kernel (L1 is not coherent).
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float ldcg1(const float* p) { return __ldcg(p); }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Grid-wide barrier for cooperative (co-resident) launches.  `counter` is zeroed by the host before
// the launch and only ever grows; `target` is the per-thread running expected value.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target, unsigned nblocks) {
    target += nblocks;
    __syncthreads();
    if (threadIdx.x == 0) {
        red_release_add_u32(counter, 1u);             // release: orders this CTA's prior global writes
        while (ld_acquire_u32(counter) < target) {}
    }
    __syncthreads();
}

// Split-phase variant: arrive after a stage's global writes, do work that only needs OLDER data, then wait.
__device__ __forceinline__ void grid_arrive(unsigned* counter) {
    __syncthreads();                                   // all warps of this CTA finished the stage's writes
    if (threadIdx.x == 0) red_release_add_u32(counter, 1u);
}
__device__ __forceinline__ void grid_wait(const unsigned* counter, unsigned target) {
    if (threadIdx.x == 0) {
        while (ld_acquire_u32(counter) < target) {}
    }
    __syncthreads();
}

// ---- mbarrier / shared-memory PTX wrappers ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine; completion is signalled on `bar` (complete_tx bytes).
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace l2s
