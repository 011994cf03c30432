// tcgen05 (5th-gen tensor core) GEMM / implicit Conv1d for sm_100a with fp32-grade accuracy (3xTF32).
//
//   C[m, n] = epi( sum_tap sum_k A[m + tap - pad, k] * W[n, tap*Kcp + k] )
//
//  * A is fp32 row-major (rows = time-major activations, channels contiguous).  Conv1d inputs use a
//    zero-padded per-sequence layout [B][P + L + P][C], so a tap is just a row shift of the TMA box and
//    out-of-range coordinates are zero-filled by TMA: no im2col buffer, no bounds code.
//  * Operands are staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a 3/4-stage mbarrier ring,
//    multiplied by tcgen05.mma.kind::tf32 (M=128, N=BN) issued by one elected thread, accumulated in TMEM
//    (fp32) and read back with tcgen05.ld for the fused epilogue.
//  * Accuracy: TF32 alone (10-bit mantissa) cannot meet the 1e-3 end-to-end parity bound through ~50
//    layers, so every product is error-compensated:  a = a_hi + a_lo (a_hi = a with the low 13 mantissa
//    bits cleared, a_lo = a - a_hi, exact), likewise w;  D += a_hi*w_hi + a_lo*w_hi + a_hi*w_lo.
//    The dropped a_lo*w_lo term is 2^-22 relative.  w_hi/w_lo are precomputed at pack time; a_hi/a_lo
//    are produced IN SHARED MEMORY by the four otherwise idle epilogue warps, so activations cross HBM/L2
//    once, as plain fp32.
//  * Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
//    warps 2..5 = operand splitters during the main loop, then the epilogue (one TMEM lane per thread).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace l2s {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                 // fp32 elements per stage row = 128 bytes = one swizzle row
constexpr int TC_THREADS = 192;

struct TcParams {
    int M;                  // rows of the (padded) input row space to cover
    int N, Kc, Kcp, taps, pad;
    // row bookkeeping: input row m -> (seq = m / Lp_in, tp = m % Lp_in); valid iff P_in <= tp < P_in + L
    int Lp_in, P_in, L;
    int Lp_out, P_out;      // output row = seq*Lp_out + P_out + t
    float* C; int ldc;
    const float* bias; int act; const float* act_w;
    const float* addrow;    // [nseq][N], added before act
    const float* addpos; int ldpos;   // [L][ldpos], added after act
    const float* resid; int ldr;      // indexed by OUTPUT row, added after act
    int cstride, coff, chalf, chp;    // channel map (see gemm.cuh)
    int transposed;         // 1: C[(seq*N + n)*L + t]
    // generic tap table (row shift per tap) — used by the Conv3d stem; conv1d uses tap - pad
    int use_shift_table; int tap_shift[24];
    // stem mode: input rows are padded space-to-depth positions (b, tp, hp, wp); outputs are compact NHWC rows
    int stem, sT, sTp, sHp, sWp, sHo, sWo;
    int debug_niter;        // >0: truncate the K loop (timing experiments only; results are wrong)
    int ksteps_last;        // valid 8-wide k-steps in the last K chunk of a tap (1..4); zero-padded steps are skipped
    // gather mode (GATHER_A): source rows of ga_lda floats, ga_rows rows; the window of a row is Kc floats wide
    // (the caller guarantees zeroed slack around the buffers so tap-shifted rows never leave the allocation)
    const float* ga_A; const float* ga_Alo; int ga_lda; long long ga_rows; int ga_unchecked;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
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
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory operand descriptor (cute::UMMA::SmemDescriptor, version 1):
// start>>4 | LBO(unused)=1 | SBO = 8 rows * 128 B = 1024 B | layout SWIZZLE_128B (2).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, K-major both, N>>3, M>>4.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN>
struct TcSmem {
    static constexpr int STAGES = (BN >= 128) ? 3 : 4;
    static constexpr int A_BYTES = TC_BM * TC_BK * 4;       // 16 KB
    static constexpr int W_BYTES = BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// GATHER_A = false: the A tile arrives by TMA and warps 2..5 split it in place.
// GATHER_A = true : warps 2..5 ARE the A producers (software im2col): each thread loads 16-byte pieces of the
//                   tap-shifted rows straight from global/L1 (rows of p.ga_lda floats, window p.Kc wide — windows of
//                   neighbouring rows may overlap, which is how the Conv3d stem reuses its input through L1 instead
//                   of re-reading it 40x through L2), splits and stores hi/lo into the swizzled tiles.
template <int BN, bool GATHER_A>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapWhi,
               const __grid_constant__ CUtensorMap mapWlo, const TcParams p) {
    using SM = TcSmem<BN>;
    constexpr int STAGES = SM::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * SM::STAGE_BYTES);
    uint64_t* full = bars;                   // TMA landed
    uint64_t* split = bars + STAGES;         // a_hi/a_lo ready
    uint64_t* empty = bars + 2 * STAGES;     // MMAs of the stage retired
    uint64_t* accum = bars + 3 * STAGES;     // all MMAs retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
    const int kchunks = (p.Kc + TC_BK - 1) / TC_BK;
    const int niter = p.debug_niter > 0 ? p.debug_niter : p.taps * kchunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&split[s], 128); mbar_init(&empty[s], 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN < 32 ? 32 : BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int it = 0; it < niter; ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                const int tap = it / kchunks, kc = it - tap * kchunks;
                uint8_t* st = smem + s * SM::STAGE_BYTES;
                mbar_expect_tx(&full[s], (GATHER_A ? 0 : SM::A_BYTES) + 2 * SM::W_BYTES);
                if (!GATHER_A) tma_load_2d(&mapA, &full[s], st, kc * TC_BK, m0 + (p.use_shift_table ? p.tap_shift[tap] : tap - p.pad));
                tma_load_2d(&mapWhi, &full[s], st + 2 * SM::A_BYTES, tap * p.Kcp + kc * TC_BK, n0);
                tma_load_2d(&mapWlo, &full[s], st + 2 * SM::A_BYTES + SM::W_BYTES, tap * p.Kcp + kc * TC_BK, n0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, BN < 16 ? 16 : BN);
            for (int it = 0; it < niter; ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(&split[s], ph);
                if (GATHER_A) { mbar_wait(&full[s], ph); fence_proxy_async_smem(); }   // cp.async (generic proxy) data -> tensor core
                tc_fence_after();
                const int kc_ = it % kchunks;
                const int ksteps = (kc_ == kchunks - 1) ? p.ksteps_last : TC_BK / 8;
                const uint32_t base = smem_u32(smem + s * SM::STAGE_BYTES);
                const uint64_t a_hi = umma_desc_sw128(base), a_lo = umma_desc_sw128(base + SM::A_BYTES);
                const uint64_t w_hi = umma_desc_sw128(base + 2 * SM::A_BYTES), w_lo = umma_desc_sw128(base + 2 * SM::A_BYTES + SM::W_BYTES);
                for (int k = 0; k < ksteps; ++k) {
                    const uint64_t adv = (uint64_t)(k * 32 >> 4);          // 8 tf32 = 32 bytes along K inside the swizzle row
                    umma_tf32(tmem_d, a_lo + adv, w_hi + adv, idesc, (it | k) != 0);
                    umma_tf32(tmem_d, a_hi + adv, w_lo + adv, idesc, 1);
                    umma_tf32(tmem_d, a_hi + adv, w_hi + adv, idesc, 1);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(accum);
        }
    } else {
        // ===== splitters (main loop) =====
        const int t = threadIdx.x - 64;                 // 0..127
        if (GATHER_A) {
            // Software im2col with cp.async: the source is already split (ga_A = hi array, ga_Alo = lo array, same layout),
            // so a stage is 16 x 16-byte asynchronous copies per thread into the swizzled tiles and no ALU work.
            // Completion is signalled on split[s] by cp.async.mbarrier.arrive (one arrival per thread).
            const int c = t & 7, r0 = t >> 3;              // 16-byte chunk inside the 128-byte K row; first of 8 rows (stride 16)
            const long long rstep = 16LL * p.ga_lda;
            const float* base_hi = p.ga_A + ((long long)m0 + r0) * p.ga_lda + c * 4;
            const float* base_lo = p.ga_Alo + ((long long)m0 + r0) * p.ga_lda + c * 4;
            const uint32_t off0 = (uint32_t)(r0 * 128 + ((c ^ (r0 & 7)) << 4));     // (r0 + 16 i) & 7 == r0 & 7
            int it = 0;
            for (int tap = 0; tap < p.taps && it < niter; ++tap) {
                const long long shift = (long long)(p.use_shift_table ? p.tap_shift[tap] : tap - p.pad) * p.ga_lda;
                for (int kc = 0; kc < kchunks && it < niter; ++kc, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    if (kc * TC_BK + c * 4 < p.Kc) {        // chunks beyond Kc are never read (ksteps_last)
                        const uint32_t dst = smem_u32(smem + s * SM::STAGE_BYTES) + off0;
                        const float* sh = base_hi + shift + kc * TC_BK;
                        const float* sl = base_lo + shift + kc * TC_BK;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + i * 2048), "l"(sh + i * rstep) : "memory");
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + SM::A_BYTES + i * 2048), "l"(sl + i * rstep) : "memory");
                        }
                    }
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&split[s])) : "memory");
                }
            }
        }
        for (int it = 0; !GATHER_A && it < niter; ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(&full[s], ph);
            float4* hi = reinterpret_cast<float4*>(smem + s * SM::STAGE_BYTES);
            float4* lo = reinterpret_cast<float4*>(smem + s * SM::STAGE_BYTES + SM::A_BYTES);
#pragma unroll
            for (int i = 0; i < SM::A_BYTES / 16 / 128; ++i) {
                const int idx = t + i * 128;             // same (swizzled) offset in both tiles
                float4 a = hi[idx];
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u); l.x = a.x - h.x;
                h.y = __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u); l.y = a.y - h.y;
                h.z = __uint_as_float(__float_as_uint(a.z) & 0xFFFFE000u); l.z = a.z - h.z;
                h.w = __uint_as_float(__float_as_uint(a.w) & 0xFFFFE000u); l.w = a.w - h.w;
                hi[idx] = h; lo[idx] = l;
            }
            fence_proxy_async_smem();                    // generic-proxy writes -> visible to the tensor core (async proxy)
            mbar_arrive(&split[s]);
        }
        // ===== epilogue =====
        mbar_wait(accum, 0);
        tc_fence_after();
        const int lg = warp & 3;                          // TMEM lane group this warp may access
        const int row = lg * 32 + lane;
        const int m = m0 + row;
        int seq, tt; bool valid; size_t orow;
        if (!p.stem) {
            seq = m / p.Lp_in;
            tt = m - seq * p.Lp_in - p.P_in;
            valid = (m < p.M) && (tt >= 0) && (tt < p.L);
            orow = (size_t)seq * p.Lp_out + p.P_out + tt;
        } else {
            const int wp = m % p.sWp; int r = m / p.sWp;
            const int hp = r % p.sHp; r /= p.sHp;
            const int tpp = r % p.sTp; const int b = r / p.sTp;
            seq = 0; tt = 0;
            valid = (m < p.M) && tpp >= 2 && tpp < p.sT + 2 && hp >= 2 && hp < p.sHo + 2 && wp >= 2 && wp < p.sWo + 2;
            orow = ((size_t)(b * p.sT + tpp - 2) * p.sHo + (hp - 2)) * p.sWo + (wp - 2);
        }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            if (n0 + c0 >= p.N) break;                    // warp-uniform
            uint32_t v[16];
            tmem_ld16(tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, v);
            if (valid) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = n0 + c0 + j;
                    if (n < p.N) {
                        float x = __uint_as_float(v[j]);
                        if (p.bias) x += __ldg(p.bias + n);
                        if (p.addrow) x += p.addrow[(size_t)seq * p.N + n];
                        x = apply_act(x, p.act, p.act_w ? __ldg(p.act_w + n) : 1.f);
                        if (p.addpos) x += __ldg(p.addpos + (size_t)tt * p.ldpos + n);
                        if (p.resid) x += p.resid[orow * p.ldr + n];
                        if (p.transposed) {
                            p.C[((size_t)seq * p.N + n) * p.L + tt] = x;
                        } else {
                            int l = n * p.cstride + p.coff;
                            if (p.chalf > 0 && l >= p.chalf) l = l - p.chalf + p.chp;
                            p.C[orow * p.ldc + l] = x;
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(BN < 32 ? 32 : BN) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2-D fp32 tensor map: inner dimension `cols` (contiguous), `rows` rows of `ld` floats; box = [32 cols][box_rows].
inline bool make_map_2d(CUtensorMap* map, const float* ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {TC_BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

struct TcOperands {
    const float* A; int a_cols, a_rows, lda;         // activations
    const float* Whi; const float* Wlo; int w_cols;  // [N][w_cols], w_cols = taps*Kcp
};

template <int BN, bool G>
inline cudaError_t launch_tc_inst(dim3 grid, const CUtensorMap& mA, const CUtensorMap& mWh, const CUtensorMap& mWl, const TcParams& p, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<BN, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<BN>::TOTAL);
    if (e != cudaSuccess) return e;
    tc_gemm_kernel<BN, G><<<grid, TC_THREADS, TcSmem<BN>::TOTAL, s>>>(mA, mWh, mWl, p);
    return cudaGetLastError();
}

inline const char* launch_tc_gemm(const TcOperands& o, TcParams p, cudaStream_t s, bool gather = false) {
    CUtensorMap mA, mWh, mWl;
    const int BN = (p.N <= 32) ? 32 : (p.N <= 64 ? 64 : 128);
    const int rem = p.Kc % TC_BK;
    p.ksteps_last = rem == 0 ? TC_BK / 8 : (rem + 7) / 8;
    p.ga_A = o.A; p.ga_lda = o.lda; p.ga_rows = o.a_rows;
    if (const char* e = getenv("L2S_TC_DEBUG_NITER")) p.debug_niter = atoi(e);
    if (gather) {
        if ((o.lda & 3) || (p.Kc & 3) || (reinterpret_cast<uintptr_t>(o.A) & 15)) return "gather-A needs 16-byte aligned rows";
        // the A map is unused in gather mode but must be a valid object: describe the W matrix again
        if (!make_map_2d(&mA, o.Whi, o.w_cols, p.N, o.w_cols, BN)) return "cuTensorMapEncodeTiled failed";
    } else if (!make_map_2d(&mA, o.A, o.a_cols, o.a_rows, o.lda, TC_BM)) return "cuTensorMapEncodeTiled(A) failed";
    if (!make_map_2d(&mWh, o.Whi, o.w_cols, p.N, o.w_cols, BN)) return "cuTensorMapEncodeTiled(Whi) failed";
    if (!make_map_2d(&mWl, o.Wlo, o.w_cols, p.N, o.w_cols, BN)) return "cuTensorMapEncodeTiled(Wlo) failed";
    dim3 grid(ceil_div(p.M, TC_BM), ceil_div(p.N, BN));
    cudaError_t e;
    if (gather) {
        e = BN == 32 ? launch_tc_inst<32, true>(grid, mA, mWh, mWl, p, s) : BN == 64 ? launch_tc_inst<64, true>(grid, mA, mWh, mWl, p, s)
                                                                                  : launch_tc_inst<128, true>(grid, mA, mWh, mWl, p, s);
    } else {
        e = BN == 32 ? launch_tc_inst<32, false>(grid, mA, mWh, mWl, p, s) : BN == 64 ? launch_tc_inst<64, false>(grid, mA, mWh, mWl, p, s)
                                                                                   : launch_tc_inst<128, false>(grid, mA, mWh, mWl, p, s);
    }
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace l2s
