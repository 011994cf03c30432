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
//  * BF16 variant (the Conv3d stem at precision = bf16, BASELINE config 2 "bf16 frontend"): operands are single bf16
//    arrays (no hi/lo pair), one tcgen05.mma.kind::f16 per 16-wide k-step, one 128-byte swizzle row (64 bf16) per tap.
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
    int debug_skip;         // timing experiments only: bit0 skip epilogue body, bit1 skip worker split, bit2 skip MMA issue, bit3 skip A TMA
    int debug_niter;        // >0: truncate the K loop (timing experiments only; results are wrong)
    int ksteps_last;        // valid 8-wide k-steps in the last K chunk of a tap (1..4); zero-padded steps are skipped
    // gather mode (GATHER_A): source rows of ga_lda floats, ga_rows rows; the window of a row is Kc floats wide
    // (the caller guarantees zeroed slack around the buffers so tap-shifted rows never leave the allocation)
    const float* ga_A; const float* ga_Alo; int ga_lda; long long ga_rows; int ga_unchecked;
};

// ---- PTX wrappers (mbarrier helpers live in common.cuh) ---------------------------------------------
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
// D=F32, A=B=BF16.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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

// Stem mode: a 128-row tile whose first and last rows both lie in temporal padding planes (tp < 2 or tp >= T + 2) holds
// no valid output row (a tile spans at most two planes) and is skipped by every warp role alike.
__device__ __forceinline__ bool tc_stem_pad_tile(const TcParams& p, int m0) {
    if (!p.stem) return false;
    const int plane = p.sHp * p.sWp;
    const int t0 = (m0 / plane) % p.sTp;
    const int m1 = min(m0 + TC_BM - 1, p.M - 1);
    const int t1 = (m1 / plane) % p.sTp;
    const bool pad0 = t0 < 2 || t0 >= p.sT + 2, pad1 = t1 < 2 || t1 >= p.sT + 2;
    return pad0 && pad1;
}

template <int BN>
struct TcSmem {
    static constexpr int STAGES = (BN >= 128) ? 3 : 4;
    static constexpr int A_BYTES = TC_BM * TC_BK * 4;       // 16 KB
    static constexpr int W_BYTES = BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Persistent, warp-specialised kernel: grid = min(#tiles, #SMs); every CTA walks tiles t = blockIdx.x, +gridDim.x, ...
// (n fastest, so neighbouring CTAs share A rows in L2) with ONE continuous smem ring and a double-buffered TMEM
// accumulator, so TMEM allocation, barrier setup, pipeline fill and the epilogue of tile i overlap tile i+1.
//   warp 0      : TMA producer (W hi/lo always; the A tile too unless GATHER_A)
//   warp 1      : TMEM allocator + MMA issuer (one elected thread)
//   warps 2..9  : A-side workers (256 threads)
//                   GATHER_A = false: split the TMA-landed fp32 tile in place into hi / lo
//                   GATHER_A = true : software im2col with cp.async from pre-split hi/lo arrays (rows of ga_lda floats whose
//                                     Kc-wide windows may overlap — the Conv3d stem re-reads its input through L1, not L2)
//   warps 10..17: epilogue (two warps per TMEM lane group; bias/act/residual/positional add, channel-mapped or transposed store)
constexpr int TC_WORKERS = 256;
constexpr int TC_EPI = 256;
constexpr int TC_THREADS_P = 64 + TC_WORKERS + TC_EPI;

template <int BN, bool GATHER_A, bool BF16 = false>
__global__ void __launch_bounds__(TC_THREADS_P, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapWhi,
               const __grid_constant__ CUtensorMap mapWlo, const TcParams p) {
    using SM = TcSmem<BN>;
    constexpr int STAGES = SM::STAGES;
    constexpr int TCOLS = (4 * BN < 32) ? 32 : 4 * BN;      // two accumulator buffers x [a*w_hi (+a_lo*w_hi) | a_hi*w_lo]
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * SM::STAGE_BYTES);
    uint64_t* full = bars;                    // TMA landed
    uint64_t* split = bars + STAGES;          // a_hi/a_lo ready
    uint64_t* empty = bars + 2 * STAGES;      // MMAs of the stage retired
    uint64_t* tfull = bars + 3 * STAGES;      // [2] accumulator complete
    uint64_t* tempty = bars + 3 * STAGES + 2; // [2] accumulator drained by the epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
    __shared__ float ep_s[TC_EPI / 32][32 * 17];            // per-warp transpose scratch
    __shared__ int4 rinfo_s[TC_EPI / 32][32];               // per-warp row info (valid, seq, t, output row)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    static_assert(!BF16 || GATHER_A, "the bf16 variant gathers its A operand");
    const int kchunks = BF16 ? 1 : (p.Kc + TC_BK - 1) / TC_BK;      // bf16: one 64-wide (128-byte) window per tap
    // bf16 stem: TWO taps per ring stage (the stage's A-lo / W-lo slots hold the second tap's tiles): half the stage hand-offs
    // (empty -> gather + TMA -> MMA issue -> commit) per 128-row tile; measured 2.96 -> 2.91 ms for the B=32 frontend.
    const int niter = p.debug_niter > 0 ? p.debug_niter : (BF16 ? (p.taps + 1) / 2 : p.taps * kchunks);
    const int ntn = (p.N + BN - 1) / BN;
    const int ntiles = ((p.M + TC_BM - 1) / TC_BM) * ntn;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&split[s], TC_WORKERS); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], TC_EPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int git = 0;                                    // global stage counter across tiles
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int m0 = (tile / ntn) * TC_BM, n0 = (tile % ntn) * BN;
                if (tc_stem_pad_tile(p, m0)) continue;
                for (int it = 0; it < niter; ++it, ++git) {
                    const int s = git % STAGES, ph = (git / STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    const int tap = it / kchunks, kc = it - tap * kchunks;
                    uint8_t* st = smem + s * SM::STAGE_BYTES;
                    const bool skipA = GATHER_A || (p.debug_skip & 8);
                    if (p.debug_skip & 32) { mbar_arrive(&full[s]); continue; }      // skip all TMA traffic
                    if (BF16) {                                  // bf16 W tiles of taps 2 it, 2 it + 1: [BN rows][64 bf16] = W_BYTES each
                        const bool two = 2 * it + 1 < p.taps;
                        mbar_expect_tx(&full[s], (two ? 2 : 1) * SM::W_BYTES);
                        tma_load_2d(&mapWhi, &full[s], st + 2 * SM::A_BYTES, 2 * it * 64, n0);
                        if (two) tma_load_2d(&mapWhi, &full[s], st + 2 * SM::A_BYTES + SM::W_BYTES, (2 * it + 1) * 64, n0);
                        continue;
                    }
                    mbar_expect_tx(&full[s], (skipA ? 0 : SM::A_BYTES) + 2 * SM::W_BYTES);
                    if (!skipA) tma_load_2d(&mapA, &full[s], st, kc * TC_BK, m0 + (p.use_shift_table ? p.tap_shift[tap] : tap - p.pad));
                    tma_load_2d(&mapWhi, &full[s], st + 2 * SM::A_BYTES, tap * p.Kcp + kc * TC_BK, n0);
                    tma_load_2d(&mapWlo, &full[s], st + 2 * SM::A_BYTES + SM::W_BYTES, tap * p.Kcp + kc * TC_BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // Per k-step (8 tf32 = 32 bytes) two MMAs instead of three: the W tiles [w_hi ; w_lo] are contiguous in shared
            // memory, so ONE N=2*BN instruction computes a_hi*w_hi (columns 0..BN) and a_hi*w_lo (columns BN..2BN) with a
            // single read of the A tile — SS-mode MMAs at small N are bound by the A-operand read, not by N — and a second
            // N=BN instruction adds a_lo*w_hi.  The epilogue sums the two column groups.
            constexpr uint32_t idesc2 = umma_idesc_tf32(TC_BM, 2 * BN);
            constexpr uint32_t idesc1 = umma_idesc_tf32(TC_BM, BN < 16 ? 16 : BN);
            uint64_t d_ahi[STAGES], d_alo[STAGES], d_w[STAGES];
#pragma unroll
            for (int s = 0; s < STAGES; ++s) {
                const uint32_t base = smem_u32(smem + s * SM::STAGE_BYTES);
                d_ahi[s] = umma_desc_sw128(base); d_alo[s] = umma_desc_sw128(base + SM::A_BYTES);
                d_w[s] = umma_desc_sw128(base + 2 * SM::A_BYTES);
            }
            int git = 0, lt = 0;                            // lt = local tile counter (accumulator buffer = lt & 1)
            int s = 0, ph = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                if (tc_stem_pad_tile(p, (tile / ntn) * TC_BM)) continue;
                const int buf = lt & 1, bph = (lt >> 1) & 1;
                mbar_wait(&tempty[buf], bph ^ 1);           // epilogue has drained this accumulator (two tiles ago)
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * 2 * BN);
                int kc_ = 0;
                for (int it = 0; it < niter; ++it, ++git) {
                    mbar_wait(&split[s], ph);
                    if (GATHER_A) { mbar_wait(&full[s], ph); fence_proxy_async_smem(); }   // cp.async (generic proxy) data -> tensor core
                    tc_fence_after();
                    const int ksteps = (p.debug_skip & 4) ? 0 : ((kc_ == kchunks - 1) ? p.ksteps_last : TC_BK / 8);
                    if (++kc_ == kchunks) kc_ = 0;
                    uint64_t ahi = d_ahi[s], alo = d_alo[s], w = d_w[s];
                    uint32_t acc = it != 0;
                    if (BF16) {                                  // 64 bf16 per stage row = four K=16 steps, one product each
                        constexpr uint32_t idesc16 = umma_idesc_bf16(TC_BM, BN < 16 ? 16 : BN);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (!(p.debug_skip & 4)) umma_f16(tmem_d, ahi, w, idesc16, acc);
                            acc = 1; ahi += 2; w += 2;               // 16 bf16 = 32 bytes = 2 descriptor units along K
                        }
                        if (2 * it + 1 < p.taps) {                   // second tap of the stage: A in the "lo" slot, W one tile further
                            w = d_w[s] + (uint64_t)(SM::W_BYTES >> 4);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (!(p.debug_skip & 4)) umma_f16(tmem_d, alo, w, idesc16, 1u);
                                alo += 2; w += 2;
                            }
                        }
                        umma_commit(&empty[s]);
                        if (++s == STAGES) { s = 0; ph ^= 1; }
                        continue;
                    }
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        if (k < ksteps) {
                            umma_tf32(tmem_d, ahi, w, idesc2, acc);
                            umma_tf32(tmem_d, alo, w, idesc1, 1);
                            acc = 1; ahi += 2; alo += 2; w += 2;           // 8 tf32 = 32 bytes = 2 descriptor units along K
                        }
                    }
                    umma_commit(&empty[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit(&tfull[buf]);
                ++lt;
            }
        }
    } else if (warp < 2 + TC_WORKERS / 32) {
        // ===== A-side workers =====
        const int t = threadIdx.x - 64;                     // 0..255
        int git = 0;
        if (BF16) {
            // bf16 gather: source rows of 16 bf16 (32 bytes) per space-to-depth position; a tap's window is 4 positions =
            // 128 contiguous bytes = one swizzle row
            const int c = t & 7, r0 = t >> 3;
            const uint32_t off0 = (uint32_t)(r0 * 128 + ((c ^ (r0 & 7)) << 4));
            const char* src0 = reinterpret_cast<const char*>(p.ga_A);
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int m0 = (tile / ntn) * TC_BM;
                if (tc_stem_pad_tile(p, m0)) continue;
                const char* base = src0 + ((long long)m0 + r0) * 32 + c * 16;
                for (int it = 0; it < niter; ++it, ++git) {
                    const int s = git % STAGES, ph = (git / STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    if (p.debug_skip & 2) { mbar_arrive(&split[s]); continue; }      // timing experiments: no gather
                    const uint32_t dst = smem_u32(smem + s * SM::STAGE_BYTES) + off0;
                    const char* sp = base + (long long)p.tap_shift[2 * it] * 32;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + i * 4096), "l"(sp + i * 1024) : "memory");
                    if (2 * it + 1 < p.taps) {
                        const char* sq = base + (long long)p.tap_shift[2 * it + 1] * 32;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + SM::A_BYTES + i * 4096), "l"(sq + i * 1024) : "memory");
                    }
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&split[s])) : "memory");
                }
            }
        } else if (GATHER_A) {
            const int c = t & 7, r0 = t >> 3;               // 16-byte chunk inside the 128-byte K row; first of 4 rows (stride 32)
            const long long rstep = 32LL * p.ga_lda;
            const uint32_t off0 = (uint32_t)(r0 * 128 + ((c ^ (r0 & 7)) << 4));     // (r0 + 32 i) & 7 == r0 & 7
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int m0 = (tile / ntn) * TC_BM;
                if (tc_stem_pad_tile(p, m0)) continue;
                const float* base_hi = p.ga_A + ((long long)m0 + r0) * p.ga_lda + c * 4;
                const float* base_lo = p.ga_Alo + ((long long)m0 + r0) * p.ga_lda + c * 4;
                int it = 0;
                for (int tap = 0; tap < p.taps && it < niter; ++tap) {
                    const long long shift = (long long)(p.use_shift_table ? p.tap_shift[tap] : tap - p.pad) * p.ga_lda;
                    for (int kc = 0; kc < kchunks && it < niter; ++kc, ++it, ++git) {
                        const int s = git % STAGES, ph = (git / STAGES) & 1;
                        mbar_wait(&empty[s], ph ^ 1);
                        if (kc * TC_BK + c * 4 < p.Kc && !(p.debug_skip & 16)) {     // chunks beyond Kc are never read (ksteps_last)
                            const uint32_t dst = smem_u32(smem + s * SM::STAGE_BYTES) + off0;
                            const float* sh = base_hi + shift + kc * TC_BK;
                            const float* sl = base_lo + shift + kc * TC_BK;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + i * 4096), "l"(sh + i * rstep) : "memory");
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + SM::A_BYTES + i * 4096), "l"(sl + i * rstep) : "memory");
                            }
                        }
                        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&split[s])) : "memory");
                    }
                }
            }
        } else {
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                if (tc_stem_pad_tile(p, (tile / ntn) * TC_BM)) continue;
                for (int it = 0; it < niter; ++it, ++git) {
                    const int s = git % STAGES, ph = (git / STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    float4* hi = reinterpret_cast<float4*>(smem + s * SM::STAGE_BYTES);
                    float4* lo = reinterpret_cast<float4*>(smem + s * SM::STAGE_BYTES + SM::A_BYTES);
                    if (p.debug_skip & 2) { mbar_arrive(&split[s]); continue; }
                    float4 a[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = hi[t + i * TC_WORKERS];      // same (swizzled) offset in both tiles
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4 h, l;
                        h.x = __uint_as_float(__float_as_uint(a[i].x) & 0xFFFFE000u); l.x = a[i].x - h.x;
                        h.y = __uint_as_float(__float_as_uint(a[i].y) & 0xFFFFE000u); l.y = a[i].y - h.y;
                        h.z = __uint_as_float(__float_as_uint(a[i].z) & 0xFFFFE000u); l.z = a[i].z - h.z;
                        h.w = __uint_as_float(__float_as_uint(a[i].w) & 0xFFFFE000u); l.w = a[i].w - h.w;
                        hi[t + i * TC_WORKERS] = h; lo[t + i * TC_WORKERS] = l;
                    }
                    fence_proxy_async_smem();                // generic-proxy writes -> visible to the tensor core (async proxy)
                    mbar_arrive(&split[s]);
                }
            }
        }
    } else {
        // ===== epilogue (8 warps: two per TMEM lane group, alternating 16-column chunks) =====
        const int e = warp - (2 + TC_WORKERS / 32);        // 0..7
        const int lg = warp & 3;                            // TMEM lane group this warp may access
        const int half = e >> 2;
        const int row = lg * 32 + lane;
        float* ep = ep_s[e];
        int4* rinfo = rinfo_s[e];
        int lt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int m0 = (tile / ntn) * TC_BM, n0 = (tile % ntn) * BN;
            if (tc_stem_pad_tile(p, m0)) continue;
            const int buf = lt & 1, bph = (lt >> 1) & 1;
            ++lt;
            mbar_wait(&tfull[buf], bph);
            tc_fence_after();
            if (p.debug_skip & 1) { tc_fence_before(); mbar_arrive(&tempty[buf]); continue; }
            const uint32_t tmem_d = tmem_base + (uint32_t)(buf * 2 * BN);
            const int m = m0 + row;
            int seq, tt; bool valid; int orow;
            if (!p.stem) {
                seq = m / p.Lp_in;
                tt = m - seq * p.Lp_in - p.P_in;
                valid = (m < p.M) && (tt >= 0) && (tt < p.L);
                orow = seq * p.Lp_out + p.P_out + tt;
            } else {
                const int wp = m % p.sWp; int r = m / p.sWp;
                const int hp = r % p.sHp; r /= p.sHp;
                const int tpp = r % p.sTp; const int b = r / p.sTp;
                seq = 0; tt = 0;
                valid = (m < p.M) && tpp >= 2 && tpp < p.sT + 2 && hp >= 2 && hp < p.sHo + 2 && wp >= 2 && wp < p.sWo + 2;
                orow = ((b * p.sT + tpp - 2) * p.sHo + (hp - 2)) * p.sWo + (wp - 2);
            }
            // Row bookkeeping is computed once per tile by the thread that owns the TMEM lane and shared with the warp;
            // each 32x16 accumulator chunk is transposed through shared memory so that lanes map to CHANNELS: bias /
            // residual loads and the output stores are then contiguous within a row (a lane-per-row store would touch
            // 32 different rows = 32 sectors per instruction).
            rinfo[lane] = make_int4(valid ? 1 : 0, seq, tt, orow);
            __syncwarp();
#pragma unroll 1
            for (int c0 = half * 16; c0 < BN; c0 += 32) {
                if (n0 + c0 >= p.N) break;                   // warp-uniform
                uint32_t v[16];
                {
                    uint32_t v2[16];
                    tmem_ld16(tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, v);
                    if (!BF16) {
                        tmem_ld16(tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)(BN + c0), v2);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
                    }
                }
                if (p.transposed) {                           // [B][N][L] output: consecutive rows are consecutive addresses already
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
                                if (p.resid) x += p.resid[(size_t)orow * p.ldr + n];
                                p.C[((size_t)seq * p.N + n) * p.L + tt] = x;
                            }
                        }
                    }
                    continue;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) ep[lane * 17 + j] = __uint_as_float(v[j]);
                __syncwarp();
                const int j = lane & 15, n = n0 + c0 + j;
                if (n < p.N) {
                    const float bias = p.bias ? __ldg(p.bias + n) : 0.f;
                    const float aw = p.act_w ? __ldg(p.act_w + n) : 1.f;
                    int l = n * p.cstride + p.coff;
                    if (p.chalf > 0 && l >= p.chalf) l = l - p.chalf + p.chp;
                    float* __restrict__ Cn = p.C + l;
                    const bool plain = !p.addrow && !p.addpos && !p.resid;
#pragma unroll 8
                    for (int rr = 0; rr < 16; ++rr) {
                        const int r = 2 * rr + (lane >> 4);
                        const int4 ri = rinfo[r];
                        float x = ep[r * 17 + j] + bias;
                        if (plain) {
                            x = apply_act(x, p.act, aw);
                        } else if (ri.x) {
                            if (p.addrow) x += p.addrow[(size_t)ri.y * p.N + n];
                            x = apply_act(x, p.act, aw);
                            if (p.addpos) x += __ldg(p.addpos + (size_t)ri.z * p.ldpos + n);
                            if (p.resid) x += p.resid[(size_t)ri.w * p.ldr + n];
                        }
                        if (ri.x) Cn[(size_t)ri.w * p.ldc] = x;
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(&tempty[buf]);                        // accumulator may be overwritten (all 256 epilogue threads arrive)
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS) : "memory");
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

// 2-D bf16 tensor map: inner dimension `cols` bf16 (contiguous), box = [64 cols = 128 bytes][box_rows].
inline bool make_map_2d_bf16(CUtensorMap* map, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
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
    static bool attr_set = false;                          // once per process (one device per process): not on every launch
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<BN, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<BN>::TOTAL);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    tc_gemm_kernel<BN, G><<<grid, TC_THREADS_P, TcSmem<BN>::TOTAL, s>>>(mA, mWh, mWl, p);
    return cudaGetLastError();
}

inline const char* launch_tc_gemm(const TcOperands& o, TcParams p, cudaStream_t s, bool gather = false) {
    CUtensorMap mA, mWh, mWl;
    const int BN = (p.N <= 32) ? 32 : (p.N <= 64 ? 64 : 128);
    const int rem = p.Kc % TC_BK;
    p.ksteps_last = rem == 0 ? TC_BK / 8 : (rem + 7) / 8;
    p.ga_A = o.A; p.ga_lda = o.lda; p.ga_rows = o.a_rows;
#ifdef L2S_DEBUG
    if (const char* e = getenv("L2S_TC_DEBUG_NITER")) p.debug_niter = atoi(e);
    if (const char* e = getenv("L2S_TC_DEBUG_SKIP")) p.debug_skip = atoi(e);
#endif
    if (gather) {
        if ((o.lda & 3) || (p.Kc & 3) || (reinterpret_cast<uintptr_t>(o.A) & 15)) return "gather-A needs 16-byte aligned rows";
        // the A map is unused in gather mode but must be a valid object: describe the W matrix again
        if (!make_map_2d(&mA, o.Whi, o.w_cols, p.N, o.w_cols, BN)) return "cuTensorMapEncodeTiled failed";
    } else if (!make_map_2d(&mA, o.A, o.a_cols, o.a_rows, o.lda, TC_BM)) return "cuTensorMapEncodeTiled(A) failed";
    if (!make_map_2d(&mWh, o.Whi, o.w_cols, p.N, o.w_cols, BN)) return "cuTensorMapEncodeTiled(Whi) failed";
    if (!make_map_2d(&mWl, o.Wlo, o.w_cols, p.N, o.w_cols, BN)) return "cuTensorMapEncodeTiled(Wlo) failed";
    static int num_sms = 0;
    if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); }
    const int ntiles = ceil_div(p.M, TC_BM) * ceil_div(p.N, BN);
    dim3 grid(ntiles < num_sms ? ntiles : num_sms);
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

// The Conv3d stem with bf16 operands: A = space-to-depth rows of 16 bf16 per position (tap windows of 64 bf16 gathered by
// cp.async), W = bf16 [32][taps*64] (rows >= N zero), N <= 32.  p.tap_shift / p.stem* describe the geometry as in the
// fp32 path.
inline const char* launch_tc_stem_bf16(const void* A16, const void* W16, int taps, TcParams p, cudaStream_t s) {
    CUtensorMap mW;
    if (p.N > 32) return "bf16 stem needs N <= 32";
    if (reinterpret_cast<uintptr_t>(A16) & 15) return "bf16 stem needs 16-byte aligned rows";
    if (!make_map_2d_bf16(&mW, W16, (uint64_t)taps * 64, 32, (uint64_t)taps * 64, 32)) return "cuTensorMapEncodeTiled(W bf16) failed";
    p.ga_A = reinterpret_cast<const float*>(A16); p.taps = taps; p.use_shift_table = 1; p.ksteps_last = 4;
#ifdef L2S_DEBUG
    if (const char* e = getenv("L2S_TC_DEBUG_SKIP")) p.debug_skip = atoi(e);
#endif
    static int num_sms = 0;
    if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); }
    const int ntiles = ceil_div(p.M, TC_BM);
    dim3 grid(ntiles < num_sms ? ntiles : num_sms);
    static bool attr_set = false;
    cudaError_t e = cudaSuccess;
    if (!attr_set) {
        e = cudaFuncSetAttribute(tc_gemm_kernel<32, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<32>::TOTAL);
        if (e != cudaSuccess) return cudaGetErrorString(e);
        attr_set = true;
    }
    tc_gemm_kernel<32, true, true><<<grid, TC_THREADS_P, TcSmem<32>::TOTAL, s>>>(mW, mW, mW, p);
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace l2s
