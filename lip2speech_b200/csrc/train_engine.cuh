// Train-mode engine: a reverse-mode tape over hand-written CUDA kernels (forward-with-saved-activations + backward) for the
// layers of Decoder.forward / VideoExtractor.forward in train() mode (reference decoder.py:320-379, video.py:76-87,
// shufflenetv2.py:42-104; autograd at train.py:184).
//
// Design: every activation is a 2-D row-major view [rows][cols] with a row stride (channels last: a Conv1d activation
// [B,C,L] lives as rows (b,l) x C columns; an NHWC frame as rows (n,h,w) x C), values in one bump arena, gradients in a second
// arena that is zero-filled once per backward.  An op runs its forward kernel(s) and pushes a closure with its backward
// kernels on the tape; `backward()` replays the tape in reverse.  Parameters are NOT copied: the library reads the caller's
// (PyTorch-owned) fp32 parameter memory and accumulates into the caller's gradient memory (l2s_train_bind), so an optimizer
// step needs no re-bind / re-pack.  All reductions have a fixed order (no floating-point atomics): results are
// run-to-run deterministic.  fp32 FMA arithmetic; the large GEMMs of the decoder run as 3xTF32 tensor-core products with fp32
// accumulation, those of the video frontend stay on FMAs (sgemm_kernel says why).
//
// Speed: a train step is ~5 K small dependent launches (8 clips per GPU: every per-step layer is a few-row GEMM).  The arenas hand
// out the same addresses for the same shapes, so the forward and the backward launch sequences of a (B, T, M) key are captured
// once as CUDA graphs and replayed (GraphSlot below; train_model.cuh stages the caller's tensors so that nothing a graph points at
// moves).  The kernels that matter are sized for bytes in flight rather than FLOPs: skinny_nt_smem / skinny_nn_strip (few-row
// GEMMs against L2-resident weights), sgemm_tn_rows (weight AND bias gradients of per-step layers gathered over all steps through
// a row-pointer list), attn_step_* and psine_chain_* (one launch per attention / activation chain of a decoder step).
#pragma once
#include <functional>
#include <string>
#include <vector>

#include "common.cuh"
#include "context.h"
#include "matvec.cuh"

namespace l2s {
namespace tr {

// ---- launches ---------------------------------------------------------------------------------------------------------------------
// A train step is a chain of ~5 K small DEPENDENT kernels; what separates two of them is mostly launch latency.  Every kernel of
// this file therefore starts with griddepcontrol.wait and is launched with programmatic stream serialization (programmatic
// dependent launch): the next kernel's CTAs are scheduled while the previous kernel drains, and block at the wait until that grid
// has completed and its memory is visible — the semantics of a plain stream dependency, without the gap.  Right after its own wait
// a kernel lets its dependents launch (they do nothing before THEIR wait, so nothing runs ahead of its inputs; one kernel deep).
// The attribute is also honoured inside captured graphs (programmatic edges).
#define TR_PDL_WAIT() asm volatile("griddepcontrol.wait;\n\tgriddepcontrol.launch_dependents;" ::: "memory")
inline bool& pdl_enabled() { static bool on = true; return on; }
template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    static_assert(sizeof...(KArgs) == sizeof...(Args), "kernel argument count");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    if (e != cudaSuccess) throw L2sError(2, std::string("train: kernel launch: ") + cudaGetErrorString(e));
}

struct TT {                       // train tensor: [rows][cols] view, element (r, c) at v[r*rs + c]
    float* v = nullptr;
    float* g = nullptr;           // gradient with the same layout (null: no gradient wanted)
    int rows = 0, cols = 0, rs = 0;
    size_t numel() const { return (size_t)rows * cols; }
    bool dense() const { return rs == cols; }
    TT colslice(int c0, int n) const { TT t = *this; t.v += c0; if (t.g) t.g += c0; t.cols = n; return t; }
    // rows r0, r0+step, ... (n of them)
    TT rowslice(int r0, int n, int step = 1) const {
        TT t = *this; t.v += (size_t)r0 * rs; if (t.g) t.g += (size_t)r0 * rs; t.rows = n; t.rs = rs * step; return t;
    }
};

// ---- kernels: GEMM ------------------------------------------------------------------------------------------------------
// C[M,N] (+)= A' B'  with A' = A (TA=0: A[m*lda + k]) or A^T (TA=1: A[k*lda + m]); B' = B (TB=0: B[k*ldb + n]) or B^T (TB=1: B[n*ldb + k]).
// 64x64 output tile per CTA, 32-deep k tiles fetched one tile ahead into registers; 8 warps = 4 (m16) x 2 (32 columns = four n8
// tiles) run mma.sync.m16n8k8 with the 3xTF32 split done on the fragments (a_lo b_hi + a_hi b_lo + a_hi b_hi, fp32 accumulate:
// the accuracy class of the inference path's GEMMs, ~2^-21 per product).  The tiles sit in shared memory in the layout in which
// BOTH the coalesced global->shared stores and the fragment loads are bank-conflict-free: the contiguous index of the operand
// in memory stays contiguous (row stride 36 when that is k, 72 when it is m / n).
// gridDim.z > 1: split over K — block z reduces K range [z*kper, (z+1)*kper) into its own [M][N] slab at C + z*M*ldc (the caller
// sums the slabs in order: a tall-skinny weight gradient, M x N small and K = tens of thousands of rows, would otherwise run on
// one or four CTAs).
// EXACT: the same tiles multiplied with fp32 FMAs (4x4 outputs per thread) instead of tensor cores.  The video frontend needs it:
// its BatchNorm + ReLU stacks on a small batch amplify a 2^-21 product error into flipped ReLU decisions, and single entries of
// some gradients then move by percents (tests/test_train_model_gpu.py: _check_against_fp64 measures exactly that).
template <int TA, int TB, bool EXACT>
__global__ void __launch_bounds__(256) sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                    float* __restrict__ C, int ldc, int accumulate, int kper) {
    TR_PDL_WAIT();
    constexpr int BM = 64, BN = 64, BK = 32;
    // EXACT keeps BOTH tiles k-major with a row stride of 68 floats: a thread then reads its 4 consecutive rows / columns of a k
    // step as one LDS.128 each (2 shared loads per 16 FMAs); the transposing fill of a k-contiguous operand pays a 4-way conflict.
    constexpr bool A_KMAJOR = EXACT || TA == 1, B_KMAJOR = EXACT || TB == 0;
    constexpr int A_LD = EXACT ? BM + 4 : (TA == 0 ? BK + 4 : BM + 8), B_LD = EXACT ? BN + 4 : (TB == 0 ? BN + 8 : BK + 4);
    __shared__ __align__(16) float As[BM * (BK + 4)];                  // every layout fits in 2304 floats
    __shared__ __align__(16) float Bs[BN * (BK + 4)];
    static_assert(BM * (BK + 4) == BK * (BM + 8) && BN * (BK + 4) == BK * (BN + 8), "tile layouts");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3, wm = warp & 3, wn = warp >> 2;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * kper;
    if (gridDim.z > 1) { C += (size_t)blockIdx.z * M * ldc; K = min(K, kbeg + kper); }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float ra[8], rb[8];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + i * 256;
            {
                int m, k;
                if (TA == 0) { k = e & 31; m = e >> 5; } else { m = e & 63; k = e >> 6; }     // contiguous index fastest
                const int gm = m0 + m, gk = k0 + k;
                ra[i] = (gm < M && gk < K) ? (TA == 0 ? A[(size_t)gm * lda + gk] : A[(size_t)gk * lda + gm]) : 0.f;
            }
            {
                int n, k;
                if (TB == 0) { n = e & 63; k = e >> 6; } else { k = e & 31; n = e >> 5; }
                const int gn = n0 + n, gk = k0 + k;
                rb[i] = (gn < N && gk < K) ? (TB == 0 ? B[(size_t)gk * ldb + gn] : B[(size_t)gn * ldb + gk]) : 0.f;
            }
        }
    };
    auto a_at = [&](int m, int k) -> float { return A_KMAJOR ? As[k * A_LD + m] : As[m * A_LD + k]; };
    auto b_at = [&](int k, int n) -> float { return B_KMAJOR ? Bs[k * B_LD + n] : Bs[n * B_LD + k]; };
    fetch(kbeg);
    for (int k0 = kbeg; k0 < K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + i * 256;
            {   // (m, k) of element e as fetched; stored at the layout's address
                const int m = TA == 0 ? e >> 5 : e & 63, k = TA == 0 ? e & 31 : e >> 6;
                As[A_KMAJOR ? k * A_LD + m : m * A_LD + k] = ra[i];
            }
            {
                const int n = TB == 0 ? e & 63 : e >> 5, k = TB == 0 ? e >> 6 : e & 31;
                Bs[B_KMAJOR ? k * B_LD + n : n * B_LD + k] = rb[i];
            }
        }
        __syncthreads();
        if (k0 + BK < K) fetch(k0 + BK);
        if (EXACT) {
            const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const float4 a4 = *reinterpret_cast<const float4*>(As + k * A_LD + 4 * ty);
                const float4 b4 = *reinterpret_cast<const float4*>(Bs + k * B_LD + 4 * tx);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        } else
#pragma unroll
        for (int ks = 0; ks < BK; ks += 8) {
            uint32_t ah[4], al[4];
            split_tf32(a_at(wm * 16 + g, ks + t), ah[0], al[0]);
            split_tf32(a_at(wm * 16 + g + 8, ks + t), ah[1], al[1]);
            split_tf32(a_at(wm * 16 + g, ks + t + 4), ah[2], al[2]);
            split_tf32(a_at(wm * 16 + g + 8, ks + t + 4), ah[3], al[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t bh[2], bl[2];
                const int n = wn * 32 + j * 8 + g;
                split_tf32(b_at(ks + t, n), bh[0], bl[0]);
                split_tf32(b_at(ks + t + 4, n), bh[1], bl[1]);
                mma_tf32(acc[j], al[0], al[1], al[2], al[3], bh[0], bh[1]);
                mma_tf32(acc[j], ah[0], ah[1], ah[2], ah[3], bl[0], bl[1]);
                mma_tf32(acc[j], ah[0], ah[1], ah[2], ah[3], bh[0], bh[1]);
            }
        }
        __syncthreads();
    }
    // tensor cores: acc[j] = {(r, c), (r, c+1), (r+8, c), (r+8, c+1)}, r = m0 + 16 wm + g, c = n0 + 32 wn + 8 j + 2 t;
    // EXACT: acc[i][j] = (m0 + 4 ty + i, n0 + 4 tx + j)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int m = EXACT ? m0 + 4 * (tid >> 4) + j : m0 + wm * 16 + g + 8 * (q >> 1);
            const int n = EXACT ? n0 + 4 * (tid & 15) + q : n0 + wn * 32 + j * 8 + 2 * t + (q & 1);
            if (m < M && n < N) {
                float* c = C + (size_t)m * ldc + n;
                *c = accumulate ? *c + acc[j][q] : acc[j][q];
            }
        }
}

// Few-row GEMM (M <= 16: the per-step layers of the decode loop at a per-GPU batch of 8): C[m][n] = sum_k A[m][k] W[n][k] (+bias),
// one warp per output column n, lanes split k; the weight row is read once and reused for every m.  VEC: rows are 16-byte
// aligned and K % 4 == 0 -> each lane takes four consecutive k per step (LDG.128 for W and A: 4x fewer, 4x wider requests
// in flight; the scalar form spent its time waiting for one 4-byte L2 load per iteration).
template <bool VEC>
__global__ void __launch_bounds__(256) skinny_nt_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                                                        const float* __restrict__ bias, float* __restrict__ C, int ldc, int accumulate) {
    TR_PDL_WAIT();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= N) return;
    float acc[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) acc[m] = 0.f;
    const float* w = W + (size_t)warp * ldw;
    if (VEC) {
#pragma unroll 2
        for (int k = 4 * lane; k < K; k += 128) {
            const float4 wv = *reinterpret_cast<const float4*>(w + k);
#pragma unroll
            for (int m = 0; m < 16; ++m)
                if (m < M) {
                    const float4 a = *reinterpret_cast<const float4*>(A + (size_t)m * lda + k);
                    acc[m] = fmaf(a.x, wv.x, acc[m]); acc[m] = fmaf(a.y, wv.y, acc[m]);
                    acc[m] = fmaf(a.z, wv.z, acc[m]); acc[m] = fmaf(a.w, wv.w, acc[m]);
                }
        }
    } else {
        for (int k = lane; k < K; k += 32) {
            const float wv = w[k];
#pragma unroll
            for (int m = 0; m < 16; ++m)
                if (m < M) acc[m] = fmaf(A[(size_t)m * lda + k], wv, acc[m]);
        }
    }
#pragma unroll
    for (int m = 0; m < 16; ++m)
        if (m < M) {
            const float v = warp_sum(acc[m]);
            if (lane == 0) {
                float* c = C + (size_t)m * ldc + warp;
                const float r = v + (bias ? bias[warp] : 0.f);
                *c = accumulate ? *c + r : r;
            }
        }
}
// Cross-lane sums of MANY per-lane partial values without one butterfly per value: each round swaps half of the remaining values
// with the partner lane (offset `off`) and adds, so V values cost V - 1 shuffles in total instead of 5 V.
#define L2S_FOLD_LANES(v, off, cnt)                                                   \
    {                                                                                 \
        const bool up_ = (lane & (off)) != 0;                                         \
        _Pragma("unroll") for (int i_ = 0; i_ < (cnt); ++i_) {                        \
            const float send_ = up_ ? v[i_] : v[i_ + (cnt)];                          \
            const float keep_ = up_ ? v[i_ + (cnt)] : v[i_];                          \
            v[i_] = keep_ + __shfl_xor_sync(0xffffffffu, send_, (off));               \
        }                                                                             \
    }
// Folds V per-lane partial values (V = 8, 16 or 32) across the warp; afterwards lane l holds the complete sum of value
// l / (32 / V) (every lane of that group of 32 / V lanes holds it).
template <int V>
__device__ __forceinline__ float warp_fold(float (&v)[V], int lane) {
    if constexpr (V == 32) {
        L2S_FOLD_LANES(v, 16, 16) L2S_FOLD_LANES(v, 8, 8) L2S_FOLD_LANES(v, 4, 4) L2S_FOLD_LANES(v, 2, 2) L2S_FOLD_LANES(v, 1, 1)
    } else if constexpr (V == 16) {
        L2S_FOLD_LANES(v, 16, 8) L2S_FOLD_LANES(v, 8, 4) L2S_FOLD_LANES(v, 4, 2) L2S_FOLD_LANES(v, 2, 1)
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    } else {
        static_assert(V == 8, "8, 16 or 32 values");
        L2S_FOLD_LANES(v, 16, 4) L2S_FOLD_LANES(v, 8, 2) L2S_FOLD_LANES(v, 4, 1)
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    }
    return v[0];
}
// Few-row GEMM, second form: the CTA first stages ALL of A (M x K, M <= MT) in shared memory, each warp owns NW weight rows and
// issues every 16-byte weight load of a 512- or 1024-wide k chunk before it touches them (the LSTM weights of a step are in
// flight at once instead of two loads per warp); the NW * MT accumulators per lane are reduced with warp_fold.
// A second operand pair (A2 [M,K2], W2 [N,K2], bias2) is treated as a continuation of the reduction: C = A W^T + A2 W2^T + bias +
// bias2 in one launch (the two halves of the LSTM gate pre-activation, x W_ih^T + h W_hh^T).  K2 = 0: single product.
template <int NW, int MT>
__global__ void __launch_bounds__(256) skinny_nt_smem_kernel(int M, int N, int K1, const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                                                             const float* __restrict__ bias, int K2, const float* __restrict__ A2, int lda2,
                                                             const float* __restrict__ W2, int ldw2, const float* __restrict__ bias2,
                                                             float* __restrict__ C, int ldc, int accumulate) {
    TR_PDL_WAIT();
    const int K = K1 + K2;
    constexpr int V = NW * MT;                                        // accumulators per lane
    constexpr int J = NW * MT <= 16 ? 8 : 4;                          // 16-byte weight loads per row kept in flight: a chunk of 128 J columns
    extern __shared__ float4 xs4[];                                   // [MT][K/4], rows >= M are zero
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = (blockIdx.x * 8 + warp) * NW;
    const int K4 = K >> 2;
    float4 wv[NW][J];
    auto load_w = [&](int kc) {
#pragma unroll
        for (int n = 0; n < NW; ++n)
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int k = kc + 128 * j + 4 * lane;
                const float* src = k < K1 ? W + (size_t)(n0 + n) * ldw + k : W2 + (size_t)(n0 + n) * ldw2 + (k - K1);
                wv[n][j] = (n0 + n < N && k < K) ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
    };
    load_w(0);
    for (int i = threadIdx.x; i < MT * K4; i += 256) {
        const int m = i / K4, q = i - m * K4;
        const float* src = 4 * q < K1 ? A + (size_t)m * lda + 4 * q : A2 + (size_t)m * lda2 + (4 * q - K1);
        xs4[i] = m < M ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    for (int kc = 0; kc < K; kc += 128 * J) {
        if (kc) load_w(kc);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int q = (kc >> 2) + 32 * j + lane;
            if (q < K4) {
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const float4 a = xs4[m * K4 + q];
#pragma unroll
                    for (int n = 0; n < NW; ++n) {
                        float t = acc[n * MT + m];
                        t = fmaf(a.x, wv[n][j].x, t); t = fmaf(a.y, wv[n][j].y, t); t = fmaf(a.z, wv[n][j].z, t); t = fmaf(a.w, wv[n][j].w, t);
                        acc[n * MT + m] = t;
                    }
                }
            }
        }
    }
    const float v = warp_fold<V>(acc, lane);                          // value index = n * MT + m
    const int idx = lane / (32 / V);
    const int n = n0 + idx / MT, m = idx % MT;
    if ((lane & (32 / V - 1)) == 0 && n < N && m < M) {
        float* c = C + (size_t)m * ldc + n;
        const float r = v + (bias ? bias[n] : 0.f) + (bias2 ? bias2[n] : 0.f);
        *c = accumulate ? *c + r : r;
    }
}

// dx[m][k] += sum_n dy[m][n] W[n][k] for M <= MT in ONE launch and without partial sums in memory: a CTA of 16 warps owns a
// strip of 8 columns k (32 bytes = one sector of every weight row) and walks ALL N rows: lane = (2 k-quads) x (16 rows per load
// instruction), the 16 warps interleave the rows, so N = 2048 is 8 loads per lane and ALL of them are issued before the first
// is used (the kernel is bound by bytes in flight: 4 MB of L2-resident weights, ~0.1 MFLOP).  dy sits in shared memory in its
// natural [m][n] layout (16-byte copies in, conflict-free scalar reads out).  The 16 row-lanes are folded with shuffles, the
// 16 warps through shared memory in index order: deterministic.
// gridDim.y = 2: a second problem (W2, K2, dX2) that shares dy (the two operands of one pre-activation).
constexpr int NN_THREADS = 512;
template <int MT>
__global__ void __launch_bounds__(NN_THREADS) skinny_nn_strip_kernel(int M, int N, int K, const float* __restrict__ dY, int ldy, const float* __restrict__ W, int ldw,
                                                                     float* __restrict__ dX, int ldx, int K2, const float* __restrict__ W2, int ldw2,
                                                                     float* __restrict__ dX2, int ldx2) {
    TR_PDL_WAIT();
    if (blockIdx.y == 1) { K = K2; W = W2; ldw = ldw2; dX = dX2; ldx = ldx2; }
    if ((int)blockIdx.x * 8 >= K) return;
    extern __shared__ float4 nn_sm4[];
    float* dys = reinterpret_cast<float*>(nn_sm4);                   // [MT][Npad]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kq = lane & 1, ny = lane >> 1;
    const int k = blockIdx.x * 8 + kq * 4;
    const bool kok = k < K;
    const int nfirst = warp * 16 + ny;                               // this lane's rows: nfirst + 256 j
    const int iters = (N + 255) / 256;
    const int Npad = iters * 256;
    float4 wv[8];
    auto fetch = [&](int j0) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int n = nfirst + 256 * (j0 + u);
            wv[u] = (kok && n < N) ? *reinterpret_cast<const float4*>(W + (size_t)n * ldw + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    fetch(0);
    if (!(ldy & 3) && !(reinterpret_cast<uintptr_t>(dY) & 15)) {
        const int N4 = Npad >> 2;
        for (int i = threadIdx.x; i < MT * N4; i += NN_THREADS) {
            const int m = i / N4, q = i - m * N4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < M && 4 * q + 3 < N) v = *reinterpret_cast<const float4*>(dY + (size_t)m * ldy + 4 * q);
            else if (m < M) {
                float t[4] = {0.f, 0.f, 0.f, 0.f};
                for (int u = 0; u < 4; ++u) if (4 * q + u < N) t[u] = dY[(size_t)m * ldy + 4 * q + u];
                v = make_float4(t[0], t[1], t[2], t[3]);
            }
            nn_sm4[i] = v;
        }
    } else {
        for (int i = threadIdx.x; i < MT * Npad; i += NN_THREADS) {
            const int m = i / Npad, n = i - m * Npad;
            dys[i] = (m < M && n < N) ? dY[(size_t)m * ldy + n] : 0.f;
        }
    }
    __syncthreads();
    float acc[MT * 4];
#pragma unroll
    for (int i = 0; i < MT * 4; ++i) acc[i] = 0.f;
    for (int j0 = 0; j0 < iters; j0 += 8) {
        if (j0) fetch(j0);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (j0 + u >= iters) break;
            const float* d = dys + nfirst + 256 * (j0 + u);
            const float4 w4 = wv[u];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const float dm = d[(size_t)m * Npad];
                float* a = acc + m * 4;
                a[0] = fmaf(dm, w4.x, a[0]); a[1] = fmaf(dm, w4.y, a[1]); a[2] = fmaf(dm, w4.z, a[2]); a[3] = fmaf(dm, w4.w, a[3]);
            }
        }
    }
    // fold the 16 row-lanes (lane bits 4..1): each round hands half of the remaining values to the partner
    L2S_FOLD_LANES(acc, 16, MT * 2) L2S_FOLD_LANES(acc, 8, MT) L2S_FOLD_LANES(acc, 4, MT / 2) L2S_FOLD_LANES(acc, 2, MT / 4)
    // MT/4 values are left: value index v = (MT*2) b4 + MT b3 + (MT/2) b2 + (MT/4) b1 + [0, MT/4), row v / 4, column v % 4
    __syncthreads();                                                 // dys is dead: reuse it for the cross-warp sums
    float* red = dys;                                                // [16 warps][MT][8]
    const int vbase = (MT * 2) * ((lane >> 4) & 1) + MT * ((lane >> 3) & 1) + (MT / 2) * ((lane >> 2) & 1) + (MT / 4) * ((lane >> 1) & 1);
#pragma unroll
    for (int u = 0; u < MT / 4; ++u) {
        const int v = vbase + u;
        red[((size_t)warp * MT + (v >> 2)) * 8 + kq * 4 + (v & 3)] = acc[u];
    }
    __syncthreads();
    if (threadIdx.x < MT * 8) {
        const int m = threadIdx.x >> 3, kk = blockIdx.x * 8 + (threadIdx.x & 7);
        if (m < M && kk < K) {
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < NN_THREADS / 32; ++w) sum += red[((size_t)w * MT + m) * 8 + (threadIdx.x & 7)];
            dX[(size_t)m * ldx + kk] += sum;
        }
    }
}

// dx[m][k] (+)= sum_n dy[m][n] W[n][k]  for M <= 16, in two deterministic stages so that the whole chip streams W once:
// stage 1: block (64 columns k) x (slice of `nslice` rows n), 256 threads = 64 kx x 4 ny -> part[slice][m][k];
// stage 2: dX (+)= sum over the slices in order.
__global__ void __launch_bounds__(256) skinny_nn_part_kernel(int M, int N, int K, int nslice, const float* __restrict__ dY, int ldy,
                                                             const float* __restrict__ W, int ldw, float* __restrict__ part) {
    TR_PDL_WAIT();
    __shared__ float red[4][16][65];
    const int kx = threadIdx.x & 63, ny = threadIdx.x >> 6;
    const int k = blockIdx.x * 64 + kx;
    const int n0 = blockIdx.y * nslice, n1 = min(N, n0 + nslice);
    float acc[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) acc[m] = 0.f;
    if (k < K)
#pragma unroll 4
        for (int n = n0 + ny; n < n1; n += 4) {
            const float wv = W[(size_t)n * ldw + k];
#pragma unroll
            for (int m = 0; m < 16; ++m)
                if (m < M) acc[m] = fmaf(__ldg(dY + (size_t)m * ldy + n), wv, acc[m]);
        }
#pragma unroll
    for (int m = 0; m < 16; ++m) red[ny][m][kx] = acc[m];
    __syncthreads();
    if (ny == 0 && k < K)
        for (int m = 0; m < M; ++m)
            part[((size_t)blockIdx.y * M + m) * K + k] = (red[0][m][kx] + red[1][m][kx]) + (red[2][m][kx] + red[3][m][kx]);
}
__global__ void skinny_nn_sum_kernel(int M, int K, int nslices, const float* __restrict__ part, float* __restrict__ dX, int ldx, int accumulate) {
    TR_PDL_WAIT();
    const size_t total = (size_t)M * K;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int m = i / K, k = i % K;
        float a = 0.f;
        for (int z = 0; z < nslices; ++z) a += part[(size_t)z * total + i];
        float* d = dX + (size_t)m * ldx + k;
        *d = accumulate ? *d + a : a;
    }
}

// dW[n][k] += sum_i A_i[n] * B_i[k] over a LIST of row pairs (A_i = a gradient row, B_i = an input row): the weight gradient of
// a layer that ran once per decoder step / LSTM time step, gathered over all its invocations in ONE GEMM instead of one
// read-modify-write of dW per step (rowsA / rowsB: device arrays of row pointers, R entries).
// db (optional): the bias gradient db[n] += sum_i A_i[n], formed by the CTAs of the first column block from the tiles they load anyway.
__global__ void __launch_bounds__(256) sgemm_tn_rows_kernel(int N, int K, int R, const float* const* __restrict__ rowsA, const float* const* __restrict__ rowsB,
                                                            float* __restrict__ C, int ldc, float* __restrict__ db) {
    TR_PDL_WAIT();
    constexpr int BM = 64, BN = 64, BK = 32, LD = 72;                  // both tiles k-major [BK][64 + 8]: see sgemm_kernel
    __shared__ float As[BK * LD];
    __shared__ float Bs[BK * LD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3, wm = warp & 3, wn = warp >> 2;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const bool sums = db != nullptr && blockIdx.x == 0 && tid < 64;
    float bsum = 0.f;
    // 32 list rows per tile, 8 elements of each operand per thread, fetched one tile ahead (pointer + data: two dependent loads)
    float ra[8], rb[8];
    auto fetch = [&](int r0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + i * 256;
            const int c = e & 63, gr = r0 + (e >> 6);
            ra[i] = 0.f; rb[i] = 0.f;
            if (gr < R) {
                if (m0 + c < N) ra[i] = rowsA[gr][m0 + c];
                if (n0 + c < K) rb[i] = rowsB[gr][n0 + c];
            }
        }
    };
    fetch(0);
    for (int r0 = 0; r0 < R; r0 += BK) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + i * 256;
            As[(e >> 6) * LD + (e & 63)] = ra[i]; Bs[(e >> 6) * LD + (e & 63)] = rb[i];
        }
        __syncthreads();
        if (r0 + BK < R) fetch(r0 + BK);
        if (sums)
#pragma unroll
            for (int k = 0; k < BK; ++k) bsum += As[k * LD + tid];
#pragma unroll
        for (int ks = 0; ks < BK; ks += 8) {
            uint32_t ah[4], al[4];
            split_tf32(As[(ks + t) * LD + wm * 16 + g], ah[0], al[0]);
            split_tf32(As[(ks + t) * LD + wm * 16 + g + 8], ah[1], al[1]);
            split_tf32(As[(ks + t + 4) * LD + wm * 16 + g], ah[2], al[2]);
            split_tf32(As[(ks + t + 4) * LD + wm * 16 + g + 8], ah[3], al[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t bh[2], bl[2];
                const int n = wn * 32 + j * 8 + g;
                split_tf32(Bs[(ks + t) * LD + n], bh[0], bl[0]);
                split_tf32(Bs[(ks + t + 4) * LD + n], bh[1], bl[1]);
                mma_tf32(acc[j], al[0], al[1], al[2], al[3], bh[0], bh[1]);
                mma_tf32(acc[j], ah[0], ah[1], ah[2], ah[3], bl[0], bl[1]);
                mma_tf32(acc[j], ah[0], ah[1], ah[2], ah[3], bh[0], bh[1]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int m = m0 + wm * 16 + g + 8 * (q >> 1), n = n0 + wn * 32 + j * 8 + 2 * t + (q & 1);
            if (m < N && n < K) C[(size_t)m * ldc + n] += acc[j][q];
        }
    if (sums && m0 + tid < N) db[m0 + tid] += bsum;
}

// ---- kernels: column reductions ---------------------------------------------------------------------------------------
// out[c] (+)= sum_r f(r, c).  Grid (ceil(cols/32), row splits): each CTA reduces its row range for 32 columns (8 row lanes x
// 32 columns, fixed order) into part[split][c]; colfinish_kernel adds the splits in order.  (Tall, narrow activations — the
// stem's 534 K rows x 24 channels — would otherwise be reduced by ONE CTA.)
enum ColOp { COL_SUM = 0, COL_SUM_XY = 1, COL_PSINE_DW = 2, COL_PRELU_DW = 3, COL_SQDEV = 4, COL_BN_DGAMMA = 5 };
template <int OP>
__global__ void __launch_bounds__(256) colreduce_kernel(int rows, int cols, const float* __restrict__ X, int xs, const float* __restrict__ Y, int ys,
                                                        const float* __restrict__ aux, const float* __restrict__ aux2, float eps, float* __restrict__ part) {
    TR_PDL_WAIT();
    __shared__ float p0[8][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const int per = (rows + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    float a = 0.f;
    if (c < cols) {
        const float mu = (OP == COL_SQDEV || OP == COL_BN_DGAMMA) ? aux[c] : 0.f;
        const float rstd = OP == COL_BN_DGAMMA ? rsqrtf(aux2[c] + eps) : 0.f;
#pragma unroll 8
        for (int r = r0 + rl; r < r1; r += 8) {
            const float x = X[(size_t)r * xs + c];
            if (OP == COL_SUM) a += x;
            else if (OP == COL_SUM_XY) a = fmaf(x, Y[(size_t)r * ys + c], a);
            else if (OP == COL_PSINE_DW) a = fmaf(sinf(x), Y[(size_t)r * ys + c], a);            // X = pre-activation, Y = dy
            else if (OP == COL_PRELU_DW) a += x < 0.f ? x * Y[(size_t)r * ys + c] : 0.f;
            else if (OP == COL_SQDEV) { const float d = x - mu; a = fmaf(d, d, a); }              // aux = mean
            else a = fmaf(Y[(size_t)r * ys + c], (x - mu) * rstd, a);                             // COL_BN_DGAMMA: aux = mean, aux2 = var, Y = dy
        }
    }
    p0[rl][cl] = a;
    __syncthreads();
    if (rl == 0 && c < cols) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += p0[i][cl];
        part[(size_t)blockIdx.y * cols + c] = s;
    }
}
// out[c] = (accumulate ? out[c] : 0) + scale * sum_split part[split][c]; one warp per column: lanes stride over the splits, then
// a fixed shuffle tree (deterministic for a given split count).
__global__ void colfinish_kernel(int cols, int splits, const float* __restrict__ part, float scale, float* __restrict__ out, int accumulate) {
    TR_PDL_WAIT();
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= cols) return;
    float s = 0.f;
    for (int z = lane; z < splits; z += 32) s += part[(size_t)z * cols + c];
    s = warp_sum(s) * scale;
    if (lane == 0) out[c] = accumulate ? out[c] + s : s;
}

// Full reduction sum(X*Y) -> out[0] (+)=, single CTA (used for the two scalar temperatures).
__global__ void __launch_bounds__(1024) dot_all_kernel(int rows, int cols, const float* __restrict__ X, int xs, const float* __restrict__ Y, int ys,
                                                       float* __restrict__ out, int accumulate) {
    TR_PDL_WAIT();
    __shared__ float part[32];
    float a = 0.f;
    const size_t n = (size_t)rows * cols;
    for (size_t i = threadIdx.x; i < n; i += 1024) { const int r = i / cols, c = i % cols; a = fmaf(X[(size_t)r * xs + c], Y[(size_t)r * ys + c], a); }
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        float s = warp_sum(part[threadIdx.x]);
        if (threadIdx.x == 0) out[0] = accumulate ? out[0] + s : s;
    }
}

// ---- kernels: elementwise ---------------------------------------------------------------------------------------------
#define TR_EW_LOOP(total) for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (total); i += (size_t)gridDim.x * blockDim.x)

enum EwOp { EW_COPY = 0, EW_ADD, EW_SILU, EW_RELU, EW_PSINE, EW_PRELU, EW_SCALE, EW_MASK, EW_ADDCONST, EW_ADDROW };
// y[r][c] = f(x[r][c], ...)   aux: per-column parameter (PSINE / PRELU), second operand (ADD / MASK / ADDCONST, row stride as),
// scalar alpha (SCALE / MASK).  ADDROW: aux is [rows/group][cols] broadcast over `group` consecutive rows.
template <int OP>
__global__ void ew_fwd_kernel(int rows, int cols, const float* X, int xs, const float* __restrict__ aux, int as, float alpha, int group,
                              float* Y, int ys) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)rows * cols) {
        const int r = i / cols, c = i % cols;
        const float x = X[(size_t)r * xs + c];
        float y;
        if (OP == EW_COPY) y = x;
        else if (OP == EW_ADD || OP == EW_ADDCONST) y = x + aux[(size_t)r * as + c];
        else if (OP == EW_SILU) y = siluf_acc(x);
        else if (OP == EW_RELU) y = x > 0.f ? x : 0.f;
        else if (OP == EW_PSINE) y = sinf(x) * aux[c];
        else if (OP == EW_PRELU) y = x >= 0.f ? x : aux[c] * x;
        else if (OP == EW_SCALE) y = x * alpha;
        else if (OP == EW_MASK) y = x * (aux[(size_t)r * as + c] * alpha);
        else y = x + aux[(size_t)(r / group) * as + c];      // EW_ADDROW
        Y[(size_t)r * ys + c] = y;
    }
}
// dX[r][c] += dY[r][c] * f'(...)   (X = the op's INPUT saved by the forward)
template <int OP>
__global__ void ew_bwd_kernel(int rows, int cols, const float* __restrict__ X, int xs, const float* __restrict__ aux, int as, float alpha,
                              const float* __restrict__ dY, int dys, float* __restrict__ dX, int dxs) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)rows * cols) {
        const int r = i / cols, c = i % cols;
        const float dy = dY[(size_t)r * dys + c];
        float d;
        if (OP == EW_COPY || OP == EW_ADD || OP == EW_ADDCONST || OP == EW_ADDROW) d = dy;
        else if (OP == EW_SILU) { const float x = X[(size_t)r * xs + c]; const float s = sigmoidf_acc(x); d = dy * (s * (1.f + x * (1.f - s))); }
        else if (OP == EW_RELU) d = X[(size_t)r * xs + c] > 0.f ? dy : 0.f;
        else if (OP == EW_PSINE) d = dy * cosf(X[(size_t)r * xs + c]) * aux[c];
        else if (OP == EW_PRELU) d = X[(size_t)r * xs + c] >= 0.f ? dy : aux[c] * dy;
        else if (OP == EW_SCALE) d = dy * alpha;
        else d = dy * (aux[(size_t)r * as + c] * alpha);      // EW_MASK
        dX[(size_t)r * dxs + c] += d;
    }
}
__global__ void select_fwd_kernel(int rows, int cols, const float* __restrict__ flag, const float* __restrict__ A, int as, const float* __restrict__ Bv, int bs,
                                  float* __restrict__ Y, int ys) {
    TR_PDL_WAIT();
    const bool pick_a = flag[0] != 0.f;
    TR_EW_LOOP((size_t)rows * cols) { const int r = i / cols, c = i % cols; Y[(size_t)r * ys + c] = pick_a ? A[(size_t)r * as + c] : Bv[(size_t)r * bs + c]; }
}
__global__ void select_bwd_kernel(int rows, int cols, const float* __restrict__ flag, const float* __restrict__ dY, int dys, float* __restrict__ dB, int dbs) {
    TR_PDL_WAIT();
    if (flag[0] != 0.f) return;
    TR_EW_LOOP((size_t)rows * cols) { const int r = i / cols, c = i % cols; dB[(size_t)r * dbs + c] += dY[(size_t)r * dys + c]; }
}
// dAux[g][c] += sum over the `group` rows of group g of dY   (backward of EW_ADDROW w.r.t. the broadcast operand)
__global__ void addrow_bwd_kernel(int groups, int group, int cols, const float* __restrict__ dY, int dys, float* __restrict__ dA, int das) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)groups * cols) {
        const int g = i / cols, c = i % cols;
        float a = 0.f;
        for (int j = 0; j < group; ++j) a += dY[(size_t)(g * group + j) * dys + c];
        dA[(size_t)g * das + c] += a;
    }
}

// y = x * w[0] (learnable scalar: the attention temperatures, decoder.py:302,237)
__global__ void scale_param_kernel(int rows, int cols, const float* __restrict__ X, int xs, const float* __restrict__ w, float* __restrict__ Y, int ys) {
    TR_PDL_WAIT();
    const float a = w[0];
    TR_EW_LOOP((size_t)rows * cols) { const int r = i / cols, c = i % cols; Y[(size_t)r * ys + c] = X[(size_t)r * xs + c] * a; }
}
__global__ void scale_param_bwd_kernel(int rows, int cols, const float* __restrict__ w, const float* __restrict__ dY, int dys, float* __restrict__ dX, int dxs) {
    TR_PDL_WAIT();
    const float a = w[0];
    TR_EW_LOOP((size_t)rows * cols) { const int r = i / cols, c = i % cols; dX[(size_t)r * dxs + c] += dY[(size_t)r * dys + c] * a; }
}

// BatchNorm (train): y = (x - mean) * rstd * gamma + beta
__global__ void bn_fwd_kernel(int rows, int cols, const float* __restrict__ X, int xs, const float* __restrict__ mean, const float* __restrict__ var, float eps,
                              const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ Y, int ys) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)rows * cols) {
        const int r = i / cols, c = i % cols;
        Y[(size_t)r * ys + c] = (X[(size_t)r * xs + c] - mean[c]) * rsqrtf(var[c] + eps) * gamma[c] + beta[c];
    }
}
// dx = gamma*rstd * (dy - dbeta/R - xhat * dgamma/R)
__global__ void bn_bwd_kernel(int rows, int cols, const float* __restrict__ X, int xs, const float* __restrict__ mean, const float* __restrict__ var, float eps,
                              const float* __restrict__ gamma, const float* __restrict__ dgamma, const float* __restrict__ dbeta,
                              const float* __restrict__ dY, int dys, float* __restrict__ dX, int dxs) {
    TR_PDL_WAIT();
    const float invR = 1.f / (float)rows;
    TR_EW_LOOP((size_t)rows * cols) {
        const int r = i / cols, c = i % cols;
        const float rstd = rsqrtf(var[c] + eps);
        const float xhat = (X[(size_t)r * xs + c] - mean[c]) * rstd;
        dX[(size_t)r * dxs + c] += gamma[c] * rstd * (dY[(size_t)r * dys + c] - dbeta[c] * invR - xhat * dgamma[c] * invR);
    }
}
// running = (1 - momentum) * running + momentum * stat  (variance: unbiased, x rows/(rows-1)); nn.BatchNorm*d in train()
__global__ void bn_running_kernel(int cols, int rows, float momentum, const float* __restrict__ mean, const float* __restrict__ var,
                                  float* __restrict__ rmean, float* __restrict__ rvar) {
    TR_PDL_WAIT();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean[c];
    const float unbiased = rows > 1 ? var[c] * ((float)rows / (float)(rows - 1)) : var[c];
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * unbiased;
}

// softmax over the columns of every row (one warp per row)
__global__ void __launch_bounds__(256) softmax_fwd_kernel(int rows, int cols, const float* __restrict__ X, int xs, float* __restrict__ Y, int ys) {
    TR_PDL_WAIT();
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* x = X + (size_t)r * xs;
    float mx = -INFINITY;
    for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, x[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += expf(x[c] - mx);
    s = warp_sum(s);
    for (int c = lane; c < cols; c += 32) Y[(size_t)r * ys + c] = expf(x[c] - mx) / s;
}
// dx += y * (dy - sum(dy*y))
__global__ void __launch_bounds__(256) softmax_bwd_kernel(int rows, int cols, const float* __restrict__ Y, int ys, const float* __restrict__ dY, int dys,
                                                          float* __restrict__ dX, int dxs) {
    TR_PDL_WAIT();
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= rows) return;
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s = fmaf(dY[(size_t)r * dys + c], Y[(size_t)r * ys + c], s);
    s = warp_sum(s);
    for (int c = lane; c < cols; c += 32) dX[(size_t)r * dxs + c] += Y[(size_t)r * ys + c] * (dY[(size_t)r * dys + c] - s);
}

// ---- kernels: per-clip attention (decoder.py:414-419, 262-271) ---------------------------------------------------------
// scores[b][t] = sum_k q[b][k] * Kmem[(b*T + t)][k]       (one warp per (b, t))
__global__ void __launch_bounds__(256) attn_scores_kernel(int B, int T, int D, const float* __restrict__ Q, int qs, const float* __restrict__ Km, int ks,
                                                          float* __restrict__ S, int ss) {
    TR_PDL_WAIT();
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= B * T) return;
    const int b = w / T, t = w % T;
    float a = 0.f;
    for (int k = lane; k < D; k += 32) a = fmaf(Q[(size_t)b * qs + k], Km[(size_t)(b * T + t) * ks + k], a);
    a = warp_sum(a);
    if (lane == 0) S[(size_t)b * ss + t] = a;
}
// dq[b][k] += sum_t dS[b][t] K[b,t,k] ;  dK[b,t,k] += dS[b][t] q[b][k]      (thread per (b, k))
__global__ void attn_scores_bwd_kernel(int B, int T, int D, const float* __restrict__ Q, int qs, const float* __restrict__ Km, int ks,
                                       const float* __restrict__ dS, int dss, float* __restrict__ dQ, int dqs, float* __restrict__ dK, int dks) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)B * D) {
        const int b = i / D, k = i % D;
        const float q = Q[(size_t)b * qs + k];
        float a = 0.f;
        for (int t = 0; t < T; ++t) {
            const float ds = dS[(size_t)b * dss + t];
            a = fmaf(ds, Km[(size_t)(b * T + t) * ks + k], a);
            if (dK) dK[(size_t)(b * T + t) * dks + k] += ds * q;
        }
        if (dQ) dQ[(size_t)b * dqs + k] += a;
    }
}
// ctx[b][k] = sum_t a[b][t] V[(b*T + t)][k]
__global__ void attn_context_kernel(int B, int T, int D, const float* __restrict__ A, int as, const float* __restrict__ V, int vs, float* __restrict__ C, int cs) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)B * D) {
        const int b = i / D, k = i % D;
        float acc = 0.f;
        for (int t = 0; t < T; ++t) acc = fmaf(A[(size_t)b * as + t], V[(size_t)(b * T + t) * vs + k], acc);
        C[(size_t)b * cs + k] = acc;
    }
}
// dA[b][t] += sum_k dC[b][k] V[b,t,k]  (warp per (b,t)) ; dV[b,t,k] += a[b][t] dC[b][k]
__global__ void __launch_bounds__(256) attn_context_bwd_kernel(int B, int T, int D, const float* __restrict__ A, int as, const float* __restrict__ V, int vs,
                                                               const float* __restrict__ dC, int dcs, float* __restrict__ dA, int das, float* __restrict__ dV, int dvs) {
    TR_PDL_WAIT();
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= B * T) return;
    const int b = w / T, t = w % T;
    const float a = A[(size_t)b * as + t];
    float s = 0.f;
    for (int k = lane; k < D; k += 32) {
        const float dc = dC[(size_t)b * dcs + k];
        s = fmaf(dc, V[(size_t)(b * T + t) * vs + k], s);
        if (dV) dV[(size_t)(b * T + t) * dvs + k] += a * dc;
    }
    s = warp_sum(s);
    if (lane == 0 && dA) dA[(size_t)b * das + t] += s;
}

// ---- kernels: one decoder-step attention in one launch -----------------------------------------------------------------------
// s_t = (w q_b) . K_{b,t}  ->  (x keep-mask / (1-p))  ->  softmax over t  ->  ctx_b = sum_t a_t V_{b,t}   (decoder.py:360-364 with
// the logit dropout, 262-271 without).  Grid (B, ATT_SPLIT): the CTAs of a clip each form the full score vector (K is small and
// L2-resident) and own one quarter of the feature columns of the read-out.  Saved for the backward pass: the probabilities a
// and the unscaled products q . K_t (the derivative with respect to the learnable temperature w).  logits (optional): the
// post-dropout scores, written straight into the caller-visible [B][M][T] tensor.  T <= 320; D, DV multiples of 4, rows 16-byte
// aligned (checked by the host).
constexpr int ATT_SPLIT = 4;
__global__ void __launch_bounds__(256) attn_step_fwd_kernel(int T, int D, int DV, const float* __restrict__ w, const float* __restrict__ Q, int qs,
                                                            const float* __restrict__ Km, int ks, const float* __restrict__ mask, float alpha,
                                                            const float* __restrict__ V, int vs, float* __restrict__ sraw, float* __restrict__ A,
                                                            float* __restrict__ logits, int ls, float* __restrict__ C, int cs) {
    TR_PDL_WAIT();
    __shared__ float sc[320];
    __shared__ float red[2];
    __shared__ float4 part[8][32];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool first = blockIdx.y == 0;
    const float wv = w[0];
    const float4* q4 = reinterpret_cast<const float4*>(Q + (size_t)b * qs);
    const int D4 = D >> 2;
    for (int t = warp; t < T; t += 8) {
        const float4* k4 = reinterpret_cast<const float4*>(Km + (size_t)(b * T + t) * ks);
        float a = 0.f, raw = 0.f;
#pragma unroll 4
        for (int k = lane; k < D4; k += 32) {
            const float4 qv = q4[k], kv = k4[k];
            a = fmaf(qv.x * wv, kv.x, a); a = fmaf(qv.y * wv, kv.y, a); a = fmaf(qv.z * wv, kv.z, a); a = fmaf(qv.w * wv, kv.w, a);
            raw = fmaf(qv.x, kv.x, raw); raw = fmaf(qv.y, kv.y, raw); raw = fmaf(qv.z, kv.z, raw); raw = fmaf(qv.w, kv.w, raw);
        }
        a = warp_sum(a); raw = warp_sum(raw);
        if (lane == 0) {
            if (mask) a *= mask[(size_t)b * T + t] * alpha;
            sc[t] = a;
            if (first) {
                sraw[(size_t)b * T + t] = raw;
                if (logits) logits[(size_t)b * ls + t] = a;
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        float mx = -INFINITY;
        for (int t = lane; t < T; t += 32) mx = fmaxf(mx, sc[t]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int t = lane; t < T; t += 32) sum += expf(sc[t] - mx);
        sum = warp_sum(sum);
        if (lane == 0) { red[0] = mx; red[1] = sum; }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += 256) {
        const float a = expf(sc[t] - red[0]) / red[1];
        sc[t] = a;
        if (first) A[(size_t)b * T + t] = a;
    }
    __syncthreads();
    // read-out: this CTA's quarter of the columns; 32 column quads x 8 position groups, folded in group order
    const int per = ((DV >> 2) + ATT_SPLIT - 1) / ATT_SPLIT;           // column quads per CTA (<= 32 for DV <= 512)
    for (int c0 = blockIdx.y * per; c0 < min((int)(blockIdx.y + 1) * per, DV >> 2); c0 += 32) {
        const int c = c0 + lane;
        const bool ok = c < min((int)(blockIdx.y + 1) * per, DV >> 2);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok)
#pragma unroll 4
            for (int t = warp; t < T; t += 8) {
                const float4 v = reinterpret_cast<const float4*>(V + (size_t)(b * T + t) * vs)[c];
                const float a = sc[t];
                acc.x = fmaf(a, v.x, acc.x); acc.y = fmaf(a, v.y, acc.y); acc.z = fmaf(a, v.z, acc.z); acc.w = fmaf(a, v.w, acc.w);
            }
        part[warp][lane] = acc;
        __syncthreads();
        if (warp == 0 && ok) {
            float4 r = part[0][lane];
#pragma unroll
            for (int g = 1; g < 8; ++g) { const float4 p4 = part[g][lane]; r.x += p4.x; r.y += p4.y; r.z += p4.z; r.w += p4.w; }
            reinterpret_cast<float4*>(C + (size_t)b * cs)[c] = r;
        }
        __syncthreads();
    }
}
// Backward of the above for clip b = blockIdx.x, column quarter blockIdx.y: every CTA forms da = V dC and the softmax / dropout
// backward for all positions (needs all columns: V is read once per CTA), then updates ITS columns of dV, dK and dq.  The dw partial
// (one float per clip, summed later in a fixed order) comes from the first CTA.
__global__ void __launch_bounds__(256) attn_step_bwd_kernel(int T, int D, int DV, const float* __restrict__ w, const float* __restrict__ Q, int qs,
                                                            const float* __restrict__ Km, int ks, const float* __restrict__ mask, float alpha,
                                                            const float* __restrict__ V, int vs, const float* __restrict__ sraw, const float* __restrict__ A,
                                                            const float* __restrict__ dC, int dcs, float* __restrict__ dQ, int dqs, float* __restrict__ dK, int dks,
                                                            float* __restrict__ dV, int dvs, float* __restrict__ dwpart) {
    TR_PDL_WAIT();
    __shared__ float da[320];
    __shared__ float ds[320];
    __shared__ float red[1];
    __shared__ float4 part[8][32];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float wv = w[0];
    const float4* dc4 = reinterpret_cast<const float4*>(dC + (size_t)b * dcs);
    const int DV4 = DV >> 2, D4 = D >> 2;
    const int vper = (DV4 + ATT_SPLIT - 1) / ATT_SPLIT, v0 = blockIdx.y * vper, v1 = min(v0 + vper, DV4);
    for (int t = warp; t < T; t += 8) {
        const float a = A[(size_t)b * T + t];
        const float4* vr = reinterpret_cast<const float4*>(V + (size_t)(b * T + t) * vs);
        float4* dvr = dV ? reinterpret_cast<float4*>(dV + (size_t)(b * T + t) * dvs) : nullptr;
        float acc = 0.f;
#pragma unroll 4
        for (int k = lane; k < DV4; k += 32) {
            const float4 d = dc4[k], v = vr[k];
            acc = fmaf(d.x, v.x, acc); acc = fmaf(d.y, v.y, acc); acc = fmaf(d.z, v.z, acc); acc = fmaf(d.w, v.w, acc);
            if (dvr && k >= v0 && k < v1) {
                float4 o = dvr[k];
                o.x = fmaf(a, d.x, o.x); o.y = fmaf(a, d.y, o.y); o.z = fmaf(a, d.z, o.z); o.w = fmaf(a, d.w, o.w);
                dvr[k] = o;
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) da[t] = acc;
    }
    __syncthreads();
    if (warp == 0) {
        float s = 0.f;
        for (int t = lane; t < T; t += 32) s = fmaf(da[t], A[(size_t)b * T + t], s);
        s = warp_sum(s);
        if (lane == 0) red[0] = s;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += 256) {
        float g = A[(size_t)b * T + t] * (da[t] - red[0]);
        if (mask) g *= mask[(size_t)b * T + t] * alpha;
        ds[t] = g;
    }
    __syncthreads();
    if (warp == 0 && dwpart && blockIdx.y == 0) {
        float s = 0.f;
        for (int t = lane; t < T; t += 32) s = fmaf(ds[t], sraw[(size_t)b * T + t], s);
        s = warp_sum(s);
        if (lane == 0) dwpart[b] = s;
    }
    // dq / dK for this CTA's column quads: 32 quads x 8 position groups
    const float4* q4 = reinterpret_cast<const float4*>(Q + (size_t)b * qs);
    const int kper = (D4 + ATT_SPLIT - 1) / ATT_SPLIT, k1 = min((int)(blockIdx.y + 1) * kper, D4);
    for (int c0 = blockIdx.y * kper; c0 < k1; c0 += 32) {
        const int c = c0 + lane;
        const bool ok = c < k1;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
            float4 qw = q4[c];
            qw.x *= wv; qw.y *= wv; qw.z *= wv; qw.w *= wv;
#pragma unroll 4
            for (int t = warp; t < T; t += 8) {
                const float g = ds[t];
                const float4 kv = reinterpret_cast<const float4*>(Km + (size_t)(b * T + t) * ks)[c];
                acc.x = fmaf(g, kv.x, acc.x); acc.y = fmaf(g, kv.y, acc.y); acc.z = fmaf(g, kv.z, acc.z); acc.w = fmaf(g, kv.w, acc.w);
                if (dK) {
                    float4* dk = reinterpret_cast<float4*>(dK + (size_t)(b * T + t) * dks) + c;
                    float4 o = *dk;
                    o.x = fmaf(g, qw.x, o.x); o.y = fmaf(g, qw.y, o.y); o.z = fmaf(g, qw.z, o.z); o.w = fmaf(g, qw.w, o.w);
                    *dk = o;
                }
            }
        }
        part[warp][lane] = acc;
        __syncthreads();
        if (warp == 0 && ok && dQ) {
            float4 r = part[0][lane];
#pragma unroll
            for (int g = 1; g < 8; ++g) { const float4 p4 = part[g][lane]; r.x += p4.x; r.y += p4.y; r.z += p4.z; r.w += p4.w; }
            float4* dq = reinterpret_cast<float4*>(dQ + (size_t)b * dqs) + c;
            float4 o = *dq;
            o.x = fmaf(r.x, wv, o.x); o.y = fmaf(r.y, wv, o.y); o.z = fmaf(r.z, wv, o.z); o.w = fmaf(r.w, wv, o.w);
            *dq = o;
        }
        __syncthreads();
    }
}
// out[0] += sum over a list of R arrays of n floats, in list order (the per-step, per-clip temperature-gradient partials)
__global__ void __launch_bounds__(256) sum_list_kernel(int R, int n, const float* const* __restrict__ list, float* __restrict__ out) {
    TR_PDL_WAIT();
    __shared__ float part[8];
    float a = 0.f;
    for (int i = threadIdx.x; i < R * n; i += 256) a += list[i / n][i % n];
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += part[i]; out[0] += t; }
}

// y = psine_w(x) [* keep-mask * alpha] [+ constant]: the activation, the dropout and the positional term that follow a per-step
// linear, in one pass (prenet: decoder.py:306-309; query: :359-360).
__global__ void psine_chain_fwd_kernel(int rows, int cols, const float* __restrict__ X, int xs, const float* __restrict__ w, const float* __restrict__ mask, int ms,
                                       float alpha, const float* __restrict__ addc, int as, float* __restrict__ Y, int ys) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)rows * cols) {
        const int r = i / cols, c = i % cols;
        float y = sinf(X[(size_t)r * xs + c]) * w[c];
        if (mask) y *= mask[(size_t)r * ms + c] * alpha;
        if (addc) y += addc[(size_t)r * as + c];
        Y[(size_t)r * ys + c] = y;
    }
}
// dx += dy m cos(x) w ; dw[c] += sum_r dy m sin(x).  CTA = 32 columns x 8 row lanes (rows <= 16: two passes), every load of
// the launch in flight at once; the column sums are folded in row order (deterministic).
__global__ void __launch_bounds__(256) psine_chain_bwd_kernel(int rows, int cols, const float* __restrict__ X, int xs, const float* __restrict__ w,
                                                              const float* __restrict__ mask, int ms, float alpha, const float* __restrict__ dY, int dys,
                                                              float* __restrict__ dX, int dxs, float* __restrict__ dw) {
    TR_PDL_WAIT();
    __shared__ float red[16][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const float wc = c < cols ? w[c] : 0.f;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int r = rl + 8 * p;
        float contrib = 0.f;
        if (c < cols && r < rows) {
            float g = dY[(size_t)r * dys + c];
            if (mask) g *= mask[(size_t)r * ms + c] * alpha;
            float sn, cs;
            sincosf(X[(size_t)r * xs + c], &sn, &cs);
            if (dX) dX[(size_t)r * dxs + c] += g * cs * wc;
            contrib = g * sn;
        }
        red[r][cl] = contrib;
    }
    __syncthreads();
    if (rl == 0 && c < cols && dw) {
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) acc += red[r][cl];
        dw[c] += acc;
    }
}

// ---- kernels: LSTM cell (gate order i, f, g, o) -------------------------------------------------------------
// gates [B][4H] (pre-activation) + c_prev [B][H] -> act [B][4H] (sigmoid/tanh applied, saved for backward), c [B][H], h [B][H]
__global__ void lstm_cell_fwd_kernel(int B, int H, const float* __restrict__ G, int gs, const float* __restrict__ Cp, int cps,
                                     float* __restrict__ Act, float* __restrict__ Cn, int cns, float* __restrict__ Hn, int hns) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)B * H) {
        const int b = i / H, j = i % H;
        const float* g = G + (size_t)b * gs;
        const float gi = sigmoidf_acc(g[j]), gf = sigmoidf_acc(g[H + j]), gg = tanhf(g[2 * H + j]), go = sigmoidf_acc(g[3 * H + j]);
        const float c = gf * Cp[(size_t)b * cps + j] + gi * gg;
        float* a = Act + (size_t)b * 4 * H;
        a[j] = gi; a[H + j] = gf; a[2 * H + j] = gg; a[3 * H + j] = go;
        Cn[(size_t)b * cns + j] = c;
        Hn[(size_t)b * hns + j] = go * tanhf(c);
    }
}
// Given dH, dC (of the outputs): dG (pre-activation gates, OVERWRITTEN) and dCp += .
__global__ void lstm_cell_bwd_kernel(int B, int H, const float* __restrict__ Act, const float* __restrict__ Cp, int cps, const float* __restrict__ Cn, int cns,
                                     const float* __restrict__ dH, int dhs, const float* __restrict__ dCn, int dcns,
                                     float* __restrict__ dG, int dgs, float* __restrict__ dCp, int dcps) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)B * H) {
        const int b = i / H, j = i % H;
        const float* a = Act + (size_t)b * 4 * H;
        const float gi = a[j], gf = a[H + j], gg = a[2 * H + j], go = a[3 * H + j];
        const float tc = tanhf(Cn[(size_t)b * cns + j]);
        const float dh = dH ? dH[(size_t)b * dhs + j] : 0.f;
        const float dc = (dCn ? dCn[(size_t)b * dcns + j] : 0.f) + dh * go * (1.f - tc * tc);
        float* dg = dG + (size_t)b * dgs;
        dg[j] = dc * gg * gi * (1.f - gi);
        dg[H + j] = dc * Cp[(size_t)b * cps + j] * gf * (1.f - gf);
        dg[2 * H + j] = dc * gi * (1.f - gg * gg);
        dg[3 * H + j] = dh * tc * go * (1.f - go);
        if (dCp) dCp[(size_t)b * dcps + j] += dc * gf;
    }
}

// ---- kernels: Conv1d support (rows (b, l) x channels) ------------------------------------------------------------------
// col[(b*Lo + lo)][ci*K + kk] = x[(b*L + lo*stride - pad + kk)][ci]  (zero outside [0, L))    — weight layout [co][ci][k]
__global__ void im2col1d_kernel(int B, int L, int Lo, int C, int K, int stride, int pad, const float* __restrict__ X, int xs, float* __restrict__ col) {
    TR_PDL_WAIT();
    const size_t total = (size_t)B * Lo * C * K;
    TR_EW_LOOP(total) {
        const int kk = i % K; size_t r = i / K;
        const int ci = r % C; r /= C;
        const int lo = r % Lo; const int b = r / Lo;
        const int l = lo * stride - pad + kk;
        col[i] = (l >= 0 && l < L) ? X[(size_t)(b * L + l) * xs + ci] : 0.f;
    }
}
// dx[(b*L + l)][ci] += sum_kk dcol[(b*Lo + lo)][ci*K + kk]  over (lo, kk) with lo*stride - pad + kk == l
__global__ void col2im1d_kernel(int B, int L, int Lo, int C, int K, int stride, int pad, const float* __restrict__ dcol, float* __restrict__ dX, int dxs) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)B * L * C) {
        const int ci = i % C; size_t r = i / C;
        const int l = r % L; const int b = r / L;
        float a = 0.f;
        for (int kk = 0; kk < K; ++kk) {
            const int num = l + pad - kk;
            if (num < 0 || num % stride) continue;
            const int lo = num / stride;
            if (lo >= Lo) continue;
            a += dcol[((size_t)(b * Lo + lo) * C + ci) * K + kk];
        }
        dX[(size_t)(b * L + l) * dxs + ci] += a;
    }
}
// adaptive_avg_pool1d over rows: y[(b*m + i)][c] = mean of x[(b*L + l)][c] for l in [floor(i L / m), ceil((i+1) L / m))
__global__ void adaptive_pool_fwd_kernel(int B, int L, int m, int C, const float* __restrict__ X, int xs, float* __restrict__ Y, int ys) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)B * m * C) {
        const int c = i % C; size_t r = i / C;
        const int j = r % m; const int b = r / m;
        const int lo = (j * L) / m, hi = ((j + 1) * L + m - 1) / m;
        float a = 0.f;
        for (int l = lo; l < hi; ++l) a += X[(size_t)(b * L + l) * xs + c];
        Y[(size_t)(b * m + j) * ys + c] = a / (float)(hi - lo);
    }
}
__global__ void adaptive_pool_bwd_kernel(int B, int L, int m, int C, const float* __restrict__ dY, int dys, float* __restrict__ dX, int dxs) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)B * L * C) {
        const int c = i % C; size_t r = i / C;
        const int l = r % L; const int b = r / L;
        float a = 0.f;
        for (int j = 0; j < m; ++j) {
            const int lo = (j * L) / m, hi = ((j + 1) * L + m - 1) / m;
            if (l >= lo && l < hi) a += dY[(size_t)(b * m + j) * dys + c] / (float)(hi - lo);
        }
        dX[(size_t)(b * L + l) * dxs + c] += a;
    }
}
// [B][C][L] <-> rows (b, l) x C
__global__ void bcl_to_rows_tr_kernel(int B, int C, int L, const float* __restrict__ X, float* __restrict__ Y, int ys) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)B * C * L) { const int l = i % L; size_t r = i / L; const int c = r % C; const int b = r / C; Y[(size_t)(b * L + l) * ys + c] = X[i]; }
}
__global__ void rows_to_bcl_tr_kernel(int B, int C, int L, const float* __restrict__ X, int xs, float* __restrict__ Y) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)B * C * L) { const int l = i % L; size_t r = i / L; const int c = r % C; const int b = r / C; Y[i] = X[(size_t)(b * L + l) * xs + c]; }
}

// ---- kernels: video frontend (NHWC rows (n, h, w) x C) ---------------------------------------------------------------
// Conv3d(3->24,(5,7,7),s(1,2,2),p(2,3,3)) forward: y[(n,ho,wo)][co] over the caller's NCDHW clip tensor; k = ((ci*5+kt)*7+kh)*7+kw.
// One thread per (position, co); the 735-tap patch is re-read through L1 by the 24 threads of a position.
__global__ void __launch_bounds__(256) stem_fwd_kernel(int B, int T, int H, int W, int Ho, int Wo, const float* __restrict__ X, const float* __restrict__ Wt,
                                                       float* __restrict__ Y) {
    TR_PDL_WAIT();
    const size_t total = (size_t)B * T * Ho * Wo * 24;
    TR_EW_LOOP(total) {
        const int co = i % 24; size_t r = i / 24;
        const int wo = r % Wo; r /= Wo;
        const int ho = r % Ho; r /= Ho;
        const int t = r % T; const int b = r / T;
        const float* w = Wt + (size_t)co * 735;
        float acc = 0.f;
        for (int ci = 0; ci < 3; ++ci)
            for (int kt = 0; kt < 5; ++kt) {
                const int ti = t + kt - 2;
                if (ti < 0 || ti >= T) continue;
                const float* plane = X + ((size_t)(b * 3 + ci) * T + ti) * H * W;
                for (int kh = 0; kh < 7; ++kh) {
                    const int hi = 2 * ho + kh - 3;
                    if (hi < 0 || hi >= H) continue;
                    for (int kw = 0; kw < 7; ++kw) {
                        const int wi = 2 * wo + kw - 3;
                        if (wi < 0 || wi >= W) continue;
                        acc = fmaf(__ldg(plane + (size_t)hi * W + wi), w[((ci * 5 + kt) * 7 + kh) * 7 + kw], acc);
                    }
                }
            }
        Y[i] = acc;
    }
}
// Tiled form for W <= 96 (the LRW / AVSpeech crops): one CTA = one output frame x 8 output rows x the full output width, for all
// 24 channels.  The 5 x 21 x (W+5) input patch of the three colour planes is staged in shared memory, split into even and odd
// columns so that the stride-2 reads of neighbouring output columns hit consecutive banks; the weights sit next to it as
// [tap][24] and are read as broadcast float4.  Each thread owns 2 output positions x 24 channels (48 accumulators): 48 FMAs per
// 2 patch words + 6 weight quads, i.e. the kernel runs on the FMA pipe instead of on L1 latency (the one-thread-per-output form
// above took 8.9 ms for 8 clips; this one is bound by 9.4 GFMA).
constexpr int STEM_EW = 51, STEM_PR = 21;
constexpr size_t STEM_TILED_SMEM = (size_t)(735 * 24 + 2 * 15 * STEM_PR * STEM_EW) * sizeof(float);
__global__ void __launch_bounds__(192) stem_fwd_tiled_kernel(int B, int T, int H, int W, int Ho, int Wo, const float* __restrict__ X, const float* __restrict__ Wt,
                                                             float* __restrict__ Y) {
    TR_PDL_WAIT();
    extern __shared__ float4 stem_sm4[];
    float* wts = reinterpret_cast<float*>(stem_sm4);              // [735][24]
    float* ev = wts + 735 * 24;                                      // [(ci*5+kt)][21][51]: input columns 2j - 3
    float* od = ev + 15 * STEM_PR * STEM_EW;                         //                      input columns 2j - 2
    const int tiles_h = (Ho + 7) / 8;
    int blk = blockIdx.x;
    const int th = blk % tiles_h; blk /= tiles_h;
    const int t = blk % T, b = blk / T;
    const int ho0 = th * 8;
    for (int i = threadIdx.x; i < 735 * 24; i += 192) { const int co = i / 735, k = i - co * 735; wts[k * 24 + co] = Wt[i]; }
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int row = warp; row < 15 * STEM_PR; row += 6) {
            const int cd = row / STEM_PR, r = row - cd * STEM_PR;
            const int ci = cd / 5, ti = t + cd % 5 - 2, hi = 2 * ho0 + r - 3;
            const bool ok = ti >= 0 && ti < T && hi >= 0 && hi < H;
            const float* src = X + (((size_t)(b * 3 + ci) * T + (ok ? ti : 0)) * H + (ok ? hi : 0)) * W;
            for (int j = lane; j < 2 * STEM_EW; j += 32) {
                const int wi = j - 3;
                const float v = (ok && wi >= 0 && wi < W) ? __ldg(src + wi) : 0.f;
                ((j & 1) ? od : ev)[row * STEM_EW + (j >> 1)] = v;
            }
        }
    }
    __syncthreads();
    const int wo = threadIdx.x % 48, rp = threadIdx.x / 48;          // output rows ho0 + rp and ho0 + rp + 4
    float acc[2][24];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int c = 0; c < 24; ++c) acc[p][c] = 0.f;
    for (int cd = 0; cd < 15; ++cd)
        for (int kh = 0; kh < 7; ++kh) {
            const int o0 = (cd * STEM_PR + 2 * rp + kh) * STEM_EW + wo, o1 = o0 + 8 * STEM_EW;
            const float4* wk = reinterpret_cast<const float4*>(wts + (cd * 7 + kh) * 7 * 24);
#pragma unroll
            for (int kw = 0; kw < 7; ++kw) {
                const float* pl = (kw & 1) ? od : ev;
                const float x0 = pl[o0 + (kw >> 1)], x1 = pl[o1 + (kw >> 1)];
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const float4 w4 = wk[kw * 6 + q];
                    acc[0][4 * q] = fmaf(x0, w4.x, acc[0][4 * q]); acc[0][4 * q + 1] = fmaf(x0, w4.y, acc[0][4 * q + 1]);
                    acc[0][4 * q + 2] = fmaf(x0, w4.z, acc[0][4 * q + 2]); acc[0][4 * q + 3] = fmaf(x0, w4.w, acc[0][4 * q + 3]);
                    acc[1][4 * q] = fmaf(x1, w4.x, acc[1][4 * q]); acc[1][4 * q + 1] = fmaf(x1, w4.y, acc[1][4 * q + 1]);
                    acc[1][4 * q + 2] = fmaf(x1, w4.z, acc[1][4 * q + 2]); acc[1][4 * q + 3] = fmaf(x1, w4.w, acc[1][4 * q + 3]);
                }
            }
        }
    if (wo >= Wo) return;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int ho = ho0 + rp + 4 * p;
        if (ho >= Ho) continue;
        float4* dst = reinterpret_cast<float4*>(Y + ((((size_t)b * T + t) * Ho + ho) * Wo + wo) * 24);
#pragma unroll
        for (int q = 0; q < 6; ++q) dst[q] = make_float4(acc[p][4 * q], acc[p][4 * q + 1], acc[p][4 * q + 2], acc[p][4 * q + 3]);
    }
}
// Weight gradient of the stem: part[chunk][co*735 + k] = sum over the chunk's positions of dY[pos][co] * patch[pos][k].
// Thread t owns taps k = t, t+256, t+512 (735 <= 768) for all 24 output channels: 72 accumulators, the patch values are
// gathered straight from the clip tensor (each thread decodes its taps once), dY rows are broadcast from shared memory.
__global__ void __launch_bounds__(256) stem_wgrad_kernel(int B, int T, int H, int W, int Ho, int Wo, int pos_per_chunk, const float* __restrict__ X,
                                                         const float* __restrict__ dY, float* __restrict__ part) {
    TR_PDL_WAIT();
    __shared__ __align__(16) float dys[32][24];
    __shared__ int4 pinfo[32];                                       // per staged position: (b, t, 2 ho - 3, 2 wo - 3)
    const int tid = threadIdx.x;
    const size_t npos = (size_t)B * T * Ho * Wo;
    const size_t p0 = (size_t)blockIdx.x * pos_per_chunk, p1 = min(npos, p0 + pos_per_chunk);
    int ci[3], kt[3], kh[3], kw[3]; bool kv[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int k = tid + 256 * j;
        kv[j] = k < 735;
        const int kk = kv[j] ? k : 0;
        ci[j] = kk / 245; const int r = kk % 245; kt[j] = r / 49 - 2; kh[j] = (r % 49) / 7; kw[j] = r % 7;
    }
    float acc[3][24];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int c = 0; c < 24; ++c) acc[j][c] = 0.f;
    for (size_t pb = p0; pb < p1; pb += 32) {
        const int nb = (int)min((size_t)32, p1 - pb);
        __syncthreads();
        for (int e = tid; e < nb * 24; e += 256) dys[e / 24][e % 24] = dY[(pb + e / 24) * 24 + e % 24];
        if (tid < nb) {
            size_t r = pb + tid;
            const int wo = r % Wo; r /= Wo;
            const int ho = r % Ho; r /= Ho;
            pinfo[tid] = make_int4((int)(r / T), (int)(r % T), 2 * ho - 3, 2 * wo - 3);
        }
        __syncthreads();
#pragma unroll 2
        for (int q = 0; q < nb; ++q) {
            const int4 pi = pinfo[q];
            float xv[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int ti = pi.y + kt[j], hi = pi.z + kh[j], wi = pi.w + kw[j];
                xv[j] = (kv[j] && ti >= 0 && ti < T && hi >= 0 && hi < H && wi >= 0 && wi < W)
                            ? __ldg(X + (((size_t)(pi.x * 3 + ci[j]) * T + ti) * H + hi) * W + wi) : 0.f;
            }
            const float4* d4 = reinterpret_cast<const float4*>(dys[q]);
#pragma unroll
            for (int c4 = 0; c4 < 6; ++c4) {
                const float4 d = d4[c4];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    acc[j][4 * c4] = fmaf(d.x, xv[j], acc[j][4 * c4]); acc[j][4 * c4 + 1] = fmaf(d.y, xv[j], acc[j][4 * c4 + 1]);
                    acc[j][4 * c4 + 2] = fmaf(d.z, xv[j], acc[j][4 * c4 + 2]); acc[j][4 * c4 + 3] = fmaf(d.w, xv[j], acc[j][4 * c4 + 3]);
                }
            }
        }
    }
    float* out = part + (size_t)blockIdx.x * (24 * 735);
#pragma unroll
    for (int j = 0; j < 3; ++j)
        if (kv[j])
#pragma unroll
            for (int c = 0; c < 24; ++c) out[c * 735 + tid + 256 * j] = acc[j][c];
}
// dst[i] += sum_chunk part[chunk][i]   (fixed order)
__global__ void sum_chunks_kernel(int n, int chunks, const float* __restrict__ part, float* __restrict__ dst) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)n) {
        float a = 0.f;
        for (int c = 0; c < chunks; ++c) a += part[(size_t)c * n + i];
        dst[i] += a;
    }
}
// MaxPool 3x3 stride 2 pad 1 over NHWC rows; idx = flat input row of the maximum (first in scan order on ties)
__global__ void maxpool_fwd_kernel(int N, int H, int W, int C, int Ho, int Wo, const float* __restrict__ X, float* __restrict__ Y, int* __restrict__ idx) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)N * Ho * Wo * C) {
        const int c = i % C; size_t r = i / C;
        const int wo = r % Wo; r /= Wo;
        const int ho = r % Ho; const int n = r / Ho;
        float best = -INFINITY; int bi = -1;
        for (int kh = 0; kh < 3; ++kh) {
            const int h = 2 * ho + kh - 1;
            if (h < 0 || h >= H) continue;
            for (int kw = 0; kw < 3; ++kw) {
                const int w = 2 * wo + kw - 1;
                if (w < 0 || w >= W) continue;
                const int row = (n * H + h) * W + w;
                const float v = X[(size_t)row * C + c];
                if (v > best) { best = v; bi = row; }
            }
        }
        Y[i] = best; idx[i] = bi;
    }
}
// gather form (deterministic): an input position receives dY of every window whose maximum it is
__global__ void maxpool_bwd_kernel(int N, int H, int W, int C, int Ho, int Wo, const float* __restrict__ dY, const int* __restrict__ idx, float* __restrict__ dX) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)N * H * W * C) {
        const int c = i % C; size_t r = i / C;
        const int row = (int)r;
        const int w = r % W; r /= W;
        const int h = r % H; const int n = r / H;
        float a = 0.f;
        for (int ho = (h + 1 - 2 + 1) / 2; ho <= (h + 1) / 2; ++ho) {
            if (ho < 0 || ho >= Ho) continue;
            for (int wo = (w + 1 - 2 + 1) / 2; wo <= (w + 1) / 2; ++wo) {
                if (wo < 0 || wo >= Wo) continue;
                const size_t o = ((size_t)(n * Ho + ho) * Wo + wo) * C + c;
                if (idx[o] == row) a += dY[o];
            }
        }
        dX[i] += a;
    }
}
// depthwise 3x3, pad 1, stride s, no bias; weight [C][9]
__global__ void dw3x3_fwd_kernel(int N, int H, int W, int C, int s, int Ho, int Wo, const float* __restrict__ X, int xs, const float* __restrict__ Wt,
                                 float* __restrict__ Y, int ys) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)N * Ho * Wo * C) {
        const int c = i % C; size_t r = i / C;
        const int wo = r % Wo; r /= Wo;
        const int ho = r % Ho; const int n = r / Ho;
        float a = 0.f;
        for (int kh = 0; kh < 3; ++kh) {
            const int h = ho * s + kh - 1;
            if (h < 0 || h >= H) continue;
            for (int kw = 0; kw < 3; ++kw) {
                const int w = wo * s + kw - 1;
                if (w < 0 || w >= W) continue;
                a = fmaf(X[(size_t)((n * H + h) * W + w) * xs + c], Wt[c * 9 + kh * 3 + kw], a);
            }
        }
        Y[(size_t)((n * Ho + ho) * Wo + wo) * ys + c] = a;
    }
}
__global__ void dw3x3_dgrad_kernel(int N, int H, int W, int C, int s, int Ho, int Wo, const float* __restrict__ dY, int dys, const float* __restrict__ Wt,
                                   float* __restrict__ dX, int dxs) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)N * H * W * C) {
        const int c = i % C; size_t r = i / C;
        const int w = r % W; r /= W;
        const int h = r % H; const int n = r / H;
        float a = 0.f;
        for (int kh = 0; kh < 3; ++kh) {
            const int hn = h + 1 - kh;
            if (hn < 0 || hn % s) continue;
            const int ho = hn / s;
            if (ho >= Ho) continue;
            for (int kw = 0; kw < 3; ++kw) {
                const int wn = w + 1 - kw;
                if (wn < 0 || wn % s) continue;
                const int wo = wn / s;
                if (wo >= Wo) continue;
                a = fmaf(dY[(size_t)((n * Ho + ho) * Wo + wo) * dys + c], Wt[c * 9 + kh * 3 + kw], a);
            }
        }
        dX[(size_t)((n * H + h) * W + w) * dxs + c] += a;
    }
}
// part[split][tap][c] = sum over the split's output positions of dY * x(tap)   — grid (ceil(C/32), splits), 8 row lanes x 32
// channels, fixed order; every thread keeps the nine taps of its channel (dY is read once per position, not once per tap);
// dw3x3_wfinish_kernel adds the splits into dW[c][tap]
__global__ void __launch_bounds__(256) dw3x3_wgrad_kernel(int N, int H, int W, int C, int s, int Ho, int Wo, const float* __restrict__ X, int xs,
                                                          const float* __restrict__ dY, int dys, float* __restrict__ part) {
    TR_PDL_WAIT();
    __shared__ float red[9][8][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const int rows = N * Ho * Wo;
    const int per = (rows + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    float a[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) a[t] = 0.f;
    if (c < C)
        for (int r = r0 + rl; r < r1; r += 8) {
            const int wo = r % Wo, ho = (r / Wo) % Ho, n = r / (Wo * Ho);
            const float dy = dY[(size_t)r * dys + c];
            const float* xn = X + (size_t)n * H * W * xs + c;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int h = ho * s + kh - 1;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int w = wo * s + kw - 1;
                    const bool ok = h >= 0 && h < H && w >= 0 && w < W;
                    const float x = ok ? xn[(size_t)(h * W + w) * xs] : 0.f;
                    a[kh * 3 + kw] = fmaf(dy, x, a[kh * 3 + kw]);
                }
            }
        }
#pragma unroll
    for (int t = 0; t < 9; ++t) red[t][rl][cl] = a[t];
    __syncthreads();
    for (int t = rl; t < 9; t += 8)
        if (c < C) {
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) v += red[t][i][cl];
            part[((size_t)blockIdx.y * 9 + t) * C + c] = v;
        }
}
__global__ void dw3x3_wfinish_kernel(int C, int splits, const float* __restrict__ part, float* __restrict__ dW) {
    TR_PDL_WAIT();
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;       // one warp per (tap, channel)
    if (i >= 9 * C) return;
    const int tap = i / C, c = i % C;
    float t = 0.f;
    for (int z = lane; z < splits; z += 32) t += part[((size_t)z * 9 + tap) * C + c];
    t = warp_sum(t);
    if (lane == 0) dW[c * 9 + tap] += t;
}
// y[:, 2j] = a[:, j], y[:, 2j+1] = b[:, j]   (torch.cat + channel_shuffle(groups=2), shufflenetv2.py:26-40,92-104)
__global__ void interleave2_fwd_kernel(int rows, int half, const float* __restrict__ A, int as, const float* __restrict__ Bm, int bs, float* __restrict__ Y, int ys) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)rows * 2 * half) {
        const int c = i % (2 * half); const int r = i / (2 * half);
        Y[(size_t)r * ys + c] = (c & 1) ? Bm[(size_t)r * bs + (c >> 1)] : A[(size_t)r * as + (c >> 1)];
    }
}
__global__ void interleave2_bwd_kernel(int rows, int half, const float* __restrict__ dY, int dys, float* __restrict__ dA, int das, float* __restrict__ dB, int dbs) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)rows * 2 * half) {
        const int c = i % (2 * half); const int r = i / (2 * half);
        const float d = dY[(size_t)r * dys + c];
        if (c & 1) { if (dB) dB[(size_t)r * dbs + (c >> 1)] += d; }
        else if (dA) dA[(size_t)r * das + (c >> 1)] += d;
    }
}
// F.normalize(p=2, dim=-1, eps=1e-12): y = x / max(||x||, eps); one warp per row; nrm saved
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(int rows, int cols, const float* __restrict__ X, int xs, float* __restrict__ Y, int ys, float* __restrict__ nrm) {
    TR_PDL_WAIT();
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= rows) return;
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) { const float x = X[(size_t)r * xs + c]; s = fmaf(x, x, s); }
    s = fmaxf(sqrtf(warp_sum(s)), 1e-12f);
    for (int c = lane; c < cols; c += 32) Y[(size_t)r * ys + c] = X[(size_t)r * xs + c] / s;
    if (lane == 0) nrm[r] = s;
}
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(int rows, int cols, const float* __restrict__ Y, int ys, const float* __restrict__ nrm,
                                                         const float* __restrict__ dY, int dys, float* __restrict__ dX, int dxs) {
    TR_PDL_WAIT();
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= rows) return;
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s = fmaf(Y[(size_t)r * ys + c], dY[(size_t)r * dys + c], s);
    s = warp_sum(s);
    const float inv = 1.f / nrm[r];
    for (int c = lane; c < cols; c += 32) dX[(size_t)r * dxs + c] += (dY[(size_t)r * dys + c] - Y[(size_t)r * ys + c] * s) * inv;
}

}  // namespace tr
}  // namespace l2s

// ========================================================================================================================
// host side: arenas, tape, ops
// ========================================================================================================================
namespace l2s {
namespace tr {

struct Arena {                     // bump allocator over cudaMalloc'ed blocks that persist across steps
    struct Block { char* p; size_t cap, used; };
    std::vector<Block> blocks;
    size_t cur = 0;
    size_t block_bytes = (size_t)256 << 20;
    float* alloc(size_t floats) {
        const size_t bytes = (std::max<size_t>(floats, 1) * sizeof(float) + 255) & ~size_t(255);
        while (cur < blocks.size() && blocks[cur].cap - blocks[cur].used < bytes) ++cur;
        if (cur == blocks.size()) {
            Block nb; nb.cap = std::max(block_bytes, bytes); nb.used = 0; nb.p = nullptr;
            L2S_CUDA(cudaMalloc(&nb.p, nb.cap));
            blocks.push_back(nb);
        }
        float* r = reinterpret_cast<float*>(blocks[cur].p + blocks[cur].used);
        blocks[cur].used += bytes;
        return r;
    }
    void reset() { for (auto& b : blocks) b.used = 0; cur = 0; }
    void zero_used(cudaStream_t s) { for (auto& b : blocks) if (b.used) L2S_CUDA(cudaMemsetAsync(b.p, 0, b.used, s)); }
    void free_all() { for (auto& b : blocks) cudaFree(b.p); blocks.clear(); cur = 0; }
};

struct Param { float* v = nullptr; float* g = nullptr; int64_t n = 0; };

// ---- CUDA-graph replay -----------------------------------------------------------------------------------------------------
// A train step launches ~13 K small kernels from host closures (~8 us of host time each).  The arenas hand out the same
// addresses for the same shapes, so the whole forward (and the whole backward) of one (B, T, M, ...) key is captured once as a
// CUDA graph and replayed: the second call with a key captures, later calls replay.  Caller tensors never appear inside a
// graph — inputs are staged into arena memory before the launch, outputs copied out after it.
struct HostTables {                // pinned host memory read by the captured H2D copies of the row-pointer tables at every replay
    char* p = nullptr; size_t cap = 0, used = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        L2S_CUDA(cudaMallocHost(&p, bytes));
        cap = bytes;
    }
    void* take(size_t bytes) {
        if (used + bytes > cap) throw L2sError(2, "train: row-pointer tables outgrew their pinned buffer");
        void* r = p + used; used += (bytes + 15) & ~size_t(15); return r;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = used = 0; }
};
struct GraphSlot {
    cudaGraphExec_t fwd = nullptr, bwd = nullptr;
    int64_t fwd_launches = 0, bwd_launches = 0;
    int seen = 0;
    size_t table_bytes = 0;        // what the last eager backward needed for its pointer tables
    HostTables tables;
    void drop_graphs() {
        if (fwd) cudaGraphExecDestroy(fwd);
        if (bwd) cudaGraphExecDestroy(bwd);
        fwd = bwd = nullptr;
    }
    void release() { drop_graphs(); tables.release(); }
};
template <class F>
inline cudaGraphExec_t capture_graph(cudaStream_t cs, F&& body) {
    L2S_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
    cudaGraph_t g = nullptr;
    try { body(); }
    catch (...) { cudaStreamEndCapture(cs, &g); if (g) cudaGraphDestroy(g); throw; }
    L2S_CUDA(cudaStreamEndCapture(cs, &g));
    cudaGraphExec_t x = nullptr;
    const cudaError_t err = cudaGraphInstantiate(&x, g, 0);
    cudaGraphDestroy(g);
    if (err != cudaSuccess) throw L2sError(2, std::string("train: cudaGraphInstantiate: ") + cudaGetErrorString(err));
    return x;
}

inline int ew_blocks(size_t total) { return (int)std::min<size_t>(std::max<size_t>((total + 255) / 256, 1), 148 * 16); }

constexpr int SKINNY_SMEM_MAX = 132 * 1024;

struct Engine {
    Context* ctx = nullptr;
    cudaStream_t s = nullptr;
    Arena vals, grads;
    std::vector<std::function<void()>> tape;
    std::map<std::string, Param>* params = nullptr;
    bool update_bn_running = true;
    bool exact_gemm = false;           // large GEMMs with fp32 FMAs instead of 3xTF32 tensor-core products (see sgemm_kernel)
    int64_t* launches = nullptr;
    bool capturing = false;            // the body runs under stream capture: no synchronisation, host tables must persist
    HostTables* tables = nullptr;      // where a captured backward keeps its row-pointer tables
    size_t table_bytes = 0;            // bytes of pointer tables the last backward uploaded

    bool ready = false;
    // One-time per-device kernel attributes; called outside any stream capture.
    void setup() {
        if (ready) return;
        {
            L2S_CUDA(cudaFuncSetAttribute(skinny_nt_smem_kernel<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKINNY_SMEM_MAX));
            L2S_CUDA(cudaFuncSetAttribute(skinny_nt_smem_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKINNY_SMEM_MAX));
            L2S_CUDA(cudaFuncSetAttribute(skinny_nt_smem_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKINNY_SMEM_MAX));
            L2S_CUDA(cudaFuncSetAttribute(skinny_nt_smem_kernel<2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKINNY_SMEM_MAX));
            L2S_CUDA(cudaFuncSetAttribute(skinny_nt_smem_kernel<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKINNY_SMEM_MAX));
            L2S_CUDA(cudaFuncSetAttribute(skinny_nn_strip_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKINNY_SMEM_MAX));
            L2S_CUDA(cudaFuncSetAttribute(skinny_nn_strip_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKINNY_SMEM_MAX));
            L2S_CUDA(cudaFuncSetAttribute(stem_fwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEM_TILED_SMEM));
            ready = true;
        }
    }
    void release() { vals.free_all(); grads.free_all(); }
    void begin(Context* c, cudaStream_t stream, std::map<std::string, Param>* p) {
        ctx = c; s = stream; params = p; launches = &c->launches;
        vals.reset(); grads.reset(); tape.clear(); deferred.clear(); deferred_scalar.clear(); deferred_scalar_n.clear();
    }
    void ck(const char* what) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) throw L2sError(2, std::string("train: ") + what + ": " + cudaGetErrorString(e));
        ++*launches;
    }
    TT make(int rows, int cols, bool grad = true) {
        TT t; t.rows = rows; t.cols = cols; t.rs = cols;
        t.v = vals.alloc((size_t)rows * cols);
        t.g = grad ? grads.alloc((size_t)rows * cols) : nullptr;
        return t;
    }
    float* scratch(size_t floats) { return vals.alloc(floats); }
    // wraps caller memory (inputs)
    TT wrap(const float* p, int rows, int cols, bool grad) {
        TT t; t.v = const_cast<float*>(p); t.rows = rows; t.cols = cols; t.rs = cols;
        t.g = grad ? grads.alloc((size_t)rows * cols) : nullptr;
        return t;
    }
    TT param(const std::string& key, int rows, int cols) {
        auto it = params->find(key);
        if (it == params->end()) throw L2sError(3, "train: parameter not bound: " + key + " (l2s_train_bind)");
        if (it->second.n != (int64_t)rows * cols) throw L2sError(1, "train: parameter " + key + " has " + std::to_string(it->second.n) + " elements, expected " + std::to_string((int64_t)rows * cols));
        TT t; t.v = it->second.v; t.g = it->second.g; t.rows = rows; t.cols = cols; t.rs = cols;
        return t;
    }
    // Runs the tape in reverse.  Gradients of the outputs must have been written into their .g before the call; the
    // gradient arena was zero-filled by begin_backward().
    void begin_backward() { grads.zero_used(s); }
    void backward() {
        for (size_t i = tape.size(); i-- > 0;) tape[i]();
        tape.clear();
        flush_deferred();
    }

    // ---- column reductions (two deterministic stages) -----------------------------------------------------------------------
    template <int OP>
    void colred(int rows, int cols, const float* X, int xs, const float* Y, int ys, const float* aux, const float* aux2, float eps, float scale,
                float* out, bool accumulate) {
        // enough CTAs to fill the chip even for 24 columns (the stem's 534 K rows were walked by 64 CTAs, 1 044 dependent loads each)
        const int colblocks = (cols + 31) / 32;
        const int splits = std::max(1, std::min(std::min(1024, (148 * 8 + colblocks - 1) / colblocks), rows / 256));
        float* part = scratch((size_t)splits * cols);
        launch(colreduce_kernel<OP>, dim3(colblocks, splits), 256, 0, s, rows, cols, X, xs, Y, ys, aux, aux2, eps, part);
        ck("column reduce");
        launch(colfinish_kernel, (cols * 32 + 255) / 256, 256, 0, s, cols, splits, part, scale, out, accumulate ? 1 : 0);
        ck("column reduce finish");
    }

    // ---- deferred weight gradients of the per-step (few-row) linears ---------------------------------------------------------
    struct Deferred { TT W; float* db = nullptr; std::vector<const float*> a, b; };
    std::map<float*, Deferred> deferred;
    std::map<float*, std::vector<const float*>> deferred_scalar;      // scalar parameter gradient <- list of partial arrays
    std::map<float*, int> deferred_scalar_n;
    void flush_deferred() {
        table_bytes = 0;
        for (auto& kv : deferred) {
            Deferred& d = kv.second;
            const int R = (int)d.a.size();
            const size_t half = (size_t)R * sizeof(float*);
            const float** tab = reinterpret_cast<const float**>(scratch((size_t)R * 4));      // 2 R pointers of 8 bytes
            if (capturing) {
                char* host = static_cast<char*>(tables->take(2 * half));
                memcpy(host, d.a.data(), half); memcpy(host + half, d.b.data(), half);
                L2S_CUDA(cudaMemcpyAsync(tab, host, 2 * half, cudaMemcpyHostToDevice, s));
            } else {
                L2S_CUDA(cudaMemcpyAsync(tab, d.a.data(), half, cudaMemcpyHostToDevice, s));
                L2S_CUDA(cudaMemcpyAsync(tab + R, d.b.data(), half, cudaMemcpyHostToDevice, s));
            }
            table_bytes += 2 * half + 16;
            const int N = d.W.rows, K = d.W.cols;
            launch(sgemm_tn_rows_kernel, dim3((K + 63) / 64, (N + 63) / 64), 256, 0, s, N, K, R, tab, tab + R, d.W.g, d.W.rs, d.db);
            ck("deferred weight gradient");
        }
        for (auto& kv : deferred_scalar) {
            const int R = (int)kv.second.size();
            const size_t bytes = (size_t)R * sizeof(float*);
            const float** tab = reinterpret_cast<const float**>(scratch((size_t)R * 2));
            if (capturing) {
                char* host = static_cast<char*>(tables->take(bytes));
                memcpy(host, kv.second.data(), bytes);
                L2S_CUDA(cudaMemcpyAsync(tab, host, bytes, cudaMemcpyHostToDevice, s));
            } else L2S_CUDA(cudaMemcpyAsync(tab, kv.second.data(), bytes, cudaMemcpyHostToDevice, s));
            table_bytes += bytes + 16;
            launch(sum_list_kernel, 1, 256, 0, s, R, deferred_scalar_n[kv.first], tab, kv.first);
            ck("deferred scalar gradient");
        }
        // eager: the host pointer tables must outlive the asynchronous copies
        if (!capturing) L2S_CUDA(cudaStreamSynchronize(s));
        deferred.clear(); deferred_scalar.clear(); deferred_scalar_n.clear();
    }

    // ---- GEMM helpers ----------------------------------------------------------------------------------------------------
    template <int TA, int TB>
    void gemm(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, bool acc) {
        if (M <= 0 || N <= 0 || K <= 0) return;
        dim3 grid((N + 63) / 64, (M + 63) / 64);
        const int tiles = grid.x * grid.y;
        if (tiles <= 148 && K >= 1024) {                     // fewer output tiles than SMs, long reduction: split K over the chip
            const int splits = std::max(2, std::min(std::min(128, 296 / tiles), (K + 255) / 256));
            const int kper = (((K + splits - 1) / splits) + 31) / 32 * 32;
            const int nz = (K + kper - 1) / kper;
            float* part = scratch((size_t)nz * M * N);
            grid.z = nz;
            if (exact_gemm) launch(sgemm_kernel<TA, TB, true>, grid, 256, 0, s, M, N, K, A, lda, B, ldb, part, N, 0, kper);
            else launch(sgemm_kernel<TA, TB, false>, grid, 256, 0, s, M, N, K, A, lda, B, ldb, part, N, 0, kper);
            ck("sgemm split-K");
            launch(skinny_nn_sum_kernel, ew_blocks((size_t)M * N), 256, 0, s, M, N, nz, part, C, ldc, acc ? 1 : 0);
            ck("sgemm split-K sum");
            return;
        }
        if (exact_gemm) launch(sgemm_kernel<TA, TB, true>, grid, 256, 0, s, M, N, K, A, lda, B, ldb, C, ldc, acc ? 1 : 0, K);
        else launch(sgemm_kernel<TA, TB, false>, grid, 256, 0, s, M, N, K, A, lda, B, ldb, C, ldc, acc ? 1 : 0, K);
        ck("sgemm");
    }

    // y = x W^T (+ b): x [R,K], W [N,K] (nn.Linear / flattened Conv1d weight), b [N] or empty
    // Few-row product(s) y = x1 W1^T (+ x2 W2^T) + biases: rows per warp chosen so that the launch has ~100+ CTAs (it is bound by
    // bytes in flight, not by FMAs).
    void launch_skinny_nt(int R, int N, int K1, const float* x1, int lx1, const float* W1, int lw1, const float* b1, int K2, const float* x2, int lx2,
                          const float* W2, int lw2, const float* b2, float* y, int ly, int accf) {
        const int MT = R <= 8 ? 8 : 16;
        const size_t smem = (size_t)MT * (K1 + K2) * sizeof(float);
#define L2S_NT(NW_, MT_) launch(skinny_nt_smem_kernel<NW_, MT_>, (N + 8 * NW_ - 1) / (8 * NW_), 256, smem, s, R, N, K1, x1, lx1, W1, lw1, b1, K2, x2, lx2, W2, lw2, b2, y, ly, accf)
        if (MT == 8) {
            if (N >= 3072) L2S_NT(4, 8); else if (N >= 1536) L2S_NT(2, 8); else L2S_NT(1, 8);
        } else {
            if (N >= 1536) L2S_NT(2, 16); else L2S_NT(1, 16);
        }
#undef L2S_NT
        ck("skinny_nt");
    }
    // dx (+ dx2) += dy W (+ dy W2), few rows
    void launch_skinny_nn(int R, int N, int K1, const float* dy, int ldy, const float* W1, int lw1, float* dx1, int ldx1, int K2, const float* W2, int lw2,
                          float* dx2, int ldx2) {
        const int MT = R <= 8 ? 8 : 16;
        const size_t smem = skinny_nn_smem(R, N);
        const dim3 grid((std::max(K1, K2) + 7) / 8, K2 ? 2 : 1);
        if (MT == 8) launch(skinny_nn_strip_kernel<8>, grid, NN_THREADS, smem, s, R, N, K1, dy, ldy, W1, lw1, dx1, ldx1, K2, W2, lw2, dx2, ldx2);
        else launch(skinny_nn_strip_kernel<16>, grid, NN_THREADS, smem, s, R, N, K1, dy, ldy, W1, lw1, dx1, ldx1, K2, W2, lw2, dx2, ldx2);
        ck("skinny_nn strip");
    }
    static size_t skinny_nn_smem(int R, int N) { return (size_t)(R <= 8 ? 8 : 16) * ((N + 255) / 256 * 256) * sizeof(float); }
    // dst: write into this [R,N] view instead of a fresh tensor; accumulate: y += (a second operand of the same pre-activation,
    // e.g. the recurrent half of the LSTM gates — the view's gradient then serves both products).
    TT linear(const TT& x, const TT& W, const TT* b, const TT* dst = nullptr, bool accumulate = false) {
        const int R = x.rows, K = x.cols, N = W.rows;
        if (W.cols != K) throw L2sError(1, "train: linear shape mismatch");
        if (dst && (dst->rows != R || dst->cols != N)) throw L2sError(1, "train: linear destination shape mismatch");
        if (accumulate && !dst) throw L2sError(1, "train: accumulating linear needs a destination");
        TT y = dst ? *dst : make(R, N);
        const int accf = accumulate ? 1 : 0;
        if (R <= 16) {
            const bool vec = !(K & 3) && !(x.rs & 3) && !(W.rs & 3) && !((reinterpret_cast<uintptr_t>(x.v) | reinterpret_cast<uintptr_t>(W.v)) & 15);
            const int MT = R <= 8 ? 8 : 16;
            const size_t smem = (size_t)MT * K * sizeof(float);
            if (vec && smem <= (size_t)SKINNY_SMEM_MAX) {
                launch_skinny_nt(R, N, K, x.v, x.rs, W.v, W.rs, b ? b->v : nullptr, 0, nullptr, 0, nullptr, 0, nullptr, y.v, y.rs, accf);
            } else if (vec) launch(skinny_nt_kernel<true>, (N * 32 + 255) / 256, 256, 0, s, R, N, K, x.v, x.rs, W.v, W.rs, b ? b->v : nullptr, y.v, y.rs, accf);
            else launch(skinny_nt_kernel<false>, (N * 32 + 255) / 256, 256, 0, s, R, N, K, x.v, x.rs, W.v, W.rs, b ? b->v : nullptr, y.v, y.rs, accf);
            if (!(vec && smem <= (size_t)SKINNY_SMEM_MAX)) ck("skinny_nt");
        } else {
            gemm<0, 1>(R, N, K, x.v, x.rs, W.v, W.rs, y.v, y.rs, accumulate);
            if (b) {
                launch(ew_fwd_kernel<EW_ADDROW>, ew_blocks(y.numel()), 256, 0, s, R, N, y.v, y.rs, b->v, 0, 0.f, R, y.v, y.rs);   // group = R: one broadcast row
                ck("bias");
            }
        }
        TT bb; if (b) bb = *b;
        const bool hasb = b != nullptr;
        const bool defer = R <= 16 && W.g != nullptr;       // per-step layer: its weight gradient is gathered over all steps at the end
        if (defer) {
            Deferred& d = deferred[W.g];
            d.W = W;
            if (hasb && bb.g) d.db = bb.g;                   // the bias gradient rides along with the gathered weight gradient
            for (int r = 0; r < R; ++r) { d.a.push_back(y.g + (size_t)r * y.rs); d.b.push_back(x.v + (size_t)r * x.rs); }
        }
        tape.push_back([=]() {
            if (x.g) {
                const bool vecb = !(K & 3) && !(W.rs & 3) && !(reinterpret_cast<uintptr_t>(W.v) & 15) && skinny_nn_smem(R, N) <= (size_t)SKINNY_SMEM_MAX;
                if (R <= 16 && vecb) {
                    launch_skinny_nn(R, N, K, y.g, y.rs, W.v, W.rs, x.g, x.rs, 0, nullptr, 0, nullptr, 0);
                } else if (R <= 16) {
                    const int nslice = 128, nslices = (N + nslice - 1) / nslice;
                    float* part = scratch((size_t)nslices * R * K);
                    launch(skinny_nn_part_kernel, dim3((K + 63) / 64, nslices), 256, 0, s, R, N, K, nslice, y.g, y.rs, W.v, W.rs, part);
                    ck("skinny_nn part");
                    launch(skinny_nn_sum_kernel, ew_blocks((size_t)R * K), 256, 0, s, R, K, nslices, part, x.g, x.rs, 1);
                    ck("skinny_nn sum");
                } else gemm<0, 0>(R, K, N, y.g, y.rs, W.v, W.rs, x.g, x.rs, true);
            }
            if (W.g && !defer) gemm<1, 0>(N, K, R, y.g, y.rs, x.v, x.rs, W.g, W.rs, true);
            if (hasb && bb.g && !defer) colred<COL_SUM>(R, N, y.g, y.rs, nullptr, 0, nullptr, nullptr, 0.f, 1.f, bb.g, true);
        });
        return y;
    }
    // y = x1 W1^T + b1 + x2 W2^T + b2 in one launch forward and one launch backward when the rows are few (the LSTM gates of one
    // step: input and recurrent product, nn.LSTM's b_ih + b_hh); otherwise two accumulating linears.
    TT linear2(const TT& x1, const TT& W1, const TT* b1, const TT& x2, const TT& W2, const TT* b2, const TT* dst = nullptr) {
        const int R = x1.rows, K1 = x1.cols, K2 = x2.cols, N = W1.rows;
        auto aligned = [](const TT& x, const TT& W) {
            return !(x.cols & 3) && !(x.rs & 3) && !(W.rs & 3) && !((reinterpret_cast<uintptr_t>(x.v) | reinterpret_cast<uintptr_t>(W.v)) & 15);
        };
        const int MT = R <= 8 ? 8 : 16;
        const size_t smem = (size_t)MT * (K1 + K2) * sizeof(float);
        const size_t smemb = skinny_nn_smem(R, N);
        const bool fused = R <= 16 && x2.rows == R && W2.rows == N && W1.cols == K1 && W2.cols == K2 && aligned(x1, W1) && aligned(x2, W2) &&
                           smem <= (size_t)SKINNY_SMEM_MAX && smemb <= (size_t)SKINNY_SMEM_MAX && x1.g && x2.g && W1.g && W2.g;
        if (!fused) {
            TT y = linear(x1, W1, b1, dst);
            linear(x2, W2, b2, &y, true);
            return y;
        }
        TT y = dst ? *dst : make(R, N);
        const float* bv1 = b1 ? b1->v : nullptr; const float* bv2 = b2 ? b2->v : nullptr;
        launch_skinny_nt(R, N, K1, x1.v, x1.rs, W1.v, W1.rs, bv1, K2, x2.v, x2.rs, W2.v, W2.rs, bv2, y.v, y.rs, 0);
        for (int which = 0; which < 2; ++which) {
            const TT& x = which ? x2 : x1; const TT& W = which ? W2 : W1; const TT* b = which ? b2 : b1;
            Deferred& d = deferred[W.g];
            d.W = W;
            if (b && b->g) d.db = b->g;
            for (int r = 0; r < R; ++r) { d.a.push_back(y.g + (size_t)r * y.rs); d.b.push_back(x.v + (size_t)r * x.rs); }
        }
        tape.push_back([=]() {
            launch_skinny_nn(R, N, K1, y.g, y.rs, W1.v, W1.rs, x1.g, x1.rs, K2, W2.v, W2.rs, x2.g, x2.rs);
        });
        return y;
    }
    // y = x E  (E [K,N] row-major parameter: z @ word_embeddings, decoder.py:258)
    TT matmul_nn(const TT& x, const TT& E) {
        const int R = x.rows, K = x.cols, N = E.cols;
        TT y = make(R, N);
        gemm<0, 0>(R, N, K, x.v, x.rs, E.v, E.rs, y.v, y.rs, false);
        tape.push_back([=]() {
            if (x.g) gemm<0, 1>(R, K, N, y.g, y.rs, E.v, E.rs, x.g, x.rs, true);
            if (E.g) gemm<1, 0>(K, N, R, x.v, x.rs, y.g, y.rs, E.g, E.rs, true);
        });
        return y;
    }

    // ---- elementwise ------------------------------------------------------------------------------------------------------
    template <int OP>
    TT ew(const TT& x, const float* aux, int as, float alpha, int group = 1) {
        TT y = make(x.rows, x.cols);
        launch(ew_fwd_kernel<OP>, ew_blocks(x.numel()), 256, 0, s, x.rows, x.cols, x.v, x.rs, aux, as, alpha, group, y.v, y.rs);
        ck("elementwise");
        tape.push_back([=]() {
            if (!x.g) return;
            launch(ew_bwd_kernel<OP>, ew_blocks(x.numel()), 256, 0, s, x.rows, x.cols, x.v, x.rs, aux, as, alpha, y.g, y.rs, x.g, x.rs);
            ck("elementwise bwd");
        });
        return y;
    }
    TT silu(const TT& x) { return ew<EW_SILU>(x, nullptr, 0, 0.f); }
    TT relu(const TT& x) { return ew<EW_RELU>(x, nullptr, 0, 0.f); }
    TT scale(const TT& x, float a) { return ew<EW_SCALE>(x, nullptr, 0, a); }
    TT add_const(const TT& x, const float* cst, int cs) { return ew<EW_ADDCONST>(x, cst, cs, 0.f); }        // + a constant tensor (positions, gumbel)
    TT dropout(const TT& x, const float* keep, int ks, float p) { return ew<EW_MASK>(x, keep, ks, 1.0f / (1.0f - p)); }
    TT psine(const TT& x, const TT& w) {
        TT y = ew<EW_PSINE>(x, w.v, 0, 0.f);
        if (w.g) tape.push_back([=]() { colred<COL_PSINE_DW>(x.rows, x.cols, x.v, x.rs, y.g, y.rs, nullptr, nullptr, 0.f, 1.f, w.g, true); });
        return y;
    }
    TT prelu(const TT& x, const TT& w) {
        TT y = ew<EW_PRELU>(x, w.v, 0, 0.f);
        if (w.g) tape.push_back([=]() { colred<COL_PRELU_DW>(x.rows, x.cols, x.v, x.rs, y.g, y.rs, nullptr, nullptr, 0.f, 1.f, w.g, true); });
        return y;
    }
    TT add(const TT& a, const TT& b, const TT* dst = nullptr) {
        TT y = dst ? *dst : make(a.rows, a.cols);
        launch(ew_fwd_kernel<EW_ADD>, ew_blocks(a.numel()), 256, 0, s, a.rows, a.cols, a.v, a.rs, b.v, b.rs, 0.f, 1, y.v, y.rs);
        ck("add");
        tape.push_back([=]() {
            for (const TT* t : {&a, &b})
                if (t->g) { launch(ew_bwd_kernel<EW_COPY>, ew_blocks(y.numel()), 256, 0, s, y.rows, y.cols, nullptr, 0, nullptr, 0, 0.f, y.g, y.rs, t->g, t->rs); ck("add bwd"); }
        });
        return y;
    }
    // y = flag[0] != 0 ? a : b, the flag read on the DEVICE (the teacher-forcing coin of decoder.py:355-357: the launch sequence
    // does not depend on the draw, so one captured graph serves every step).  Only b carries a gradient (a = teacher frames).
    TT select(const float* flag, const TT& a, const TT& b) {
        TT y = make(a.rows, a.cols);
        launch(select_fwd_kernel, ew_blocks(a.numel()), 256, 0, s, a.rows, a.cols, flag, a.v, a.rs, b.v, b.rs, y.v, y.rs);
        ck("select");
        tape.push_back([=]() {
            if (!b.g) return;
            launch(select_bwd_kernel, ew_blocks(y.numel()), 256, 0, s, y.rows, y.cols, flag, y.g, y.rs, b.g, b.rs);
            ck("select bwd");
        });
        return y;
    }
    // y[(g*group + j)] = x[(g*group + j)] + v[g]     (attention_site broadcast over the T frames of a clip, decoder.py:327)
    TT add_rows(const TT& x, const TT& v, int group) {
        TT y = make(x.rows, x.cols);
        launch(ew_fwd_kernel<EW_ADDROW>, ew_blocks(x.numel()), 256, 0, s, x.rows, x.cols, x.v, x.rs, v.v, v.rs, 0.f, group, y.v, y.rs);
        ck("add rows");
        tape.push_back([=]() {
            if (x.g) { launch(ew_bwd_kernel<EW_COPY>, ew_blocks(y.numel()), 256, 0, s, y.rows, y.cols, nullptr, 0, nullptr, 0, 0.f, y.g, y.rs, x.g, x.rs); ck("add rows bwd"); }
            if (v.g) { launch(addrow_bwd_kernel, ew_blocks((size_t)v.rows * v.cols), 256, 0, s, v.rows, group, v.cols, y.g, y.rs, v.g, v.rs); ck("add rows bwd v"); }
        });
        return y;
    }
    TT scale_param(const TT& x, const TT& w) {
        TT y = make(x.rows, x.cols);
        launch(scale_param_kernel, ew_blocks(x.numel()), 256, 0, s, x.rows, x.cols, x.v, x.rs, w.v, y.v, y.rs);
        ck("scale param");
        tape.push_back([=]() {
            if (x.g) { launch(scale_param_bwd_kernel, ew_blocks(x.numel()), 256, 0, s, x.rows, x.cols, w.v, y.g, y.rs, x.g, x.rs); ck("scale param bwd"); }
            if (w.g) { launch(dot_all_kernel, 1, 1024, 0, s, x.rows, x.cols, x.v, x.rs, y.g, y.rs, w.g, 1); ck("scale param dw"); }
        });
        return y;
    }
    // dense copy of a (possibly strided) view, gradient flows back
    TT copy(const TT& x) { return ew<EW_COPY>(x, nullptr, 0, 0.f); }
    // [a | b | ...] along the columns
    TT concat_cols(const std::vector<TT>& parts) {
        int cols = 0;
        for (auto& p : parts) cols += p.cols;
        TT y = make(parts[0].rows, cols);
        int c0 = 0;
        for (auto& p : parts) {
            launch(ew_fwd_kernel<EW_COPY>, ew_blocks(p.numel()), 256, 0, s, p.rows, p.cols, p.v, p.rs, nullptr, 0, 0.f, 1, y.v + c0, y.rs);
            ck("concat");
            c0 += p.cols;
        }
        std::vector<TT> ps = parts;
        tape.push_back([=]() {
            int c = 0;
            for (auto& p : ps) {
                if (p.g) { launch(ew_bwd_kernel<EW_COPY>, ew_blocks(p.numel()), 256, 0, s, p.rows, p.cols, nullptr, 0, nullptr, 0, 0.f, y.g + c, y.rs, p.g, p.rs); ck("concat bwd"); }
                c += p.cols;
            }
        });
        return y;
    }

    // BatchNorm{1,2,3}d in train(): batch statistics over the rows; running statistics updated in place when bound.
    TT batchnorm(const TT& x, const std::string& name, float eps = 1e-5f) {
        const int R = x.rows, C = x.cols;
        TT gamma = param(name + ".weight", 1, C), beta = param(name + ".bias", 1, C);
        float* mean = scratch(C); float* var = scratch(C);
        colred<COL_SUM>(R, C, x.v, x.rs, nullptr, 0, nullptr, nullptr, 0.f, 1.f / (float)R, mean, false);
        colred<COL_SQDEV>(R, C, x.v, x.rs, nullptr, 0, mean, nullptr, 0.f, 1.f / (float)R, var, false);       // biased variance, two passes
        TT y = make(R, C);
        launch(bn_fwd_kernel, ew_blocks(x.numel()), 256, 0, s, R, C, x.v, x.rs, mean, var, eps, gamma.v, beta.v, y.v, y.rs);
        ck("bn fwd");
        if (update_bn_running) {
            auto rm = params->find(name + ".running_mean"), rv = params->find(name + ".running_var");
            if (rm != params->end() && rv != params->end()) {
                launch(bn_running_kernel, (C + 255) / 256, 256, 0, s, C, R, 0.1f, mean, var, rm->second.v, rv->second.v);
                ck("bn running");
            }
        }
        float* dgamma = scratch(C); float* dbeta = scratch(C);
        tape.push_back([=]() {
            colred<COL_BN_DGAMMA>(R, C, x.v, x.rs, y.g, y.rs, mean, var, eps, 1.f, dgamma, false);
            colred<COL_SUM>(R, C, y.g, y.rs, nullptr, 0, nullptr, nullptr, 0.f, 1.f, dbeta, false);
            if (x.g) { launch(bn_bwd_kernel, ew_blocks(x.numel()), 256, 0, s, R, C, x.v, x.rs, mean, var, eps, gamma.v, dgamma, dbeta, y.g, y.rs, x.g, x.rs); ck("bn bwd"); }
            if (gamma.g) { launch(ew_bwd_kernel<EW_COPY>, ew_blocks(C), 256, 0, s, 1, C, nullptr, 0, nullptr, 0, 0.f, dgamma, C, gamma.g, C); ck("bn dgamma acc"); }
            if (beta.g) { launch(ew_bwd_kernel<EW_COPY>, ew_blocks(C), 256, 0, s, 1, C, nullptr, 0, nullptr, 0, 0.f, dbeta, C, beta.g, C); ck("bn dbeta acc"); }
        });
        return y;
    }

    TT softmax(const TT& x) {
        TT y = make(x.rows, x.cols);
        launch(softmax_fwd_kernel, (x.rows * 32 + 255) / 256, 256, 0, s, x.rows, x.cols, x.v, x.rs, y.v, y.rs);
        ck("softmax");
        tape.push_back([=]() {
            if (!x.g) return;
            launch(softmax_bwd_kernel, (x.rows * 32 + 255) / 256, 256, 0, s, x.rows, x.cols, y.v, y.rs, y.g, y.rs, x.g, x.rs);
            ck("softmax bwd");
        });
        return y;
    }
    // scores [B,T] = q [B,D] . Kmem rows (b,t) [B*T, D]
    TT attn_scores(const TT& q, const TT& Km, int T) {
        const int B = q.rows, D = q.cols;
        TT y = make(B, T);
        launch(attn_scores_kernel, (B * T * 32 + 255) / 256, 256, 0, s, B, T, D, q.v, q.rs, Km.v, Km.rs, y.v, y.rs);
        ck("attn scores");
        tape.push_back([=]() {
            launch(attn_scores_bwd_kernel, ew_blocks((size_t)B * D), 256, 0, s, B, T, D, q.v, q.rs, Km.v, Km.rs, y.g, y.rs, q.g, q.rs, Km.g, Km.rs);
            ck("attn scores bwd");
        });
        return y;
    }
    TT attn_context(const TT& a, const TT& V, int T, const TT* dst = nullptr) {
        const int B = a.rows, D = V.cols;
        TT y = dst ? *dst : make(B, D);
        launch(attn_context_kernel, ew_blocks((size_t)B * D), 256, 0, s, B, T, D, a.v, a.rs, V.v, V.rs, y.v, y.rs);
        ck("attn context");
        tape.push_back([=]() {
            launch(attn_context_bwd_kernel, (B * T * 32 + 255) / 256, 256, 0, s, B, T, D, a.v, a.rs, V.v, V.rs, y.g, y.rs, a.g, a.rs, V.g, V.rs);
            ck("attn context bwd");
        });
        return y;
    }

    // One attention read-out of a decoder step in one launch each way (see attn_step_fwd_kernel).  w: the learnable scalar
    // temperature; mask (optional): keep mask [B,T] of the logit dropout with scale alpha; logits (optional): caller-visible
    // copy of the post-dropout scores, row stride ls.
    TT attn_step(const TT& q, const TT& w, const TT& Km, const TT& V, int T, const float* mask, float alpha, float* logits, int ls, const TT* dst = nullptr) {
        const int B = q.rows, D = q.cols, DV = V.cols;
        if (T > 320) throw L2sError(1, "train: attention over more than 320 positions");
        TT y = dst ? *dst : make(B, DV);
        float* sraw = scratch((size_t)B * T); float* probs = scratch((size_t)B * T);
        const bool vec = !(D & 3) && !(DV & 3) && !(q.rs & 3) && !(Km.rs & 3) && !(V.rs & 3) && !(y.rs & 3) &&
                         !((reinterpret_cast<uintptr_t>(q.v) | reinterpret_cast<uintptr_t>(Km.v) | reinterpret_cast<uintptr_t>(V.v) | reinterpret_cast<uintptr_t>(y.v)) & 15);
        if (!vec) throw L2sError(1, "train: attention operands must be 16-byte aligned with multiples of 4 columns");
        launch(attn_step_fwd_kernel, dim3(B, ATT_SPLIT), 256, 0, s, T, D, DV, w.v, q.v, q.rs, Km.v, Km.rs, mask, alpha, V.v, V.rs, sraw, probs, logits, ls, y.v, y.rs);
        ck("attention step");
        tape.push_back([=]() {
            float* dwp = nullptr;
            if (w.g) { dwp = scratch(B); deferred_scalar[w.g].push_back(dwp); deferred_scalar_n[w.g] = B; }
            launch(attn_step_bwd_kernel, dim3(B, ATT_SPLIT), 256, 0, s, T, D, DV, w.v, q.v, q.rs, Km.v, Km.rs, mask, alpha, V.v, V.rs, sraw, probs, y.g, y.rs, q.g, q.rs, Km.g, Km.rs,
                                                   V.g, V.rs, dwp);
            ck("attention step bwd");
        });
        return y;
    }
    // psine (+ dropout) (+ constant) after a few-row linear, one launch each way; the PSine weight gradient is formed in the
    // same backward launch (rows <= 16).
    TT psine_chain(const TT& x, const TT& w, const float* mask, int ms, float p, const float* addc, int as) {
        if (x.rows > 16) {
            TT y = psine(x, w);
            if (mask) y = dropout(y, mask, ms, p);
            if (addc) y = add_const(y, addc, as);
            return y;
        }
        const float alpha = 1.0f / (1.0f - p);
        TT y = make(x.rows, x.cols);
        launch(psine_chain_fwd_kernel, ew_blocks(x.numel()), 256, 0, s, x.rows, x.cols, x.v, x.rs, w.v, mask, ms, alpha, addc, as, y.v, y.rs);
        ck("psine chain");
        tape.push_back([=]() {
            launch(psine_chain_bwd_kernel, (x.cols + 31) / 32, 256, 0, s, x.rows, x.cols, x.v, x.rs, w.v, mask, ms, alpha, y.g, y.rs, x.g, x.rs, w.g);
            ck("psine chain bwd");
        });
        return y;
    }

    // LSTM cell on pre-activation gates [B,4H] and previous cell state; returns (h, c)
    void lstm_cell(const TT& gates, const TT& cprev, TT& h, TT& c, const TT* hdst = nullptr, const TT* cdst = nullptr) {
        const int B = gates.rows, H = gates.cols / 4;
        float* act = scratch((size_t)B * 4 * H);
        h = hdst ? *hdst : make(B, H); c = cdst ? *cdst : make(B, H);
        launch(lstm_cell_fwd_kernel, ew_blocks((size_t)B * H), 256, 0, s, B, H, gates.v, gates.rs, cprev.v, cprev.rs, act, c.v, c.rs, h.v, h.rs);
        ck("lstm cell");
        TT hh = h, cc = c;
        tape.push_back([=]() {
            // gates.g is overwritten (a gates tensor feeds exactly one cell), then flows on through the tape
            launch(lstm_cell_bwd_kernel, ew_blocks((size_t)B * H), 256, 0, s, B, H, act, cprev.v, cprev.rs, cc.v, cc.rs, hh.g, hh.rs, cc.g, cc.rs, gates.g, gates.rs,
                                                                          cprev.g, cprev.rs);
            ck("lstm cell bwd");
        });
    }

    // Conv1d over rows (b,l) x Cin -> rows (b,lo) x Cout; weight [Cout][Cin*K] (the nn.Conv1d layout flattened), bias [Cout]
    TT conv1d(const TT& x, int B, int L, const TT& W, const TT* b, int K, int stride, int pad, int* Lout = nullptr) {
        const int Cin = x.cols, Lo = (L + 2 * pad - K) / stride + 1;
        if (Lout) *Lout = Lo;
        if (K == 1 && stride == 1 && pad == 0) return linear(x, W, b);
        TT col = make(B * Lo, Cin * K);
        launch(im2col1d_kernel, ew_blocks(col.numel()), 256, 0, s, B, L, Lo, Cin, K, stride, pad, x.v, x.rs, col.v);
        ck("im2col");
        tape.push_back([=]() {
            if (!x.g) return;
            launch(col2im1d_kernel, ew_blocks((size_t)B * L * Cin), 256, 0, s, B, L, Lo, Cin, K, stride, pad, col.g, x.g, x.rs);
            ck("col2im");
        });
        return linear(col, W, b);
    }
    // ---- video frontend ops ---------------------------------------------------------------------------------------------------
    // Conv3d stem over the caller's NCDHW clips -> rows (b,t,ho,wo) x 24.  Only the weight gradient exists (the input is data).
    TT stem_conv(const float* video, int B, int T, int H, int W, const TT& Wt) {
        const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
        TT y = make(B * T * Ho * Wo, 24);
        if (!(W & 1) && W <= 96 && Wo <= 48)
            launch(stem_fwd_tiled_kernel, B * T * ((Ho + 7) / 8), 192, STEM_TILED_SMEM, s, B, T, H, W, Ho, Wo, video, Wt.v, y.v);
        else
            launch(stem_fwd_kernel, ew_blocks(y.numel()), 256, 0, s, B, T, H, W, Ho, Wo, video, Wt.v, y.v);
        ck("stem conv");
        tape.push_back([=]() {
            if (!Wt.g) return;
            const size_t npos = (size_t)B * T * Ho * Wo;
            const int chunks = (int)std::min<size_t>(148 * 4, (npos + 255) / 256);
            const int per = (int)((npos + chunks - 1) / chunks);
            float* part = grads.alloc((size_t)chunks * 24 * 735);
            launch(stem_wgrad_kernel, chunks, 256, 0, s, B, T, H, W, Ho, Wo, per, video, y.g, part);
            ck("stem wgrad");
            launch(sum_chunks_kernel, ew_blocks(24 * 735), 256, 0, s, 24 * 735, chunks, part, Wt.g);
            ck("stem wgrad sum");
        });
        return y;
    }
    TT maxpool3x3s2(const TT& x, int N, int H, int W, int* Hout, int* Wout) {
        const int C = x.cols, Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
        *Hout = Ho; *Wout = Wo;
        TT y = make(N * Ho * Wo, C);
        int* idx = reinterpret_cast<int*>(scratch(y.numel()));
        launch(maxpool_fwd_kernel, ew_blocks(y.numel()), 256, 0, s, N, H, W, C, Ho, Wo, x.v, y.v, idx);
        ck("maxpool");
        tape.push_back([=]() {
            if (!x.g) return;
            launch(maxpool_bwd_kernel, ew_blocks(x.numel()), 256, 0, s, N, H, W, C, Ho, Wo, y.g, idx, x.g);
            ck("maxpool bwd");
        });
        return y;
    }
    TT dwconv3x3(const TT& x, int N, int H, int W, int stride, const TT& Wt, int* Hout, int* Wout) {
        const int C = x.cols, Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
        *Hout = Ho; *Wout = Wo;
        TT y = make(N * Ho * Wo, C);
        launch(dw3x3_fwd_kernel, ew_blocks(y.numel()), 256, 0, s, N, H, W, C, stride, Ho, Wo, x.v, x.rs, Wt.v, y.v, y.rs);
        ck("dw conv");
        tape.push_back([=]() {
            if (x.g) { launch(dw3x3_dgrad_kernel, ew_blocks((size_t)N * H * W * C), 256, 0, s, N, H, W, C, stride, Ho, Wo, y.g, y.rs, Wt.v, x.g, x.rs); ck("dw dgrad"); }
            if (Wt.g) {
                const int rows = N * Ho * Wo, colblocks = (C + 31) / 32;
                const int splits = std::max(1, std::min(std::min(1024, (148 * 8 + colblocks - 1) / colblocks), rows / 64));
                float* part = scratch((size_t)splits * 9 * C);
                launch(dw3x3_wgrad_kernel, dim3(colblocks, splits), 256, 0, s, N, H, W, C, stride, Ho, Wo, x.v, x.rs, y.g, y.rs, part);
                ck("dw wgrad");
                launch(dw3x3_wfinish_kernel, (9 * C * 32 + 255) / 256, 256, 0, s, C, splits, part, Wt.g);
                ck("dw wgrad finish");
            }
        });
        return y;
    }
    TT interleave2(const TT& a, const TT& b) {
        const int half = a.cols;
        TT y = make(a.rows, 2 * half);
        launch(interleave2_fwd_kernel, ew_blocks(y.numel()), 256, 0, s, a.rows, half, a.v, a.rs, b.v, b.rs, y.v, y.rs);
        ck("interleave");
        tape.push_back([=]() {
            launch(interleave2_bwd_kernel, ew_blocks(y.numel()), 256, 0, s, a.rows, half, y.g, y.rs, a.g, a.rs, b.g, b.rs);
            ck("interleave bwd");
        });
        return y;
    }
    TT l2normalize(const TT& x) {
        TT y = make(x.rows, x.cols);
        float* nrm = scratch(x.rows);
        launch(l2norm_fwd_kernel, (x.rows * 32 + 255) / 256, 256, 0, s, x.rows, x.cols, x.v, x.rs, y.v, y.rs, nrm);
        ck("l2 normalize");
        tape.push_back([=]() {
            if (!x.g) return;
            launch(l2norm_bwd_kernel, (x.rows * 32 + 255) / 256, 256, 0, s, x.rows, x.cols, y.v, y.rs, nrm, y.g, y.rs, x.g, x.rs);
            ck("l2 normalize bwd");
        });
        return y;
    }
    TT adaptive_pool(const TT& x, int B, int L, int m) {
        if (L == m) return x;
        TT y = make(B * m, x.cols);
        launch(adaptive_pool_fwd_kernel, ew_blocks(y.numel()), 256, 0, s, B, L, m, x.cols, x.v, x.rs, y.v, y.rs);
        ck("adaptive pool");
        tape.push_back([=]() {
            if (!x.g) return;
            launch(adaptive_pool_bwd_kernel, ew_blocks((size_t)B * L * x.cols), 256, 0, s, B, L, m, x.cols, y.g, y.rs, x.g, x.rs);
            ck("adaptive pool bwd");
        });
        return y;
    }
};

}  // namespace tr
}  // namespace l2s
