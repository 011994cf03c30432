// Persistent multi-layer / bidirectional LSTM recurrence (cooperative launch, one CTA per SM).
//
// Restates nn.LSTM (gate order i,f,g,o) for
//   * the decoder's encoder_rnn  : 1 layer, 2 directions, H=512, input projection precomputed
//                                  (reference decoder.py:296,392)
//   * the speaker encoder's lstm : 3 layers, H=256 (reference audio.py:114-119,135); the layers wave-front by themselves
//                                  (layer l, time t starts as soon as layer l-1, time t and its own time t-1 have landed)
// Every CTA owns ONE block of <= 8 hidden units (32 gate rows) of one (layer, direction): its [W_ih | W_hh] rows stay
// resident in shared memory for the whole sequence, and a time step is ONE tensor-core pass (matvec.cuh: mv32 tiles, 3xTF32
// mma.sync over up to 32 clips) + the gate epilogue.
// Exchange (round 2): NO grid barrier.  Hidden states go to a HISTORY buffer hist[time slot][plane][H][Bpad] (slot 0 = initial
// state, slot t+1 = h after step t; every word is written exactly once per launch, so there is no write-after-read hazard
// even though upper layers may lag behind lower ones), and every word carries a "written" flag in its two mantissa LSBs
// (decode3.cuh's flag-carrying exchange; the buffer is zero-filled before the launch): a consumer polls the words it needs
// instead of fence -> grid barrier -> load.  Round 1 (one grid barrier per step): 6.5 us per step.
#pragma once
#include "matvec.cuh"

namespace l2s {

constexpr int LSTM_MAX_UNITS = 8;

struct LstmBlock {
    int layer, dir, u0, nu;      // nu == 0: this CTA has no work (it still takes part in the barriers)
    int K0, K1;                  // segment widths: layer 0 -> K0 = H (h_prev), K1 = 0 ; layer>0 -> K0 = H (below), K1 = H
    int w_off;                   // float offset of this block's [4*nu][K0+K1] rows in the packed weights
    int pad_;
    float bias[4 * LSTM_MAX_UNITS];   // b_ih + b_hh for layer > 0 (layer 0 bias is folded into xproj)
};

struct LstmParams {
    const float* xproj; int ldx;      // [B][T][ldx]; layer-0 gate pre-activations, column dir*4H + g*H + u
    const float* wpk;                 // packed block weights
    const LstmBlock* blocks;          // [grid]
    float* hist;                      // [T+1][L*dirs][H][Bpad] tagged words, zero-filled + slot 0 initialised before the launch
    float* cbuf;                      // [L*dirs][H][Bpad] cell state (private to the owning CTA; initial value set by the caller)
    float* out; int ldo;              // [B][T][ldo] outputs of the LAST layer at column dir*H + u (may be null)
    int T, B, Bpad, H, L, dirs;
    unsigned* abort_word;             // a wait that never ends (a bug) raises it and every CTA stops waiting
};

constexpr uint32_t LSTM_TAG = 1u;

inline size_t lstm_smem_bytes(int H) {
    return ((size_t)4 * LSTM_MAX_UNITS * (2 * H + 16) + MV_WARPS * 16 * MV_CLIPS + 4 * LSTM_MAX_UNITS * MV_CLIPS) * sizeof(float);
}

__device__ __forceinline__ float4 lstm_ld4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool lstm_fresh(const float4& v) {
    return ((__float_as_uint(v.x) & __float_as_uint(v.y) & __float_as_uint(v.z) & __float_as_uint(v.w)) & 3u) == LSTM_TAG &&
           (((__float_as_uint(v.x) | __float_as_uint(v.y) | __float_as_uint(v.z) | __float_as_uint(v.w)) & 3u) == LSTM_TAG);
}
__device__ __forceinline__ void lstm_st(float* p, float v) {
    const uint32_t w = (__float_as_uint(v) & ~3u) | LSTM_TAG;
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(w) : "memory");
}

// One K-segment of the pass: requests the chunks this warp owns (tagged words, layout as mv32_accumulate), waits until all
// of them are written, multiplies.  `dead`: a previous wait gave up (abort), do not wait any more.
template <int RT>
__device__ __forceinline__ void lstm_segment(const float* __restrict__ W, int ldw, int wcol0, int R, const float* __restrict__ X, int K, int ldb, int b0,
                                             float (&acc)[RT][4][4], unsigned* abort_word, bool& dead) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int npw = K / (MV_KC * MV_WARPS);                  // 1 or 2 chunks per warp
    float4 x[2][4];
#pragma unroll
    for (int d = 0; d < 2; ++d)
        if (d < npw) {
            const float* xp = X + (size_t)((warp + d * MV_WARPS) * MV_KC + 4 * t) * ldb + b0 + 4 * g;
#pragma unroll
            for (int i = 0; i < 4; ++i) x[d][i] = lstm_ld4(xp + (size_t)i * ldb);
        }
    unsigned spins = 0;
    bool again = !dead;
    while (again) {
        again = false;
#pragma unroll
        for (int d = 0; d < 2; ++d)
            if (d < npw) {
                const float* xp = X + (size_t)((warp + d * MV_WARPS) * MV_KC + 4 * t) * ldb + b0 + 4 * g;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (!lstm_fresh(x[d][i])) { x[d][i] = lstm_ld4(xp + (size_t)i * ldb); again = true; }
            }
        if (again && (++spins & 1023u) == 0u) {
            if (spins > (1u << 19)) atomicExch(abort_word, 1u);
            if (*reinterpret_cast<volatile unsigned*>(abort_word) != 0u) { dead = true; break; }
        }
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

__global__ void __launch_bounds__(MV_THREADS, 1) lstm_persistent_kernel(const LstmParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    __shared__ LstmBlock bk;
    if (tid == 0) bk = p.blocks[blockIdx.x];
    __syncthreads();
    if (bk.nu == 0) return;                                     // nothing to wait for: there is no barrier to take part in
    const int K = bk.K0 + bk.K1, R = 4 * bk.nu;
    const int ldw = K + 16;                                     // 16 mod 32: conflict-free LDS.128 fragment loads
    float* wsm = smem;                                          // [32][2H + 16]
    float* red = wsm + (size_t)4 * LSTM_MAX_UNITS * (2 * p.H + 16);   // [16 warps][16][32]
    float* gsm = red + MV_WARPS * 16 * MV_CLIPS;                // [32][32]
    for (int i = tid; i < R * (K / 4); i += MV_THREADS) {
        const int r = i / (K / 4), c4 = i - r * (K / 4);
        *reinterpret_cast<float4*>(wsm + (size_t)r * ldw + 4 * c4) = __ldg(reinterpret_cast<const float4*>(p.wpk + bk.w_off + (size_t)r * K) + c4);
    }
    __syncthreads();

    const size_t lstride = (size_t)p.H * p.Bpad;                // one (layer,dir) plane
    const size_t sstride = lstride * p.L * p.dirs;              // one time slot
    const int plane = bk.layer * p.dirs + bk.dir;
    const int RT = (R + 15) / 16;
    bool dead = false;
    for (int ts = 0; ts < p.T; ++ts) {
        const int t = bk.dir ? (p.T - 1 - ts) : ts;
        const float* own_prev = p.hist + (size_t)ts * sstride + plane * lstride;                    // own h after step ts-1 (slot ts)
        const float* below = p.hist + (size_t)(ts + 1) * sstride + (plane - p.dirs) * lstride;      // layer below after ITS step ts
        float* hout = p.hist + (size_t)(ts + 1) * sstride + plane * lstride;
        const float* x0 = (bk.layer == 0) ? own_prev : below;
        for (int b0 = 0; b0 < p.Bpad; b0 += MV_CLIPS) {
            // operands of the epilogue that do not depend on this step's pass are requested first: the layer-0 input
            // projection (or nothing) for this thread's two gate rows, and the cell state of its (unit, clip)
            float xin[2] = {0.f, 0.f};
            if (bk.layer == 0) {
#pragma unroll
                for (int rt = 0; rt < 2; ++rt) {
                    const int r = 16 * rt + (tid >> 5), b = b0 + (tid & 31);
                    if ((r >> 2) < bk.nu && b < p.B)
                        xin[rt] = __ldg(p.xproj + ((size_t)b * p.T + t) * p.ldx + bk.dir * 4 * p.H + (r & 3) * p.H + bk.u0 + (r >> 2));
                }
            }
            float c_prev = 0.f;
            if (tid < bk.nu * MV_CLIPS)
                c_prev = __ldcg(p.cbuf + (size_t)plane * lstride + (size_t)(bk.u0 + (tid >> 5)) * p.Bpad + b0 + (tid & 31));
            float acc[2][4][4];                          // two row tiles (rows >= R re-read row R-1, results unused)
            mv32_zero<2>(acc);
            // the own previous state is one step old, the layer below's is this step's: own first
            if (bk.K1 > 0) lstm_segment<2>(wsm, ldw, bk.K0, R, own_prev, bk.K1, p.Bpad, b0, acc, p.abort_word, dead);
            lstm_segment<2>(wsm, ldw, 0, R, x0, bk.K0, p.Bpad, b0, acc, p.abort_word, dead);
#pragma unroll
            for (int rt = 0; rt < 2; ++rt) {
                if (rt < RT) {
                    float v = mv32_reduce_tile(acc[rt], red);
                    const int r = 16 * rt + (tid >> 5), bb = tid & 31, b = b0 + bb;
                    const int ul = r >> 2, g = r & 3;
                    if (ul < bk.nu && b < p.B) v += (bk.layer == 0) ? xin[rt] : bk.bias[r];
                    gsm[r * MV_CLIPS + bb] = (g == 2) ? tanhf(v) : sigmoidf_acc(v);      // i, f, o: sigmoid; g: tanh
                    __syncthreads();
                }
            }
            for (int o = tid; o < bk.nu * MV_CLIPS; o += MV_THREADS) {
                const int ul = o >> 5, bb = o & 31, b = b0 + bb, r = 4 * ul;
                const float gi = gsm[(r + 0) * MV_CLIPS + bb], gf = gsm[(r + 1) * MV_CLIPS + bb];
                const float gg = gsm[(r + 2) * MV_CLIPS + bb], go = gsm[(r + 3) * MV_CLIPS + bb];
                const int u = bk.u0 + ul;
                const size_t si = (size_t)u * p.Bpad + b;
                const float c = gf * c_prev + gi * gg;
                const float h = go * tanhf(c);
                p.cbuf[(size_t)plane * lstride + si] = c;
                lstm_st(hout + si, h);                     // padding clips too: consumers wait for every word of a 32-clip tile
                if (b < p.B && p.out && bk.layer == p.L - 1) p.out[((size_t)b * p.T + t) * p.ldo + bk.dir * p.H + u] = h;
            }
            __syncthreads();
        }
    }
}

}  // namespace l2s
