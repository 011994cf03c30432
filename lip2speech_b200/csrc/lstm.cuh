// Persistent multi-layer / bidirectional LSTM recurrence (cooperative launch, one CTA per SM).
//
// Restates nn.LSTM (gate order i,f,g,o) for
//   * the decoder's encoder_rnn  : 1 layer, 2 directions, H=512, input projection precomputed
//                                  (reference decoder.py:296,392)
//   * the speaker encoder's lstm : 3 layers wave-fronted (layer l runs time step s-l at global step s),
//                                  H=256 (reference audio.py:114-119,135)
// Every CTA owns ONE block of <= 8 hidden units (32 gate rows) of one (layer, direction): its [W_ih | W_hh] rows stay
// resident in shared memory for the whole sequence, and a time step is ONE tensor-core pass (matvec.cuh: mv32_*, 3xTF32
// mma.sync over up to 32 clips) + the gate epilogue.  Hidden state is exchanged through L2 in feature-major ping-pong
// buffers with one grid barrier per time step.  (First version: chunks of 4 units, up to two FMA passes per CTA and
// step: 8.8 us per step; the step is a latency chain, so one pass per step matters more than its arithmetic.)
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
    float* hbuf;                      // [2][L*dirs][H][Bpad] ping-pong, feature-major
    float* cbuf;                      // [L*dirs][H][Bpad]
    float* out; int ldo;              // [B][T][ldo] outputs of the LAST layer at column dir*H + u (may be null)
    int T, B, Bpad, H, L, dirs;
    unsigned* barrier;
};

inline size_t lstm_smem_bytes(int H) {
    return ((size_t)4 * LSTM_MAX_UNITS * (2 * H + 16) + MV_WARPS * 16 * MV_CLIPS + 4 * LSTM_MAX_UNITS * MV_CLIPS) * sizeof(float);
}

__global__ void __launch_bounds__(MV_THREADS, 1) lstm_persistent_kernel(const LstmParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    __shared__ LstmBlock bk;
    if (tid == 0) bk = p.blocks[blockIdx.x];
    __syncthreads();
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

    unsigned target = 0;
    const size_t lstride = (size_t)p.H * p.Bpad;                // one (layer,dir) plane
    const size_t pstride = lstride * p.L * p.dirs;              // one parity
    const int nsteps = p.T + p.L - 1;
    const int plane = bk.layer * p.dirs + bk.dir;
    const int RT = (R + 15) / 16;
    for (int s = 0; s < nsteps; ++s) {
        const float* hcur = p.hbuf + (size_t)(s & 1) * pstride;
        float* hnext = p.hbuf + (size_t)((s + 1) & 1) * pstride;
        const int ts = s - bk.layer;
        if (bk.nu > 0 && ts >= 0 && ts < p.T) {                  // CTA-uniform
            const int t = bk.dir ? (p.T - 1 - ts) : ts;
            const float* x0 = (bk.layer == 0) ? hcur + plane * lstride : hcur + (plane - p.dirs) * lstride;
            const float* x1 = hcur + plane * lstride;
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
                if (tid < bk.nu * MV_CLIPS && b0 + (tid & 31) < p.B)
                    c_prev = __ldcg(p.cbuf + (size_t)plane * lstride + (size_t)(bk.u0 + (tid >> 5)) * p.Bpad + b0 + (tid & 31));
                float acc[2][4][4];                          // two row tiles (rows >= R re-read row R-1, results unused)
                mv32_zero<2>(acc);
                mv32_accumulate<2>(wsm, ldw, 0, R, x0, bk.K0, p.Bpad, b0, acc);
                if (bk.K1 > 0) mv32_accumulate<2>(wsm, ldw, bk.K0, R, x1, bk.K1, p.Bpad, b0, acc);
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
                    if (b < p.B) {
                        const float gi = gsm[(r + 0) * MV_CLIPS + bb], gf = gsm[(r + 1) * MV_CLIPS + bb];
                        const float gg = gsm[(r + 2) * MV_CLIPS + bb], go = gsm[(r + 3) * MV_CLIPS + bb];
                        const int u = bk.u0 + ul;
                        const size_t si = (size_t)plane * lstride + (size_t)u * p.Bpad + b;
                        const float c = gf * c_prev + gi * gg;
                        const float h = go * tanhf(c);
                        p.cbuf[si] = c;
                        hnext[si] = h;
                        if (p.out && bk.layer == p.L - 1) p.out[((size_t)b * p.T + t) * p.ldo + bk.dir * p.H + u] = h;
                    }
                }
                __syncthreads();
            }
        } else if (bk.nu > 0) {
            // a layer that is idle this step carries its state across the ping-pong
            for (int i = tid; i < bk.nu * p.Bpad; i += MV_THREADS) {
                const size_t si = (size_t)plane * lstride + (size_t)(bk.u0 + i / p.Bpad) * p.Bpad + (i % p.Bpad);
                hnext[si] = ldcg1(hcur + si);
            }
        }
        grid_barrier(p.barrier, target, gridDim.x);
    }
}

}  // namespace l2s
