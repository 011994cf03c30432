// Persistent multi-layer / bidirectional LSTM recurrence (cooperative launch, one CTA per SM).
//
// Restates nn.LSTM (gate order i,f,g,o) for
//   * the decoder's encoder_rnn  : 1 layer, 2 directions, H=512, input projection precomputed
//                                  (reference decoder.py:296,392)
//   * the speaker encoder's lstm : 3 layers wave-fronted (layer l runs time step s-l at global step s),
//                                  H=256 (reference audio.py:114-119,135)
// Hidden units are partitioned over the CTAs in chunks of <=4 units (16 gate rows); each chunk's
// [W_ih | W_hh] rows stay resident in shared memory for the whole sequence.  Hidden state is exchanged
// through L2 in feature-major ping-pong buffers with one grid barrier per time step.
#pragma once
#include "matvec.cuh"

namespace l2s {

struct LstmChunk {
    int layer, dir, u0, nu;
    int K0, K1;          // segment widths: layer 0 -> K0 = H (h_prev), K1 = 0 ; layer>0 -> K0 = H (below), K1 = H
    int w_off;           // float offset of this chunk's [16][K0+K1] block in the packed weights
    int pad_;
    float bias[16];      // b_ih + b_hh for layer > 0 (layer 0 bias is folded into xproj)
};

struct LstmParams {
    const float* xproj; int ldx;      // [B][T][ldx]; layer-0 gate pre-activations, column dir*4H + g*H + u
    const float* wpk;                 // packed chunk weights
    const LstmChunk* chunks; int nchunks;
    float* hbuf;                      // [2][L*dirs][H][Bpad] ping-pong, feature-major
    float* cbuf;                      // [L*dirs][H][Bpad]
    float* out; int ldo;              // [B][T][ldo] outputs of the LAST layer at column dir*H + u (may be null)
    int T, B, Bpad, H, L, dirs;
    unsigned* barrier;
    int max_chunks;                   // chunks per CTA upper bound (smem sizing)
};

constexpr int LSTM_MAX_CHUNKS = 2;

__global__ void __launch_bounds__(MV_THREADS, 1) lstm_persistent_kernel(const LstmParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int Kmax = 2 * p.H;
    float* wsm = smem;                                          // [max_chunks][16][Kmax]
    float* red = wsm + (size_t)p.max_chunks * 16 * Kmax;        // [16 warps][16][32]
    float* gsm = red + MV_WARPS * 16 * MV_CLIPS;                // [16][32]
    __shared__ LstmChunk cks[LSTM_MAX_CHUNKS];
    __shared__ int nmine;

    if (tid == 0) {
        int n = 0;
        for (int c = blockIdx.x; c < p.nchunks && n < LSTM_MAX_CHUNKS; c += gridDim.x) cks[n++] = p.chunks[c];
        nmine = n;
    }
    __syncthreads();
    for (int j = 0; j < nmine; ++j) {
        const int K = cks[j].K0 + cks[j].K1;
        const float4* src = reinterpret_cast<const float4*>(p.wpk + cks[j].w_off);
        float4* dst = reinterpret_cast<float4*>(wsm + (size_t)j * 16 * Kmax);
        for (int i = tid; i < 16 * K / 4; i += MV_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    unsigned target = 0;
    const size_t lstride = (size_t)p.H * p.Bpad;                // one (layer,dir) plane
    const size_t pstride = lstride * p.L * p.dirs;              // one parity
    const int nsteps = p.T + p.L - 1;
    for (int s = 0; s < nsteps; ++s) {
        const float* hcur = p.hbuf + (size_t)(s & 1) * pstride;
        float* hnext = p.hbuf + (size_t)((s + 1) & 1) * pstride;
        for (int j = 0; j < nmine; ++j) {
            const LstmChunk& ck = cks[j];
            const int ts = s - ck.layer;
            if (ts < 0 || ts >= p.T) continue;                  // CTA-uniform
            const int t = ck.dir ? (p.T - 1 - ts) : ts;
            const int plane = ck.layer * p.dirs + ck.dir;
            const int K = ck.K0 + ck.K1;
            Seg s0, s1;
            if (ck.layer == 0) { s0.x = hcur + plane * lstride; s0.K = ck.K0; s1.x = nullptr; s1.K = 0; }
            else { s0.x = hcur + (plane - p.dirs) * lstride; s0.K = ck.K0; s1.x = hcur + plane * lstride; s1.K = ck.K1; }
            for (int b0 = 0; b0 < p.Bpad; b0 += MV_CLIPS) {
                float v = mv_pass<16>(wsm + (size_t)j * 16 * Kmax, K, s0, s1, p.Bpad, b0, red);
                const int r = tid >> 5, bb = tid & 31, b = b0 + bb;
                const int ul = r >> 2, g = r & 3, u = ck.u0 + ul;
                if (ul < ck.nu && b < p.B) {
                    if (ck.layer == 0) v += __ldg(p.xproj + ((size_t)b * p.T + t) * p.ldx + ck.dir * 4 * p.H + g * p.H + u);
                    else v += ck.bias[r];
                }
                gsm[r * MV_CLIPS + bb] = v;
                __syncthreads();
                if (g == 0 && ul < ck.nu && b < p.B) {
                    const float gi = gsm[(r + 0) * MV_CLIPS + bb], gf = gsm[(r + 1) * MV_CLIPS + bb];
                    const float gg = gsm[(r + 2) * MV_CLIPS + bb], go = gsm[(r + 3) * MV_CLIPS + bb];
                    const size_t si = (size_t)plane * lstride + (size_t)u * p.Bpad + b;
                    const float c = sigmoidf_acc(gf) * p.cbuf[si] + sigmoidf_acc(gi) * tanhf(gg);
                    const float h = sigmoidf_acc(go) * tanhf(c);
                    p.cbuf[si] = c;
                    hnext[si] = h;
                    if (p.out && ck.layer == p.L - 1) p.out[((size_t)b * p.T + t) * p.ldo + ck.dir * p.H + u] = h;
                }
                __syncthreads();
            }
        }
        // layers that are idle this step must carry their state across the ping-pong
        for (int j = 0; j < nmine; ++j) {
            const LstmChunk& ck = cks[j];
            const int ts = s - ck.layer;
            if (ts >= 0 && ts < p.T) continue;
            const int plane = ck.layer * p.dirs + ck.dir;
            for (int i = tid; i < ck.nu * p.Bpad; i += MV_THREADS) {
                const size_t si = (size_t)plane * lstride + (size_t)(ck.u0 + i / p.Bpad) * p.Bpad + (i % p.Bpad);
                hnext[si] = ldcg1(hcur + si);
            }
        }
        grid_barrier(p.barrier, target, gridDim.x);
    }
}

}  // namespace l2s
