// Persistent autoregressive decode loop: ALL decoder steps of Decoder.inference in ONE cooperative
// kernel (reference decoder.py:403-435; one step = SURVEY.md §3.4).
//
// Design (B200-first, not a translation of the reference's ~25 library calls + 1 host sync per step):
//  * The 5.4 M step weights (prenet, Q, content.Q, attention_proj, 2 LSTM cells, fc_out, stop) are
//    partitioned BY OUTPUT ROW over the 148 SMs and stay resident in shared memory for all 300 steps:
//    HBM weight traffic is paid once per batch instead of once per step.
//  * Activations are exchanged through L2 in feature-major [feature][clip] buffers; every dependent
//    layer boundary is one grid barrier (5 per step):
//        A : fc_out -> mel frame, stop token ; prenet layer 1 (fused with fc_out: W_p1*W_fc) ;
//            Q = PSine(W_q [h0;h1]) + pos[i+1] ; content query = SiLU(W_cq [c0;c1])
//        B : prenet layer 2 ; per-clip dot-product attention over T (K,V) and over the content slots
//        C : x2 = prenet2 + attention_proj(ctx)
//        D : LSTM layer 0 (gates = [W_ih|W_hh] [cv;x2;h0] + b) -> h0', c0'
//        E : LSTM layer 1 -> h1', c1'
//  * Which CTA owns which rows is decided on the host (api.cu: pack_decode) and handed over as a list
//    of `DecPass` descriptors, so load-balancing policy is not baked into the kernel.
//  * Stop-token bookkeeping (output_lengths) happens on the device; there is no host sync in the loop.
#pragma once
#include "matvec.cuh"

namespace l2s {

enum DecOp { OP_NONE = 0, OP_FC, OP_P1, OP_STOP, OP_Q, OP_CQ, OP_P2, OP_X2, OP_GATE0, OP_GATE1 };
enum DecSrc { SRC_NONE = 0, SRC_H1NEW, SRC_HNEW, SRC_C, SRC_P1, SRC_CTX, SRC_XD, SRC_H0OLD, SRC_H0NEW, SRC_H1OLD };
enum DecStage { ST_A = 0, ST_B, ST_C, ST_D, ST_E, ST_COUNT };

struct DecPass {
    int stage, R, K0, K1, src0, src1, w_off, pad_;
    int op[16];
    int idx[16];
    float bias[16];
    float aux[16];     // PSine w
    float aux2[16];    // OP_P1: prenet layer-1 output for the BOS frame (step 0 input)
};

constexpr int DEC_MAX_PASSES = 10;

struct DecodeParams {
    // per-CTA program
    const DecPass* passes;        // [grid][DEC_MAX_PASSES]
    const int* npasses;           // [grid]
    const float* wimg;            // [grid][wimg_floats] shared-memory weight image per CTA
    int wimg_floats;
    // state / activations (feature-major, ld = Bpad)
    float* S;                     // [2][1024][Bpad]  (h0 rows 0..511, h1 rows 512..1023), parity ping-pong
    float* Cst;                   // [1024][Bpad]     (c0, c1)
    float* P1; float* P2; float* Q; float* CQ; float* CTX; float* XD;
    // per-clip memories
    const float* Kmem; const float* Vmem;      // [B][T][512]
    const float* ckey; const float* cval;      // [B][minT][256]
    const float* stop_const;                   // [B]  W_stop[512:1024].enc_cell + b_stop
    const float* pos;                          // [300][512]
    float temp, ctemp;
    float* outputs;               // [B][steps][80]
    long long* lengths;           // [B]
    float* attn;                  // [B][steps][T] or null
    int B, Bpad, T, minT, steps;
    unsigned* barrier;
};

__device__ __forceinline__ const float* dec_src(const DecodeParams& p, int src, int parity_new) {
    const size_t plane = (size_t)1024 * p.Bpad;
    const float* Snew = p.S + (size_t)parity_new * plane;
    const float* Sold = p.S + (size_t)(parity_new ^ 1) * plane;
    switch (src) {
        case SRC_H1NEW: return Snew + (size_t)512 * p.Bpad;
        case SRC_HNEW: return Snew;
        case SRC_C: return p.Cst;
        case SRC_P1: return p.P1;
        case SRC_CTX: return p.CTX;
        case SRC_XD: return p.XD;
        case SRC_H0OLD: return Sold;
        case SRC_H0NEW: return Snew;
        case SRC_H1OLD: return Sold + (size_t)512 * p.Bpad;
        default: return nullptr;
    }
}

// Dot-product attention over the T encoder positions and over the minT content slots for one clip
// (reference decoder.py:414-419 and Content.forward 262-271).
__device__ void dec_attend_clip(const DecodeParams& p, int b, int step, float* qs, float* sc, float* cqs, float* csc) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    qs[tid] = ldcg1(p.Q + (size_t)tid * p.Bpad + b) * p.temp;
    if (tid < 256) cqs[tid] = ldcg1(p.CQ + (size_t)tid * p.Bpad + b) * p.ctemp;
    __syncthreads();
    for (int t = warp; t < p.T; t += MV_WARPS) {
        const float4* kr = reinterpret_cast<const float4*>(p.Kmem + ((size_t)b * p.T + t) * 512);
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 k = __ldg(kr + lane + 32 * i);
            const float4 q = *reinterpret_cast<const float4*>(qs + 4 * (lane + 32 * i));
            a = fmaf(q.x, k.x, a); a = fmaf(q.y, k.y, a); a = fmaf(q.z, k.z, a); a = fmaf(q.w, k.w, a);
        }
        a = warp_sum(a);
        if (lane == 0) sc[t] = a;
    }
    for (int m = warp; m < p.minT; m += MV_WARPS) {
        const float4* kr = reinterpret_cast<const float4*>(p.ckey + ((size_t)b * p.minT + m) * 256);
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float4 k = __ldg(kr + lane + 32 * i);
            const float4 q = *reinterpret_cast<const float4*>(cqs + 4 * (lane + 32 * i));
            a = fmaf(q.x, k.x, a); a = fmaf(q.y, k.y, a); a = fmaf(q.z, k.z, a); a = fmaf(q.w, k.w, a);
        }
        a = warp_sum(a);
        if (lane == 0) csc[m] = a;
    }
    __syncthreads();
    if (warp == 0) {
        float mx = -INFINITY;
        for (int t = lane; t < p.T; t += 32) mx = fmaxf(mx, sc[t]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int t = lane; t < p.T; t += 32) { float e = expf(sc[t] - mx); sc[t] = e; sum += e; }
        sum = warp_sum(sum);
        for (int t = lane; t < p.T; t += 32) {
            float a = sc[t] / sum;
            sc[t] = a;
            if (p.attn) p.attn[((size_t)b * p.steps + step) * p.T + t] = a;
        }
    } else if (warp == 1) {
        float v = lane < p.minT ? csc[lane] : -INFINITY;
        float mx = warp_max(v);
        float e = lane < p.minT ? expf(v - mx) : 0.f;
        float sum = warp_sum(e);
        if (lane < p.minT) csc[lane] = e / sum;
    }
    __syncthreads();
    {
        const float* vr = p.Vmem + (size_t)b * p.T * 512 + tid;
        float a = 0.f;
        for (int t = 0; t < p.T; ++t) a = fmaf(sc[t], __ldg(vr + (size_t)t * 512), a);
        p.CTX[(size_t)tid * p.Bpad + b] = a;
    }
    if (tid < 256) {
        const float* vr = p.cval + (size_t)b * p.minT * 256 + tid;
        float a = 0.f;
        for (int m = 0; m < p.minT; ++m) a = fmaf(csc[m], __ldg(vr + (size_t)m * 256), a);
        p.XD[(size_t)tid * p.Bpad + b] = a;
    }
    __syncthreads();
}

template <int R>
__device__ __forceinline__ void dec_run_pass(const DecodeParams& p, const DecPass& ps, const float* wsm, float* red, float* gsm,
                                             int step, int parity_new) {
    const int tid = threadIdx.x;
    Seg s0, s1;
    s0.x = dec_src(p, ps.src0, parity_new); s0.K = ps.K0;
    s1.x = dec_src(p, ps.src1, parity_new); s1.K = ps.K1;
    const size_t plane = (size_t)1024 * p.Bpad;
    float* Snew = p.S + (size_t)parity_new * plane;
    for (int b0 = 0; b0 < p.Bpad; b0 += MV_CLIPS) {
        float v = mv_pass<R>(wsm + ps.w_off, ps.K0 + ps.K1, s0, s1, p.Bpad, b0, red);
        const int r = tid >> 5, bb = tid & 31, b = b0 + bb;
        const bool live = (tid < R * MV_CLIPS) && (b < p.B);
        const int op = live ? ps.op[r] : OP_NONE;
        const int idx = live ? ps.idx[r] : 0;
        if (live) v += ps.bias[r];
        bool gate_pass = false;
        switch (op) {
            case OP_FC:
                if (step >= 0) p.outputs[((size_t)b * p.steps + step) * 80 + idx] = v;
                break;
            case OP_P1:
                p.P1[(size_t)idx * p.Bpad + b] = (step >= 0) ? sinf(v) * ps.aux[r] : ps.aux2[r];
                break;
            case OP_STOP:
                if (step >= 0 && (v + p.stop_const[b]) > 0.f && p.lengths[b] == (long long)p.steps) p.lengths[b] = step + 1;
                break;
            case OP_Q: {
                float q = sinf(v) * ps.aux[r];
                if (step + 1 < p.steps) q += __ldg(p.pos + (size_t)(step + 1) * 512 + idx);
                p.Q[(size_t)idx * p.Bpad + b] = q;
            } break;
            case OP_CQ:
                p.CQ[(size_t)idx * p.Bpad + b] = siluf_acc(v);
                break;
            case OP_P2:
                p.P2[(size_t)idx * p.Bpad + b] = sinf(v) * ps.aux[r];
                break;
            case OP_X2:
                p.XD[(size_t)(256 + idx) * p.Bpad + b] = v + ldcg1(p.P2 + (size_t)idx * p.Bpad + b);
                break;
            default: break;
        }
        // LSTM passes: rows are (unit, gate) = (r>>2, r&3); all 16 rows of the pass are gate rows.
        if (ps.op[0] == OP_GATE0 || ps.op[0] == OP_GATE1) {
            gate_pass = true;
            if (tid < R * MV_CLIPS) gsm[r * MV_CLIPS + bb] = v;
        }
        __syncthreads();
        if (gate_pass && live && (r & 3) == 0 && idx >= 0) {
            const int layer = (op == OP_GATE1) ? 1 : 0;
            const float gi = gsm[(r + 0) * MV_CLIPS + bb], gf = gsm[(r + 1) * MV_CLIPS + bb];
            const float gg = gsm[(r + 2) * MV_CLIPS + bb], go = gsm[(r + 3) * MV_CLIPS + bb];
            const size_t si = (size_t)(layer * 512 + idx) * p.Bpad + b;
            const float c = sigmoidf_acc(gf) * p.Cst[si] + sigmoidf_acc(gi) * tanhf(gg);
            const float h = sigmoidf_acc(go) * tanhf(c);
            p.Cst[si] = c;
            Snew[si] = h;
        }
        if (gate_pass) __syncthreads();
    }
}

__device__ __forceinline__ void dec_dispatch(const DecodeParams& p, const DecPass& ps, const float* wsm, float* red, float* gsm,
                                             int step, int parity_new) {
    if (ps.R == 16) dec_run_pass<16>(p, ps, wsm, red, gsm, step, parity_new);
    else if (ps.R == 8) dec_run_pass<8>(p, ps, wsm, red, gsm, step, parity_new);
    else dec_run_pass<4>(p, ps, wsm, red, gsm, step, parity_new);
}

__global__ void __launch_bounds__(MV_THREADS, 1) decode_persistent_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) float smem[];
    float* wsm = smem;                                   // weight image
    float* red = wsm + p.wimg_floats;                    // [16 warps][16][32]
    float* gsm = red + MV_WARPS * 16 * MV_CLIPS;         // [16][32]
    float* qs = gsm + 16 * MV_CLIPS;                     // [512]
    float* sc = qs + 512;                                // [320]
    float* cqs = sc + 320;                               // [256]
    float* csc = cqs + 256;                              // [32]
    __shared__ DecPass passes[DEC_MAX_PASSES];
    __shared__ int np;

    const int tid = threadIdx.x;
    if (tid == 0) np = p.npasses[blockIdx.x];
    {
        const int* src = reinterpret_cast<const int*>(p.passes + (size_t)blockIdx.x * DEC_MAX_PASSES);
        int* dst = reinterpret_cast<int*>(passes);
        for (int i = tid; i < (int)(sizeof(DecPass) * DEC_MAX_PASSES / 4); i += MV_THREADS) dst[i] = src[i];
        const float4* wsrc = reinterpret_cast<const float4*>(p.wimg + (size_t)blockIdx.x * p.wimg_floats);
        float4* wdst = reinterpret_cast<float4*>(wsm);
        for (int i = tid; i < p.wimg_floats / 4; i += MV_THREADS) wdst[i] = __ldg(wsrc + i);
    }
    __syncthreads();

    unsigned target = 0;
    // pre-stage A(-1): Q, content query and prenet(BOS) from the initial state in S[0]
    for (int j = 0; j < np; ++j)
        if (passes[j].stage == ST_A) dec_dispatch(p, passes[j], wsm, red, gsm, -1, 0);
    grid_barrier(p.barrier, target, gridDim.x);

    for (int step = 0; step < p.steps; ++step) {
        const int parity_new = (step + 1) & 1;
        // ---- B: prenet layer 2 + attention ----
        for (int j = 0; j < np; ++j)
            if (passes[j].stage == ST_B) dec_dispatch(p, passes[j], wsm, red, gsm, step, parity_new);
        for (int b = blockIdx.x; b < p.B; b += gridDim.x) dec_attend_clip(p, b, step, qs, sc, cqs, csc);
        grid_barrier(p.barrier, target, gridDim.x);
        // ---- C, D, E, A ----
#pragma unroll 1
        for (int st = ST_C; st <= ST_E + 1; ++st) {
            const int stage = (st == ST_E + 1) ? ST_A : st;
            for (int j = 0; j < np; ++j)
                if (passes[j].stage == stage) dec_dispatch(p, passes[j], wsm, red, gsm, step, parity_new);
            grid_barrier(p.barrier, target, gridDim.x);
        }
    }
}

}  // namespace l2s
