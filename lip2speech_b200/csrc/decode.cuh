// Persistent autoregressive decode loop: ALL decoder steps of Decoder.inference in ONE cooperative
// kernel (reference decoder.py:403-435; one step = SURVEY.md §3.4).
//
// Design (B200-first, not a translation of the reference's ~25 library calls + 1 host sync per step):
//  * The step weights (prenet, Q, content.Q, attention_proj, 2 LSTM cells, fc_out, stop) are partitioned
//    BY OUTPUT ROW over the 148 SMs and stay resident in shared memory for all 300 steps: HBM weight
//    traffic is paid once per batch instead of once per step.
//  * Activations are exchanged through L2 in feature-major [feature][clip] buffers.  Every dependent layer
//    boundary is one grid barrier; linear∘linear pairs are merged on the host so that only FOUR remain:
//        A : fc_out -> mel frame, stop token ; prenet-1 (fused with fc_out: W_p1·W_fc) ;
//            Q = PSine(W_q [h0;h1]) + pos[i+1] ; content query = SiLU(W_cq [c0;c1])
//        B : prenet-2 ; per-clip dot-product attention over T (K,V) and over the content slots
//        D : LSTM-0 with attention_proj folded in: gates = [W_a|W_b|W_b·W_ap|W_hh] [cv;p2;ctx;h0] + b'
//        E : LSTM-1: gates = [W_ih|W_hh] [h0';h1] + b
//  * Barriers are split-phase (arrive / wait).  Each pass has an EARLY segment whose operands were produced
//    two or more stages ago (h_prev for the LSTMs, h0'/c0' for the queries) and a LATE segment that needs
//    the previous stage: the early FMAs run between arrive and wait and hide the barrier latency.
//  * Which CTA owns which rows is decided on the host (pack.h: pack_decode_program) and handed over as a
//    list of `DecPass` descriptors, so load-balancing policy is not baked into the kernel.
//  * Stop-token bookkeeping (output_lengths) happens on the device; there is no host sync in the loop.
#pragma once
#include "matvec.cuh"

namespace l2s {

enum DecOp { OP_NONE = 0, OP_FC, OP_P1, OP_STOP, OP_Q, OP_CQ, OP_P2, OP_GATE0, OP_GATE1 };
enum DecSrc { SRC_NONE = 0, SRC_H0NEW, SRC_H1NEW, SRC_C0, SRC_C1, SRC_P1, SRC_XD, SRC_H0OLD, SRC_H1OLD };
enum DecStage { ST_A = 0, ST_B, ST_D, ST_E, ST_COUNT };

struct DecPass {
    int stage, R;
    int Ke, src_e, wcol_e;     // early segment (may be empty)
    int Kl, src_l, wcol_l;     // late segment
    int ldw, w_off, pad0_, pad1_;
    int op[16];
    int idx[16];
    float bias[16];
    float aux[16];     // PSine w
    float aux2[16];    // OP_P1: prenet layer-1 output for the BOS frame (step 0 input)
};

constexpr int DEC_MAX_PASSES = 8;
constexpr int DEC_RED_ROWS = 8;           // cross-warp reduction buffer holds 8 rows (R=16 passes reduce in two rounds)
constexpr int DEC_TIMING_SLOTS = 12;      // per stage: early compute, barrier wait, late compute

struct DecodeParams {
    // per-CTA program
    const DecPass* passes;        // [grid][DEC_MAX_PASSES]
    const int* npasses;           // [grid]
    const float* wimg;            // [grid][wimg_floats] shared-memory weight image per CTA
    int wimg_floats;
    // state / activations (feature-major, ld = Bpad)
    float* S;                     // [2][1024][Bpad]  (h0 rows 0..511, h1 rows 512..1023), parity ping-pong
    float* Cst;                   // [1024][Bpad]     (c0, c1)
    float* P1; float* Q; float* CQ;
    float* XD;                    // [1024][Bpad]: content value cv (0..255), prenet-2 (256..511), attention ctx (512..1023)
    // per-clip memories
    const float* Kmem; const float* Vmem;      // [B][T][512]
    const float* ckey; const float* cval;      // [B][minT][256]
    const float* stop_const;                   // [B]  W_stop[512:1024].enc_cell
    const float* pos;                          // [300][512]
    float temp, ctemp;
    float* outputs;               // [B][steps][80]
    long long* lengths;           // [B]
    float* attn;                  // [B][steps][T] or null
    int B, Bpad, T, minT, steps;
    int nsplit;                   // CTAs cooperating on one clip's attention (1, 2 or 4)
    unsigned* barrier;
    float* timing;                // optional [grid][DEC_TIMING_SLOTS] SM cycles
    // Decoder.forward flavour (decoder.py:353-375): teacher forcing and per-step logits; all null for inference
    const unsigned char* tf_mask; // [steps]: 1 = step i consumes the teacher frame (BOS for i=0, mels[:, i-1] otherwise)
    const float* p1_teacher;      // [steps][256][Bpad] prenet layer-1 output of the teacher frames
    float* stop_out;              // [B][steps] raw stop logits
    float* attn_logits;           // [B][steps][T] PRE-softmax attention logits
};

struct DecSmem {
    float* wsm; float* red; float* gsm; float* qs; float* sc; float* cqs; float* csc;
};

__device__ __forceinline__ const float* dec_src(const DecodeParams& p, int src, int parity_new) {
    const size_t plane = (size_t)1024 * p.Bpad;
    const float* Snew = p.S + (size_t)parity_new * plane;
    const float* Sold = p.S + (size_t)(parity_new ^ 1) * plane;
    switch (src) {
        case SRC_H0NEW: return Snew;
        case SRC_H1NEW: return Snew + (size_t)512 * p.Bpad;
        case SRC_C0: return p.Cst;
        case SRC_C1: return p.Cst + (size_t)512 * p.Bpad;
        case SRC_P1: return p.P1;
        case SRC_XD: return p.XD;
        case SRC_H0OLD: return Sold;
        case SRC_H1OLD: return Sold + (size_t)512 * p.Bpad;
        default: return nullptr;
    }
}

struct StageSync {
    const unsigned* counter; unsigned target; bool waited;
    float* tacc; long long tmark; int slot0; bool timing;
    __device__ __forceinline__ void lap(int slot) {
        if (timing && threadIdx.x == 0) { long long now = clock64(); tacc[slot] += (float)(now - tmark); tmark = now; }
    }
    // Block until every CTA has finished the previous stage (idempotent within a stage).
    __device__ __forceinline__ void wait() {
        if (!waited) {
            lap(slot0);
            grid_wait(counter, target);
            lap(slot0 + 1);
            waited = true;
        }
    }
};

// Dot-product attention over the T encoder positions and over the minT content slots for one clip
// (reference decoder.py:414-419 and Content.forward 262-271).  `nsplit` CTAs share a clip: each recomputes
// the (cheap) scores and produces its 512/nsplit slice of ctx and 256/nsplit slice of the content value.
__device__ void dec_attend(const DecodeParams& p, const DecSmem& sm, int b, int part, int step) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    sm.qs[tid] = ldcg1(p.Q + (size_t)tid * p.Bpad + b) * p.temp;
    if (tid < 256) sm.cqs[tid] = ldcg1(p.CQ + (size_t)tid * p.Bpad + b) * p.ctemp;
    __syncthreads();
    for (int t0 = warp; t0 < p.T; t0 += 2 * MV_WARPS) {
        const int t1 = t0 + MV_WARPS;
        const float4* kr0 = reinterpret_cast<const float4*>(p.Kmem + ((size_t)b * p.T + t0) * 512);
        const float4* kr1 = reinterpret_cast<const float4*>(p.Kmem + ((size_t)b * p.T + (t1 < p.T ? t1 : t0)) * 512);
        float4 k0[4], k1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { k0[i] = __ldg(kr0 + lane + 32 * i); k1[i] = __ldg(kr1 + lane + 32 * i); }
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 q = *reinterpret_cast<const float4*>(sm.qs + 4 * (lane + 32 * i));
            a0 = fmaf(q.x, k0[i].x, a0); a0 = fmaf(q.y, k0[i].y, a0); a0 = fmaf(q.z, k0[i].z, a0); a0 = fmaf(q.w, k0[i].w, a0);
            a1 = fmaf(q.x, k1[i].x, a1); a1 = fmaf(q.y, k1[i].y, a1); a1 = fmaf(q.z, k1[i].z, a1); a1 = fmaf(q.w, k1[i].w, a1);
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1);
        if (lane == 0) { sm.sc[t0] = a0; if (t1 < p.T) sm.sc[t1] = a1; }
    }
    for (int m = warp; m < p.minT; m += MV_WARPS) {
        const float4* kr = reinterpret_cast<const float4*>(p.ckey + ((size_t)b * p.minT + m) * 256);
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float4 k = __ldg(kr + lane + 32 * i);
            const float4 q = *reinterpret_cast<const float4*>(sm.cqs + 4 * (lane + 32 * i));
            a = fmaf(q.x, k.x, a); a = fmaf(q.y, k.y, a); a = fmaf(q.z, k.z, a); a = fmaf(q.w, k.w, a);
        }
        a = warp_sum(a);
        if (lane == 0) sm.csc[m] = a;
    }
    __syncthreads();
    if (warp == 0) {
        if (p.attn_logits && part == 0)
            for (int t = lane; t < p.T; t += 32) p.attn_logits[((size_t)b * p.steps + step) * p.T + t] = sm.sc[t];
        float mx = -INFINITY;
        for (int t = lane; t < p.T; t += 32) mx = fmaxf(mx, sm.sc[t]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int t = lane; t < p.T; t += 32) { float e = expf(sm.sc[t] - mx); sm.sc[t] = e; sum += e; }
        sum = warp_sum(sum);
        for (int t = lane; t < p.T; t += 32) {
            float a = sm.sc[t] / sum;
            sm.sc[t] = a;
            if (p.attn && part == 0) p.attn[((size_t)b * p.steps + step) * p.T + t] = a;
        }
    } else if (warp == 1) {
        float v = lane < p.minT ? sm.csc[lane] : -INFINITY;
        float mx = warp_max(v);
        float e = lane < p.minT ? expf(v - mx) : 0.f;
        float sum = warp_sum(e);
        if (lane < p.minT) sm.csc[lane] = e / sum;
    }
    __syncthreads();
    // ctx slice in chunks of 128 features: thread (fq = tid&31, tg = tid>>5) sums its t-group, then the 16
    // t-groups are combined in a fixed order (independent of nsplit, so results do not depend on B).
    const int FW = 512 / p.nsplit;
    float* part_smem = sm.red;                               // [16][128]
    for (int fc = 0; fc < FW; fc += 128) {
        const int f0 = part * FW + fc;
        const int fq = tid & 31, tg = tid >> 5;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = tg; t < p.T; t += MV_WARPS) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.Vmem + ((size_t)b * p.T + t) * 512 + f0) + fq);
            const float w = sm.sc[t];
            a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
        }
        *reinterpret_cast<float4*>(part_smem + tg * 128 + fq * 4) = a;
        __syncthreads();
        if (tid < 128) {
            float s = 0.f;
#pragma unroll
            for (int g = 0; g < MV_WARPS; ++g) s += part_smem[g * 128 + tid];
            p.XD[(size_t)(512 + f0 + tid) * p.Bpad + b] = s;
        }
        __syncthreads();
    }
    const int CW = 256 / p.nsplit;
    if (tid < CW) {
        const int f = part * CW + tid;
        const float* vr = p.cval + (size_t)b * p.minT * 256 + f;
        float a = 0.f;
        for (int m = 0; m < p.minT; ++m) a = fmaf(sm.csc[m], __ldg(vr + (size_t)m * 256), a);
        p.XD[(size_t)f * p.Bpad + b] = a;
    }
    __syncthreads();
}

template <int R>
__device__ __forceinline__ void dec_run_pass(const DecodeParams& p, const DecPass& ps, const DecSmem& sm, StageSync& sync,
                                             int step, int parity_new) {
    const int tid = threadIdx.x;
    const float* xe = dec_src(p, ps.src_e, parity_new);
    const float* xl = dec_src(p, ps.src_l, parity_new);
    const float* W = sm.wsm + ps.w_off;
    const size_t plane = (size_t)1024 * p.Bpad;
    float* Snew = p.S + (size_t)parity_new * plane;
    for (int b0 = 0; b0 < p.Bpad; b0 += MV_CLIPS) {
        float acc[R][2];
        mv_zero<R>(acc);
        const bool narrow = p.B <= 2;                        // CTA-uniform: single-clip inference uses the k-split lane mapping
        if (narrow) {
            if (ps.Ke > 0) mv_accumulate_narrow<R>(W, ps.ldw, ps.wcol_e, xe, ps.Ke, p.Bpad, acc);
            sync.wait();
            mv_accumulate_narrow<R>(W, ps.ldw, ps.wcol_l, xl, ps.Kl, p.Bpad, acc);
        } else {
            if (ps.Ke > 0) mv_accumulate<R>(W, ps.ldw, ps.wcol_e, xe, ps.Ke, p.Bpad, b0, acc);
            sync.wait();
            mv_accumulate<R>(W, ps.ldw, ps.wcol_l, xl, ps.Kl, p.Bpad, b0, acc);
        }
        const int r = tid >> 5, bb = tid & 31, b = b0 + bb;
        const bool live = (tid < R * MV_CLIPS) && (b < p.B);
        const bool gate_pass = (ps.op[0] == OP_GATE0 || ps.op[0] == OP_GATE1);      // CTA-uniform
        float c_prev = 0.f;                                  // own cell state: fetched before the reduction to hide its latency
        if (gate_pass && live && (r & 3) == 0 && ps.idx[r] >= 0)
            c_prev = p.Cst[(size_t)((ps.op[0] == OP_GATE1 ? 512 : 0) + ps.idx[r]) * p.Bpad + b];
        float v = narrow ? mv_reduce_narrow<R>(acc, sm.red) : mv_reduce<R, (R == 16 ? DEC_RED_ROWS : R)>(acc, sm.red);
        const int op = live ? ps.op[r] : OP_NONE;
        const int idx = live ? ps.idx[r] : 0;
        if (live) v += ps.bias[r];
        switch (op) {
            case OP_FC:
                if (step >= 0) p.outputs[((size_t)b * p.steps + step) * 80 + idx] = v;
                break;
            case OP_P1: {
                float p1 = (step >= 0) ? sinf(v) * ps.aux[r] : ps.aux2[r];
                if (p.tf_mask && step + 1 < p.steps && p.tf_mask[step + 1])      // next step is teacher-forced
                    p1 = p.p1_teacher[((size_t)(step + 1) * 256 + idx) * p.Bpad + b];
                p.P1[(size_t)idx * p.Bpad + b] = p1;
            } break;
            case OP_STOP:
                if (step >= 0) {
                    const float logit = v + p.stop_const[b];
                    if (p.stop_out) p.stop_out[(size_t)b * p.steps + step] = logit;
                    if (logit > 0.f && p.lengths[b] == (long long)p.steps) p.lengths[b] = step + 1;
                }
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
                p.XD[(size_t)(256 + idx) * p.Bpad + b] = sinf(v) * ps.aux[r];
                break;
            default: break;
        }
        // LSTM passes: rows are (unit, gate) = (r>>2, r&3); all 16 rows of the pass are gate rows.
        if (gate_pass && tid < R * MV_CLIPS) sm.gsm[r * MV_CLIPS + bb] = v;
        __syncthreads();
        if (gate_pass) {
            if (live && (r & 3) == 0 && idx >= 0) {
                const int layer = (op == OP_GATE1) ? 1 : 0;
                const float gi = sm.gsm[(r + 0) * MV_CLIPS + bb], gf = sm.gsm[(r + 1) * MV_CLIPS + bb];
                const float gg = sm.gsm[(r + 2) * MV_CLIPS + bb], go = sm.gsm[(r + 3) * MV_CLIPS + bb];
                const size_t si = (size_t)(layer * 512 + idx) * p.Bpad + b;
                const float c = sigmoidf_acc(gf) * c_prev + sigmoidf_acc(gi) * tanhf(gg);
                const float h = sigmoidf_acc(go) * tanhf(c);
                p.Cst[si] = c;
                Snew[si] = h;
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ void dec_dispatch(const DecodeParams& p, const DecPass& ps, const DecSmem& sm, StageSync& sync,
                                             int step, int parity_new) {
    if (ps.R == 16) dec_run_pass<16>(p, ps, sm, sync, step, parity_new);
    else if (ps.R == 8) dec_run_pass<8>(p, ps, sm, sync, step, parity_new);
    else dec_run_pass<4>(p, ps, sm, sync, step, parity_new);
}

__global__ void __launch_bounds__(MV_THREADS, 1) decode_persistent_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) float smem[];
    DecSmem sm;
    sm.wsm = smem;                                           // weight image
    sm.red = sm.wsm + p.wimg_floats;                         // [16 warps][8][32]
    sm.gsm = sm.red + MV_WARPS * DEC_RED_ROWS * MV_CLIPS;    // [16][32]
    sm.qs = sm.gsm + 16 * MV_CLIPS;                          // [512]
    sm.sc = sm.qs + 512;                                     // [320]
    sm.cqs = sm.sc + 320;                                    // [256]
    sm.csc = sm.cqs + 256;                                   // [32]
    __shared__ DecPass passes[DEC_MAX_PASSES];
    __shared__ int np;
    __shared__ float tacc[DEC_TIMING_SLOTS];

    const int tid = threadIdx.x;
    if (tid == 0) np = p.npasses[blockIdx.x];
    if (tid < DEC_TIMING_SLOTS) tacc[tid] = 0.f;
    {
        const int* src = reinterpret_cast<const int*>(p.passes + (size_t)blockIdx.x * DEC_MAX_PASSES);
        int* dst = reinterpret_cast<int*>(passes);
        for (int i = tid; i < (int)(sizeof(DecPass) * DEC_MAX_PASSES / 4); i += MV_THREADS) dst[i] = src[i];
        const float4* wsrc = reinterpret_cast<const float4*>(p.wimg + (size_t)blockIdx.x * p.wimg_floats);
        float4* wdst = reinterpret_cast<float4*>(sm.wsm);
        for (int i = tid; i < p.wimg_floats / 4; i += MV_THREADS) wdst[i] = __ldg(wsrc + i);
    }
    __syncthreads();

    StageSync sync;
    sync.counter = p.barrier; sync.target = 0; sync.waited = true;     // nothing to wait for before the prologue
    sync.tacc = tacc; sync.tmark = 0; sync.slot0 = 0; sync.timing = false;
    const unsigned n = gridDim.x;
    const int njobs = p.B * p.nsplit;

    // prologue A(-1): Q, content query and prenet(BOS) from the initial state in S[0]
    for (int j = 0; j < np; ++j)
        if (passes[j].stage == ST_A) dec_dispatch(p, passes[j], sm, sync, -1, 0);
    grid_arrive(p.barrier);
    sync.target += n;
    sync.timing = (p.timing != nullptr);
    if (tid == 0) sync.tmark = clock64();

    for (int step = 0; step < p.steps; ++step) {
        const int parity_new = (step + 1) & 1;
#pragma unroll 1
        for (int st = ST_B; st <= ST_E + 1; ++st) {
            const int stage = (st == ST_E + 1) ? ST_A : st;
            sync.waited = false;
            sync.slot0 = 3 * (st - ST_B);
            for (int j = 0; j < np; ++j)
                if (passes[j].stage == stage) dec_dispatch(p, passes[j], sm, sync, step, parity_new);
            if (stage == ST_B) {
                for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
                    sync.wait();
                    dec_attend(p, sm, job / p.nsplit, job % p.nsplit, step);
                }
            }
            sync.wait();                 // barriers must complete in order even for CTAs idle in this stage
            sync.lap(sync.slot0 + 2);
            grid_arrive(p.barrier);
            sync.target += n;
        }
    }
    __syncthreads();
    if (p.timing && tid < DEC_TIMING_SLOTS) p.timing[blockIdx.x * DEC_TIMING_SLOTS + tid] = tacc[tid];
}

}  // namespace l2s
