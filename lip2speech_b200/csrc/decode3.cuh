// Stage-pipelined persistent decode loop (B <= 32; clip groups without clips are skipped): same arithmetic per clip as decode.cuh (reference
// decoder.py:403-435, one step = SURVEY.md §3.4), different mapping onto the chip.
//
// Why: in decode.cuh every SM owns rows of EVERY layer, so every SM reads every activation of every clip from L2
// in every step: 148 x 545 KB = 80 MB per step, and the L2 -> SM fabric tops out near 7.6 TB/s (measured,
// tools/mvt_bench.cu) -> >= 10.6 us per step before any arithmetic, plus a shared-memory-bound FMA pass.
//
// Here the four dependent stages of a step are given to four disjoint groups of SMs, and the batch is cut into four
// clip groups of 8 that travel through the stages one turn apart (a software pipeline over SM groups):
//
//        turn:        t        t+1      t+2      t+3
//   B SMs (attention, prenet-2)  g0       g3       g2       g1
//   D SMs (LSTM-0)               g1       g0       g3       g2
//   E SMs (LSTM-1)               g2       g1       g0       g3
//   A SMs (fc_out, queries)      g3       g2       g1       g0
//
//  * an SM holds the weights of ONE stage (the same ~190 KB of shared memory) and reads only that stage's inputs for
//    8 clips per turn: L2 -> SM traffic drops to ~26 MB per step;
//  * the 16-row x 8-clip tiles go to the tensor cores as error-compensated 3xTF32 (matvec.cuh: mv8_*), every weight is
//    read from shared memory once per turn;
//  * the attention SMs hold no weights: they keep the encoder keys/values of the clip they serve in shared memory,
//    double-buffered (the next turn's clip is fetched by the TMA engine, cp.async.bulk, while the current one is attended);
//  * one grid barrier per turn (four per step, as before), split-phase with the same early/late segments.
// Recurrent state is GROUP-MAJOR ([group][feature][8 clips], see matvec.cuh: mv8_accumulate); queries are clip-major.
// Outputs do not depend on which other clips share the batch (each clip's columns are independent in the MMA).
#pragma once
#include "decode.cuh"

namespace l2s {

constexpr int D3_MAXRT = 8;           // 16-row tiles per CTA (instantiated: 2, 3, 4, 5, 6, 8)
constexpr int D3_ROWS = 16 * D3_MAXRT;
constexpr int D3_CG = 8;              // clips per group
constexpr int D3_NG = 4;              // clip groups == pipeline stages
constexpr int D3_NSPLIT = 2;          // attention CTAs per clip (each: all scores, half of the context features)
constexpr int D3_TIMING_SLOTS = 6;    // early compute, barrier wait, late accumulate, reduce + epilogue / attention, idle turns, arrive

enum D3Role { ROLE_B = 0, ROLE_D = 1, ROLE_E = 2, ROLE_A = 3 };

// The ONE pass a CTA runs every turn: R rows (RT tiles) x [early segment | late segment].
struct Dec3Pass {
    int R, RT;
    int Ke, src_e, wcol_e;
    int Kl, src_l, wcol_l;
    int ldw, wfloats, pad0_, pad1_;
    int op[D3_ROWS];
    int idx[D3_ROWS];
    float bias[D3_ROWS];
    float aux[D3_ROWS];
    float aux2[D3_ROWS];
};

struct Decode3Params {
    DecodeParams d;               // sizes, attention memories, outputs; d.S / d.Cst / d.P1 / d.XD are group-major here,
                                  // d.Q [B][512] and d.CQ [B][256] clip-major (d.passes / d.npasses / d.wimg unused)
    const Dec3Pass* passes;       // [grid]  (R == 0: no pass)
    const int* role;              // [grid] D3Role
    const int* job;               // [grid] attention job inside a clip group (clip*D3_NSPLIT + part) or -1
    const float* wimg;            // [grid][wimg_floats]
    int wimg_floats;
    int kv_smem;                  // 1: two clips' K + V-half fit in shared memory (T <= 31)
    const float* Vsplit;          // [B][2][T][256]: the two feature halves of V, each contiguous (one bulk copy per image)
    const float* cvsplit;         // [B][2][minT][128]: likewise for the content values
    float* timing;                // optional [grid][D3_TIMING_SLOTS]
};

// group-major addressing: buffer of `rows` features, clip group g
__device__ __forceinline__ float* d3_group(float* buf, int rows, int g) { return buf + (size_t)g * rows * D3_CG; }

__device__ __forceinline__ const float* d3_src(const DecodeParams& p, int src, int parity_new, int g) {
    const size_t plane = (size_t)1024 * p.Bpad;
    const float* Snew = p.S + (size_t)parity_new * plane + (size_t)g * 1024 * D3_CG;
    const float* Sold = p.S + (size_t)(parity_new ^ 1) * plane + (size_t)g * 1024 * D3_CG;
    switch (src) {
        case SRC_H0NEW: return Snew;
        case SRC_H1NEW: return Snew + 512 * D3_CG;
        case SRC_C0: return p.Cst + (size_t)g * 1024 * D3_CG;
        case SRC_C1: return p.Cst + (size_t)g * 1024 * D3_CG + 512 * D3_CG;
        case SRC_P1: return p.P1 + (size_t)g * 256 * D3_CG;
        case SRC_XD: return p.XD + (size_t)g * 1024 * D3_CG;
        case SRC_H0OLD: return Sold;
        case SRC_H1OLD: return Sold + 512 * D3_CG;
        default: return nullptr;
    }
}

// Per-row epilogue of reduction round `round` (rows 48*round .. 48*round+47) for clip group g; threads tid < 384.
__device__ __forceinline__ void d3_epilogue(const DecodeParams& p, const Dec3Pass& ps, const DecSmem& sm, float v, float c_prev,
                                            int round, int g, int step, int parity_new) {
    const int tid = threadIdx.x;
    const int r = 16 * MV8_RTILES * round + (tid >> 3), bb = tid & 7, b = g * D3_CG + bb;
    const bool live = (tid < 128 * MV8_RTILES) && (r < ps.R) && (b < p.B);
    const bool gate_pass = (ps.op[0] == OP_GATE0 || ps.op[0] == OP_GATE1);      // CTA-uniform
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
            d3_group(p.P1, 256, g)[idx * D3_CG + bb] = p1;
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
            p.Q[(size_t)b * 512 + idx] = q;
        } break;
        case OP_CQ:
            p.CQ[(size_t)b * 256 + idx] = siluf_acc(v);
            break;
        case OP_P2:
            d3_group(p.XD, 1024, g)[(256 + idx) * D3_CG + bb] = sinf(v) * ps.aux[r];
            break;
        default: break;
    }
    if (gate_pass) {                                          // rows are (unit, gate) = (r>>2, r&3); 12 units per round
        // every thread applies its own gate non-linearity (i, f, o: sigmoid; g: tanh), the gate-0 thread combines
        if (tid < 128 * MV8_RTILES) sm.gsm[tid] = ((r & 3) == 2) ? tanhf(v) : sigmoidf_acc(v);
        __syncthreads();
        if (live && (r & 3) == 0 && idx >= 0) {
            const int layer = (op == OP_GATE1) ? 1 : 0;
            const float gi = sm.gsm[tid], gf = sm.gsm[tid + 8], gg = sm.gsm[tid + 16], go = sm.gsm[tid + 24];
            const size_t si = (size_t)g * 1024 * D3_CG + (size_t)(layer * 512 + idx) * D3_CG + bb;
            const float c = gf * c_prev + gi * gg;
            const float h = go * tanhf(c);
            p.Cst[si] = c;
            (p.S + (size_t)parity_new * 1024 * p.Bpad)[si] = h;
        }
    }
}

constexpr int D3_EDEPTH = 2;          // chunks per warp of an early segment (Ke <= 512)

// What this CTA's stage does in a given turn of the software pipeline.
struct D3Slot { int g, step; bool active; };
__device__ __forceinline__ D3Slot d3_slot(int turn, int role, int steps, int B) {
    D3Slot s;
    s.g = (turn - role) & (D3_NG - 1);                       // clip group served by this stage in this turn
    const int u = turn - s.g;                                // stages completed by that group (u % 4 == role)
    s.step = u >> 2;
    s.active = (u >= 0) && (s.step < steps) && (s.g * D3_CG < B);        // empty clip groups (B <= 24) are skipped
    return s;
}

// The pass of this CTA for one (clip group, step): early segment, barrier wait, late segment, reduction + epilogue.
// xe: the early segment's activations.  They are at least two turns old when they are used, so they are requested one
// turn ahead (right after the late MMAs of the previous turn were issued): `xe_valid` says whether that happened; on
// return xe holds the early activations of `next` (if next.active) and xe_valid is updated.
template <int RT, bool EARLY>
__device__ __forceinline__ void d3_turn(const DecodeParams& p, const Dec3Pass& ps, const DecSmem& sm, StageSync& sync,
                                        int g, int step, int parity_new, float (&xe)[D3_EDEPTH][4], bool& xe_valid, const D3Slot& next) {
    constexpr int ROUNDS = (RT + MV8_RTILES - 1) / MV8_RTILES;
    const int tid = threadIdx.x;
    const bool gate_pass = (ps.op[0] == OP_GATE0 || ps.op[0] == OP_GATE1);
    // own cell state (written by this very CTA four turns ago): requested first, consumed in the epilogue
    float c_prev[ROUNDS];
#pragma unroll
    for (int round = 0; round < ROUNDS; ++round) {
        const int r = 16 * MV8_RTILES * round + (tid >> 3), b = g * D3_CG + (tid & 7);
        c_prev[round] = 0.f;
        if (gate_pass && tid < 128 * MV8_RTILES && (r & 3) == 0 && r < ps.R && b < p.B)
            c_prev[round] = __ldcg(p.Cst + (size_t)g * 1024 * D3_CG + (size_t)((ps.op[0] == OP_GATE1 ? 512 : 0) + ps.idx[r]) * D3_CG + (tid & 7));
    }
    float acc[RT][4];
    mv8_zero<RT>(acc);
    if (EARLY) {
        if (!xe_valid) mv8_load<D3_EDEPTH>(d3_src(p, ps.src_e, parity_new, g), ps.Ke, xe);
        mv8_mma<RT, D3_EDEPTH>(sm.wsm, ps.ldw, ps.wcol_e, ps.R, ps.Ke, xe, acc);
    }
    sync.wait();
    {
        float xl[MV8_DEPTH][4];
        mv8_load<MV8_DEPTH>(d3_src(p, ps.src_l, parity_new, g), ps.Kl, xl);
        mv8_mma<RT, MV8_DEPTH>(sm.wsm, ps.ldw, ps.wcol_l, ps.R, ps.Kl, xl, acc);
    }
    if (EARLY) {
        xe_valid = next.active;
        if (next.active) mv8_load<D3_EDEPTH>(d3_src(p, ps.src_e, (next.step + 1) & 1, next.g), ps.Ke, xe);
    }
    sync.lap(2);
#pragma unroll
    for (int round = 0; round < ROUNDS; ++round) {
        if (16 * MV8_RTILES * round < ps.R) {
            const float v = mv8_reduce_round<RT>(acc, round, sm.red);
            d3_epilogue(p, ps, sm, v, c_prev[round], round, g, step, parity_new);
            __syncthreads();
        }
    }
}

// ---- attention CTAs ------------------------------------------------------------------------------------------------
// Shared-memory image of one clip: K [T][512], the CTA's half of V [T][256], content keys [minT][256] and the CTA's half
// of the content values [minT][128].
__device__ __forceinline__ int d3_kv_floats(int T, int minT) { return T * 768 + minT * 384; }

// Issued by one thread: expect_tx + four bulk copies (K, this CTA's half of V, content keys, its half of the content values;
// the halves come from the pre-split copies so that each is one contiguous piece — 35 row-wise bulk copies per image
// cost the issuing warp 1.2 us per turn).
__device__ __forceinline__ void d3_prefetch_kv(const Decode3Params& q, float* buf, uint64_t* bar, int b, int part) {
    if (threadIdx.x != 0) return;
    const DecodeParams& p = q.d;
    float* vb = buf + (size_t)p.T * 512;
    float* ckb = vb + (size_t)p.T * 256;
    float* cvb = ckb + (size_t)p.minT * 256;
    mbar_expect_tx(bar, (uint32_t)d3_kv_floats(p.T, p.minT) * 4u);
    bulk_load_1d(buf, p.Kmem + (size_t)b * p.T * 512, (uint32_t)p.T * 2048u, bar);
    bulk_load_1d(vb, q.Vsplit + ((size_t)b * 2 + part) * p.T * 256, (uint32_t)p.T * 1024u, bar);
    bulk_load_1d(ckb, p.ckey + (size_t)b * p.minT * 256, (uint32_t)p.minT * 1024u, bar);
    bulk_load_1d(cvb, q.cvsplit + ((size_t)b * 2 + part) * p.minT * 128, (uint32_t)p.minT * 512u, bar);
}

// Dot-product attention over the T encoder positions and over the minT content slots for clip b (reference
// decoder.py:414-419 and Content.forward 262-271).  Both CTAs of a clip compute all scores; CTA `part` produces context
// features [256*part, 256*part+256) and content-value features [128*part, 128*part+128).
// Kb: [T][512]; Vb: this CTA's 256 features of row 0, row stride vstride; ck: content keys [minT][256]; cv: this CTA's
// 128 content-value features of slot 0, slot stride cvstride (all four in shared memory or all in global memory).
__device__ void d3_attend(const DecodeParams& p, const DecSmem& sm, const float* Kb, const float* Vb, int vstride,
                          const float* ck, const float* cv, int cvstride, int b, int part, int step) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = b / D3_CG, bb = b % D3_CG;
    sm.qs[tid] = ldcg1(p.Q + (size_t)b * 512 + tid) * p.temp;
    if (tid < 256) sm.cqs[tid] = ldcg1(p.CQ + (size_t)b * 256 + tid) * p.ctemp;
    __syncthreads();
    for (int t = warp; t < p.T; t += MV_WARPS) {
        const float4* kr = reinterpret_cast<const float4*>(Kb + (size_t)t * 512);
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 k = kr[lane + 32 * i];
            const float4 q = *reinterpret_cast<const float4*>(sm.qs + 4 * (lane + 32 * i));
            a = fmaf(q.x, k.x, a); a = fmaf(q.y, k.y, a); a = fmaf(q.z, k.z, a); a = fmaf(q.w, k.w, a);
        }
        a = warp_sum(a);
        if (lane == 0) sm.sc[t] = a;
    }
    for (int m = warp; m < p.minT; m += MV_WARPS) {
        const float4* kr = reinterpret_cast<const float4*>(ck + (size_t)m * 256);
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float4 k = kr[lane + 32 * i];
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
    // context: thread (fq = tid&63: 4 features, tg = tid>>6: every 8th position) -> partial sums, combined in a fixed order
    {
        const int fq = tid & 63, tg = tid >> 6;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = tg; t < p.T; t += 8) {
            const float4 v = *reinterpret_cast<const float4*>(Vb + (size_t)t * vstride + 4 * fq);
            const float w = sm.sc[t];
            a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
        }
        *reinterpret_cast<float4*>(sm.red + tg * 256 + fq * 4) = a;
    }
    __syncthreads();
    float* xd = d3_group(p.XD, 1024, g);
    if (tid < 256) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += sm.red[j * 256 + tid];
        xd[(512 + part * 256 + tid) * D3_CG + bb] = s;
    } else if (tid < 256 + 128) {
        const int f = tid - 256;
        float a = 0.f;
        for (int m = 0; m < p.minT; ++m) a = fmaf(sm.csc[m], cv[(size_t)m * cvstride + f], a);
        xd[(part * 128 + f) * D3_CG + bb] = a;
    }
    __syncthreads();
}

template <int RT, bool EARLY>
__device__ __forceinline__ void d3_loop(const Decode3Params& q, const Dec3Pass& ps, const DecSmem& sm, StageSync& sync,
                                        int role, int job, uint64_t* kvbar) {
    const DecodeParams& p = q.d;
    const unsigned n = gridDim.x;
    const bool has_pass = ps.R > 0;
    const int aclip = job / D3_NSPLIT, apart = job % D3_NSPLIT;         // attention CTAs: clip inside the group, half
    float* kvbuf = sm.csc + 32;                                         // [2][d3_kv_floats] when q.kv_smem
    const int kvfloats = d3_kv_floats(p.T, p.minT);
    float xe[D3_EDEPTH][4];
    bool xe_valid = false;
    D3Slot none; none.g = 0; none.step = 0; none.active = false;
    // prologue A(-1): Q, content query and prenet(BOS) of every clip group from the initial state in S[0]
    if (role == ROLE_A && has_pass)
        for (int g = 0; g * D3_CG < p.B; ++g) d3_turn<RT, EARLY>(p, ps, sm, sync, g, -1, 0, xe, xe_valid, none);
    if (job >= 0 && q.kv_smem) d3_prefetch_kv(q, kvbuf, &kvbar[0], min(aclip, p.B - 1), apart);          // turn 0 serves group 0
    grid_arrive(p.barrier);
    sync.target += n;
    sync.timing = (q.timing != nullptr);
    if (threadIdx.x == 0) sync.tmark = clock64();

    const int nturns = D3_NG * p.steps + D3_NG - 1;
    // attention CTAs: K/V image number n lives in buffer n & 1 and is that buffer's (n >> 1)-th use (mbarrier parity);
    // image n + 1 (the clip of this CTA's NEXT ACTIVE turn — empty clip groups are skipped) is requested when image n is
    // about to be consumed, i.e. after the previous reader of its buffer has finished.
    int kv_issued = (job >= 0 && q.kv_smem) ? 1 : 0, kv_consumed = 0;
#pragma unroll 1
    for (int turn = 0; turn < nturns; ++turn) {
        const D3Slot cur = d3_slot(turn, role, p.steps, p.B);
        const D3Slot next = (turn + 1 < nturns) ? d3_slot(turn + 1, role, p.steps, p.B) : none;
        const int g = cur.g, step = cur.step;
        sync.waited = false;
        if (cur.active) {
            const int parity_new = (step + 1) & 1;
            if (has_pass) d3_turn<RT, EARLY>(p, ps, sm, sync, g, step, parity_new, xe, xe_valid, next);
            if (job >= 0) {
                const int b = g * D3_CG + aclip;
                if (q.kv_smem) {
                    D3Slot na = next;                        // this CTA's next active turn (at most D3_NG turns ahead)
                    for (int d = 2; d <= D3_NG && !na.active && turn + d < nturns; ++d) na = d3_slot(turn + d, role, p.steps, p.B);
                    if (na.active) {
                        d3_prefetch_kv(q, kvbuf + (size_t)(kv_issued & 1) * kvfloats, &kvbar[kv_issued & 1], min(na.g * D3_CG + aclip, p.B - 1), apart);
                        ++kv_issued;
                    }
                }
                sync.wait();
                if (q.kv_smem) mbar_wait(&kvbar[kv_consumed & 1], (kv_consumed >> 1) & 1);     // this turn's clip has landed
                sync.lap(2);
                if (b < p.B) {
                    if (q.kv_smem) {
                        const float* kb = kvbuf + (size_t)(kv_consumed & 1) * kvfloats;
                        const float* ckb = kb + (size_t)p.T * 768;
                        d3_attend(p, sm, kb, kb + (size_t)p.T * 512, 256, ckb, ckb + (size_t)p.minT * 256, 128, b, apart, step);
                    } else {
                        d3_attend(p, sm, p.Kmem + (size_t)b * p.T * 512, p.Vmem + (size_t)b * p.T * 512 + apart * 256, 512,
                                  p.ckey + (size_t)b * p.minT * 256, p.cval + (size_t)b * p.minT * 256 + apart * 128, 256, b, apart, step);
                    }
                }
                ++kv_consumed;
            }
            sync.wait();
            sync.lap(3);
        } else {
            sync.wait();
            sync.lap(4);
        }
        grid_arrive(p.barrier);
        sync.lap(5);
        sync.target += n;
    }
}

__global__ void __launch_bounds__(MV_THREADS, 1) decode3_kernel(const Decode3Params q) {
    const DecodeParams& p = q.d;
    extern __shared__ __align__(16) float smem[];
    __shared__ Dec3Pass pass;
    __shared__ float tacc[D3_TIMING_SLOTS];
    __shared__ __align__(8) uint64_t kvbar[2];              // attention CTAs: K/V image landed (one per buffer)
    const int tid = threadIdx.x;
    const int role = q.role[blockIdx.x], job = q.job[blockIdx.x];
    if (tid == 0) {
        mbar_init(&kvbar[0], 1); mbar_init(&kvbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    DecSmem sm;
    sm.red = smem;                                           // [16 warps][3 tiles][128]; attention partials [8][256]
    sm.gsm = sm.red + MV_WARPS * MV8_RTILES * 128;           // [48][8]
    sm.wsm = sm.gsm + 16 * MV8_RTILES * D3_CG;                            // weight image (absent on attention CTAs, whose scratch + K/V live here)
    sm.qs = sm.wsm;                                          // [512]
    sm.sc = sm.qs + 512;                                     // [320]
    sm.cqs = sm.sc + 320;                                    // [256]
    sm.csc = sm.cqs + 256;                                   // [32]; then [2][d3_kv_floats] K/V images
    if (tid < D3_TIMING_SLOTS) tacc[tid] = 0.f;
    {
        const int* src = reinterpret_cast<const int*>(q.passes + blockIdx.x);
        int* dst = reinterpret_cast<int*>(&pass);
        for (int i = tid; i < (int)(sizeof(Dec3Pass) / 4); i += MV_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    {
        const float4* wsrc = reinterpret_cast<const float4*>(q.wimg + (size_t)blockIdx.x * q.wimg_floats);
        float4* wdst = reinterpret_cast<float4*>(sm.wsm);
        for (int i = tid; i < pass.wfloats / 4; i += MV_THREADS) wdst[i] = __ldg(wsrc + i);
    }
    __syncthreads();

    StageSync sync;
    sync.counter = p.barrier; sync.target = 0; sync.waited = true;     // nothing to wait for before the prologue
    sync.tacc = tacc; sync.tmark = 0; sync.slot0 = 0; sync.timing = false;
    const bool early = pass.Ke > 0;                        // (Ke <= 256 * D3_EDEPTH is checked by the packer)
    if (pass.RT <= 2) { if (early) d3_loop<2, true>(q, pass, sm, sync, role, job, kvbar); else d3_loop<2, false>(q, pass, sm, sync, role, job, kvbar); }
    else if (pass.RT == 3) { if (early) d3_loop<3, true>(q, pass, sm, sync, role, job, kvbar); else d3_loop<3, false>(q, pass, sm, sync, role, job, kvbar); }
    else if (pass.RT == 4) d3_loop<4, false>(q, pass, sm, sync, role, job, kvbar);
    else if (pass.RT == 5) d3_loop<5, false>(q, pass, sm, sync, role, job, kvbar);
    else if (pass.RT == 6) d3_loop<6, false>(q, pass, sm, sync, role, job, kvbar);
    else d3_loop<8, false>(q, pass, sm, sync, role, job, kvbar);
    __syncthreads();
    if (q.timing && tid < D3_TIMING_SLOTS) q.timing[blockIdx.x * D3_TIMING_SLOTS + tid] = tacc[tid];
}

// src [B][rows][2*half] -> dst [B][2][rows][half] (feature halves made contiguous for the attention CTAs' bulk copies)
__global__ void split_halves_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int rows, int half) {
    const size_t total = (size_t)B * rows * 2 * half;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int f = i % (2 * half); size_t r = i / (2 * half);
        const int row = r % rows; const int b = r / rows;
        dst[(((size_t)b * 2 + f / half) * rows + row) * half + f % half] = src[i];
    }
}

// recurrent state of the row-partitioned layout ([feature][Bpad]) -> group-major ([group][feature][8])
__global__ void fm_to_group_major_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int Bpad) {
    const size_t total = (size_t)rows * Bpad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = i % Bpad; const int k = i / Bpad;
        dst[((size_t)(b / D3_CG) * rows + k) * D3_CG + (b % D3_CG)] = src[i];
    }
}

}  // namespace l2s
