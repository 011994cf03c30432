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
//  * NO grid barrier: every exchanged activation is a {value, turn tag} word that consumers poll (the LL protocol of
//    collective libraries): release fence + barrier counter + acquire poll + data load collapse into one L2 round trip;
//    write-after-read safety follows from the ring dependency B -> D -> E -> A -> B of a clip group.
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
constexpr int D3_TIMING_SLOTS = 6;    // early MMAs, wait for the late activations, late MMAs, reduce + epilogue / attention, (unused), idle turns

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

// ---- the exchange: flag-carrying words ---------------------------------------------------------------------------------
// Every activation that crosses CTAs carries its own "ready" flag: the two least significant mantissa bits of the fp32 word
// hold a 2-bit GENERATION tag (how many times this slot has been written in this launch, + 1, mod 4; the exchange buffers are
// zero-filled before every launch, so an unwritten slot reads tag 0).  A producer stores value-with-tag in ONE 4-byte store;
// a consumer loads the word and knows from the tag whether it holds the value it is waiting for — no release fence, no
// barrier counter, no second round trip, and no extra bytes: data and flag travel together (the idea of the LL protocol
// of collective libraries, with the flag folded into the payload).  The tagged word IS the activation everywhere (producer
// and every consumer see the same bits), so results stay deterministic and batch-invariant; the cost is 2 mantissa bits
// (relative 2.4e-7) on recurrent state that is already exchanged at 3xTF32 accuracy.  A 2-bit tag suffices because a slot is
// never rewritten before every reader of the previous value has finished (ring dependency B -> D -> E -> A -> B of a clip
// group), so a reader can only ever observe the previous generation or the one it waits for.
__device__ __forceinline__ void lx_store(float* p, float v, uint32_t tag) {
    const uint32_t w = (__float_as_uint(v) & ~3u) | tag;
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(w) : "memory");
}
__device__ __forceinline__ uint32_t lx_load(const float* p) {
    uint32_t w;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(w) : "l"(p) : "memory");
    return w;
}
// Generation tag of the value a stage produced at step `sp` (-1: prologue / initial state).  The recurrent state S has two
// parity planes, each written every second step; every other buffer is written once per step.
__device__ __forceinline__ uint32_t d3_tag(bool is_state, int sp) {
    const int gen = is_state ? ((sp + 1) >> 1) : (sp + 1);
    return (uint32_t)(gen + 1) & 3u;
}

// Spin bookkeeping: a wait that lasts absurdly long (a bug, never a legal schedule: a turn takes microseconds) raises the
// launch-wide abort word, and every thread that sees it stops waiting for the rest of the launch, so the kernel always
// terminates (results are then garbage and `abort` tells the host).
struct LLWait {
    unsigned* abort_word;
    bool dead;
    __device__ __forceinline__ bool give_up(unsigned& spins) {
        if ((++spins & 1023u) == 0u) {
            if (spins > (1u << 19)) atomicExch(abort_word, 1u);
            if (*reinterpret_cast<volatile unsigned*>(abort_word) != 0u) dead = true;
        }
        return dead;
    }
};

struct Decode3Params {
    DecodeParams d;               // sizes, attention memories, outputs
    // exchange buffers (tagged fp32 words, zero-filled before the launch)
    float* S;                     // [2 parity][groups][1024][8]   h0 rows 0..511, h1 rows 512..1023     (group-major, see matvec.cuh)
    float* Cst;                   // [groups][1024][8]             c0, c1
    float* P1;                    // [groups][256][8]              prenet layer 1
    float* XD;                    // [groups][1024][8]             content value (0..255), prenet-2 (256..511), attention context (512..1023)
    float* Q;                     // [clips][512]                  attention queries (clip-major)
    float* CQ;                    // [clips][256]                  content queries
    unsigned* abort_word;
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

// Which stage wrote a source; the tag a reader of stage `role` at step `step` waits for.  Stages run B, D, E, A inside a step:
// a producer that comes earlier in the step wrote at this step, a later (or the same) stage wrote at the previous one.
__device__ __forceinline__ int d3_producer(int src) {
    switch (src) {
        case SRC_H0NEW: case SRC_H0OLD: case SRC_C0: return ROLE_D;
        case SRC_H1NEW: case SRC_H1OLD: case SRC_C1: return ROLE_E;
        case SRC_P1: return ROLE_A;
        default: return ROLE_B;       // SRC_XD
    }
}
__device__ __forceinline__ uint32_t d3_src_tag(int src, int role, int step) {
    const bool is_state = src == SRC_H0NEW || src == SRC_H0OLD || src == SRC_H1NEW || src == SRC_H1OLD;
    return d3_tag(is_state, d3_producer(src) < role ? step : step - 1);
}

__device__ __forceinline__ const float* d3_src(const Decode3Params& q, int src, int parity_new, int g) {
    const size_t plane = (size_t)1024 * q.d.Bpad;
    const float* Snew = q.S + (size_t)parity_new * plane + (size_t)g * 1024 * D3_CG;
    const float* Sold = q.S + (size_t)(parity_new ^ 1) * plane + (size_t)g * 1024 * D3_CG;
    switch (src) {
        case SRC_H0NEW: return Snew;
        case SRC_H1NEW: return Snew + 512 * D3_CG;
        case SRC_C0: return q.Cst + (size_t)g * 1024 * D3_CG;
        case SRC_C1: return q.Cst + (size_t)g * 1024 * D3_CG + 512 * D3_CG;
        case SRC_P1: return q.P1 + (size_t)g * 256 * D3_CG;
        case SRC_XD: return q.XD + (size_t)g * 1024 * D3_CG;
        case SRC_H0OLD: return Sold;
        case SRC_H1OLD: return Sold + 512 * D3_CG;
        default: return nullptr;
    }
}

// Requests the chunks of one segment that this warp owns (see mv8_load) as tagged words.
template <int DEPTH>
__device__ __forceinline__ void d3_request(const float* __restrict__ X, int K, uint32_t (&w)[DEPTH][4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int npw = K / (MV_KC * MV_WARPS);
    const float* xp = X + (size_t)(warp * MV_KC + t) * 8 + g;
    constexpr int xstride = MV_WARPS * MV_KC * 8;
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
        if (d < npw) {
#pragma unroll
            for (int i = 0; i < 4; ++i) w[d][i] = lx_load(xp + d * xstride + i * 32);
        }
}
// Waits until every requested word carries `tag` (re-requesting the ones that do not yet) and hands the values out.
template <int DEPTH>
__device__ __forceinline__ void d3_collect(const float* __restrict__ X, int K, uint32_t tag, uint32_t (&w)[DEPTH][4], float (&x)[DEPTH][4], LLWait& lw) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int npw = K / (MV_KC * MV_WARPS);
    const float* xp = X + (size_t)(warp * MV_KC + t) * 8 + g;
    constexpr int xstride = MV_WARPS * MV_KC * 8;
    unsigned spins = 0;
    bool again = !lw.dead;
    while (again) {
        again = false;
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
            if (d < npw) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if ((w[d][i] & 3u) != tag) { w[d][i] = lx_load(xp + d * xstride + i * 32); again = true; }
            }
        if (again && lw.give_up(spins)) break;
    }
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
#pragma unroll
        for (int i = 0; i < 4; ++i) x[d][i] = __uint_as_float(w[d][i]);
}

// Per-row epilogue of reduction round `round` (rows 48*round .. 48*round+47) for clip group g; threads tid < 384.
// Exchanged outputs are written for all 8 clip slots of the group (padding clips included: consumers wait for every
// word); only the caller-visible outputs are masked by b < B.
__device__ __forceinline__ void d3_epilogue(const Decode3Params& q, const Dec3Pass& ps, const DecSmem& sm, float v, float c_prev,
                                            int round, int g, int step, int parity_new) {
    const DecodeParams& p = q.d;
    const uint32_t tag = d3_tag(false, step), stag = d3_tag(true, step);      // this stage's outputs of this step
    const int tid = threadIdx.x;
    const int r = 16 * MV8_RTILES * round + (tid >> 3), bb = tid & 7, b = g * D3_CG + bb;
    const bool rowlive = (tid < 128 * MV8_RTILES) && (r < ps.R);
    const bool real = b < p.B;
    const bool gate_pass = (ps.op[0] == OP_GATE0 || ps.op[0] == OP_GATE1);      // CTA-uniform
    const int op = rowlive ? ps.op[r] : OP_NONE;
    const int idx = rowlive ? ps.idx[r] : 0;
    if (rowlive) v += ps.bias[r];
    switch (op) {
        case OP_FC:
            if (step >= 0 && real) p.outputs[((size_t)b * p.steps + step) * 80 + idx] = v;
            break;
        case OP_P1: {
            float p1 = (step >= 0) ? sinf(v) * ps.aux[r] : ps.aux2[r];
            if (p.tf_mask && step + 1 < p.steps && p.tf_mask[step + 1])      // next step is teacher-forced
                p1 = p.p1_teacher[((size_t)(step + 1) * 256 + idx) * p.Bpad + b];
            lx_store(q.P1 + ((size_t)g * 256 + idx) * D3_CG + bb, p1, tag);
        } break;
        case OP_STOP:
            if (step >= 0 && real) {
                const float logit = v + p.stop_const[b];
                if (p.stop_out) p.stop_out[(size_t)b * p.steps + step] = logit;
                if (logit > 0.f && p.lengths[b] == (long long)p.steps) p.lengths[b] = step + 1;
            }
            break;
        case OP_Q: {
            float qv = sinf(v) * ps.aux[r];
            if (step + 1 < p.steps) qv += __ldg(p.pos + (size_t)(step + 1) * 512 + idx);
            lx_store(q.Q + (size_t)b * 512 + idx, qv, tag);
        } break;
        case OP_CQ:
            lx_store(q.CQ + (size_t)b * 256 + idx, siluf_acc(v), tag);
            break;
        case OP_P2:
            lx_store(q.XD + ((size_t)g * 1024 + 256 + idx) * D3_CG + bb, sinf(v) * ps.aux[r], tag);
            break;
        default: break;
    }
    if (gate_pass) {                                          // rows are (unit, gate) = (r>>2, r&3); 12 units per round
        // every thread applies its own gate non-linearity (i, f, o: sigmoid; g: tanh), the gate-0 thread combines
        if (tid < 128 * MV8_RTILES) sm.gsm[tid] = ((r & 3) == 2) ? tanhf(v) : sigmoidf_acc(v);
        __syncthreads();
        if (rowlive && (r & 3) == 0 && idx >= 0) {
            const int layer = (op == OP_GATE1) ? 1 : 0;
            const float gi = sm.gsm[tid], gf = sm.gsm[tid + 8], gg = sm.gsm[tid + 16], go = sm.gsm[tid + 24];
            const size_t si = (size_t)g * 1024 * D3_CG + (size_t)(layer * 512 + idx) * D3_CG + bb;
            const float c = gf * c_prev + gi * gg;
            const float h = go * tanhf(c);
            lx_store(q.Cst + si, c, tag);
            lx_store(q.S + (size_t)parity_new * 1024 * p.Bpad + si, h, stag);
        }
    }
}

constexpr int D3_EDEPTH = 2;          // chunks per warp of an early segment (Ke <= 512)

// What this CTA's stage does in a given turn of the software pipeline.
struct D3Slot { int g, step, turn; bool active; };
__device__ __forceinline__ D3Slot d3_slot(int turn, int role, int steps, int B) {
    D3Slot s;
    s.turn = turn;
    s.g = (turn - role) & (D3_NG - 1);                       // clip group served by this stage in this turn
    const int u = turn - s.g;                                // stages completed by that group (u % 4 == role)
    s.step = u >> 2;
    s.active = (u >= 0) && (s.step < steps) && (s.g * D3_CG < B);        // empty clip groups (B <= 24) are skipped
    return s;
}

struct D3Timing {
    float* tacc; long long tmark; bool on;
    __device__ __forceinline__ void lap(int slot) {
        if (on && threadIdx.x == 0) { long long now = clock64(); tacc[slot] += (float)(now - tmark); tmark = now; }
    }
};

// The pass of this CTA for one (clip group, step): early segment (operands at least two turns old), late segment (the
// previous stage's output of the previous turn), reduction + epilogue.
// xe: the early segment's words.  They are requested one turn ahead (right after the late MMAs of the previous turn were
// issued): `xe_valid` says whether that happened; on return xe holds the request for `next` (if next.active).
template <int RT, bool EARLY>
__device__ __forceinline__ void d3_turn(const Decode3Params& q, const Dec3Pass& ps, const DecSmem& sm, D3Timing& tm, LLWait& lw, int role,
                                        int g, int step, int parity_new, uint32_t (&xe)[D3_EDEPTH][4], bool& xe_valid, const D3Slot& next) {
    constexpr int ROUNDS = (RT + MV8_RTILES - 1) / MV8_RTILES;
    const DecodeParams& p = q.d;
    const int tid = threadIdx.x;
    const bool gate_pass = (ps.op[0] == OP_GATE0 || ps.op[0] == OP_GATE1);
    // the late request goes out first (it is the one the turn waits for; measured 3 % faster than issuing it after the
    // early tags are checked): the words travel while the early MMAs run
    const float* Xl = d3_src(q, ps.src_l, parity_new, g);
    uint32_t wl[MV8_DEPTH][4];
    d3_request<MV8_DEPTH>(Xl, ps.Kl, wl);
    // own cell state (written by this very thread four turns ago): consumed in the epilogue
    float c_prev[ROUNDS];
#pragma unroll
    for (int round = 0; round < ROUNDS; ++round) {
        const int r = 16 * MV8_RTILES * round + (tid >> 3);
        c_prev[round] = 0.f;
        if (gate_pass && tid < 128 * MV8_RTILES && (r & 3) == 0 && r < ps.R)
            c_prev[round] = __uint_as_float(lx_load(q.Cst + (size_t)g * 1024 * D3_CG + (size_t)((ps.op[0] == OP_GATE1 ? 512 : 0) + ps.idx[r]) * D3_CG + (tid & 7)));
    }
    float acc[RT][4];
    mv8_zero<RT>(acc);
    if (EARLY) {
        const float* Xe = d3_src(q, ps.src_e, parity_new, g);
        if (!xe_valid) d3_request<D3_EDEPTH>(Xe, ps.Ke, xe);
        float xv[D3_EDEPTH][4];
        d3_collect<D3_EDEPTH>(Xe, ps.Ke, d3_src_tag(ps.src_e, role, step), xe, xv, lw);
        mv8_mma<RT, D3_EDEPTH>(sm.wsm, ps.ldw, ps.wcol_e, ps.R, ps.Ke, xv, acc);
    }
    tm.lap(0);
    {
        float xl[MV8_DEPTH][4];
        d3_collect<MV8_DEPTH>(Xl, ps.Kl, d3_src_tag(ps.src_l, role, step), wl, xl, lw);
        tm.lap(1);
        mv8_mma<RT, MV8_DEPTH>(sm.wsm, ps.ldw, ps.wcol_l, ps.R, ps.Kl, xl, acc);
    }
    if (EARLY) {
        xe_valid = next.active;
        if (next.active) d3_request<D3_EDEPTH>(d3_src(q, ps.src_e, (next.step + 1) & 1, next.g), ps.Ke, xe);
    }
    tm.lap(2);
#pragma unroll
    for (int round = 0; round < ROUNDS; ++round) {
        if (16 * MV8_RTILES * round < ps.R) {
            const float v = mv8_reduce_round<RT>(acc, round, sm.red);
            d3_epilogue(q, ps, sm, v, c_prev[round], round, g, step, parity_new);
            __syncthreads();
        }
    }
    tm.lap(3);
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

// Dot-product attention over the T encoder positions and over the minT content slots for the clip in slot `b` of the batch
// (reference decoder.py:414-419 and Content.forward 262-271); `bsrc` = the clip whose keys/values are used (b itself, or
// the last real clip for a padding slot).  Both CTAs of a clip compute all scores; CTA `part` produces context features
// [256*part, 256*part+256) and content-value features [128*part, 128*part+128).
// Kb: [T][512]; Vb: this CTA's 256 features of row 0, row stride vstride; ck: content keys [minT][256]; cv: this CTA's
// 128 content-value features of slot 0, slot stride cvstride (all four in shared memory or all in global memory).
// Three CTA barriers: queries in shared memory | scores | (softmax recomputed by every warp) context + stores.
__device__ void d3_attend(const Decode3Params& q, const DecSmem& sm, const float* Kb, const float* Vb, int vstride,
                          const float* ck, const float* cv, int cvstride, int b, int part, int step, LLWait& lw, D3Timing& tm) {
    const DecodeParams& p = q.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = b / D3_CG, bb = b % D3_CG;
    const bool real = b < p.B;
    const uint32_t tag_in = d3_tag(false, step - 1);         // stage A wrote the queries at the end of the previous step
    const uint32_t tag = d3_tag(false, step);
    {
        const float* qp = q.Q + (size_t)b * 512 + tid;
        const float* cp = q.CQ + (size_t)b * 256 + (tid & 255);
        uint32_t wq = lx_load(qp), wc = lx_load(cp);
        unsigned spins = 0;
        while (!lw.dead && ((wq & 3u) != tag_in || (wc & 3u) != tag_in)) {
            if ((wq & 3u) != tag_in) wq = lx_load(qp);
            if ((wc & 3u) != tag_in) wc = lx_load(cp);
            if (lw.give_up(spins)) break;
        }
        sm.qs[tid] = __uint_as_float(wq) * p.temp;
        if (tid < 256) sm.cqs[tid] = __uint_as_float(wc) * p.ctemp;
    }
    tm.lap(1);
    __syncthreads();
    for (int t = warp; t < p.T; t += MV_WARPS) {
        const float4* kr = reinterpret_cast<const float4*>(Kb + (size_t)t * 512);
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 k = kr[lane + 32 * i];
            const float4 qv = *reinterpret_cast<const float4*>(sm.qs + 4 * (lane + 32 * i));
            a = fmaf(qv.x, k.x, a); a = fmaf(qv.y, k.y, a); a = fmaf(qv.z, k.z, a); a = fmaf(qv.w, k.w, a);
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
            const float4 qv = *reinterpret_cast<const float4*>(sm.cqs + 4 * (lane + 32 * i));
            a = fmaf(qv.x, k.x, a); a = fmaf(qv.y, k.y, a); a = fmaf(qv.z, k.z, a); a = fmaf(qv.w, k.w, a);
        }
        a = warp_sum(a);
        if (lane == 0) sm.csc[m] = a;
    }
    __syncthreads();
    float* xd = q.XD + (size_t)g * 1024 * D3_CG;
    if (vstride != 256) {
        // keys / values streamed from L2 (T > 31: the images do not fit in shared memory twice): the context sum is split over
        // the positions as well (thread = 4 features x every 8th position, float4 loads) so that each thread waits for ~T/8
        // dependent L2 loads instead of T, partials combined in a fixed order
        if (warp == 0) {
            float mx = -INFINITY;
            for (int t = lane; t < p.T; t += 32) mx = fmaxf(mx, sm.sc[t]);
            mx = warp_max(mx);
            float sum = 0.f;
            for (int t = lane; t < p.T; t += 32) sum += expf(sm.sc[t] - mx);
            sum = warp_sum(sum);
            for (int t = lane; t < p.T; t += 32) {
                const float a = expf(sm.sc[t] - mx) / sum;
                if (part == 0 && real) {
                    if (p.attn_logits) p.attn_logits[((size_t)b * p.steps + step) * p.T + t] = sm.sc[t];
                    if (p.attn) p.attn[((size_t)b * p.steps + step) * p.T + t] = a;
                }
                sm.sc[t] = a;
            }
        } else if (warp == 1) {
            const float v = lane < p.minT ? sm.csc[lane] : -INFINITY;
            const float mx = warp_max(v);
            const float e = lane < p.minT ? expf(v - mx) : 0.f;
            const float sum = warp_sum(e);
            if (lane < p.minT) sm.csc[lane] = e / sum;
        }
        __syncthreads();
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
        if (tid < 256) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) s += sm.red[j * 256 + tid];
            lx_store(xd + (size_t)(512 + part * 256 + tid) * D3_CG + bb, s, tag);
        } else if (tid < 256 + 128) {
            const int f = tid - 256;
            float a = 0.f;
            for (int m = 0; m < p.minT; ++m) a = fmaf(sm.csc[m], cv[(size_t)m * cvstride + f], a);
            lx_store(xd + (size_t)(part * 128 + f) * D3_CG + bb, a, tag);
        }
    } else if (warp < 8 || warp == 12) {
        // softmax over the T positions, recomputed by every warp that needs it (same instructions, same order: identical bits)
        float mx = -INFINITY;
        for (int t = lane; t < p.T; t += 32) mx = fmaxf(mx, sm.sc[t]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int t = lane; t < p.T; t += 32) sum += expf(sm.sc[t] - mx);
        sum = warp_sum(sum);
        if (warp == 12) {                                    // the caller-visible maps (one CTA of the pair writes them)
            if (part == 0 && real) {
                for (int t = lane; t < p.T; t += 32) {
                    if (p.attn_logits) p.attn_logits[((size_t)b * p.steps + step) * p.T + t] = sm.sc[t];
                    if (p.attn) p.attn[((size_t)b * p.steps + step) * p.T + t] = expf(sm.sc[t] - mx) / sum;
                }
            }
        } else {
            // context feature f = tid (256 per CTA): positions in ascending order, weights broadcast lane by lane
            const float* vcol = Vb + tid;
            float acc = 0.f;
            for (int t0 = 0; t0 < p.T; t0 += 32) {
                const int n = min(32, p.T - t0);
                const float mine = (lane < n) ? expf(sm.sc[t0 + lane] - mx) / sum : 0.f;
                for (int j = 0; j < n; ++j) acc = fmaf(__shfl_sync(0xffffffffu, mine, j), vcol[(size_t)(t0 + j) * vstride], acc);
            }
            lx_store(xd + (size_t)(512 + part * 256 + tid) * D3_CG + bb, acc, tag);
        }
    } else if (warp < 12) {
        // content slots (minT <= 32): softmax per warp, value feature f = tid - 256 (128 per CTA)
        const float v = lane < p.minT ? sm.csc[lane] : -INFINITY;
        const float mx = warp_max(v);
        const float e = lane < p.minT ? expf(v - mx) : 0.f;
        const float mine = e / warp_sum(e);
        const int f = tid - 256;
        float acc = 0.f;
        for (int m = 0; m < p.minT; ++m) acc = fmaf(__shfl_sync(0xffffffffu, mine, m), cv[(size_t)m * cvstride + f], acc);
        lx_store(xd + (size_t)(part * 128 + f) * D3_CG + bb, acc, tag);
    }
    __syncthreads();                                         // scratch (qs, sc) is rewritten by the next turn
}

template <int RT, bool EARLY>
__device__ __forceinline__ void d3_loop(const Decode3Params& q, const Dec3Pass& ps, const DecSmem& sm, D3Timing& tm,
                                        int role, int job, uint64_t* kvbar) {
    const DecodeParams& p = q.d;
    const bool has_pass = ps.R > 0;
    const int aclip = job / D3_NSPLIT, apart = job % D3_NSPLIT;         // attention CTAs: clip inside the group, half
    float* kvbuf = sm.csc + 32;                                         // [2][d3_kv_floats] when q.kv_smem
    const int kvfloats = d3_kv_floats(p.T, p.minT);
    uint32_t xe[D3_EDEPTH][4];
    bool xe_valid = false;
    LLWait lw; lw.abort_word = q.abort_word; lw.dead = false;
    D3Slot none; none.g = 0; none.step = 0; none.turn = 0; none.active = false;
    // prologue A(-1): Q, content query and prenet(BOS) of every clip group from the initial state in S[0]
    if (role == ROLE_A && has_pass)
        for (int g = 0; g * D3_CG < p.B; ++g) d3_turn<RT, EARLY>(q, ps, sm, tm, lw, role, g, -1, 0, xe, xe_valid, none);
    if (job >= 0 && q.kv_smem) d3_prefetch_kv(q, kvbuf, &kvbar[0], min(aclip, p.B - 1), apart);          // turn 0 serves group 0
    tm.on = (q.timing != nullptr);
    if (threadIdx.x == 0) tm.tmark = clock64();

    const int nturns = D3_NG * p.steps + D3_NG - 1;
    // attention CTAs: K/V image number n lives in buffer n & 1 and is that buffer's (n >> 1)-th use (mbarrier parity);
    // image n + 1 (the clip of this CTA's NEXT ACTIVE turn — empty clip groups are skipped) is requested when image n is
    // about to be consumed, i.e. after the previous reader of its buffer has finished.
    int kv_issued = (job >= 0 && q.kv_smem) ? 1 : 0, kv_consumed = 0;
#pragma unroll 1
    for (int turn = 0; turn < nturns; ++turn) {
        const D3Slot cur = d3_slot(turn, role, p.steps, p.B);
        if (!cur.active) continue;
        D3Slot next = none;                                  // this CTA's next active turn (at most D3_NG turns ahead)
        for (int d = 1; d <= D3_NG && !next.active && turn + d < nturns; ++d) next = d3_slot(turn + d, role, p.steps, p.B);
        const int g = cur.g, step = cur.step;
        const int parity_new = (step + 1) & 1;
        if (has_pass) d3_turn<RT, EARLY>(q, ps, sm, tm, lw, role, g, step, parity_new, xe, xe_valid, next);
        if (job >= 0) {
            const int b = g * D3_CG + aclip, bsrc = min(b, p.B - 1);
            if (q.kv_smem) {
                if (next.active) {
                    d3_prefetch_kv(q, kvbuf + (size_t)(kv_issued & 1) * kvfloats, &kvbar[kv_issued & 1], min(next.g * D3_CG + aclip, p.B - 1), apart);
                    ++kv_issued;
                }
                mbar_wait(&kvbar[kv_consumed & 1], (kv_consumed >> 1) & 1);     // this turn's clip has landed
                tm.lap(0);
                const float* kb = kvbuf + (size_t)(kv_consumed & 1) * kvfloats;
                const float* ckb = kb + (size_t)p.T * 768;
                d3_attend(q, sm, kb, kb + (size_t)p.T * 512, 256, ckb, ckb + (size_t)p.minT * 256, 128, b, apart, step, lw, tm);
            } else {
                tm.lap(0);
                d3_attend(q, sm, p.Kmem + (size_t)bsrc * p.T * 512, p.Vmem + (size_t)bsrc * p.T * 512 + apart * 256, 512,
                          p.ckey + (size_t)bsrc * p.minT * 256, p.cval + (size_t)bsrc * p.minT * 256 + apart * 128, 256, b, apart, step, lw, tm);
            }
            ++kv_consumed;
            tm.lap(3);
        }
    }
}

__global__ void __launch_bounds__(MV_THREADS, 1) decode3_kernel(const Decode3Params q) {
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
    sm.red = smem;                                           // [16 warps][3 tiles][128]
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

    D3Timing tm; tm.tacc = tacc; tm.tmark = 0; tm.on = false;
    const bool early = pass.Ke > 0;                        // (Ke <= 256 * D3_EDEPTH is checked by the packer)
    if (pass.RT <= 2) { if (early) d3_loop<2, true>(q, pass, sm, tm, role, job, kvbar); else d3_loop<2, false>(q, pass, sm, tm, role, job, kvbar); }
    else if (pass.RT == 3) { if (early) d3_loop<3, true>(q, pass, sm, tm, role, job, kvbar); else d3_loop<3, false>(q, pass, sm, tm, role, job, kvbar); }
    else if (pass.RT == 4) d3_loop<4, false>(q, pass, sm, tm, role, job, kvbar);
    else if (pass.RT == 5) d3_loop<5, false>(q, pass, sm, tm, role, job, kvbar);
    else if (pass.RT == 6) d3_loop<6, false>(q, pass, sm, tm, role, job, kvbar);
    else d3_loop<8, false>(q, pass, sm, tm, role, job, kvbar);
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

// Initial recurrent state: h (feature-major [1024][Bpad], the Bi-LSTM's final hidden states) -> S[parity 0], zero cell state
// -> Cst, group-major tagged words standing for "the output of LSTM-0 / LSTM-1 at step -1".
__global__ void d3_init_state_kernel(const float* __restrict__ h, float* __restrict__ S, float* __restrict__ Cst, int Bpad) {
    const size_t total = (size_t)1024 * Bpad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = i % Bpad; const int k = i / Bpad;
        const size_t o = ((size_t)(b / D3_CG) * 1024 + k) * D3_CG + (b % D3_CG);
        lx_store(S + o, h[i], d3_tag(true, -1));
        lx_store(Cst + o, 0.f, d3_tag(false, -1));
    }
}

}  // namespace l2s
