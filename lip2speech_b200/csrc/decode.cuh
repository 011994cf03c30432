// Shared definitions of the persistent autoregressive decode loop (reference decoder.py:403-435; one step = SURVEY.md §3.4):
// row operations, activation sources and the parameter block.  The kernel itself is decode3.cuh.
//
// A step has four dependent stages (linear∘linear pairs are merged on the host so that only FOUR remain):
//        A : fc_out -> mel frame, stop token ; prenet-1 (fused with fc_out: W_p1·W_fc) ;
//            Q = PSine(W_q [h0;h1]) + pos[i+1] ; content query = SiLU(W_cq [c0;c1])
//        B : prenet-2 ; per-clip dot-product attention over T (K,V) and over the content slots
//        D : LSTM-0 with attention_proj folded in: gates = [W_a|W_b|W_b·W_ap|W_hh] [cv;p2;ctx;h0] + b'
//        E : LSTM-1: gates = [W_ih|W_hh] [h0';h1] + b
// Each pass has an EARLY segment whose operands were produced two or more stages ago (h_prev for the LSTMs, h0'/c0' for
// the queries) and a LATE segment that needs the previous stage.  Which CTA owns which rows is decided on the host
// (pack.h: pack_decode_program3).  Stop-token bookkeeping (output_lengths) happens on the device; no host sync in the loop.
#pragma once
#include "matvec.cuh"

namespace l2s {

enum DecOp { OP_NONE = 0, OP_FC, OP_P1, OP_STOP, OP_Q, OP_CQ, OP_P2, OP_GATE0, OP_GATE1 };
enum DecSrc { SRC_NONE = 0, SRC_H0NEW, SRC_H1NEW, SRC_C0, SRC_C1, SRC_P1, SRC_XD, SRC_H0OLD, SRC_H1OLD };
enum DecStage { ST_A = 0, ST_B, ST_D, ST_E, ST_COUNT };

struct DecodeParams {
    // per-clip memories
    const float* Kmem; const float* Vmem;      // [B][T][512]
    const float* ckey; const float* cval;      // [B][minT][256]
    const float* stop_const;                   // [B]  W_stop[512:1024].enc_cell
    const float* pos;                          // [300][512]
    float temp, ctemp;
    float* outputs;               // [B][steps][80]
    long long* lengths;           // [B]
    float* attn;                  // [B][steps][T] or null
    int B, Bpad, T, minT, steps;  // B: clips of this launch (<= 32); Bpad: clip stride of the whole batch's planes
    // Decoder.forward flavour (decoder.py:353-375): teacher forcing and per-step logits; all null for inference
    const unsigned char* tf_mask; // [steps]: 1 = step i consumes the teacher frame (BOS for i=0, mels[:, i-1] otherwise)
    const float* p1_teacher;      // [steps][256][Bpad] prenet layer-1 output of the teacher frames
    float* stop_out;              // [B][steps] raw stop logits
    float* attn_logits;           // [B][steps][T] PRE-softmax attention logits
};

struct DecSmem {
    float* wsm; float* red; float* gsm; float* qs; float* sc; float* cqs; float* csc;
};

}  // namespace l2s
