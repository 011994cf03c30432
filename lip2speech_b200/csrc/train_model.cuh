// Train-mode forward + backward of the reference's Decoder (decoder.py:320-379) and VideoExtractor (video.py:76-87) on the tape
// engine (train_engine.cuh).  Every RNG site of the reference is an explicit input: KEEP masks (1 = kept) and
// the gumbel tensor, so both this path and the oracle (oracle/train_oracle.py) consume identical noise.
#pragma once
#include "train_engine.cuh"

namespace l2s {
namespace tr {

// out[(b*T + t)][c] = pos[t][c]
__global__ void tile_pos_kernel(int B, int T, int C, const float* __restrict__ pos, float* __restrict__ out) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)B * T * C) { const int c = i % C; const int t = (i / C) % T; out[i] = pos[(size_t)t * C + c]; }
}
// dst rows (strided view) = src ; backward: src.g += dst.g
// [B][C][M] keep-mask -> rows (b, m) x C
__global__ void zero_kernel(float* __restrict__ p, size_t n) {
    TR_PDL_WAIT(); TR_EW_LOOP(n) p[i] = 0.f; }
// stop[b*M + i] = per_step[i*B + b] + per_clip[b]   (stop-token logit = W_h . h1_i + (W_c . enc_cell + bias), decoder.py:373)
__global__ void stop_combine_kernel(int B, int M, const float* __restrict__ step, const float* __restrict__ clip, float* __restrict__ out) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)B * M) { const int b = i / M, m = i % M; out[i] = step[(size_t)m * B + b] + clip[b]; }
}
__global__ void stop_combine_bwd_kernel(int B, int M, const float* __restrict__ g, float* __restrict__ gstep, float* __restrict__ gclip) {
    TR_PDL_WAIT();
    TR_EW_LOOP((size_t)B * M) { const int m = i / B, b = i % B; gstep[i] += g[(size_t)b * M + m]; }
    TR_EW_LOOP((size_t)B) { float a = 0.f; for (int m = 0; m < M; ++m) a += g[(size_t)i * M + m]; gclip[i] += a; }
}

struct DecoderTrainIO {
    const float* visual; const float* spk; const float* mels;      // [B,T,1024], [B,256], [B,80,M]
    const unsigned char* tf_mask;                                   // host [M]
    const float* gumbel;                                            // [B*minT,501]
    const float* prenet_mask; const float* attn_mask; const float* lstm_mask;   // [M,B,256], [M,B,T], [M,B,512] keep masks
    const float* post_mask[5];                                      // [B,C,M] keep masks (C = 512 x4, 80)
    float* out_mel; float* out_post; float* out_stop; float* out_attn_logits; float* out_content_dis;
};

// Which way a forward ran, and therefore how its backward has to run.
enum GraphMode { RUN_EAGER = 0, RUN_CAPTURED = 1, RUN_REPLAY = 2 };

inline cudaStream_t capture_stream(cudaStream_t& cs) {            // capture needs a non-legacy stream; the caller's may be stream 0
    if (!cs) L2S_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    return cs;
}
inline int content_min_t(int T) { return T / 7; }                 // Content.agg: four Conv1d(k, stride=k), k = 1,3,5,7 -> shortest output (decoder.py:222-233)
inline void stage_in(float* dst, const float* src, size_t floats, cudaStream_t s) {      // null source: zeros
    if (src) L2S_CUDA(cudaMemcpyAsync(dst, src, floats * sizeof(float), cudaMemcpyDeviceToDevice, s));
    else L2S_CUDA(cudaMemsetAsync(dst, 0, floats * sizeof(float), s));
}

struct DecoderTrain {
    Engine e;
    Arena io;                                                       // staged copies of the caller's tensors (same addresses per shape)
    int B = 0, T = 0, M = 0;
    struct Handles { TT visual, spk, outputs, post, stops, cdis; int minT = 0; } h;
    struct Slot : GraphSlot { Handles h; };
    std::map<std::string, Slot> slots;
    Slot* slot = nullptr;
    GraphMode mode = RUN_EAGER;
    bool use_graphs = true;
    uint64_t bind_gen = 0;
    cudaStream_t cs = nullptr;
    bool live = false;
    bool want_logits = false;
    struct Staged {
        float *visual, *spk, *mels, *tf, *gumbel, *prenet, *attn, *lstm, *post[5];
        float *out_mel, *out_post, *out_stop, *out_logits, *out_cdis;
        float *g_mel, *g_post, *g_stop, *g_cdis;
    } st{};

    void drop_graphs() { for (auto& kv : slots) kv.second.release(); slots.clear(); slot = nullptr; }
    void release() { drop_graphs(); e.release(); io.free_all(); if (cs) cudaStreamDestroy(cs); cs = nullptr; }

    static TT assign(Engine& e, const TT& dst_view, const TT& src) {
        launch(ew_fwd_kernel<EW_COPY>, ew_blocks(src.numel()), 256, 0, e.s, src.rows, src.cols, src.v, src.rs, nullptr, 0, 0.f, 1, dst_view.v, dst_view.rs);
        e.ck("assign");
        Engine* pe = &e;
        e.tape.push_back([=]() {
            if (!src.g || !dst_view.g) return;
            launch(ew_bwd_kernel<EW_COPY>, ew_blocks(src.numel()), 256, 0, pe->s, src.rows, src.cols, nullptr, 0, nullptr, 0, 0.f, dst_view.g, dst_view.rs, src.g, src.rs);
            pe->ck("assign bwd");
        });
        return dst_view;
    }
    TT lin(const TT& x, const std::string& name, int N) {          // LinearNorm / nn.Linear / k=1 Conv1d named `name`(.weight/.bias)
        TT W = e.param(name + ".weight", N, x.cols), b = e.param(name + ".bias", 1, N);
        return e.linear(x, W, &b);
    }
    TT multihop(const TT& x, const std::string& p) {               // MultiHopConv, decoder.py:159-196
        static const int ks[4] = {1, 3, 7, 11};
        std::vector<TT> feats{x};
        for (int j = 0; j < 4; ++j) {
            const std::string n = p + "conv." + std::to_string(j);
            TT W = e.param(n + ".0.weight", 512, 512 * ks[j]), b = e.param(n + ".0.bias", 1, 512);
            TT y = e.conv1d(x, B, T, W, &b, ks[j], 1, ks[j] / 2);
            feats.push_back(e.silu(e.batchnorm(y, n + ".1")));
        }
        return lin(e.concat_cols(feats), p + "bottleneck", 512);
    }
    // one direction of the encoder Bi-LSTM (decoder.py:296,325); writes h_t into rnn_out[:, dir*512 : dir*512+512]
    void encoder_dir(const TT& x, const TT& site, const TT& rnn_out, int dir, TT& h_final, TT& c_final) {
        const std::string sfx = dir ? "_reverse" : "";
        const std::string p = "decoder.encoder_rnn.";
        TT Wih = e.param(p + "weight_ih_l0" + sfx, 2048, 1024), bih = e.param(p + "bias_ih_l0" + sfx, 1, 2048);
        TT Whh = e.param(p + "weight_hh_l0" + sfx, 2048, 512), bhh = e.param(p + "bias_hh_l0" + sfx, 1, 2048);
        TT xproj = e.linear(x, Wih, &bih);                          // [B*T, 2048], all time steps at once
        TT h = site, c = site;
        for (int k = 0; k < T; ++k) {
            const int t = dir ? T - 1 - k : k;
            TT g = xproj.rowslice(t, B, T);                          // the gates of step t, completed in place: += h W_hh^T + b_hh
            e.linear(h, Whh, &bhh, &g, true);
            TT hn, cn;
            TT hdst = rnn_out.rowslice(t, B, T).colslice(dir * 512, 512);
            e.lstm_cell(g, c, hn, cn, &hdst);                       // h_t lands in its slot of the Bi-LSTM output
            h = hn; c = cn;
        }
        h_final = h; c_final = c;
    }

    void forward(Context& ctx, std::map<std::string, Param>& params, uint64_t gen, const DecoderTrainIO& in, int B_, int T_, int M_, bool want_input_grads,
                 cudaStream_t s) {
        B = B_; T = T_; M = M_;
        if (B <= 0 || T < 7 || T > 300 || M <= 0 || M > 300) throw L2sError(1, "decoder_train_fwd: need 7<=T<=300, 1<=M<=300");
        live = false;
        e.setup();
        if (gen != bind_gen) { drop_graphs(); bind_gen = gen; }     // parameter / gradient memory moved: the graphs point at the old one
        want_logits = in.out_attn_logits != nullptr;
        // ---- stage the caller's tensors ---------------------------------------------------------------------------------------
        io.reset();
        const size_t nBM = (size_t)B * M;
        auto in_copy = [&](const float* src, size_t floats) { float* d = io.alloc(floats); stage_in(d, src, floats, s); return d; };
        st.visual = in_copy(in.visual, (size_t)B * T * 1024);
        st.spk = in_copy(in.spk, (size_t)B * 256);
        st.mels = in_copy(in.mels, nBM * 80);
        st.gumbel = in_copy(in.gumbel, (size_t)B * content_min_t(T) * 501);
        st.prenet = in_copy(in.prenet_mask, nBM * 256);
        st.attn = in_copy(in.attn_mask, nBM * T);
        st.lstm = in_copy(in.lstm_mask, nBM * 512);
        for (int i = 0; i < 5; ++i) st.post[i] = in_copy(in.post_mask[i], nBM * (i == 4 ? 80 : 512));
        {
            std::vector<float> tf(M);
            for (int i = 0; i < M; ++i) tf[i] = in.tf_mask[i] ? 1.f : 0.f;
            st.tf = io.alloc(M);
            L2S_CUDA(cudaMemcpyAsync(st.tf, tf.data(), (size_t)M * sizeof(float), cudaMemcpyHostToDevice, s));     // pageable source: staged before the call returns
        }
        st.out_mel = io.alloc(nBM * 80); st.out_post = io.alloc(nBM * 80); st.out_stop = io.alloc(nBM);
        st.out_logits = io.alloc(nBM * T); st.out_cdis = io.alloc((size_t)B * T * 501);
        st.g_mel = io.alloc(nBM * 80); st.g_post = io.alloc(nBM * 80); st.g_stop = io.alloc(nBM); st.g_cdis = io.alloc((size_t)B * T * 501);
        // ---- run: replay, capture + launch, or eager ------------------------------------------------------------------------------
        const std::string key = std::to_string(B) + "," + std::to_string(T) + "," + std::to_string(M) + "," + (want_input_grads ? "g" : "-") +
                                (want_logits ? "l" : "-") + (e.update_bn_running ? "r" : "-");
        if (!slots.count(key) && slots.size() >= 16) drop_graphs();       // ragged workloads (every batch its own M): bound the cache
        slot = &slots[key];
        if (use_graphs && slot->fwd && slot->bwd) {
            mode = RUN_REPLAY;
            h = slot->h;
            L2S_CUDA(cudaGraphLaunch(slot->fwd, s));
            ctx.launches += slot->fwd_launches;
        } else if (use_graphs && slot->seen >= 1) {
            mode = RUN_CAPTURED;
            slot->drop_graphs();
            const int64_t n0 = ctx.launches;
            e.capturing = true;
            try { slot->fwd = capture_graph(capture_stream(cs), [&]() { body(ctx, params, want_input_grads, cs); }); }
            catch (...) { e.capturing = false; throw; }
            e.capturing = false;
            slot->fwd_launches = ctx.launches - n0;
            slot->h = h;
            L2S_CUDA(cudaGraphLaunch(slot->fwd, s));
        } else {
            mode = RUN_EAGER;
            body(ctx, params, want_input_grads, s);
        }
        ++slot->seen;
        // ---- caller-visible outputs ------------------------------------------------------------------------------------------------
        auto out_copy = [&](float* dst, const float* src, size_t floats) { if (dst) L2S_CUDA(cudaMemcpyAsync(dst, src, floats * sizeof(float), cudaMemcpyDeviceToDevice, s)); };
        out_copy(in.out_mel, st.out_mel, nBM * 80);
        out_copy(in.out_post, st.out_post, nBM * 80);
        out_copy(in.out_stop, st.out_stop, nBM);
        out_copy(in.out_attn_logits, st.out_logits, nBM * T);
        out_copy(in.out_content_dis, st.out_cdis, (size_t)B * h.minT * 501);
        live = true;
    }

    // The launch sequence of one forward pass; reads and writes staged memory only (so it can be captured).
    void body(Context& ctx, std::map<std::string, Param>& params, bool want_input_grads, cudaStream_t s) {
        const Staged& io = st;
        TT &visual = h.visual, &spk = h.spk, &outputs = h.outputs, &post = h.post, &stops = h.stops, &cdis = h.cdis;
        int& minT = h.minT;
        e.begin(&ctx, s, &params);
        const std::string P = "decoder.";
        visual = e.wrap(io.visual, B * T, 1024, want_input_grads);
        spk = e.wrap(io.spk, B, 256, want_input_grads);
        const float* pos = e.param(P + "positional_encodings.pos_table", 300, 512).v;
        // ---- pre-loop (decoder.py:321-340) -------------------------------------------------------------------------------
        TT residual = lin(visual, P + "residual_bottleneck", 512);
        TT enc_site = e.psine(lin(spk, P + "encoder_site.0.linear_layer", 512), e.param(P + "encoder_site.1.w", 1, 512));
        TT att_site = e.psine(lin(spk, P + "attention_site.0.linear_layer", 512), e.param(P + "attention_site.1.w", 1, 512));
        TT rnn_out = e.make(B * T, 1024);
        TT hf, cf, hb, cb;
        encoder_dir(visual, enc_site, rnn_out, 0, hf, cf);
        encoder_dir(visual, enc_site, rnn_out, 1, hb, cb);
        TT enc_cell = lin(e.concat_cols({cf, cb}), P + "E_C.linear_layer", 512);
        TT enc = e.add(e.add_rows(lin(rnn_out, P + "encoder_proj.linear_layer", 512), att_site, T), residual);
        float* pos_tile = e.scratch((size_t)B * T * 512);
        launch(tile_pos_kernel, ew_blocks((size_t)B * T * 512), 256, 0, s, B, T, 512, pos, pos_tile);
        e.ck("pos tile");
        TT Kmem = e.add_const(e.psine(multihop(enc, P + "K.0."), e.param(P + "K.1.w", 1, 512)), pos_tile, 512);
        TT Vmem = e.add_const(e.psine(multihop(enc, P + "V.0."), e.param(P + "V.1.w", 1, 512)), pos_tile, 512);
        // ---- Content.encode (decoder.py:239-260) --------------------------------------------------------------------------
        TT ckey, cval;
        {
            static const int ks[4] = {1, 3, 5, 7};
            std::vector<TT> feats{enc};
            std::vector<int> Ls{T};
            minT = T;
            for (int j = 0; j < 4; ++j) {
                const std::string n = P + "content.agg." + std::to_string(j);
                TT W = e.param(n + ".0.weight", 512, 512 * ks[j]), b = e.param(n + ".0.bias", 1, 512);
                int Lo = 0;
                TT y = e.conv1d(enc, B, T, W, &b, ks[j], ks[j], 0, &Lo);
                feats.push_back(e.silu(e.batchnorm(y, n + ".1")));
                Ls.push_back(Lo);
                minT = std::min(minT, Lo);
            }
            std::vector<TT> pooled;
            for (size_t j = 0; j < feats.size(); ++j) pooled.push_back(e.adaptive_pool(feats[j], B, Ls[j], minT));
            TT w = lin(e.concat_cols(pooled), P + "content.bottleneck", 256);                  // rows (b, m) x 256
            ckey = e.silu(lin(e.silu(lin(w, P + "content.K.0", 256)), P + "content.K.2", 256));
            TT l = e.silu(lin(w, P + "content.location_fc.0", 256));
            l = e.silu(lin(l, P + "content.location_fc.2", 256));
            l = e.silu(lin(l, P + "content.location_fc.4", 501));
            TT z = e.softmax(e.scale(e.add_const(l, io.gumbel, 501), 1.0f / 0.1f));             // F.gumbel_softmax(w_y, 0.1), soft
            cval = e.matmul_nn(z, e.param(P + "content.word_embeddings", 501, 256));
            cdis = e.softmax(l);
        }
        // ---- the M-step loop (decoder.py:343-375) ---------------------------------------------------------------------------
        TT mel_rows = e.make(B * M, 80, false);                     // teacher frames as rows (b, i)
        launch(bcl_to_rows_tr_kernel, ew_blocks((size_t)B * 80 * M), 256, 0, s, B, 80, M, io.mels, mel_rows.v, 80);
        e.ck("mels -> rows");
        TT zeros80 = e.make(B, 80, false);
        launch(zero_kernel, ew_blocks((size_t)B * 80), 256, 0, s, zeros80.v, (size_t)B * 80);
        e.ck("zeros");
        TT bos = e.add_rows(zeros80, e.param(P + "BOS", 1, 80), B);                            // torch.tile(self.BOS, (N,1,1))
        outputs = e.make(B * M, 80);
        stops = e.make(B * M, 1);
        // State of all M steps in two buffers, rows (step, clip): [h0 | h1] and [c0 | c1] sit side by side, so the concatenations
        // the reference feeds to Q / content.Q (decoder.py:359,366) are plain views; hidden = encoder h_n, cell = 0 (347).
        TT HH = e.make(M * B, 1024), CC = e.make(M * B, 1024);
        TT hh_prev = e.concat_cols({hf, hb});
        TT cc_prev = e.make(B, 1024, false);
        launch(zero_kernel, ew_blocks((size_t)B * 1024), 256, 0, s, cc_prev.v, (size_t)B * 1024);
        e.ck("zeros");
        TT ys = bos;
        TT temp = e.param(P + "temperature", 1, 1), ctemp = e.param(P + "content.temperature", 1, 1);
        TT Wq_w = e.param(P + "Q.1.w", 1, 512), p1w = e.param(P + "prenet.1.w", 1, 256), p4w = e.param(P + "prenet.4.w", 1, 256);
        const std::string R = P + "decoder_rnn.";
        TT Wih0 = e.param(R + "weight_ih_l0", 2048, 512), bih0 = e.param(R + "bias_ih_l0", 1, 2048), Whh0 = e.param(R + "weight_hh_l0", 2048, 512), bhh0 = e.param(R + "bias_hh_l0", 1, 2048);
        TT Wih1 = e.param(R + "weight_ih_l1", 2048, 512), bih1 = e.param(R + "bias_ih_l1", 1, 2048), Whh1 = e.param(R + "weight_hh_l1", 2048, 512), bhh1 = e.param(R + "bias_hh_l1", 1, 2048);
        TT Wfc = e.param(P + "fc_out.linear_layer.weight", 80, 512), bfc = e.param(P + "fc_out.linear_layer.bias", 1, 80);
        for (int i = 0; i < M; ++i) {
            TT hh = HH.rowslice(i * B, B), cc = CC.rowslice(i * B, B);
            if (i > 0) ys = e.select(io.tf + i, mel_rows.rowslice(i - 1, B, M), ys);            // teacher_input[:, i] or the previous output (355-357)
            TT p1d = e.psine_chain(lin(ys, P + "prenet.0.linear_layer", 256), p1w, io.prenet + (size_t)i * B * 256, 256, 0.2f, nullptr, 0);
            TT p2 = e.psine_chain(lin(p1d, P + "prenet.3.linear_layer", 256), p4w, nullptr, 0, 0.f, nullptr, 0);
            TT q = e.psine_chain(lin(hh_prev, P + "Q.0.linear_layer", 512), Wq_w, nullptr, 0, 0.f, pos + (size_t)i * 512, 0);
            TT xy = e.make(B, 512);                                 // [content read-out | prenet + attention], the input of LSTM layer 0
            TT xy_c = xy.colslice(0, 256), xy_y = xy.colslice(256, 256);
            // location attention with the logit dropout (360-364); the post-dropout logits go straight to out[b][i][:]
            TT ctx = e.attn_step(q, temp, Kmem, Vmem, T, io.attn + (size_t)i * B * T, 1.0f / 0.9f, want_logits ? io.out_logits + (size_t)i * T : nullptr, M * T);
            TT o = lin(ctx, P + "attention_proj.linear_layer", 256);
            e.add(p2, o, &xy_y);
            TT cq = e.silu(lin(cc_prev, P + "content.Q.0", 256));
            e.attn_step(cq, ctemp, ckey, cval, minT, nullptr, 1.f, nullptr, 0, &xy_c);
            TT h0n, c0n, h1n, c1n;
            TT h0dst = hh.colslice(0, 512), c0dst = cc.colslice(0, 512), h1dst = hh.colslice(512, 512), c1dst = cc.colslice(512, 512);
            TT g0 = e.linear2(xy, Wih0, &bih0, hh_prev.colslice(0, 512), Whh0, &bhh0);
            e.lstm_cell(g0, cc_prev.colslice(0, 512), h0n, c0n, &h0dst, &c0dst);
            TT h0d = e.dropout(h0n, io.lstm + (size_t)i * B * 512, 512, 0.1f);                 // nn.LSTM(dropout=0.1): layer 1's input only
            TT g1 = e.linear2(h0d, Wih1, &bih1, hh_prev.colslice(512, 512), Whh1, &bhh1);
            e.lstm_cell(g1, cc_prev.colslice(512, 512), h1n, c1n, &h1dst, &c1dst);
            TT ydst = outputs.rowslice(i, B, M);
            ys = e.linear(h1n, Wfc, &bfc, &ydst);                   // straight into row (b, i) of the output
            hh_prev = hh; cc_prev = cc;
        }
        {   // stop-token logits of all steps at once: they feed nothing inside the loop (decoder.py:373)
            TT Wst = e.param(P + "stop_token_layer.linear_layer.weight", 1, 1024), bst = e.param(P + "stop_token_layer.linear_layer.bias", 1, 1);
            TT Wst_h = Wst.colslice(0, 512), Wst_c = Wst.colslice(512, 512);
            TT per_step = e.linear(HH.colslice(512, 512), Wst_h, nullptr);                     // [M*B, 1], rows (step, clip)
            TT per_clip = e.linear(enc_cell, Wst_c, &bst);                                     // [B, 1]
            launch(stop_combine_kernel, ew_blocks((size_t)B * M), 256, 0, s, B, M, per_step.v, per_clip.v, stops.v);
            e.ck("stop combine");
            Engine* pe = &e; const int Bc = B, Mc = M; TT st = stops;
            e.tape.push_back([=]() {
                launch(stop_combine_bwd_kernel, ew_blocks((size_t)Bc * Mc), 256, 0, pe->s, Bc, Mc, st.g, per_step.g, per_clip.g);
                pe->ck("stop combine bwd");
            });
        }
        // ---- postnet (decoder.py:143-156, 377-378) -------------------------------------------------------------------------
        {
            const std::string PP = P + "postnet.";
            TT x = outputs;
            for (int i = 0; i < 5; ++i) {
                const int cin = i == 0 ? 80 : 512, cout = i == 4 ? 80 : 512;
                const std::string n = PP + "convolutions." + std::to_string(i);
                TT W = e.param(n + ".0.conv.weight", cout, cin * 5), b = e.param(n + ".0.conv.bias", 1, cout);
                TT y = e.batchnorm(e.conv1d(x, B, M, W, &b, 5, 1, 2), n + ".1");
                if (i < 4) {
                    y = e.psine(y, e.param(PP + "sin_activation." + std::to_string(i) + ".w", 1, 512));
                    if (i != 0) y = e.add(y, x);
                }
                TT mrows = e.make(B * M, cout, false);              // keep mask [B,C,M] -> rows (b, m) x C
                launch(bcl_to_rows_tr_kernel, ew_blocks((size_t)B * cout * M), 256, 0, s, B, cout, M, io.post[i], mrows.v, cout);
                e.ck("post mask rows");
                x = e.dropout(y, mrows.v, cout, 0.5f);
            }
            post = e.add(x, outputs);
        }
        // ---- outputs in the caller's layouts (staged) ------------------------------------------------------------------------------
        launch(rows_to_bcl_tr_kernel, ew_blocks((size_t)B * 80 * M), 256, 0, s, B, 80, M, outputs.v, outputs.rs, io.out_mel); e.ck("out mel");
        launch(rows_to_bcl_tr_kernel, ew_blocks((size_t)B * 80 * M), 256, 0, s, B, 80, M, post.v, post.rs, io.out_post); e.ck("out post");
        L2S_CUDA(cudaMemcpyAsync(io.out_stop, stops.v, (size_t)B * M * sizeof(float), cudaMemcpyDeviceToDevice, s));
        L2S_CUDA(cudaMemcpyAsync(io.out_cdis, cdis.v, (size_t)B * minT * 501 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    void backward_body(cudaStream_t s) {
        e.s = s;
        e.begin_backward();
        launch(bcl_to_rows_tr_kernel, ew_blocks((size_t)B * 80 * M), 256, 0, s, B, 80, M, st.g_mel, h.outputs.g, 80); e.ck("g mel");
        launch(bcl_to_rows_tr_kernel, ew_blocks((size_t)B * 80 * M), 256, 0, s, B, 80, M, st.g_post, h.post.g, 80); e.ck("g post");
        L2S_CUDA(cudaMemcpyAsync(h.stops.g, st.g_stop, (size_t)B * M * sizeof(float), cudaMemcpyDeviceToDevice, s));
        L2S_CUDA(cudaMemcpyAsync(h.cdis.g, st.g_cdis, (size_t)B * h.minT * 501 * sizeof(float), cudaMemcpyDeviceToDevice, s));
        e.backward();
    }

    // Gradients of the scalar objective w.r.t. the four outputs the loss reads (any may be null = zero) -> parameter gradients
    // (accumulated into the bound gradient memory) and, optionally, the gradients of the inputs.
    void backward(Context& ctx, const float* g_mel, const float* g_post, const float* g_stop, const float* g_cdis, float* g_visual, float* g_spk, cudaStream_t s) {
        if (!live) throw L2sError(1, "decoder_train_bwd: no forward pass to differentiate (call l2s_decoder_train_fwd first)");
        if ((g_visual || g_spk) && !h.visual.g) throw L2sError(1, "decoder_train_bwd: input gradients were not requested in the forward call");
        const size_t nBM = (size_t)B * M;
        stage_in(st.g_mel, g_mel, nBM * 80, s);
        stage_in(st.g_post, g_post, nBM * 80, s);
        stage_in(st.g_stop, g_stop, nBM, s);
        stage_in(st.g_cdis, g_cdis, (size_t)B * h.minT * 501, s);
        if (mode == RUN_REPLAY) {
            L2S_CUDA(cudaGraphLaunch(slot->bwd, s));
            ctx.launches += slot->bwd_launches;
        } else if (mode == RUN_CAPTURED && slot->table_bytes) {
            slot->tables.reserve(slot->table_bytes);
            slot->tables.used = 0;
            e.tables = &slot->tables;
            const int64_t n0 = ctx.launches;
            e.capturing = true;
            try { slot->bwd = capture_graph(cs, [&]() { backward_body(cs); }); }
            catch (...) { e.capturing = false; throw; }
            e.capturing = false;
            slot->bwd_launches = ctx.launches - n0;
            L2S_CUDA(cudaGraphLaunch(slot->bwd, s));
        } else {                                                    // eager (also: a captured forward whose table size is not known yet)
            backward_body(s);
            slot->table_bytes = e.table_bytes;
        }
        if (g_visual) L2S_CUDA(cudaMemcpyAsync(g_visual, h.visual.g, (size_t)B * T * 1024 * sizeof(float), cudaMemcpyDeviceToDevice, s));
        if (g_spk) L2S_CUDA(cudaMemcpyAsync(g_spk, h.spk.g, (size_t)B * 256 * sizeof(float), cudaMemcpyDeviceToDevice, s));
        live = false;
    }
};

// ---- VideoExtractor.forward in train mode (video.py:76-87; shufflenetv2.py:42-104; model.py:26 dropout) ---------------
struct VideoTrain {
    Engine e;
    Arena io;
    int B = 0, T = 0, H = 0, W = 0;
    TT feat;
    struct Slot : GraphSlot { TT feat; };
    std::map<std::string, Slot> slots;
    Slot* slot = nullptr;
    GraphMode mode = RUN_EAGER;
    bool use_graphs = true;
    uint64_t bind_gen = 0;
    cudaStream_t cs = nullptr;
    bool live = false;
    float *st_video = nullptr, *st_mask = nullptr, *st_out = nullptr, *st_g = nullptr;

    void drop_graphs() { for (auto& kv : slots) kv.second.release(); slots.clear(); slot = nullptr; }
    void release() { drop_graphs(); e.release(); io.free_all(); if (cs) cudaStreamDestroy(cs); cs = nullptr; }

    TT pw(const TT& x, const std::string& name, int cout) {        // 1x1 Conv2d, bias=False
        return e.linear(x, e.param(name + ".weight", cout, x.cols), nullptr);
    }
    void forward(Context& ctx, std::map<std::string, Param>& params, uint64_t gen, const float* video, const float* drop_mask, int B_, int T_, int H_, int W_,
                 float* out_feat, cudaStream_t s) {
        B = B_; T = T_; H = H_; W = W_;
        if (B <= 0 || T <= 0 || (H & 3) || (W & 3)) throw L2sError(1, "video_train_fwd: bad shape");
        live = false;
        e.setup();
        e.exact_gemm = true;                                        // BatchNorm + ReLU stacks: see sgemm_kernel
        if (gen != bind_gen) { drop_graphs(); bind_gen = gen; }
        const size_t N = (size_t)B * T;
        io.reset();
        st_video = io.alloc(N * 3 * H * W); stage_in(st_video, video, N * 3 * H * W, s);
        st_mask = io.alloc(N * 768); if (drop_mask) stage_in(st_mask, drop_mask, N * 768, s);
        st_out = io.alloc(N * 768); st_g = io.alloc(N * 768);
        const bool masked = drop_mask != nullptr;
        const std::string key = std::to_string(B) + "," + std::to_string(T) + "," + std::to_string(H) + "," + std::to_string(W) + (masked ? "m" : "-") +
                                (e.update_bn_running ? "r" : "-");
        if (!slots.count(key) && slots.size() >= 16) drop_graphs();
        slot = &slots[key];
        if (use_graphs && slot->fwd && slot->bwd) {
            mode = RUN_REPLAY;
            feat = slot->feat;
            L2S_CUDA(cudaGraphLaunch(slot->fwd, s));
            ctx.launches += slot->fwd_launches;
        } else if (use_graphs && slot->seen >= 1) {
            mode = RUN_CAPTURED;
            slot->drop_graphs();
            const int64_t n0 = ctx.launches;
            e.capturing = true;
            try { slot->fwd = capture_graph(capture_stream(cs), [&]() { body(ctx, params, masked, cs); }); }
            catch (...) { e.capturing = false; throw; }
            e.capturing = false;
            slot->fwd_launches = ctx.launches - n0;
            slot->feat = feat;
            L2S_CUDA(cudaGraphLaunch(slot->fwd, s));
        } else {
            mode = RUN_EAGER;
            body(ctx, params, masked, s);
        }
        ++slot->seen;
        L2S_CUDA(cudaMemcpyAsync(out_feat, st_out, N * 768 * sizeof(float), cudaMemcpyDeviceToDevice, s));
        live = true;
    }
    void body(Context& ctx, std::map<std::string, Param>& params, bool masked, cudaStream_t s) {
        const float* video = st_video;
        const float* drop_mask = masked ? st_mask : nullptr;
        float* out_feat = st_out;
        e.begin(&ctx, s, &params);
        const std::string P = "encoder.";
        const int N = B * T;
        TT x = e.stem_conv(video, B, T, H, W, e.param(P + "frontend3D.0.weight", 24, 735));
        x = e.prelu(e.batchnorm(x, P + "frontend3D.1"), e.param(P + "frontend3D.2.weight", 1, 24));
        int h = (H - 1) / 2 + 1, w = (W - 1) / 2 + 1;
        x = e.maxpool3x3s2(x, N, h, w, &h, &w);
        for (int blk = 0;; ++blk) {
            const std::string q = P + "trunk.0." + std::to_string(blk) + ".";
            if (!params.count(q + "banch2.0.weight")) break;
            const bool down = params.count(q + "banch1.0.weight") != 0;
            if (down) {                                             // benchmodel 2: both branches see x, stride 2
                const int cin = x.cols, half = (int)(params.at(q + "banch2.0.weight").n / cin);
                int ho, wo;
                TT b1 = e.dwconv3x3(x, N, h, w, 2, e.param(q + "banch1.0.weight", cin, 9), &ho, &wo);
                b1 = e.relu(e.batchnorm(pw(e.batchnorm(b1, q + "banch1.1"), q + "banch1.2", half), q + "banch1.3"));
                TT b2 = e.relu(e.batchnorm(pw(x, q + "banch2.0", half), q + "banch2.1"));
                b2 = e.batchnorm(e.dwconv3x3(b2, N, h, w, 2, e.param(q + "banch2.3.weight", half, 9), &ho, &wo), q + "banch2.4");
                b2 = e.relu(e.batchnorm(pw(b2, q + "banch2.5", half), q + "banch2.6"));
                x = e.interleave2(b1, b2);
                h = ho; w = wo;
            } else {                                                // benchmodel 1: x1 passes through, x2 -> branch 2
                const int half = x.cols / 2;
                int ho, wo;
                TT x1 = x.colslice(0, half), x2 = x.colslice(half, half);
                TT b2 = e.relu(e.batchnorm(pw(x2, q + "banch2.0", half), q + "banch2.1"));
                b2 = e.batchnorm(e.dwconv3x3(b2, N, h, w, 1, e.param(q + "banch2.3.weight", half, 9), &ho, &wo), q + "banch2.4");
                b2 = e.relu(e.batchnorm(pw(b2, q + "banch2.5", half), q + "banch2.6"));
                x = e.interleave2(x1, b2);
            }
        }
        if (h != 3 || w != 3) throw L2sError(1, "video_train_fwd: trunk output must be 3x3 (H, W in {88, 96}) for AvgPool2d(3)");
        x = e.relu(e.batchnorm(pw(x, P + "trunk.1.0", 768), P + "trunk.1.1"));
        x = e.adaptive_pool(x, N, 9, 1);                            // AvgPool2d(3) on the 3x3 map
        x = e.l2normalize(x);                                       // F.normalize(p=2, dim=2), video.py:85
        if (drop_mask) x = e.dropout(x, drop_mask, 768, 0.1f);      // F.dropout(video_features, 0.1, training), model.py:26
        feat = x;
        L2S_CUDA(cudaMemcpyAsync(out_feat, x.v, (size_t)N * 768 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    void backward_body(cudaStream_t s) {
        e.s = s;
        e.begin_backward();
        L2S_CUDA(cudaMemcpyAsync(feat.g, st_g, (size_t)B * T * 768 * sizeof(float), cudaMemcpyDeviceToDevice, s));
        e.backward();
    }
    void backward(Context& ctx, const float* g_feat, cudaStream_t s) {
        if (!live) throw L2sError(1, "video_train_bwd: no forward pass to differentiate (call l2s_video_train_fwd first)");
        stage_in(st_g, g_feat, (size_t)B * T * 768, s);
        if (mode == RUN_REPLAY) {
            L2S_CUDA(cudaGraphLaunch(slot->bwd, s));
            ctx.launches += slot->bwd_launches;
        } else if (mode == RUN_CAPTURED) {                          // no per-step layers in the frontend: the backward uploads no pointer tables
            slot->tables.used = 0;
            e.tables = &slot->tables;
            const int64_t n0 = ctx.launches;
            e.capturing = true;
            try { slot->bwd = capture_graph(cs, [&]() { backward_body(cs); }); }
            catch (...) { e.capturing = false; throw; }
            e.capturing = false;
            slot->bwd_launches = ctx.launches - n0;
            L2S_CUDA(cudaGraphLaunch(slot->bwd, s));
        } else backward_body(s);
        live = false;
    }
};

}  // namespace tr
}  // namespace l2s
