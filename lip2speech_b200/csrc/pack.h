// Host-side weight packing: eval-BatchNorm folding, layout changes, per-SM row partition of the decode
// step.  Runs once per l2s_commit_weights on the CPU in plain C++ (38 M parameters, ~100 ms).
#pragma once
#include <algorithm>
#include <cmath>
#include <numeric>

#include "context.h"
#include "decode.cuh"
#include "decode3.cuh"
#include "lstm.cuh"

namespace l2s {

constexpr float BN_EPS = 1e-5f;

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// scale/shift of an eval BatchNorm: y = x*scale + shift  (reference: nn.BatchNorm*, eps=1e-5)
inline void bn_affine(const Context& c, const std::string& bn, std::vector<float>& scale, std::vector<float>& shift) {
    const auto& g = c.W(bn + ".weight").f; const auto& b = c.W(bn + ".bias").f;
    const auto& m = c.W(bn + ".running_mean").f; const auto& v = c.W(bn + ".running_var").f;
    scale.resize(g.size()); shift.resize(g.size());
    for (size_t i = 0; i < g.size(); ++i) {
        float s = g[i] / std::sqrt(v[i] + BN_EPS);
        scale[i] = s; shift[i] = b[i] - m[i] * s;
    }
}

// Conv1d weight [co][ci][k] (+bias) followed by BN -> [co][k*ci_n + ci] with BN folded; bias out.
inline void pack_conv1d(const Context& c, const std::string& conv, const std::string& bn /* "" = none */,
                        std::vector<float>& w, std::vector<float>& bias) {
    const HostTensor& W = c.W(conv + ".weight");
    const int co = (int)W.shape[0], ci = (int)W.shape[1], k = (int)W.shape[2];
    const auto& b = c.W(conv + ".bias").f;
    std::vector<float> scale(co, 1.f), shift(co, 0.f);
    if (!bn.empty()) bn_affine(c, bn, scale, shift);
    w.assign((size_t)co * k * ci, 0.f); bias.assign(co, 0.f);
    for (int o = 0; o < co; ++o) {
        for (int i = 0; i < ci; ++i)
            for (int t = 0; t < k; ++t) w[(size_t)o * k * ci + (size_t)t * ci + i] = W.f[((size_t)o * ci + i) * k + t] * scale[o];
        bias[o] = b[o] * scale[o] + shift[o];
    }
}

// Tensor-core (3xTF32) operand form of a [N][taps*Kc] weight: every tap padded to Kcp = roundup(Kc,32)
// (so a 32-wide K chunk never straddles taps), split into w_hi (low 13 mantissa bits cleared) and
// w_lo = w - w_hi (exact).  Uploaded as name.hi / name.lo; meta name.kcp.
// round-to-nearest-even fp32 -> bf16 (bit pattern)
inline uint16_t bf16_bits(float v) {
    uint32_t b; std::memcpy(&b, &v, 4);
    if ((b & 0x7f800000u) == 0x7f800000u) return (uint16_t)(b >> 16);      // inf / nan
    b += 0x7fffu + ((b >> 16) & 1u);
    return (uint16_t)(b >> 16);
}

inline void upload_tc(Context& c, const std::string& name, const std::vector<float>& w, int N, int taps, int Kc) {
    const int Kcp = round_up(Kc, 32);
    std::vector<float> hi((size_t)N * taps * Kcp, 0.f), lo((size_t)N * taps * Kcp, 0.f);
    for (int n = 0; n < N; ++n)
        for (int t = 0; t < taps; ++t)
            for (int k = 0; k < Kc; ++k) {
                const float v = w[((size_t)n * taps + t) * Kc + k];
                uint32_t bits; std::memcpy(&bits, &v, 4);
                bits &= 0xFFFFE000u;
                float h; std::memcpy(&h, &bits, 4);
                hi[((size_t)n * taps + t) * Kcp + k] = h;
                lo[((size_t)n * taps + t) * Kcp + k] = v - h;
            }
    c.upload(name + ".hi", hi); c.upload(name + ".lo", lo);
    c.meta[name + ".kcp"] = Kcp;
}

// ------------------------------------------------------------------------------------------------
// video frontend
// ------------------------------------------------------------------------------------------------
struct TrunkBlockMeta { int down, cin_phys, half, hp, in_half, in_hp; };

inline int phys_ch(int l, int half, int hp) { return (half > 0 && l >= half) ? l - half + hp : l; }

inline void pack_video(Context& c) {
    const std::string p = "encoder.";
    {   // stem: Conv3d weight [24][3*5*7*7] * bn scale, bias = shift, PReLU slopes
        const HostTensor& W = c.W(p + "frontend3D.0.weight");
        std::vector<float> scale, shift; bn_affine(c, p + "frontend3D.1", scale, shift);
        std::vector<float> w(W.f.size());
        const int K = 735;
        for (int o = 0; o < 24; ++o) for (int k = 0; k < K; ++k) w[(size_t)o * K + k] = W.f[(size_t)o * K + k] * scale[o];
        c.upload("v.stem.w", w); c.upload("v.stem.b", shift); c.upload("v.stem.prelu", c.W(p + "frontend3D.2.weight").f);
        // tensor-core form: stride-2 7x7 -> stride-1 4x4 over the space-to-depth input (12 = 2x2 phases x 3 channels);
        // tap = (kt, jh), K within a tap = jw*12 + (ph*2+pw)*3 + ci, original kh = 2*jh + ph - 1, kw = 2*jw + pw - 1.
        std::vector<float> ws((size_t)24 * 20 * 48, 0.f);
        for (int o = 0; o < 24; ++o)
            for (int ci = 0; ci < 3; ++ci)
                for (int kt = 0; kt < 5; ++kt)
                    for (int jh = 0; jh < 4; ++jh)
                        for (int ph = 0; ph < 2; ++ph)
                            for (int jw = 0; jw < 4; ++jw)
                                for (int pw = 0; pw < 2; ++pw) {
                                    const int kh = 2 * jh + ph - 1, kw = 2 * jw + pw - 1;
                                    if (kh < 0 || kw < 0) continue;
                                    const float v = w[(size_t)o * K + ((ci * 5 + kt) * 7 + kh) * 7 + kw];
                                    ws[((size_t)o * 20 + kt * 4 + jh) * 48 + jw * 12 + (ph * 2 + pw) * 3 + ci] = v;
                                }
        upload_tc(c, "v.stem.tc", ws, 24, 20, 48);
        // bf16 form (precision = bf16): [32 rows][20 taps][4 positions x 16 channels], 12 real channels per position
        std::vector<uint16_t> w16((size_t)32 * 20 * 64, 0);
        for (int o = 0; o < 24; ++o)
            for (int tap = 0; tap < 20; ++tap)
                for (int jw = 0; jw < 4; ++jw)
                    for (int ch = 0; ch < 12; ++ch)
                        w16[((size_t)o * 20 + tap) * 64 + jw * 16 + ch] = bf16_bits(ws[((size_t)o * 20 + tap) * 48 + jw * 12 + ch]);
        c.upload_raw("v.stem.tc16", w16.data(), w16.size());
    }
    int blk = 0;
    int in_half = 0, in_hp = 0, cin = 24, cin_phys = 24;       // stem output: identity layout
    std::vector<TrunkBlockMeta> metas;
    while (c.has(p + "trunk.0." + std::to_string(blk) + ".banch2.0.weight")) {
        const std::string q = p + "trunk.0." + std::to_string(blk) + ".";
        const std::string n = "v.b" + std::to_string(blk) + ".";
        const bool down = c.has(q + "banch1.0.weight");
        const int half = (int)c.W(q + "banch2.5.weight").shape[0];
        const int hp = round_up(half, 4);
        std::vector<float> scale, shift;
        auto pack_pw = [&](const std::string& conv, const std::string& bn, int K_phys, bool map_in, const std::string& name) {
            // 1x1 conv [co][ci] -> [co][K_phys] over PHYSICAL input channels, BN folded
            const HostTensor& W = c.W(conv + ".weight");
            const int co = (int)W.shape[0], ci = (int)W.shape[1];
            bn_affine(c, bn, scale, shift);
            std::vector<float> w((size_t)co * K_phys, 0.f);
            for (int o = 0; o < co; ++o)
                for (int i = 0; i < ci; ++i) {
                    int pi = map_in ? phys_ch(i, in_half, in_hp) : i;
                    w[(size_t)o * K_phys + pi] = W.f[(size_t)o * ci + i] * scale[o];
                }
            c.upload(name + ".w", w); c.upload(name + ".b", shift);
            upload_tc(c, name, w, co, 1, K_phys);
        };
        auto pack_dw = [&](const std::string& conv, const std::string& bn, int C_phys, bool map_in, const std::string& name) {
            const HostTensor& W = c.W(conv + ".weight");            // [C][1][3][3]
            const int C = (int)W.shape[0];
            bn_affine(c, bn, scale, shift);
            std::vector<float> w((size_t)9 * C_phys, 0.f), b(C_phys, 0.f);
            for (int ch = 0; ch < C; ++ch) {
                int pc = map_in ? phys_ch(ch, in_half, in_hp) : ch;
                for (int k = 0; k < 9; ++k) w[(size_t)k * C_phys + pc] = W.f[(size_t)ch * 9 + k] * scale[ch];
                b[pc] = shift[ch];
            }
            c.upload(name + ".w", w); c.upload(name + ".b", b);
        };
        if (down) {
            pack_dw(q + "banch1.0", q + "banch1.1", cin_phys, true, n + "b1dw");
            pack_pw(q + "banch1.2", q + "banch1.3", cin_phys, true, n + "b1pw");
            pack_pw(q + "banch2.0", q + "banch2.1", cin_phys, true, n + "b2pw1");
        } else {
            pack_pw(q + "banch2.0", q + "banch2.1", hp, false, n + "b2pw1");
        }
        pack_dw(q + "banch2.3", q + "banch2.4", hp, false, n + "b2dw");
        pack_pw(q + "banch2.5", q + "banch2.6", hp, false, n + "b2pw2");
        metas.push_back({down ? 1 : 0, cin_phys, half, hp, in_half, in_hp});
        cin = 2 * half; cin_phys = 2 * hp; in_half = half; in_hp = hp;
        ++blk;
    }
    {   // conv_last 1x1 cin->768 over physical channels
        const HostTensor& W = c.W(p + "trunk.1.0.weight");
        const int co = (int)W.shape[0], ci = (int)W.shape[1];
        std::vector<float> scale, shift; bn_affine(c, p + "trunk.1.1", scale, shift);
        std::vector<float> w((size_t)co * cin_phys, 0.f);
        for (int o = 0; o < co; ++o)
            for (int i = 0; i < ci; ++i) w[(size_t)o * cin_phys + phys_ch(i, in_half, in_hp)] = W.f[(size_t)o * ci + i] * scale[o];
        c.upload("v.last.w", w); c.upload("v.last.b", shift);
        upload_tc(c, "v.last", w, co, 1, cin_phys);
        c.meta["v.last.k"] = cin_phys; c.meta["v.last.n"] = co;
    }
    c.meta["v.nblocks"] = blk;
    for (int i = 0; i < blk; ++i) {
        const std::string n = "v.b" + std::to_string(i) + ".";
        c.meta[n + "down"] = metas[i].down; c.meta[n + "cin_phys"] = metas[i].cin_phys;
        c.meta[n + "half"] = metas[i].half; c.meta[n + "hp"] = metas[i].hp;
    }
    (void)cin;
}

// ------------------------------------------------------------------------------------------------
// LSTM chunking shared by the decoder's encoder_rnn and the speaker encoder
// ------------------------------------------------------------------------------------------------
struct LstmPack { std::vector<LstmBlock> blocks; std::vector<float> w; bool ok = true; };

// layers x dirs planes of H units; the CTAs are divided evenly among the planes and every CTA gets ONE block of
// <= LSTM_MAX_UNITS consecutive units of its plane.  get_w(layer, dir, which /*0 ih, 1 hh*/) returns the torch-layout
// [4H][in] matrix, get_b(layer, dir) the summed bias.
template <typename GW, typename GB>
inline LstmPack pack_lstm(int L, int dirs, int H, int num_ctas, GW get_w, GB get_b) {
    LstmPack pk;
    pk.blocks.assign(num_ctas, LstmBlock{});
    const int planes = L * dirs;
    int cta = 0;
    for (int l = 0; l < L; ++l)
        for (int d = 0; d < dirs; ++d) {
            const int pl = l * dirs + d;
            const int nc = num_ctas / planes + (pl < num_ctas % planes ? 1 : 0);      // CTAs of this plane
            const int per = (H + nc - 1) / nc;
            if (nc == 0 || per > LSTM_MAX_UNITS) { pk.ok = false; return pk; }
            const std::vector<float>* wih = (l == 0) ? nullptr : get_w(l, d, 0);
            const std::vector<float>* whh = get_w(l, d, 1);
            const std::vector<float> bsum = get_b(l, d);
            for (int j = 0; j < nc; ++j, ++cta) {
                LstmBlock& bk = pk.blocks[cta];
                bk.layer = l; bk.dir = d; bk.u0 = std::min(H, j * per); bk.nu = std::min(per, H - bk.u0);
                bk.K0 = H; bk.K1 = (l == 0) ? 0 : H;
                bk.w_off = (int)pk.w.size();
                const int K = bk.K0 + bk.K1;
                pk.w.resize(pk.w.size() + (size_t)4 * bk.nu * K, 0.f);
                for (int ul = 0; ul < bk.nu; ++ul)
                    for (int g = 0; g < 4; ++g) {
                        const int r = 4 * ul + g, row = g * H + bk.u0 + ul;
                        float* dst = pk.w.data() + bk.w_off + (size_t)r * K;
                        if (l == 0) {
                            std::copy(whh->begin() + (size_t)row * H, whh->begin() + (size_t)(row + 1) * H, dst);
                        } else {
                            std::copy(wih->begin() + (size_t)row * H, wih->begin() + (size_t)(row + 1) * H, dst);
                            std::copy(whh->begin() + (size_t)row * H, whh->begin() + (size_t)(row + 1) * H, dst + H);
                        }
                        bk.bias[r] = (l == 0) ? 0.f : bsum[row];
                    }
            }
        }
    return pk;
}

inline std::vector<float> vadd(const std::vector<float>& a, const std::vector<float>& b) {
    std::vector<float> r(a.size());
    for (size_t i = 0; i < a.size(); ++i) r[i] = a[i] + b[i];
    return r;
}

inline void pack_speaker(Context& c) {
    const std::string p = "speaker_encoder.";
    c.upload("s.window", c.W(p + "mel_spec.spectrogram.window").f);
    {   // DFT table for the GEMM form of the spectrogram: rows 0..200 cos(2 pi k n / 400), rows 201..401 -sin (sign irrelevant for |.|^2)
        std::vector<float> dft((size_t)402 * 400);
        for (int k = 0; k < 201; ++k)
            for (int n = 0; n < 400; ++n) {
                const int j = (int)(((long long)k * n) % 400);                 // exact argument reduction
                const double a = 2.0 * M_PI * (double)j / 400.0;
                dft[(size_t)k * 400 + n] = (float)std::cos(a);
                dft[(size_t)(201 + k) * 400 + n] = (float)std::sin(a);
            }
        c.upload("s.dft.w", dft);
        upload_tc(c, "s.dft", dft, 402, 1, 400);
    }
    c.upload("s.fb", c.W(p + "mel_spec.mel_scale.fb").f);
    c.upload("s.wih0.w", c.W(p + "lstm.weight_ih_l0").f);
    upload_tc(c, "s.wih0", c.W(p + "lstm.weight_ih_l0").f, 1024, 1, 40);
    c.upload("s.b0", vadd(c.W(p + "lstm.bias_ih_l0").f, c.W(p + "lstm.bias_hh_l0").f));
    c.upload("s.lin.w", c.W(p + "linear.weight").f);
    c.upload("s.lin.b", c.W(p + "linear.bias").f);
    auto gw = [&](int l, int, int which) { return &c.W(p + "lstm.weight_" + (which ? "hh" : "ih") + "_l" + std::to_string(l)).f; };
    auto gb = [&](int l, int) { return vadd(c.W(p + "lstm.bias_ih_l" + std::to_string(l)).f, c.W(p + "lstm.bias_hh_l" + std::to_string(l)).f); };
    LstmPack pk = pack_lstm(3, 1, 256, c.num_sms, gw, gb);
    if (!pk.ok) throw L2sError(1, "speaker LSTM does not fit this SM count");
    c.upload("s.lstm.w", pk.w);
    c.upload_raw("s.lstm.blocks", pk.blocks.data(), pk.blocks.size());
}

// ------------------------------------------------------------------------------------------------
// decoder: pre-loop, postnet, decode-step program
// ------------------------------------------------------------------------------------------------
// Row sources of the decode step after the host-side linear∘linear merges (pointers into vectors that outlive the packers).
struct StepRows {
    const float *Wfc, *bfc, *Wpf, *bpf, *ps1, *p1bos, *Wp2, *bp2, *ps2, *Wq, *bq, *psq, *Wcq, *bcq, *Wst, *bst;
    const float *Wih0, *Wx, *Whh0, *b0x, *Wih1, *Whh1, *b1;
};

// Program for the stage-pipelined kernel (decode3.cuh): every CTA serves ONE stage with ONE pass.  LSTM-0: 8 units per
// CTA (32 gate rows x 1536); LSTM-1: 12 units (48 x 1024); stage A: query rows (48 x 1024), content-query rows
// (43 x 1024) and fc_out / fc_out∘prenet-1 / stop rows (68 x 512) on separate CTAs; stage B: 8 clips x nsplit attention
// CTAs + prenet-2 rows (86 x 256).  The split follows the measured per-turn critical path (tools/dec3_debug.py).  Row widths are padded by 16 floats so the tensor-core fragment loads (LDS.128,
// rows g / g+8) are bank-conflict free.
inline void pack_decode_program3(Context& c, const StepRows& w) {
    const int nC = c.num_sms;
    const int nD = 64, nE = 43, nQ = 11, nCQ = 6, nF = 5, nP2 = 3, nsplit = D3_NSPLIT;
    const int nAttn = D3_CG * nsplit;
    c.meta["d.step3.ok"] = 0;
    if (nD + nE + nQ + nCQ + nF + nP2 + nAttn > nC) return;    // not enough SMs (the decoder then refuses to run: B200 has 148)
    struct Row { int op, idx; float bias, aux, aux2; std::vector<std::pair<const float*, int>> w; };
    struct PB { int Ke, src_e, wcol_e, Kl, src_l, wcol_l; std::vector<Row> rows; };
    std::vector<PB> per_cta(nC, PB{0, SRC_NONE, 0, 0, SRC_NONE, 0, {}});
    std::vector<int> role(nC, ROLE_B), job(nC, -1);
    int cta = 0;
    for (int j = 0; j < nD; ++j, ++cta) {                      // LSTM-0: [W_ih0[:, :256] | W_ih0[:,256:] | W_ih0[:,256:] W_ap | W_hh0] x [cv; p2; ctx; h0]
        role[cta] = ROLE_D;
        PB pb{512, SRC_H0OLD, 1024, 1024, SRC_XD, 0, {}};
        for (int u = 8 * j; u < std::min(512, 8 * j + 8); ++u)
            for (int g = 0; g < 4; ++g) {
                const int row = g * 512 + u;
                pb.rows.push_back({OP_GATE0, u, w.b0x[row], 0.f, 0.f,
                                   {{w.Wih0 + (size_t)row * 512, 512}, {w.Wx + (size_t)row * 512, 512}, {w.Whh0 + (size_t)row * 512, 512}}});
            }
        per_cta[cta] = pb;
    }
    for (int j = 0; j < nE; ++j, ++cta) {                      // LSTM-1: [W_ih1 | W_hh1] x [h0'; h1]
        role[cta] = ROLE_E;
        PB pb{512, SRC_H1OLD, 512, 512, SRC_H0NEW, 0, {}};
        for (int u = 12 * j; u < std::min(512, 12 * j + 12); ++u)
            for (int g = 0; g < 4; ++g) {
                const int row = g * 512 + u;
                pb.rows.push_back({OP_GATE1, u, w.b1[row], 0.f, 0.f, {{w.Wih1 + (size_t)row * 512, 512}, {w.Whh1 + (size_t)row * 512, 512}}});
            }
        per_cta[cta] = pb;
    }
    {
        const int qn = ceil_div(512, nQ), cn = ceil_div(256, nCQ), fn = ceil_div(337, nF);
        for (int j = 0; j < nQ; ++j, ++cta) {
            role[cta] = ROLE_A;
            PB q{512, SRC_H0NEW, 0, 512, SRC_H1NEW, 512, {}};
            for (int r = qn * j; r < std::min(512, qn * (j + 1)); ++r) q.rows.push_back({OP_Q, r, w.bq[r], w.psq[r], 0.f, {{w.Wq + (size_t)r * 1024, 1024}}});
            per_cta[cta] = q;
        }
        for (int j = 0; j < nCQ; ++j, ++cta) {
            role[cta] = ROLE_A;
            PB cq{512, SRC_C0, 0, 512, SRC_C1, 512, {}};
            for (int r = cn * j; r < std::min(256, cn * (j + 1)); ++r) cq.rows.push_back({OP_CQ, r, w.bcq[r], 0.f, 0.f, {{w.Wcq + (size_t)r * 1024, 1024}}});
            per_cta[cta] = cq;
        }
        for (int j = 0; j < nF; ++j, ++cta) {
            role[cta] = ROLE_A;
            PB f{0, SRC_NONE, 0, 512, SRC_H1NEW, 0, {}};
            for (int r = fn * j; r < std::min(337, fn * (j + 1)); ++r) {
                if (r < 80) f.rows.push_back({OP_FC, r, w.bfc[r], 0.f, 0.f, {{w.Wfc + (size_t)r * 512, 512}}});
                else if (r < 336) f.rows.push_back({OP_P1, r - 80, w.bpf[r - 80], w.ps1[r - 80], w.p1bos[r - 80], {{w.Wpf + (size_t)(r - 80) * 512, 512}}});
                else f.rows.push_back({OP_STOP, 0, w.bst[0], 0.f, 0.f, {{w.Wst, 512}}});
            }
            per_cta[cta] = f;
        }
    }
    for (int j = 0; j < nAttn; ++j, ++cta) job[cta] = j;       // attention CTAs hold no weights
    {
        const int pn = ceil_div(256, nP2);
        for (int j = 0; j < nP2; ++j, ++cta) {
            PB pb{0, SRC_NONE, 0, 256, SRC_P1, 0, {}};
            for (int r = pn * j; r < std::min(256, pn * (j + 1)); ++r)
                pb.rows.push_back({OP_P2, r, w.bp2[r], w.ps2[r], 0.f, {{w.Wp2 + (size_t)r * 256, 256}}});
            per_cta[cta] = pb;
        }
    }
    size_t wimg_floats = 512 + 320 + 256 + 32;                 // attention scratch shares the weight region
    for (int i = 0; i < nC; ++i) {
        if (per_cta[i].rows.size() > (size_t)D3_ROWS) return;
        if (per_cta[i].Ke > 256 * D3_EDEPTH || per_cta[i].Kl > 256 * MV8_DEPTH) return;
        if (per_cta[i].Ke > 0 && per_cta[i].rows.size() > 48) return;      // early segments are instantiated for <= 3 row tiles
        wimg_floats = std::max(wimg_floats, per_cta[i].rows.size() * (size_t)(per_cta[i].Ke + per_cta[i].Kl + 16));
    }
    const size_t smem = (wimg_floats + MV_WARPS * MV8_RTILES * 128 + 16 * MV8_RTILES * D3_CG) * sizeof(float);
    if (smem + sizeof(Dec3Pass) + 1024 > (size_t)c.max_smem_optin) return;
    std::vector<float> wimg((size_t)nC * wimg_floats, 0.f);
    std::vector<Dec3Pass> passes(nC);
    std::memset(passes.data(), 0, passes.size() * sizeof(Dec3Pass));
    for (int i = 0; i < nC; ++i) {
        const PB& pb = per_cta[i];
        Dec3Pass& d = passes[i];
        d.R = (int)pb.rows.size(); d.RT = ceil_div(d.R, 16);
        d.Ke = pb.Ke; d.src_e = pb.src_e; d.wcol_e = pb.wcol_e; d.Kl = pb.Kl; d.src_l = pb.src_l; d.wcol_l = pb.wcol_l;
        d.ldw = pb.Ke + pb.Kl + 16; d.wfloats = d.R * d.ldw;
        for (int r = 0; r < D3_ROWS; ++r) {
            if (r < d.R) {
                const Row& row = pb.rows[r];
                d.op[r] = row.op; d.idx[r] = row.idx; d.bias[r] = row.bias; d.aux[r] = row.aux; d.aux2[r] = row.aux2;
                float* dst = wimg.data() + (size_t)i * wimg_floats + (size_t)r * d.ldw;
                int col = 0;
                for (auto& piece : row.w) { std::copy(piece.first, piece.first + piece.second, dst + col); col += piece.second; }
                if (col != pb.Ke + pb.Kl) throw L2sError(1, "internal: decode3 row width mismatch");
                for (int c0 = 0; c0 < col; c0 += 16) {         // fragment order inside every 16-chunk: position 4t+i <- column 4i+t
                    float tmp[16];
                    std::copy(dst + c0, dst + c0 + 16, tmp);
                    for (int t = 0; t < 4; ++t)
                        for (int i2 = 0; i2 < 4; ++i2) dst[c0 + 4 * t + i2] = tmp[4 * i2 + t];
                }
            } else { d.op[r] = OP_NONE; d.idx[r] = -1; }
        }
    }
    c.upload("d.step3.wimg", wimg);
    c.upload_raw("d.step3.passes", passes.data(), passes.size());
    c.upload_raw("d.step3.role", role.data(), role.size());
    c.upload_raw("d.step3.job", job.data(), job.size());
    c.meta["d.step3.wimg_floats"] = (int64_t)wimg_floats;
    c.meta["d.step3.smem"] = (int64_t)smem;
    c.meta["d.step3.ok"] = 1;
}

inline void pack_decode_program(Context& c) {
    const std::string p = "decoder.";
    const int nC = c.num_sms;
    const auto& Wfc = c.W(p + "fc_out.linear_layer.weight").f;   const auto& bfc = c.W(p + "fc_out.linear_layer.bias").f;
    const auto& Wp1 = c.W(p + "prenet.0.linear_layer.weight").f; const auto& bp1 = c.W(p + "prenet.0.linear_layer.bias").f;
    const auto& ps1 = c.W(p + "prenet.1.w").f;
    const auto& Wp2 = c.W(p + "prenet.3.linear_layer.weight").f; const auto& bp2 = c.W(p + "prenet.3.linear_layer.bias").f;
    const auto& ps2 = c.W(p + "prenet.4.w").f;
    const auto& Wq = c.W(p + "Q.0.linear_layer.weight").f;       const auto& bq = c.W(p + "Q.0.linear_layer.bias").f;
    const auto& psq = c.W(p + "Q.1.w").f;
    const auto& Wcq = c.W(p + "content.Q.0.weight").f;           const auto& bcq = c.W(p + "content.Q.0.bias").f;
    const auto& Wap = c.W(p + "attention_proj.linear_layer.weight").f; const auto& bap = c.W(p + "attention_proj.linear_layer.bias").f;
    const auto& Wst = c.W(p + "stop_token_layer.linear_layer.weight").f; const auto& bst = c.W(p + "stop_token_layer.linear_layer.bias").f;
    const auto& bos = c.W(p + "BOS").f;
    // prenet layer 1 fused with fc_out (linear o linear): Wpf = Wp1 * Wfc [256x512], bpf = Wp1*bfc + bp1
    std::vector<float> Wpf((size_t)256 * 512, 0.f), bpf(256, 0.f), p1bos(256, 0.f);   // outlive the row pointers below
    for (int j = 0; j < 256; ++j) {
        std::vector<double> acc(512, 0.0);
        double bb = bp1[j], pb = bp1[j];
        for (int m = 0; m < 80; ++m) {
            const double w = Wp1[(size_t)j * 80 + m];
            const float* fr = Wfc.data() + (size_t)m * 512;
            for (int k = 0; k < 512; ++k) acc[k] += w * fr[k];
            bb += w * bfc[m];
            pb += w * bos[m];
        }
        for (int k = 0; k < 512; ++k) Wpf[(size_t)j * 512 + k] = (float)acc[k];
        bpf[j] = (float)bb;
        p1bos[j] = std::sin((float)pb) * ps1[j];
    }
    std::vector<float> b0 = vadd(c.W(p + "decoder_rnn.bias_ih_l0").f, c.W(p + "decoder_rnn.bias_hh_l0").f);
    std::vector<float> b1 = vadd(c.W(p + "decoder_rnn.bias_ih_l1").f, c.W(p + "decoder_rnn.bias_hh_l1").f);
    const auto& Wih0 = c.W(p + "decoder_rnn.weight_ih_l0").f; const auto& Whh0 = c.W(p + "decoder_rnn.weight_hh_l0").f;
    const auto& Wih1 = c.W(p + "decoder_rnn.weight_ih_l1").f; const auto& Whh1 = c.W(p + "decoder_rnn.weight_hh_l1").f;
    // attention_proj folded into LSTM-0 (linear o linear): the LSTM input is cat([cv, p2 + W_ap ctx + b_ap]), so
    // W_ih0[:,256:] (p2 + W_ap ctx + b_ap) = W_ih0b p2 + (W_ih0b W_ap) ctx + W_ih0b b_ap.
    std::vector<float> Wx((size_t)2048 * 512, 0.f), b0x(2048, 0.f);
    {
        std::vector<double> acc(512);
        for (int row = 0; row < 2048; ++row) {
            std::fill(acc.begin(), acc.end(), 0.0);
            double bb = b0[row];
            const float* wb = Wih0.data() + (size_t)row * 512 + 256;
            for (int m = 0; m < 256; ++m) {
                const double w = wb[m];
                const float* ar = Wap.data() + (size_t)m * 512;
                for (int k = 0; k < 512; ++k) acc[k] += w * ar[k];
                bb += w * bap[m];
            }
            for (int k = 0; k < 512; ++k) Wx[(size_t)row * 512 + k] = (float)acc[k];
            b0x[row] = (float)bb;
        }
    }

    // stop token: the encoder_cell half of the weight row (runtime GEMM N=1) — bias already in the pass
    std::vector<float> wst2(Wst.begin() + 512, Wst.begin() + 1024);
    c.upload("d.stop2.w", wst2);
    upload_tc(c, "d.stop2", wst2, 1, 1, 512);
    StepRows sr{Wfc.data(), bfc.data(), Wpf.data(), bpf.data(), ps1.data(), p1bos.data(), Wp2.data(), bp2.data(), ps2.data(),
                Wq.data(), bq.data(), psq.data(), Wcq.data(), bcq.data(), Wst.data(), bst.data(),
                Wih0.data(), Wx.data(), Whh0.data(), b0x.data(), Wih1.data(), Whh1.data(), b1.data()};
    pack_decode_program3(c, sr);
}


inline void pack_decoder(Context& c) {
    const std::string p = "decoder.";
    std::vector<float> w, b;
    // postnet
    for (int i = 0; i < 5; ++i) {
        const std::string s = std::to_string(i);
        pack_conv1d(c, p + "postnet.convolutions." + s + ".0.conv", p + "postnet.convolutions." + s + ".1", w, b);
        c.upload("d.post" + s + ".w", w); c.upload("d.post" + s + ".b", b);
        upload_tc(c, "d.post" + s, w, (int)b.size(), 5, (int)(w.size() / b.size() / 5));
        if (i < 4) c.upload("d.post" + s + ".psw", c.W(p + "postnet.sin_activation." + s + ".w").f);
    }
    // encoder pre-loop linears
    auto up_lin = [&](const std::string& key, const std::string& name) {
        const HostTensor& W = c.W(key + ".weight");
        c.upload(name + ".w", W.f); c.upload(name + ".b", c.W(key + ".bias").f);
        upload_tc(c, name, W.f, (int)W.shape[0], 1, (int)W.shape[1]);
    };
    up_lin(p + "encoder_proj.linear_layer", "d.encproj");
    up_lin(p + "encoder_site.0.linear_layer", "d.encsite"); c.upload("d.encsite.psw", c.W(p + "encoder_site.1.w").f);
    up_lin(p + "attention_site.0.linear_layer", "d.attsite"); c.upload("d.attsite.psw", c.W(p + "attention_site.1.w").f);
    up_lin(p + "residual_bottleneck", "d.resid");            // [512][1024][1] == [512][1024]
    up_lin(p + "E_C.linear_layer", "d.ec");
    {   // Bi-LSTM: input projection for both directions as one [4096][1024] GEMM, recurrent chunks
        std::vector<float> wih = c.W(p + "encoder_rnn.weight_ih_l0").f;
        const auto& wr = c.W(p + "encoder_rnn.weight_ih_l0_reverse").f;
        wih.insert(wih.end(), wr.begin(), wr.end());
        std::vector<float> bsum = vadd(c.W(p + "encoder_rnn.bias_ih_l0").f, c.W(p + "encoder_rnn.bias_hh_l0").f);
        std::vector<float> br = vadd(c.W(p + "encoder_rnn.bias_ih_l0_reverse").f, c.W(p + "encoder_rnn.bias_hh_l0_reverse").f);
        bsum.insert(bsum.end(), br.begin(), br.end());
        c.upload("d.ernn.wih.w", wih); c.upload("d.ernn.b", bsum);
        upload_tc(c, "d.ernn.wih", wih, 4096, 1, 1024);
        auto gw = [&](int, int d, int which) { return &c.W(p + "encoder_rnn.weight_" + (which ? "hh" : "ih") + "_l0" + (d ? "_reverse" : "")).f; };
        auto gb = [&](int, int) { return std::vector<float>(); };
        LstmPack pk = pack_lstm(1, 2, 512, c.num_sms, gw, gb);
        if (!pk.ok) throw L2sError(1, "encoder LSTM does not fit this SM count");
        c.upload("d.ernn.w", pk.w);
        c.upload_raw("d.ernn.blocks", pk.blocks.data(), pk.blocks.size());
    }
    // K / V MultiHopConv: the four convolutions of K and of V read the same input, so each pair is packed as ONE
    // conv with 1024 output channels (rows 0..511 = K branch, 512..1023 = V branch)
    for (int j = 0; j < 4; ++j) {
        const std::string s = std::to_string(j);
        std::vector<float> wk, bk, wv, bv;
        pack_conv1d(c, p + "K.0.conv." + s + ".0", p + "K.0.conv." + s + ".1", wk, bk);
        pack_conv1d(c, p + "V.0.conv." + s + ".0", p + "V.0.conv." + s + ".1", wv, bv);
        wk.insert(wk.end(), wv.begin(), wv.end()); bk.insert(bk.end(), bv.begin(), bv.end());
        c.upload("d.KV.c" + s + ".w", wk); c.upload("d.KV.c" + s + ".b", bk);
        upload_tc(c, "d.KV.c" + s, wk, 1024, (int)(wk.size() / 1024 / 512), 512);
    }
    for (const char* kv : {"K", "V"}) {
        pack_conv1d(c, p + kv + ".0.bottleneck", "", w, b);
        c.upload(std::string("d.") + kv + ".bn.w", w); c.upload(std::string("d.") + kv + ".bn.b", b);
        upload_tc(c, std::string("d.") + kv + ".bn", w, 512, 1, 2560);
        c.upload(std::string("d.") + kv + ".psw", c.W(p + kv + ".1.w").f);
    }
    // Content.encode
    for (int j = 0; j < 4; ++j) {
        const std::string s = std::to_string(j);
        pack_conv1d(c, p + "content.agg." + s + ".0", p + "content.agg." + s + ".1", w, b);
        c.upload("d.cagg" + s + ".w", w); c.upload("d.cagg" + s + ".b", b);
        upload_tc(c, "d.cagg" + s, w, 512, 1, (int)(w.size() / 512));       // k == stride: a dense GEMM over k*512-wide rows
    }
    pack_conv1d(c, p + "content.bottleneck", "", w, b);
    c.upload("d.cbn.w", w); c.upload("d.cbn.b", b);
    upload_tc(c, "d.cbn", w, 256, 1, 2560);
    up_lin(p + "content.location_fc.0", "d.cloc0"); up_lin(p + "content.location_fc.2", "d.cloc2"); up_lin(p + "content.location_fc.4", "d.cloc4");
    up_lin(p + "content.K.0", "d.ck0"); up_lin(p + "content.K.2", "d.ck2");
    c.upload("d.cemb", c.W(p + "content.word_embeddings").f);
    c.upload("d.bos", c.W(p + "BOS").f);
    up_lin(p + "prenet.0.linear_layer", "d.prenet0"); c.upload("d.prenet0.psw", c.W(p + "prenet.1.w").f);   // teacher-forced steps
    c.upload("d.pos", c.W(p + "positional_encodings.pos_table").f);
    c.meta["d.temp_bits"] = 0;
    pack_decode_program(c);
}

}  // namespace l2s
