"""Pin the CPU oracle (oracle/l2s_oracle.py) to the outputs of the unmodified reference modules
(tests/golden/golden_synthetic.pt, made by tests/golden/make_golden.py)."""
import hashlib

import pytest
import torch

from conftest import rel_err
from lip2speech_b200 import spec, synth
from oracle import l2s_oracle as O

TOL = 2e-5   # oracle vs reference: same torch build, only op-order differences


def test_seeded_weights_match_golden_sha(golden, weights):
    for k, h in golden["weight_sha"].items():
        got = hashlib.sha256(weights[k].contiguous().numpy().tobytes()).hexdigest()[:16]
        assert got == h, f"seeded tensor {k} differs from the one the goldens were made with"


def test_speaker_encoder(golden, spk_weights):
    raw = O.speaker_forward(spk_weights, synth.wav(2))
    assert rel_err(raw, golden["A_spk_raw"]) < TOL
    emb = O.speaker_inference(spk_weights, synth.wav(2))
    assert rel_err(emb, golden["A_spk_emb"]) < TOL
    raw = O.speaker_forward(spk_weights, synth.wav(1, 48000, seed=9))
    assert rel_err(raw, golden["F_spk_raw"]) < TOL


def test_video_features(golden, weights):
    f = O.video_features(weights, synth.video(2, 29))
    assert f.shape == (2, 29, 768)
    assert rel_err(f, golden["A_video_feat"]) < TOL
    f = O.video_features(weights, synth.video(1, 5, 88, 88, seed=5))
    assert rel_err(f, golden["D_video_feat"]) < TOL


def test_postnet(golden, weights):
    y = O.postnet(weights, synth.mel_like(2, 77))
    assert rel_err(y, golden["E_postnet"]) < TOL


def test_decoder_inference_t29(golden, weights):
    visual, face = synth.visual_features(3, 29)
    mel, lengths, attn = O.decoder_inference(weights, visual, face, synth.gumbel(3, 29), return_attention=True)
    assert mel.shape == (3, 80, 300) and attn.shape == (3, 300, 29)
    assert torch.equal(lengths, golden["B_lengths"])
    assert rel_err(mel, golden["B_mel"]) < 1e-4
    assert rel_err(attn, golden["B_attn"]) < 2e-3   # one-hot-sharp softmax (temperature sqrt(512)): fp32 op-order noise x ~200
    pre = O.decoder_preloop(weights, visual, face[:, 0], synth.gumbel(3, 29))
    assert rel_err(pre["ckey"], golden["B_content_key"]) < TOL
    assert rel_err(pre["cval"], golden["B_content_value"]) < 1e-4


def test_decoder_inference_t75(golden, weights):
    visual, face = synth.visual_features(1, 75, seed=77)
    mel, lengths = O.decoder_inference(weights, visual, face, synth.gumbel(1, 75, seed=77))
    assert torch.equal(lengths, golden["C_lengths"])
    assert rel_err(mel, golden["C_mel"]) < 1e-4


def test_full_span(golden, weights, spk_weights):
    mel, lengths = O.demo_span(weights, spk_weights, synth.video(2, 29), synth.wav(2), synth.gumbel(2, 29))
    assert torch.equal(lengths, golden["A_lengths"])
    assert rel_err(mel, golden["A_mel"]) < 1e-4


def test_lstm_op_matches_explicit_cell(weights):
    """The ATen lstm op used by the oracle == the explicit i,f,g,o cell."""
    p = "decoder.decoder_rnn."
    x = torch.randn(4, 1, 512, generator=torch.Generator().manual_seed(0))
    h0 = torch.randn(2, 4, 512, generator=torch.Generator().manual_seed(1)) * 0.1
    c0 = torch.randn(2, 4, 512, generator=torch.Generator().manual_seed(2)) * 0.1
    out, h, c = O._lstm(x, h0, c0, weights, "decoder.decoder_rnn", 2)
    h_a, c_a = O.lstm_cell(x[:, 0], h0[0], c0[0], *[weights[p + n] for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")])
    h_b, c_b = O.lstm_cell(h_a, h0[1], c0[1], *[weights[p + n] for n in ("weight_ih_l1", "weight_hh_l1", "bias_ih_l1", "bias_hh_l1")])
    assert rel_err(h[1], h_b) < 1e-5 and rel_err(c[0], c_a) < 1e-5


def test_min_t():
    assert spec.content_min_t(29) == 4 and spec.content_min_t(75) == 10


def test_decoder_forward_eval(golden, weights):
    """Decoder.forward in eval mode (evaluate.py path): explicit teacher-forcing mask == the reference's seeded coin flips."""
    visual, face = synth.visual_features(2, 29, seed=11)
    g = synth.gumbel(2, 29, seed=11)
    mels = synth.mel_like(2, 24, seed=11)
    for name, tf in (("G5", 0.5), ("G1", 1.0)):
        mask = O.teacher_forcing_mask(tf, 24, torch.Generator().manual_seed(4321))
        if tf == 1.0:
            assert not mask.any()
        else:
            assert mask.any()
        o = O.decoder_forward(weights, visual, face, mels, mask, g)
        assert rel_err(o[0], golden[name + "_outputs"]) < 1e-4
        assert rel_err(o[1], golden[name + "_post"]) < 1e-4
        assert rel_err(o[2], golden[name + "_stop"]) < 1e-4
        assert rel_err(o[4], golden[name + "_attn_logits"]) < 1e-4
        assert rel_err(o[5], golden[name + "_cdis"]) < 1e-4
