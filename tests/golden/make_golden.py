"""Generate golden vectors by running the UNMODIFIED reference modules (build container only).

    python tests/golden/make_golden.py

Imports /root/reference (with the sys.modules shims of oracle/ref_import.py), loads the seeded
weights of lip2speech_b200.spec with strict=True into the reference's own Decoder /
VideoExtractor / SpeakerEncoder, replaces F.gumbel_softmax by the explicit-noise form
(bit-identical to the seeded original) and stores what the reference returns.
The outputs (not the 154 MB of weights, which are a pure function of the seed) are committed
under tests/golden/ and are what pins oracle/l2s_oracle.py and the CUDA path.
"""
import hashlib
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from lip2speech_b200 import spec, synth          # noqa: E402
from oracle import ref_import                     # noqa: E402

SEED = 1234


def build_reference():
    Decoder, VideoExtractor, SpeakerEncoder, Lip2Speech = ref_import.import_reference()
    import torchaudio.transforms as AT

    dec, vid = Decoder().eval(), VideoExtractor().eval()
    dec.load_state_dict(spec.seeded_state_dict(spec.decoder_spec(), SEED), strict=True)
    vid.load_state_dict(spec.seeded_state_dict(spec.encoder_spec(), SEED), strict=True)
    spk_sd = spec.seeded_state_dict(spec.speaker_spec(), SEED)
    spk = SpeakerEncoder(state_dict=spk_sd).eval()       # ctor loads with strict=True (audio.py:129)
    return dec, vid, spk


class ExplicitGumbel:
    """Route decoder.py:257 `F.gumbel_softmax(w_y, 0.1, dim=-1)` to an explicit noise tensor."""

    def __init__(self, noise):
        self.noise = noise

    def __enter__(self):
        self.orig = F.gumbel_softmax
        F.gumbel_softmax = lambda logits, tau=1, hard=False, eps=1e-10, dim=-1: ((logits + self.noise) / tau).softmax(dim)

    def __exit__(self, *a):
        F.gumbel_softmax = self.orig


def sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()[:16]


def main():
    torch.set_grad_enabled(False)
    dec, vid, spk = build_reference()

    man = {"decoder": {k: list(v.shape) for k, v in dec.state_dict().items()},
           "encoder": {k: list(v.shape) for k, v in vid.state_dict().items()},
           "speaker_encoder": {k: list(v.shape) for k, v in spk.state_dict().items()}}
    with open(os.path.join(HERE, "state_manifest.json"), "w") as f:
        json.dump(man, f, indent=0, sort_keys=True)

    full = spec.seeded_state_dict(spec.full_spec(), SEED)
    weight_sha = {k: sha(full[k]) for k in ("encoder.frontend3D.0.weight", "decoder.Q.0.linear_layer.weight",
                                            "decoder.decoder_rnn.weight_hh_l1", "speaker_encoder.lstm.weight_ih_l0",
                                            "decoder.positional_encodings.pos_table",
                                            "speaker_encoder.mel_spec.mel_scale.fb")}

    # sanity: the fixed buffers of the spec equal what the reference/torchaudio construct
    Decoder, VideoExtractor, SpeakerEncoder, _ = ref_import.import_reference()
    ref_pos = Decoder().state_dict()["positional_encodings.pos_table"]
    assert torch.equal(ref_pos, full["decoder.positional_encodings.pos_table"]), "pos_table mismatch"
    import torchaudio.transforms as AT
    ms = AT.MelSpectrogram(sample_rate=16000, n_fft=400, hop_length=160, n_mels=40)
    assert torch.allclose(ms.mel_scale.fb, full["speaker_encoder.mel_spec.mel_scale.fb"], atol=1e-5)  # torchaudio builds it in fp32
    assert torch.allclose(ms.spectrogram.window, full["speaker_encoder.mel_spec.spectrogram.window"], atol=1e-7)

    out = {"seed": SEED, "weight_sha": weight_sha}

    # --- case A: B=2, T=29, 96x96, full demo span (speaker enc -> video -> decoder) --------
    B, T = 2, 29
    video, wav, g = synth.video(B, T), synth.wav(B), synth.gumbel(B, T)
    emb = spk.inference(wav)
    out["A_spk_raw"] = spk(wav)
    out["A_spk_emb"] = emb
    out["A_video_feat"] = vid(video)
    face = emb.unsqueeze(1).repeat(1, T, 1)
    visual = torch.cat([out["A_video_feat"], face], dim=2)
    with ExplicitGumbel(g):
        mel, lengths, attn = dec.inference(visual, face, return_attention_map=True)
    out["A_mel"], out["A_lengths"], out["A_attn"] = mel, lengths, attn
    # unpatched + seeded must give the same thing (A.7): gumbel is the first RNG consumer
    # (checked for the decoder-only case below where we control the seed)

    # --- case B: decoder only, B=3, T=29, synthetic visual features -------------------------
    B = 3
    visual, face = synth.visual_features(B, T)
    g = synth.gumbel(B, T)
    with ExplicitGumbel(g):
        mel, lengths, attn = dec.inference(visual, face, return_attention_map=True)
    out["B_mel"], out["B_lengths"], out["B_attn"] = mel, lengths, attn
    out["B_content_key"] = dec.content.key.clone()
    out["B_content_value"] = dec.content.value.clone()

    # --- case C: decoder only, B=1, T=75 (AVSpeech shape, minT=10) --------------------------
    visual, face = synth.visual_features(1, 75, seed=77)
    g = synth.gumbel(1, 75, seed=77)
    with ExplicitGumbel(g):
        mel, lengths = dec.inference(visual, face)
    out["C_mel"], out["C_lengths"] = mel, lengths

    # --- case D: video frontend at 88x88 (north_star wording), B=1, T=5 ---------------------
    out["D_video_feat"] = vid(synth.video(1, 5, 88, 88, seed=5))

    # --- case E: postnet alone -------------------------------------------------------------
    x = synth.mel_like(2, 77)
    out["E_postnet"] = dec.postnet(x.clone())

    # --- case F: speaker encoder on a longer utterance (S=48000 -> 301 frames) ---------------
    out["F_spk_raw"] = spk(synth.wav(1, 48000, seed=9))

    # --- case G: Decoder.forward in eval mode (evaluate.py path), tf_ratio 0.5 and 1.0, B=2, M=24 ---------------
    visual, face = synth.visual_features(2, 29, seed=11)
    g = synth.gumbel(2, 29, seed=11)
    mels = synth.mel_like(2, 24, seed=11)
    lens = torch.full((2,), 29, dtype=torch.long)
    for name, tf in (("G5", 0.5), ("G1", 1.0)):
        torch.manual_seed(4321)                       # the per-step torch.rand(1) coin of decoder.py:355 uses the CPU generator
        with ExplicitGumbel(g):
            o = dec(visual, face, mels, lens, lens, tf)
        out[name + "_outputs"], out[name + "_post"], out[name + "_stop"] = o[0], o[1], o[2]
        out[name + "_attn_logits"], out[name + "_cdis"] = o[4], o[5]

    # --- A.7 check: seeded, unpatched reference == explicit-noise reference ------------------
    visual, face = synth.visual_features(2, 29, seed=3)
    torch.manual_seed(99)
    g = -torch.empty(2 * 4, 501).exponential_().log()
    with ExplicitGumbel(g):
        m1, _ = dec.inference(visual, face)
    torch.manual_seed(99)
    m2, _ = dec.inference(visual, face)
    assert torch.equal(m1, m2), "explicit gumbel noise is not bit-identical to the seeded reference"

    out = {k: (v.detach().clone().contiguous() if torch.is_tensor(v) else v) for k, v in out.items()}
    torch.save(out, os.path.join(HERE, "golden_synthetic.pt"))
    for k, v in out.items():
        if torch.is_tensor(v):
            print(f"{k:20s} {tuple(v.shape)} absmax={v.abs().max().item():.4f}" if v.is_floating_point() else f"{k:20s} {v.tolist()}")
    print("bytes:", os.path.getsize(os.path.join(HERE, "golden_synthetic.pt")))


if __name__ == "__main__":
    main()
