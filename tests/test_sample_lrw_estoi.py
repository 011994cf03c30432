"""Real inputs: the reference's 10 SAMPLE_LRW clips (tests/golden/sample_lrw.npz, extracted by make_sample_lrw.py).
North-star criterion: ESTOI computed from the CUDA path's mels equals ESTOI computed from the reference-path (oracle) mels
to +-0.001 — same weights, same gumbel noise, same vocoder + seed, same ESTOI implementation.

Vocoder setting: Griffin-Lim is run for 32 iterations here, not the 256 of demo.py/evaluate.py.  Measured in the build
container: at 256 iterations the phase retrieval is chaotic — perturbing the REFERENCE's own mel by 3e-5 relative (fp32
summation-order noise level of a 50-layer network) changes the waveform by 30 % and ESTOI by up to 4e-3, so "+-0.001" is
not a property any two fp32 implementations share at that setting; at 32 iterations the same perturbation moves ESTOI by
<= 4e-4, which makes the criterion meaningful."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from lip2speech_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def load_sample_lrw():
    d = np.load(os.path.join(HERE, "golden", "sample_lrw.npz"))
    v = d["video"].astype(np.float32) / 255.0                       # datasets/lrw/dataset.py:82-86
    v = (v - MEAN) / STD
    video = torch.from_numpy(v).permute(0, 4, 1, 2, 3).contiguous()  # [B,3,T,H,W]
    return video, torch.from_numpy(d["audio"])


def test_fixture_shapes_and_estoi_sanity():
    from oracle import audio_metrics as AM
    video, audio = load_sample_lrw()
    assert video.shape == (10, 3, 29, 96, 96) and audio.shape == (10, 19456)
    a = audio[0].numpy()
    assert AM.estoi(a, a) > 0.99                                      # identical signals
    rng = np.random.default_rng(0)
    assert AM.estoi(a, rng.standard_normal(a.shape)) < 0.3           # unrelated noise


@pytest.mark.gpu
def test_estoi_matches_reference_path(weights, spk_weights):
    from lip2speech_b200 import _lib, build
    from oracle import audio_metrics as AM
    from oracle import l2s_oracle as O
    build.build()
    be = _lib.backend(0)
    be.bind_state_dict(weights, "", 7)
    video, audio = load_sample_lrw()
    g = synth.gumbel(10, 29, seed=2024)
    mel_gpu, len_gpu = be.infer(video.cuda(), audio.cuda(), g.cuda())
    mel_ref, len_ref = O.demo_span(weights, spk_weights, video, audio, g)
    assert torch.equal(len_gpu.cpu(), len_ref)
    assert rel_err(mel_gpu.cpu(), mel_ref) < 1e-3
    L = 77                                                            # LRW clips carry 77 mel frames (S=19456, hop 256)
    wav_gpu = AM.mel_to_audio(mel_gpu.cpu()[:, :, :L], seed=7, n_iter=32).numpy()
    wav_ref = AM.mel_to_audio(mel_ref[:, :, :L], seed=7, n_iter=32).numpy()
    e_gpu = np.array([AM.estoi(audio[i].numpy(), wav_gpu[i]) for i in range(10)])
    e_ref = np.array([AM.estoi(audio[i].numpy(), wav_ref[i]) for i in range(10)])
    assert np.abs(e_gpu - e_ref).max() <= 1e-3, (e_gpu, e_ref)
    assert abs(e_gpu.mean() - e_ref.mean()) <= 1e-3
